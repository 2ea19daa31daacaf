#!/bin/bash
# cell transforms with all moment loads in flight: moment tests, cfg3 + cfg4 rows, launch list of the cfg4 assembly
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fit.py -x -q -m gpu -k "moment or refinement or normal_equations or streaming" > gpurun_out/r5i_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r5i_tests.log
timeout 600 python scripts/config_times.py cfg3 cfg4 2>&1 | tail -2 | tee gpurun_out/r5i_cfg.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spl_(moments4|cell_transform4|classify|perm)" -c 8 --csv --log-file gpurun_out/r5i_launches.csv python scripts/cfg4_fit_once.py 1e7 1 > gpurun_out/r5i_ncu_launch.log 2>&1
python scripts/launch_shares.py gpurun_out/r5i_launches.csv | head -12
