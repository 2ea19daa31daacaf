#!/bin/bash
# two-level partition, second version: moment tests, cfg3 row A/B, launch list of the binning kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit.py -x -q -m gpu -k "moment or normal_equations or streaming or deterministic" > gpurun_out/r5v_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r5v_tests.log
timeout 600 python scripts/config_times.py cfg3 2>&1 | tail -1 | tee gpurun_out/r5v_cfg.md
SPLPAK_B200_BINNING=atomic timeout 600 python scripts/config_times.py cfg3 2>&1 | tail -1 | sed "s/^/atomic /" | tee -a gpurun_out/r5v_cfg.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spl_(part|perm|scan|items|classify)" -c 12 --csv --log-file gpurun_out/r5v_launches.csv python scripts/gpu_time.py 1e8 1e6 > gpurun_out/r5v_ncu.log 2>&1
python scripts/launch_shares.py gpurun_out/r5v_launches.csv | head -12
