#!/bin/bash
# two-level partition binning: fit / scale tests, cfg3 + cfg4 rows with the partition and with the per-point atomics
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fit.py tests/test_gpu_scale.py -x -q -m gpu > gpurun_out/r5u_tests.log 2>&1; echo "fit+scale tests rc=$?"; tail -4 gpurun_out/r5u_tests.log
timeout 600 python scripts/config_times.py cfg3 cfg4 2>&1 | tail -2 | tee gpurun_out/r5u_cfg.md
SPLPAK_B200_BINNING=atomic timeout 600 python scripts/config_times.py cfg3 cfg4 2>&1 | tail -2 | sed "s/^/atomic /" | tee -a gpurun_out/r5u_cfg.md
