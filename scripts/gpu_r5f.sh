#!/bin/bash
# 4-D assembly by cell moments: fit / scale tests, cfg4 row (moments vs direct)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fit.py tests/test_gpu_scale.py -x -q -m gpu > gpurun_out/r5f_tests.log 2>&1; echo "fit+scale tests rc=$?"; tail -15 gpurun_out/r5f_tests.log
timeout 600 python scripts/config_times.py cfg4 2>&1 | tail -1 | tee gpurun_out/r5f_cfg4.md
SPLPAK_B200_ASSEMBLY=direct timeout 600 python scripts/config_times.py cfg4 2>&1 | tail -1 | sed "s/^/direct /" | tee -a gpurun_out/r5f_cfg4.md
