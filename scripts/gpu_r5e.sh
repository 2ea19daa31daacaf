#!/bin/bash
# new band expansion (CTA per column, stencil offsets) + KB = 8 default: fit / scale tests, four-config table
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fit.py tests/test_gpu_scale.py -x -q -m gpu > gpurun_out/r5e_tests.log 2>&1; echo "fit+scale tests rc=$?"; tail -3 gpurun_out/r5e_tests.log
timeout 600 python scripts/config_times.py > gpurun_out/r5e_configs.md 2>&1; cat gpurun_out/r5e_configs.md
