#!/bin/bash
# 2 GPUs: C-ABI multi-GPU test (default + deterministic replicas); four-config table of the final build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r5o_tests_multi.log 2>&1; echo "multi tests rc=$?"; tail -3 gpurun_out/r5o_tests_multi.log
timeout 600 python scripts/config_times.py > gpurun_out/r5o_configs.md 2>&1; cat gpurun_out/r5o_configs.md
