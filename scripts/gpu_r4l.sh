#!/bin/bash
# kernel-per-phase driver: 256-thread update tile vs the 128-thread one (cfg4 is the shape that uses it); all four configs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fit.py tests/test_gpu_scale.py -x -q -m gpu > gpurun_out/r4l_tests.log 2>&1; echo "fit+scale tests rc=$?"; tail -3 gpurun_out/r4l_tests.log
SPLPAK_B200_SYRK=128 timeout 600 python scripts/config_times.py > gpurun_out/r4l_configs_syrk128.md 2>&1; tail -4 gpurun_out/r4l_configs_syrk128.md
timeout 600 python scripts/config_times.py > gpurun_out/r4l_configs.md 2>&1; tail -4 gpurun_out/r4l_configs.md
