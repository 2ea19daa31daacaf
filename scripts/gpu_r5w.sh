#!/bin/bash
# two-level partition, final form: cfg3 row A/B, cfg2/cfg1 sanity, deterministic-mode cost
mkdir -p gpurun_out
timeout 600 python scripts/config_times.py cfg3 2>&1 | tail -1 | tee gpurun_out/r5w_cfg.md
SPLPAK_B200_BINNING=atomic timeout 600 python scripts/config_times.py cfg3 2>&1 | tail -1 | sed "s/^/atomic /" | tee -a gpurun_out/r5w_cfg.md
timeout 900 python -m pytest tests/test_gpu_fit.py -x -q -m gpu -k "moment or normal_equations or streaming or deterministic or histogram" > gpurun_out/r5w_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r5w_tests.log
timeout 600 python scripts/det_check.py 2>&1 | grep cfg | tee gpurun_out/r5w_det.log
