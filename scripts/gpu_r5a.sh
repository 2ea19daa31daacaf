#!/bin/bash
# K-blocked trailing update of the kernel-per-phase factor loop: solver tests, cfg4 scale test, cfg4 timings per block size
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit.py -x -q -m gpu -k "solver_paths_agree or solver_failure" > gpurun_out/r5a_tests_solver.log 2>&1; echo "solver tests rc=$?"; tail -3 gpurun_out/r5a_tests_solver.log
timeout 600 python -m pytest tests/test_gpu_scale.py -x -q -m gpu -k "cfg4" > gpurun_out/r5a_tests_cfg4.log 2>&1; echo "cfg4 tests rc=$?"; tail -3 gpurun_out/r5a_tests_cfg4.log
for kb in 1 2 4 8; do
  SPLPAK_B200_KBLOCK=$kb timeout 300 python scripts/config_times.py cfg4 2>&1 | tail -1 | sed "s/^/KB=$kb /" | tee -a gpurun_out/r5a_cfg4.log
done
