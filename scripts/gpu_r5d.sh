#!/bin/bash
# K-blocked update, second version (test-free operand loads for interior tiles, one barrier per half-chunk) + stream priority A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit.py -x -q -m gpu -k "solver_paths_agree or solver_failure" > gpurun_out/r5d_tests_solver.log 2>&1; echo "solver tests rc=$?"; tail -3 gpurun_out/r5d_tests_solver.log
timeout 600 python -m pytest tests/test_gpu_scale.py -x -q -m gpu -k "cfg4" > gpurun_out/r5d_tests_cfg4.log 2>&1; echo "cfg4 tests rc=$?"; tail -3 gpurun_out/r5d_tests_cfg4.log
for prio in 1 0; do for kb in 4 8; do
  SPLPAK_B200_STPRIO=$prio SPLPAK_B200_KBLOCK=$kb timeout 300 python scripts/cfg4_fit_once.py 1e6 3 2>&1 | tail -1 | sed "s/^/PRIO=$prio KB=$kb /" | tee -a gpurun_out/r5d_cfg4.log
done; done
