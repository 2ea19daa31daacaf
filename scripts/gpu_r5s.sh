#!/bin/bash
# per-node priorities in the captured factor graph: A/B at cfg4 + solver tests
mkdir -p gpurun_out
for prio in 1 0; do for kb in 8 4; do
  SPLPAK_B200_GRAPHPRIO=$prio SPLPAK_B200_KBLOCK=$kb timeout 300 python scripts/cfg4_fit_once.py 1e6 3 2>&1 | tail -1 | sed "s/^/GRAPHPRIO=$prio KB=$kb /" | tee -a gpurun_out/r5s_cfg4.log
done; done
timeout 900 python -m pytest tests/test_gpu_fit.py -x -q -m gpu -k "solver_paths_agree or solver_failure or reuses" > gpurun_out/r5s_tests_solver.log 2>&1; echo "solver tests rc=$?"; tail -3 gpurun_out/r5s_tests_solver.log
