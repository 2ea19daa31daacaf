#!/bin/bash
# 8 GPUs: new edge-shape solver tests, the 2-GPU C-ABI test, the contract bench at N = 8 (weak line + strong sub-record + e2e)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fit.py -x -q -m gpu -k "solver_paths_agree or solver_failure" > gpurun_out/r4o_tests_solver.log 2>&1; echo "solver tests rc=$?"; tail -2 gpurun_out/r4o_tests_solver.log
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r4o_tests_multi.log 2>&1; echo "multi tests rc=$?"; tail -2 gpurun_out/r4o_tests_multi.log
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r4o_bench_n8.json 2> gpurun_out/r4o_bench_n8.err; echo "bench n8 rc=$?"; tail -3 gpurun_out/r4o_bench_n8.err
