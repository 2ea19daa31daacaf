#!/bin/bash
# last build of round 2: whole GPU suite, smoke, four-config table, contract bench at N = 1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r5t_tests.log 2>&1; echo "gpu suite rc=$?"; tail -4 gpurun_out/r5t_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r5t_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r5t_smoke.log
timeout 600 python scripts/config_times.py > gpurun_out/r5t_configs.md 2>&1; cat gpurun_out/r5t_configs.md
timeout 900 python bench.py > gpurun_out/r5t_bench.json 2> gpurun_out/r5t_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r5t_bench.err; head -c 600 gpurun_out/r5t_bench.json
