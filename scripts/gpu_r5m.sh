#!/bin/bash
# final build of round 2: whole GPU suite, smoke, contract bench at N = 1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r5m_tests.log 2>&1; echo "gpu suite rc=$?"; tail -4 gpurun_out/r5m_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r5m_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r5m_smoke.log
timeout 900 python bench.py > gpurun_out/r5m_bench.json 2> gpurun_out/r5m_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r5m_bench.err; head -c 1500 gpurun_out/r5m_bench.json
