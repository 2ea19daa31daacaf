#!/bin/bash
# checkpoint: smoke, full GPU test suite, both bench arms, then the ncu evidence for the bench command
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4m_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r4m_smoke.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r4m_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r4m_tests.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r4m_bench_ref.json 2> gpurun_out/r4m_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r4m_bench.json 2> gpurun_out/r4m_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r4m_bench.err
bash scripts/gpu_profile_r02.sh
