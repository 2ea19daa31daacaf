#!/bin/bash
# data-flow factor kernel, second cut (triangular products, one release per step, C-tile prefetch) + moment kernel A/B
mkdir -p gpurun_out
echo "== solver_check"; timeout 300 python scripts/solver_check.py 2>&1 | tail -9
timeout 900 python -m pytest tests/test_gpu_fit.py tests/test_gpu_scale.py -x -q -m gpu > gpurun_out/r4k_tests.log 2>&1; echo "fit+scale tests rc=$?"; tail -3 gpurun_out/r4k_tests.log
SPLPAK_B200_PANELCLK=1 timeout 300 python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu > gpurun_out/r4k_clk.json 2> gpurun_out/r4k_clk.err; echo "clk rc=$?"; grep -E "data-flow|communication|panel worker|helper 0|B1 arrivals" gpurun_out/r4k_clk.err | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r4k_bench.json 2> gpurun_out/r4k_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r4k_bench.json') if l.startswith('{')][-1]); s=d['stages_ms']; print('fit %.2f eval %.2f' % (d['fit_ms'], d['eval_ms']), {k: round(v,2) for k,v in s.items()}, 'chk', d['checksum'])"
