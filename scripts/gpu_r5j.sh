#!/bin/bash
# deterministic mode: new test + the fit suite (the flush sites changed for everybody)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit.py -x -q -m gpu -k "deterministic" > gpurun_out/r5j_det.log 2>&1; echo "det tests rc=$?"; tail -30 gpurun_out/r5j_det.log
timeout 1200 python -m pytest tests/test_gpu_fit.py tests/test_gpu_scale.py -x -q -m gpu -k "not deterministic" > gpurun_out/r5j_tests.log 2>&1; echo "fit+scale tests rc=$?"; tail -5 gpurun_out/r5j_tests.log
