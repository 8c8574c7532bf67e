#!/bin/bash
# Round 2, GPU call 6: new section-8(f) pieces (device initial conditions, DEM resample, topographies, source tables) + fixed tests
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1500 python -m pytest tests/test_gpu_initial_conditions.py tests/test_gpu_dem.py tests/test_gpu_topography.py tests/test_gpu_parity2.py tests/test_gpu_acceptance.py -m gpu -q -k "not c5_" -s 2>&1 | grep -v "^$" | tail -60 ) > gpurun_out/r02_tests6.log 2>&1
cat gpurun_out/r02_tests6.log
