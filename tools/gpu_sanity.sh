#!/bin/bash
# last run of the round: default-build sanity + compute-sanitizer memcheck over this session's new kernels
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 300 python -m pytest tests/test_gpu_topography.py tests/test_gpu_fast.py -m gpu -q 2>&1 | tail -5; timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3 ) > gpurun_out/sanity.log 2>&1
( timeout 270 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest \
    "tests/test_gpu_parity.py::test_morpho_redistribution_single_and_global_walk" \
    "tests/test_gpu_topography.py::test_every_topography_function_matches_the_host" tests/test_gpu_output.py -m gpu -q -x 2>&1 | tail -15; echo "exit=$?" ) > gpurun_out/sanitizer2.log 2>&1
cat gpurun_out/sanity.log gpurun_out/sanitizer2.log
