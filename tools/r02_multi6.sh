#!/bin/bash
# round 2, final build: tests/test_gpu_multi.py through pytest on a 2-GPU box (world 4 / 8 cases skip)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs 2>&1 | tail -25 ) > gpurun_out/r02_multigpu_pytest_n2.log 2>&1
cat gpurun_out/r02_multigpu_pytest_n2.log
