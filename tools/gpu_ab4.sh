#!/bin/bash
# A/B run 4: phase D's inputs staged by TMA (KGPU_STAGE_D: 0 none, unset = what fits)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 ) > gpurun_out/ab4_tests.log 2>&1
: > gpurun_out/ab4_bench.log
b() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --size 4096 --steps 60 --warmup 10 --no-cpu --no-e2e --no-faithful $EXTRA 2>&1 | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$label $EXTRA size=4096 value=%.4g ms=%.3f kernel_ms=%.4f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))" \
    >> gpurun_out/ab4_bench.log 2>&1
}
EXTRA=""
b stageD=0 KGPU_STAGE_D=0
b stageD=auto A=1
b stageD=1 KGPU_STAGE_D=1
b stageD=2 KGPU_STAGE_D=2
b stageD=auto A=1
EXTRA="--arithmetic 0"
b stageD=0 KGPU_STAGE_D=0
b stageD=auto A=1
EXTRA="--workload morpho --steps 15"
b stageD=0 KGPU_STAGE_D=0
b stageD=auto A=1
cat gpurun_out/ab4_tests.log gpurun_out/ab4_bench.log
