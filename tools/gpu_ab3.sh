#!/bin/bash
# A/B run 3: face loop split by direction, phase-D loads hoisted above the barrier
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
V=$PWD/kestrel_b200/lib/variants
: > gpurun_out/ab3_bench.log
b() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --size 4096 --steps 60 --warmup 10 --no-cpu --no-e2e $ARITH 2>&1 | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); o=d.get('other_arithmetic') or {}; print('$label size=4096 value=%.4g ms=%.3f kernel_ms=%.4f faithful=%.4g' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], o.get('value', 0)))" \
    >> gpurun_out/ab3_bench.log 2>&1
}
ARITH=""
b default A=1
for v in split hoist hoistsplit; do b $v KGPU_LIB=$V/$v/libkestrel_gpu.so; done
b default A=1
for v in split hoist hoistsplit; do
  ( KGPU_LIB=$V/$v/libkestrel_gpu.so timeout 600 python -m pytest tests/test_gpu_fast.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/ab3_tests_$v.log 2>&1
done
cat gpurun_out/ab3_bench.log gpurun_out/ab3_tests_*.log
