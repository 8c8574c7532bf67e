#!/bin/bash
# A/B run 6: phase D evaluates the drag coefficient before the flux divergence (variant dragfirst)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
V=$PWD/kestrel_b200/lib/variants
: > gpurun_out/ab6_bench.log
b() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --size 4096 --steps 60 --warmup 10 --no-cpu --no-e2e $EXTRA 2>&1 | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); o=d.get('other_arithmetic') or {}; print('$label $EXTRA size=4096 value=%.4g ms=%.3f kernel_ms=%.4f faithful=%.4g' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], o.get('value',0)))" \
    >> gpurun_out/ab6_bench.log 2>&1
}
EXTRA=""
b default A=1
b dragfirst KGPU_LIB=$V/dragfirst/libkestrel_gpu.so
b default A=1
b dragfirst KGPU_LIB=$V/dragfirst/libkestrel_gpu.so
( KGPU_LIB=$V/dragfirst/libkestrel_gpu.so timeout 600 python -m pytest tests/test_gpu_fast.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -4 ) > gpurun_out/ab6_tests.log 2>&1
cat gpurun_out/ab6_bench.log gpurun_out/ab6_tests.log
