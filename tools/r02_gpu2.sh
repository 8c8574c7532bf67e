#!/bin/bash
# Round 2, GPU call 2: A/B of the stage-kernel changes (always-late failed check, block CFL reduced before phase D, block origin
# re-derived in phase D, SPEC instantiation, L1 prefetch, hoisted q0 loads, drag before the barrier) against the round-1 kernel.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
V=$PWD/kestrel_b200/lib/variants
: > gpurun_out/r02_ab1.log
b() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --size 4096 --steps 60 --warmup 10 --no-cpu --no-e2e 2>&1 | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); o=d.get('other_arithmetic') or {}; print('$label size=4096 value=%.4g ms=%.3f kernel_ms=%.4f faithful=%.4g' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], o.get('value',0)))" \
    >> gpurun_out/r02_ab1.log 2>&1
}
b base KGPU_LIB=$V/base/libkestrel_gpu.so
b default A=1
b tune63_L1prefetch KGPU_TUNE=63
b tune95_nospec KGPU_TUNE=95
b dhoist KGPU_LIB=$V/dhoist/libkestrel_gpu.so
b predrag KGPU_LIB=$V/predrag/libkestrel_gpu.so
b base KGPU_LIB=$V/base/libkestrel_gpu.so
b default A=1
cat gpurun_out/r02_ab1.log
( timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/r02_tests2.log 2>&1
cat gpurun_out/r02_tests2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hydro_stage_kernel -s 12 -c 4 -f -o gpurun_out/r02_stage_modes_v2 \
   python bench.py --size 4096 --steps 3 --warmup 3 --no-cpu --no-e2e --no-faithful > gpurun_out/r02_ncu_full_v2.log 2>&1
timeout 900 python bench.py --no-cpu > gpurun_out/r02_bench_16384_v2.json 2> gpurun_out/r02_bench_16384_v2.err
tail -1 gpurun_out/r02_bench_16384_v2.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('16384', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], (d.get('e2e') or {}).get('value'), d['clocks'], d.get('other_arithmetic'))"
