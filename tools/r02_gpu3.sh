#!/bin/bash
# Round 2, GPU call 3: persistent software-pipelined stage kernel against the one-CTA-per-tile schedule and the previous build.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
V=$PWD/kestrel_b200/lib/variants
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fast.py -m gpu -q -x 2>&1 | tail -8 ) > gpurun_out/r02_tests3.log 2>&1
cat gpurun_out/r02_tests3.log
: > gpurun_out/r02_ab2.log
b() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --size 4096 --steps 60 --warmup 10 --no-cpu --no-e2e 2>&1 | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); o=d.get('other_arithmetic') or {}; print('$label size=4096 value=%.4g ms=%.3f kernel_ms=%.4f faithful=%.4g' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], o.get('value',0)))" \
    >> gpurun_out/r02_ab2.log 2>&1
}
b v1 KGPU_LIB=$V/v1/libkestrel_gpu.so
b persistent A=1
b one_cta_per_tile KGPU_TUNE=159
b persistent_pd222 KGPU_PREFETCH_DISTANCE=222
b persistent_pd888 KGPU_PREFETCH_DISTANCE=888
b v1 KGPU_LIB=$V/v1/libkestrel_gpu.so
b persistent A=1
cat gpurun_out/r02_ab2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hydro_stage_kernel -s 12 -c 4 -f -o gpurun_out/r02_stage_modes_v3 \
   python bench.py --size 4096 --steps 3 --warmup 3 --no-cpu --no-e2e --no-faithful > gpurun_out/r02_ncu_full_v3.log 2>&1
