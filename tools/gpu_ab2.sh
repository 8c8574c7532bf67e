#!/bin/bash
# A/B run 2: CTA-shape variants (tools/build_variant.py), prefetch distance, maxima prefetch; then the full default bench.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
V=$PWD/kestrel_b200/lib/variants
: > gpurun_out/ab2_bench.log
b() {  # label, env..., then bench args
  local label=$1; shift
  env "$@" timeout 300 python bench.py --size 4096 --steps 60 --warmup 10 --no-cpu --no-e2e --no-faithful 2>&1 | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$label size=4096 value=%.4g ms=%.3f kernel_ms=%.4f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))" \
    >> gpurun_out/ab2_bench.log 2>&1
}
for v in t288 t192 t224; do
  ( KGPU_LIB=$V/$v/libkestrel_gpu.so timeout 600 python -m pytest tests/test_gpu_fast.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/ab2_tests_$v.log 2>&1
done
b default KGPU_TUNE=15
b tune31 KGPU_TUNE=31
b pd222 KGPU_TUNE=31 KGPU_PREFETCH_DISTANCE=222
b pd888 KGPU_TUNE=31 KGPU_PREFETCH_DISTANCE=888
b t288 KGPU_TUNE=31 KGPU_LIB=$V/t288/libkestrel_gpu.so
b t192 KGPU_TUNE=31 KGPU_LIB=$V/t192/libkestrel_gpu.so KGPU_PREFETCH_DISTANCE=592
b t224 KGPU_TUNE=31 KGPU_LIB=$V/t224/libkestrel_gpu.so KGPU_PREFETCH_DISTANCE=592
b default KGPU_TUNE=15
timeout 900 python bench.py > gpurun_out/bench_16384.json 2> gpurun_out/bench_16384.err
cat gpurun_out/ab2_tests_*.log gpurun_out/ab2_bench.log; tail -c 1500 gpurun_out/bench_16384.json
