#!/bin/bash
# A/B run 5: drag coefficient computed in phase A (variant draga) + the device-topography tests
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
V=$PWD/kestrel_b200/lib/variants
( KGPU_LIB=$V/draga/libkestrel_gpu.so timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/ab5_tests_draga.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_topography.py tests/test_gpu_fast.py tests/test_gpu_output.py -m gpu -q 2>&1 | tail -8 ) > gpurun_out/ab5_tests_default.log 2>&1
: > gpurun_out/ab5_bench.log
b() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --size 4096 --steps 60 --warmup 10 --no-cpu --no-e2e --no-faithful $EXTRA 2>&1 | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$label $EXTRA size=4096 value=%.4g ms=%.3f kernel_ms=%.4f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))" \
    >> gpurun_out/ab5_bench.log 2>&1
}
EXTRA=""
b default A=1
b draga KGPU_LIB=$V/draga/libkestrel_gpu.so
b default A=1
b draga KGPU_LIB=$V/draga/libkestrel_gpu.so
EXTRA="--workload morpho --steps 15"
b default A=1
b draga KGPU_LIB=$V/draga/libkestrel_gpu.so
cat gpurun_out/ab5_tests_draga.log gpurun_out/ab5_tests_default.log gpurun_out/ab5_bench.log
