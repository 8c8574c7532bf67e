#!/bin/bash
# A/B run of the stage-kernel tuning bits (KGPU_TUNE) on one B200: parity suite with every bit on,
# then the bench at 4096^2 and 8192^2 per setting.  Output under gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( KGPU_TUNE=15 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/ab_tests.log 2>&1
for t in 0 1 2 4 8 15 0 15; do
  KGPU_TUNE=$t timeout 300 python bench.py --size 4096 --steps 60 --warmup 10 --no-cpu --no-e2e --no-faithful 2>&1 | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tune=$t size=4096 value=%.4g ms=%.3f kernel_ms=%.4f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))" \
    >> gpurun_out/ab_bench.log 2>&1
done
for t in 0 15; do
  KGPU_TUNE=$t timeout 300 python bench.py --size 8192 --steps 40 --warmup 5 --no-cpu --no-e2e --arithmetic 0 --no-faithful 2>&1 | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tune=$t faithful size=8192 value=%.4g ms=%.3f kernel_ms=%.4f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))" \
    >> gpurun_out/ab_bench.log 2>&1
done
for t in 0 15; do
  KGPU_LIB=$PWD/kestrel_b200/lib/variants/mb2/libkestrel_gpu.so KGPU_TUNE=$t timeout 300 python bench.py --size 4096 --steps 60 --warmup 10 --no-cpu --no-e2e --no-faithful 2>&1 | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mb2 tune=$t size=4096 value=%.4g ms=%.3f kernel_ms=%.4f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))" \
    >> gpurun_out/ab_bench.log 2>&1
done
cat gpurun_out/ab_tests.log gpurun_out/ab_bench.log
