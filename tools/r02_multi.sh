#!/bin/bash
# Round 2, multi-GPU call: N = $1 GPUs.  Decomposed parity tests, weak and strong scaling lines (with the decomposed-vs-single bitwise check).
set -u
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -12 ) > gpurun_out/r02_multi_tests_n$N.log 2>&1
cat gpurun_out/r02_multi_tests_n$N.log
run() { local tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
     bench.py --gpus $N "$@" > gpurun_out/r02_${tag}_n$N.json 2> gpurun_out/r02_${tag}_n$N.err
  tail -1 gpurun_out/r02_${tag}_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$tag', d['n_gpus'], d['scaling'], d['config']['grid'], 'value %.4g' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e %.4g' % ((d.get('e2e') or {}).get('value') or 0), 'bitwise', d.get('decomposed_bitwise'), d.get('decomposed_bitwise_runs'))" 2>&1 | tail -2
  grep -iE "error|Traceback" gpurun_out/r02_${tag}_n$N.err | head -3
}
run weak --steps 20 --warmup 5 --no-cpu --no-faithful
run strong --scaling strong --size 16384 --steps 20 --warmup 5 --no-cpu --no-faithful --no-bitwise --no-e2e
run morpho_weak --workload morpho --size 4096 --steps 10 --warmup 3 --no-cpu --no-faithful --no-e2e
