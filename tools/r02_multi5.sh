#!/bin/bash
# round 2, final build: the 2-GPU weak-scaling bench line exactly as the driver launches it (decomposed-vs-single bitwise check included)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 \
   bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_weak_final_n2.json 2> gpurun_out/r02_weak_final_n2.err
tail -1 gpurun_out/r02_weak_final_n2.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('weak', d['n_gpus'], d['scaling'], d['config']['grid'], 'value %.4g' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e %.4g' % ((d.get('e2e') or {}).get('value') or 0), 'bitwise', d.get('decomposed_bitwise'), d.get('decomposed_bitwise_runs'), d['clocks'])"
grep -iE "error|Traceback" gpurun_out/r02_weak_final_n2.err | head -3
