#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
: > gpurun_out/multi2.log
r() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
   tests/run_multigpu.py "$@" 2>&1 | grep MULTIGPU | sed "s/^/[$*] /" >> gpurun_out/multi2.log; }
r --tiles 8 --per 64 --steps 20 --morpho --arithmetic 0
r --tiles 8 --per 64 --steps 20 --arithmetic 1
r --tiles 8 --per 32 --steps 12 --morpho --arithmetic 1
r --tiles 8 --per 64 --steps 1 --morpho --arithmetic 1
r --tiles 8 --per 64 --steps 3 --morpho --arithmetic 1
r --tiles 8 --per 64 --steps 8 --morpho --arithmetic 1
KGPU_TUNE=0 r --tiles 8 --per 64 --steps 8 --morpho --arithmetic 1
cat gpurun_out/multi2.log
