#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_parity.py::test_morpho_redistribution_single_and_global_walk" tests/test_gpu_parity.py::test_morpho_dambreak_periodic -m gpu -q 2>&1 | tail -30 ) > gpurun_out/multi3_tests.log 2>&1
: > gpurun_out/multi3.log
r() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
   tests/run_multigpu.py "$@" 2>&1 | grep -E "MULTIGPU|Error|error" | head -5 | sed "s/^/[$*] /" >> gpurun_out/multi3.log; }
r --tiles 8 --per 64 --steps 20 --arithmetic 1
r --tiles 8 --per 64 --steps 20 --morpho --arithmetic 1
r --tiles 4 --per 32 --steps 20 --thin --arithmetic 1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
   bench.py --gpus 2 --workload morpho --size 4096 --steps 10 --warmup 3 --no-cpu --no-e2e --no-faithful > gpurun_out/morpho_n2.json 2> gpurun_out/morpho_n2.err
cat gpurun_out/multi3_tests.log gpurun_out/multi3.log
tail -1 gpurun_out/morpho_n2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['rolled_back_attempts'])"
grep -i "KestrelError" gpurun_out/morpho_n2.err | head -3
