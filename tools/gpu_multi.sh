#!/bin/bash
# 2-GPU run: multi-GPU parity tests (hydraulic + morphodynamic), the new single-GPU tests, a morphodynamic weak-scaling point.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_output.py -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/multi_tests.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
   tests/run_multigpu.py --tiles 8 --per 64 --steps 20 --morpho --arithmetic 1 > gpurun_out/multi_morpho_fast.log 2>&1
timeout 600 python bench.py --workload morpho --size 4096 --steps 10 --warmup 3 --no-cpu --no-e2e --no-faithful > gpurun_out/morpho_n1.json 2> gpurun_out/morpho_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
   bench.py --gpus 2 --workload morpho --size 4096 --steps 10 --warmup 3 --no-cpu --no-e2e --no-faithful > gpurun_out/morpho_n2.json 2> gpurun_out/morpho_n2.err
cat gpurun_out/multi_tests.log; grep MULTIGPU gpurun_out/multi_morpho_fast.log; tail -3 gpurun_out/multi_morpho_fast.log
for f in gpurun_out/morpho_n1.json gpurun_out/morpho_n2.json; do tail -1 $f | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['rolled_back_attempts'])"; done
tail -3 gpurun_out/morpho_n2.err
