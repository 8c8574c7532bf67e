#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/last_tests.log 2>&1
timeout 200 python bench.py --size 8192 --steps 40 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 > gpurun_out/last_bench_8192.json
cat gpurun_out/last_tests.log; python -c "
import json; d=json.loads(open('gpurun_out/last_bench_8192.json').read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['other_arithmetic']['value'])"
