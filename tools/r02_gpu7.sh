#!/bin/bash
# Round 2, GPU call 7: fp64 microbenchmark, N = 1 line with the scaling run's flags, device-IC test rerun
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
./kestrel_b200/bin/fp64_peak > gpurun_out/r02_fp64_peak.json 2>&1; cat gpurun_out/r02_fp64_peak.json
./kestrel_b200/bin/fp64_peak >> gpurun_out/r02_fp64_peak.json 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-faithful > gpurun_out/r02_weak_n1.json 2> gpurun_out/r02_weak_n1.err
tail -1 gpurun_out/r02_weak_n1.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('weak n1', 'value %.4g' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], d['e2e']['interval_s'], d['clocks'])"
( timeout 600 python -m pytest tests/test_gpu_initial_conditions.py -m gpu -q 2>&1 | tail -4 ) 
