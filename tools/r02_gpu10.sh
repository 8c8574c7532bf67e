#!/bin/bash
# round 2, call 10: maxima split (values in the RHS launch, stamps in the final stage), three-kernel morpho stage as default
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 ) > gpurun_out/r02_tests10.log 2>&1
cat gpurun_out/r02_tests10.log
timeout 600 python bench.py --no-cpu > gpurun_out/r02_bench_16384_v8.json 2> gpurun_out/r02_bench_16384_v8.err
tail -1 gpurun_out/r02_bench_16384_v8.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('hydro v8', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('kernel_ms'), d['e2e']['value'])"
timeout 900 python bench.py --workload morpho --size 8192 --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/r02_bench_morpho_8192_v8.json 2> gpurun_out/r02_bench_morpho_8192_v8.err
tail -1 gpurun_out/r02_bench_morpho_8192_v8.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('morpho v8', d['value'], d['ms_per_step'], d['roofline']['step_frac_of_hbm_roofline'], d['config'].get('rolled_back_attempts'), d.get('other_arithmetic'))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_4096_v8.csv \
   python bench.py --size 4096 --steps 2 --warmup 3 --no-cpu --no-e2e --no-faithful > gpurun_out/r02_ncu_list_v8.log 2>&1
