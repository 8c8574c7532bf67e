#!/bin/bash
# round 2, call 11: periodic index wrap without the run-time modulos (morpho_bed_kernel was issue-bound on them)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_acceptance.py -m gpu -q -x -k "morpho or redistribution or depositional or lake or maxima" 2>&1 | tail -5 ) > gpurun_out/r02_tests11.log 2>&1
cat gpurun_out/r02_tests11.log
timeout 900 python bench.py --workload morpho --size 8192 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_morpho_8192_v9.json 2> gpurun_out/r02_bench_morpho_8192_v9.err
tail -1 gpurun_out/r02_bench_morpho_8192_v9.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('morpho v9', d['value'], d['ms_per_step'], d['roofline']['step_frac_of_hbm_roofline'], d['config'].get('rolled_back_attempts'), d.get('e2e',{}).get('value'), d['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_morpho_4096_v9.csv \
   python bench.py --workload morpho --size 4096 --steps 2 --warmup 3 --no-cpu --no-e2e --no-faithful > gpurun_out/r02_ncu_list_morpho_v9.log 2>&1
