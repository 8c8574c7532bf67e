#!/bin/bash
# round 2, final single-GPU verification of the committed build: full GPU test suite, the two bench lines, launch lists,
# one ncu --set full capture of the four stage launches of a step for each arithmetic variant
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 ) > gpurun_out/r02_tests_final.log 2>&1
cat gpurun_out/r02_tests_final.log
timeout 600 python bench.py > gpurun_out/r02_bench_16384_final.json 2> gpurun_out/r02_bench_16384_final.err
tail -1 gpurun_out/r02_bench_16384_final.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('hydro final', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('kernel_ms'), d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])"
timeout 900 python bench.py --workload morpho --size 8192 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_morpho_8192_final.json 2> gpurun_out/r02_bench_morpho_8192_final.err
tail -1 gpurun_out/r02_bench_morpho_8192_final.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('morpho final', d['value'], d['ms_per_step'], d['roofline']['step_frac_of_hbm_roofline'], d['config'].get('rolled_back_attempts'), d.get('e2e',{}).get('value'))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_4096_final.csv \
   python bench.py --size 4096 --steps 2 --warmup 3 --no-cpu --no-e2e --no-faithful > gpurun_out/r02_ncu_list_final.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_morpho_4096_final.csv \
   python bench.py --workload morpho --size 4096 --steps 2 --warmup 3 --no-cpu --no-e2e --no-faithful > gpurun_out/r02_ncu_list_morpho_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hydro_stage_kernel -s 12 -c 4 -f -o gpurun_out/r02_stage_modes_faithful \
   python bench.py --size 4096 --steps 3 --warmup 3 --no-cpu --no-e2e --no-faithful --arithmetic 0 > gpurun_out/r02_ncu_full_faithful.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
