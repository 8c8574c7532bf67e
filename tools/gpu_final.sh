#!/bin/bash
# Round-end 1-GPU run: full GPU test-suite, the default bench, the morphodynamic bench, the output-interval leg and the ncu captures.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/final_tests.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_16384.json 2> gpurun_out/bench_16384.err
timeout 600 python bench.py --workload morpho --size 8192 --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_morpho_8192.json 2> gpurun_out/bench_morpho_8192.err
timeout 600 python bench.py --size 8192 --steps 20 --warmup 5 --output-intervals 4 --no-cpu --no-e2e --no-faithful > gpurun_out/bench_output_8192.json 2> gpurun_out/bench_output_8192.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_4096_contracted.csv \
   python bench.py --size 4096 --steps 2 --warmup 3 --no-cpu --no-e2e --no-faithful > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hydro_stage_kernel -s 9 -c 1 -f -o gpurun_out/stage_contracted_v7 \
   python bench.py --size 2048 --steps 3 --warmup 3 --no-cpu --no-e2e --no-faithful > gpurun_out/ncu_full_c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hydro_stage_kernel -s 9 -c 1 -f -o gpurun_out/stage_faithful_v9 \
   python bench.py --size 2048 --steps 3 --warmup 3 --no-cpu --no-e2e --no-faithful --arithmetic 0 > gpurun_out/ncu_full_f.log 2>&1
cat gpurun_out/final_tests.log
for f in bench_16384 bench_morpho_8192 bench_output_8192; do tail -1 gpurun_out/$f.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$f', d['value'], d['ms_per_step'], d['roofline']['frac'], (d.get('e2e') or {}).get('value'), d.get('output_intervals'), d['clocks'])"; done
ls -la gpurun_out/*.ncu-rep
