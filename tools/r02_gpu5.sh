#!/bin/bash
# Round 2, GPU call 5: full GPU suite (incl. the acceptance suite at full length) on the v4 stage kernel, headline bench, morphodynamic bench + launch list.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 2400 python -m pytest tests -m gpu -q --durations=25 2>&1 | tail -70 ) > gpurun_out/r02_tests5.log 2>&1
cat gpurun_out/r02_tests5.log | tail -45
timeout 900 python bench.py > gpurun_out/r02_bench_16384_v4.json 2> gpurun_out/r02_bench_16384_v4.err
timeout 900 python bench.py --workload morpho --size 8192 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_morpho_8192_v4.json 2> gpurun_out/r02_bench_morpho_8192_v4.err
for f in r02_bench_16384_v4 r02_bench_morpho_8192_v4; do tail -1 gpurun_out/$f.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$f', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], (d.get('e2e') or {}).get('value'), d['config'].get('rolled_back_attempts'), d['clocks'], d.get('other_arithmetic'))"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_morpho_4096.csv \
   python bench.py --workload morpho --size 4096 --steps 2 --warmup 3 --no-cpu --no-e2e --no-faithful > gpurun_out/r02_ncu_list_morpho.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_4096_contracted_v4.csv \
   python bench.py --size 4096 --steps 2 --warmup 3 --no-cpu --no-e2e --no-faithful > gpurun_out/r02_ncu_list_v4.log 2>&1
