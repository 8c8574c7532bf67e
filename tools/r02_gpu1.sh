#!/bin/bash
# Round 2, GPU call 1: full GPU test-suite (with the round-2 parity cases), the default bench (first measurement of f90b2af),
# the morphodynamic bench with the unbounded redistribution list, the launch list and one ncu --set full capture of each stage mode.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nproc > gpurun_out/r02_host.txt; free -g >> gpurun_out/r02_host.txt; nvidia-smi -L >> gpurun_out/r02_host.txt
( timeout 1500 python -m pytest tests -m gpu -q --durations=15 2>&1 | tail -60 ) > gpurun_out/r02_tests1.log 2>&1
timeout 900 python bench.py > gpurun_out/r02_bench_16384.json 2> gpurun_out/r02_bench_16384.err
timeout 900 python bench.py --workload morpho --size 8192 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_morpho_8192.json 2> gpurun_out/r02_bench_morpho_8192.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_4096_contracted.csv \
   python bench.py --size 4096 --steps 2 --warmup 3 --no-cpu --no-e2e --no-faithful > gpurun_out/r02_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hydro_stage_kernel -s 12 -c 4 -f -o gpurun_out/r02_stage_modes_contracted \
   python bench.py --size 4096 --steps 3 --warmup 3 --no-cpu --no-e2e --no-faithful > gpurun_out/r02_ncu_full_c.log 2>&1
cat gpurun_out/r02_tests1.log
for f in r02_bench_16384 r02_bench_morpho_8192; do tail -1 gpurun_out/$f.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$f', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], (d.get('e2e') or {}).get('value'), d['config'].get('rolled_back_attempts'), d['clocks'], d.get('other_arithmetic'))"; done
tail -3 gpurun_out/*.err
ls -la gpurun_out/*.ncu-rep
