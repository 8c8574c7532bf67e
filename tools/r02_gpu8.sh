#!/bin/bash
# Round 2, GPU call 8: fused morphodynamic stage (parity against the three-kernel path), second maxima pass dropped, then the benches
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1500 python -m pytest tests/test_gpu_parity2.py tests/test_gpu_parity.py tests/test_gpu_fast.py tests/test_gpu_initial_conditions.py -m gpu -q -x -k "not c5_parity_subset_1024" 2>&1 | tail -25 ) > gpurun_out/r02_tests8.log 2>&1
cat gpurun_out/r02_tests8.log
timeout 900 python bench.py --workload morpho --size 8192 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_morpho_8192_v5.json 2> gpurun_out/r02_bench_morpho_8192_v5.err
timeout 600 python bench.py --size 8192 --steps 40 --warmup 10 --no-cpu --no-e2e > gpurun_out/r02_bench_8192_v5.json 2> gpurun_out/r02_bench_8192_v5.err
for f in r02_bench_morpho_8192_v5 r02_bench_8192_v5; do tail -1 gpurun_out/$f.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$f', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['step_frac_of_hbm_roofline'], d['config'].get('rolled_back_attempts'), d.get('other_arithmetic'))"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_morpho_4096_v5.csv \
   python bench.py --workload morpho --size 4096 --steps 2 --warmup 3 --no-cpu --no-e2e --no-faithful > gpurun_out/r02_ncu_list_morpho_v5.log 2>&1
