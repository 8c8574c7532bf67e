#!/bin/bash
# round 2: ncu --set full of the morphodynamic kernels of one stage (E - D, bed, cells, check) at 4096^2
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"morpho_(emd|bed|cell|check|prepare)_kernel|topo_planes_kernel|stage1_update_kernel" -s 10 -c 7 -f -o gpurun_out/r02_morpho_kernels \
   python bench.py --workload morpho --size 4096 --steps 2 --warmup 2 --no-cpu --no-e2e --no-faithful > gpurun_out/r02_ncu_morpho_kernels.log 2>&1
ls -la gpurun_out/r02_morpho_kernels.ncu-rep
ncu -i gpurun_out/r02_morpho_kernels.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
keys=['Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','dram__throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum']
for r in rows[2:]:
    print([ (r[h.index(k)][:40] if k in h else None) for k in keys])
"
