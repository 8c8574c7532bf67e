#!/bin/bash
# round 2: dynamic tiles across 4 ranks (2 x 2; 1-D: 4 x 1) + the periodic redistribution check after the patch layout change
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
W=4
run() {
   local name=$1; shift
   timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 --master-port 29571 \
      tests/run_multigpu_dynamic.py "$@" > gpurun_out/r02_dyn_${name}_n$W.log 2>&1
   echo "== $name rc=$?"; grep -E "MULTIGPU|^    (output|rank|ghost)|Error|KestrelError" gpurun_out/r02_dyn_${name}_n$W.log | head -20
}
S2="--set nXpertile=10 --set nYpertile=10 --set Xtilesize=10.0 --set Ytilesize=None --set Nout=2"
run flux2d --case case_flux_hydro_2d.txt $S2
run capm2d --case case_cap_morpho_2d.txt $S2 --set tend=2.0
run fluxm2d --case case_flux_morpho_2d.txt $S2 --set tend=5.0
run fluxm1d --case case_flux_morpho.txt --set nXpertile=10 --set Xtilesize=10.0 --set tend=10.0 --set Nout=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 --master-port 29573 tests/run_multigpu.py --tiles 4 --per 32 --steps 20 --thin 2>&1 | grep MULTIGPU | tee gpurun_out/r02_periodic_thin_n4.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 --master-port 29574 tests/run_multigpu.py --tiles 8 --per 32 --steps 12 --morpho 2>&1 | grep MULTIGPU | tee gpurun_out/r02_periodic_morpho_n4.log
