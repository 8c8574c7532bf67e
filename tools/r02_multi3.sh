#!/bin/bash
# round 2: domain edges in dynamic decomposed runs (dirichlet ghost tiles, halt)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
W=${W:-2}
run() {
   local name=$1; shift
   timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 --master-port 29571 \
      tests/run_multigpu_dynamic.py "$@" > gpurun_out/r02_dyn_${name}_n$W.log 2>&1
   echo "== $name rc=$?"; grep -E "MULTIGPU|^    (output|rank|ghost)|Error|KestrelError" gpurun_out/r02_dyn_${name}_n$W.log | head -20
}
S2="--set nXpertile=10 --set nYpertile=10 --set Xtilesize=10.0 --set Ytilesize=None --set Nout=2"
S8="$S2 --set nXtiles=8 --set nYtiles=8 --set tend=20.0"
run dirichlet --case case_flux_hydro_2d.txt $S8 --set bcs=dirichlet --set bcsHnval=0.02 --set bcsuval=0.1
run dirichlet_fast --case case_flux_hydro_2d.txt $S8 --set bcs=dirichlet --set bcsHnval=0.02 --set bcsuval=0.1 --arithmetic 1
run halt --case case_flux_hydro_2d.txt $S8 --expect-halt
