#!/bin/bash
# round 2: dynamic tiles across ranks (2 GPUs)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
W=${W:-2}
run() {  # name, args...
   local name=$1; shift
   timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 --master-port 29571 \
      tests/run_multigpu_dynamic.py "$@" > gpurun_out/r02_dyn_${name}_n$W.log 2>&1
   echo "== $name rc=$?"; grep -E "MULTIGPU|^    (output|rank|ghost)|Error|KestrelError" gpurun_out/r02_dyn_${name}_n$W.log | head -20
}
S2="--set nXpertile=10 --set nYpertile=10 --set Xtilesize=10.0 --set Ytilesize=None --set Nout=2"
if [ "${HYDRO:-1}" = 1 ]; then
run flux2d --case case_flux_hydro_2d.txt $S2
run cap2d --case case_cap_conc_2d.txt $S2 --set tend=8.0
run flux1d --case case_flux_hydro.txt --set nXpertile=10 --set Xtilesize=10.0 --set tend=30.0 --set Nout=2
run flux2d_fast --case case_flux_hydro_2d.txt $S2 --arithmetic 1
fi
# morphodynamics: redistribution with dynamic tiles across the seam, a rank without active tiles, 1-D
run capm2d --case case_cap_morpho_2d.txt $S2 --set tend=2.0
run indep20 --case case_tile_indep_dynamic_20m.txt --set tend=2.0 --set Nout=2
run fluxm2d --case case_flux_morpho_2d.txt $S2 --set tend=5.0
run capm1d --case case_cap_morpho.txt --set nXpertile=20 --set Xtilesize=20.0 --set tend=5.0 --set Nout=2
run fluxm1d --case case_flux_morpho.txt --set nXpertile=10 --set Xtilesize=10.0 --set tend=10.0 --set Nout=2
run capm2d_fast --case case_cap_morpho_2d.txt $S2 --set tend=2.0 --arithmetic 1
S8="$S2 --set nXtiles=8 --set nYtiles=8 --set tend=20.0"
run dirichlet --case case_flux_hydro_2d.txt $S8 --set bcs=dirichlet --set bcsHnval=0.02 --set bcsuval=0.1
run halt --case case_flux_hydro_2d.txt $S8 --expect-halt
# the periodic all-active path after the non-finite allreduce
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 --master-port 29572 tests/run_multigpu.py --tiles 8 --per 32 --steps 25 2>&1 | grep MULTIGPU
