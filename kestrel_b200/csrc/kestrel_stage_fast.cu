// kestrel_stage_fast.cu -- contracted-arithmetic instantiations of the fused stage kernel
// (params.arithmetic = 1).  This translation unit alone is compiled with -fmad=true: products
// and sums fuse into DFMA, divisions sharing a denominator become one reciprocal, the CFL
// minimum is tracked as a maximum rate.  Results differ from the faithful variant in the last
// bits only; tests hold it to the north-star tolerance (rel-Linf 1e-10 per field).
#define KGPU_STAGE_ONLY 1
#include "kgpu_hydro.cuh"

namespace kgpu {

constexpr int FBX2 = 32, FBY2 = KGPU_STAGE_BY2, FBX1 = 128, FBY1 = 1;

template <bool ONED, bool HASBT, int LIM, int SPEC>
static void launchFastK(dim3 nblocks, cudaStream_t s, const DevParams &P, const StageArgs &a) {
   constexpr int BX = ONED ? FBX1 : FBX2, BY = ONED ? FBY1 : FBY2;
   using G = StageGeom<BX, BY, ONED>;
   hydro_stage_kernel<BX, BY, ONED, HASBT, LIM, true, SPEC>
      <<<nblocks, KGPU_STAGE_THREADS, G::smemBytes(HASBT, true, stageFluxPlanes(SPEC)), s>>>(P, a);
}

// spec: geometric factors on and no eddy viscosity -- the 2-D instantiation that has both as compile-time constants
void launch_stage_fast(bool oneD, bool hasBt, bool mm2, bool spec, dim3 nblocks, cudaStream_t s, const DevParams &P, const StageArgs &a) {
   if (oneD) {
      if (hasBt) { if (mm2) launchFastK<true, true, KGPU_LIM_MINMOD2, 0>(nblocks, s, P, a); else launchFastK<true, true, -1, 0>(nblocks, s, P, a); }
      else       { if (mm2) launchFastK<true, false, KGPU_LIM_MINMOD2, 0>(nblocks, s, P, a); else launchFastK<true, false, -1, 0>(nblocks, s, P, a); }
   } else if (spec) {
      if (hasBt) { if (mm2) launchFastK<false, true, KGPU_LIM_MINMOD2, 1>(nblocks, s, P, a); else launchFastK<false, true, -1, 1>(nblocks, s, P, a); }
      else       { if (mm2) launchFastK<false, false, KGPU_LIM_MINMOD2, 1>(nblocks, s, P, a); else launchFastK<false, false, -1, 1>(nblocks, s, P, a); }
   } else {
      if (hasBt) { if (mm2) launchFastK<false, true, KGPU_LIM_MINMOD2, 0>(nblocks, s, P, a); else launchFastK<false, true, -1, 0>(nblocks, s, P, a); }
      else       { if (mm2) launchFastK<false, false, KGPU_LIM_MINMOD2, 0>(nblocks, s, P, a); else launchFastK<false, false, -1, 0>(nblocks, s, P, a); }
   }
}

template <bool ONED, bool HASBT, int LIM, int SPEC>
static void setAttr() {
   constexpr int BX = ONED ? FBX1 : FBX2, BY = ONED ? FBY1 : FBY2;
   using G = StageGeom<BX, BY, ONED>;
   cudaFuncSetAttribute(hydro_stage_kernel<BX, BY, ONED, HASBT, LIM, true, SPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                        (int)G::smemBytes(true, true, stageFluxPlanes(SPEC)));
}
void stage_fast_set_attributes() {
   setAttr<false, false, KGPU_LIM_MINMOD2, 0>(); setAttr<false, false, -1, 0>(); setAttr<false, true, KGPU_LIM_MINMOD2, 0>(); setAttr<false, true, -1, 0>();
   setAttr<false, false, KGPU_LIM_MINMOD2, 1>(); setAttr<false, false, -1, 1>(); setAttr<false, true, KGPU_LIM_MINMOD2, 1>(); setAttr<false, true, -1, 1>();
   setAttr<true, false, KGPU_LIM_MINMOD2, 0>(); setAttr<true, false, -1, 0>(); setAttr<true, true, KGPU_LIM_MINMOD2, 0>(); setAttr<true, true, -1, 0>();
}

}  // namespace kgpu
