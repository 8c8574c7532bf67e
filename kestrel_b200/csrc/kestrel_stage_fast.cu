// kestrel_stage_fast.cu -- contracted-arithmetic instantiations of the fused stage kernel
// (params.arithmetic = 1).  This translation unit alone is compiled with -fmad=true: products
// and sums fuse into DFMA, divisions sharing a denominator become one reciprocal, the CFL
// minimum is tracked as a maximum rate.  Results differ from the faithful variant in the last
// bits only; tests hold it to the north-star tolerance (rel-Linf 1e-10 per field).
#define KGPU_STAGE_ONLY 1
#include "kgpu_hydro.cuh"

namespace kgpu {

constexpr int FBX2 = 32, FBY2 = KGPU_STAGE_BY2, FBX1 = 128, FBY1 = 1;

template <bool ONED, bool HASBT, int LIM>
static void launchFastK(dim3 nblocks, cudaStream_t s, const DevParams &P, const StageArgs &a) {
   constexpr int BX = ONED ? FBX1 : FBX2, BY = ONED ? FBY1 : FBY2;
   using G = StageGeom<BX, BY, ONED>;
   hydro_stage_kernel<BX, BY, ONED, HASBT, LIM, true><<<nblocks, KGPU_STAGE_THREADS, G::smemBytes(HASBT, true), s>>>(P, a);
}

void launch_stage_fast(bool oneD, bool hasBt, bool mm2, dim3 nblocks, cudaStream_t s, const DevParams &P, const StageArgs &a) {
   if (oneD) {
      if (hasBt) { if (mm2) launchFastK<true, true, KGPU_LIM_MINMOD2>(nblocks, s, P, a); else launchFastK<true, true, -1>(nblocks, s, P, a); }
      else       { if (mm2) launchFastK<true, false, KGPU_LIM_MINMOD2>(nblocks, s, P, a); else launchFastK<true, false, -1>(nblocks, s, P, a); }
   } else {
      if (hasBt) { if (mm2) launchFastK<false, true, KGPU_LIM_MINMOD2>(nblocks, s, P, a); else launchFastK<false, true, -1>(nblocks, s, P, a); }
      else       { if (mm2) launchFastK<false, false, KGPU_LIM_MINMOD2>(nblocks, s, P, a); else launchFastK<false, false, -1>(nblocks, s, P, a); }
   }
}

void stage_fast_set_attributes() {
   int s2 = (int)StageGeom<FBX2, FBY2, false>::smemBytes(true, true), s1 = (int)StageGeom<FBX1, FBY1, true>::smemBytes(true, true);
   cudaFuncSetAttribute(hydro_stage_kernel<FBX2, FBY2, false, false, KGPU_LIM_MINMOD2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2);
   cudaFuncSetAttribute(hydro_stage_kernel<FBX2, FBY2, false, false, -1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2);
   cudaFuncSetAttribute(hydro_stage_kernel<FBX2, FBY2, false, true, KGPU_LIM_MINMOD2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2);
   cudaFuncSetAttribute(hydro_stage_kernel<FBX2, FBY2, false, true, -1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2);
   cudaFuncSetAttribute(hydro_stage_kernel<FBX1, FBY1, true, false, KGPU_LIM_MINMOD2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s1);
   cudaFuncSetAttribute(hydro_stage_kernel<FBX1, FBY1, true, false, -1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s1);
   cudaFuncSetAttribute(hydro_stage_kernel<FBX1, FBY1, true, true, KGPU_LIM_MINMOD2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s1);
   cudaFuncSetAttribute(hydro_stage_kernel<FBX1, FBY1, true, true, -1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s1);
}

}  // namespace kgpu
