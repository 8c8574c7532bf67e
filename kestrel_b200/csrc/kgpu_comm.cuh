// kgpu_comm.cuh -- halo strips for the 2-D block decomposition: pack / unpack kernels.
// The exchange itself is ncclSend / ncclRecv between the (up to four) neighbouring ranks,
// issued on a communication stream so that it overlaps the interior of the stage kernel.
#pragma once
#include "kgpu_device.cuh"

namespace kgpu {

struct StripArgs {
   double *f[8];
   int nf;
};

// two columns [c0, c0+2) over rows [r0, r0+nr): buffer layout [field][row][2]
__global__ void strip_cols_kernel(const DevParams P, const StripArgs A, double *buf, int c0, int r0, int nr, int unpack) {
   int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k >= nr * 2) return;
   int r = k >> 1, c = k & 1;
   size_t g = (size_t)(r0 + r + YO) * P.pitch + (c0 + c + XO);
   for (int f = 0; f < A.nf; f++) {
      size_t b = ((size_t)f * nr + r) * 2 + c;
      if (unpack) A.f[f][g] = buf[b];
      else buf[b] = A.f[f][g];
   }
}
// two rows [r0, r0+2) over columns [c0, c0+nc): buffer layout [field][2][nc]
__global__ void strip_rows_kernel(const DevParams P, const StripArgs A, double *buf, int r0, int c0, int nc, int unpack) {
   int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k >= nc * 2) return;
   int r = k / nc, c = k % nc;
   size_t g = (size_t)(r0 + r + YO) * P.pitch + (c0 + c + XO);
   for (int f = 0; f < A.nf; f++) {
      size_t b = ((size_t)f * 2 + r) * nc + c;
      if (unpack) A.f[f][g] = buf[b];
      else buf[b] = A.f[f][g];
   }
}

}  // namespace kgpu
