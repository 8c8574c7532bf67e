// fp64_peak.cu -- measured fp64 pipe rates of the device (hard part 0): DFMA / DADD / DMUL issue rate with
// enough independent chains to fill the pipe, the dependent-issue latency of a DFMA chain, and the cost of the IEEE division and
// square root sequences and of the MUFU-seeded Newton reciprocal the contracted variant uses.  Prints one JSON object.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int CHAINS>
__global__ void __launch_bounds__(256) rate_kernel(double *out, double a, double b, int iters) {
   double x[CHAINS];
#pragma unroll
   for (int c = 0; c < CHAINS; c++) x[c] = a + c * 1e-3 + threadIdx.x * 1e-6;
   for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int c = 0; c < CHAINS; c++) {
         if (OP == 0) x[c] = fma(x[c], a, b);
         else if (OP == 1) x[c] = x[c] + b;
         else if (OP == 2) x[c] = x[c] * a;
         else if (OP == 3) x[c] = b / x[c];                 // IEEE division (div.rn.f64)
         else if (OP == 4) x[c] = sqrt(x[c] + b);           // IEEE square root
         else if (OP == 5) {                                // MUFU.RCP64H + 2 Newton steps (rcpFast of kgpu_hydro.cuh)
            double r;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[c]));
            double e = fma(-x[c], r, 1.0);
            e = fma(e, e, e);
            r = fma(r, e, r);
            e = fma(-x[c], r, 1.0);
            x[c] = fma(r, e, r) + b;
         }
      }
   }
   double s = 0.0;
#pragma unroll
   for (int c = 0; c < CHAINS; c++) s += x[c];
   if (s == 123.456) out[0] = s;   // keep the work
}

template <int OP, int CHAINS>
static double run(double *d, int blocksPerSm, int nsm, int iters, double a, double b) {
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   rate_kernel<OP, CHAINS><<<nsm * blocksPerSm, 256>>>(d, a, b, iters / 10);
   cudaDeviceSynchronize();
   cudaEventRecord(e0);
   rate_kernel<OP, CHAINS><<<nsm * blocksPerSm, 256>>>(d, a, b, iters);
   cudaEventRecord(e1);
   cudaEventSynchronize(e1);
   float ms = 0.f;
   cudaEventElapsedTime(&ms, e0, e1);
   return (double)nsm * blocksPerSm * 256.0 * CHAINS * iters / (ms * 1e-3);   // thread-operations per second
}

int main() {
   cudaDeviceProp p;
   cudaGetDeviceProperties(&p, 0);
   int nsm = p.multiProcessorCount, khz = 0;
   cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
   double *d;
   cudaMalloc(&d, 64);
   const int it = 20000;
   double dfma = run<0, 8>(d, 8, nsm, it, 0.999999, 1e-9);
   double dadd = run<1, 8>(d, 8, nsm, it, 1.0, 1e-9);
   double dmul = run<2, 8>(d, 8, nsm, it, 1.0000001, 0.0);
   double dfma1 = run<0, 1>(d, 1, nsm, it, 0.999999, 1e-9);     // one warp-chain per scheduler at most: dependent-issue latency
   double ddiv = run<3, 4>(d, 8, nsm, it / 10, 1.0, 1.5);
   double dsqrt = run<4, 4>(d, 8, nsm, it / 10, 1.0, 1.5);
   double drcp = run<5, 4>(d, 8, nsm, it / 10, 1.0, 1.5);
   // one block of 256 threads per SM = 2 warps per scheduler, 1 chain: ops per second per thread -> cycles per dependent DFMA
   double lat_cycles = (double)khz * 1e3 / (dfma1 / ((double)nsm * 256.0));
   printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz_attr\": %.0f, \"dfma_per_s\": %.4g, \"dadd_per_s\": %.4g, \"dmul_per_s\": %.4g, "
          "\"fp64_tflops_fma\": %.2f, \"dfma_per_clk_per_sm\": %.1f, \"dependent_dfma_latency_cycles_2_warps_per_scheduler\": %.1f, "
          "\"ieee_div_per_s\": %.4g, \"ieee_sqrt_per_s\": %.4g, \"newton_rcp_per_s\": %.4g, "
          "\"fp64_instr_equiv_div\": %.1f, \"fp64_instr_equiv_sqrt\": %.1f, \"fp64_instr_equiv_newton_rcp\": %.1f}\n",
          p.name, nsm, khz / 1e3, dfma, dadd, dmul, 2.0 * dfma / 1e12, dfma / ((double)khz * 1e3) / nsm, lat_cycles, ddiv, dsqrt, drcp,
          dfma / ddiv, dfma / dsqrt, dfma / drcp);
   return 0;
}
