// kgpu_tiles.cuh -- device-side tile bookkeeping, maxima tracking and host transfer
// staging.  Replaces the per-tile loops of UpdateTiles.f90 / TimeStepper.f90 that touch
// field data; the ordered-list logic itself (which tile is switched on when) is
// replayed on the host from four flag bits per tile.
#pragma once
#include "kgpu_device.cuh"

namespace kgpu {

// cell-centred topography straight from the global vertex arrays
// (BT: const double * or const volatile double * -- the redistribution wave reads a bed other threads are correcting)
template <class BT>
__device__ __forceinline__ void centreTopoGlobal(const DevParams &P, const double *b0v, BT btv, int ci, int cj,
                                                 double &b0c, double &btc, double &bx, double &by) {
   size_t g = (size_t)(cj + YO) * P.pitch + (ci + XO);
   if (!P.oneD) {
      double a = b0v[g], b = b0v[g + 1], c = b0v[g + P.pitch], d = b0v[g + P.pitch + 1];
      double ta = 0.0, tb = 0.0, tc = 0.0, td = 0.0;
      if (btv) { ta = btv[g]; tb = btv[g + 1]; tc = btv[g + P.pitch]; td = btv[g + P.pitch + 1]; }
      b0c = 0.25 * kahan4(a, b, c, d);
      btc = 0.25 * kahan4(ta, tb, tc, td);
      bx = 0.5 * P.dxR * kahan8(b, tb, -a, -ta, d, td, -c, -tc);
      by = 0.5 * P.dyR * kahan8(c, tc, -a, -ta, d, td, -b, -tb);
   } else {
      double a = b0v[g], b = b0v[g + 1], ta = 0.0, tb = 0.0;
      if (btv) { ta = btv[g]; tb = btv[g + 1]; }
      b0c = 0.5 * (a + b);
      btc = 0.5 * (ta + tb);
      bx = P.dxR * kahan4(b, tb, -a, -ta);
      by = 0.0;
   }
}

struct StatePtrs {
   double *q[4];
};

// SetDefaultTileData / SetDomainBoundaryData / ActivateTile's "w = b0" (UpdateTiles.f90:342-370, 595-664).
// kind: 0 = default ghost data into every state buffer, 1 = dirichlet ghost, 2 = activation (w = b0c in S0 only)
struct AllStates {
   double *q[5][4];
   int n;  // buffers in use; q[0] is the current state
};
__global__ void tile_default_kernel(const DevParams P, AllStates S, const double *b0v,
                                    const double *btv, int tx, int ty, int kind, double bcH, double bcU, double bcV, double bcPsi) {
   int li = blockIdx.x * blockDim.x + threadIdx.x;
   int lj = blockIdx.y;
   if (li >= P.nX || lj >= P.nY) return;
   int ci = tx * P.nX + li, cj = ty * P.nY + lj;
   size_t g = (size_t)(cj + YO) * P.pitch + (ci + XO);
   double b0c, btc, bx, by;
   centreTopoGlobal(P, b0v, btv, ci, cj, b0c, btc, bx, by);
   if (kind == 2) {
      S.q[0][QW][g] = b0c;
      return;
   }
   double w = b0c, hu = 0.0, hv = 0.0, hpsi = 0.0;
   if (kind == 1) {
      double rho = P.rhow + (P.rhos - P.rhow) * bcPsi;
      double hpval = bcH / gamma2(P, bx, by);
      w = btc + hpval;
      w = w + b0c;
      hu = rho * bcH * bcU; hv = rho * bcH * bcV; hpsi = bcH * bcPsi;
   }
   for (int k = 0; k < S.n; k++) {
      S.q[k][QW][g] = w; S.q[k][QHU][g] = hu; S.q[k][QHV][g] = hv; S.q[k][QHPSI][g] = hpsi;
   }
}

// Near{N,S,E,W}Boundary (TimeStepper.f90:948-1150) reduced to four bits per active tile:
// bit0 N: wet cell with jj > nY - buf; bit1 S: jj <= buf; bit2 E: ii > nX - buf; bit3 W: ii <= buf.
__global__ void tile_flags_kernel(const DevParams P, StatePtrs S0, const double *b0v, const double *btv, const int *tiles,
                                  int buf, int *flagsOut) {
   int t = tiles[blockIdx.x];  // local tile index ty*nXt + tx
   int tx = t % P.nXt, ty = t / P.nXt;
   __shared__ int s_flags;
   if (threadIdx.x == 0) s_flags = 0;
   __syncthreads();
   int f = 0;
   int ncell = P.nX * P.nY;
   for (int k = threadIdx.x; k < ncell; k += blockDim.x) {
      int li = k % P.nX, lj = k / P.nX;
      bool n = lj >= P.nY - buf, s = lj < buf, e = li >= P.nX - buf, w = li < buf;
      if (P.oneD) { n = false; s = false; }
      if (!(n || s || e || w)) continue;
      int ci = tx * P.nX + li, cj = ty * P.nY + lj;
      size_t g = (size_t)(cj + YO) * P.pitch + (ci + XO);
      double b0c, btc, bx, by;
      centreTopoGlobal(P, b0v, btv, ci, cj, b0c, btc, bx, by);
      double Hn = computeHn(S0.q[QW][g], b0c, btc, gamma2(P, bx, by));
      if (Hn > P.Hneps) f |= (n ? 1 : 0) | (s ? 2 : 0) | (e ? 4 : 0) | (w ? 8 : 0);
   }
   if (f) atomicOr(&s_flags, f);
   __syncthreads();
   if (threadIdx.x == 0) flagsOut[blockIdx.x] = s_flags;
}

// Running maxima (TimeStepper.f90:1155-1303), evaluated on the state at the START of the
// step and stamped with its END time (quirk Q1, TimeStepper.f90:519-524).
template <int BX, int BY>
__global__ void __launch_bounds__(256) maxima_kernel(const DevParams P, StatePtrs S0, const double *b0v, const double *btv,
                                                        MaximaPtrs M, const uint8_t *tileMask, const int2 *blockList,
                                                        const Ctrl *ctrl, int allActive) {
   if (ctrl->failed) return;
   if (threadIdx.x >= BX * BY) return;
   const int2 bo = blockList[blockIdx.x];
   int ci = bo.x * BX + threadIdx.x % BX, cj = bo.y * BY + threadIdx.x / BX;
   if (ci >= P.NX || cj >= P.NY) return;
   if (!allActive) {
      int tx = ci / P.nX + 1, ty = cj / P.nY + 1;
      if (tileMask[ty * (P.nXt + 2) + tx] != 2) return;
   }
   size_t g = (size_t)(cj + YO) * P.pitch + (ci + XO);
   double tt = ctrl->t + ctrl->dt;  // nextT
   CellState q;
   q.w = S0.q[QW][g]; q.hu = S0.q[QHU][g]; q.hv = S0.q[QHV][g]; q.hpsi = S0.q[QHPSI][g];
   centreTopoGlobal(P, b0v, btv, ci, cj, q.b0, q.bt, q.bx, q.by);
   desingularise(P, q, true);
   updateMaxima(P, M, g, tt, q.Hn, sqrt(speed2(P, q.u, q.v, q.bx, q.by)), q.bt, q.psi);
}

// ---- host transfer staging: one tile <-> one contiguous staging buffer
// staging layout: [u13 (13,nX,nY)] [maxima 5*(nX,nY,2)] [tfirst (nX,nY)] [b0v (nX+1,nY+1)] [btv (nX+1,nY+1)]
__host__ __device__ inline size_t stageU13(const DevParams &) { return 0; }

__global__ void import_tile_kernel(const DevParams P, StatePtrs S0, MaximaPtrs M, const double *stage, int tx, int ty,
                                   int hasMaxima, int hasTfirst) {
   int li = blockIdx.x * blockDim.x + threadIdx.x;
   int lj = blockIdx.y;
   if (li >= P.nX || lj >= P.nY) return;
   int ci = tx * P.nX + li, cj = ty * P.nY + lj;
   size_t g = (size_t)(cj + YO) * P.pitch + (ci + XO);
   size_t ncell = (size_t)P.nX * P.nY, k = (size_t)lj * P.nX + li;
   const double *u = stage + k * 13;
   S0.q[QW][g] = u[0]; S0.q[QHU][g] = u[1]; S0.q[QHV][g] = u[2]; S0.q[QHPSI][g] = u[3];
   const double *mx = stage + ncell * 13;
   if (hasMaxima) {
      M.Hnmax[g] = mx[0 * 2 * ncell + k]; M.HnmaxT[g] = mx[0 * 2 * ncell + ncell + k];
      M.umax[g] = mx[1 * 2 * ncell + k];  M.umaxT[g] = mx[1 * 2 * ncell + ncell + k];
      M.emax[g] = mx[2 * 2 * ncell + k];  M.emaxT[g] = mx[2 * 2 * ncell + ncell + k];
      M.dmax[g] = mx[3 * 2 * ncell + k];  M.dmaxT[g] = mx[3 * 2 * ncell + ncell + k];
      M.psimax[g] = mx[4 * 2 * ncell + k]; M.psimaxT[g] = mx[4 * 2 * ncell + ncell + k];
   } else {
      M.Hnmax[g] = M.HnmaxT[g] = M.umax[g] = M.umaxT[g] = M.emax[g] = M.emaxT[g] = 0.0;
      M.dmax[g] = M.dmaxT[g] = M.psimax[g] = M.psimaxT[g] = 0.0;
   }
   M.tfirst[g] = hasTfirst ? stage[ncell * 23 + k] : -1.0;
}

// u13 as the reference's writers expect it.  qpre = state before the final implicit
// momentum correction: u, v are desingularised from it (quirk Q2, TimeStepper.f90:501-517).
__global__ void export_tile_kernel(const DevParams P, StatePtrs S0, StatePtrs Qpre, int usePre, const double *b0v,
                                   const double *btv, MaximaPtrs M, double *stage, int tx, int ty) {
   int li = blockIdx.x * blockDim.x + threadIdx.x;
   int lj = blockIdx.y;
   if (li >= P.nX || lj >= P.nY) return;
   int ci = tx * P.nX + li, cj = ty * P.nY + lj;
   size_t g = (size_t)(cj + YO) * P.pitch + (ci + XO);
   size_t ncell = (size_t)P.nX * P.nY, k = (size_t)lj * P.nX + li;
   CellState q;
   q.w = S0.q[QW][g]; q.hu = S0.q[QHU][g]; q.hv = S0.q[QHV][g]; q.hpsi = S0.q[QHPSI][g];
   centreTopoGlobal(P, b0v, btv, ci, cj, q.b0, q.bt, q.bx, q.by);
   double hu = q.hu, hv = q.hv;
   if (usePre) { q.hu = Qpre.q[QHU][g]; q.hv = Qpre.q[QHV][g]; }
   desingularise(P, q, true);
   double *u = stage + k * 13;
   u[0] = q.w; u[1] = hu; u[2] = hv; u[3] = q.hpsi; u[4] = q.Hn; u[5] = q.u; u[6] = q.v; u[7] = q.psi; u[8] = q.rho;
   u[9] = q.b0; u[10] = q.bt; u[11] = q.bx; u[12] = q.by;
   double *mx = stage + ncell * 13;
   mx[0 * 2 * ncell + k] = M.Hnmax[g]; mx[0 * 2 * ncell + ncell + k] = M.HnmaxT[g];
   mx[1 * 2 * ncell + k] = M.umax[g];  mx[1 * 2 * ncell + ncell + k] = M.umaxT[g];
   mx[2 * 2 * ncell + k] = M.emax[g];  mx[2 * 2 * ncell + ncell + k] = M.emaxT[g];
   mx[3 * 2 * ncell + k] = M.dmax[g];  mx[3 * 2 * ncell + ncell + k] = M.dmaxT[g];
   mx[4 * 2 * ncell + k] = M.psimax[g]; mx[4 * 2 * ncell + ncell + k] = M.psimaxT[g];
   stage[ncell * 23 + k] = M.tfirst[g];
}

// vertices of one tile: (nX+1) x (nY+1), i fastest.  dir 0: device -> stage, 1: stage -> device
// writeMask (dir 1): bit0 write right column, bit1 write top row, bit2 write top-right corner
__global__ void tile_vertices_kernel(const DevParams P, double *vfield, double *stage, int tx, int ty, int dir, int writeMask) {
   int li = blockIdx.x * blockDim.x + threadIdx.x;
   int lj = blockIdx.y;
   int nvy = P.oneD ? 1 : P.nY + 1;
   if (li > P.nX || lj >= nvy) return;
   int vi = tx * P.nX + li, vj = ty * P.nY + lj;
   size_t k = (size_t)lj * (P.nX + 1) + li;
   if (dir == 0) { stage[k] = vfield[(size_t)(vj + YO) * P.pitch + (vi + XO)]; return; }
   // periodic: vertex NX aliases vertex 0 (EqualiseTopographicBoundaryData across the wrap)
   if (P.periodic) { if (vi == P.NX) vi = 0; if (!P.oneD && vj == P.NY) vj = 0; }
   size_t g = (size_t)(vj + YO) * P.pitch + (vi + XO);
   bool right = (li == P.nX), top = (!P.oneD && lj == P.nY);
   if (right && top) { if (!(writeMask & 4)) return; }
   else if (right) { if (!(writeMask & 1)) return; }
   else if (top) { if (!(writeMask & 2)) return; }
   vfield[g] = stage[k];
}

// ------------------------------------------------------------------ analytic topography on the device
// TopogFuncs.f90 evaluated at the vertices of one tile, with the coordinates of Grid.f90:339-353 /
// UpdateTiles.f90:288-325 (cell centre of local index ii in tile gi: -xSize/2 + dx ((gi-1) nX + ii - 1/2); a vertex
// is half a cell below its cell, the last one half a cell above the last cell).  Same operation order as the host
// implementations (kestrel_b200/host/topog.py, host_cpp GetHeights): the algebraic functions give the host's bits,
// the transcendental ones differ by libm vs libdevice rounding.
struct TopogFn {
   int func, n;
   double p[8];
};
__global__ void tile_topog_kernel(const DevParams P, const TopogFn F, int gi, int gj, double *stage) {
   int li = blockIdx.x * blockDim.x + threadIdx.x;
   int lj = blockIdx.y;
   int nvy = P.oneD ? 1 : P.nY + 1;
   if (li > P.nX || lj >= nvy) return;
   const double PI = 3.141592653589793238462643383279502884;
   auto xc = [&](int ii) { return -0.5 * P.xSize + P.dx * ((gi - 1.0) * P.nX + (ii - 0.5)); };
   auto yc = [&](int jj) { return -0.5 * P.ySize + P.dy * ((gj - 1.0) * P.nY + (jj - 0.5)); };
   const double X = li < P.nX ? xc(li + 1) - 0.5 * P.dx : xc(P.nX) + 0.5 * P.dx;
   const double Y = lj < P.nY ? yc(lj + 1) - 0.5 * P.dy : yc(P.nY) + 0.5 * P.dy;
   const double *p = F.p;
   double b = 0.0;
   switch (F.func) {
      case KGPU_TOPOG_FLAT: b = 0.0; break;
      case KGPU_TOPOG_XSLOPE: b = p[0] * X; break;
      case KGPU_TOPOG_YSLOPE: b = p[0] * Y; break;
      case KGPU_TOPOG_XYSLOPE: b = p[0] * X + p[1] * Y; break;
      case KGPU_TOPOG_XSINSLOPE: b = p[0] * sin(X * (2.0 * PI / P.xSize)); break;
      case KGPU_TOPOG_XYSINSLOPE: b = p[0] * sin(X * (2.0 * PI / P.xSize)) * sin(Y * (2.0 * PI / P.ySize)); break;
      case KGPU_TOPOG_XHUMP: b = (X > -p[1] && X < p[1]) ? 0.5 * p[0] * (1.0 + cos(PI * X / p[1])) : 0.0; break;
      case KGPU_TOPOG_XTANH: b = p[1] * (1.0 + tanh((X - p[0]) / p[2])); break;
      case KGPU_TOPOG_XPARAB: b = p[0] * X * X; break;
      case KGPU_TOPOG_XYPARAB: b = p[0] * X * X + p[1] * Y * Y; break;
      case KGPU_TOPOG_XBISLOPE: {
         double phi1 = p[0] * PI / 180.0, phi2 = p[1] * PI / 180.0, lam = p[2];
         double a1 = tan(phi1), a2 = tan(phi2);
         b = -0.5 * (a1 + a2) * X + 0.5 * (a1 - a2) * lam * log(cosh(X / lam));
         break;
      }
      case KGPU_TOPOG_X2SLOPES: {
         double alpha = p[0], beta = p[1], R = p[2];
         double sa = sqrt(1.0 + alpha * alpha), sb = sqrt(1.0 + beta * beta);
         double xc0 = (sa - sb) * R / (alpha - beta);
         double zc0 = (alpha * sb - beta * sa) * R / (alpha - beta);
         double x1 = xc0 - alpha * R / sa, x2 = xc0 - beta * R / sb;
         double arc = zc0 - sqrt(fmax(R * R - (X - xc0) * (X - xc0), 0.0));
         b = X < x1 ? -alpha * X : (X > x2 ? -beta * X : arc);
         break;
      }
      default: break;
   }
   stage[(size_t)lj * (P.nX + 1) + li] = b;
}

}  // namespace kgpu
