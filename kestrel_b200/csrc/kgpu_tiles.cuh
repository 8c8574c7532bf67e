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

// Decomposed runs with dynamic tiles replay the tile operations of the ranks next door on their own halo (tile
// coordinates -1 or nXt / nYt, kgpu_dyn_host.inl): everything outside the two-cell halo of the storage is dropped.
__device__ __forceinline__ bool cellInStorage(const DevParams &P, int ci, int cj) {
   return ci >= -2 && ci < P.NX + 2 && (P.oneD ? cj == 0 : (cj >= -2 && cj < P.NY + 2));
}
__device__ __forceinline__ bool vertexInStorage(const DevParams &P, int vi, int vj) {
   return vi >= -2 && vi <= P.NX + 2 && (P.oneD ? vj == 0 : (vj >= -2 && vj <= P.NY + 2));
}

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
   if (!cellInStorage(P, ci, cj)) return;
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

// (Running maxima, TimeStepper.f90:1155-1303: updateMaxima in kgpu_device.cuh, called from the final stage of K1.)

// ---- host transfer staging: one tile <-> one contiguous staging buffer
// staging layout: [u13 (13,nX,nY)] [maxima 5*(nX,nY,2)] [tfirst (nX,nY)] [b0v (nX+1,nY+1)] [btv (nX+1,nY+1)]
__host__ __device__ inline size_t stageU13(const DevParams &) { return 0; }

__global__ void import_tile_kernel(const DevParams P, StatePtrs S0, MaximaPtrs M, const double *stage, int tx, int ty,
                                   int hasMaxima, int hasTfirst) {
   int li = blockIdx.x * blockDim.x + threadIdx.x;
   int lj = blockIdx.y;
   if (li >= P.nX || lj >= P.nY) return;
   int ci = tx * P.nX + li, cj = ty * P.nY + lj;
   if (!cellInStorage(P, ci, cj)) return;
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
   if (!vertexInStorage(P, vi, vj)) return;
   size_t g = (size_t)(vj + YO) * P.pitch + (vi + XO);
   bool right = (li == P.nX), top = (!P.oneD && lj == P.nY);
   if (right && top) { if (!(writeMask & 4)) return; }
   else if (right) { if (!(writeMask & 1)) return; }
   else if (top) { if (!(writeMask & 2)) return; }
   vfield[g] = stage[k];
}

// ------------------------------------------------------------------ DEM ingest on the device
// TileHeightData (dem.f90:260-356): the height of a tile vertex is the mean of four bicubic interpolations of the
// raster at the vertex +- half a cell in x and y; Bicubic_r (Interp2d.f90:214-292) builds the 16 coefficients of the
// bicubic patch from the values and centred finite differences at the four pixels around the point with the constant
// matrix m (Interp2d.f90:57-74) and evaluates it in Horner form.  The raster is held on the device as Elev(i, j),
// i fastest, 1-based pixel indices as in the reference; image coordinates imgx = (E - originX) / pixelW + 1.
// (The reference reads one raster section per tile through GDAL; here the library holds the section that covers the
// domain and every tile indexes into it.  Stencil indices are clamped to the raster.)
struct RasterDesc {
   const double *elev;
   int nx, ny;
   double originX, originY, pixelW, pixelH, centreE, centreN;
};
__constant__ double c_bicubicM[256] = {
   1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
   0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0,
   -3, 0, 3, 0, 0, 0, 0, 0, -2, 0, -1, 0, 0, 0, 0, 0,
   2, 0, -2, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0,
   0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
   0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0,
   0, 0, 0, 0, -3, 0, 3, 0, 0, 0, 0, 0, -2, 0, -1, 0,
   0, 0, 0, 0, 2, 0, -2, 0, 0, 0, 0, 0, 1, 0, 1, 0,
   -3, 3, 0, 0, -2, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
   0, 0, 0, 0, 0, 0, 0, 0, -3, 3, 0, 0, -2, -1, 0, 0,
   9, -9, -9, 9, 6, 3, -6, -3, 6, -6, 3, -3, 4, 2, 2, 1,
   -6, 6, 6, -6, -4, -2, 4, 2, -3, 3, -3, 3, -2, -1, -2, -1,
   2, -2, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
   0, 0, 0, 0, 0, 0, 0, 0, 2, -2, 0, 0, 1, 1, 0, 0,
   -6, 6, 6, -6, -3, -3, 3, 3, -4, 4, -2, 2, -2, -2, -1, -1,
   4, -4, -4, 4, 2, 2, -2, -2, 2, -2, 2, -2, 1, 1, 1, 1};
__device__ inline double bicubicR(const RasterDesc &R, double x, double y) {
   const int x0 = (int)floor(x), x1 = (int)ceil(x), y0 = (int)floor(y), y1 = (int)ceil(y);
   auto z = [&](int i, int j) -> double {
      i = i < 1 ? 1 : (i > R.nx ? R.nx : i);
      j = j < 1 ? 1 : (j > R.ny ? R.ny : j);
      return R.elev[(size_t)(j - 1) * R.nx + (i - 1)];
   };
   double f[16];
   const double dx0 = (double)(x1 - x0 + 1), dx1 = (double)(x1 + 1 - x0), dy0 = (double)(y1 - y0 + 1), dy1 = (double)(y1 + 1 - y0);
   f[0] = z(x0, y0); f[1] = z(x1, y0); f[2] = z(x0, y1); f[3] = z(x1, y1);
   f[4] = (z(x1, y0) - z(x0 - 1, y0)) / dx0;
   f[5] = (z(x1 + 1, y0) - z(x0, y0)) / dx1;
   f[6] = (z(x1, y1) - z(x0 - 1, y1)) / dx0;
   f[7] = (z(x1 + 1, y1) - z(x0, y1)) / dx1;
   f[8] = (z(x0, y1) - z(x0, y0 - 1)) / dy0;
   f[9] = (z(x1, y1) - z(x1, y0 - 1)) / dy0;
   f[10] = (z(x0, y1 + 1) - z(x0, y0)) / dy1;
   f[11] = (z(x1, y1 + 1) - z(x1, y0)) / dy1;
   f[12] = ((z(x1, y1) - z(x0 - 1, y1)) - (z(x1, y0 - 1) - z(x0 - 1, y0 - 1))) / dx0 / dy0;
   f[13] = ((z(x1 + 1, y1) - z(x0, y1)) - (z(x1 + 1, y0 - 1) - z(x0, y0 - 1))) / dx1 / dy0;
   f[14] = ((z(x1, y1 + 1) - z(x0 - 1, y1 + 1)) - (z(x1, y0) - z(x0 - 1, y0))) / dx0 / dy1;
   f[15] = ((z(x1 + 1, y1 + 1) - z(x0, y1 + 1)) - (z(x1 + 1, y0) - z(x0, y0))) / dx1 / dy1;
   double a[16];   // aVec = matmul(m, fVec); a(i, j) = aVec(4 (i - 1) + j)
   for (int r = 0; r < 16; r++) {
      double acc = 0.0;
      for (int c = 0; c < 16; c++) acc = acc + c_bicubicM[r * 16 + c] * f[c];
      a[r] = acc;
   }
   const double t = (x == (double)x0) ? 0.0 : (x - (double)x0) / (double)(x1 - x0);
   const double u = (y == (double)y0) ? 0.0 : (y - (double)y0) / (double)(y1 - y0);
   double ans = 0.0;
   for (int i = 3; i >= 0; i--) ans = t * ans + ((a[i * 4 + 3] * u + a[i * 4 + 2]) * u + a[i * 4 + 1]) * u + a[i * 4 + 0];
   return ans;
}
// heights of one tile's vertices from the raster, into the staging buffer (same layout as tile_topog_kernel's)
__global__ void tile_raster_kernel(const DevParams P, const RasterDesc R, int gi, int gj, double *stage) {
   int li = blockIdx.x * blockDim.x + threadIdx.x;
   int lj = blockIdx.y;
   int nvy = P.oneD ? 1 : P.nY + 1;
   if (li > P.nX || lj >= nvy) return;
   auto xc = [&](int ii) { return -0.5 * P.xSize + P.dx * ((gi - 1.0) * P.nX + (ii - 0.5)); };
   auto yc = [&](int jj) { return -0.5 * P.ySize + P.dy * ((gj - 1.0) * P.nY + (jj - 0.5)); };
   const double X = li < P.nX ? xc(li + 1) - 0.5 * P.dx : xc(P.nX) + 0.5 * P.dx;   // x_vertex, y_vertex (UpdateTiles.f90:288-325)
   const double Y = lj < P.nY ? yc(lj + 1) - 0.5 * P.dy : yc(P.nY) + 0.5 * P.dy;
   double s = 0.0;   // 0.25 * sum(hcorner), hcorner(k), k = jj + 2 (ii - 1)
   for (int ii = 1; ii <= 2; ii++)
      for (int jj = 1; jj <= 2; jj++) {
         const double E = R.centreE + X + ((double)(ii - 1) - 0.5) * P.dx;
         const double N = R.centreN + Y + ((double)(jj - 1) - 0.5) * P.dy;
         const double imgx = (E - R.originX) / R.pixelW + 1, imgy = (N - R.originY) / R.pixelH + 1;
         s = s + bicubicR(R, imgx, imgy);
      }
   stage[(size_t)lj * (P.nX + 1) + li] = 0.25 * s;
}

// ------------------------------------------------------------------ initial conditions on the device
// LoadSourceConditions (SetSources.f90:47-392).  Cell coordinates: GridToPhysical (Grid.f90:339-353) = cellX / cellY.
struct ShapeTable {
   const kgpu_cap *caps;
   const kgpu_cube *cubes;
   const DevSource *src;
   int ncaps, ncubes, nsrc;
};
__device__ __forceinline__ bool capHit(const DevParams &P, const kgpu_cap &c, double x, double y, double &R2) {
   R2 = P.oneD ? (x - c.x) * (x - c.x) : (x - c.x) * (x - c.x) + (y - c.y) * (y - c.y);
   return R2 <= c.radius * c.radius;
}
__device__ __forceinline__ bool cubeHit(const DevParams &P, const kgpu_cube &c, double x, double y) {
   return (fabs(x - c.x) <= 0.5 * c.length) && (P.oneD || fabs(y - c.y) <= 0.5 * c.width);
}
__device__ __forceinline__ bool sourceHit(const DevParams &P, const DevSource &s, double x, double y) {
   const double R2 = P.oneD ? (x - s.x) * (x - s.x) : (x - s.x) * (x - s.x) + (y - s.y) * (y - s.y);
   return R2 <= s.radius * s.radius;   // <= here, strict < in the RHS (quirk Q9)
}
// pass 1, one block per tile of the local domain: bit 0 = some cell centre lies in a cap or cube, bit 1 = in a source disc
__global__ void shape_touch_kernel(const DevParams P, const ShapeTable T, int *touch) {
   const int t = blockIdx.x, tx = t % P.nXt, ty = t / P.nXt;
   __shared__ int s_f;
   if (threadIdx.x == 0) s_f = 0;
   __syncthreads();
   int f = 0;
   for (int k = threadIdx.x; k < P.nX * P.nY; k += blockDim.x) {
      const int ci = tx * P.nX + k % P.nX, cj = ty * P.nY + k / P.nX;
      const double x = cellX(P, ci), y = cellY(P, cj);
      double R2;
      for (int c = 0; c < T.ncaps; c++) if (capHit(P, T.caps[c], x, y, R2)) f |= 1;
      for (int c = 0; c < T.ncubes; c++) if (cubeHit(P, T.cubes[c], x, y)) f |= 1;
      for (int c = 0; c < T.nsrc; c++) if (sourceHit(P, T.src[c], x, y)) f |= 2;
   }
   if (f) atomicOr(&s_f, f);
   __syncthreads();
   if (threadIdx.x == 0) touch[t] = s_f;
}
// pass 2, one block per ACTIVE tile: the shapes cell by cell (the tile's cells start from ActivateTile's state
// w = b0, everything else 0: UpdateTiles.f90:342-370), NumCellsInSrc, and the seed of the first activation scan
__global__ void shape_raster_kernel(const DevParams P, const ShapeTable T, StatePtrs S0, MaximaPtrs M, const double *b0v, const int *tiles,
                                    int buf, int *seedFlags, int *numCells) {
   const int t = tiles[blockIdx.x], tx = t % P.nXt, ty = t / P.nXt;
   __shared__ int s_f;
   if (threadIdx.x == 0) s_f = 0;
   __syncthreads();
   int flags = 0;
   for (int k = threadIdx.x; k < P.nX * P.nY; k += blockDim.x) {
      const int li = k % P.nX, lj = k / P.nX;
      const int ci = tx * P.nX + li, cj = ty * P.nY + lj;
      const size_t g = (size_t)(cj + YO) * P.pitch + (ci + XO);
      const double x = cellX(P, ci), y = cellY(P, cj);
      double b0c, btc, bx, by;
      centreTopoGlobal(P, b0v, (const double *)nullptr, ci, cj, b0c, btc, bx, by);
      const double gam = gamma2(P, bx, by);
      double w = b0c, hu = 0.0, hv = 0.0, hpsi = 0.0, Hn = 0.0, Hnmax = 0.0, psimax = 0.0;
      for (int c = 0; c < T.ncaps; c++) {
         const kgpu_cap &C = T.caps[c];
         double R2;
         if (!capHit(P, C, x, y, R2)) continue;
         const double rho = P.rhow + (P.rhos - P.rhow) * C.psi;
         if (C.shape == KGPU_SHAPE_FLAT) {
            w = w + C.height / gam; Hn = Hn + C.height; Hnmax = Hnmax + C.height;
            hu = hu + rho * C.height * C.u;
            if (P.oneD) psimax = psimax + C.psi; else hv = hv + rho * C.height * C.v;
            hpsi = hpsi + C.psi * C.height;
         } else if (C.shape == KGPU_SHAPE_PARA) {
            const double prof = C.height * (1.0 - R2 / C.radius / C.radius);
            w = w + prof / gam; Hn = Hn + prof; Hnmax = Hnmax + prof;
            hu = hu + rho * C.height * C.u;          // no parabola factor on the momentum (quirk Q10)
            if (P.oneD) psimax = psimax + C.psi; else hv = hv + rho * C.height * C.v;
            hpsi = hpsi + C.psi * C.height * (1.0 - R2 / C.radius / C.radius);
         } else {   // level: `height` is the free-surface elevation
            const double hp = C.height - b0c;
            if (P.oneD) {
               if (hp > 0.0) {
                  w = w + hp; Hnmax = Hnmax + hp * gam; Hn = Hn + hp * gam;
                  hu = hu + rho * hp * gam * C.u; hpsi = hpsi + C.psi * hp * gam; psimax = psimax + C.psi;
               }
            } else {
               const double H = hp * gam;
               if (H > 0.0) {
                  w = w + hp; Hn = Hn + H; Hnmax = Hnmax + H;
                  hu = hu + rho * H * C.u; hv = hv + rho * H * C.v; hpsi = hpsi + C.psi * H;
               }
            }
         }
      }
      for (int c = 0; c < T.ncubes; c++) {
         const kgpu_cube &C = T.cubes[c];
         if (!cubeHit(P, C, x, y)) continue;
         if (C.shape == KGPU_SHAPE_LEVEL) {
            const double hp = C.height - b0c, H = hp * gam;
            if (H > 0.0) { w = w + hp; Hn = Hn + H; Hnmax = Hnmax + H; hpsi = hpsi + (P.oneD ? C.psi * H : H * C.psi); }
         } else {
            w = w + C.height / gam; Hn = Hn + C.height; Hnmax = Hnmax + C.height; hpsi = hpsi + C.psi * C.height;
         }
      }
      for (int c = 0; c < T.nsrc; c++) if (sourceHit(P, T.src[c], x, y)) atomicAdd(&numCells[c], 1);
      S0.q[QW][g] = w; S0.q[QHU][g] = hu; S0.q[QHV][g] = hv; S0.q[QHPSI][g] = hpsi;
      M.Hnmax[g] = Hnmax; M.psimax[g] = psimax;
      if (Hn > P.Hneps) {
         if (!P.oneD) { if (lj >= P.nY - buf) flags |= 1; if (lj < buf) flags |= 2; }
         if (li >= P.nX - buf) flags |= 4;
         if (li < buf) flags |= 8;
      }
   }
   if (flags) atomicOr(&s_f, flags);
   __syncthreads();
   if (threadIdx.x == 0) seedFlags[blockIdx.x] = s_f;
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
      case KGPU_TOPOG_USGS:
      case KGPU_TOPOG_FLUME: {   // TopogFuncs.f90:145-242
         double theta0 = 31.0, theta1 = 2.4, xwall = 8.5, wallW = 2.0, wallH = p[0], sigma = p[1];
         if (F.func == KGPU_TOPOG_FLUME) { theta0 = p[0]; theta1 = p[1]; xwall = p[2]; wallW = p[3]; wallH = p[4]; sigma = p[5]; }
         double alpha = 8.5 / (asinh(-tan(4.0 * PI / 180.0)) - asinh(-tan(theta0 * PI / 180.0)));
         double xc0 = -alpha * asinh(-tan(theta0 * PI / 180.0));
         double zc0 = -alpha * cosh((-xc0) / alpha);
         double x1 = xc0 + alpha * asinh(-tan(theta1 * PI / 180.0));
         if (X < 0.0) b = -tan(theta0 * PI / 180.0) * X;
         else if (X > x1) b = zc0 + alpha * cosh((x1 - xc0) / alpha) - tan(theta1 * PI / 180.0) * (X - x1);
         else b = zc0 + alpha * cosh((X - xc0) / alpha);
         if (X < xwall)
            b = b + 0.5 * wallH * (tanh(sigma * (Y - 0.5 * wallW)) - tanh(sigma * (Y - 1.5 * wallW)) + tanh(sigma * (Y + 1.5 * wallW)) -
                                   tanh(sigma * (Y + 0.5 * wallW)));
         break;
      }
      case KGPU_TOPOG_CHANNEL_POWERLAW: b = p[0] * X + cos(atan(p[0])) * pow(fabs(Y) / p[1], p[2]); break;                 // :255-276
      case KGPU_TOPOG_CHANNEL_TRAPEZIUM: b = p[0] * X + cos(atan(p[0])) * fmax(0.0, p[2] * (fabs(Y) - 0.5 * p[1])); break;  // :288-308
      case KGPU_TOPOG_XTRISLOPE: {   // TopogFuncs.f90:346-388
         double phi1 = p[0] * PI / 180.0, phi2 = p[1] * PI / 180.0, phi3 = p[2] * PI / 180.0, lam = p[3], x1 = p[4], x2 = p[5];
         double s1 = tan(phi1), s2 = tan(phi2), s3 = tan(phi3);
         double c2 = (x1 - 0.5 * lam) * 0.5 * (s1 - s2);
         double c3 = (x1 + 0.5 * lam) * 0.5 * (s1 - s2) + c2;
         double c4 = (x2 - 0.5 * lam) * 0.5 * (s2 - s3) + c3;
         double c5 = (x2 + 0.5 * lam) * 0.5 * (s2 - s3) + c4;
         if (X < x1 - 0.5 * lam) b = s1 * X;
         else if (X < x1 + 0.5 * lam) { double A = 0.5 * (s2 - s1) * lam / PI; b = A * sin((X - x1) * PI / lam - 0.5 * PI) + 0.5 * (s1 + s2) * X + c2; }
         else if (X < x2 - 0.5 * lam) b = s2 * X + c3;
         else if (X < x2 + 0.5 * lam) { double A = 0.5 * (s3 - s2) * lam / PI; b = A * sin((X - x2) * PI / lam - 0.5 * PI) + 0.5 * (s2 + s3) * X + c4; }
         else b = s3 * X + c5;
         break;
      }
      default: break;
   }
   stage[(size_t)lj * (P.nX + 1) + li] = b;
}

}  // namespace kgpu
