// kgpu_morpho.cuh -- the morphodynamic operator M of the Strang split
// (MorphodynamicRHS.f90:72-305, TimeStepper.f90:532-781, Redistribute.f90).
//
// Per Runge-Kutta stage: (1) E - D at cell centres, zeroed beside dry cells; (2) vertex
// source -gamma_v/(4 psi_b) * Kahan(E-D of the 4 cells) and the bed update with the erosion
// depth clamp; (3) the linear, conservative update of w and Hnpsi from the new bed.  The
// velocities u, v are frozen over M at the values of the last hydraulic RHS evaluation
// (ComputeDesingularisedVariables(computeVelocities=.false.), TimeStepper.f90:609).
#pragma once
#include "kgpu_device.cuh"
#include "kgpu_tiles.cuh"

namespace kgpu {

struct RedistEntry {
   double excess;
   int i, j;
};

// periodic image of an index.  Nearly every call is already in range: the two run-time modulos (≈ 40 instructions, and four
// such indices per vertex made morpho_bed_kernel issue-bound: 81 % of the issue slots busy at 2 TB/s) are only paid at the wrap
__device__ __forceinline__ int wrapIdx(int i, int n, int periodic) {
   if (!periodic || (unsigned)i < (unsigned)n) return i;
   return ((i % n) + n) % n;
}

struct MorphoArgs {
   const double *w, *hpsi;      // stage state (w, Hnpsi)
   const double *w0, *hpsi0;    // state at the start of M (intermed0)
   const double *U, *V;         // frozen velocities
   const double *b0v;
   const double *btk;           // bed of the stage state
   const double *bt0;           // bed at the start of M
   double *btn;                 // bed being written
   double *wn, *hpsin;          // stage state being written
   double *EmD;
   const uint8_t *tileMask;
   const int2 *blockList;
   Ctrl *ctrl;
   int allActive;
   double a0, a1;               // RK weights: bt_new = a0*bt0 + a1*(btk + dt*rhs); a0 = 0 selects stage 1
   double dtMorpho;
   // cell-centred planes of the stage bed btk (bt, bx, by at centres and the clamped Hn of the stage
   // state), written once per cell by the previous stage's cell kernel (stage 1: the hydraulic
   // topography planes + morpho_prepare) instead of being re-derived from the vertices by every reader
   const double *b0c;                   // static
   const double *cBt, *cBx, *cBy, *cHn; // of (btk, stage state)
   const double *zBt, *zBx, *zBy, *zGam; // of bt0 (= the hydraulic topography planes during M)
   double *nBt, *nBx, *nBy, *nHn;       // of (btn, state being written)
};

__device__ __forceinline__ bool cellTileActive(const DevParams &P, const uint8_t *mask, int allActive, int ci, int cj) {
   if (allActive) return ci >= -2 && ci < P.NX + 2 && cj >= -2 && cj < P.NY + 2;
   int tx = (ci >= 0 ? ci / P.nX : -1) + 1, ty = (cj >= 0 ? cj / P.nY : -1) + 1;
   if (tx > P.nXt + 1 || ty > P.nYt + 1) return false;
   return mask[ty * (P.nXt + 2) + tx] == 2;
}

// stored u(Hn) of a container: ComputeHn clamped at zero (HydraulicRHS.f90:808, 840-842)
__device__ __forceinline__ double storedHn(const DevParams &P, const double *w, const double *b0v, const double *btv, int ci, int cj) {
   double b0c, btc, bx, by;
   centreTopoGlobal(P, b0v, btv, ci, cj, b0c, btc, bx, by);
   double Hn = computeHn(w[(size_t)(cj + YO) * P.pitch + (ci + XO)], b0c, btc, gamma2(P, bx, by));
   return Hn < 0.0 ? 0.0 : Hn;
}

// velocities of the hydraulic result as its 4th RHS evaluation left them (pre-correction momenta)
template <int BX, int BY>
__global__ void __launch_bounds__(256) morpho_prepare_kernel(const DevParams P, const double *w, const double *hpsi, const double *huPre,
                                                                const double *hvPre, const double *b0c, const double *btc,
                                                                const double *bxc, const double *byc, double *U, double *V, double *HnOut,
                                                                const uint8_t *tileMask, const int2 *blockList, int allActive) {
   if (threadIdx.x >= BX * BY) return;
   const int2 bo = blockList[blockIdx.x];
   int ci = bo.x * BX + threadIdx.x % BX, cj = bo.y * BY + threadIdx.x / BX;
   if (ci >= P.NX || cj >= P.NY) return;
   if (!cellTileActive(P, tileMask, allActive, ci, cj)) return;
   size_t g = (size_t)(cj + YO) * P.pitch + (ci + XO);
   CellState q;
   q.w = w[g]; q.hpsi = hpsi[g]; q.hu = huPre[g]; q.hv = hvPre[g];
   q.b0 = b0c[g]; q.bt = btc[g]; q.bx = bxc[g]; q.by = byc[g];   // centre topography of bt0 (hydraulic planes)
   desingularise(P, q, true);
   U[g] = q.u; V[g] = q.v;
   HnOut[g] = q.Hn;   // clamped at zero: the stored u(Hn) the dry test of the first stage reads
}

// CalculateMorphodynamicRHS, cell part (MorphodynamicRHS.f90:96-145)
template <int BX, int BY>
__global__ void __launch_bounds__(256) morpho_emd_kernel(const DevParams P, const MorphoArgs A) {
   if (threadIdx.x >= BX * BY) return;
   const int2 bo = A.blockList[blockIdx.x];
   int ci = bo.x * BX + threadIdx.x % BX, cj = bo.y * BY + threadIdx.x / BX;
   if (ci >= P.NX || cj >= P.NY) return;
   if (!cellTileActive(P, A.tileMask, A.allActive, ci, cj)) return;
   size_t g = (size_t)(cj + YO) * P.pitch + (ci + XO);
   CellState q;
   q.w = A.w[g]; q.hpsi = A.hpsi[g]; q.hu = 0.0; q.hv = 0.0;
   q.b0 = A.b0c[g]; q.bt = A.cBt[g]; q.bx = A.cBx[g]; q.by = A.cBy[g];
   desingularise(P, q, false);
   q.u = A.U[g]; q.v = A.V[g];
   double eps = P.Hneps;
   // neighbours: the plane inside active tiles (their halo images included), the vertices elsewhere
   auto nbHn = [&](int i, int j) -> double {
      if (cellTileActive(P, A.tileMask, A.allActive, i, j)) return A.cHn[(size_t)(j + YO) * P.pitch + (i + XO)];
      return storedHn(P, A.w, A.b0v, A.btk, i, j);
   };
   bool dry = q.Hn < eps || nbHn(ci - 1, cj) < eps || nbHn(ci + 1, cj) < eps;
   if (!P.oneD) dry = dry || nbHn(ci, cj - 1) < eps || nbHn(ci, cj + 1) < eps;
   A.EmD[g] = dry ? 0.0 : erosionMinusDeposition(P, q);
}

// BtSourceTerm (MorphodynamicRHS.f90:154-305) + the bed stage update (TimeStepper.f90:574-581 etc.)
__global__ void morpho_bed_kernel(const DevParams P, const MorphoArgs A) {
   int vi = blockIdx.x * blockDim.x + threadIdx.x;
   int vj = blockIdx.y;
   int nvy = P.oneD ? 1 : P.NY + 1;
   if (vi > P.NX || vj >= nvy) return;
   if (P.periodic && (vi == P.NX || (!P.oneD && vj == P.NY))) return;  // aliases, refreshed by the halo fill
   // decomposed run: cells beyond the block are images held in the halo (E - D included, exchanged by the
   // host after the cell kernel); the seam vertices are then computed identically by both ranks
   const bool wrapOrHalo = P.periodic || P.haloValid;
   // vertex of an active tile?
   bool any = false;
   for (int dj = (P.oneD ? 0 : -1); dj <= 0; dj++)
      for (int di = -1; di <= 0; di++) {
         int i = vi + di, j = vj + dj;
         if (!wrapOrHalo && (i < 0 || i >= P.NX || j < 0 || j >= P.NY)) continue;
         if (cellTileActive(P, A.tileMask, A.allActive, i, j)) any = true;
      }
   if (!any) return;
   size_t gv = (size_t)(vj + YO) * P.pitch + (vi + XO);
   double psib = 1.0 - P.BedPorosity;
   double rhs;
   auto emd = [&](int i, int j) -> double {
      if (!wrapOrHalo && (i < 0 || i >= P.NX || j < 0 || j >= P.NY)) return 0.0;
      i = wrapIdx(i, P.NX, P.periodic); j = wrapIdx(j, P.NY, P.periodic);
      return A.EmD[(size_t)(j + YO) * P.pitch + (i + XO)];
   };
   // centre slopes of a cell next to the vertex: the stage's plane inside active tiles (halo images
   // included), the vertex arrays elsewhere
   auto slopes = [&](int i, int j, double &bx_, double &by_) {
      if (cellTileActive(P, A.tileMask, A.allActive, i, j)) {
         size_t gc = (size_t)(j + YO) * P.pitch + (i + XO);
         bx_ = A.cBx[gc]; by_ = A.cBy[gc];
      } else {
         double b0c_, btc_;
         centreTopoGlobal(P, A.b0v, A.btk, i, j, b0c_, btc_, bx_, by_);
      }
   };
   if (!P.oneD) {
      double bx[4], by[4];
      slopes(vi - 1, vj - 1, bx[0], by[0]);
      slopes(vi - 1, vj, bx[1], by[1]);
      slopes(vi, vj - 1, bx[2], by[2]);
      slopes(vi, vj, bx[3], by[3]);
      double dbdx = 0.25 * kahan4(bx[0], bx[1], bx[2], bx[3]);
      double dbdy = 0.25 * kahan4(by[0], by[1], by[2], by[3]);
      double gam = gamma2(P, dbdx, dbdy);
      rhs = -0.25 * gam / psib * kahan4(emd(vi - 1, vj - 1), emd(vi - 1, vj), emd(vi, vj - 1), emd(vi, vj));
   } else {
      bool lAct = (wrapOrHalo || vi - 1 >= 0) && cellTileActive(P, A.tileMask, A.allActive, vi - 1, 0);
      bool rAct = (wrapOrHalo || vi < P.NX) && cellTileActive(P, A.tileMask, A.allActive, vi, 0);
      double bxl = 0.0, bxr = 0.0, byd;
      if (lAct) slopes(vi - 1, 0, bxl, byd);
      if (rAct) slopes(vi, 0, bxr, byd);
      if (lAct && rAct) {
         double dbdx = 0.5 * (bxl + bxr);
         double gam = gamma2(P, dbdx, 0.0);
         rhs = -0.5 * gam * (emd(vi - 1, 0) + emd(vi, 0)) / psib;
      } else {
         double dbdx = 0.5 * (lAct ? bxl : bxr);
         double gam = gamma2(P, dbdx, 0.0);
         rhs = -0.5 * gam * (lAct ? emd(vi - 1, 0) : emd(vi, 0)) / psib;
      }
   }
   double val;
   if (A.a0 == 0.0) val = A.bt0[gv] + A.dtMorpho * rhs;
   else val = A.a0 * A.bt0[gv] + A.a1 * (A.btk[gv] + A.dtMorpho * rhs);
   A.btn[gv] = fmax(-P.EroDepth, val);
}

// linear updates of w and Hnpsi from the new bed (TimeStepper.f90:587-610)
template <int BX, int BY>
__global__ void __launch_bounds__(256) morpho_cell_kernel(const DevParams P, const MorphoArgs A) {
   if (threadIdx.x >= BX * BY) return;
   const int2 bo = A.blockList[blockIdx.x];
   int ci = bo.x * BX + threadIdx.x % BX, cj = bo.y * BY + threadIdx.x / BX;
   if (ci >= P.NX || cj >= P.NY) return;
   if (!cellTileActive(P, A.tileMask, A.allActive, ci, cj)) return;
   size_t g = (size_t)(cj + YO) * P.pitch + (ci + XO);
   double b0c, btnc, bxn, byn;
   const double bt0c = A.zBt[g];                       // centre planes of bt0
   centreTopoGlobal(P, A.b0v, A.btn, ci, cj, b0c, btnc, bxn, byn);
   double gamold = A.zGam[g], gamnew = gamma2(P, bxn, byn);
   double Hn_old = computeHn(A.w0[g], b0c, bt0c, gamold);
   double db = btnc - bt0c;
   double w = -db / gamnew / gamnew;
   w = w + btnc;
   w = w + Hn_old * gamold / gamnew / gamnew;
   w = w + b0c;
   A.wn[g] = w;
   double Hnpsi = -(1.0 - P.BedPorosity) * db / gamnew;
   Hnpsi = Hnpsi + A.hpsi0[g] * gamold / gamnew;
   A.hpsin[g] = Hnpsi;
   // centre planes of the new bed and the clamped depth of the new state, for the next stage's readers
   A.nBt[g] = btnc; A.nBx[g] = bxn; A.nBy[g] = byn;
   double Hn_new = computeHn(w, b0c, btnc, gamnew);
   A.nHn[g] = Hn_new < 0.0 ? 0.0 : Hn_new;
}

struct CheckArgs {
   const double *w0, *hpsi0, *w3;
   const double *b0v, *bt0, *bt3;
   const double *b0c, *zBt, *zGam;        // centre planes of bt0 (hydraulic topography planes)
   const double *c3Bt, *c3Bx, *c3By;      // centre planes of bt3 (written by the third stage's cell kernel)
   const uint8_t *tileMask;
   const int2 *blockList;
   Ctrl *ctrl;
   RedistEntry *list;
   int listCap;
   int allActive;
};

// Redistribute.f90:158-198
__device__ __forceinline__ void excessDeposition(const DevParams &P, double hpsi0, double Hn_old, double gamold, double deltaBt, double &excess) {
   excess = -(hpsi0 * gamold / (1.0 - P.BedPorosity) - deltaBt);
   excess = fmax(excess, -(Hn_old * gamold - deltaBt));
}

// the two checks on the morphodynamic update (TimeStepper.f90:709-753), order-free form:
// refine = OR over cells; the redistribution list is only used when nothing refined.
template <int BX, int BY>
__global__ void __launch_bounds__(256) morpho_check_kernel(const DevParams P, const CheckArgs A) {
   if (threadIdx.x >= BX * BY) return;
   const int2 bo = A.blockList[blockIdx.x];
   int ci = bo.x * BX + threadIdx.x % BX, cj = bo.y * BY + threadIdx.x / BX;
   if (ci >= P.NX || cj >= P.NY) return;
   if (!cellTileActive(P, A.tileMask, A.allActive, ci, cj)) return;
   size_t g = (size_t)(cj + YO) * P.pitch + (ci + XO);
   const double EPS = 2.220446049250313e-16;
   const double b0c = A.b0c[g], bt0c = A.zBt[g], bt3c = A.c3Bt[g];
   double gamold = A.zGam[g], gamnew = gamma2(P, A.c3Bx[g], A.c3By[g]);
   double Hn_old = computeHn(A.w0[g], b0c, bt0c, gamold);
   double Hn_new = computeHn(A.w3[g], b0c, bt3c, gamnew);
   double deltaBt = bt3c - bt0c;
   // psiold = intermed0's stored psi (desingularised, HydraulicRHS.f90:847)
   double Hc = Hn_old < 0.0 ? 0.0 : Hn_old, hs = A.hpsi0[g] < 0.0 ? 0.0 : A.hpsi0[g];
   double psiold = fmin(2.0 * Hc * hs / (Hc * Hc + fmax(Hc * Hc, P.Hneps * P.Hneps)), P.maxPack);
   double excess;
   excessDeposition(P, A.hpsi0[g], Hn_old, gamold, deltaBt, excess);
   if (excess > EPS && deltaBt > EPS && psiold > -EPS) {
      if (Hn_old < P.EroCritH || fabs(Hn_new - Hn_old) < P.EroCritH) {
         int k = atomicAdd(&A.ctrl->nRedist, 1);
         if (k < A.listCap) { A.list[k].excess = excess; A.list[k].i = ci; A.list[k].j = cj; }
      } else {
         A.ctrl->refineMorpho = 1;
      }
      return;
   }
   if (Hn_old < P.EroCritH) return;
   double rel = fabs(Hn_new - Hn_old) / fabs(Hn_old);
   if (rel > 0.1) A.ctrl->refineMorpho = 1;
}

struct RedistArgs {
   const double *w0, *hpsi0;
   double *w3, *hpsi3;
   const double *b0v, *bt0;
   double *bt3;
   const uint8_t *tileMask;
   const RedistEntry *list;
   int n;
   int allActive;
   Ctrl *ctrl;
};

// RedistributeCell (Redistribute.f90:249-475) for one listed cell.  The corrected bed and the refreshed
// cells are read and written through volatile pointers: in the wave kernel below other threads are
// correcting cells further away at the same time.  Returns false when the cell has no depositional
// vertex (the reference then asks for a smaller time step, Redistribute.f90:304-309).
__device__ __forceinline__ bool redistributeCell(const DevParams &P, const RedistArgs &A, int i, int j) {
   const double EPS = 2.220446049250313e-16;
   const int pitch = P.pitch;
   volatile double *bt3 = A.bt3, *w3 = A.w3, *hpsi3 = A.hpsi3;
   const volatile double *bt3r = A.bt3;
   auto vix = [&](int vi, int vj) -> size_t {
      if (P.periodic) { vi = wrapIdx(vi, P.NX, 1); if (!P.oneD) vj = wrapIdx(vj, P.NY, 1); }
      return (size_t)(vj + YO) * pitch + (vi + XO);
   };
   // periodic aliases of a vertex that was just modified (keeps the halo copies current)
   auto storeBt = [&](int vi, int vj, double val) {
      bt3[vix(vi, vj)] = val;
      if (P.periodic) {
         int bi = wrapIdx(vi, P.NX, 1), bj = P.oneD ? 0 : wrapIdx(vj, P.NY, 1);
         for (int oj = -1; oj <= 1; oj++)
            for (int oi = -1; oi <= 1; oi++) {
               if (P.oneD && oj != 0) continue;
               int ai = bi + oi * P.NX, aj = bj + oj * P.NY;
               if (ai < -2 || ai > P.NX + 2 || aj < (P.oneD ? 0 : -2) || aj > (P.oneD ? 0 : P.NY + 2)) continue;
               bt3[(size_t)(aj + YO) * pitch + (ai + XO)] = val;
            }
      }
   };
   size_t g = (size_t)(j + YO) * pitch + (i + XO);
   double b0c, bt0c, bx0, by0, bt3c, bx3, by3;
   centreTopoGlobal(P, A.b0v, A.bt0, i, j, b0c, bt0c, bx0, by0);
   centreTopoGlobal(P, A.b0v, bt3r, i, j, b0c, bt3c, bx3, by3);
   double gamold = gamma2(P, bx0, by0);
   double Hn_old = computeHn(A.w0[g], b0c, bt0c, gamold);
   double corr;
   excessDeposition(P, A.hpsi0[g], Hn_old, gamold, bt3c - bt0c, corr);
   if (!(corr > EPS)) return true;
   // depositional vertices of the cell (Redistribute.f90:276-302)
   double b_diff[4] = {0, 0, 0, 0}, sum_b_diff = 0.0;
   int dvi[4], dvj[4], N = 0;
   const int oi[4] = {0, 1, 0, 1}, oj[4] = {0, 0, 1, 1};
   for (int k = 0; k < (P.oneD ? 2 : 4); k++) {
      size_t v = vix(i + oi[k], j + oj[k]);
      b_diff[N] = bt3[v] - A.bt0[v];
      if (b_diff[N] > 0.0) { dvi[N] = i + oi[k]; dvj[N] = j + oj[k]; sum_b_diff = sum_b_diff + b_diff[N]; N++; }
   }
   if (N == 0) return false;
   double delta = 4.0 * corr / sum_b_diff;
   double Hnold = Hn_old < 0.0 ? 0.0 : Hn_old;  // intermed0%u(iHn)
   double db = bt3c - bt0c;
   double tol = EPS * w3[g] * 10.0;
   double adjustment = 0.0;
   double discrepancy = kahan3(Hnold * gamold, -db, corr);
   if (fabs(discrepancy) < tol) {
      adjustment = kahan3(tol, -Hnold * gamold, db);
      adjustment = adjustment * (4.0 / sum_b_diff);
      adjustment = adjustment - delta;
      adjustment = fmax(adjustment, 0.0);
   }
   double Hg = A.hpsi0[g] * gamold / (1.0 - P.BedPorosity);
   tol = EPS * 10.0;
   discrepancy = kahan3(Hg, -db, corr);
   if (fabs(discrepancy) < tol) {
      double adj = kahan3(tol, -Hg, db);
      adj = adj * (4.0 / sum_b_diff);
      adj = adj - delta;
      adjustment = fmax(adjustment, adj);
   }
   delta = delta + adjustment;
   for (int k = 0; k < N; k++) {
      size_t v = vix(dvi[k], dvj[k]);
      storeBt(dvi[k], dvj[k], bt3[v] - delta * b_diff[k]);
   }
   // refresh the surrounding 3^D cells (Redistribute.f90:404-472)
   for (int ci = i - 1; ci <= i + 1; ci++)
      for (int cj = (P.oneD ? 0 : j - 1); cj <= (P.oneD ? 0 : j + 1); cj++) {
         if (!P.periodic && (ci < 0 || ci >= P.NX || cj < 0 || cj >= P.NY)) continue;
         int wi = wrapIdx(ci, P.NX, P.periodic), wj = P.oneD ? 0 : wrapIdx(cj, P.NY, P.periodic);
         if (!cellTileActive(P, A.tileMask, A.allActive, wi, wj)) continue;
         size_t gc = (size_t)(wj + YO) * pitch + (wi + XO);
         double c_b0, c_bt0, c_bx0, c_by0, c_bt3, c_bx3, c_by3;
         centreTopoGlobal(P, A.b0v, A.bt0, wi, wj, c_b0, c_bt0, c_bx0, c_by0);
         centreTopoGlobal(P, A.b0v, bt3r, wi, wj, c_b0, c_bt3, c_bx3, c_by3);
         double dbc = c_bt3 - c_bt0;
         double go = gamma2(P, c_bx0, c_by0), gn = gamma2(P, c_bx3, c_by3);
         double Ho = computeHn(A.w0[gc], c_b0, c_bt0, go);
         if (Ho < 0.0) Ho = 0.0;
         double w;
         if (!P.oneD) {
            w = c_bt3;
            w = w + (Ho * go / gn - dbc / gn) / gn;
            w = w + c_b0;
         } else {
            w = -dbc / gn / gn;
            w = w + Ho * go / gn / gn;
            w = w + c_bt3;
            w = w + c_b0;
         }
         w3[gc] = w;
         hpsi3[gc] = A.hpsi0[gc] * go / gn - (1.0 - P.BedPorosity) * dbc / gn;
      }
   return true;
}

// RedistributeGrid (Redistribute.f90:203-247) as the reference runs it: one thread walks the sorted list.
// Kept as the yardstick of the wave kernel below (kgpu_debug_sequential_walk).
__global__ void morpho_redistribute_kernel(const DevParams P, const RedistArgs A) {
   if (threadIdx.x != 0 || blockIdx.x != 0) return;
   for (int e = 0; e < A.n; e++)
      if (!redistributeCell(P, A, A.list[e].i, A.list[e].j)) { A.ctrl->refineMorpho = 1; return; }
}

// The same walk, in parallel, with the reference's order preserved where it matters.  Entry e (its position
// in the sorted list) reads the bed at the vertices (i-1 .. i+2, j-1 .. j+2) and writes the vertices of its
// own cell and the 3 x 3 cells around it, so two entries commute exactly when their cells are more than two
// cells apart (Chebyshev distance, through the periodic wrap).  rankMap holds, per listed cell, its list
// position (INT_MAX elsewhere and once the entry is done): an entry runs as soon as no cell within distance 2
// carries a smaller rank.  Thread = list position; blocks take a ticket, so every entry a thread can wait
// for belongs to a block that has already started -- no deadlock -- and the lowest unfinished entry is always
// runnable.  Results are those of the sequential walk bit for bit (tests/test_gpu_parity.py).
struct RedistWaveArgs {
   RedistArgs R;
   int *rankMap;      // one int per padded cell position, INT_MAX outside a walk
   int *ticket;       // block ticket counter, zero before the launch
};
__global__ void redist_rank_kernel(const DevParams P, const RedistEntry *list, int n, int *rankMap, int *ticket) {
   int e = blockIdx.x * blockDim.x + threadIdx.x;
   if (e == 0) *ticket = 0;
   if (e < n) rankMap[(size_t)(list[e].j + YO) * P.pitch + (list[e].i + XO)] = e;
}
__global__ void __launch_bounds__(128) morpho_redistribute_wave_kernel(const DevParams P, const RedistWaveArgs W) {
   __shared__ int s_block;
   if (threadIdx.x == 0) s_block = atomicAdd(W.ticket, 1);
   __syncthreads();
   const int e = s_block * blockDim.x + threadIdx.x;
   if (e >= W.R.n) return;
   const int i = W.R.list[e].i, j = W.R.list[e].j;
   volatile int *rank = W.rankMap;
   volatile int *abortFlag = &W.R.ctrl->refineMorpho;
   const size_t g = (size_t)(j + YO) * P.pitch + (i + XO);
   const int INF = 0x7f7f7f7f;   // the value the map is memset to
   bool done = false;
   while (!done) {
      if (*abortFlag) break;
      bool ready = true;
      for (int dj = (P.oneD ? 0 : -2); dj <= (P.oneD ? 0 : 2); dj++)
         for (int di = -2; di <= 2; di++) {
            int ni = i + di, nj = j + dj;
            if (P.periodic) { ni = wrapIdx(ni, P.NX, 1); if (!P.oneD) nj = wrapIdx(nj, P.NY, 1); }
            else if (ni < 0 || ni >= P.NX || nj < 0 || nj >= P.NY) continue;
            if (ni == i && nj == j) continue;
            if (rank[(size_t)(nj + YO) * P.pitch + (ni + XO)] < e) ready = false;
         }
      if (ready) {
         __threadfence();   // the data of the finished neighbours is visible before it is read
         if (!redistributeCell(P, W.R, i, j)) *abortFlag = 1;
         done = true;
      }
   }
   __threadfence();         // this entry's corrections are visible before its rank is released
   rank[g] = INF;
}

// ------------------------------------------------------------------ redistribution across ranks
// RedistributeGrid is one sequential walk over a list sorted over the WHOLE domain (Redistribute.f90:203-247),
// and every correction changes the bed its neighbours' corrections start from, so a decomposed run cannot
// process "its" entries independently.  Instead every rank replays the whole global list on a sparse copy of
// just the data the walk can touch: per listed cell the 4 x 4 vertices and 3 x 3 cells around it ("patch"),
// packed by the owning rank, all-gathered, and addressed through slot tables the host builds from the cell
// indices alone (vertices / cells shared by several patches resolve to ONE canonical slot, periodic images
// included).  All ranks do identical arithmetic in identical order, then each scatters the canonical values
// that fall into its block.  The arithmetic is that of morpho_redistribute_kernel, statement for statement.
constexpr int RP_V = 16, RP_C = 9;                    // vertices / cells per patch
constexpr int RP_B0 = 0, RP_BT0 = 16, RP_BT3 = 32;    // vertex fields
constexpr int RP_W0 = 48, RP_HPSI0 = 57, RP_W3 = 66, RP_HPSI3 = 75;   // cell fields
constexpr int RP_ACT = 84;                                             // 1.0 where the cell's tile is active (dynamic tiles), else 0.0
constexpr int RP_DOUBLES = 93;
static_assert(RP_DOUBLES == RT_DOUBLES && RP_W0 == RT_W0 && RP_V == RT_V && RP_C == RT_C, "patch layout of kgpu_redist_tables.hpp");

struct RedistPackArgs {
   const double *b0v, *bt0, *bt3, *w0, *hpsi0, *w3, *hpsi3;
   const RedistEntry *list;   // local entries, local indices
   int n;
   const uint8_t *tileMask;
   int allActive;
};
// one thread per (entry, patch element)
__global__ void redist_pack_kernel(const DevParams P, const RedistPackArgs A, double *out) {
   int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k >= A.n * RP_DOUBLES) return;
   int e = k / RP_DOUBLES, o = k % RP_DOUBLES;
   int i = A.list[e].i, j = A.list[e].j;
   const double *src;
   int a, b;
   if (o < RP_W0) {
      src = o < RP_BT0 ? A.b0v : (o < RP_BT3 ? A.bt0 : A.bt3);
      int q = o % RP_V; a = q % 4 - 1; b = q / 4 - 1;
   } else {
      int f = (o - RP_W0) / RP_C;
      src = f == 0 ? A.w0 : f == 1 ? A.hpsi0 : f == 2 ? A.w3 : A.hpsi3;
      int q = (o - RP_W0) % RP_C; a = q % 3 - 1; b = q / 3 - 1;
   }
   int jj = P.oneD ? 0 : j + b;
   if (o >= RP_ACT) {   // RedistributeGrid refreshes only cells of active tiles (Redistribute.f90:404-472; the ring of the mask
      out[k] = cellTileActive(P, A.tileMask, A.allActive, i + a, jj) ? 1.0 : 0.0;   // knows the tiles of the ranks next door)
      return;
   }
   out[k] = src[(size_t)(jj + YO) * P.pitch + (i + a + XO)];
}

__device__ __forceinline__ void centreTopoVals(const DevParams &P, double a, double b, double c, double d, double ta, double tb, double tc,
                                               double td, double &b0c, double &btc, double &bx, double &by) {
   // centreTopoGlobal (kgpu_tiles.cuh) on values: a = (i,j), b = (i+1,j), c = (i,j+1), d = (i+1,j+1)
   if (!P.oneD) {
      b0c = 0.25 * kahan4(a, b, c, d);
      btc = 0.25 * kahan4(ta, tb, tc, td);
      bx = 0.5 * P.dxR * kahan8(b, tb, -a, -ta, d, td, -c, -tc);
      by = 0.5 * P.dyR * kahan8(c, tc, -a, -ta, d, td, -b, -tb);
   } else {
      b0c = 0.5 * (a + b);
      btc = 0.5 * (ta + tb);
      bx = P.dxR * kahan4(b, tb, -a, -ta);
      by = 0.0;
   }
}

struct RedistGlobalArgs {
   double *G;            // gathered patches, canonical values live at the slot offsets below
   const int *vslot;     // [n][16] offset of the b0 value of vertex (i-1+a, j-1+b), a + 4 b; bt0 at +16, bt3 at +32
   const int *cslot;     // [n][9]  offset of the w0 value of cell (i-1+a, j-1+b), a + 3 b; hpsi0 +9, w3 +18, hpsi3 +27, active +36
   int n;
   Ctrl *ctrl;
};
__global__ void redist_global_kernel(const DevParams P, const RedistGlobalArgs A) {
   if (threadIdx.x != 0 || blockIdx.x != 0) return;
   const double EPS = 2.220446049250313e-16;
   double *G = A.G;
   const int rowV = P.oneD ? 0 : 4, rowC = P.oneD ? 0 : 3;   // 1-D: every row of the patch is row 0
   for (int e = 0; e < A.n; e++) {
      const int *vs = A.vslot + (size_t)e * RP_V, *cs = A.cslot + (size_t)e * RP_C;
      // the patch rows are j-1, j, j+1(, j+2): the cell itself sits at a = 1, b = 1
      auto V = [&](int a, int b) -> int { return vs[a + (P.oneD ? 1 : b) * 4]; };
      auto C = [&](int a, int b) -> int { return cs[a + (P.oneD ? 1 : b) * 3]; };
      (void)rowV; (void)rowC;
      auto centre = [&](int a, int b, int fieldOff, double &b0c, double &btc, double &bx, double &by) {
         int v00 = V(a, b), v10 = V(a + 1, b), v01 = P.oneD ? v00 : V(a, b + 1), v11 = P.oneD ? v10 : V(a + 1, b + 1);
         centreTopoVals(P, G[v00], G[v10], G[v01], G[v11], G[v00 + fieldOff], G[v10 + fieldOff], G[v01 + fieldOff], G[v11 + fieldOff],
                        b0c, btc, bx, by);
      };
      const int c = C(1, 1);
      double b0c, bt0c, bx0, by0, bt3c, bx3, by3;
      centre(1, 1, RP_BT0, b0c, bt0c, bx0, by0);
      centre(1, 1, RP_BT3, b0c, bt3c, bx3, by3);
      double gamold = gamma2(P, bx0, by0);
      double Hn_old = computeHn(G[c], b0c, bt0c, gamold);
      double corr;
      excessDeposition(P, G[c + 9], Hn_old, gamold, bt3c - bt0c, corr);
      if (!(corr > EPS)) continue;
      double b_diff[4] = {0, 0, 0, 0}, sum_b_diff = 0.0;
      int dv[4], N = 0;
      const int oi[4] = {0, 1, 0, 1}, oj[4] = {0, 0, 1, 1};
      for (int k = 0; k < (P.oneD ? 2 : 4); k++) {
         int v = V(1 + oi[k], 1 + oj[k]);
         b_diff[N] = G[v + RP_BT3] - G[v + RP_BT0];
         if (b_diff[N] > 0.0) { dv[N] = v; sum_b_diff = sum_b_diff + b_diff[N]; N++; }
      }
      if (N == 0) { A.ctrl->refineMorpho = 1; return; }
      double delta = 4.0 * corr / sum_b_diff;
      double Hnold = Hn_old < 0.0 ? 0.0 : Hn_old;
      double db = bt3c - bt0c;
      double tol = EPS * G[c + 18] * 10.0;
      double adjustment = 0.0;
      double discrepancy = kahan3(Hnold * gamold, -db, corr);
      if (fabs(discrepancy) < tol) {
         adjustment = kahan3(tol, -Hnold * gamold, db);
         adjustment = adjustment * (4.0 / sum_b_diff);
         adjustment = adjustment - delta;
         adjustment = fmax(adjustment, 0.0);
      }
      double Hg = G[c + 9] * gamold / (1.0 - P.BedPorosity);
      tol = EPS * 10.0;
      discrepancy = kahan3(Hg, -db, corr);
      if (fabs(discrepancy) < tol) {
         double adj = kahan3(tol, -Hg, db);
         adj = adj * (4.0 / sum_b_diff);
         adj = adj - delta;
         adjustment = fmax(adjustment, adj);
      }
      delta = delta + adjustment;
      for (int k = 0; k < N; k++) G[dv[k] + RP_BT3] = G[dv[k] + RP_BT3] - delta * b_diff[k];
      // refresh the surrounding 3^D cells (Redistribute.f90:404-472)
      for (int a = 0; a < 3; a++)
         for (int b = (P.oneD ? 1 : 0); b < (P.oneD ? 2 : 3); b++) {
            const int cc = C(a, b);
            if (G[cc + 36] == 0.0) continue;   // cell of a ghost or inactive tile: left alone
            double c_b0, c_bt0, c_bx0, c_by0, c_bt3, c_bx3, c_by3;
            centre(a, b, RP_BT0, c_b0, c_bt0, c_bx0, c_by0);
            centre(a, b, RP_BT3, c_b0, c_bt3, c_bx3, c_by3);
            double dbc = c_bt3 - c_bt0;
            double go = gamma2(P, c_bx0, c_by0), gn = gamma2(P, c_bx3, c_by3);
            double Ho = computeHn(G[cc], c_b0, c_bt0, go);
            if (Ho < 0.0) Ho = 0.0;
            double w;
            if (!P.oneD) {
               w = c_bt3;
               w = w + (Ho * go / gn - dbc / gn) / gn;
               w = w + c_b0;
            } else {
               w = -dbc / gn / gn;
               w = w + Ho * go / gn / gn;
               w = w + c_bt3;
               w = w + c_b0;
            }
            G[cc + 18] = w;
            G[cc + 27] = G[cc + 9] * go / gn - (1.0 - P.BedPorosity) * dbc / gn;
         }
   }
}

// canonical values back into the local planes: key = global index, images = every local position it maps to
struct RedistScatterArgs {
   const double *G;
   const int *vkey, *vbase;   // unique vertices: (gi, gj) pairs, slot offset
   const int *ckey, *cbase;   // unique cells
   int nv, nc;
   double *bt3, *w3, *hpsi3;
   int gx0, gy0, NXg, NYg;    // origin of the local block in global cells, global extent
};
__global__ void redist_scatter_kernel(const DevParams P, const RedistScatterArgs A) {
   int k = blockIdx.x * blockDim.x + threadIdx.x;
   const bool isV = k < A.nv;
   if (!isV) { k -= A.nv; if (k >= A.nc) return; }
   const int gi = isV ? A.vkey[2 * k] : A.ckey[2 * k], gj = isV ? A.vkey[2 * k + 1] : A.ckey[2 * k + 1];
   const int base = isV ? A.vbase[k] : A.cbase[k];
   const int hiX = isV ? P.NX + 2 : P.NX + 1, hiY = P.oneD ? 0 : (isV ? P.NY + 2 : P.NY + 1);
   for (int oy = -1; oy <= 1; oy++)
      for (int ox = -1; ox <= 1; ox++) {
         if (P.oneD && oy != 0) continue;
         int li = gi - A.gx0 + ox * A.NXg, lj = P.oneD ? 0 : gj - A.gy0 + oy * A.NYg;
         if (li < -2 || li > hiX || lj < (P.oneD ? 0 : -2) || lj > hiY) continue;
         size_t g = (size_t)(lj + YO) * P.pitch + (li + XO);
         if (isV) A.bt3[g] = A.G[base + RP_BT3];
         else { A.w3[g] = A.G[base + 18]; A.hpsi3[g] = A.G[base + 27]; }
      }
}

// ------------------------------------------------------------------ one Runge-Kutta stage of M in ONE launch
// morpho_emd_kernel + morpho_bed_kernel + morpho_cell_kernel fused (single device): E - D is evaluated over the tile
// plus one cell (the four cells around every vertex of the tile) into shared memory, the stage bed at the tile's
// (BX+1) x (BY+1) vertices from it, and the linear update of w and Hn psi of the tile's cells from that bed -- the
// E - D plane and the second read of the new bed never touch HBM, and two of the three halo fills between the
// kernels disappear.  Every statement is the one of the kernel it comes from (same functions, same operand order),
// so the results are those of the three-kernel path bit for bit; decomposed runs keep that path (they exchange E - D
// and the bed between the kernels), and tests/test_gpu_multi.py requires both to agree.
// What must be valid around the tile: w, Hn psi, b0c, cBt, cBx, cBy, U, V one cell out and cHn two cells out (the
// periodic images are filled by the host after every stage; out-of-domain and inactive cells contribute E - D = 0 as
// the never-written E - D plane did), the vertex arrays b0v, bt0, btk two vertices out.
// Tile: 32 x 13 cells for the 2-D kernel -- (32+2) x (13+2) = 510 E - D evaluations fill two passes of 256 threads
// (32 x 14 would need a third pass for 32 leftover cells of the most expensive phase: measured +30 % on the kernel).
constexpr int MORPHO_STAGE_BY = 13;
// EMDPLANE = true: E - D comes from the plane morpho_emd_kernel has just written (one evaluation per cell: the closures
// -- pow, tanh, log -- are compute-bound and the 23 % of extra evaluations in the ring around the tile cost more than
// the plane's 16 B per cell: measured); false: evaluated here over tile + 1 cell.
template <int BX, int BY, bool ONED, bool EMDPLANE>
__global__ void __launch_bounds__(256, 4) morpho_stage_kernel(const DevParams P, const MorphoArgs A, const int2 *blocks) {
   constexpr int EX = BX + 2, EY = ONED ? 1 : BY + 2;   // E - D: tile + 1 cell
   constexpr int VX = BX + 1, VY = ONED ? 1 : BY + 1;   // vertices of the tile
   __shared__ double s_emd[EX * EY];
   __shared__ double s_bt[VX * VY];
   const int2 bo = blocks[blockIdx.x];
   const int x0 = bo.x * BX, y0 = ONED ? 0 : bo.y * BY;
   const int tid = threadIdx.x;
   const bool wrapOrHalo = P.periodic || P.haloValid;
   const double eps = P.Hneps;
   // ---- E - D (morpho_emd_kernel; MorphodynamicRHS.f90:96-145)
   for (int k = tid; k < EX * EY; k += blockDim.x) {
      const int ci = x0 - 1 + k % EX, cj = ONED ? 0 : y0 - 1 + k / EX;
      double val = 0.0;
      const bool inDomain = ci >= 0 && ci < P.NX && cj >= 0 && cj < P.NY;
      const bool image = wrapOrHalo && ci >= -1 && ci <= P.NX && cj >= (ONED ? 0 : -1) && cj <= (ONED ? 0 : P.NY);
      if (EMDPLANE) {   // morpho_bed_kernel's read of the plane: wrapped index on a periodic device, zero outside the domain
         if (inDomain || image) {
            int i = ci, j = cj;
            i = wrapIdx(i, P.NX, P.periodic); j = wrapIdx(j, P.NY, P.periodic);
            val = A.EmD[(size_t)(j + YO) * P.pitch + (i + XO)];
         }
      } else if ((inDomain || image) && cellTileActive(P, A.tileMask, A.allActive, ci, cj)) {
         const size_t g = (size_t)(cj + YO) * P.pitch + (ci + XO);
         CellState q;
         q.w = A.w[g]; q.hpsi = A.hpsi[g]; q.hu = 0.0; q.hv = 0.0;
         q.b0 = A.b0c[g]; q.bt = A.cBt[g]; q.bx = A.cBx[g]; q.by = A.cBy[g];
         desingularise(P, q, false);
         q.u = A.U[g]; q.v = A.V[g];
         auto nbHn = [&](int i, int j) -> double {
            if (cellTileActive(P, A.tileMask, A.allActive, i, j)) return A.cHn[(size_t)(j + YO) * P.pitch + (i + XO)];
            return storedHn(P, A.w, A.b0v, A.btk, i, j);
         };
         bool dry = q.Hn < eps || nbHn(ci - 1, cj) < eps || nbHn(ci + 1, cj) < eps;
         if (!P.oneD) dry = dry || nbHn(ci, cj - 1) < eps || nbHn(ci, cj + 1) < eps;
         val = dry ? 0.0 : erosionMinusDeposition(P, q);
      }
      s_emd[k] = val;
   }
   __syncthreads();
   // ---- the stage bed at the tile's vertices (morpho_bed_kernel; MorphodynamicRHS.f90:154-305, TimeStepper.f90:574-581)
   const int nvy = P.oneD ? 1 : P.NY + 1;
   for (int k = tid; k < VX * VY; k += blockDim.x) {
      const int lvx = k % VX, lvy = k / VX;
      const int vi = x0 + lvx, vj = ONED ? 0 : y0 + lvy;
      double val = 0.0;
      if (vi <= P.NX && vj < nvy) {
         const size_t gv = (size_t)(vj + YO) * P.pitch + (vi + XO);
         bool any = false;
         for (int dj = (P.oneD ? 0 : -1); dj <= 0; dj++)
            for (int di = -1; di <= 0; di++) {
               int i = vi + di, j = vj + dj;
               if (!wrapOrHalo && (i < 0 || i >= P.NX || j < 0 || j >= P.NY)) continue;
               if (cellTileActive(P, A.tileMask, A.allActive, i, j)) any = true;
            }
         if (!any) val = A.btn[gv];   // not a vertex of an active tile: the array keeps what it holds
         else {
            const double psib = 1.0 - P.BedPorosity;
            double rhs;
            // E - D of the cell (vi + di, vj + dj), di, dj in {-1, 0}: tile-local (lvx + di + 1, lvy + dj + 1)
            auto emd = [&](int di, int dj) -> double { return s_emd[(ONED ? 0 : lvy + dj + 1) * EX + lvx + di + 1]; };
            auto slopes = [&](int i, int j, double &bx_, double &by_) {
               if (cellTileActive(P, A.tileMask, A.allActive, i, j)) {
                  size_t gc = (size_t)(j + YO) * P.pitch + (i + XO);
                  bx_ = A.cBx[gc]; by_ = A.cBy[gc];
               } else {
                  double b0c_, btc_;
                  centreTopoGlobal(P, A.b0v, A.btk, i, j, b0c_, btc_, bx_, by_);
               }
            };
            if (!P.oneD) {
               double bx[4], by[4];
               slopes(vi - 1, vj - 1, bx[0], by[0]);
               slopes(vi - 1, vj, bx[1], by[1]);
               slopes(vi, vj - 1, bx[2], by[2]);
               slopes(vi, vj, bx[3], by[3]);
               double dbdx = 0.25 * kahan4(bx[0], bx[1], bx[2], bx[3]);
               double dbdy = 0.25 * kahan4(by[0], by[1], by[2], by[3]);
               double gam = gamma2(P, dbdx, dbdy);
               rhs = -0.25 * gam / psib * kahan4(emd(-1, -1), emd(-1, 0), emd(0, -1), emd(0, 0));
            } else {
               bool lAct = (wrapOrHalo || vi - 1 >= 0) && cellTileActive(P, A.tileMask, A.allActive, vi - 1, 0);
               bool rAct = (wrapOrHalo || vi < P.NX) && cellTileActive(P, A.tileMask, A.allActive, vi, 0);
               double bxl = 0.0, bxr = 0.0, byd;
               if (lAct) slopes(vi - 1, 0, bxl, byd);
               if (rAct) slopes(vi, 0, bxr, byd);
               if (lAct && rAct) {
                  double dbdx = 0.5 * (bxl + bxr);
                  double gam = gamma2(P, dbdx, 0.0);
                  rhs = -0.5 * gam * (emd(-1, 0) + emd(0, 0)) / psib;
               } else {
                  double dbdx = 0.5 * (lAct ? bxl : bxr);
                  double gam = gamma2(P, dbdx, 0.0);
                  rhs = -0.5 * gam * (lAct ? emd(-1, 0) : emd(0, 0)) / psib;
               }
            }
            if (A.a0 == 0.0) val = A.bt0[gv] + A.dtMorpho * rhs;
            else val = A.a0 * A.bt0[gv] + A.a1 * (A.btk[gv] + A.dtMorpho * rhs);
            val = fmax(-P.EroDepth, val);
            // (neighbouring tiles write the vertices they share with this one, with the same bits; the periodic aliases
            // NX, NY are left to the halo fill, as in morpho_bed_kernel)
            if (!(P.periodic && (vi == P.NX || (!P.oneD && vj == P.NY)))) A.btn[gv] = val;
         }
      }
      s_bt[k] = val;
   }
   __syncthreads();
   // ---- w, Hn psi and the centre planes of the new bed (morpho_cell_kernel; TimeStepper.f90:587-610)
   for (int k = tid; k < BX * BY; k += blockDim.x) {
      const int tx = k % BX, ty = ONED ? 0 : k / BX;
      const int ci = x0 + tx, cj = ONED ? 0 : y0 + ty;
      if (ci >= P.NX || cj >= P.NY) continue;
      if (!cellTileActive(P, A.tileMask, A.allActive, ci, cj)) continue;
      const size_t g = (size_t)(cj + YO) * P.pitch + (ci + XO);
      const double va = A.b0v[g], vb = A.b0v[g + 1], vc = ONED ? 0.0 : A.b0v[g + P.pitch], vd = ONED ? 0.0 : A.b0v[g + P.pitch + 1];
      const double ta = s_bt[ty * VX + tx], tb = s_bt[ty * VX + tx + 1];
      const double tc = ONED ? 0.0 : s_bt[(ty + 1) * VX + tx], td = ONED ? 0.0 : s_bt[(ty + 1) * VX + tx + 1];
      double b0c, btnc, bxn, byn;
      centreTopoVals(P, va, vb, vc, vd, ta, tb, tc, td, b0c, btnc, bxn, byn);
      const double bt0c = A.zBt[g];
      double gamold = A.zGam[g], gamnew = gamma2(P, bxn, byn);
      double Hn_old = computeHn(A.w0[g], b0c, bt0c, gamold);
      double db = btnc - bt0c;
      double w = -db / gamnew / gamnew;
      w = w + btnc;
      w = w + Hn_old * gamold / gamnew / gamnew;
      w = w + b0c;
      A.wn[g] = w;
      double Hnpsi = -(1.0 - P.BedPorosity) * db / gamnew;
      Hnpsi = Hnpsi + A.hpsi0[g] * gamold / gamnew;
      A.hpsin[g] = Hnpsi;
      A.nBt[g] = btnc; A.nBx[g] = bxn; A.nBy[g] = byn;
      double Hn_new = computeHn(w, b0c, btnc, gamnew);
      A.nHn[g] = Hn_new < 0.0 ? 0.0 : Hn_new;
   }
}

}  // namespace kgpu
