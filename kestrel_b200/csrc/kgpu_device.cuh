// kgpu_device.cuh -- device-side data model and pointwise physics of libkestrel_gpu.
//
// B200-native layout (nothing here mirrors the reference's tile-of-arrays design):
// one flat struct-of-arrays fp64 grid per field over the whole (sub)domain with a
// halo, so that a warp reads 32 consecutive doubles of one field (256 B, two full
// sectors quads) and the fused stage kernel can stage a tile + 2-cell halo into
// shared memory.  Cell (i,j), i in [-2,NX+2), lives at (j+YO)*pitch + (i+XO);
// XO = 16 doubles keeps column 0 of every row 128-byte aligned.  Vertex (i,j)
// (the lower-left corner of cell (i,j)) uses the same indexing.
//
// Pointwise formulas follow the reference operation by operation (cited inline);
// compiled with -fmad=false in the faithful variant so every +,-,*,/,sqrt rounds
// exactly as the CPU reference's IEEE arithmetic does.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/kestrel_gpu.h"

#ifndef KGPU_FAST
#define KGPU_FAST 0
#endif

namespace kgpu {

constexpr int XO = 16;   // x offset of cell 0 in a padded row (doubles)
constexpr int YO = 2;    // y offset of row 0
constexpr int HALO = 2;  // cells

// variable slots of the primary state
enum { QW = 0, QHU = 1, QHV = 2, QHPSI = 3 };

// One flux source (type Sources, RunSettings.f90:101-109).  The time series of ALL sources live in one pool
// (any number of sources, any series length, as in the reference's allocatable arrays): source J owns
// pool[off .. off + 3 n): n times, n fluxes, n solids fractions.
struct DevSource {
   double x, y, radius;
   int numCells, n;
   long long off;
};

// Everything a kernel needs that is constant over a run.
struct DevParams {
   int NX, NY;            // cells owned by this device
   int nX, nY;            // cells per reference tile
   int nXt, nYt;          // reference tiles owned by this device
   int gtx0, gty0;        // global tile offset of this device's block (multi-GPU)
   int gnXt, gnYt;        // global tile counts
   int pitch;             // doubles per padded row
   int rows;              // padded rows
   int oneD, periodic, geom, morpho;
   int haloValid;   // decomposed periodic run: cells / vertices beyond the local block are images owned by another rank (or the local wrap)
   int limiter, drag, erosion, deposition, eroTrans, damp, fswitch;
   int nSources;
   double dx, dy, dxR, dyR, xSize, ySize;
   double g, rhow, rhos, gred;
   double ChezyCo, ManningCo, CoulombCo, PoulMin, PoulMax, PoulInt, PoulBeta;
   double EdBetastar, EdKappa, EdGamma, SwitchRate, SwitchValue;
   double EroRate, EroRateGranular, CriticalShields, EroDepth, EroCritH, BedPorosity, maxPack, SolidDiameter, ws0, nsettling, nu;
   double Hneps, cfl, diffusiveTimeScale, maxdt;
   int bcDirichlet;      // ghost cells of domain-edge tiles carry the boundary values (UpdateTiles.f90:571-664)
   double bcU, bcV, bcPsi;
   double mm2HalfTheta;  // 0.5 * 1.3: MinMod2 half-slope factor of the contracted variant (constant-bank operand)
   double halfGRhow;     // 0.5 * g * rhow: hydrostatic flux factor of pure-water tiles in the contracted variant
};

// Device-resident control block: dt selection and rollback flags never leave the GPU
// inside a step (TimeStepper.f90:155-169, 393-404, 452-463).
struct Ctrl {
   unsigned long long cflBits[4];  // per-stage min unit-CFL step, as ordered bits of a positive double
   double t;                       // grid%t of the current H operator
   double dt;                      // dt_hydro in use
   double dtAdvised;               // advisedTimeStep of the latest substep-1 RHS
   double dtNew;                   // 0.9*dt_k proposed by a failed stage
   int failed;                     // stage whose dt_k < dt (0 = none)
   int nonfinite;
   int refineMorpho;               // morphodynamic refine flag
   int nRedist;                    // length of the redistribution list
   int gRefine;                    // decomposed runs: max over ranks of refineMorpho ...
   int gRedistMax;                 // ... and of nRedist (the local values above stay intact)
};

__device__ __forceinline__ int cidx(const DevParams &P, int i, int j) { return (j + YO) * P.pitch + (i + XO); }

// ---------------------------------------------------------------- small helpers
// utilities.f90:439-448
__device__ __forceinline__ void kahanAdd(double x, double &s, double &c) {
   double y = x - c;
   double t = s + y;
   c = (t - s) - y;
   s = t;
}
__device__ __forceinline__ double kahan3(double a, double b, double c3) {
   double s = 0.0, c = 0.0;
   kahanAdd(a, s, c); kahanAdd(b, s, c); kahanAdd(c3, s, c);
   return s;
}
__device__ __forceinline__ double kahan4(double a, double b, double c4, double d) {
   double s = 0.0, c = 0.0;
   kahanAdd(a, s, c); kahanAdd(b, s, c); kahanAdd(c4, s, c); kahanAdd(d, s, c);
   return s;
}
__device__ __forceinline__ double kahan8(double a, double b, double c8, double d, double e, double f, double g, double h) {
   double s = 0.0, c = 0.0;
   kahanAdd(a, s, c); kahanAdd(b, s, c); kahanAdd(c8, s, c); kahanAdd(d, s, c);
   kahanAdd(e, s, c); kahanAdd(f, s, c); kahanAdd(g, s, c); kahanAdd(h, s, c);
   return s;
}

// Closures.f90:269-305
__device__ __forceinline__ double gamma2(const DevParams &P, double bx, double by) {
   return P.geom ? sqrt(1.0 + bx * bx + by * by) : 1.0;
}
// Closures.f90:158-171
__device__ __forceinline__ double computeHn(double w, double b0, double bt, double gam) {
   double Hn = -bt;
   Hn = Hn + (w - b0);
   Hn = Hn * gam;
   return Hn;
}
// HydraulicRHS.f90:741-753
__device__ __forceinline__ double interpolateB(double b0_0, double b0_1, double bt_0, double bt_1) {
   double b = bt_0 + bt_1;
   b = b + b0_0;
   b = b + b0_1;
   b = b * 0.5;
   return b;
}

// Limiters.f90:83-184 -- runtime-selected device-function variant
__device__ __forceinline__ double limiter(const DevParams &P, double a, double b) {
   switch (P.limiter) {
      case KGPU_LIM_MINMOD1:
         if (a * b <= 0.0) return 0.0;
         return a > 0.0 ? fmin(a, b) : fmax(a, b);
      default:
      case KGPU_LIM_MINMOD2: {
         const double theta = 1.3;
         if (a * b <= 0.0) return 0.0;
         if (a > 0.0) return fmin(theta * a, fmin(theta * b, 0.5 * (a + b)));
         return fmax(theta * a, fmax(theta * b, 0.5 * (a + b)));
      }
      case KGPU_LIM_NONE: return 0.5 * (a + b);
      case KGPU_LIM_VANALBADA: {
         double den = a * a + b * b;
         if (den == 0.0) return 0.0;
         return (a * a * b + a * b * b) / den;
      }
      case KGPU_LIM_WENO: {
         const double eps = 1.0e-6;
         double ea = a * a + eps, eb = b * b + eps;
         double wa = 1.0 / (ea * ea), wb = 1.0 / (eb * eb);
         return (wa * a + wb * b) / (wa + wb);
      }
   }
}

// Closures.f90:178-205
__device__ __forceinline__ double speed2(const DevParams &P, double u, double v, double bx, double by) {
   double m = P.geom ? u * u * (1.0 + bx * bx) : u * u;
   if (P.oneD) return m;
   if (P.geom) m = m + v * v * (1.0 + by * by) + 2.0 * bx * by * u * v;
   else m = m + v * v;
   return m;
}

// cell-centred state handed to the closures (the 13-vector of the reference)
struct CellState {
   double w, hu, hv, hpsi, Hn, u, v, psi, rho, b0, bt, bx, by;
};

// Closures.f90:797-915
__device__ inline double fswitch(const DevParams &P, double psi) {
   const double rate = P.SwitchRate, val = P.SwitchValue;
   const double PI = 3.141592653589793238462643383279502884;
   switch (P.fswitch) {
      default:
      case KGPU_SWITCH_TANH: return 0.5 * (1.0 + tanh(rate * (psi - val)));
      case KGPU_SWITCH_RAT3: {
         double a = val - 1.5 / rate, b = val + 1.5 / rate;
         if (psi <= a) return 0.0;
         if (psi >= b) return 1.0;
         double x = (psi - a) / (b - a);
         return (x * x * x) / ((1 - x) * (1 - x) * (1 - x) + x * x * x);
      }
      case KGPU_SWITCH_COS: {
         double a = val - 0.25 * PI / rate, b = val + 0.25 * PI / rate;
         if (psi <= a) return 0.0;
         if (psi >= b) return 1.0;
         double x = (psi - a) / (b - a);
         return 0.5 * (1.0 - cos(PI * x));
      }
      case KGPU_SWITCH_LINEAR: return psi / P.maxPack;
      case KGPU_SWITCH_EQUAL: return 0.5;
      case KGPU_SWITCH_ZERO: return 0.0;
      case KGPU_SWITCH_ONE: return 1.0;
      case KGPU_SWITCH_STEP: return psi < val ? 0.0 : 1.0;
   }
}
// Closures.f90:441-464
__device__ inline double pouliquenMu(const DevParams &P, double gcos, double Hn, double modu) {
   double mu1 = P.PoulMin, mu2 = P.PoulMax, beta = P.PoulBeta;
   if (Hn > P.Hneps) {
      double Fr = modu / sqrt(gcos * Hn);
      double I = Fr * P.SolidDiameter / Hn;
      return mu1 + (mu2 - mu1) * I / (beta + I);
   }
   return mu1;
}
__device__ inline double chezyDrag(const DevParams &P, const CellState &q) { return P.ChezyCo * speed2(P, q.u, q.v, q.bx, q.by); }
__device__ inline double coulombDrag(const DevParams &P, const CellState &q) {
   double gam = gamma2(P, q.bx, q.by);
   double g = P.g / gam;
   return P.CoulombCo * g * q.Hn;
}
__device__ inline double pouliquenDrag(const DevParams &P, const CellState &q) {
   double gam = gamma2(P, q.bx, q.by);
   double g = P.g / gam;
   double modu2 = speed2(P, q.u, q.v, q.bx, q.by);
   double mu = pouliquenMu(P, g, q.Hn, sqrt(modu2));
   return modu2 > 0 ? mu * g * q.Hn : 0.0;
}
// Closures.f90:365-558 -- basal drag variants selected at runtime
__device__ inline double dragClosure(const DevParams &P, const CellState &q) {
   switch (P.drag) {
      default:
      case KGPU_DRAG_CHEZY: return chezyDrag(P, q);
      case KGPU_DRAG_COULOMB: return coulombDrag(P, q);
      case KGPU_DRAG_VOELLMY: return chezyDrag(P, q) + coulombDrag(P, q);
      case KGPU_DRAG_POULIQUEN: return pouliquenDrag(P, q);
      case KGPU_DRAG_EDWARDS2019: {
         double Hn = q.Hn;
         double gam = gamma2(P, q.bx, q.by);
         double gperp = P.g / gam;
         double modu = sqrt(speed2(P, q.u, q.v, q.bx, q.by));
         double Fr = modu / sqrt(gperp * Hn);
         double mu1 = P.PoulMin, mu2 = P.PoulMax, mu3 = P.PoulInt;
         double beta = P.PoulBeta, betastar = P.EdBetastar, kappa = P.EdKappa, capgam = P.EdGamma, L = P.SolidDiameter;
         double fr;
         if (Fr > betastar) {
            fr = mu1 + (mu2 - mu1) / (1.0 + Hn * beta / (L * (Fr + capgam)));
         } else {
            fr = (pow(Fr / betastar, kappa)) *
                     (mu1 + (mu2 - mu1) / (1.0 + Hn * beta / (L * (betastar + capgam))) - mu3 - (mu2 - mu1) / (1.0 + Hn / L)) +
                 mu3 + (mu2 - mu1) / (1.0 + Hn / L);
         }
         return fr * gperp * Hn;
      }
      case KGPU_DRAG_VARIABLE: {
         double cf = chezyDrag(P, q), pf = pouliquenDrag(P, q), fc = fswitch(P, q.psi);
         return cf * (1.0 - fc) + pf * fc;
      }
      case KGPU_DRAG_MANNING: {
         double gam = gamma2(P, q.bx, q.by);
         double g = P.g / gam;
         if (q.Hn > P.Hneps) return g * P.ManningCo * P.ManningCo / pow(q.Hn, 1.0 / 3.0);
         return 0.0;
      }
   }
}

// Closures.f90:208-241
__device__ inline double shieldsNumber(const DevParams &P, const CellState &q) {
   double gred = P.gred / gamma2(P, q.bx, q.by);
   double cf = P.ChezyCo * speed2(P, q.u, q.v, q.bx, q.by);
   return cf / (gred * P.SolidDiameter);
}
__device__ inline double particleSpeed(const DevParams &P, const CellState &q) {
   double gred = P.gred / gamma2(P, q.bx, q.by);
   return sqrt(gred * P.SolidDiameter);
}
// Closures.f90:566-675
__device__ inline double fluidErosion(const DevParams &P, const CellState &q) {
   double s = shieldsNumber(P, q);
   if (s > P.CriticalShields) {
      double ero = P.EroRate * (s - P.CriticalShields);
      return ero * particleSpeed(P, q);
   }
   return 0.0;
}
__device__ inline double granularErosion(const DevParams &P, const CellState &q) {
   const double PI = 3.141592653589793238462643383279502884;
   double modu2 = speed2(P, q.u, q.v, q.bx, q.by);
   double mn = P.PoulMin;
   double t1 = tan(PI / 180.0);
   double stat = (mn + t1) / (1.0 - mn * t1);
   double gcos = P.g / gamma2(P, q.bx, q.by);
   double mu = pouliquenMu(P, gcos, q.Hn, sqrt(modu2));
   double r = q.Hn / 25.0 / P.SolidDiameter;
   double muN = mn + (stat - mn) / (1.0 + r * r);
   if (mu > muN) {
      double ero = P.EroRateGranular * (mu - muN);
      return ero * particleSpeed(P, q);
   }
   return 0.0;
}
__device__ inline double erosionClosure(const DevParams &P, const CellState &q) {
   switch (P.erosion) {
      default:
      case KGPU_ERO_OFF: return 0.0;
      case KGPU_ERO_SIMPLE: {
         double ero = P.EroRate * shieldsNumber(P, q);
         return ero * particleSpeed(P, q);
      }
      case KGPU_ERO_FLUID: return fluidErosion(P, q);
      case KGPU_ERO_GRANULAR: return granularErosion(P, q);
      case KGPU_ERO_MIXED: {
         double fc = fswitch(P, q.psi);
         double fe = fluidErosion(P, q), ge = granularErosion(P, q);
         return (1.0 - fc) * fe + fc * ge;
      }
   }
}
// Closures.f90:684-732
__device__ inline double erosionTransition(const DevParams &P, const CellState &q) {
   switch (P.eroTrans) {
      default:
      case KGPU_EROTRANS_SMOOTH: return 0.5 * (1.0 + tanh(1e5 * (q.bt + P.EroDepth)));
      case KGPU_EROTRANS_STEP: return q.bt < -P.EroDepth ? 0.0 : 1.0;
      case KGPU_EROTRANS_OFF: return 1.0;
   }
}
// Closures.f90:320-356
__device__ inline double depositionClosure(const DevParams &P, double psi) {
   switch (P.deposition) {
      case KGPU_DEP_NONE: return 0.0;
      case KGPU_DEP_SIMPLE: return psi * (1.0 - psi / P.maxPack);
      default:
      case KGPU_DEP_SPEARMAN_MANNING: {
         double a = 2.7 - 0.15 * P.nsettling;
         double b = 0.62 * P.nsettling - 1.46;
         return psi * pow(1.0 - psi, a) * pow(1.0 - psi / P.maxPack, b);
      }
   }
}
// Closures.f90:744-791
__device__ inline double morphoDamping(const DevParams &P, double Hn) {
   double Hc = P.EroCritH;
   switch (P.damp) {
      case KGPU_DAMP_NONE: return 1.0;
      default:
      case KGPU_DAMP_TANH: return 0.5 * (1.0 + tanh(10.0 * log(Hn / Hc)));
      case KGPU_DAMP_RAT3: {
         if (Hn < Hc) return 0.0;
         if (Hn > 2.0 * Hc) return 1.0;
         double tt = Hn / Hc - 1.0;
         return (tt * tt * tt) / ((1.0 - tt) * (1.0 - tt) * (1.0 - tt) + tt * tt * tt);
      }
   }
}
// Equations.f90:385-448: returns E - D
__device__ inline double erosionMinusDeposition(const DevParams &P, const CellState &q) {
   double E = erosionClosure(P, q) * erosionTransition(P, q);
   double alpha;
   if (q.psi >= P.maxPack) alpha = 0.0;
   else if (q.psi > 0.0) alpha = depositionClosure(P, q.psi);
   else alpha = 0.0;
   double D = P.ws0 * alpha;
   double damping = morphoDamping(P, q.Hn);
   E = E * damping;
   D = D * damping;
   return E - D;
}

// HydraulicRHS.f90:762-878 -- desingularised Hn, psi, rho, u, v from the primary variables
__device__ __forceinline__ void desingularise(const DevParams &P, CellState &q, bool velocities) {
   double gam = gamma2(P, q.bx, q.by);
   double Hn = computeHn(q.w, q.b0, q.bt, gam);
   double Hnpsi = q.hpsi;
   if (Hn < 0.0) Hn = 0.0;
   if (Hnpsi < 0.0) Hnpsi = 0.0;
   double den = Hn * Hn + fmax(Hn * Hn, P.Hneps * P.Hneps);
   double psi = fmin(2.0 * Hn * Hnpsi / den, P.maxPack);
   double rho = P.rhow + (P.rhos - P.rhow) * psi;
   q.Hn = Hn; q.psi = psi; q.rho = rho;
   if (velocities) {
      q.u = 2.0 * Hn * q.hu / den / rho;
      q.v = P.oneD ? 0.0 : 2.0 * Hn * q.hv / den / rho;
   }
}

// Grid.f90:339-353 evaluated tile-relative exactly as the reference does
__device__ __forceinline__ double cellX(const DevParams &P, int i) {
   int gi = P.gtx0 + i / P.nX + 1, ti = i % P.nX + 1;
   return -0.5 * P.xSize + P.dx * ((gi - 1.0) * P.nX + (ti - 0.5));
}
__device__ __forceinline__ double cellY(const DevParams &P, int j) {
   int gj = P.gty0 + j / P.nY + 1, tj = j % P.nY + 1;
   return -0.5 * P.ySize + P.dy * ((gj - 1.0) * P.nY + (tj - 0.5));
}

// Equations.f90:501-599 -- total volumetric and solids flux of all sources at a cell centre
__device__ inline void fluxSources(const DevParams &P, const DevSource *src, const double *pool, double tEval, double tGrid, double x, double y,
                                   double &Qt, double &psiQt) {
   double s = 0.0, sp = 0.0;
   for (int J = 0; J < P.nSources; J++) {
      const DevSource &S = src[J];
      const double *Stime = pool + S.off, *Sflux = Stime + S.n, *Spsi = Sflux + S.n;
      double Qf = 0.0, psiQf = 0.0;
      if (((x - S.x) * (x - S.x) + (y - S.y) * (y - S.y)) < S.radius * S.radius) {
         int n = S.n;
         if (n == 1) {
            if (tEval < Stime[0] || (tEval == Stime[0] && tGrid < tEval)) {
               Qf = 0.0; psiQf = 0.0;
            } else {
               Qf = Sflux[0];
               psiQf = Spsi[0] * Qf;
            }
         } else {
            if (tEval < Stime[0] || tEval > Stime[n - 1] || (tEval == Stime[0] && tGrid < tEval) ||
                (tEval == Stime[n - 1] && tGrid == tEval)) {
               Qf = 0.0; psiQf = 0.0;
            } else {
               for (int K = 1; K < n; K++) {
                  if (tEval >= Stime[K - 1] && tEval <= Stime[K]) {
                     double Qa = Sflux[K - 1], psia = Spsi[K - 1], ta = Stime[K - 1];
                     double Qb = Sflux[K], psib = Spsi[K], tb = Stime[K];
                     Qf = Qa + (Qb - Qa) * (tEval - ta) / (tb - ta);
                     double psif = psia + (psib - psia) * (tEval - ta) / (tb - ta);
                     psiQf = psif * Qf;
                  }
               }
            }
         }
      }
      if (P.oneD) {
         Qf = Qf / S.numCells / P.dx;
         psiQf = psiQf / S.numCells / P.dx;
      } else {
         Qf = Qf / S.numCells / P.dx / P.dy;
         psiQf = psiQf / S.numCells / P.dx / P.dy;
      }
      s += Qf;
      sp += psiQf;
   }
   Qt = 0.0 + s;
   psiQt = 0.0 + sp;
}

// ---- running maxima (UpdateMaximumHeights/Speeds/Erosion/Deposit/SolidsFraction,
// TimeStepper.f90:1155-1303): value plane + time-of-maximum plane per field, first-inundation time.
// Planes are only touched where they can change: bt == 0 on a static bed, and psimax >= 0
// always (zero-initialised running maximum), so psi == 0 never updates it.
struct MaximaPtrs {
   double *Hnmax, *HnmaxT, *umax, *umaxT, *emax, *emaxT, *dmax, *dmaxT, *psimax, *psimaxT, *tfirst;
};
__device__ __forceinline__ void updateMaxima(const DevParams &P, const MaximaPtrs &M, size_t g, double tt, double Hn, double spd,
                                             double bt, double psi) {
   const bool wet = Hn > P.Hneps;
   if (wet) {
      if (M.tfirst[g] == -1) M.tfirst[g] = tt;
      if (Hn > M.Hnmax[g]) { M.Hnmax[g] = Hn; M.HnmaxT[g] = tt; }
      if (spd > M.umax[g]) { M.umax[g] = spd; M.umaxT[g] = tt; }
      if (psi > 0.0) { if (psi > M.psimax[g]) { M.psimax[g] = psi; M.psimaxT[g] = tt; } }
   }
   if (bt < 0) { if (-bt > M.emax[g]) { M.emax[g] = -bt; M.emaxT[g] = tt; } }
   if (bt > 0) { if (bt > M.dmax[g]) { M.dmax[g] = bt; M.dmaxT[g] = tt; } }
}

}  // namespace kgpu
