// kestrel_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (fp64, C++17) of Kestrel's explicit finite-volume time step, used
// only as the parity checker for the CUDA path and as the timed CPU baseline:
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.  The product (kestrel_b200/) never does.
//
// PARITY UNPINNED: the reference ships no golden vectors and cannot be compiled
// in this image (no Fortran compiler).  What pins this file is
// (a) line-by-line correspondence with the Fortran, cited per function below,
// and (b) the reference's own property tests re-expressed in tests/
// (conservation 1e-10, Hn >= -1e-14, lake at rest, tile-layout independence).
//
// Layout: one flat cell grid NX x NY (NX = nXtiles*nXpertile) with a per-tile
// active mask instead of the reference's array of tiles; the arithmetic, the
// order of operations inside every formula, the Kahan sums, the sweep structure
// of CalculateHydraulicRHS and the control flow of IntegrateTo follow the
// reference.  Known, documented deviations of the flat layout (all within the
// reference's own tile-independence tolerance, tests/runall.jl:45-46):
//   * one copy of every seam vertex / seam face (the reference keeps one per tile);
//     MorphodynamicRHS.f90:210,266 pre-add two terms at W edges / NE corners (Q4),
//     here the interior 4-term Kahan form is used everywhere;
//   * limited slopes in ghost cells are recomputed each sweep for w (the reference
//     freezes them at ghost creation, UpdateTiles.f90:669-750) and are zero for
//     every other variable (as in the reference, UpdateTiles.f90:245-252);
//   * CorrectSlopes applies the interior per-cell rule at tile seams
//     (HydraulicRHS.f90:585-611 replicate the neighbour's rule there).
//
// Exposes the same C API as include/kestrel_gpu.h with the prefix kor_.

#include "../include/kestrel_gpu.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using std::vector;
typedef vector<double> Arr;

const double EPS = std::numeric_limits<double>::epsilon();
const double HUGE_D = std::numeric_limits<double>::max();
const double PI = 3.141592653589793238462643383279502884;  // SetPrecision.f90:38

// indices into u[] (main.f90:76-101, zero-based here)
enum { W = 0, HU = 1, HV = 2, HPSI = 3, HN = 4, U = 5, V = 6, PSI = 7, RHO = 8, B0 = 9, BT = 10, BX = 11, BY = 12 };

// utilities.f90:439-448
inline void kahanAdd(double x, double &s, double &c) {
   double y = x - c;
   double t = s + y;
   c = (t - s) - y;
   s = t;
}
// utilities.f90:418-433
template <int N> inline double kahanSum(const double (&v)[N]) {
   double s = 0.0, c = 0.0;
   for (int i = 0; i < N; i++) kahanAdd(v[i], s, c);
   return s;
}

struct Source {
   double x, y, radius;
   int numCells;
   vector<double> time, flux, psi;
};

// One solution container: the flat analogue of tileContainer / intermed0..3
// (Grid.f90:141,155-156) holding what CopySolutionData moves (TimeStepper.f90:811-835).
struct Cont {
   Arr u[13];   // cell centred
   Arr btv;     // bt at vertices
   Arr E[4];    // ddtExplicit
   Arr I;       // ddtImplicit for the two momenta (identical, Equations.f90:653-654)
   Arr EBt;     // ddtExplicitBt at vertices
   Arr EmD;     // EminusD at cells
};

struct Oracle {
   kgpu_params P;
   vector<Source> src;
   std::string err;

   int nX, nY, nXt, nYt, NX, NY, NXV, NYV, nTiles;
   bool oneD, periodic, geom;
   double dx, dy, dxR, dyR;

   // tile state: 0 untouched, 1 ghost, 2 active
   vector<int> tstate;
   vector<char> hasSource;
   vector<int> activeList;  // ascending 1-based ids (utilities.f90:260)
   vector<int> ghostList;   // 1-based ids in order of creation
   vector<char> loaded;     // heights loaded

   Arr b0v;  // static bed at vertices
   Cont C, I0, I1, I2, I3;

   // scratch of the RHS sweeps (TileType uLim*, u{Plus,Minus}*, *Flux; Grid.f90:103-115)
   Arr limX[9], limY[9];
   Arr pX[13], mX[13], pY[13], mY[13];  // uPlusX, uMinusX, uPlusY, uMinusY
   Arr hX[4], hY[4], gX, gY, dfX[2], dfY[2];  // g flux identical for both momenta; p only momenta

   // maxima (Grid.f90:96-103)
   Arr Hnmax[2], umax[2], emax[2], dmax[2], psimax[2], tfirst;
   Arr HnSeed;      // Hn as uploaded: first tile-activation scan reads the IC's u(Hn)
   bool firstScan = true;

   // Q2: u, v desingularised before the final implicit correction
   double t, t0, dtgrid;
   int64_t nsteps = 0, nrefines = 0, ntilesAdded = 0;
   int nthreads = 1;
   int64_t nRedistributed = 0;
   bool haltError = false;
   // loop bounds: bounding box of the active tiles + 2 cells (cells [ilo,ihi) x [jlo,jhi))
   int ilo = 0, ihi = 0, jlo = 0, jhi = 0;
   void updateBounds() {
      refreshCellAct();
      if (activeList.empty()) { ilo = ihi = jlo = jhi = 0; return; }
      int a = NX, b = 0, c = NY, d = 0;
      for (int id : activeList) {
         int i0, i1, j0, j1;
         forTileCells(id - 1, i0, i1, j0, j1);
         a = std::min(a, i0); b = std::max(b, i1); c = std::min(c, j0); d = std::max(d, j1);
      }
      if (periodic) { a = 0; b = NX; c = 0; d = NY; }
      ilo = std::max(a - 2, 0); ihi = std::min(b + 2, NX);
      jlo = std::max(c - 2, 0); jhi = std::min(d + 2, NY);
   }

   // ---------------------------------------------------------------- indexing
   // wrap / clamp tables for i in [-4, NX+4) and the halo'd per-cell activity mask: the
   // inner loops index these instead of dividing (the reference follows tile pointers)
   vector<int> wxT, wyT, vxT, vyT;
   vector<unsigned char> cellAct;  // (NX+4) x (NY+4), halo 2
   inline int wrapi(int i) const { return wxT[i + 4]; }
   inline int wrapj(int j) const { return wyT[j + 4]; }
   inline int cidx(int i, int j) const { return wyT[j + 4] * NX + wxT[i + 4]; }
   inline int vwi(int i) const { return vxT[i + 4]; }
   inline int vwj(int j) const { return oneD ? 0 : vyT[j + 4]; }
   inline int vidx(int i, int j) const { return (oneD ? 0 : vyT[j + 4]) * NXV + vxT[i + 4]; }
   inline int fxidx(int fi, int j) const { return j * (NX + 1) + fi; }  // x-faces (NX+1) x NY
   inline int fyidx(int i, int fj) const { return fj * NX + i; }        // y-faces NX x (NY+1)
   inline int tileOfCell(int i, int j) const { return (wrapi(i) / nX) + (wrapj(j) / nY) * nXt; }
   inline bool cellActive(int i, int j) const { return cellAct[(size_t)(j + 2) * (NX + 4) + (i + 2)] != 0; }
   void buildTables() {
      wxT.resize(NX + 8); wyT.resize(NY + 8); vxT.resize(NX + 9); vyT.resize(NY + 9);
      for (int i = -4; i < NX + 4; i++) wxT[i + 4] = periodic ? ((i % NX) + NX) % NX : std::min(std::max(i, 0), NX - 1);
      for (int j = -4; j < NY + 4; j++) wyT[j + 4] = periodic ? ((j % NY) + NY) % NY : std::min(std::max(j, 0), NY - 1);
      for (int i = -4; i < NX + 5; i++) vxT[i + 4] = periodic ? ((i % NX) + NX) % NX : std::min(std::max(i, 0), NXV - 1);
      for (int j = -4; j < NY + 5; j++) vyT[j + 4] = periodic ? ((j % NY) + NY) % NY : std::min(std::max(j, 0), std::max(NYV - 1, 0));
      cellAct.assign((size_t)(NX + 4) * (NY + 4), 0);
   }
   void refreshCellAct() {
      for (int j = -2; j < NY + 2; j++)
         for (int i = -2; i < NX + 2; i++) {
            bool in = periodic || (i >= 0 && i < NX && j >= 0 && j < NY);
            unsigned char a = 0;
            if (in) a = tstate[(wxT[i + 4] / nX) + (wyT[j + 4] / nY) * nXt] == 2 ? 1 : 0;
            cellAct[(size_t)(j + 2) * (NX + 4) + (i + 2)] = a;
         }
   }

   // Closures.f90:269-305
   inline double gamma2(double bx, double by) const { return geom ? std::sqrt(1.0 + bx * bx + by * by) : 1.0; }
   inline double gammaC(const Cont &T, int c) const { return gamma2(T.u[BX][c], T.u[BY][c]); }
   // Closures.f90:158-171
   static inline double computeHn(double w, double b0, double bt, double gam) {
      double Hn = -bt;
      Hn = Hn + (w - b0);
      Hn = Hn * gam;
      return Hn;
   }
   // Closures.f90:245-258
   inline double density(double psi) const { return P.rhow + (P.rhos - P.rhow) * psi; }

   // ---------------------------------------------------------------- limiters (Limiters.f90:83-184)
   inline double limiter(double a, double b) const {
      switch (P.limiter) {
         case KGPU_LIM_MINMOD1:
            if (a * b <= 0.0) return 0.0;
            return a > 0.0 ? std::min(a, b) : std::max(a, b);
         default:
         case KGPU_LIM_MINMOD2: {
            const double theta = 1.3;
            if (a * b <= 0.0) return 0.0;
            if (a > 0.0) return std::min(theta * a, std::min(theta * b, 0.5 * (a + b)));
            return std::max(theta * a, std::max(theta * b, 0.5 * (a + b)));
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

   // ---------------------------------------------------------------- closures
   // Closures.f90:178-205
   inline double speed2(double u, double v, double bx, double by) const {
      double m = geom ? u * u * (1.0 + bx * bx) : u * u;
      if (oneD) return m;
      if (geom) m = m + v * v * (1.0 + by * by) + 2.0 * bx * by * u * v;
      else m = m + v * v;
      return m;
   }
   // Closures.f90:797-915
   double fswitch(double psi) const {
      const double rate = P.VoellmySwitchRate, val = P.VoellmySwitchValue;
      switch (P.fswitch) {
         default:
         case KGPU_SWITCH_TANH: return 0.5 * (1.0 + std::tanh(rate * (psi - val)));
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
            return 0.5 * (1.0 - std::cos(PI * x));
         }
         case KGPU_SWITCH_LINEAR: return psi / P.maxPack;
         case KGPU_SWITCH_EQUAL: return 0.5;
         case KGPU_SWITCH_ZERO: return 0.0;
         case KGPU_SWITCH_ONE: return 1.0;
         case KGPU_SWITCH_STEP: return psi < val ? 0.0 : 1.0;
      }
   }
   // Closures.f90:441-464
   double pouliquenMu(double gcos, double Hn, double modu) const {
      double mu1 = P.PouliquenMinSlope, mu2 = P.PouliquenMaxSlope, beta = P.PouliquenBeta;
      if (Hn > P.heightThreshold) {
         double Fr = modu / std::sqrt(gcos * Hn);
         double I = Fr * P.SolidDiameter / Hn;
         return mu1 + (mu2 - mu1) * I / (beta + I);
      }
      return mu1;
   }
   double chezy(const double *q) const { return P.ChezyCo * speed2(q[U], q[V], q[BX], q[BY]); }
   double coulomb(const double *q) const {
      double gam = gamma2(q[BX], q[BY]);
      double g = P.g / gam;
      return P.CoulombCo * g * q[HN];
   }
   double pouliquen(const double *q) const {
      double Hn = q[HN];
      double gam = gamma2(q[BX], q[BY]);
      double g = P.g / gam;
      double modu2 = speed2(q[U], q[V], q[BX], q[BY]);
      double mu = pouliquenMu(g, Hn, std::sqrt(modu2));
      return modu2 > 0 ? mu * g * Hn : 0.0;
   }
   // Closures.f90:365-558
   double drag(const double *q) const {
      switch (P.drag) {
         default:
         case KGPU_DRAG_CHEZY: return chezy(q);
         case KGPU_DRAG_COULOMB: return coulomb(q);
         case KGPU_DRAG_VOELLMY: return chezy(q) + coulomb(q);
         case KGPU_DRAG_POULIQUEN: return pouliquen(q);
         case KGPU_DRAG_EDWARDS2019: {
            double Hn = q[HN];
            double gam = gamma2(q[BX], q[BY]);
            double gperp = P.g / gam;
            double modu = std::sqrt(speed2(q[U], q[V], q[BX], q[BY]));
            double Fr = modu / std::sqrt(gperp * Hn);
            double mu1 = P.PouliquenMinSlope, mu2 = P.PouliquenMaxSlope, mu3 = P.PouliquenIntermediateSlope;
            double beta = P.PouliquenBeta, betastar = P.Edwards2019betastar, kappa = P.Edwards2019kappa;
            double capgam = P.Edwards2019Gamma, L = P.SolidDiameter;
            double fr;
            if (Fr > betastar) {
               fr = mu1 + (mu2 - mu1) / (1.0 + Hn * beta / (L * (Fr + capgam)));
            } else {
               fr = (std::pow(Fr / betastar, kappa)) *
                        (mu1 + (mu2 - mu1) / (1.0 + Hn * beta / (L * (betastar + capgam))) - mu3 -
                         (mu2 - mu1) / (1.0 + Hn / L)) +
                    mu3 + (mu2 - mu1) / (1.0 + Hn / L);
            }
            return fr * gperp * Hn;
         }
         case KGPU_DRAG_VARIABLE: {
            double cf = chezy(q), pf = pouliquen(q), fc = fswitch(q[PSI]);
            return cf * (1.0 - fc) + pf * fc;
         }
         case KGPU_DRAG_MANNING: {
            double Hn = q[HN];
            double gam = gamma2(q[BX], q[BY]);
            double g = P.g / gam;
            if (Hn > P.heightThreshold) return g * P.ManningCo * P.ManningCo / std::pow(Hn, 1.0 / 3.0);
            return 0.0;
         }
      }
   }
   // Closures.f90:208-241
   double shields(const double *q) const {
      double gred = P.gred / gamma2(q[BX], q[BY]);
      double cf = P.ChezyCo * speed2(q[U], q[V], q[BX], q[BY]);
      return cf / (gred * P.SolidDiameter);
   }
   double particleSpeed(const double *q) const {
      double gred = P.gred / gamma2(q[BX], q[BY]);
      return std::sqrt(gred * P.SolidDiameter);
   }
   // Closures.f90:566-675
   double fluidErosion(const double *q) const {
      double s = shields(q);
      if (s > P.CriticalShields) {
         double ero = P.EroRate * (s - P.CriticalShields);
         return ero * particleSpeed(q);
      }
      return 0.0;
   }
   double granularErosion(const double *q) const {
      double Hn = q[HN];
      double modu2 = speed2(q[U], q[V], q[BX], q[BY]);
      double mn = P.PouliquenMinSlope;
      double t1 = std::tan(PI / 180.0);
      double stat = (mn + t1) / (1.0 - mn * t1);
      double gcos = P.g / gamma2(q[BX], q[BY]);
      double mu = pouliquenMu(gcos, Hn, std::sqrt(modu2));
      double r = Hn / 25.0 / P.SolidDiameter;
      double muN = mn + (stat - mn) / (1.0 + r * r);  // (..)**2.0_wp
      if (mu > muN) {
         double ero = P.EroRateGranular * (mu - muN);
         return ero * particleSpeed(q);
      }
      return 0.0;
   }
   double erosionClosure(const double *q) const {
      switch (P.erosion) {
         default:
         case KGPU_ERO_OFF: return 0.0;
         case KGPU_ERO_SIMPLE: {
            double ero = P.EroRate * shields(q);
            return ero * particleSpeed(q);
         }
         case KGPU_ERO_FLUID: return fluidErosion(q);
         case KGPU_ERO_GRANULAR: return granularErosion(q);
         case KGPU_ERO_MIXED: {
            double fc = fswitch(q[PSI]);
            double fe = fluidErosion(q), ge = granularErosion(q);
            return (1.0 - fc) * fe + fc * ge;
         }
      }
   }
   // Closures.f90:684-732
   double erosionTransition(const double *q) const {
      switch (P.erosion_transition) {
         default:
         case KGPU_EROTRANS_SMOOTH: return 0.5 * (1.0 + std::tanh(1e5 * (q[BT] + P.EroDepth)));
         case KGPU_EROTRANS_STEP: return q[BT] < -P.EroDepth ? 0.0 : 1.0;
         case KGPU_EROTRANS_OFF: return 1.0;
      }
   }
   // Closures.f90:320-356
   double depositionClosure(double psi) const {
      switch (P.deposition) {
         case KGPU_DEP_NONE: return 0.0;
         case KGPU_DEP_SIMPLE: return psi * (1.0 - psi / P.maxPack);
         default:
         case KGPU_DEP_SPEARMAN_MANNING: {
            double a = 2.7 - 0.15 * P.nsettling;
            double b = 0.62 * P.nsettling - 1.46;
            return psi * std::pow(1.0 - psi, a) * std::pow(1.0 - psi / P.maxPack, b);
         }
      }
   }
   // Closures.f90:744-791
   double morphoDamping(double Hn) const {
      double Hc = P.EroCriticalHeight;
      switch (P.morpho_damp) {
         case KGPU_DAMP_NONE: return 1.0;
         default:
         case KGPU_DAMP_TANH: return 0.5 * (1.0 + std::tanh(10.0 * std::log(Hn / Hc)));
         case KGPU_DAMP_RAT3: {
            if (Hn < Hc) return 0.0;
            if (Hn > 2.0 * Hc) return 1.0;
            double tt = Hn / Hc - 1.0;
            return (tt * tt * tt) / ((1.0 - tt) * (1.0 - tt) * (1.0 - tt) + tt * tt * tt);
         }
      }
   }
   // Equations.f90:385-448
   void erosionDeposition(const double *q, double &E, double &D) const {
      E = erosionClosure(q) * erosionTransition(q);
      double psi = q[PSI], alpha;
      if (psi >= P.maxPack) alpha = 0.0;
      else if (psi > 0.0) alpha = depositionClosure(psi);
      else alpha = 0.0;
      D = P.ws0 * alpha;
      double damping = morphoDamping(q[HN]);
      E = E * damping;
      D = D * damping;
   }

   // ---------------------------------------------------------------- geometry
   // Grid.f90:339-353 (tile-relative, as the reference evaluates it)
   inline double cellX(int i) const {
      int gi = i / nX + 1, ti = i % nX + 1;
      return -0.5 * P.xSize + dx * ((gi - 1.0) * nX + (ti - 0.5));
   }
   inline double cellY(int j) const {
      int gj = j / nY + 1, tj = j % nY + 1;
      return -0.5 * P.ySize + dy * ((gj - 1.0) * nY + (tj - 0.5));
   }

   void forTileCells(int tile0, int &i0, int &i1, int &j0, int &j1) const {
      int tx = tile0 % nXt, ty = tile0 / nXt;
      i0 = tx * nX; i1 = i0 + nX; j0 = ty * nY; j1 = j0 + nY;
   }

   // MorphodynamicRHS.f90:308-368 for one cell; withB0 = also refresh b0 centre
   void centreTopo(Cont &T, int i, int j, bool withB0) {
      int c = j * NX + i;
      const Arr &bt = T.btv;
      if (!oneD) {
         if (withB0) {
            double a[4] = {b0v[vidx(i, j)], b0v[vidx(i + 1, j)], b0v[vidx(i, j + 1)], b0v[vidx(i + 1, j + 1)]};
            T.u[B0][c] = 0.25 * kahanSum(a);
         }
         double b[4] = {bt[vidx(i, j)], bt[vidx(i + 1, j)], bt[vidx(i, j + 1)], bt[vidx(i + 1, j + 1)]};
         T.u[BT][c] = 0.25 * kahanSum(b);
         double ax[8] = {b0v[vidx(i + 1, j)], bt[vidx(i + 1, j)], -b0v[vidx(i, j)], -bt[vidx(i, j)],
                         b0v[vidx(i + 1, j + 1)], bt[vidx(i + 1, j + 1)], -b0v[vidx(i, j + 1)], -bt[vidx(i, j + 1)]};
         T.u[BX][c] = 0.5 * dxR * kahanSum(ax);
         double ay[8] = {b0v[vidx(i, j + 1)], bt[vidx(i, j + 1)], -b0v[vidx(i, j)], -bt[vidx(i, j)],
                         b0v[vidx(i + 1, j + 1)], bt[vidx(i + 1, j + 1)], -b0v[vidx(i + 1, j)], -bt[vidx(i + 1, j)]};
         T.u[BY][c] = 0.5 * dyR * kahanSum(ay);
      } else {
         if (withB0) T.u[B0][c] = 0.5 * (b0v[vidx(i, 0)] + b0v[vidx(i + 1, 0)]);
         T.u[BT][c] = 0.5 * (bt[vidx(i, 0)] + bt[vidx(i + 1, 0)]);
         double ax[4] = {b0v[vidx(i + 1, 0)], bt[vidx(i + 1, 0)], -b0v[vidx(i, 0)], -bt[vidx(i, 0)]};
         T.u[BX][c] = dxR * kahanSum(ax);
      }
   }
   void centreTopoTile(Cont &T, int tile0, bool withB0) {
      int i0, i1, j0, j1;
      forTileCells(tile0, i0, i1, j0, j1);
      for (int j = j0; j < j1; j++)
         for (int i = i0; i < i1; i++) centreTopo(T, i, j, withB0);
   }

   // HydraulicRHS.f90:741-753
   static inline double interpolateB(double b0_0, double b0_1, double bt_0, double bt_1) {
      double b = bt_0 + bt_1;
      b = b + b0_0;
      b = b + b0_1;
      b = b * 0.5;
      return b;
   }

   // ---------------------------------------------------------------- sweeps of CalculateHydraulicRHS
   // HydraulicRHS.f90:762-878
   void desingularise(Cont &T, int c, bool velocities) const {
      double Hneps = P.heightThreshold;
      double rhoHnu = T.u[HU][c], Hnpsi = T.u[HPSI][c];
      double gam = gammaC(T, c);
      double Hn = computeHn(T.u[W][c], T.u[B0][c], T.u[BT][c], gam);
      if (Hn < 0.0) Hn = 0.0;
      if (Hnpsi < 0.0) Hnpsi = 0.0;
      double psi = std::min(2.0 * Hn * Hnpsi / (Hn * Hn + std::max(Hn * Hn, Hneps * Hneps)), P.maxPack);
      double rho = density(psi);
      T.u[HN][c] = Hn;
      T.u[PSI][c] = psi;
      T.u[RHO][c] = rho;
      if (velocities) {
         T.u[U][c] = 2.0 * Hn * rhoHnu / (Hn * Hn + std::max(Hn * Hn, Hneps * Hneps)) / rho;
         if (!oneD) {
            double rhoHnv = T.u[HV][c];
            T.u[V][c] = 2.0 * Hn * rhoHnv / (Hn * Hn + std::max(Hn * Hn, Hneps * Hneps)) / rho;
         }
      }
   }

   // needX(i,j): cell whose x-slope / x-face values are consumed by an active tile
   inline bool needX(int i, int j) const { return cellActive(i, j) || cellActive(i - 1, j) || cellActive(i + 1, j); }
   inline bool needY(int i, int j) const { return cellActive(i, j) || cellActive(i, j - 1) || cellActive(i, j + 1); }

   // HydraulicRHS.f90:181-386.  d0..d1 = variable range.  In ghost cells only w gets
   // a slope (UpdateTiles.f90:669-750); everything else stays 0 (UpdateTiles.f90:245-252).
   void limitedDerivs(Cont &T, int d0, int d1) {
#pragma omp parallel for num_threads(nthreads) schedule(static)
      for (int j = jlo; j < jhi; j++)
         for (int i = ilo; i < ihi; i++) {
            int c = j * NX + i;
            bool act = cellActive(i, j);
            bool nx = act || cellActive(i - 1, j) || cellActive(i + 1, j);
            bool ny = !oneD && (act || cellActive(i, j - 1) || cellActive(i, j + 1));
            if (!nx && !ny) continue;
            for (int d = d0; d <= d1; d++) {
               bool live = act || d == W;
               if (nx) {
                  double s = 0.0;
                  if (live) {
                     int e = cidx(i + 1, j), w = cidx(i - 1, j);
                     s = dxR * limiter(T.u[d][e] - T.u[d][c], T.u[d][c] - T.u[d][w]);
                  }
                  limX[d][c] = s;
               }
               if (oneD) {
                  limY[d][c] = 0.0;
               } else if (ny) {
                  double s = 0.0;
                  if (live) {
                     int n = cidx(i, j + 1), so = cidx(i, j - 1);
                     s = dyR * limiter(T.u[d][n] - T.u[d][c], T.u[d][c] - T.u[d][so]);
                  }
                  limY[d][c] = s;
               }
            }
         }
   }

   // face topography (MorphodynamicRHS.f90:419-578, dem.f90:380-392); same on both sides
   void faceTopoX(const Cont &T, int fi, int j, double &b0f, double &btf, double &bxf, double &byf) const {
      const Arr &bt = T.btv;
      if (!oneD) {
         b0f = 0.5 * (b0v[vidx(fi, j)] + b0v[vidx(fi, j + 1)]);
         btf = 0.5 * (bt[vidx(fi, j)] + bt[vidx(fi, j + 1)]);
         double ay[4] = {b0v[vidx(fi, j + 1)], bt[vidx(fi, j + 1)], -b0v[vidx(fi, j)], -bt[vidx(fi, j)]};
         byf = dyR * kahanSum(ay);
         double ax[8] = {b0v[vidx(fi + 1, j)], bt[vidx(fi + 1, j)], b0v[vidx(fi + 1, j + 1)], bt[vidx(fi + 1, j + 1)],
                         -b0v[vidx(fi - 1, j)], -bt[vidx(fi - 1, j)], -b0v[vidx(fi - 1, j + 1)], -bt[vidx(fi - 1, j + 1)]};
         bxf = 0.25 * dxR * kahanSum(ax);
      } else {
         b0f = b0v[vidx(fi, 0)];
         btf = bt[vidx(fi, 0)];
         double ax[4] = {b0v[vidx(fi + 1, 0)], bt[vidx(fi + 1, 0)], -b0v[vidx(fi - 1, 0)], -bt[vidx(fi - 1, 0)]};
         bxf = 0.5 * dxR * kahanSum(ax);
         byf = 0.0;
      }
   }
   void faceTopoY(const Cont &T, int i, int fj, double &b0f, double &btf, double &bxf, double &byf) const {
      const Arr &bt = T.btv;
      b0f = 0.5 * (b0v[vidx(i, fj)] + b0v[vidx(i + 1, fj)]);
      btf = 0.5 * (bt[vidx(i, fj)] + bt[vidx(i + 1, fj)]);
      double ax[4] = {b0v[vidx(i + 1, fj)], bt[vidx(i + 1, fj)], -b0v[vidx(i, fj)], -bt[vidx(i, fj)]};
      bxf = dxR * kahanSum(ax);
      double ay[8] = {b0v[vidx(i + 1, fj + 1)], bt[vidx(i + 1, fj + 1)], b0v[vidx(i, fj + 1)], bt[vidx(i, fj + 1)],
                      -b0v[vidx(i + 1, fj - 1)], -bt[vidx(i + 1, fj - 1)], -b0v[vidx(i, fj - 1)], -bt[vidx(i, fj - 1)]};
      byf = 0.25 * dyR * kahanSum(ay);
   }

   // HydraulicRHS.f90:397-552.  pass1: d in {w,rhoHnu,rhoHnv,Hnpsi}; pass2: Hn..rho + overwrite
   void reconstruct(const Cont &T, bool pass2) {
      int d0 = pass2 ? HN : W, d1 = pass2 ? RHO : HPSI;
#pragma omp parallel for num_threads(nthreads) schedule(static)
      for (int j = jlo; j < jhi; j++)
         for (int fi = ilo; fi <= ihi; fi++) {
            // a cell's two face values are needed whenever its slope is (CorrectSlopes reads both)
            bool doP = needX(fi, j), doM = needX(fi - 1, j);
            if (!periodic) { if (fi == NX) doP = false; if (fi == 0) doM = false; }
            if (!doP && !doM) continue;
            int cl = cidx(fi - 1, j), cr = cidx(fi, j);
            int f = fxidx(fi, j);
            for (int d = d0; d <= d1; d++) {
               if (doP) pX[d][f] = T.u[d][cr] - limX[d][cr] * 0.5 * dx;
               if (doM) mX[d][f] = T.u[d][cl] + limX[d][cl] * 0.5 * dx;
            }
            if (!pass2) {
               double b0f, btf, bxf, byf;
               faceTopoX(T, fi, j, b0f, btf, bxf, byf);
               pX[B0][f] = mX[B0][f] = b0f;
               pX[BT][f] = mX[BT][f] = btf;
               pX[BX][f] = mX[BX][f] = bxf;
               pX[BY][f] = mX[BY][f] = byf;
            } else {
               double gam = gamma2(pX[BX][f], pX[BY][f]);
               pX[HN][f] = computeHn(pX[W][f], pX[B0][f], pX[BT][f], gam);
               mX[HN][f] = computeHn(mX[W][f], mX[B0][f], mX[BT][f], gam);
               pX[HU][f] = pX[RHO][f] * pX[HN][f] * pX[U][f];
               mX[HU][f] = mX[RHO][f] * mX[HN][f] * mX[U][f];
               if (!oneD) {
                  pX[HV][f] = pX[RHO][f] * pX[HN][f] * pX[V][f];
                  mX[HV][f] = mX[RHO][f] * mX[HN][f] * mX[V][f];
               }
            }
         }
      if (oneD) return;
#pragma omp parallel for num_threads(nthreads) schedule(static)
      for (int fj = jlo; fj <= jhi; fj++)
         for (int i = ilo; i < ihi; i++) {
            bool doP = needY(i, fj), doM = needY(i, fj - 1);
            if (!periodic) { if (fj == NY) doP = false; if (fj == 0) doM = false; }
            if (!doP && !doM) continue;
            int cl = cidx(i, fj - 1), cr = cidx(i, fj);
            int f = fyidx(i, fj);
            for (int d = d0; d <= d1; d++) {
               if (doP) pY[d][f] = T.u[d][cr] - limY[d][cr] * 0.5 * dy;
               if (doM) mY[d][f] = T.u[d][cl] + limY[d][cl] * 0.5 * dy;
            }
            if (!pass2) {
               double b0f, btf, bxf, byf;
               faceTopoY(T, i, fj, b0f, btf, bxf, byf);
               pY[B0][f] = mY[B0][f] = b0f;
               pY[BT][f] = mY[BT][f] = btf;
               pY[BX][f] = mY[BX][f] = bxf;
               pY[BY][f] = mY[BY][f] = byf;
            } else {
               double gam = gamma2(pY[BX][f], pY[BY][f]);
               pY[HN][f] = computeHn(pY[W][f], pY[B0][f], pY[BT][f], gam);
               mY[HN][f] = computeHn(mY[W][f], mY[B0][f], mY[BT][f], gam);
               pY[HU][f] = pY[RHO][f] * pY[HN][f] * pY[U][f];
               mY[HU][f] = mY[RHO][f] * mY[HN][f] * mY[U][f];
               pY[HV][f] = pY[RHO][f] * pY[HN][f] * pY[V][f];
               mY[HV][f] = mY[RHO][f] * mY[HN][f] * mY[V][f];
            }
         }
   }

   // HydraulicRHS.f90:560-736, per-cell rule (:613-639, :693-712)
   void correctSlopes(const Cont &T) {
      const Arr &bt = T.btv;
#pragma omp parallel for num_threads(nthreads) schedule(static)
      for (int j = jlo; j < jhi; j++)
         for (int i = ilo; i < ihi; i++) {
            int c = j * NX + i;
            if (needX(i, j)) {
               double bl, br;
               if (!oneD) {
                  bl = interpolateB(b0v[vidx(i, j)], b0v[vidx(i, j + 1)], bt[vidx(i, j)], bt[vidx(i, j + 1)]);
                  br = interpolateB(b0v[vidx(i + 1, j)], b0v[vidx(i + 1, j + 1)], bt[vidx(i + 1, j)], bt[vidx(i + 1, j + 1)]);
               } else {
                  bl = b0v[vidx(i, 0)] + bt[vidx(i, 0)];
                  br = b0v[vidx(i + 1, 0)] + bt[vidx(i + 1, 0)];
               }
               int fl = fxidx(i, j), fr = fxidx(i + 1, j);
               // faces of this cell: left = uPlusX(i), right = uMinusX(i+1).  With periodic wrap
               // the face arrays hold both copies (fi = 0 and fi = NX); keep them identical.
               double &wR = mX[W][fr], &wL = pX[W][fl];
               if ((wR < br) || (wL < bl)) {
                  wR = T.u[W][c] + 0.5 * (br - bl);
                  wL = T.u[W][c] + 0.5 * (bl - br);
               }
               double &sR = mX[HPSI][fr], &sL = pX[HPSI][fl];
               if ((sR < 0.0) || (sL < 0.0)) {
                  sR = T.u[HPSI][c];
                  sL = T.u[HPSI][c];
               }
               if (periodic) {
                  if (i == 0) { pX[W][fxidx(NX, j)] = wL; pX[HPSI][fxidx(NX, j)] = sL; }
                  if (i == NX - 1) { mX[W][fxidx(0, j)] = wR; mX[HPSI][fxidx(0, j)] = sR; }
               }
            }
            if (!oneD && needY(i, j)) {
               double bb = interpolateB(b0v[vidx(i, j)], b0v[vidx(i + 1, j)], bt[vidx(i, j)], bt[vidx(i + 1, j)]);
               double btp = interpolateB(b0v[vidx(i, j + 1)], b0v[vidx(i + 1, j + 1)], bt[vidx(i, j + 1)], bt[vidx(i + 1, j + 1)]);
               int fb = fyidx(i, j), ft = fyidx(i, j + 1);
               double &wT = mY[W][ft], &wB = pY[W][fb];
               if ((wT < btp) || (wB < bb)) {
                  wT = T.u[W][c] + 0.5 * (btp - bb);
                  wB = T.u[W][c] + 0.5 * (bb - btp);
               }
               double &sT = mY[HPSI][ft], &sB = pY[HPSI][fb];
               if ((sT < 0.0) || (sB < 0.0)) {
                  sT = T.u[HPSI][c];
                  sB = T.u[HPSI][c];
               }
               if (periodic) {
                  if (j == 0) { pY[W][fyidx(i, NY)] = wB; pY[HPSI][fyidx(i, NY)] = sB; }
                  if (j == NY - 1) { mY[W][fyidx(i, 0)] = wT; mY[HPSI][fyidx(i, 0)] = sT; }
               }
            }
         }
   }

   // Equations.f90:249-381
   inline double waveC(double Hn, double gam, double bt2) const {
      if (Hn <= 0.0) Hn = 0.0;
      if (geom) return std::sqrt(P.g * Hn * (1.0 + bt2 * bt2) / (gam * gam * gam));
      return std::sqrt(P.g * Hn);
   }

   // HydraulicRHS.f90:896-1182.  Returns min over faces of the unit-CFL step.
   double calculateFluxes(const Cont &T) {
      double cflmin = HUGE_D;
      const double nu = P.EddyViscosity;
#pragma omp parallel for num_threads(nthreads) schedule(static) reduction(min : cflmin)
      for (int j = jlo; j < jhi; j++)
         for (int fi = ilo; fi <= ihi; fi++) {
            if (!(cellActive(fi - 1, j) || cellActive(fi, j))) continue;
            if (!periodic && (fi == 0 || fi == NX)) continue;
            int f = fxidx(fi, j);
            int cl = cidx(fi - 1, j), cr = cidx(fi, j);
            double gamf = gamma2(pX[BX][f], pX[BY][f]);
            double by = pX[BY][f];
            double cP = waveC(pX[HN][f], gamf, by), cM = waveC(mX[HN][f], gamf, by);
            double wsP = pX[U][f] + cP, wsM = mX[U][f] + cM;
            double aPos = wsP > wsM ? wsP : wsM;
            if (aPos < 0.0) aPos = 0.0;
            wsP = pX[U][f] - cP; wsM = mX[U][f] - cM;
            double aNeg = wsP < wsM ? wsP : wsM;
            if (aNeg > 0.0) aNeg = 0.0;
            if (aPos > EPS) {
               double gr = std::min(gammaC(T, cl) / gamf, 1.0);
               cflmin = std::min(gr * gr * dx / aPos, cflmin);
            }
            if (std::fabs(aNeg) > EPS) {
               double gr = std::min(gammaC(T, cr) / gamf, 1.0);
               cflmin = std::min(gr * gr * dx / std::fabs(aNeg), cflmin);
            }
            // convection (Equations.f90:53-77), hydrostatic (:109-138), diffusion (:176-208)
            double cvP[4], cvM[4];
            cvP[W] = pX[HN][f] * pX[U][f] * gamf;       cvM[W] = mX[HN][f] * mX[U][f] * gamf;
            cvP[HPSI] = pX[HPSI][f] * pX[U][f] * gamf;  cvM[HPSI] = mX[HPSI][f] * mX[U][f] * gamf;
            cvP[HU] = pX[HU][f] * pX[U][f];             cvM[HU] = mX[HU][f] * mX[U][f];
            cvP[HV] = pX[HV][f] * pX[U][f];             cvM[HV] = mX[HV][f] * mX[U][f];
            double hp = -pX[BT][f]; hp = hp + (pX[W][f] - pX[B0][f]);
            double hyP = 0.5 * P.g * pX[RHO][f] * hp * hp;
            hp = -mX[BT][f]; hp = hp + (mX[W][f] - mX[B0][f]);
            double hyM = 0.5 * P.g * mX[RHO][f] * hp * hp;
            double dP[2], dM[2];
            if (pX[HN][f] < 0.0) { dP[0] = dP[1] = 0.0; }
            else { dP[0] = nu * pX[RHO][f] * pX[HN][f] * limX[U][cr]; dP[1] = nu * pX[RHO][f] * pX[HN][f] * limX[V][cr]; }
            if (mX[HN][f] < 0.0) { dM[0] = dM[1] = 0.0; }
            else { dM[0] = nu * mX[RHO][f] * mX[HN][f] * limX[U][cl]; dM[1] = nu * mX[RHO][f] * mX[HN][f] * limX[V][cl]; }
            double dif = aPos - aNeg;
            if (dif < 1e-10) {
               for (int d = 0; d < 4; d++) hX[d][f] = 0.0;
               gX[f] = 0.0; dfX[0][f] = dfX[1][f] = 0.0;
            } else {
               double h;
               h = pX[HN][f] * gamf - mX[HN][f] * gamf;
               h = h * aPos * aNeg; h = h + (aPos * cvM[W] - aNeg * cvP[W]); h = h / dif; hX[W][f] = h;
               h = pX[HPSI][f] * gamf - mX[HPSI][f] * gamf;
               h = h * aPos * aNeg; h = h + (aPos * cvM[HPSI] - aNeg * cvP[HPSI]); h = h / dif; hX[HPSI][f] = h;
               h = pX[HU][f] - mX[HU][f];
               h = h * aPos * aNeg; h = h + (aPos * cvM[HU] - aNeg * cvP[HU]); h = h / dif; hX[HU][f] = h;
               h = pX[HV][f] - mX[HV][f];
               h = h * aPos * aNeg; h = h + (aPos * cvM[HV] - aNeg * cvP[HV]); h = h / dif; hX[HV][f] = h;
               gX[f] = (aPos * hyM - aNeg * hyP) / dif;
               dfX[0][f] = 0.5 * (dP[0] + dM[0]);
               dfX[1][f] = 0.5 * (dP[1] + dM[1]);
            }
         }
      if (oneD) return cflmin;
#pragma omp parallel for num_threads(nthreads) schedule(static) reduction(min : cflmin)
      for (int fj = jlo; fj <= jhi; fj++)
         for (int i = ilo; i < ihi; i++) {
            if (!(cellActive(i, fj - 1) || cellActive(i, fj))) continue;
            if (!periodic && (fj == 0 || fj == NY)) continue;
            int f = fyidx(i, fj);
            int cl = cidx(i, fj - 1), cr = cidx(i, fj);
            double gamf = gamma2(pY[BX][f], pY[BY][f]);
            double bx = pY[BX][f];
            double cP = waveC(pY[HN][f], gamf, bx), cM = waveC(mY[HN][f], gamf, bx);
            double wsP = pY[V][f] + cP, wsM = mY[V][f] + cM;
            double aPos = wsP > wsM ? wsP : wsM;
            if (aPos < 0.0) aPos = 0.0;
            wsP = pY[V][f] - cP; wsM = mY[V][f] - cM;
            double aNeg = wsP < wsM ? wsP : wsM;
            if (aNeg > 0.0) aNeg = 0.0;
            if (aPos > EPS) {
               double gr = std::min(gammaC(T, cl) / gamf, 1.0);
               cflmin = std::min(gr * gr * dy / aPos, cflmin);
            }
            if (std::fabs(aNeg) > EPS) {
               double gr = std::min(gammaC(T, cr) / gamf, 1.0);
               cflmin = std::min(gr * gr * dy / std::fabs(aNeg), cflmin);
            }
            double cvP[4], cvM[4];
            cvP[W] = pY[HN][f] * pY[V][f] * gamf;       cvM[W] = mY[HN][f] * mY[V][f] * gamf;
            cvP[HPSI] = pY[HPSI][f] * pY[V][f] * gamf;  cvM[HPSI] = mY[HPSI][f] * mY[V][f] * gamf;
            cvP[HU] = pY[HU][f] * pY[V][f];             cvM[HU] = mY[HU][f] * mY[V][f];
            cvP[HV] = pY[HV][f] * pY[V][f];             cvM[HV] = mY[HV][f] * mY[V][f];
            double hp = -pY[BT][f]; hp = hp + (pY[W][f] - pY[B0][f]);
            double hyP = 0.5 * P.g * pY[RHO][f] * hp * hp;
            hp = -mY[BT][f]; hp = hp + (mY[W][f] - mY[B0][f]);
            double hyM = 0.5 * P.g * mY[RHO][f] * hp * hp;
            double dP[2], dM[2];
            if (pY[HN][f] < 0.0) { dP[0] = dP[1] = 0.0; }
            else { dP[0] = nu * pY[RHO][f] * pY[HN][f] * limY[U][cr]; dP[1] = nu * pY[RHO][f] * pY[HN][f] * limY[V][cr]; }
            if (mY[HN][f] < 0.0) { dM[0] = dM[1] = 0.0; }
            else { dM[0] = nu * mY[RHO][f] * mY[HN][f] * limY[U][cl]; dM[1] = nu * mY[RHO][f] * mY[HN][f] * limY[V][cl]; }
            double dif = aPos - aNeg;
            if (dif < 1e-10) {
               for (int d = 0; d < 4; d++) hY[d][f] = 0.0;
               gY[f] = 0.0; dfY[0][f] = dfY[1][f] = 0.0;
            } else {
               double h;
               h = pY[HN][f] * gamf - mY[HN][f] * gamf;
               h = h * aPos * aNeg; h = h + (aPos * cvM[W] - aNeg * cvP[W]); h = h / dif; hY[W][f] = h;
               h = pY[HPSI][f] * gamf - mY[HPSI][f] * gamf;
               h = h * aPos * aNeg; h = h + (aPos * cvM[HPSI] - aNeg * cvP[HPSI]); h = h / dif; hY[HPSI][f] = h;
               h = pY[HU][f] - mY[HU][f];
               h = h * aPos * aNeg; h = h + (aPos * cvM[HU] - aNeg * cvP[HU]); h = h / dif; hY[HU][f] = h;
               h = pY[HV][f] - mY[HV][f];
               h = h * aPos * aNeg; h = h + (aPos * cvM[HV] - aNeg * cvP[HV]); h = h / dif; hY[HV][f] = h;
               gY[f] = (aPos * hyM - aNeg * hyP) / dif;
               dfY[0][f] = 0.5 * (dP[0] + dM[0]);
               dfY[1][f] = 0.5 * (dP[1] + dM[1]);
            }
         }
      return cflmin;
   }

   // Equations.f90:456-620: flux-source part; returns Qt and psiQt for a cell centre
   void fluxSources(double tEval, double x, double y, double &Qt, double &psiQt) const {
      Qt = 0.0; psiQt = 0.0;
      int nSrc = (int)src.size();
      if (nSrc < 1) return;
      vector<double> Qf(nSrc, 0.0), psiQf(nSrc, 0.0);
      for (int J = 0; J < nSrc; J++) {
         const Source &S = src[J];
         if (((x - S.x) * (x - S.x) + (y - S.y) * (y - S.y)) < S.radius * S.radius) {
            int n = (int)S.time.size();
            if (n == 1) {
               if (tEval < S.time[0] || (tEval == S.time[0] && t < tEval)) {
                  Qf[J] = 0.0; psiQf[J] = 0.0;
               } else {
                  Qf[J] = S.flux[0];
                  psiQf[J] = S.psi[0] * Qf[J];
               }
            } else {
               if (tEval < S.time[0] || tEval > S.time[n - 1] || (tEval == S.time[0] && t < tEval) ||
                   (tEval == S.time[n - 1] && t == tEval)) {
                  Qf[J] = 0.0; psiQf[J] = 0.0;
               } else {
                  for (int K = 1; K < n; K++) {
                     if (tEval >= S.time[K - 1] && tEval <= S.time[K]) {
                        double Qa = S.flux[K - 1], psia = S.psi[K - 1], ta = S.time[K - 1];
                        double Qb = S.flux[K], psib = S.psi[K], tb = S.time[K];
                        Qf[J] = Qa + (Qb - Qa) * (tEval - ta) / (tb - ta);
                        double psif = psia + (psib - psia) * (tEval - ta) / (tb - ta);
                        psiQf[J] = psif * Qf[J];
                     }
                  }
               }
            }
         }
         if (oneD) {
            Qf[J] = Qf[J] / S.numCells / dx;
            psiQf[J] = psiQf[J] / S.numCells / dx;
         } else {
            Qf[J] = Qf[J] / S.numCells / dx / dy;
            psiQf[J] = psiQf[J] / S.numCells / dx / dy;
         }
      }
      double s = 0.0, sp = 0.0;
      for (int J = 0; J < nSrc; J++) { s += Qf[J]; sp += psiQf[J]; }
      Qt = Qt + s;
      psiQt = psiQt + sp;
   }

   // HydraulicRHS.f90:1190-1304
   void constructRHS(Cont &T, double tEval) {
#pragma omp parallel for num_threads(nthreads) schedule(static)
      for (int j = jlo; j < jhi; j++)
         for (int i = ilo; i < ihi; i++) {
            if (!cellActive(i, j)) continue;
            int c = j * NX + i;
            double q[13];
            for (int d = 0; d < 13; d++) q[d] = T.u[d][c];
            double gam = gamma2(q[BX], q[BY]);
            double STF[4];
            int fl = fxidx(i, j), fr = fxidx(i + 1, j);
            if (!oneD) {
               int fb = fyidx(i, j), ft = fyidx(i, j + 1);
               double bx = q[BX], by = q[BY];
               double gXu, gXv, gYu, gYv;
               if (geom) {
                  gXu = (1.0 + by * by) / gam; gXv = -bx * by / gam;
                  gYu = -bx * by / gam;        gYv = (1.0 + bx * bx) / gam;
               } else { gXu = 1.0; gXv = 0.0; gYu = 0.0; gYv = 1.0; }
               STF[W] = (hX[W][fl] - hX[W][fr]) * dxR / (gam * gam) + (hY[W][fb] - hY[W][ft]) * dyR / (gam * gam);
               STF[HPSI] = (hX[HPSI][fl] - hX[HPSI][fr]) * dxR / gam + (hY[HPSI][fb] - hY[HPSI][ft]) * dyR / gam;
               {
                  double a[3] = {hX[HU][fl] - hX[HU][fr], (gX[fl] - gX[fr]) * gXu, dfX[0][fr] - dfX[0][fl]};
                  double b[3] = {hY[HU][fb] - hY[HU][ft], (gY[fb] - gY[ft]) * gYu, dfY[0][ft] - dfY[0][fb]};
                  double s = kahanSum(a) * dxR;
                  STF[HU] = s + kahanSum(b) * dyR;
               }
               {
                  double a[3] = {hX[HV][fl] - hX[HV][fr], (gX[fl] - gX[fr]) * gXv, dfX[1][fr] - dfX[1][fl]};
                  double b[3] = {hY[HV][fb] - hY[HV][ft], (gY[fb] - gY[ft]) * gYv, dfY[1][ft] - dfY[1][fb]};
                  double s = kahanSum(a) * dxR;
                  STF[HV] = s + kahanSum(b) * dyR;
               }
            } else {
               STF[W] = (hX[W][fl] - hX[W][fr]) * dxR / (gam * gam);
               STF[HPSI] = (hX[HPSI][fl] - hX[HPSI][fr]) * dxR / gam;
               double a[3] = {hX[HU][fl] - hX[HU][fr], (gX[fl] - gX[fr]) / gam, dfX[0][fr] - dfX[0][fl]};
               STF[HU] = kahanSum(a) * dxR;
               double b[3] = {hX[HV][fl] - hX[HV][fr], (gX[fl] - gX[fr]) / gam, dfX[1][fr] - dfX[1][fl]};
               STF[HV] = kahanSum(b) * dxR;
            }
            // ExplicitSourceTerms, Equations.f90:456-620
            double STE[4] = {0.0, 0.0, 0.0, 0.0};
            double Qt = 0.0, psiQt = 0.0;
            if (hasSource[tileOfCell(i, j)]) fluxSources(tEval, cellX(i), cellY(j), Qt, psiQt);
            STE[W] = STE[W] + Qt / (gam * gam);
            STE[HPSI] = STE[HPSI] + psiQt / gam;
            double hpg = -q[BT];
            hpg = hpg + (q[W] - q[B0]);
            hpg = hpg / gam;
            STE[HU] = STE[HU] - P.g * q[RHO] * hpg * q[BX];
            STE[HV] = STE[HV] - P.g * q[RHO] * hpg * q[BY];
            for (int d = 0; d < 4; d++) T.E[d][c] = STF[d] + STE[d];
            // DragClosure + ImplicitSourceTerms, Equations.f90:627-658
            double fr_ = drag(q);
            double sti = 0.0;
            if (q[HN] > P.heightThreshold) {
               double modu = std::sqrt(speed2(q[U], q[V], q[BX], q[BY]));
               if (modu > 1.0e-8) {
                  double hr = 1.0 / q[HN];
                  sti = -fr_ * hr / modu;
               }
            }
            T.I[c] = sti;
         }
   }

   // HydraulicRHS.f90:64-174
   double hydraulicRHS(Cont &T, double tEval, int substep) {
      limitedDerivs(T, W, HPSI);
      reconstruct(T, false);
      correctSlopes(T);
#pragma omp parallel for num_threads(nthreads) schedule(static)
      for (int j = jlo; j < jhi; j++)
         for (int i = ilo; i < ihi; i++)
            if (cellActive(i, j)) desingularise(T, j * NX + i, true);
      limitedDerivs(T, HN, RHO);
      reconstruct(T, true);
      double unit = calculateFluxes(T);
      // unitCFLTimeStep(:) = maxdt for every tile, active ones overwritten (HydraulicRHS.f90:82, 946)
      if ((int)activeList.size() < nTiles) unit = std::min(unit, P.maxdt);
      double maxTimeStep = std::min(P.cfl * unit, P.diffusiveTimeScale);
      maxTimeStep = std::min(maxTimeStep, P.maxdt);
      double advised;
      if (substep == 1) {
         advised = 0.9 * maxTimeStep;
         dtgrid = advised;
      } else {
         advised = maxTimeStep;
      }
      constructRHS(T, tEval);
      return advised;
   }

   // ---------------------------------------------------------------- copies (TimeStepper.f90:785-875)
   template <class F> void forActiveTiles(F f) {
      for (int id : activeList) f(id - 1);
   }
   void copyTileCells(const Arr &from, Arr &to, int tile0) {
      int i0, i1, j0, j1;
      forTileCells(tile0, i0, i1, j0, j1);
      for (int j = j0; j < j1; j++) std::memcpy(&to[j * NX + i0], &from[j * NX + i0], sizeof(double) * nX);
   }
   void copyTileVerts(const Arr &from, Arr &to, int tile0) {
      int i0, i1, j0, j1;
      forTileCells(tile0, i0, i1, j0, j1);
      int jn = oneD ? 1 : nY + 1;
      for (int lj = 0; lj < jn; lj++)
         for (int li = 0; li <= nX; li++) {
            int v = vidx(i0 + li, j0 + lj);
            to[v] = from[v];
         }
   }
   void copyMutableTopo(const Cont &from, Cont &to) {
      for (int id : activeList) {
         int t0_ = id - 1;
         copyTileCells(from.u[BT], to.u[BT], t0_);
         copyTileCells(from.u[BX], to.u[BX], t0_);
         if (!oneD) copyTileCells(from.u[BY], to.u[BY], t0_);
         copyTileVerts(from.btv, to.btv, t0_);
      }
   }
   void copySolution(const Cont &from, Cont &to) {
      copyMutableTopo(from, to);
      for (int id : activeList) {
         int t0_ = id - 1;
         for (int d = 0; d < 13; d++) copyTileCells(from.u[d], to.u[d], t0_);
         for (int d = 0; d < 4; d++) copyTileCells(from.E[d], to.E[d], t0_);
         copyTileCells(from.I, to.I, t0_);
         if (P.MorphodynamicsOn) copyTileVerts(from.EBt, to.EBt, t0_);
      }
   }
   void copyWholeTile(const Cont &from, Cont &to, int tile0) {
      for (int d = 0; d < 13; d++) copyTileCells(from.u[d], to.u[d], tile0);
      for (int d = 0; d < 4; d++) copyTileCells(from.E[d], to.E[d], tile0);
      copyTileCells(from.I, to.I, tile0);
      copyTileCells(from.EmD, to.EmD, tile0);
      copyTileVerts(from.btv, to.btv, tile0);
      copyTileVerts(from.EBt, to.EBt, tile0);
   }
   void initialiseTimeSteppingArrays() {
      for (int id : activeList) {
         copyWholeTile(C, I0, id - 1); copyWholeTile(C, I1, id - 1);
         copyWholeTile(C, I2, id - 1); copyWholeTile(C, I3, id - 1);
      }
      for (int id : ghostList) {
         copyWholeTile(C, I0, id - 1); copyWholeTile(C, I1, id - 1);
         copyWholeTile(C, I2, id - 1); copyWholeTile(C, I3, id - 1);
      }
   }

   // ---------------------------------------------------------------- tiles (UpdateTiles.f90)
   int tileW(int t0_) const {
      int tx = t0_ % nXt, ty = t0_ / nXt;
      if (tx == 0) return periodic ? (nXt - 1) + ty * nXt : -1;
      return t0_ - 1;
   }
   int tileE(int t0_) const {
      int tx = t0_ % nXt, ty = t0_ / nXt;
      if (tx == nXt - 1) return periodic ? ty * nXt : -1;
      return t0_ + 1;
   }
   int tileS(int t0_) const {
      int tx = t0_ % nXt, ty = t0_ / nXt;
      if (ty == 0) return periodic ? tx + (nYt - 1) * nXt : -1;
      return t0_ - nXt;
   }
   int tileN(int t0_) const {
      int tx = t0_ % nXt, ty = t0_ / nXt;
      if (ty == nYt - 1) return periodic ? tx : -1;
      return t0_ + nXt;
   }
   // Grid.f90:322-335
   bool onDomainEdge(int t0_) const {
      int tx = t0_ % nXt, ty = t0_ / nXt;
      bool on = (tx == 0 || tx == nXt - 1);
      on = on || (nYt > 1 && (ty == 0 || ty == nYt - 1));
      return on;
   }

   // GetHeights + EqualiseTopographicBoundaryData (dem.f90:360, MorphodynamicRHS.f90:588):
   // a shared vertex takes the value of the tile for which it is local index 1.
   int loadHeights(int t0_, const double *given) {
      if (loaded[t0_] && !given) return 0;
      vector<double> hb((size_t)(nX + 1) * (oneD ? 1 : nY + 1));
      if (given) {
         std::memcpy(hb.data(), given, hb.size() * sizeof(double));
      } else {
         if (!P.heights) { err = "no heights callback and no b0 given"; return KGPU_ERR_ARG; }
         vector<double> full((size_t)(nX + 1) * (nY + 1));
         if (P.heights(P.heights_ctx, t0_ + 1, full.data()) != 0) { err = "heights callback failed"; return KGPU_ERR_ARG; }
         for (size_t k = 0; k < hb.size(); k++) hb[k] = full[k];
      }
      int i0, i1, j0, j1;
      forTileCells(t0_, i0, i1, j0, j1);
      int tE = tileE(t0_), tN = oneD ? -1 : tileN(t0_);
      int tNE = (tE >= 0 && !oneD) ? tileN(tE) : -1;
      bool eL = tE >= 0 && loaded[tE], nL = tN >= 0 && loaded[tN], neL = tNE >= 0 && loaded[tNE];
      if (tE == t0_) eL = false;
      if (tN == t0_) nL = false;
      int jn = oneD ? 1 : nY + 1;
      for (int lj = 0; lj < jn; lj++)
         for (int li = 0; li <= nX; li++) {
            bool right = (li == nX), top = (!oneD && lj == nY);
            if (right && top) { if (eL || nL || neL) continue; }
            else if (right) { if (eL) continue; }
            else if (top) { if (nL) continue; }
            if (periodic && ((right && nXt == 1) || (top && nYt == 1))) continue;  // aliases own index 1
            b0v[vidx(i0 + li, j0 + lj)] = hb[(size_t)lj * (nX + 1) + li];
         }
      loaded[t0_] = 1;
      // refresh cell-centred topography of this tile and of the W, S, SW tiles that share the seam
      centreTopoTile(C, t0_, true);
      int tW = tileW(t0_), tS = oneD ? -1 : tileS(t0_);
      int tSW = (tW >= 0 && !oneD) ? tileS(tW) : -1;
      for (int tt : {tW, tS, tSW})
         if (tt >= 0 && tt != t0_ && loaded[tt]) {
            centreTopoTile(C, tt, true);
            if (tstate[tt] == 1) setGhostData(tt);  // w = b0 follows the refreshed centre value
         }
      return 0;
   }

   // UpdateTiles.f90:571-664
   void setGhostData(int t0_) {
      int i0, i1, j0, j1;
      forTileCells(t0_, i0, i1, j0, j1);
      bool dirichlet = (P.bcs == KGPU_BC_DIRICHLET) && onDomainEdge(t0_);
      for (int j = j0; j < j1; j++)
         for (int i = i0; i < i1; i++) {
            int c = j * NX + i;
            for (int d = 0; d < 9; d++) C.u[d][c] = 0.0;
            C.u[RHO][c] = P.rhow;
            C.u[W][c] = C.u[B0][c];
            if (dirichlet) {
               double rho = density(P.bcspsival);
               double hpval = P.bcsHnval / gammaC(C, c);
               double w = C.u[BT][c] + hpval;
               w = w + C.u[B0][c];
               C.u[W][c] = w;
               C.u[HU][c] = rho * P.bcsHnval * P.bcsuval;
               C.u[HV][c] = rho * P.bcsHnval * P.bcsvval;
               C.u[HPSI][c] = P.bcsHnval * P.bcspsival;
               C.u[HN][c] = P.bcsHnval;
               C.u[U][c] = P.bcsuval;
               C.u[V][c] = P.bcsvval;
               C.u[PSI][c] = P.bcspsival;
               C.u[RHO][c] = rho;
            }
         }
   }

   // UpdateTiles.f90:389-481
   int addGhostTiles(int t0_) {
      int nb[8];
      int n = 0;
      nb[n++] = tileW(t0_); nb[n++] = tileE(t0_);
      if (!oneD) {
         nb[n++] = tileN(t0_); nb[n++] = tileS(t0_);
         if (!periodic) {
            int s = tileS(t0_), nn = tileN(t0_);
            nb[n++] = s >= 0 ? tileW(s) : -1; nb[n++] = s >= 0 ? tileE(s) : -1;
            nb[n++] = nn >= 0 ? tileW(nn) : -1; nb[n++] = nn >= 0 ? tileE(nn) : -1;
         }
      }
      for (int k = 0; k < n; k++) {
         int tt = nb[k];
         if (tt < 0) { err = "ghost tile out of bounds (UpdateTiles.f90:423)"; return KGPU_ERR_ARG; }
         if (tstate[tt] != 0) continue;
         tstate[tt] = 1;
         ghostList.push_back(tt + 1);
         int rc = loadHeights(tt, nullptr);
         if (rc) return rc;
         setGhostData(tt);
      }
      return 0;
   }

   // UpdateTiles.f90:56-78 (+ AddToActiveTiles :81, AllocateTile :120, ActivateTile :342)
   int addTile(int t0_, bool initialise) {
      if (t0_ < 0 || t0_ >= nTiles || (onDomainEdge(t0_) && !periodic)) {
         if (P.bcs == KGPU_BC_HALT) {
            haltError = true;
            err = "tried to add a tile outside the domain (bcs = halt)";
            return KGPU_ERR_HALT_BC;
         }
         return 0;
      }
      if (tstate[t0_] == 2) return 0;
      bool wasGhost = tstate[t0_] == 1;
      tstate[t0_] = 2;
      activeList.insert(std::upper_bound(activeList.begin(), activeList.end(), t0_ + 1), t0_ + 1);
      if (wasGhost) ghostList.erase(std::find(ghostList.begin(), ghostList.end(), t0_ + 1));
      hasSource[t0_] = 0;
      int rc = loadHeights(t0_, nullptr);
      if (rc) return rc;
      if (initialise) {
         int i0, i1, j0, j1;
         forTileCells(t0_, i0, i1, j0, j1);
         for (int j = j0; j < j1; j++)
            for (int i = i0; i < i1; i++) {
               int c = j * NX + i;
               if (!wasGhost) {
                  for (int d = 0; d < 9; d++) C.u[d][c] = 0.0;
                  C.u[RHO][c] = P.rhow;
               }
               C.u[W][c] = C.u[B0][c];
               C.u[BT][c] = 0.0;
            }
      }
      updateBounds();
      rc = addGhostTiles(t0_);
      ntilesAdded++;
      return rc;
   }

   // TimeStepper.f90:924-1150 with the list-mutation semantics of Q3: the trip count is
   // fixed at loop entry while AddTile inserts into the ordered list.
   int checkIfNearBoundaries() {
      const Arr &Hn = firstScan ? HnSeed : C.u[HN];
      double thr = P.heightThreshold;
      int buf = P.TileBuffer;
      auto wetIn = [&](int t0_, int li0, int li1, int lj0, int lj1) {
         int i0, i1, j0, j1;
         forTileCells(t0_, i0, i1, j0, j1);
         for (int lj = std::max(lj0, 0); lj < std::min(lj1, nY); lj++)
            for (int li = std::max(li0, 0); li < std::min(li1, nX); li++)
               if (Hn[(j0 + lj) * NX + i0 + li] > thr) return true;
         return false;
      };
      // dir: 0 N, 1 S, 2 E, 3 W
      for (int dir = (oneD ? 2 : 0); dir < 4; dir++) {
         int trip = (int)activeList.size();
         for (int tt = 0; tt < trip; tt++) {
            int t0_ = activeList[tt] - 1;
            int nbr;
            // neighbour ids are k+-1, k+-nXtiles without bounds checks unless periodic
            // (UpdateTiles.f90:500-503); "ttN > 0" is the only guard (TimeStepper.f90:979).
            int tx = t0_ % nXt, ty = t0_ / nXt;
            switch (dir) {
               case 0: nbr = (periodic && ty == nYt - 1) ? tx : t0_ + nXt; break;
               case 1: nbr = (periodic && ty == 0) ? tx + (nYt - 1) * nXt : t0_ - nXt; break;
               case 2: nbr = (periodic && tx == nXt - 1) ? ty * nXt : t0_ + 1; break;
               default: nbr = (periodic && tx == 0) ? (nXt - 1) + ty * nXt : t0_ - 1; break;
            }
            if (nbr + 1 <= 0) continue;                       // ttN > 0
            if (nbr < nTiles && tstate[nbr] == 2) continue;   // neighbour already on
            bool trig;
            switch (dir) {
               case 0: trig = wetIn(t0_, 0, nX, nY - buf, nY) || (0 > nY - buf); break;
               case 1: trig = wetIn(t0_, 0, nX, 0, buf) || (nY <= buf); break;
               case 2: trig = wetIn(t0_, nX - buf, nX, 0, nY) || (0 > nX - buf); break;
               default: trig = wetIn(t0_, 0, buf, 0, nY) || (nX <= buf); break;
            }
            if (trig) {
               int rc = addTile(nbr, true);
               if (rc) return rc;
            }
         }
      }
      firstScan = false;
      return 0;
   }

   // ---------------------------------------------------------------- maxima (TimeStepper.f90:1155-1303)
   void updateMaxima(double tt) {
      // ComputeDesingularisedVariables(tileContainer) first (TimeStepper.f90:519)
      for (int id : activeList) {
         int i0, i1, j0, j1;
         forTileCells(id - 1, i0, i1, j0, j1);
         for (int j = j0; j < j1; j++)
            for (int i = i0; i < i1; i++) {
               int c = j * NX + i;
               desingularise(C, c, true);
               double Hn = C.u[HN][c];
               if (Hn > P.heightThreshold) {
                  if (tfirst[c] == -1) tfirst[c] = tt;
                  if (Hn > Hnmax[0][c]) { Hnmax[0][c] = Hn; Hnmax[1][c] = tt; }
               }
               double spd = std::sqrt(speed2(C.u[U][c], C.u[V][c], C.u[BX][c], C.u[BY][c]));
               if (spd > umax[0][c] && Hn > P.heightThreshold) { umax[0][c] = spd; umax[1][c] = tt; }
               double bt = C.u[BT][c];
               if (bt < 0) { if (-bt > emax[0][c]) { emax[0][c] = -bt; emax[1][c] = tt; } }
               if (bt > 0) { if (bt > dmax[0][c]) { dmax[0][c] = bt; dmax[1][c] = tt; } }
               double psi = C.u[PSI][c];
               if (Hn > P.heightThreshold) { if (psi > psimax[0][c]) { psimax[0][c] = psi; psimax[1][c] = tt; } }
            }
      }
   }

   // TimeStepper.f90:879-920 (Q13: literal test on neighbour index == 0)
   bool spongeTile(int t0_) const {
      if (!P.SpongeLayer) return false;
      int id = t0_ + 1;
      int north = id + nXt, south = id - nXt, east = id + 1, west = id - 1;
      return ((!oneD) && (north == 0 || south == 0)) || east == 0 || west == 0;
   }

   // ---------------------------------------------------------------- HydraulicTimeStepper (TimeStepper.f90:333-527)
   bool hydraulicTimeStepper(double &thisdt, double &nextT) {
      nextT = t + thisdt;
      dtgrid = thisdt;
      auto sponge = [&](Cont &T) {
         for (int id : activeList)
            if (spongeTile(id - 1)) {
               int i0, i1, j0, j1;
               forTileCells(id - 1, i0, i1, j0, j1);
               for (int j = j0; j < j1; j++)
                  for (int i = i0; i < i1; i++) {
                     int c = j * NX + i;
                     for (int d = 0; d < 4; d++) T.E[d][c] = T.E[d][c] - P.SpongeStrength * T.u[d][c];
                  }
            }
      };
      // stage 1
      sponge(I0);
#pragma omp parallel for num_threads(nthreads) schedule(static)
      for (int j = jlo; j < jhi; j++)
         for (int i = ilo; i < ihi; i++) {
            if (!cellActive(i, j)) continue;
            int c = j * NX + i;
            I1.u[W][c] = I0.u[W][c] + thisdt * I0.E[W][c];
            I1.u[HU][c] = (I0.u[HU][c] + thisdt * I0.E[HU][c]) / (1.0 - thisdt * I0.I[c]);
            I1.u[HV][c] = (I0.u[HV][c] + thisdt * I0.E[HV][c]) / (1.0 - thisdt * I0.I[c]);
            I1.u[HPSI][c] = I0.u[HPSI][c] + thisdt * I0.E[HPSI][c];
         }
      double dt1 = hydraulicRHS(I1, nextT, 2);
      if (dt1 < thisdt) { thisdt = 0.9 * dt1; return true; }
      // stage 2
      sponge(I1);
#pragma omp parallel for num_threads(nthreads) schedule(static)
      for (int j = jlo; j < jhi; j++)
         for (int i = ilo; i < ihi; i++) {
            if (!cellActive(i, j)) continue;
            int c = j * NX + i;
            I2.u[HU][c] = 0.75 * I0.u[HU][c] + 0.25 * (I1.u[HU][c] + thisdt * I1.E[HU][c]) / (1.0 - thisdt * I1.I[c]);
            I2.u[HV][c] = 0.75 * I0.u[HV][c] + 0.25 * (I1.u[HV][c] + thisdt * I1.E[HV][c]) / (1.0 - thisdt * I1.I[c]);
            double hp_old = -I0.u[BT][c];
            hp_old = hp_old + (I0.u[W][c] - I0.u[B0][c]);
            double hp_new = -I1.u[BT][c];
            hp_new = hp_new + (I1.u[W][c] - I1.u[B0][c]);
            double wu = I0.u[BT][c];
            wu = wu + 0.25 * hp_new;
            wu = wu + 0.75 * hp_old;
            wu = wu + 0.25 * thisdt * I1.E[W][c];
            wu = wu + I0.u[B0][c];
            I2.u[W][c] = wu;
            I2.u[HPSI][c] = 0.75 * I0.u[HPSI][c] + 0.25 * (I1.u[HPSI][c] + thisdt * I1.E[HPSI][c]);
         }
      double dt2 = hydraulicRHS(I2, t + 0.5 * thisdt, 3);
      if (dt2 < thisdt) { thisdt = 0.9 * dt2; return true; }
      // stage 3
      sponge(I2);
#pragma omp parallel for num_threads(nthreads) schedule(static)
      for (int j = jlo; j < jhi; j++)
         for (int i = ilo; i < ihi; i++) {
            if (!cellActive(i, j)) continue;
            int c = j * NX + i;
            const double third = 1.0 / 3.0, twothird = 2.0 / 3.0;
            I3.u[HU][c] = third * I0.u[HU][c] + twothird * (I2.u[HU][c] + thisdt * I2.E[HU][c]) / (1.0 - thisdt * I2.I[c]);
            I3.u[HV][c] = third * I0.u[HV][c] + twothird * (I2.u[HV][c] + thisdt * I2.E[HV][c]) / (1.0 - thisdt * I2.I[c]);
            double hp_old = -I0.u[BT][c];
            hp_old = hp_old + (I0.u[W][c] - I0.u[B0][c]);
            double hp_new = -I2.u[BT][c];
            hp_new = hp_new + (I2.u[W][c] - I2.u[B0][c]);
            double wu = I0.u[BT][c];
            wu = wu + third * hp_old;
            wu = wu + twothird * hp_new;
            wu = wu + twothird * thisdt * I2.E[W][c];
            wu = wu + I0.u[B0][c];
            I3.u[W][c] = wu;
            I3.u[HPSI][c] = third * I0.u[HPSI][c] + twothird * (I2.u[HPSI][c] + thisdt * I2.E[HPSI][c]);
         }
      hydraulicRHS(I3, nextT, 4);
      // final implicit substep
      sponge(I3);
#pragma omp parallel for num_threads(nthreads) schedule(static)
      for (int j = jlo; j < jhi; j++)
         for (int i = ilo; i < ihi; i++) {
            if (!cellActive(i, j)) continue;
            int c = j * NX + i;
            I3.u[HU][c] = (I3.u[HU][c] - thisdt * thisdt * I3.E[HU][c] * I3.I[c]) / (1.0 + thisdt * thisdt * I3.I[c] * I3.I[c]);
            I3.u[HV][c] = (I3.u[HV][c] - thisdt * thisdt * I3.E[HV][c] * I3.I[c]) / (1.0 + thisdt * thisdt * I3.I[c] * I3.I[c]);
         }
      updateMaxima(nextT);
      return false;
   }

   // ---------------------------------------------------------------- morphodynamics
   // MorphodynamicRHS.f90:72-305
   void morphodynamicRHS(Cont &T) {
      double Hneps = P.heightThreshold;
#pragma omp parallel for num_threads(nthreads) schedule(static)
      for (int j = jlo; j < jhi; j++)
         for (int i = ilo; i < ihi; i++) {
            if (!cellActive(i, j)) continue;
            int c = j * NX + i;
            double HnW = T.u[HN][cidx(i - 1, j)], HnE = T.u[HN][cidx(i + 1, j)];
            double HnS = 0.0, HnN = 0.0;
            if (!oneD) { HnS = T.u[HN][cidx(i, j - 1)]; HnN = T.u[HN][cidx(i, j + 1)]; }
            if (T.u[HN][c] < Hneps || HnW < Hneps || HnE < Hneps || ((!oneD) && (HnS < Hneps || HnN < Hneps))) {
               for (int d = 0; d < 4; d++) T.E[d][c] = 0.0;
               T.EmD[c] = 0.0;
            } else {
               double q[13];
               for (int d = 0; d < 13; d++) q[d] = T.u[d][c];
               double Ero, Depo;
               erosionDeposition(q, Ero, Depo);
               T.EmD[c] = Ero - Depo;
            }
         }
      // BtSourceTerm: vertices of active tiles
      double psib = 1.0 - P.BedPorosity;
      int jn = oneD ? 1 : NYV;
      auto emd = [&](int i, int j) -> double {
         if (!periodic && (i < 0 || i >= NX || j < 0 || j >= NY)) return 0.0;
         return T.EmD[cidx(i, j)];
      };
#pragma omp parallel for num_threads(nthreads) schedule(static)
      for (int vj = (oneD ? 0 : jlo); vj < (oneD ? 1 : jhi + 1); vj++)
         for (int vi = ilo; vi < ihi + 1; vi++) {
            if (periodic && (vi == NX || (!oneD && vj == NY))) continue;
            bool any;
            if (!oneD) any = vertTouchesActive(vi, vj);
            else any = vertTouchesActive1D(vi);
            if (!any) continue;
            int v = vidx(vi, vj);
            if (!oneD) {
               double ax[4] = {T.u[BX][cidx(vi - 1, vj - 1)], T.u[BX][cidx(vi - 1, vj)], T.u[BX][cidx(vi, vj - 1)], T.u[BX][cidx(vi, vj)]};
               double ay[4] = {T.u[BY][cidx(vi - 1, vj - 1)], T.u[BY][cidx(vi - 1, vj)], T.u[BY][cidx(vi, vj - 1)], T.u[BY][cidx(vi, vj)]};
               double dbdx = 0.25 * kahanSum(ax), dbdy = 0.25 * kahanSum(ay);
               double gam = gamma2(dbdx, dbdy);
               double e[4] = {emd(vi - 1, vj - 1), emd(vi - 1, vj), emd(vi, vj - 1), emd(vi, vj)};
               T.EBt[v] = -0.25 * gam / psib * kahanSum(e);
            } else {
               // MorphodynamicRHS.f90:274-302 incl. the one-sided form next to an inactive tile (Q5)
               bool lAct = (periodic || vi - 1 >= 0) && cellActive(vi - 1, 0);
               bool rAct = (periodic || vi < NX) && cellActive(vi, 0);
               if (lAct && rAct) {
                  double dbdx = 0.5 * (T.u[BX][cidx(vi - 1, 0)] + T.u[BX][cidx(vi, 0)]);
                  double gam = gamma2(dbdx, 0.0);
                  T.EBt[v] = -0.5 * gam * (T.EmD[cidx(vi - 1, 0)] + T.EmD[cidx(vi, 0)]) / psib;
               } else {
                  int cc = lAct ? cidx(vi - 1, 0) : cidx(vi, 0);
                  double dbdx = 0.5 * T.u[BX][cc];
                  double gam = gamma2(dbdx, 0.0);
                  T.EBt[v] = -0.5 * gam * T.EmD[cc] / psib;
               }
            }
         }
   }
   bool vertTouchesActive(int vi, int vj) const {
      for (int dj = -1; dj <= 0; dj++)
         for (int di = -1; di <= 0; di++) {
            int i = vi + di, j = vj + dj;
            if (!periodic && (i < 0 || i >= NX || j < 0 || j >= NY)) continue;
            if (cellActive(i, j)) return true;
         }
      return false;
   }
   bool vertTouchesActive1D(int vi) const {
      for (int di = -1; di <= 0; di++) {
         int i = vi + di;
         if (!periodic && (i < 0 || i >= NX)) continue;
         if (cellActive(i, 0)) return true;
      }
      return false;
   }

   // one RK stage of the bed + linear updates (TimeStepper.f90:569-610 / 615-655 / 660-699)
   void morphoStage(Cont &Tn, const Cont &Tk, double a0, double a1, double thisdt) {
      int jn = oneD ? 1 : NYV;
#pragma omp parallel for num_threads(nthreads) schedule(static)
      for (int vj = (oneD ? 0 : jlo); vj < (oneD ? 1 : jhi + 1); vj++)
         for (int vi = ilo; vi < ihi + 1; vi++) {
            if (periodic && (vi == NX || (!oneD && vj == NY))) continue;
            bool any = oneD ? vertTouchesActive1D(vi) : vertTouchesActive(vi, vj);
            if (!any) continue;
            int v = vidx(vi, vj);
            double val;
            if (a0 == 0.0) val = I0.btv[v] + thisdt * I0.EBt[v];
            else val = a0 * I0.btv[v] + a1 * (Tk.btv[v] + thisdt * Tk.EBt[v]);
            Tn.btv[v] = std::max(-P.EroDepth, val);
         }
#pragma omp parallel for num_threads(nthreads) schedule(static)
      for (int j = jlo; j < jhi; j++)
         for (int i = ilo; i < ihi; i++) {
            if (!cellActive(i, j)) continue;
            int c = j * NX + i;
            centreTopo(Tn, i, j, true);
            double gamold = gammaC(I0, c), gamnew = gammaC(Tn, c);
            double Hn_old = computeHn(I0.u[W][c], I0.u[B0][c], I0.u[BT][c], gamold);
            double db = Tn.u[BT][c] - I0.u[BT][c];
            double w = -db / gamnew / gamnew;
            w = w + Tn.u[BT][c];
            w = w + Hn_old * gamold / gamnew / gamnew;
            w = w + Tn.u[B0][c];
            Tn.u[W][c] = w;
            double Hnpsi_old = I0.u[HPSI][c];
            double Hnpsi = -(1.0 - P.BedPorosity) * db / gamnew;
            Hnpsi = Hnpsi + Hnpsi_old * gamold / gamnew;
            Tn.u[HPSI][c] = Hnpsi;
            desingularise(Tn, c, false);
         }
   }

   // Redistribute.f90:158-198
   void excessDeposition(int c, double &excess, double &deltaBt, double &psiold) const {
      double gamold = gammaC(I0, c);
      double Hn_old = computeHn(I0.u[W][c], I0.u[B0][c], I0.u[BT][c], gamold);
      deltaBt = I3.u[BT][c] - I0.u[BT][c];
      psiold = I0.u[PSI][c];
      excess = -(I0.u[HPSI][c] * gamold / (1.0 - P.BedPorosity) - deltaBt);
      excess = std::max(excess, -(Hn_old * gamold - deltaBt));
   }

   // Redistribute.f90:249-475
   bool redistributeCell(int i, int j, double corr) {
      int c = j * NX + i;
      double b_diff[4] = {0, 0, 0, 0}, sum_b_diff = 0.0;
      int depv[4];
      int N = 0;
      int vs[4] = {vidx(i, j), vidx(i + 1, j), oneD ? -1 : vidx(i, j + 1), oneD ? -1 : vidx(i + 1, j + 1)};
      for (int k = 0; k < (oneD ? 2 : 4); k++) {
         b_diff[N] = I3.btv[vs[k]] - I0.btv[vs[k]];
         if (b_diff[N] > 0.0) { depv[N] = vs[k]; sum_b_diff = sum_b_diff + b_diff[N]; N++; }
      }
      if (N == 0) return false;
      double delta = 4.0 * corr / sum_b_diff;
      double Hnold = I0.u[HN][c];
      double gamold = gammaC(I0, c);
      double db = I3.u[BT][c] - I0.u[BT][c];
      double tol = EPS * I3.u[W][c] * 10.0;
      double adjustment = 0.0;
      { double a[3] = {Hnold * gamold, -db, corr};
        double discrepancy = kahanSum(a);
        if (std::fabs(discrepancy) < tol) {
           double b[3] = {tol, -Hnold * gamold, db};
           adjustment = kahanSum(b);
           adjustment = adjustment * (4.0 / sum_b_diff);
           adjustment = adjustment - delta;
           adjustment = std::max(adjustment, 0.0);
        } }
      double Hg = I0.u[HPSI][c] * gamold / (1.0 - P.BedPorosity);
      tol = EPS * 10.0;  // epsilon(Hngampsi_o_psib_old) * 10
      { double a[3] = {Hg, -db, corr};
        double discrepancy = kahanSum(a);
        if (std::fabs(discrepancy) < tol) {
           double b[3] = {tol, -Hg, db};
           double adj = kahanSum(b);
           adj = adj * (4.0 / sum_b_diff);
           adj = adj - delta;
           adjustment = std::max(adjustment, adj);
        } }
      delta = delta + adjustment;
      for (int k = 0; k < N; k++) I3.btv[depv[k]] = I3.btv[depv[k]] - delta * b_diff[k];
      // refresh the 3^D surrounding cells (only cells of active tiles carry state)
      for (int ci = i - 1; ci <= i + 1; ci++)
         for (int cj = (oneD ? 0 : j - 1); cj <= (oneD ? 0 : j + 1); cj++) {
            if (!periodic && (ci < 0 || ci >= NX || cj < 0 || cj >= NY)) continue;
            int wi = wrapi(ci), wj = wrapj(cj);
            if (!cellActive(wi, wj)) continue;
            int cc = wj * NX + wi;
            centreTopo(I3, wi, wj, false);
            double dbc = I3.u[BT][cc] - I0.u[BT][cc];
            double go = gammaC(I0, cc), gn = gammaC(I3, cc);
            double Ho = I0.u[HN][cc];
            double w;
            if (!oneD) {
               w = I3.u[BT][cc];
               w = w + (Ho * go / gn - dbc / gn) / gn;
               w = w + I3.u[B0][cc];
            } else {
               w = -dbc / gn / gn;
               w = w + Ho * go / gn / gn;
               w = w + I3.u[BT][cc];
               w = w + I3.u[B0][cc];
            }
            I3.u[W][cc] = w;
            double Hpo = I0.u[HPSI][cc];
            I3.u[HPSI][cc] = Hpo * go / gn - (1.0 - P.BedPorosity) * dbc / gn;
         }
      return true;
   }

   // TimeStepper.f90:532-781
   bool morphodynamicTimeStepper(double &thisdt) {
      bool refine = false;
      morphodynamicRHS(I0);
      morphoStage(I1, I0, 0.0, 1.0, thisdt);
      morphodynamicRHS(I1);
      morphoStage(I2, I1, 0.75, 0.25, thisdt);
      morphodynamicRHS(I2);
      morphoStage(I3, I2, 1.0 / 3.0, 2.0 / 3.0, thisdt);
      // checks (TimeStepper.f90:709-753); list kept sorted ascending, stable (Redistribute.f90:69-101)
      struct Node { double ex; int i, j; };
      vector<Node> list;
      for (int id : activeList) {
         int i0, i1, j0, j1;
         forTileCells(id - 1, i0, i1, j0, j1);
         for (int j = j0; j < j1; j++)
            for (int i = i0; i < i1; i++) {
               int c = j * NX + i;
               double gamold = gammaC(I0, c), gamnew = gammaC(I3, c);
               double Hn_old = computeHn(I0.u[W][c], I0.u[B0][c], I0.u[BT][c], gamold);
               double Hn_new = computeHn(I3.u[W][c], I3.u[B0][c], I3.u[BT][c], gamnew);
               double ex, dBt, psiold;
               excessDeposition(c, ex, dBt, psiold);
               if ((!refine) && ex > EPS && dBt > EPS && psiold > -EPS) {
                  if (Hn_old < P.EroCriticalHeight || std::fabs(Hn_new - Hn_old) < P.EroCriticalHeight) {
                     Node nd{ex, i, j};
                     auto it = std::upper_bound(list.begin(), list.end(), nd, [](const Node &a, const Node &b) { return a.ex < b.ex; });
                     list.insert(it, nd);
                  } else {
                     refine = true;
                     break;
                  }
                  continue;
               }
               if (Hn_old < P.EroCriticalHeight) continue;
               double rel = std::fabs(Hn_new - Hn_old) / std::fabs(Hn_old);
               if (rel > 0.1) { refine = true; break; }
            }
      }
      if (!list.empty() && !refine) {
         for (const Node &nd : list) {
            double ex, dBt, psiold;
            excessDeposition(nd.j * NX + nd.i, ex, dBt, psiold);
            if (ex > EPS) {
               nRedistributed++;   // test diagnostics only (kor_debug_redistributed)
               if (!redistributeCell(nd.i, nd.j, ex)) { refine = true; break; }
            }
         }
      }
      if (refine) { thisdt = 0.5 * thisdt; return true; }
      return false;
   }

   // TimeStepper.f90:281-305
   double nextFluxSeriesTime(double tt) const {
      double nextT = HUGE_D, tdiff = HUGE_D;
      for (const Source &S : src)
         for (double ts : S.time) {
            double tmp = ts - tt;
            if (tmp > 0.0 && tmp < tdiff) { tdiff = tmp; nextT = ts; }
         }
      return nextT;
   }

   // TimeStepper.f90:116-277
   int integrateTo(double tend, int64_t maxSteps, kgpu_step_info *info) {
      int64_t done = 0;
      double dt_hydro = dtgrid;
      bool integrating = tend > t;
      while (integrating) {
         int rc = checkIfNearBoundaries();
         if (rc) return rc;
         initialiseTimeSteppingArrays();
         double advised = hydraulicRHS(C, t, 1);
         double tmax = std::min(tend, nextFluxSeriesTime(t));
         if (!P.MorphodynamicsOn) advised = std::min(advised, tmax - t);
         else advised = std::min(advised, 0.5 * (tmax - t));
         dtgrid = advised;
         dt_hydro = advised;
         t0 = t;
         int guard = 0;
         while (true) {
            if (++guard > 200 || !(dt_hydro > 0.0) || !std::isfinite(dt_hydro)) { err = "time step underflow"; return KGPU_ERR_DT; }
            copySolution(C, I0);
            copyMutableTopo(C, I1);
            copyMutableTopo(C, I2);
            copySolution(C, I3);
            double nextT;
            bool refine = hydraulicTimeStepper(dt_hydro, nextT);
            if (refine) { nrefines++; continue; }
            if (!P.MorphodynamicsOn) break;
            copySolution(I3, I0);
            copySolution(I3, I1);
            copySolution(I3, I2);
            double dt_morpho = 2.0 * dt_hydro;
            refine = morphodynamicTimeStepper(dt_morpho);
            if (refine) {
               dt_hydro = 0.5 * dt_morpho;
               dtgrid = dt_hydro;
               nrefines++;
               continue;
            }
            copySolution(I3, I0);
            copyMutableTopo(I3, I1);
            copyMutableTopo(I3, I2);
            t = t + dt_hydro;
            advised = hydraulicRHS(I0, t, 1);
            if (advised < dt_hydro) {
               if ((dt_hydro - advised) / dt_hydro < 0.1) dt_hydro = 0.9 * dt_hydro;
               else dt_hydro = advised;
               dtgrid = dt_hydro;
               t = t0;
               nrefines++;
               continue;
            }
            refine = hydraulicTimeStepper(dt_hydro, nextT);
            if (!refine) break;
            t = t0;
            nrefines++;
         }
         copySolution(I3, C);
         if (!P.MorphodynamicsOn) t = t0 + dt_hydro;
         else t = t0 + 2.0 * dt_hydro;
         nsteps++;
         done++;
         if (t >= tend) integrating = false;
         if (maxSteps > 0 && done >= maxSteps) integrating = false;
      }
      if (info) {
         info->t = t; info->dt_last = dt_hydro; info->nsteps = nsteps; info->nrefines = nrefines; info->ntiles_added = ntilesAdded;
      }
      return 0;
   }

   // ---------------------------------------------------------------- setup
   void allocCont(Cont &T) {
      size_t nc = (size_t)NX * NY, nv = (size_t)NXV * NYV;
      for (int d = 0; d < 13; d++) T.u[d].assign(nc, 0.0);
      T.u[RHO].assign(nc, P.rhow);
      T.btv.assign(nv, 0.0);
      for (int d = 0; d < 4; d++) T.E[d].assign(nc, 0.0);
      T.I.assign(nc, 0.0);
      T.EBt.assign(nv, 0.0);
      T.EmD.assign(nc, 0.0);
   }
   int init(const kgpu_params *p) {
      P = *p;
      src.clear();
      for (int s = 0; s < P.n_sources; s++) {
         Source S;
         const kgpu_source &k = p->sources[s];
         S.x = k.x; S.y = k.y; S.radius = k.radius; S.numCells = k.num_cells_in_src;
         S.time.assign(k.time, k.time + k.n_series);
         S.flux.assign(k.flux, k.flux + k.n_series);
         S.psi.assign(k.psi, k.psi + k.n_series);
         src.push_back(S);
      }
      P.sources = nullptr;
      nX = P.nXpertile; nY = P.nYpertile; nXt = P.nXtiles; nYt = P.nYtiles;
      NX = nX * nXt; NY = nY * nYt; nTiles = nXt * nYt;
      oneD = P.isOneD != 0; periodic = P.bcs == KGPU_BC_PERIODIC; geom = P.geometric_factors != 0;
      NXV = NX + 1; NYV = oneD ? 1 : NY + 1;
      dx = P.deltaX; dy = P.deltaY; dxR = 1.0 / dx; dyR = 1.0 / dy;
      buildTables();
      tstate.assign(nTiles, 0); hasSource.assign(nTiles, 0); loaded.assign(nTiles, 0);
      size_t nc = (size_t)NX * NY, nv = (size_t)NXV * NYV;
      b0v.assign(nv, 0.0);
      allocCont(C); allocCont(I0); allocCont(I1); allocCont(I2); allocCont(I3);
      for (int d = 0; d < 9; d++) { limX[d].assign(nc, 0.0); limY[d].assign(nc, 0.0); }
      size_t nfx = (size_t)(NX + 1) * NY, nfy = (size_t)NX * (NY + 1);
      for (int d = 0; d < 13; d++) { pX[d].assign(nfx, 0.0); mX[d].assign(nfx, 0.0); pY[d].assign(nfy, 0.0); mY[d].assign(nfy, 0.0); }
      for (int d = 0; d < 4; d++) { hX[d].assign(nfx, 0.0); hY[d].assign(nfy, 0.0); }
      gX.assign(nfx, 0.0); gY.assign(nfy, 0.0);
      for (int d = 0; d < 2; d++) { dfX[d].assign(nfx, 0.0); dfY[d].assign(nfy, 0.0); }
      for (int k = 0; k < 2; k++) { Hnmax[k].assign(nc, 0.0); umax[k].assign(nc, 0.0); emax[k].assign(nc, 0.0); dmax[k].assign(nc, 0.0); psimax[k].assign(nc, 0.0); }
      tfirst.assign(nc, -1.0);
      HnSeed.assign(nc, 0.0);
      t = P.tstart; t0 = t; dtgrid = 1.0e-5;
      return 0;
   }
};

}  // namespace

// ============================================================================ C API
extern "C" {

typedef struct Oracle kor_handle;

int kor_create(const kgpu_params *p, kor_handle **h) {
   if (!p || !h || p->struct_bytes != (int32_t)sizeof(kgpu_params)) return KGPU_ERR_ARG;
   Oracle *o = new Oracle();
   int rc = o->init(p);
   if (rc) { delete o; return rc; }
   *h = o;
   return 0;
}
int kor_destroy(kor_handle *h) { delete h; return 0; }
// test diagnostics: how many RedistributeCell calls the run has made so far (lets a test assert that the
// redistribution path was really exercised)
int64_t kor_debug_redistributed(const kor_handle *h) { return h ? h->nRedistributed : 0; }
const char *kor_last_error(const kor_handle *h) { return h ? h->err.c_str() : "null handle"; }
int kor_set_threads(kor_handle *h, int n) { h->nthreads = n < 1 ? 1 : n; return 0; }

int kor_upload_tile(kor_handle *h, int32_t tile_id, const double *u13, const double *b0_vertices,
                    const double *bt_vertices, const double *maxima, const double *tfirst, int32_t contains_source) {
   Oracle &o = *h;
   int t0_ = tile_id - 1;
   if (t0_ < 0 || t0_ >= o.nTiles || !u13) return KGPU_ERR_ARG;
   if (b0_vertices) { int rc = o.loadHeights(t0_, b0_vertices); if (rc) return rc; }
   int rc = o.addTile(t0_, true);
   if (rc) return rc;
   if (o.tstate[t0_] != 2) { o.err = "tile cannot be active (domain edge)"; return KGPU_ERR_ARG; }
   o.ntilesAdded--;
   o.hasSource[t0_] = contains_source ? 1 : 0;
   int i0, i1, j0, j1;
   o.forTileCells(t0_, i0, i1, j0, j1);
   int nX = o.nX, nY = o.nY;
   if (bt_vertices) {
      int jn = o.oneD ? 1 : nY + 1;
      for (int lj = 0; lj < jn; lj++)
         for (int li = 0; li <= nX; li++) o.C.btv[o.vidx(i0 + li, j0 + lj)] = bt_vertices[(size_t)lj * (nX + 1) + li];
      o.centreTopoTile(o.C, t0_, false);
   }
   for (int lj = 0; lj < nY; lj++)
      for (int li = 0; li < nX; li++) {
         int c = (j0 + lj) * o.NX + i0 + li;
         const double *q = u13 + ((size_t)lj * nX + li) * 13;
         for (int d = 0; d < 9; d++) o.C.u[d][c] = q[d];
         o.HnSeed[c] = q[HN];
         if (maxima) {
            size_t blk = (size_t)nX * nY * 2, k1 = (size_t)lj * nX + li, k2 = (size_t)nX * nY + k1;
            o.Hnmax[0][c] = maxima[0 * blk + k1]; o.Hnmax[1][c] = maxima[0 * blk + k2];
            o.umax[0][c] = maxima[1 * blk + k1];  o.umax[1][c] = maxima[1 * blk + k2];
            o.emax[0][c] = maxima[2 * blk + k1];  o.emax[1][c] = maxima[2 * blk + k2];
            o.dmax[0][c] = maxima[3 * blk + k1];  o.dmax[1][c] = maxima[3 * blk + k2];
            o.psimax[0][c] = maxima[4 * blk + k1]; o.psimax[1][c] = maxima[4 * blk + k2];
         }
         if (tfirst) o.tfirst[c] = tfirst[(size_t)lj * nX + li];
      }
   return 0;
}

int kor_upload_domain(kor_handle *h, const double *q4, const double *b0_vertices, const double *bt_vertices) {
   Oracle &o = *h;
   if (!o.periodic || !q4 || !b0_vertices) { o.err = "upload_domain needs periodic bcs, q4 and b0"; return KGPU_ERR_ARG; }
   size_t nc = (size_t)o.NX * o.NY;
   int jn = o.oneD ? 1 : o.NY + 1;
   for (int vj = 0; vj < jn; vj++)
      for (int vi = 0; vi <= o.NX; vi++) {
         if (vi == o.NX || (!o.oneD && vj == o.NY)) continue;  // aliases of index 0
         size_t k = (size_t)vj * (o.NX + 1) + vi;
         o.b0v[o.vidx(vi, vj)] = b0_vertices[k];
         if (bt_vertices) o.C.btv[o.vidx(vi, vj)] = bt_vertices[k];
      }
   o.activeList.clear(); o.ghostList.clear();
   for (int t0_ = 0; t0_ < o.nTiles; t0_++) { o.tstate[t0_] = 2; o.loaded[t0_] = 1; o.activeList.push_back(t0_ + 1); }
   // containsSource: a tile with a cell centre inside a source disc, <= as at load (SetSources.f90:367-372)
   std::fill(o.hasSource.begin(), o.hasSource.end(), 0);
   for (const Source &S : o.src)
      for (int j = 0; j < o.NY; j++)
         for (int i = 0; i < o.NX; i++) {
            double x = o.cellX(i), y = o.cellY(j);
            double R2 = o.oneD ? (x - S.x) * (x - S.x) : (x - S.x) * (x - S.x) + (y - S.y) * (y - S.y);
            if (R2 <= S.radius * S.radius) o.hasSource[o.tileOfCell(i, j)] = 1;
         }
   o.updateBounds();
   for (int j = 0; j < o.NY; j++)
      for (int i = 0; i < o.NX; i++) {
         int c = j * o.NX + i;
         o.centreTopo(o.C, i, j, true);
         for (int d = 0; d < 4; d++) o.C.u[d][c] = q4[d * nc + c];
         o.desingularise(o.C, c, true);
         o.HnSeed[c] = o.C.u[HN][c];
      }
   return 0;
}

int kor_integrate_to(kor_handle *h, double tend, int64_t max_steps, kgpu_step_info *info) {
   return h->integrateTo(tend, max_steps, info);
}

int kor_active_tiles(kor_handle *h, int32_t *n, int32_t *ids) {
   *n = (int32_t)h->activeList.size();
   if (ids) for (size_t k = 0; k < h->activeList.size(); k++) ids[k] = h->activeList[k];
   return 0;
}
int kor_ghost_tiles(kor_handle *h, int32_t *n, int32_t *ids) {
   *n = (int32_t)h->ghostList.size();
   if (ids) for (size_t k = 0; k < h->ghostList.size(); k++) ids[k] = h->ghostList[k];
   return 0;
}

int kor_download_tile(kor_handle *h, int32_t tile_id, double *u13, double *b0_vertices, double *bt_vertices,
                      double *maxima, double *tfirst) {
   Oracle &o = *h;
   int t0_ = tile_id - 1;
   if (t0_ < 0 || t0_ >= o.nTiles) return KGPU_ERR_ARG;
   int i0, i1, j0, j1;
   o.forTileCells(t0_, i0, i1, j0, j1);
   int nX = o.nX, nY = o.nY;
   for (int lj = 0; lj < nY; lj++)
      for (int li = 0; li < nX; li++) {
         int c = (j0 + lj) * o.NX + i0 + li;
         if (u13) for (int d = 0; d < 13; d++) u13[((size_t)lj * nX + li) * 13 + d] = o.C.u[d][c];
         if (maxima) {
            size_t blk = (size_t)nX * nY * 2, k1 = (size_t)lj * nX + li, k2 = (size_t)nX * nY + k1;
            maxima[0 * blk + k1] = o.Hnmax[0][c]; maxima[0 * blk + k2] = o.Hnmax[1][c];
            maxima[1 * blk + k1] = o.umax[0][c];  maxima[1 * blk + k2] = o.umax[1][c];
            maxima[2 * blk + k1] = o.emax[0][c];  maxima[2 * blk + k2] = o.emax[1][c];
            maxima[3 * blk + k1] = o.dmax[0][c];  maxima[3 * blk + k2] = o.dmax[1][c];
            maxima[4 * blk + k1] = o.psimax[0][c]; maxima[4 * blk + k2] = o.psimax[1][c];
         }
         if (tfirst) tfirst[(size_t)lj * nX + li] = o.tfirst[c];
      }
   int jn = o.oneD ? 1 : nY + 1;
   for (int lj = 0; lj < jn; lj++)
      for (int li = 0; li <= nX; li++) {
         int v = o.vidx(i0 + li, j0 + lj);
         if (b0_vertices) b0_vertices[(size_t)lj * (nX + 1) + li] = o.b0v[v];
         if (bt_vertices) bt_vertices[(size_t)lj * (nX + 1) + li] = o.C.btv[v];
      }
   return 0;
}

int kor_download_domain(kor_handle *h, double *q4, double *bt_vertices) {
   Oracle &o = *h;
   size_t nc = (size_t)o.NX * o.NY;
   if (q4) for (int d = 0; d < 4; d++) std::memcpy(q4 + d * nc, o.C.u[d].data(), nc * sizeof(double));
   if (bt_vertices) {
      int jn = o.oneD ? 1 : o.NY + 1;
      for (int vj = 0; vj < jn; vj++)
         for (int vi = 0; vi <= o.NX; vi++) bt_vertices[(size_t)vj * (o.NX + 1) + vi] = o.C.btv[o.vidx(vi, vj)];
   }
   return 0;
}

// Whole-domain views for parity tests: any of the 13 cell fields, flat (NX*NY).
int kor_download_field(kor_handle *h, int32_t d, double *out) {
   Oracle &o = *h;
   if (d < 0 || d >= 13) return KGPU_ERR_ARG;
   std::memcpy(out, o.C.u[d].data(), (size_t)o.NX * o.NY * sizeof(double));
   return 0;
}
// One evaluation of CalculateHydraulicRHS on the current state (substep as given):
// E4 = ddtExplicit planes, I = ddtImplicit (momenta), dt = advisedTimeStep.
int kor_debug_rhs(kor_handle *h, int32_t substep, double *E4, double *I, double *dt) {
   Oracle &o = *h;
   size_t nc = (size_t)o.NX * o.NY;
   double save = o.dtgrid;
   double adv = o.hydraulicRHS(o.C, o.t, substep);
   o.dtgrid = save;
   if (E4) for (int d = 0; d < 4; d++) std::memcpy(E4 + d * nc, o.C.E[d].data(), nc * sizeof(double));
   if (I) std::memcpy(I, o.C.I.data(), nc * sizeof(double));
   if (dt) *dt = adv;
   return 0;
}
const char *kor_version(void) { return "kestrel-oracle 0.1 (restates jakelangham/kestrel v1.1.1)"; }

}  // extern "C"
