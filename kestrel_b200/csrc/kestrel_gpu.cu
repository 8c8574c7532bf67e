// kestrel_gpu.cu -- host side of libkestrel_gpu: the C-ABI of include/kestrel_gpu.h.
//
// Replaces the body of IntegrateTo (TimeStepper.f90:116-277).  Field data never leaves
// the device between kgpu_upload_* and kgpu_download_*; dt selection, the refine test of
// every Runge-Kutta stage and the rollback flag live in a device-resident control block,
// so one step costs one host synchronisation.  q0 is never overwritten inside a step:
// a rolled-back attempt restarts from the retained stage-1 RHS (three state buffers are
// rotated by pointer instead of the reference's five deep-copied tile containers,
// TimeStepper.f90:785-875).
//
// There is no CPU execution path in this library.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

#include <cuda.h>

#include "../../include/kestrel_gpu_debug.h"
#include "kgpu_comm.cuh"
#include "kgpu_hydro.cuh"
#include "kgpu_redist_tables.hpp"
#include "kgpu_morpho.cuh"
#include "kgpu_tile_table.hpp"
#include "kgpu_tiles.cuh"

using namespace kgpu;

namespace kgpu {
// contracted-arithmetic instantiations live in kestrel_stage_fast.cu (compiled with -fmad=true)
void launch_stage_fast(bool oneD, bool hasBt, bool mm2, bool spec, dim3 grid, cudaStream_t s, const DevParams &P, const StageArgs &a);
void stage_fast_set_attributes();
}  // namespace kgpu

#define CUDA_TRY(h, call)                                                                         \
   do {                                                                                           \
      cudaError_t e_ = (call);                                                                    \
      if (e_ != cudaSuccess) {                                                                    \
         (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
         return KGPU_ERR_CUDA;                                                                    \
      }                                                                                           \
   } while (0)

namespace {
constexpr int BX2 = 32, BY2 = KGPU_STAGE_BY2;   // 2-D stage tile: 224 cells, 487 faces = 2 passes of 256 threads
constexpr int NTHREADS = 256;                   // elementwise kernels over a stage tile
constexpr int STAGE_THREADS = KGPU_STAGE_THREADS;
constexpr int BX1 = 128, BY1 = 1;  // 1-D stage tile
const double HUGE_D = std::numeric_limits<double>::max();
}  // namespace

// 2-D block decomposition over NCCL ranks (one process per GPU)
struct kgpu_comm {
   bool active = false;
   int rank = 0, size = 1, px = 1, py = 1, rx = 0, ry = 0;
   int west = -1, east = -1, south = -1, north = -1;  // neighbour ranks (-1: none / local wrap)
   void *nccl = nullptr;
   double *sendBuf[4] = {}, *recvBuf[4] = {};
   cudaStream_t stream = nullptr;       // communication stream
   cudaEvent_t evBoundary = nullptr, evHalo = nullptr;
};

// device buffers of the redistribution walk across ranks (kgpu_morpho_host.inl), grown on demand
struct RedistGlobalBufs {
   size_t cap = 0;                 // entries (all ranks) the buffers hold
   RedistEntry *dEntries = nullptr;
   double *dPatch = nullptr, *dSend = nullptr;
   int *dCounts = nullptr, *dVslot = nullptr, *dCslot = nullptr, *dVkey = nullptr, *dVbase = nullptr, *dCkey = nullptr, *dCbase = nullptr;
   int sendCap = 0;
};

struct kgpu_handle {
   kgpu_params P;
   DevParams D;
   std::vector<DevSource> src;
   std::vector<double> srcPool;   // time | flux | psi series of every source (DevSource::off)
   std::string err;
   int dev = 0;
   cudaStream_t stream = nullptr;

   int NX = 0, NY = 0, nX = 0, nY = 0, nXt = 0, nYt = 0, nTiles = 0, pitch = 0, rows = 0;
   bool oneD = false, periodic = false, morpho = false;
   bool globalPeriodic = false;  // bcs = periodic on the whole domain; `periodic` = this device wraps onto itself
   kgpu_comm comm;
   int gtx0 = 0, gty0 = 0, gnXt = 0, gnYt = 0;
   int2 *d_blockBoundary = nullptr, *d_blockInterior = nullptr;
   int nBoundary = 0, nInterior = 0;
   size_t fieldElems = 0;

   // device fields
   double *S[5][4] = {};       // rotating primary states: 3 (hydraulic) or 5 (Strang split)
   int nStates = 3;
   int i0 = 0, ia = 1, ib = 2; // indices of q0 / stage buffer A / stage buffer B
   int ic = 3, id = 4;         // morphodynamics: state after M, second H result
   bool e0Valid = false;       // E0/I0 hold the substep-1 RHS of S[i0]
   double *Um = nullptr, *Vm = nullptr;  // velocities frozen over M
   double *mc[2][4] = {};      // morphodynamic stages: centre bt, bx, by and clamped Hn of the stage (ping-pong)
   double *mcHn0 = nullptr;    // clamped Hn of the state M starts from
   double *E0[4] = {}, *I0 = nullptr;
   double *b0v = nullptr;
   double *btv[4] = {};        // morphodynamics: bed change at vertices for q0 and the three stages
   int bt0 = 0, bt1 = 1, bt2 = 2, bt3 = 3;
   double *EBt = nullptr, *EmD = nullptr;
   double *mx[11] = {};
   TopoPlanes topo = {};       // precomputed cell / face topography for the stage kernel
   TmaDesc *d_maps = nullptr;  // tensor maps of the state and topography planes (TmaSlot)
   int prefetchDistance = 0;   // L2 prefetch distance of the stage kernel in CTAs (one resident wave)
   int tune = 0;               // StageArgs::tune bits; bit 3 here: 2-D grid without the block list when every block is listed
   bool useSpec = true;        // take the (geometric factors, nu == 0) instantiation when the run allows it; KGPU_TUNE bit 6 turns it off
   int nbxAll = 0, nbyAll = 0; // CTA tiles per row / column of the local domain
   int topoBtIdx = -1;         // which bt array the planes were computed from (-1: stale)
   bool havePre = false;       // S[ia] holds q3 before the implicit correction (quirk Q2)

   uint8_t *d_tileMask = nullptr, *d_tileSource = nullptr;
   int2 *d_blockList = nullptr;
   int nBlocks = 0;
   int2 *d_blockListM = nullptr;   // tiles of the fused morphodynamic stage kernel (BX2 x MORPHO_STAGE_BY cells)
   int nBlocksM = 0;
   Ctrl *d_ctrl = nullptr, *h_ctrl = nullptr;
   DevSource *d_sources = nullptr;
   double *d_srcPool = nullptr;
   double *d_stage = nullptr, *h_stage = nullptr;
   size_t stageElems = 0;
   int *d_tileList = nullptr, *d_flags = nullptr, *h_flags = nullptr;
   RedistEntry *d_redist = nullptr, *h_redist = nullptr;
   int redistCap = 0;
   int64_t nRedistCells = 0, nRedistGrows = 0;   // cells handed to RedistributeGrid / enlargements of its list buffer
   bool debugGlobalWalk = false;                 // kgpu_debug_global_walk (kestrel_gpu_debug.h)
   bool debugSequentialWalk = false;             // kgpu_debug_sequential_walk: one thread walks the list, as the reference does
   int morphoFusion = 0;   // morphodynamic Runge-Kutta stage on a single device: 0 = E - D, bed, cells as three kernels (fastest
                           // measured, and what the decomposed runs use), 1 = E - D, then bed + cells fused, 2 = one kernel
                           // (KGPU_TUNE bits 7 / 8, kgpu_debug_morpho_fusion; DESIGN.md section 5 has the timings)
   int *d_rankMap = nullptr;                     // redistribution wave: list position per cell + the block ticket

   // host tile bookkeeping (UpdateTiles.f90)
   std::vector<int> tstate;  // 0 untouched, 1 ghost, 2 active
   std::vector<char> hasSource, loaded;
   std::vector<int> activeList, ghostList;  // 1-based ids
   std::vector<int> seedFlags;
   bool firstScan = true, masksDirty = true;
   // decomposed run on a non-periodic domain: dynamic tiles through the replicated table (kgpu_dyn_host.inl)
   bool dyn = false;
   kgpu::TileTable gt;
   size_t opsDone = 0;               // entries of gt.ops already executed on this device
   std::vector<int> gSeed, gSource;  // per tile of the whole grid: seed flags of the first scan, containsSource
   int *d_gflags = nullptr, *h_gflags = nullptr;

   double t = 0.0, dtgrid = 1.0e-5;
   int64_t nsteps = 0, nrefines = 0, ntilesAdded = 0, launches = 0;
   // RHS kernel timing
   cudaEvent_t evA = nullptr, evB = nullptr;
   bool timeRhs = false;
   double rhsMs = 0.0;
   int64_t rhsLaunches = 0;

   // asynchronous output gather: device snapshot + copy stream
   double *snap[5] = {};
   cudaStream_t copyStream = nullptr;
   cudaEvent_t evSnap = nullptr, evCopied = nullptr;
   bool outputPending = false;
   RedistGlobalBufs rg;
   TopogFn topogFn = {-1, 0, {0, 0, 0, 0, 0, 0, 0, 0}};   // analytic topography evaluated on the device (func < 0: heights callback)
   RasterDesc raster = {nullptr, 0, 0, 0, 0, 0, 0, 0, 0};  // DEM section resampled on the device (elev == nullptr: not set)
   double *d_raster = nullptr;

   // (a block of a dynamic decomposed run is never treated as all-active: the tiles around it need not be)
   bool allActive() const { return !dyn && (int)activeList.size() == nTiles; }
   StatePtrs sp(int k) const { StatePtrs s; for (int d = 0; d < 4; d++) s.q[d] = S[k][d]; return s; }
   MaximaPtrs mp() const {
      MaximaPtrs m;
      m.Hnmax = mx[0]; m.HnmaxT = mx[1]; m.umax = mx[2]; m.umaxT = mx[3]; m.emax = mx[4]; m.emaxT = mx[5];
      m.dmax = mx[6]; m.dmaxT = mx[7]; m.psimax = mx[8]; m.psimaxT = mx[9]; m.tfirst = mx[10];
      return m;
   }
};

// =========================================================================== small kernels
__global__ void fill_kernel(double *p, size_t n, double v) {
   size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (k < n) p[k] = v;
}
__global__ void ctrl_begin_kernel(Ctrl *c, double t) {
   c->t = t;
   c->cflBits[0] = 0x7FEFFFFFFFFFFFFFull;
}
__global__ void ctrl_set_dt_kernel(Ctrl *c, double dt) {
   c->dt = dt;
   c->failed = 0;
   c->cflBits[1] = c->cflBits[2] = c->cflBits[3] = 0x7FEFFFFFFFFFFFFFull;
}

// =========================================================================== helpers
static int roundUp(int a, int b) { return (a + b - 1) / b * b; }

static void tileXY(const kgpu_handle *h, int t0, int &tx, int &ty) { tx = t0 % h->nXt; ty = t0 / h->nXt; }
// Tile ids across the C-ABI are ids of the WHOLE tile grid (1-based, tx fastest), also in a decomposed run, where
// the handle stores its own block of nXt x nYt tiles starting at (gtx0, gty0).
static int globalTileId(const kgpu_handle *h, int t0) { return (h->gty0 + t0 / h->nXt) * h->gnXt + h->gtx0 + t0 % h->nXt + 1; }
static int localTile0(const kgpu_handle *h, int tile_id) {   // -1: out of range, -2: another rank's tile
   int g0 = tile_id - 1;
   if (g0 < 0 || g0 >= h->gnXt * h->gnYt) return -1;
   int tx = g0 % h->gnXt - h->gtx0, ty = g0 / h->gnXt - h->gty0;
   if (tx < 0 || tx >= h->nXt || ty < 0 || ty >= h->nYt) return -2;
   return ty * h->nXt + tx;
}
static int tileW(const kgpu_handle *h, int t0) {
   int tx, ty; tileXY(h, t0, tx, ty);
   if (tx == 0) return h->periodic ? (h->nXt - 1) + ty * h->nXt : -1;
   return t0 - 1;
}
static int tileE(const kgpu_handle *h, int t0) {
   int tx, ty; tileXY(h, t0, tx, ty);
   if (tx == h->nXt - 1) return h->periodic ? ty * h->nXt : -1;
   return t0 + 1;
}
static int tileS(const kgpu_handle *h, int t0) {
   int tx, ty; tileXY(h, t0, tx, ty);
   if (ty == 0) return h->periodic ? tx + (h->nYt - 1) * h->nXt : -1;
   return t0 - h->nXt;
}
static int tileN(const kgpu_handle *h, int t0) {
   int tx, ty; tileXY(h, t0, tx, ty);
   if (ty == h->nYt - 1) return h->periodic ? tx : -1;
   return t0 + h->nXt;
}
// Grid.f90:322-335
static bool onDomainEdge(const kgpu_handle *h, int t0) {
   int tx, ty; tileXY(h, t0, tx, ty);
   bool on = (tx == 0 || tx == h->nXt - 1);
   return on || (h->nYt > 1 && (ty == 0 || ty == h->nYt - 1));
}

static int syncStream(kgpu_handle *h) {
   CUDA_TRY(h, cudaStreamSynchronize(h->stream));
   return 0;
}

// Rebuild the device tile masks (with their ring) and the list of CUDA blocks that
// overlap at least one active tile.  Only runs when the active set changed.
static int refreshMasks(kgpu_handle *h) {
   if (!h->masksDirty) return 0;
   int mw = h->nXt + 2, mh = h->nYt + 2;
   std::vector<uint8_t> mask((size_t)mw * mh, 0), srcm((size_t)mw * mh, 0);
   for (int ty = -1; ty <= h->nYt; ty++)
      for (int tx = -1; tx <= h->nXt; tx++) {
         int sx = tx, sy = ty;
         if (h->periodic) { sx = (tx + h->nXt) % h->nXt; sy = (ty + h->nYt) % h->nYt; }
         if (sx < 0 || sx >= h->nXt || sy < 0 || sy >= h->nYt) {
            // ring tile owned by a neighbouring rank: periodic decomposed runs keep every tile active, dynamic ones
            // read the replicated table (outside the domain: nothing)
            if (h->comm.active && h->globalPeriodic) mask[(size_t)(ty + 1) * mw + tx + 1] = 2;
            else if (h->dyn) {
               int gx = h->gtx0 + tx, gy = h->gty0 + ty;
               if (gx >= 0 && gx < h->gnXt && gy >= 0 && gy < h->gnYt) mask[(size_t)(ty + 1) * mw + tx + 1] = (uint8_t)h->gt.tstate[gy * h->gnXt + gx];
            }
            continue;
         }
         int t0 = sy * h->nXt + sx;
         mask[(size_t)(ty + 1) * mw + tx + 1] = (uint8_t)h->tstate[t0];
         srcm[(size_t)(ty + 1) * mw + tx + 1] = (uint8_t)h->hasSource[t0];
      }
   CUDA_TRY(h, cudaMemcpyAsync(h->d_tileMask, mask.data(), mask.size(), cudaMemcpyHostToDevice, h->stream));
   CUDA_TRY(h, cudaMemcpyAsync(h->d_tileSource, srcm.data(), srcm.size(), cudaMemcpyHostToDevice, h->stream));
   int BX = h->oneD ? BX1 : BX2, BY = h->oneD ? BY1 : BY2;
   int nbx = (h->NX + BX - 1) / BX, nby = (h->NY + BY - 1) / BY;
   h->nbxAll = nbx; h->nbyAll = nby;
   std::vector<int2> list;
   list.reserve((size_t)nbx * nby);
   for (int by = 0; by < nby; by++)
      for (int bx = 0; bx < nbx; bx++) {
         int tx0 = (bx * BX) / h->nX, tx1 = std::min(bx * BX + BX - 1, h->NX - 1) / h->nX;
         int ty0 = (by * BY) / h->nY, ty1 = std::min(by * BY + BY - 1, h->NY - 1) / h->nY;
         bool any = false;
         for (int ty = ty0; ty <= ty1 && !any; ty++)
            for (int tx = tx0; tx <= tx1; tx++)
               if (h->tstate[ty * h->nXt + tx] == 2) { any = true; break; }
         if (any) list.push_back(make_int2(bx, by));
      }
   h->nBlocks = (int)list.size();
   if (h->nBlocks) CUDA_TRY(h, cudaMemcpyAsync(h->d_blockList, list.data(), list.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
   std::vector<int2> listM;
   if (h->morpho && !h->oneD) {   // fused morphodynamic tiles: BX2 x MORPHO_STAGE_BY cells, listed when they overlap an active tile
      const int BYM = MORPHO_STAGE_BY, nbyM = (h->NY + BYM - 1) / BYM;
      for (int by = 0; by < nbyM; by++)
         for (int bx = 0; bx < nbx; bx++) {
            int tx0 = (bx * BX) / h->nX, tx1 = std::min(bx * BX + BX - 1, h->NX - 1) / h->nX;
            int ty0 = (by * BYM) / h->nY, ty1 = std::min(by * BYM + BYM - 1, h->NY - 1) / h->nY;
            bool any = false;
            for (int ty = ty0; ty <= ty1 && !any; ty++)
               for (int tx = tx0; tx <= tx1; tx++)
                  if (h->tstate[ty * h->nXt + tx] == 2) { any = true; break; }
            if (any) listM.push_back(make_int2(bx, by));
         }
      h->nBlocksM = (int)listM.size();
      if (h->nBlocksM) CUDA_TRY(h, cudaMemcpyAsync(h->d_blockListM, listM.data(), listM.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
   }
   if (h->comm.active) {
      // blocks touching the edge of the local domain are computed first so that their strips can travel
      // while the interior is still being computed
      std::vector<int2> bnd, inr;
      // (every block that holds one of the two outermost cell columns / rows: with NY % BY == 1 the last block row
      // has a single row of cells and the second-to-last row of the strip lives in the block below it)
      for (const int2 &b : list) {
         bool edge = b.x * BX < 2 || (b.x + 1) * BX > h->NX - 2 || (!h->oneD && (b.y * BY < 2 || (b.y + 1) * BY > h->NY - 2));
         (edge ? bnd : inr).push_back(b);
      }
      h->nBoundary = (int)bnd.size(); h->nInterior = (int)inr.size();
      if (h->nBoundary) CUDA_TRY(h, cudaMemcpyAsync(h->d_blockBoundary, bnd.data(), bnd.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
      if (h->nInterior) CUDA_TRY(h, cudaMemcpyAsync(h->d_blockInterior, inr.data(), inr.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
   }
   CUDA_TRY(h, cudaStreamSynchronize(h->stream));
   h->masksDirty = false;
   h->topoBtIdx = -1;  // new blocks need their planes
   return 0;
}

// periodic wrap of the halo (single device).  Non-periodic domains need nothing: the
// cells around the active region are static ghost data.
static int exchangeHalo(kgpu_handle *h, double *const *planes, int nf, bool vertices, cudaStream_t s);
static int allreduceCfl(kgpu_handle *h, int slot);
static int allreduceNonfinite(kgpu_handle *h);

static int fillHaloCells(kgpu_handle *h, int k) {
   if (h->comm.active) return exchangeHalo(h, h->S[k], 4, false, h->stream);
   if (!h->periodic) return 0;
   HaloArgs a; a.nf = 4;
   for (int d = 0; d < 4; d++) a.f[d] = h->S[k][d];
   if (!h->oneD) {
      halo_periodic_x_kernel<<<(h->NY + 127) / 128, 128, 0, h->stream>>>(h->D, a, 0);
      halo_periodic_y_kernel<<<(h->NX + 4 + 127) / 128, 128, 0, h->stream>>>(h->D, a, 0);
      h->launches += 2;
   } else {
      halo_periodic_x_kernel<<<1, 32, 0, h->stream>>>(h->D, a, 0);
      h->launches += 1;
   }
   return 0;
}
static int fillHaloVertices(kgpu_handle *h, double *v) {
   if (h->comm.active) { double *pl[1] = {v}; return exchangeHalo(h, pl, 1, true, h->stream); }
   if (!h->periodic) return 0;
   HaloArgs a; a.nf = 1; a.f[0] = v;
   if (!h->oneD) {
      halo_periodic_x_kernel<<<(h->NY + 1 + 127) / 128, 128, 0, h->stream>>>(h->D, a, 1);
      halo_periodic_y_kernel<<<(h->NX + 5 + 127) / 128, 128, 0, h->stream>>>(h->D, a, 1);
      h->launches += 2;
   } else {
      halo_periodic_x_kernel<<<1, 32, 0, h->stream>>>(h->D, a, 1);
      h->launches += 1;
   }
   return 0;
}

template <bool ONED, bool HASBT, int LIM, int SPEC>
static void launchStageK(kgpu_handle *h, const StageArgs &a, dim3 grid) {
   constexpr int BX = ONED ? BX1 : BX2, BY = ONED ? BY1 : BY2;
   using G = StageGeom<BX, BY, ONED>;
   hydro_stage_kernel<BX, BY, ONED, HASBT, LIM, false, SPEC>
      <<<grid, STAGE_THREADS, G::smemBytes(HASBT, false, stageFluxPlanes(SPEC)), h->stream>>>(h->D, a);
}
template <bool ONED, bool HASBT, int LIM, int SPEC>
static void setStageAttr() {
   constexpr int BX = ONED ? BX1 : BX2, BY = ONED ? BY1 : BY2;
   using G = StageGeom<BX, BY, ONED>;
   cudaFuncSetAttribute(hydro_stage_kernel<BX, BY, ONED, HASBT, LIM, false, SPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                        (int)G::smemBytes(true, false, stageFluxPlanes(SPEC)));
}
template <bool ONED, int SPEC>
static void launchStageS(kgpu_handle *h, const StageArgs &a, dim3 grid, bool mm2) {
   if (h->morpho) { if (mm2) launchStageK<ONED, true, KGPU_LIM_MINMOD2, SPEC>(h, a, grid); else launchStageK<ONED, true, -1, SPEC>(h, a, grid); }
   else           { if (mm2) launchStageK<ONED, false, KGPU_LIM_MINMOD2, SPEC>(h, a, grid); else launchStageK<ONED, false, -1, SPEC>(h, a, grid); }
}
// nblocks CTAs taken from a.blockList, or -- when the list is the whole local domain in row-major order --
// a 2-D grid whose block indices are the tile coordinates (no dependent load before the TMA issue)
template <bool ONED>
static void launchStageT(kgpu_handle *h, StageArgs a, int nblocks) {
   if (nblocks <= 0) return;
   dim3 grid(nblocks);
   a.directNbx = 0;
   a.tune = h->tune & ~8;
   if ((h->tune & 8) && a.blockList == h->d_blockList && nblocks == h->nbxAll * h->nbyAll && h->nbyAll <= 65535) {
      a.directNbx = h->nbxAll;
      grid = dim3(h->nbxAll, h->nbyAll);
   }
   const bool mm2 = h->P.limiter == KGPU_LIM_MINMOD2;  // the default limiter gets a branch-free instantiation
   // geometric factors on, no eddy viscosity (the reference's defaults): the instantiation with both compiled in
   const bool spec = !ONED && h->useSpec && h->D.geom && !(h->D.nu > 0.0);
   h->launches++;
   if (h->P.arithmetic == 1) { launch_stage_fast(ONED, h->morpho, mm2, spec, grid, h->stream, h->D, a); return; }
   if (!ONED && spec) launchStageS<false, 1>(h, a, grid, mm2);
   else launchStageS<ONED, 0>(h, a, grid, mm2);
}

// Tensor maps for the TMA staging of the stage kernel: every plane is a (rows x pitch) fp64
// matrix; the box is the halo'd CTA tile (cells), BY face rows (x faces) or BY + 3 rows (y faces).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool buildTensorMaps(kgpu_handle *h) {
   static EncodeTiledFn encode = nullptr;
   if (!encode) {
      void *fn = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return false;
      encode = (EncodeTiledFn)fn;
   }
   const int BX = h->oneD ? BX1 : BX2, BY = h->oneD ? BY1 : BY2;
   const cuuint32_t RX = BX + 4, RY = h->oneD ? 1 : BY + 4, FY = h->oneD ? 1 : BY + 3;
   std::vector<TmaDesc> maps(TMA_NSLOTS);
   std::memset(maps.data(), 0, sizeof(TmaDesc) * maps.size());
   static_assert(sizeof(CUtensorMap) == sizeof(TmaDesc), "CUtensorMap is 128 bytes");
   auto put = [&](int slot, double *base, cuuint32_t boxRows) {
      if (!base) return true;
      CUtensorMap m;
      cuuint64_t dims[2] = {(cuuint64_t)h->pitch, (cuuint64_t)h->rows};
      cuuint64_t strides[1] = {(cuuint64_t)h->pitch * sizeof(double)};
      cuuint32_t box[2] = {RX, boxRows}, estr[2] = {1, 1};
      if (encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
         return false;
      std::memcpy(&maps[slot], &m, sizeof(m));
      return true;
   };
   bool ok = true;
   for (int k = 0; k < h->nStates; k++) for (int d = 0; d < 4; d++) ok = ok && put(TMA_STATE0 + 4 * k + d, h->S[k][d], RY);
   ok = ok && put(TMA_B0C, h->topo.b0c, RY) && put(TMA_GAMC, h->topo.gamc, RY) && put(TMA_BTC, h->topo.btc, RY);
   ok = ok && put(TMA_XB0, h->topo.xb0, BY) && put(TMA_XTAN, h->topo.xtan, BY) && put(TMA_XGAM, h->topo.xgam, BY) &&
        put(TMA_XB, h->topo.xB, BY) && put(TMA_XBT, h->topo.xbt, BY);
   ok = ok && put(TMA_YB0, h->topo.yb0, FY) && put(TMA_YTAN, h->topo.ytan, FY) && put(TMA_YGAM, h->topo.ygam, FY) &&
        put(TMA_YB, h->topo.yB, FY) && put(TMA_YBT, h->topo.ybt, FY);
   if (!ok) return false;
   if (cudaMalloc(&h->d_maps, sizeof(TmaDesc) * maps.size()) != cudaSuccess) return false;
   return cudaMemcpy(h->d_maps, maps.data(), sizeof(TmaDesc) * maps.size(), cudaMemcpyHostToDevice) == cudaSuccess;
}

// (Re)compute the topography planes from the vertex arrays for every listed block.
static int computeTopo(kgpu_handle *h, int kbt) {
   if (h->nBlocks == 0) return 0;
   const double *btv = h->morpho ? h->btv[kbt] : nullptr;
   if (h->oneD) topo_planes_kernel<BX1, BY1, true><<<h->nBlocks, NTHREADS, 0, h->stream>>>(h->D, h->b0v, btv, h->topo, h->d_blockList, h->P.arithmetic == 1 ? 1 : 0);
   else topo_planes_kernel<BX2, BY2, false><<<h->nBlocks, NTHREADS, 0, h->stream>>>(h->D, h->b0v, btv, h->topo, h->d_blockList, h->P.arithmetic == 1 ? 1 : 0);
   h->launches++;
   h->topoBtIdx = kbt;
   CUDA_TRY(h, cudaGetLastError());
   return 0;
}

// One fused RHS(+stage update) launch.  mode: StageMode; qin / qout = state buffer indices
// (MODE_RHS writes E0/I0 instead of a state).
static int launchStage(kgpu_handle *h, int mode, int kin, int kout, int kq0, int kbt) {
   if (h->nBlocks == 0 && !h->comm.active) return 0;   // (a rank without active blocks still takes part in the exchange)
   StageArgs a;
   for (int d = 0; d < 4; d++) {
      a.qin[d] = h->S[kin][d];
      a.q0[d] = h->S[kq0][d];
      a.qout[d] = (mode == MODE_RHS) ? h->E0[d] : h->S[kout][d];
   }
   a.Iout = h->I0;
   a.T = h->topo;
   a.maps = h->d_maps; a.mapIn = TMA_STATE0 + 4 * kin;
   a.mx = h->mp(); a.doMaxima = (mode == MODE_FINAL && kq0 == h->i0) ? 1 : 0;
   a.prefetchDistance = h->prefetchDistance;
   if (h->topoBtIdx != kbt) { int rct = computeTopo(h, kbt); if (rct) return rct; }
   a.tileMask = h->d_tileMask; a.tileSource = h->d_tileSource; a.blockList = h->d_blockList;
   a.ctrl = h->d_ctrl; a.sources = h->d_sources; a.sourcePool = h->d_srcPool;
   a.mode = mode;
   a.allActive = h->allActive() ? 1 : 0;
   if (h->timeRhs) cudaEventRecord(h->evA, h->stream);
   bool overlap = h->comm.active && mode != MODE_RHS;
   if (!overlap) {
      if (h->oneD) launchStageT<true>(h, a, h->nBlocks); else launchStageT<false>(h, a, h->nBlocks);
   } else {
      // edge blocks first; their strips travel on the communication stream while the interior runs
      a.blockList = h->d_blockBoundary;
      if (h->oneD) launchStageT<true>(h, a, h->nBoundary); else launchStageT<false>(h, a, h->nBoundary);
      CUDA_TRY(h, cudaEventRecord(h->comm.evBoundary, h->stream));
      CUDA_TRY(h, cudaStreamWaitEvent(h->comm.stream, h->comm.evBoundary, 0));
      int rc = exchangeHalo(h, h->S[kout], 4, false, h->comm.stream);
      if (rc) return rc;
      CUDA_TRY(h, cudaEventRecord(h->comm.evHalo, h->comm.stream));
      a.blockList = h->d_blockInterior;
      if (h->oneD) launchStageT<true>(h, a, h->nInterior); else launchStageT<false>(h, a, h->nInterior);
      CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->comm.evHalo, 0));
   }
   if (h->timeRhs) {
      cudaEventRecord(h->evB, h->stream);
      cudaEventSynchronize(h->evB);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, h->evA, h->evB);
      h->rhsMs += ms;
      h->rhsLaunches++;
   }
   CUDA_TRY(h, cudaGetLastError());
   // single device: the periodic image of what was just written
   if (!overlap && mode != MODE_RHS) return fillHaloCells(h, kout);
   return 0;
}

static int readCtrl(kgpu_handle *h) {
   CUDA_TRY(h, cudaMemcpyAsync(h->h_ctrl, h->d_ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, h->stream));
   CUDA_TRY(h, cudaStreamSynchronize(h->stream));
   return 0;
}

// =========================================================================== topography + tiles
static int defaultTileXY(kgpu_handle *h, int tx, int ty, int kind) {
   dim3 grid((h->nX + 127) / 128, h->nY);
   const kgpu_params &P = h->P;
   AllStates all;
   all.n = h->nStates;
   const int order[5] = {h->i0, h->ia, h->ib, h->ic, h->id};
   for (int k = 0; k < h->nStates; k++) for (int d = 0; d < 4; d++) all.q[k][d] = h->S[order[k]][d];
   tile_default_kernel<<<grid, 128, 0, h->stream>>>(h->D, all, h->b0v, h->morpho ? h->btv[h->bt0] : nullptr, tx, ty, kind,
                                                    P.bcsHnval, P.bcsuval, P.bcsvval, P.bcspsival);
   h->launches++;
   CUDA_TRY(h, cudaGetLastError());
   return 0;
}
static int defaultTile(kgpu_handle *h, int t0, int kind) {
   int tx, ty; tileXY(h, t0, tx, ty);
   return defaultTileXY(h, tx, ty, kind);
}
static int ghostData(kgpu_handle *h, int t0) {
   bool dirichlet = (h->P.bcs == KGPU_BC_DIRICHLET) && onDomainEdge(h, t0);
   return defaultTile(h, t0, dirichlet ? 1 : 0);
}

// GetHeights + EqualiseTopographicBoundaryData (dem.f90:360, MorphodynamicRHS.f90:588-687): a
// shared vertex takes the value of the tile for which it is local index 1.
static int loadHeights(kgpu_handle *h, int t0, const double *given) {
   if (h->loaded[t0] && !given) return 0;
   int nX = h->nX, nY = h->nY;
   size_t nv = (size_t)(nX + 1) * (nY + 1);
   double *hb = h->h_stage;
   const bool fromRaster = !given && h->raster.elev != nullptr;
   const bool onDevice = !given && (fromRaster || h->topogFn.func >= 0);
   if (given) {
      std::memcpy(hb, given, sizeof(double) * (size_t)(nX + 1) * (h->oneD ? 1 : nY + 1));
   } else if (!onDevice) {
      if (!h->P.heights) { h->err = "no heights callback registered and no b0_vertices given"; return KGPU_ERR_ARG; }
      if (h->P.heights(h->P.heights_ctx, globalTileId(h, t0), hb) != 0) { h->err = "heights callback failed"; return KGPU_ERR_ARG; }
   }
   int tE = tileE(h, t0), tN = h->oneD ? -1 : tileN(h, t0);
   int tNE = (tE >= 0 && !h->oneD) ? tileN(h, tE) : -1;
   bool eL = tE >= 0 && tE != t0 && h->loaded[tE], nL = tN >= 0 && tN != t0 && h->loaded[tN], neL = tNE >= 0 && h->loaded[tNE];
   int mask = 0;
   if (!eL && !(h->periodic && h->nXt == 1)) mask |= 1;
   if (!nL && !(h->periodic && h->nYt == 1)) mask |= 2;
   if (!(eL || nL || neL) && (mask & 1) && (mask & 2)) mask |= 4;
   int tx, ty; tileXY(h, t0, tx, ty);
   dim3 grid((nX + 1 + 127) / 128, h->oneD ? 1 : nY + 1);
   if (onDevice) {   // TopogFuncs.f90 at the tile's vertices, straight into the staging buffer (global 1-based tile indices)
      if (fromRaster) tile_raster_kernel<<<grid, 128, 0, h->stream>>>(h->D, h->raster, h->gtx0 + tx + 1, h->gty0 + ty + 1, h->d_stage);
      else tile_topog_kernel<<<grid, 128, 0, h->stream>>>(h->D, h->topogFn, h->gtx0 + tx + 1, h->gty0 + ty + 1, h->d_stage);
      h->launches++;
   } else CUDA_TRY(h, cudaMemcpyAsync(h->d_stage, hb, nv * sizeof(double), cudaMemcpyHostToDevice, h->stream));
   tile_vertices_kernel<<<grid, 128, 0, h->stream>>>(h->D, h->b0v, h->d_stage, tx, ty, 1, mask);
   h->launches++;
   CUDA_TRY(h, cudaStreamSynchronize(h->stream));  // h_stage is reused
   int rc = fillHaloVertices(h, h->b0v);
   if (rc) return rc;
   h->loaded[t0] = 1;
   h->topoBtIdx = -1;
   // ghost tiles sharing the refreshed seam keep w = b0 at the cell centre
   int tW = tileW(h, t0), tS = h->oneD ? -1 : tileS(h, t0);
   int tSW = (tW >= 0 && !h->oneD) ? tileS(h, tW) : -1;
   for (int tt : {tW, tS, tSW})
      if (tt >= 0 && tt != t0 && h->loaded[tt] && h->tstate[tt] == 1) {
         rc = ghostData(h, tt);
         if (rc) return rc;
      }
   return 0;
}

// UpdateTiles.f90:389-481
static int addGhostTiles(kgpu_handle *h, int t0) {
   int nb[8], n = 0;
   nb[n++] = tileW(h, t0); nb[n++] = tileE(h, t0);
   if (!h->oneD) {
      nb[n++] = tileN(h, t0); nb[n++] = tileS(h, t0);
      if (!h->periodic) {
         int s = tileS(h, t0), nn = tileN(h, t0);
         nb[n++] = s >= 0 ? tileW(h, s) : -1; nb[n++] = s >= 0 ? tileE(h, s) : -1;
         nb[n++] = nn >= 0 ? tileW(h, nn) : -1; nb[n++] = nn >= 0 ? tileE(h, nn) : -1;
      }
   }
   for (int k = 0; k < n; k++) {
      int tt = nb[k];
      if (tt < 0) { h->err = "ghost tile out of bounds (UpdateTiles.f90:423)"; return KGPU_ERR_ARG; }
      if (h->tstate[tt] != 0) continue;
      h->tstate[tt] = 1;
      h->ghostList.push_back(tt + 1);
      int rc = loadHeights(h, tt, nullptr);
      if (rc) return rc;
      rc = ghostData(h, tt);
      if (rc) return rc;
   }
   return 0;
}

// AddTile (UpdateTiles.f90:56-78): AddToActiveTiles + AllocateTile + ActivateTile
static int addTile(kgpu_handle *h, int t0, bool countIt) {
   if (t0 < 0 || t0 >= h->nTiles || (onDomainEdge(h, t0) && !h->periodic)) {
      if (h->P.bcs == KGPU_BC_HALT) {
         h->err = "tried to add a tile outside the domain (Boundary Conditions = halt)";
         return KGPU_ERR_HALT_BC;
      }
      return 0;
   }
   if (h->tstate[t0] == 2) return 0;
   bool wasGhost = h->tstate[t0] == 1;
   h->tstate[t0] = 2;
   h->activeList.insert(std::upper_bound(h->activeList.begin(), h->activeList.end(), t0 + 1), t0 + 1);
   if (wasGhost) h->ghostList.erase(std::find(h->ghostList.begin(), h->ghostList.end(), t0 + 1));
   h->hasSource[t0] = 0;
   h->masksDirty = true;
   int rc = loadHeights(h, t0, nullptr);
   if (rc) return rc;
   rc = defaultTile(h, t0, wasGhost ? 2 : 0);  // fresh tile: zeros + w = b0; promoted ghost: w = b0
   if (rc) return rc;
   rc = addGhostTiles(h, t0);
   if (countIt) h->ntilesAdded++;
   return rc;
}

// CheckIfNearBoundaries (TimeStepper.f90:924-1150): flags on the device, list replay on the host
static int dynCheckIfNearBoundaries(kgpu_handle *h);
static int checkIfNearBoundaries(kgpu_handle *h) {
   if (h->dyn) return dynCheckIfNearBoundaries(h);
   if (h->allActive()) { h->firstScan = false; return 0; }
   int nAct = (int)h->activeList.size();
   if (nAct == 0) return 0;
   std::vector<int> flags(h->nTiles, 0);
   if (h->firstScan) {
      for (int id : h->activeList) flags[id - 1] = h->seedFlags[id - 1];
   } else {
      std::vector<int> tl(nAct);
      for (int k = 0; k < nAct; k++) tl[k] = h->activeList[k] - 1;
      CUDA_TRY(h, cudaMemcpyAsync(h->d_tileList, tl.data(), nAct * sizeof(int), cudaMemcpyHostToDevice, h->stream));
      tile_flags_kernel<<<nAct, 128, 0, h->stream>>>(h->D, h->sp(h->i0), h->b0v, h->morpho ? h->btv[h->bt0] : nullptr,
                                                     h->d_tileList, h->P.TileBuffer, h->d_flags);
      h->launches++;
      CUDA_TRY(h, cudaMemcpyAsync(h->h_flags, h->d_flags, nAct * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      for (int k = 0; k < nAct; k++) flags[tl[k]] = h->h_flags[k];
   }
   int buf = h->P.TileBuffer;
   // replay of the four passes with the trip count fixed at loop entry (quirk Q3)
   for (int dir = (h->oneD ? 2 : 0); dir < 4; dir++) {
      int trip = (int)h->activeList.size();
      for (int tt = 0; tt < trip; tt++) {
         int t0 = h->activeList[tt] - 1;
         int tx, ty; tileXY(h, t0, tx, ty);
         int nbr;
         switch (dir) {
            case 0: nbr = (h->periodic && ty == h->nYt - 1) ? tx : t0 + h->nXt; break;
            case 1: nbr = (h->periodic && ty == 0) ? tx + (h->nYt - 1) * h->nXt : t0 - h->nXt; break;
            case 2: nbr = (h->periodic && tx == h->nXt - 1) ? ty * h->nXt : t0 + 1; break;
            default: nbr = (h->periodic && tx == 0) ? (h->nXt - 1) + ty * h->nXt : t0 - 1; break;
         }
         if (nbr + 1 <= 0) continue;
         if (nbr < h->nTiles && h->tstate[nbr] == 2) continue;
         int f = flags[t0];
         bool trig;
         switch (dir) {
            case 0: trig = (f & 1) || (0 > h->nY - buf); break;
            case 1: trig = (f & 2) || (h->nY <= buf); break;
            case 2: trig = (f & 4) || (0 > h->nX - buf); break;
            default: trig = (f & 8) || (h->nX <= buf); break;
         }
         if (trig) {
            int rc = addTile(h, nbr, true);
            if (rc) return rc;
         }
      }
   }
   h->firstScan = false;
   return 0;
}

// TimeStepper.f90:281-305
static double nextFluxSeriesTime(const kgpu_handle *h, double t) {
   double nextT = HUGE_D, tdiff = HUGE_D;
   for (const DevSource &S : h->src)
      for (int j = 0; j < S.n; j++) {
         double tmp = h->srcPool[S.off + j] - t;
         if (tmp > 0.0 && tmp < tdiff) { tdiff = tmp; nextT = h->srcPool[S.off + j]; }
      }
   return nextT;
}

// =========================================================================== the step
// substep-1 RHS of state kin -> E0, I0 and the advised dt (HydraulicRHS.f90:64-174)
static int firstRHS(kgpu_handle *h, int kin, int kbt, double tNow, double tmax, int setDt) {
   int rc;
   ctrl_begin_kernel<<<1, 1, 0, h->stream>>>(h->d_ctrl, tNow);
   rc = launchStage(h, MODE_RHS, kin, -1, kin, kbt);
   if (rc) return rc;
   if ((rc = allreduceCfl(h, 0))) return rc;
   ctrl_advise_kernel<<<1, 1, 0, h->stream>>>(h->D, h->d_ctrl, h->allActive() ? 0 : 1, tmax, setDt);
   h->launches += 2;
   return 0;
}

// HydraulicTimeStepper (TimeStepper.f90:333-527) with dt taken from the control block.
// On return h_ctrl is current; h_ctrl->failed != 0 means "refine" with dtNew.
static int hydraulicTimeStepper(kgpu_handle *h, int kq0, int ka, int kb, int kbt) {
   int BX = h->oneD ? BX1 : BX2, BY = h->oneD ? BY1 : BY2;
   int some = h->allActive() ? 0 : 1;
   // stage 1 (elementwise; E0/I0 retained for cheap retries)
   Update1Args u;
   for (int d = 0; d < 4; d++) { u.q0[d] = h->S[kq0][d]; u.E[d] = h->E0[d]; u.q1[d] = h->S[ka][d]; }
   u.I = h->I0; u.tileMask = h->d_tileMask; u.blockList = h->d_blockList; u.ctrl = h->d_ctrl; u.allActive = h->allActive() ? 1 : 0;
   int rc;
   auto update1 = [&](const int2 *list, int n) {
      if (n <= 0) return;
      u.blockList = list;
      if (h->oneD) stage1_update_kernel<BX1, BY1><<<n, NTHREADS, 0, h->stream>>>(h->D, u);
      else stage1_update_kernel<BX2, BY2><<<n, NTHREADS, 0, h->stream>>>(h->D, u);
      h->launches++;
   };
   if (h->comm.active) {
      // edge blocks first: their strips travel on the communication stream while the interior is updated, exactly
      // as in the fused stages (launchStage); stage 2 waits for the halo event
      update1(h->d_blockBoundary, h->nBoundary);
      CUDA_TRY(h, cudaEventRecord(h->comm.evBoundary, h->stream));
      CUDA_TRY(h, cudaStreamWaitEvent(h->comm.stream, h->comm.evBoundary, 0));
      if ((rc = exchangeHalo(h, h->S[ka], 4, false, h->comm.stream))) return rc;
      CUDA_TRY(h, cudaEventRecord(h->comm.evHalo, h->comm.stream));
      update1(h->d_blockInterior, h->nInterior);
      CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->comm.evHalo, 0));
   } else {
      update1(h->d_blockList, h->nBlocks);
      if ((rc = fillHaloCells(h, ka))) return rc;
   }
   if ((rc = launchStage(h, MODE_STAGE2, ka, kb, kq0, kbt))) return rc;
   if ((rc = allreduceCfl(h, 1))) return rc;
   ctrl_check_kernel<<<1, 1, 0, h->stream>>>(h->D, h->d_ctrl, some, 1);
   if ((rc = launchStage(h, MODE_STAGE3, kb, ka, kq0, kbt))) return rc;
   if ((rc = allreduceCfl(h, 2))) return rc;
   ctrl_check_kernel<<<1, 1, 0, h->stream>>>(h->D, h->d_ctrl, some, 2);
   if ((rc = launchStage(h, MODE_FINAL, ka, kb, kq0, kbt))) return rc;
   h->launches += 2;
   // Maxima (UpdateMaximum*, TimeStepper.f90:519-524, quirk Q1) are those of the state at the start of the whole step,
   // tileContainer, and ride in the final stage launch of the FIRST hydraulic operator (its q0 is that state).  The
   // reference runs them again at the end of the second H of a Strang step, on the same tileContainer with a later
   // stamp: every test there is a strict `>` (or tfirst == -1) against values the first pass has just stored, so that
   // pass cannot change anything and is not launched (round 1 ran a maxima_kernel here: 2 % of the Strang step).
   CUDA_TRY(h, cudaGetLastError());
   if ((rc = allreduceNonfinite(h))) return rc;
   return readCtrl(h);
}

#include "kgpu_comm_host.inl"
#include "kgpu_dyn_host.inl"
#include "kgpu_morpho_host.inl"

static int integrateTo(kgpu_handle *h, double tend, int64_t maxSteps, kgpu_step_info *info) {
   int64_t done = 0;
   double dt_hydro = h->dtgrid;
   bool integrating = tend > h->t;
   int rc;
   while (integrating) {
      if ((rc = checkIfNearBoundaries(h))) return rc;
      if (h->masksDirty) {  // tile data changed: refresh the wrapped image of the current state
         if ((rc = refreshMasks(h))) return rc;
         if ((rc = fillHaloCells(h, h->i0))) return rc;
      }
      double t0 = h->t;
      double tmax = std::min(tend, nextFluxSeriesTime(h, h->t));
      if ((rc = firstRHS(h, h->i0, h->bt0, h->t, tmax, h->morpho ? 2 : 1))) return rc;
      h->e0Valid = true;
      bool haveDt = false;  // false: dt comes from ctrl_advise; true: the host dictates dt_hydro
      int guard = 0, kres = h->ib;
      while (true) {
         if (++guard > 200) { h->err = "time step underflow (200 refinements of one step)"; return KGPU_ERR_DT; }
         if (h->morpho && h->topoBtIdx != h->bt0 && (rc = computeTopo(h, h->bt0))) return rc;
         if (!h->e0Valid) {  // E0/I0 were reused by the second H operator: re-evaluate (rare)
            if ((rc = firstRHS(h, h->i0, h->bt0, t0, tmax, 0))) return rc;
            h->e0Valid = true;
         }
         if (haveDt) {
            ctrl_set_dt_kernel<<<1, 1, 0, h->stream>>>(h->d_ctrl, dt_hydro);
            h->launches++;
         }
         if ((rc = hydraulicTimeStepper(h, h->i0, h->ia, h->ib, h->bt0))) return rc;
         if (h->h_ctrl->nonfinite) { h->err = "non-finite state"; return KGPU_ERR_DT; }
         dt_hydro = h->h_ctrl->dt;
         if (!(dt_hydro > 0.0) || !std::isfinite(dt_hydro)) { h->err = "time step underflow"; return KGPU_ERR_DT; }
         if (h->h_ctrl->failed) {
            h->nrefines++;
            dt_hydro = h->h_ctrl->dtNew;
            haveDt = true;
            continue;
         }
         if (!h->morpho) break;
         bool again = false;
         if ((rc = strangRemainder(h, t0, dt_hydro, again))) return rc;
         if (!again) { kres = h->id; break; }
         haveDt = true;
      }
      // CopySolutionData(intermed3 -> tileContainer) by pointer rotation
      if (!h->morpho) std::swap(h->i0, h->ib);
      else { std::swap(h->i0, h->id); std::swap(h->bt0, h->bt3); }
      (void)kres;
      h->havePre = true;
      h->t = h->morpho ? t0 + 2.0 * dt_hydro : t0 + dt_hydro;
      h->dtgrid = dt_hydro;
      h->nsteps++;
      done++;
      if (h->t >= tend) integrating = false;
      if (maxSteps > 0 && done >= maxSteps) integrating = false;
   }
   if (info) {
      info->t = h->t; info->dt_last = dt_hydro; info->nsteps = h->nsteps; info->nrefines = h->nrefines; info->ntiles_added = h->ntilesAdded;
   }
   return 0;
}

// =========================================================================== C-ABI
extern "C" {

const char *kgpu_version(void) { return "kestrel-b200 0.1 (sm_100a)"; }
const char *kgpu_last_error(const kgpu_handle *h) { return h ? h->err.c_str() : "null handle"; }
int64_t kgpu_launch_count(const kgpu_handle *h) { return h ? h->launches : 0; }
void *kgpu_stream(kgpu_handle *h) { return h ? (void *)h->stream : nullptr; }

int kgpu_rhs_timing(kgpu_handle *h, double *ms, int64_t *launches, int32_t reset) {
   if (!h) return KGPU_ERR_ARG;
   if (ms) *ms = h->rhsMs;
   if (launches) *launches = h->rhsLaunches;
   if (reset == 1) { h->rhsMs = 0.0; h->rhsLaunches = 0; h->timeRhs = true; }
   if (reset == 2) { h->timeRhs = false; }
   return 0;
}

int kgpu_destroy(kgpu_handle *h) {
   if (!h) return 0;
   cudaSetDevice(h->dev);
   if (h->stream) cudaStreamSynchronize(h->stream);
   for (int k = 0; k < 5; k++) for (int d = 0; d < 4; d++) cudaFree(h->S[k][d]);
   cudaFree(h->Um); cudaFree(h->Vm); cudaFree(h->mcHn0);
   for (int a = 0; a < 2; a++) for (int b = 0; b < 4; b++) cudaFree(h->mc[a][b]);
   for (int d = 0; d < 4; d++) { cudaFree(h->E0[d]); cudaFree(h->btv[d]); }
   cudaFree(h->I0); cudaFree(h->b0v); cudaFree(h->EBt); cudaFree(h->EmD);
   for (int k = 0; k < 11; k++) cudaFree(h->mx[k]);
   cudaFree(h->d_tileMask); cudaFree(h->d_tileSource); cudaFree(h->d_blockList); cudaFree(h->d_ctrl);
   cudaFree(h->d_rankMap); cudaFree(h->d_raster);
   cudaFree(h->d_sources); cudaFree(h->d_srcPool); cudaFree(h->d_stage); cudaFree(h->d_tileList); cudaFree(h->d_flags); cudaFree(h->d_gflags); cudaFree(h->d_redist); cudaFree(h->d_maps);
   cudaFreeHost(h->h_ctrl); cudaFreeHost(h->h_stage); cudaFreeHost(h->h_flags); cudaFreeHost(h->h_gflags); cudaFreeHost(h->h_redist);
   {
      double *pl[15] = {h->topo.b0c, h->topo.bxc, h->topo.byc, h->topo.gamc, h->topo.xb0, h->topo.xB, h->topo.xtan, h->topo.xgam,
                        h->topo.yb0, h->topo.yB, h->topo.ytan, h->topo.ygam, h->topo.btc, h->topo.xbt, h->topo.ybt};
      for (int k = 0; k < 15; k++) cudaFree(pl[k]);
   }
   cudaFree(h->d_blockBoundary); cudaFree(h->d_blockInterior); cudaFree(h->d_blockListM);
   for (int k = 0; k < 4; k++) { cudaFree(h->comm.sendBuf[k]); cudaFree(h->comm.recvBuf[k]); }
   if (h->comm.nccl && g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)h->comm.nccl);
   if (h->comm.evBoundary) cudaEventDestroy(h->comm.evBoundary);
   if (h->comm.evHalo) cudaEventDestroy(h->comm.evHalo);
   if (h->comm.stream) cudaStreamDestroy(h->comm.stream);
   if (h->evA) cudaEventDestroy(h->evA);
   if (h->evB) cudaEventDestroy(h->evB);
   if (h->copyStream) { cudaStreamSynchronize(h->copyStream); cudaStreamDestroy(h->copyStream); }
   if (h->evSnap) cudaEventDestroy(h->evSnap);
   if (h->evCopied) cudaEventDestroy(h->evCopied);
   for (int k = 0; k < 5; k++) cudaFree(h->snap[k]);
   cudaFree(h->rg.dEntries); cudaFree(h->rg.dPatch); cudaFree(h->rg.dSend); cudaFree(h->rg.dCounts); cudaFree(h->rg.dVslot);
   cudaFree(h->rg.dCslot); cudaFree(h->rg.dVkey); cudaFree(h->rg.dVbase); cudaFree(h->rg.dCkey); cudaFree(h->rg.dCbase);
   if (h->stream) cudaStreamDestroy(h->stream);
   delete h;
   return 0;
}

int kgpu_create(const kgpu_params *p, kgpu_handle **out) {
   if (!p || !out || p->struct_bytes != (int32_t)sizeof(kgpu_params)) return KGPU_ERR_ARG;
   if (p->n_sources < 0 || (p->n_sources > 0 && !p->sources)) return KGPU_ERR_ARG;
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return KGPU_ERR_CUDA;  // no CPU fallback
   kgpu_handle *h = new kgpu_handle();
   h->P = *p;
   h->P.sources = nullptr;
   if (p->device >= 0) { h->dev = p->device; if (cudaSetDevice(h->dev) != cudaSuccess) { delete h; return KGPU_ERR_CUDA; } }
   else cudaGetDevice(&h->dev);
   for (int s = 0; s < p->n_sources; s++) {
      const kgpu_source &k = p->sources[s];
      if (k.n_series < 1 || !k.time || !k.flux || !k.psi) { delete h; return KGPU_ERR_ARG; }
      DevSource S;
      std::memset(&S, 0, sizeof(S));
      S.x = k.x; S.y = k.y; S.radius = k.radius; S.numCells = k.num_cells_in_src; S.n = k.n_series;
      S.off = (long long)h->srcPool.size();
      h->srcPool.insert(h->srcPool.end(), k.time, k.time + k.n_series);
      h->srcPool.insert(h->srcPool.end(), k.flux, k.flux + k.n_series);
      h->srcPool.insert(h->srcPool.end(), k.psi, k.psi + k.n_series);
      h->src.push_back(S);
   }
   h->nX = p->nXpertile; h->nY = p->nYpertile; h->nXt = p->nXtiles; h->nYt = p->nYtiles;
   h->gnXt = p->nXtiles; h->gnYt = p->nYtiles;
   h->oneD = p->isOneD != 0; h->globalPeriodic = p->bcs == KGPU_BC_PERIODIC; h->morpho = p->MorphodynamicsOn != 0;
   h->periodic = h->globalPeriodic;
   if (p->comm_size > 1) {
      kgpu_comm &c = h->comm;
      c.size = p->comm_size; c.rank = p->comm_rank; c.px = p->comm_px; c.py = p->comm_py;
      bool okc = c.px >= 1 && c.py >= 1 && c.px * c.py == c.size && c.rank >= 0 && c.rank < c.size &&
                 p->nXtiles % c.px == 0 && p->nYtiles % c.py == 0 && !(h->oneD && c.py != 1);
      // periodic domains run all-active (kgpu_upload_domain); the others keep dynamic tiles through the replicated
      // tile table (kgpu_dyn_host.inl)
      const bool okd = h->globalPeriodic || (p->nXpertile >= 3 && (h->oneD || p->nYpertile >= 3));
      if (!okc || !okd) {
         fprintf(stderr, "kgpu_create: decomposition needs px*py = size and tiles divisible by px, py; without periodic bcs also "
                         "tiles of at least 3 cells\n");
         delete h;
         return okc ? KGPU_ERR_UNSUPPORTED : KGPU_ERR_ARG;
      }
      h->dyn = !h->globalPeriodic;
      c.rx = c.rank % c.px; c.ry = c.rank / c.px;
      h->nXt = p->nXtiles / c.px; h->nYt = p->nYtiles / c.py;
      h->gtx0 = c.rx * h->nXt; h->gty0 = c.ry * h->nYt;
      auto rk = [&](int x, int y) {   // -1: the domain ends there
         if (h->dyn && (x < 0 || x >= c.px || y < 0 || y >= c.py)) return -1;
         return ((y + c.py) % c.py) * c.px + ((x + c.px) % c.px);
      };
      if (c.px > 1) { c.west = rk(c.rx - 1, c.ry); c.east = rk(c.rx + 1, c.ry); }
      if (c.py > 1) { c.south = rk(c.rx, c.ry - 1); c.north = rk(c.rx, c.ry + 1); }
      h->periodic = false;  // the local block does not wrap onto itself (per direction handled in exchangeHalo)
   }
   h->NX = h->nX * h->nXt; h->NY = h->nY * h->nYt; h->nTiles = h->nXt * h->nYt;
   int BX = h->oneD ? BX1 : BX2, BY = h->oneD ? BY1 : BY2;
   h->pitch = roundUp(XO + roundUp(h->NX, BX) + 8, 16);
   h->rows = h->oneD ? (YO + 1 + 3) : (YO + roundUp(h->NY, BY) + 4);
   h->fieldElems = (size_t)h->pitch * h->rows;

   DevParams &D = h->D;
   std::memset(&D, 0, sizeof(D));
   D.NX = h->NX; D.NY = h->NY; D.nX = h->nX; D.nY = h->nY; D.nXt = h->nXt; D.nYt = h->nYt;
   D.gtx0 = h->gtx0; D.gty0 = h->gty0; D.gnXt = h->gnXt; D.gnYt = h->gnYt;
   D.mm2HalfTheta = 0.5 * 1.3;
   D.halfGRhow = 0.5 * p->g * p->rhow;
   D.pitch = h->pitch; D.rows = h->rows;
   D.haloValid = p->comm_size > 1 ? 1 : 0;
   D.oneD = h->oneD; D.periodic = h->periodic; D.geom = p->geometric_factors != 0; D.morpho = h->morpho;
   D.bcDirichlet = p->bcs == KGPU_BC_DIRICHLET ? 1 : 0; D.bcU = p->bcsuval; D.bcV = p->bcsvval; D.bcPsi = p->bcspsival;
   D.limiter = p->limiter; D.drag = p->drag; D.erosion = p->erosion; D.deposition = p->deposition;
   D.eroTrans = p->erosion_transition; D.damp = p->morpho_damp; D.fswitch = p->fswitch;
   D.nSources = p->n_sources;
   D.dx = p->deltaX; D.dy = p->deltaY; D.dxR = 1.0 / p->deltaX; D.dyR = 1.0 / p->deltaY; D.xSize = p->xSize; D.ySize = p->ySize;
   D.g = p->g; D.rhow = p->rhow; D.rhos = p->rhos; D.gred = p->gred;
   D.ChezyCo = p->ChezyCo; D.ManningCo = p->ManningCo; D.CoulombCo = p->CoulombCo;
   D.PoulMin = p->PouliquenMinSlope; D.PoulMax = p->PouliquenMaxSlope; D.PoulInt = p->PouliquenIntermediateSlope; D.PoulBeta = p->PouliquenBeta;
   D.EdBetastar = p->Edwards2019betastar; D.EdKappa = p->Edwards2019kappa; D.EdGamma = p->Edwards2019Gamma;
   D.SwitchRate = p->VoellmySwitchRate; D.SwitchValue = p->VoellmySwitchValue;
   D.EroRate = p->EroRate; D.EroRateGranular = p->EroRateGranular; D.CriticalShields = p->CriticalShields; D.EroDepth = p->EroDepth;
   D.EroCritH = p->EroCriticalHeight; D.BedPorosity = p->BedPorosity; D.maxPack = p->maxPack; D.SolidDiameter = p->SolidDiameter;
   D.ws0 = p->ws0; D.nsettling = p->nsettling; D.nu = p->EddyViscosity;
   D.Hneps = p->heightThreshold; D.cfl = p->cfl; D.diffusiveTimeScale = p->diffusiveTimeScale; D.maxdt = p->maxdt;

   h->tstate.assign(h->nTiles, 0); h->hasSource.assign(h->nTiles, 0); h->loaded.assign(h->nTiles, 0); h->seedFlags.assign(h->nTiles, 0);
   h->t = p->tstart;
   if (h->dyn) {
      h->gt.init(h->gnXt, h->gnYt, false, h->oneD, p->bcs == KGPU_BC_HALT);
      h->gSeed.assign(h->gt.nTiles(), 0); h->gSource.assign(h->gt.nTiles(), 0);
   }

   auto fail = [&](const char *what) { fprintf(stderr, "kgpu_create: %s: %s\n", what, cudaGetErrorString(cudaGetLastError())); kgpu_destroy(h); return KGPU_ERR_CUDA; };
   if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return fail("stream");
   cudaEventCreate(&h->evA); cudaEventCreate(&h->evB);
   size_t fb = h->fieldElems * sizeof(double);
   auto allocField = [&](double **ptr, double fillv) -> bool {
      if (cudaMalloc(ptr, fb) != cudaSuccess) return false;
      if (fillv == 0.0) return cudaMemsetAsync(*ptr, 0, fb, h->stream) == cudaSuccess;
      fill_kernel<<<(unsigned)((h->fieldElems + 255) / 256), 256, 0, h->stream>>>(*ptr, h->fieldElems, fillv);
      return true;
   };
   h->nStates = h->morpho ? 5 : 3;
   for (int k = 0; k < h->nStates; k++) for (int d = 0; d < 4; d++) if (!allocField(&h->S[k][d], 0.0)) return fail("state");
   for (int d = 0; d < 4; d++) if (!allocField(&h->E0[d], 0.0)) return fail("E0");
   if (!allocField(&h->I0, 0.0) || !allocField(&h->b0v, 0.0)) return fail("I0/b0v");
   if (h->morpho) {
      for (int d = 0; d < 4; d++) if (!allocField(&h->btv[d], 0.0)) return fail("btv");
      if (!allocField(&h->EmD, 0.0) || !allocField(&h->Um, 0.0) || !allocField(&h->Vm, 0.0)) return fail("EmD");
      for (int a = 0; a < 2; a++) for (int b = 0; b < 4; b++) if (!allocField(&h->mc[a][b], 0.0)) return fail("morphodynamic centre planes");
      if (!allocField(&h->mcHn0, 0.0)) return fail("morphodynamic centre planes");
   }
   for (int k = 0; k < 10; k++) if (!allocField(&h->mx[k], 0.0)) return fail("maxima");
   if (!allocField(&h->mx[10], -1.0)) return fail("tfirst");
   size_t msz = (size_t)(h->nXt + 2) * (h->nYt + 2);
   int nbx = (h->NX + BX - 1) / BX, nby = (h->NY + BY - 1) / BY;
   if (cudaMalloc(&h->d_tileMask, msz) != cudaSuccess || cudaMalloc(&h->d_tileSource, msz) != cudaSuccess) return fail("masks");
   if (cudaMalloc(&h->d_blockList, sizeof(int2) * (size_t)nbx * nby) != cudaSuccess) return fail("blocklist");
   if (h->morpho && cudaMalloc(&h->d_blockListM, sizeof(int2) * (size_t)nbx * nby) != cudaSuccess) return fail("blocklist");
   if (p->comm_size > 1 && (cudaMalloc(&h->d_blockBoundary, sizeof(int2) * (size_t)nbx * nby) != cudaSuccess ||
                            cudaMalloc(&h->d_blockInterior, sizeof(int2) * (size_t)nbx * nby) != cudaSuccess)) return fail("blocklists");
   if (cudaMalloc(&h->d_ctrl, sizeof(Ctrl)) != cudaSuccess || cudaMallocHost(&h->h_ctrl, sizeof(Ctrl)) != cudaSuccess) return fail("ctrl");
   cudaMemsetAsync(h->d_ctrl, 0, sizeof(Ctrl), h->stream);
   std::memset(h->h_ctrl, 0, sizeof(Ctrl));
   if (cudaMalloc(&h->d_sources, sizeof(DevSource) * std::max<size_t>(1, h->src.size())) != cudaSuccess) return fail("sources");
   if (!h->src.empty()) cudaMemcpyAsync(h->d_sources, h->src.data(), sizeof(DevSource) * h->src.size(), cudaMemcpyHostToDevice, h->stream);
   if (cudaMalloc(&h->d_srcPool, sizeof(double) * std::max<size_t>(1, h->srcPool.size())) != cudaSuccess) return fail("source series");
   if (!h->srcPool.empty()) cudaMemcpyAsync(h->d_srcPool, h->srcPool.data(), sizeof(double) * h->srcPool.size(), cudaMemcpyHostToDevice, h->stream);
   size_t ncell = (size_t)h->nX * h->nY, nv = (size_t)(h->nX + 1) * (h->nY + 1);
   h->stageElems = 24 * ncell + 2 * nv;
   if (cudaMalloc(&h->d_stage, h->stageElems * sizeof(double)) != cudaSuccess || cudaMallocHost(&h->h_stage, h->stageElems * sizeof(double)) != cudaSuccess) return fail("stage");
   if (cudaMalloc(&h->d_tileList, sizeof(int) * h->nTiles) != cudaSuccess || cudaMalloc(&h->d_flags, sizeof(int) * h->nTiles) != cudaSuccess ||
       cudaMallocHost(&h->h_flags, sizeof(int) * h->nTiles) != cudaSuccess) return fail("flags");
   if (h->dyn) {
      size_t nb = sizeof(int) * (size_t)h->gt.nTiles();
      if (cudaMalloc(&h->d_gflags, nb) != cudaSuccess || cudaMallocHost(&h->h_gflags, nb) != cudaSuccess) return fail("global flags");
      std::memset(h->h_gflags, 0, nb);
   }
   if (h->morpho) {
      h->redistCap = 1 << 16;
      if (cudaMalloc(&h->d_redist, sizeof(RedistEntry) * h->redistCap) != cudaSuccess ||
          cudaMallocHost(&h->h_redist, sizeof(RedistEntry) * h->redistCap) != cudaSuccess) return fail("redist");
   }
   // opt in to > 48 KB dynamic shared memory for the stage kernel
   setStageAttr<false, false, KGPU_LIM_MINMOD2, 0>(); setStageAttr<false, false, -1, 0>(); setStageAttr<false, true, KGPU_LIM_MINMOD2, 0>(); setStageAttr<false, true, -1, 0>();
   setStageAttr<false, false, KGPU_LIM_MINMOD2, 1>(); setStageAttr<false, false, -1, 1>(); setStageAttr<false, true, KGPU_LIM_MINMOD2, 1>(); setStageAttr<false, true, -1, 1>();
   setStageAttr<true, false, KGPU_LIM_MINMOD2, 0>(); setStageAttr<true, false, -1, 0>(); setStageAttr<true, true, KGPU_LIM_MINMOD2, 0>(); setStageAttr<true, true, -1, 0>();
   stage_fast_set_attributes();
   // topography planes (bt planes only when the bed moves)
   {
      double **pl[15] = {&h->topo.b0c, &h->topo.bxc, &h->topo.byc, &h->topo.gamc, &h->topo.xb0, &h->topo.xB, &h->topo.xtan, &h->topo.xgam,
                         &h->topo.yb0, &h->topo.yB, &h->topo.ytan, &h->topo.ygam, &h->topo.btc, &h->topo.xbt, &h->topo.ybt};
      int npl = h->morpho ? 15 : 12;
      for (int k = 0; k < npl; k++) if (!allocField(pl[k], 0.0)) return fail("topography planes");
   }
   if (!buildTensorMaps(h)) return fail("tensor maps (cuTensorMapEncodeTiled)");
   {
      cudaDeviceProp prop;
      int nsm = 148;
      if (cudaGetDeviceProperties(&prop, h->dev) == cudaSuccess) nsm = prop.multiProcessorCount;
      h->prefetchDistance = 3 * nsm;
      if (const char *e = std::getenv("KGPU_PREFETCH_DISTANCE")) h->prefetchDistance = std::atoi(e);  // tuning knob
      h->tune = 31;  // measured on B200 at 4096^2 (round 1): bit 1 +3.9 %, bit 2 +1.8 %, bit 3 +0.5 %, bit 0 +-0, bit 4 +2.2 %; all five +7.9 %
      if (const char *e = std::getenv("KGPU_TUNE")) h->tune = std::atoi(e);                          // tuning knob (StageArgs::tune)
      h->useSpec = !(h->tune & 64);
      h->morphoFusion = (h->tune & 256) ? 2 : (h->tune & 128) ? 1 : 0;
   }
   if (cudaStreamSynchronize(h->stream) != cudaSuccess) return fail("init sync");
   *out = h;
   return KGPU_OK;
}

int kgpu_upload_tile(kgpu_handle *h, int32_t tile_id, const double *u13, const double *b0_vertices, const double *bt_vertices,
                     const double *maxima, const double *tfirst, int32_t contains_source) {
   if (!h || !u13) return KGPU_ERR_ARG;
   cudaSetDevice(h->dev);
   if (h->dyn) {
      if (!h->comm.active) { h->err = "call kgpu_comm_attach before uploading"; return KGPU_ERR_ARG; }
      return dynUploadTile(h, tile_id, u13, b0_vertices, bt_vertices, maxima, tfirst, contains_source);
   }
   int t0 = localTile0(h, tile_id);
   if (t0 < 0) { h->err = t0 == -1 ? "tile id out of range" : "tile belongs to another rank's block (kgpu_comm_block)"; return KGPU_ERR_ARG; }
   int rc;
   if (b0_vertices && (rc = loadHeights(h, t0, b0_vertices))) return rc;
   if ((rc = addTile(h, t0, false))) return rc;
   if (h->tstate[t0] != 2) { h->err = "tile lies on the domain edge and cannot be active"; return KGPU_ERR_ARG; }
   h->hasSource[t0] = contains_source ? 1 : 0;
   h->masksDirty = true;
   int nX = h->nX, nY = h->nY;
   size_t ncell = (size_t)nX * nY;
   int tx, ty; tileXY(h, t0, tx, ty);
   if (bt_vertices && h->morpho) {
      size_t nv = (size_t)(nX + 1) * (h->oneD ? 1 : nY + 1);
      std::memcpy(h->h_stage, bt_vertices, nv * sizeof(double));
      CUDA_TRY(h, cudaMemcpyAsync(h->d_stage, h->h_stage, nv * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      dim3 gridv((nX + 1 + 127) / 128, h->oneD ? 1 : nY + 1);
      tile_vertices_kernel<<<gridv, 128, 0, h->stream>>>(h->D, h->btv[h->bt0], h->d_stage, tx, ty, 1, 7);
      h->launches++;
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      if ((rc = fillHaloVertices(h, h->btv[h->bt0]))) return rc;
   }
   std::memcpy(h->h_stage, u13, ncell * 13 * sizeof(double));
   if (maxima) std::memcpy(h->h_stage + ncell * 13, maxima, ncell * 10 * sizeof(double));
   if (tfirst) std::memcpy(h->h_stage + ncell * 23, tfirst, ncell * sizeof(double));
   CUDA_TRY(h, cudaMemcpyAsync(h->d_stage, h->h_stage, ncell * 24 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
   dim3 grid((nX + 127) / 128, nY);
   import_tile_kernel<<<grid, 128, 0, h->stream>>>(h->D, h->sp(h->i0), h->mp(), h->d_stage, tx, ty, maxima ? 1 : 0, tfirst ? 1 : 0);
   h->launches++;
   CUDA_TRY(h, cudaGetLastError());
   // seed of the first tile-activation scan: the host's u(Hn) (TimeStepper.f90:982)
   int buf = h->P.TileBuffer, f = 0;
   for (int lj = 0; lj < nY; lj++)
      for (int li = 0; li < nX; li++) {
         if (u13[((size_t)lj * nX + li) * 13 + 4] > h->P.heightThreshold) {
            if (!h->oneD) { if (lj >= nY - buf) f |= 1; if (lj < buf) f |= 2; }
            if (li >= nX - buf) f |= 4;
            if (li < buf) f |= 8;
         }
      }
   h->seedFlags[t0] = f;
   h->havePre = false;
   CUDA_TRY(h, cudaStreamSynchronize(h->stream));
   return KGPU_OK;
}

int kgpu_upload_domain(kgpu_handle *h, const double *q4, const double *b0_vertices, const double *bt_vertices) {
   if (!h || !q4 || !b0_vertices) return KGPU_ERR_ARG;
   if (!h->globalPeriodic) { h->err = "kgpu_upload_domain needs Boundary Conditions = periodic (UpdateTiles.f90:61-69)"; return KGPU_ERR_ARG; }
   if (h->P.comm_size > 1 && !h->comm.active) { h->err = "call kgpu_comm_attach before uploading"; return KGPU_ERR_ARG; }
   cudaSetDevice(h->dev);
   size_t nc = (size_t)h->NX * h->NY;
   size_t dp = (size_t)h->pitch * sizeof(double);
   for (int d = 0; d < 4; d++)
      CUDA_TRY(h, cudaMemcpy2DAsync(h->S[h->i0][d] + (size_t)YO * h->pitch + XO, dp, q4 + d * nc, (size_t)h->NX * sizeof(double),
                                    (size_t)h->NX * sizeof(double), h->NY, cudaMemcpyHostToDevice, h->stream));
   int nvy = h->oneD ? 1 : h->NY + 1;
   CUDA_TRY(h, cudaMemcpy2DAsync(h->b0v + (size_t)YO * h->pitch + XO, dp, b0_vertices, (size_t)(h->NX + 1) * sizeof(double),
                                 (size_t)(h->NX + 1) * sizeof(double), nvy, cudaMemcpyHostToDevice, h->stream));
   int rc;
   if ((rc = fillHaloVertices(h, h->b0v))) return rc;
   if (h->morpho && bt_vertices) {
      CUDA_TRY(h, cudaMemcpy2DAsync(h->btv[h->bt0] + (size_t)YO * h->pitch + XO, dp, bt_vertices, (size_t)(h->NX + 1) * sizeof(double),
                                    (size_t)(h->NX + 1) * sizeof(double), nvy, cudaMemcpyHostToDevice, h->stream));
      if ((rc = fillHaloVertices(h, h->btv[h->bt0]))) return rc;
   }
   h->activeList.clear(); h->ghostList.clear();
   for (int t0 = 0; t0 < h->nTiles; t0++) { h->tstate[t0] = 2; h->loaded[t0] = 1; h->activeList.push_back(t0 + 1); }
   // containsSource: a tile with a cell centre inside a source disc, <= as at load (SetSources.f90:367-372);
   // cell coordinates are the global ones of Grid.f90:339-353 (the device's cellX / cellY)
   std::fill(h->hasSource.begin(), h->hasSource.end(), 0);
   for (const DevSource &S : h->src)
      for (int j = 0; j < h->NY; j++) {
         const int gj = h->gty0 + j / h->nY + 1, tj = j % h->nY + 1;
         const double y = -0.5 * h->P.ySize + h->P.deltaY * ((gj - 1.0) * h->nY + (tj - 0.5));
         const double dy2 = h->oneD ? 0.0 : (y - S.y) * (y - S.y);
         if (dy2 > S.radius * S.radius) continue;
         for (int i = 0; i < h->NX; i++) {
            const int gi = h->gtx0 + i / h->nX + 1, ti = i % h->nX + 1;
            const double x = -0.5 * h->P.xSize + h->P.deltaX * ((gi - 1.0) * h->nX + (ti - 0.5));
            const double R2 = h->oneD ? (x - S.x) * (x - S.x) : (x - S.x) * (x - S.x) + dy2;
            if (R2 <= S.radius * S.radius) h->hasSource[(j / h->nY) * h->nXt + i / h->nX] = 1;
         }
      }
   h->masksDirty = true;
   h->firstScan = false;
   h->havePre = false;
   if ((rc = refreshMasks(h))) return rc;
   if ((rc = fillHaloCells(h, h->i0))) return rc;
   CUDA_TRY(h, cudaStreamSynchronize(h->stream));
   return KGPU_OK;
}

// LoadSourceConditions (SetSources.f90:47-392) on the device: see include/kestrel_gpu.h
int kgpu_load_source_conditions(kgpu_handle *h, const kgpu_cap *caps, int32_t ncaps, const kgpu_cube *cubes, int32_t ncubes,
                                int32_t *num_cells_in_src) {
   if (!h || ncaps < 0 || ncubes < 0 || (ncaps > 0 && !caps) || (ncubes > 0 && !cubes)) return KGPU_ERR_ARG;
   if (h->comm.active) { h->err = "kgpu_load_source_conditions: single device only (decomposed runs upload their blocks)"; return KGPU_ERR_UNSUPPORTED; }
   cudaSetDevice(h->dev);
   const int nsrc = (int)h->src.size();
   kgpu_cap *dCaps = nullptr;
   kgpu_cube *dCubes = nullptr;
   int *dTouch = nullptr, *dCount = nullptr;
   std::vector<int> touch(h->nTiles), counts(std::max(1, nsrc), 0);
   auto cleanup = [&]() { cudaFree(dCaps); cudaFree(dCubes); cudaFree(dTouch); cudaFree(dCount); };
#define LSC_TRY(call)                                                                   \
   do {                                                                                 \
      cudaError_t e_ = (call);                                                          \
      if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); cleanup(); return KGPU_ERR_CUDA; } \
   } while (0)
   LSC_TRY(cudaMalloc(&dCaps, sizeof(kgpu_cap) * std::max(1, ncaps)));
   LSC_TRY(cudaMalloc(&dCubes, sizeof(kgpu_cube) * std::max(1, ncubes)));
   LSC_TRY(cudaMalloc(&dTouch, sizeof(int) * h->nTiles));
   LSC_TRY(cudaMalloc(&dCount, sizeof(int) * std::max(1, nsrc)));
   if (ncaps) LSC_TRY(cudaMemcpyAsync(dCaps, caps, sizeof(kgpu_cap) * ncaps, cudaMemcpyHostToDevice, h->stream));
   if (ncubes) LSC_TRY(cudaMemcpyAsync(dCubes, cubes, sizeof(kgpu_cube) * ncubes, cudaMemcpyHostToDevice, h->stream));
   LSC_TRY(cudaMemsetAsync(dCount, 0, sizeof(int) * std::max(1, nsrc), h->stream));
   ShapeTable T{dCaps, dCubes, h->d_sources, ncaps, ncubes, nsrc};
   // pass 1: which tiles does a shape reach?
   shape_touch_kernel<<<h->nTiles, 128, 0, h->stream>>>(h->D, T, dTouch);
   h->launches++;
   LSC_TRY(cudaMemcpyAsync(touch.data(), dTouch, sizeof(int) * h->nTiles, cudaMemcpyDeviceToHost, h->stream));
   LSC_TRY(cudaStreamSynchronize(h->stream));
   // AddTile in the reference's loop order: do i = 1, nXtiles; do j = 1, nYtiles (SetSources.f90:96, 229-231)
   for (int tx = 0; tx < h->nXt; tx++)
      for (int ty = 0; ty < h->nYt; ty++) {
         const int t0 = ty * h->nXt + tx;
         if (!touch[t0]) continue;
         int rc = addTile(h, t0, false);
         if (rc) { cleanup(); return rc; }
         if (h->tstate[t0] == 2) h->hasSource[t0] = (touch[t0] & 2) ? 1 : 0;   // sponge / dirichlet edge tiles stay off silently
      }
   h->masksDirty = true;
   const int nAct = (int)h->activeList.size();
   if (nAct > 0) {
      std::vector<int> tl(nAct);
      for (int k = 0; k < nAct; k++) tl[k] = h->activeList[k] - 1;
      LSC_TRY(cudaMemcpyAsync(h->d_tileList, tl.data(), nAct * sizeof(int), cudaMemcpyHostToDevice, h->stream));
      // pass 2: the shapes, cell by cell
      shape_raster_kernel<<<nAct, 128, 0, h->stream>>>(h->D, T, h->sp(h->i0), h->mp(), h->b0v, h->d_tileList, h->P.TileBuffer, h->d_flags, dCount);
      h->launches++;
      LSC_TRY(cudaMemcpyAsync(h->h_flags, h->d_flags, nAct * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      LSC_TRY(cudaMemcpyAsync(counts.data(), dCount, sizeof(int) * std::max(1, nsrc), cudaMemcpyDeviceToHost, h->stream));
      LSC_TRY(cudaStreamSynchronize(h->stream));
      for (int k = 0; k < nAct; k++) h->seedFlags[tl[k]] = h->h_flags[k];
   }
   // NumCellsInSrc (SetSources.f90:372): every cell of the domain counts, and every such cell's tile is active by now
   for (int s_ = 0; s_ < nsrc; s_++) {
      h->src[s_].numCells = counts[s_];
      if (num_cells_in_src) num_cells_in_src[s_] = counts[s_];
   }
   if (nsrc) LSC_TRY(cudaMemcpyAsync(h->d_sources, h->src.data(), sizeof(DevSource) * nsrc, cudaMemcpyHostToDevice, h->stream));
   h->havePre = false;
   h->firstScan = true;
   LSC_TRY(cudaStreamSynchronize(h->stream));
#undef LSC_TRY
   cleanup();
   return KGPU_OK;
}

int kgpu_integrate_to(kgpu_handle *h, double tend, int64_t max_steps, kgpu_step_info *info) {
   if (!h) return KGPU_ERR_ARG;
   cudaSetDevice(h->dev);
   return integrateTo(h, tend, max_steps, info);
}

int kgpu_active_tiles(kgpu_handle *h, int32_t *n, int32_t *ids) {
   if (!h || !n) return KGPU_ERR_ARG;
   *n = (int32_t)h->activeList.size();
   if (ids) for (size_t k = 0; k < h->activeList.size(); k++) ids[k] = globalTileId(h, h->activeList[k] - 1);
   return KGPU_OK;
}
int kgpu_ghost_tiles(kgpu_handle *h, int32_t *n, int32_t *ids) {
   if (!h || !n) return KGPU_ERR_ARG;
   *n = (int32_t)h->ghostList.size();
   if (ids) for (size_t k = 0; k < h->ghostList.size(); k++) ids[k] = globalTileId(h, h->ghostList[k] - 1);
   return KGPU_OK;
}

int kgpu_download_tile(kgpu_handle *h, int32_t tile_id, double *u13, double *b0_vertices, double *bt_vertices, double *maxima, double *tfirst) {
   if (!h) return KGPU_ERR_ARG;
   cudaSetDevice(h->dev);
   int t0 = localTile0(h, tile_id);
   if (t0 < 0) { h->err = t0 == -1 ? "tile id out of range" : "tile belongs to another rank's block (kgpu_comm_block)"; return KGPU_ERR_ARG; }
   int nX = h->nX, nY = h->nY;
   size_t ncell = (size_t)nX * nY, nv = (size_t)(nX + 1) * (nY + 1), nvu = (size_t)(nX + 1) * (h->oneD ? 1 : nY + 1);
   int tx, ty; tileXY(h, t0, tx, ty);
   dim3 grid((nX + 127) / 128, nY);
   export_tile_kernel<<<grid, 128, 0, h->stream>>>(h->D, h->sp(h->i0), h->sp(h->ia), h->havePre ? 1 : 0, h->b0v,
                                                   h->morpho ? h->btv[h->bt0] : nullptr, h->mp(), h->d_stage, tx, ty);
   dim3 gridv((nX + 1 + 127) / 128, h->oneD ? 1 : nY + 1);
   tile_vertices_kernel<<<gridv, 128, 0, h->stream>>>(h->D, h->b0v, h->d_stage + 24 * ncell, tx, ty, 0, 0);
   if (h->morpho) tile_vertices_kernel<<<gridv, 128, 0, h->stream>>>(h->D, h->btv[h->bt0], h->d_stage + 24 * ncell + nv, tx, ty, 0, 0);
   else CUDA_TRY(h, cudaMemsetAsync(h->d_stage + 24 * ncell + nv, 0, nv * sizeof(double), h->stream));
   h->launches += h->morpho ? 3 : 2;
   CUDA_TRY(h, cudaMemcpyAsync(h->h_stage, h->d_stage, h->stageElems * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
   CUDA_TRY(h, cudaStreamSynchronize(h->stream));
   if (u13) std::memcpy(u13, h->h_stage, ncell * 13 * sizeof(double));
   if (maxima) std::memcpy(maxima, h->h_stage + ncell * 13, ncell * 10 * sizeof(double));
   if (tfirst) std::memcpy(tfirst, h->h_stage + ncell * 23, ncell * sizeof(double));
   if (b0_vertices) std::memcpy(b0_vertices, h->h_stage + 24 * ncell, nvu * sizeof(double));
   if (bt_vertices) std::memcpy(bt_vertices, h->h_stage + 24 * ncell + nv, nvu * sizeof(double));
   return KGPU_OK;
}

int kgpu_download_domain(kgpu_handle *h, double *q4, double *bt_vertices) {
   if (!h) return KGPU_ERR_ARG;
   cudaSetDevice(h->dev);
   size_t nc = (size_t)h->NX * h->NY;
   size_t dp = (size_t)h->pitch * sizeof(double);
   if (q4)
      for (int d = 0; d < 4; d++)
         CUDA_TRY(h, cudaMemcpy2DAsync(q4 + d * nc, (size_t)h->NX * sizeof(double), h->S[h->i0][d] + (size_t)YO * h->pitch + XO, dp,
                                       (size_t)h->NX * sizeof(double), h->NY, cudaMemcpyDeviceToHost, h->stream));
   if (bt_vertices) {
      int nvy = h->oneD ? 1 : h->NY + 1;
      if (h->morpho)
         CUDA_TRY(h, cudaMemcpy2DAsync(bt_vertices, (size_t)(h->NX + 1) * sizeof(double), h->btv[h->bt0] + (size_t)YO * h->pitch + XO, dp,
                                       (size_t)(h->NX + 1) * sizeof(double), nvy, cudaMemcpyDeviceToHost, h->stream));
      else std::memset(bt_vertices, 0, sizeof(double) * (size_t)(h->NX + 1) * nvy);
   }
   CUDA_TRY(h, cudaStreamSynchronize(h->stream));
   return KGPU_OK;
}

// Asynchronous output gather: snapshot on the compute stream (device copy), transfer on a copy stream.
int kgpu_output_wait(kgpu_handle *h) {
   if (!h) return KGPU_ERR_ARG;
   if (!h->outputPending) return KGPU_OK;
   cudaSetDevice(h->dev);
   CUDA_TRY(h, cudaEventSynchronize(h->evCopied));
   h->outputPending = false;
   return KGPU_OK;
}

int kgpu_output_begin(kgpu_handle *h, double *q4, double *bt_vertices) {
   if (!h || !q4) return KGPU_ERR_ARG;
   cudaSetDevice(h->dev);
   int rc = kgpu_output_wait(h);   // one output in flight: the snapshot planes are about to be rewritten
   if (rc) return rc;
   if (!h->copyStream) {
      CUDA_TRY(h, cudaStreamCreateWithFlags(&h->copyStream, cudaStreamNonBlocking));
      CUDA_TRY(h, cudaEventCreateWithFlags(&h->evSnap, cudaEventDisableTiming));
      CUDA_TRY(h, cudaEventCreateWithFlags(&h->evCopied, cudaEventDisableTiming | cudaEventBlockingSync));
   }
   const bool wantBt = bt_vertices && h->morpho;
   size_t fb = h->fieldElems * sizeof(double);
   for (int k = 0; k < (wantBt ? 5 : 4); k++)
      if (!h->snap[k]) CUDA_TRY(h, cudaMalloc(&h->snap[k], fb));
   // the state buffer of this step is recycled two steps from now: keep a copy (HBM speed, a few ms at 16384^2)
   for (int d = 0; d < 4; d++) CUDA_TRY(h, cudaMemcpyAsync(h->snap[d], h->S[h->i0][d], fb, cudaMemcpyDeviceToDevice, h->stream));
   if (wantBt) CUDA_TRY(h, cudaMemcpyAsync(h->snap[4], h->btv[h->bt0], fb, cudaMemcpyDeviceToDevice, h->stream));
   CUDA_TRY(h, cudaEventRecord(h->evSnap, h->stream));
   CUDA_TRY(h, cudaStreamWaitEvent(h->copyStream, h->evSnap, 0));
   size_t nc = (size_t)h->NX * h->NY;
   size_t dp = (size_t)h->pitch * sizeof(double);
   for (int d = 0; d < 4; d++)
      CUDA_TRY(h, cudaMemcpy2DAsync(q4 + d * nc, (size_t)h->NX * sizeof(double), h->snap[d] + (size_t)YO * h->pitch + XO, dp,
                                    (size_t)h->NX * sizeof(double), h->NY, cudaMemcpyDeviceToHost, h->copyStream));
   if (bt_vertices) {
      int nvy = h->oneD ? 1 : h->NY + 1;
      if (wantBt)
         CUDA_TRY(h, cudaMemcpy2DAsync(bt_vertices, (size_t)(h->NX + 1) * sizeof(double), h->snap[4] + (size_t)YO * h->pitch + XO, dp,
                                       (size_t)(h->NX + 1) * sizeof(double), nvy, cudaMemcpyDeviceToHost, h->copyStream));
      else std::memset(bt_vertices, 0, sizeof(double) * (size_t)(h->NX + 1) * nvy);
   }
   CUDA_TRY(h, cudaEventRecord(h->evCopied, h->copyStream));
   h->outputPending = true;
   return KGPU_OK;
}

int kgpu_set_topography_function(kgpu_handle *h, int32_t func, const double *params, int32_t nparams) {
   if (!h || nparams < 0 || nparams > 8 || (nparams > 0 && !params) || func > KGPU_TOPOG_XTRISLOPE) return KGPU_ERR_ARG;
   static const int need[] = {0, 1, 1, 2, 1, 1, 2, 3, 1, 2, 3, 3, 2, 6, 3, 3, 6};   // parameters each function reads (TopogFuncs.f90)
   if (func >= 0 && nparams < need[func]) { h->err = "too few topography parameters"; return KGPU_ERR_ARG; }
   h->topogFn.func = func; h->topogFn.n = nparams;
   for (int k = 0; k < 8; k++) h->topogFn.p[k] = k < nparams ? params[k] : 0.0;
   return KGPU_OK;
}

int kgpu_set_topography_raster(kgpu_handle *h, const double *elev, int32_t nx, int32_t ny, double origin_x, double origin_y,
                               double pixel_w, double pixel_h, double centre_e, double centre_n) {
   if (!h || !elev || nx < 2 || ny < 2 || pixel_w == 0.0 || pixel_h == 0.0) return KGPU_ERR_ARG;
   cudaSetDevice(h->dev);
   cudaFree(h->d_raster);
   h->d_raster = nullptr; h->raster.elev = nullptr;
   CUDA_TRY(h, cudaMalloc(&h->d_raster, sizeof(double) * (size_t)nx * ny));
   CUDA_TRY(h, cudaMemcpy(h->d_raster, elev, sizeof(double) * (size_t)nx * ny, cudaMemcpyHostToDevice));
   h->raster = RasterDesc{h->d_raster, nx, ny, origin_x, origin_y, pixel_w, pixel_h, centre_e, centre_n};
   return KGPU_OK;
}

// Test probe of the host bookkeeping of RedistributeGrid across ranks (kgpu_redist_tables.hpp): no device needed.
int kgpu_debug_redist_tables(const int32_t *geometry10, const int32_t *counts, const double *excess, const int32_t *li, const int32_t *lj,
                             int32_t *n_out, int32_t *patch, int32_t *vslot, int32_t *cslot, int32_t *n_unique2) {
   if (!geometry10 || !counts || !excess || !li || !lj || !n_out) return KGPU_ERR_ARG;
   RedistGeometry g{geometry10[0], geometry10[1], geometry10[2], geometry10[3], geometry10[4], geometry10[5], geometry10[6], geometry10[7],
                    geometry10[8], geometry10[9]};
   RedistTables T;
   buildRedistTables(g, counts, excess, li, lj, T);
   *n_out = (int32_t)T.patch.size();
   if (patch) std::copy(T.patch.begin(), T.patch.end(), patch);
   if (vslot) std::copy(T.vslot.begin(), T.vslot.end(), vslot);
   if (cslot) std::copy(T.cslot.begin(), T.cslot.end(), cslot);
   if (n_unique2) { n_unique2[0] = (int32_t)T.vbase.size(); n_unique2[1] = (int32_t)T.cbase.size(); }
   return KGPU_OK;
}

// Test probes of the replicated tile table (kgpu_tile_table.hpp): no device needed.
struct kgpu_tiletable { TileTable T; };
kgpu_tiletable *kgpu_debug_tiletable_new(int32_t nXtiles, int32_t nYtiles, int32_t periodic, int32_t isOneD, int32_t halt_bc) {
   if (nXtiles < 1 || nYtiles < 1) return nullptr;
   kgpu_tiletable *t = new kgpu_tiletable;
   t->T.init(nXtiles, nYtiles, periodic != 0, isOneD != 0, halt_bc != 0);
   return t;
}
void kgpu_debug_tiletable_free(kgpu_tiletable *t) { delete t; }
int kgpu_debug_tiletable_add(kgpu_tiletable *t, int32_t tile_id) {
   if (!t) return KGPU_ERR_ARG;
   bool ok = t->T.addTile(tile_id - 1, false, true);
   return ok ? KGPU_OK : (t->T.haltViolation ? KGPU_ERR_HALT_BC : KGPU_ERR_ARG);
}
int kgpu_debug_tiletable_replay(kgpu_tiletable *t, const int32_t *flags, int32_t nXpertile, int32_t nYpertile, int32_t tile_buffer) {
   if (!t || !flags) return KGPU_ERR_ARG;
   std::vector<int> f(flags, flags + t->T.nTiles());
   bool ok = t->T.replay(f.data(), nXpertile, nYpertile, tile_buffer);
   return ok ? KGPU_OK : (t->T.haltViolation ? KGPU_ERR_HALT_BC : KGPU_ERR_ARG);
}
int kgpu_debug_tiletable_lists(const kgpu_tiletable *t, int32_t *n_active, int32_t *active, int32_t *n_ghost, int32_t *ghost,
                               int64_t *n_added, int32_t *n_ops) {
   if (!t) return KGPU_ERR_ARG;
   if (n_active) *n_active = (int32_t)t->T.activeList.size();
   if (active) std::copy(t->T.activeList.begin(), t->T.activeList.end(), active);
   if (n_ghost) *n_ghost = (int32_t)t->T.ghostList.size();
   if (ghost) std::copy(t->T.ghostList.begin(), t->T.ghostList.end(), ghost);
   if (n_added) *n_added = t->T.ntilesAdded;
   if (n_ops) *n_ops = (int32_t)t->T.ops.size();
   return KGPU_OK;
}

// One evaluation of CalculateHydraulicRHS on the current state (parity probe for K1).
int kgpu_debug_rhs(kgpu_handle *h, int32_t substep, double *E4, double *I, double *dt) {
   if (!h) return KGPU_ERR_ARG;
   cudaSetDevice(h->dev);
   int rc;
   if ((rc = refreshMasks(h))) return rc;
   if ((rc = fillHaloCells(h, h->i0))) return rc;
   if ((rc = firstRHS(h, h->i0, h->bt0, h->t, HUGE_D, 0))) return rc;
   if ((rc = readCtrl(h))) return rc;
   size_t nc = (size_t)h->NX * h->NY;
   size_t dp = (size_t)h->pitch * sizeof(double);
   if (E4)
      for (int d = 0; d < 4; d++)
         CUDA_TRY(h, cudaMemcpy2DAsync(E4 + d * nc, (size_t)h->NX * sizeof(double), h->E0[d] + (size_t)YO * h->pitch + XO, dp,
                                       (size_t)h->NX * sizeof(double), h->NY, cudaMemcpyDeviceToHost, h->stream));
   if (I)
      CUDA_TRY(h, cudaMemcpy2DAsync(I, (size_t)h->NX * sizeof(double), h->I0 + (size_t)YO * h->pitch + XO, dp,
                                    (size_t)h->NX * sizeof(double), h->NY, cudaMemcpyDeviceToHost, h->stream));
   CUDA_TRY(h, cudaStreamSynchronize(h->stream));
   if (dt) *dt = (substep == 1) ? h->h_ctrl->dtAdvised : h->h_ctrl->dtAdvised / 0.9;
   return KGPU_OK;
}

int kgpu_debug_sequential_walk(kgpu_handle *h, int32_t on) {
   if (!h) return KGPU_ERR_ARG;
   h->debugSequentialWalk = on != 0;
   return KGPU_OK;
}
int kgpu_debug_morpho_fusion(kgpu_handle *h, int32_t level) {
   if (!h || level < 0 || level > 2) return KGPU_ERR_ARG;
   h->morphoFusion = level;
   return KGPU_OK;
}
int kgpu_debug_global_walk(kgpu_handle *h, int32_t on) {
   if (!h) return KGPU_ERR_ARG;
   h->debugGlobalWalk = on != 0;
   return KGPU_OK;
}
int kgpu_debug_redist_capacity(kgpu_handle *h, int32_t entries) {
   if (!h || entries < 1 || !h->morpho) return KGPU_ERR_ARG;
   cudaSetDevice(h->dev);
   cudaFree(h->d_redist); cudaFreeHost(h->h_redist);
   h->d_redist = nullptr; h->h_redist = nullptr;
   h->redistCap = entries;
   CUDA_TRY(h, cudaMalloc(&h->d_redist, sizeof(RedistEntry) * (size_t)entries));
   CUDA_TRY(h, cudaMallocHost(&h->h_redist, sizeof(RedistEntry) * (size_t)entries));
   return KGPU_OK;
}
int kgpu_morpho_stats(const kgpu_handle *h, int64_t *redistributed_cells, int64_t *list_enlargements) {
   if (!h) return KGPU_ERR_ARG;
   if (redistributed_cells) *redistributed_cells = h->nRedistCells;
   if (list_enlargements) *list_enlargements = h->nRedistGrows;
   return KGPU_OK;
}

int kgpu_comm_id_bytes(void) { return (int)sizeof(ncclUniqueId); }

int kgpu_comm_create_id(void *id_out) {
   std::string err;
   if (!id_out || !loadNccl(err)) { fprintf(stderr, "kgpu_comm_create_id: %s\n", err.c_str()); return KGPU_ERR_CUDA; }
   ncclUniqueId id;
   if (g_nccl.GetUniqueId(&id) != ncclSuccess) return KGPU_ERR_CUDA;
   std::memcpy(id_out, &id, sizeof(id));
   return KGPU_OK;
}

int kgpu_comm_attach(kgpu_handle *h, const void *id_in) {
   if (!h || !id_in) return KGPU_ERR_ARG;
   if (h->P.comm_size <= 1) { h->err = "params.comm_size <= 1: nothing to attach"; return KGPU_ERR_ARG; }
   cudaSetDevice(h->dev);
   if (!loadNccl(h->err)) return KGPU_ERR_CUDA;
   ncclUniqueId id;
   std::memcpy(&id, id_in, sizeof(id));
   ncclComm_t comm;
   NCCL_TRY(h, g_nccl.CommInitRank(&comm, h->comm.size, id, h->comm.rank));
   h->comm.nccl = comm;
   size_t nx = (size_t)8 * (h->NY + 1) * 2, ny = (size_t)8 * (h->NX + 5) * 2;   // up to 8 fields per exchange
   for (int k = 0; k < 4; k++) {
      size_t n = (k < 2 ? nx : ny) * sizeof(double);
      CUDA_TRY(h, cudaMalloc(&h->comm.sendBuf[k], n));
      CUDA_TRY(h, cudaMalloc(&h->comm.recvBuf[k], n));
   }
   CUDA_TRY(h, cudaStreamCreateWithFlags(&h->comm.stream, cudaStreamNonBlocking));
   CUDA_TRY(h, cudaEventCreateWithFlags(&h->comm.evBoundary, cudaEventDisableTiming));
   CUDA_TRY(h, cudaEventCreateWithFlags(&h->comm.evHalo, cudaEventDisableTiming));
   h->comm.active = true;
   h->masksDirty = true;
   return KGPU_OK;
}

int kgpu_comm_block(kgpu_handle *h, int32_t *tx0, int32_t *ty0, int32_t *ntx, int32_t *nty) {
   if (!h) return KGPU_ERR_ARG;
   if (tx0) *tx0 = h->gtx0;
   if (ty0) *ty0 = h->gty0;
   if (ntx) *ntx = h->nXt;
   if (nty) *nty = h->nYt;
   return KGPU_OK;
}

}  // extern "C"
