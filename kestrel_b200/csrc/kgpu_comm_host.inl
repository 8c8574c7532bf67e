// kgpu_comm_host.inl -- NCCL plumbing of the 2-D decomposition; included by kestrel_gpu.cu.
//
// libnccl is resolved at run time (dlopen of the soname already loaded by the host process,
// e.g. the one torch ships) so that single-GPU users need no NCCL at all.
#include <dlfcn.h>
#if __has_include(<nccl.h>) && !defined(KGPU_NO_NCCL_HEADER)
#include <nccl.h>
#else
// No NCCL headers on the build machine: the handful of declarations the plumbing needs, with the values of NCCL 2.x's
// public ABI (nccl.h: ncclResult_t, ncclDataType_t, ncclRedOp_t, NCCL_UNIQUE_ID_BYTES).  The library itself is still
// only resolved at run time.
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclChar = 0, ncclInt = 2, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
#endif

struct NcclApi {
   void *lib = nullptr;
   ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
   ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
   ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
   ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*GroupStart)() = nullptr;
   ncclResult_t (*GroupEnd)() = nullptr;
   const char *(*GetErrorString)(ncclResult_t) = nullptr;
   bool ok = false;
};
static NcclApi g_nccl;

static bool loadNccl(std::string &err) {
   if (g_nccl.ok) return true;
   const char *names[] = {"libnccl.so.2", "libnccl.so"};
   for (const char *n : names) {
      g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (g_nccl.lib) break;
   }
   if (!g_nccl.lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define KSYM(field, name)                                                                  \
   g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.lib, name));       \
   if (!g_nccl.field) { err = std::string("libnccl lacks ") + name; return false; }
   KSYM(GetUniqueId, "ncclGetUniqueId") KSYM(CommInitRank, "ncclCommInitRank") KSYM(CommDestroy, "ncclCommDestroy")
   KSYM(Send, "ncclSend") KSYM(Recv, "ncclRecv") KSYM(AllReduce, "ncclAllReduce") KSYM(AllGather, "ncclAllGather") KSYM(GroupStart, "ncclGroupStart")
   KSYM(GroupEnd, "ncclGroupEnd") KSYM(GetErrorString, "ncclGetErrorString")
#undef KSYM
   g_nccl.ok = true;
   return true;
}

#define NCCL_TRY(h, call)                                                                  \
   do {                                                                                    \
      ncclResult_t r_ = (call);                                                            \
      if (r_ != ncclSuccess) {                                                             \
         (h)->err = std::string(#call) + ": " + g_nccl.GetErrorString(r_);                 \
         return KGPU_ERR_CUDA;                                                             \
      }                                                                                    \
   } while (0)

// Exchange the halo of `nf` fields (cells: vertices = false; vertex arrays: true) on stream s.
// Directions without a neighbouring rank wrap locally (periodic, one rank in that direction)
// or are left alone (non-periodic domain edge: static ghost data).
static int exchangeHalo(kgpu_handle *h, double *const *planes, int nf, bool vertices, cudaStream_t s) {
   kgpu_comm &c = h->comm;
   const DevParams &D = h->D;
   StripArgs a;
   a.nf = nf;
   for (int d = 0; d < nf; d++) a.f[d] = planes[d];
   HaloArgs ha;
   ha.nf = nf;
   for (int d = 0; d < nf; d++) ha.f[d] = planes[d];
   int ex = vertices ? 1 : 0;
   int NX = h->NX, NY = h->NY;
   ncclComm_t comm = (ncclComm_t)c.nccl;
   // ---- x phase: two columns each side, owned rows only
   if (c.px == 1) {
      if (h->globalPeriodic) {
         halo_periodic_x_kernel<<<(NY + ex + 127) / 128, 128, 0, s>>>(D, ha, ex);
         h->launches++;
      }
   } else {
      int nr = h->oneD ? 1 : NY + ex;
      size_t cnt = (size_t)nf * nr * 2;
      int grid = (nr * 2 + 127) / 128;
      int sendW = vertices ? 1 : 0;          // cols sent west: cells [0,2), vertices [1,3)
      int recvE = vertices ? NX + 1 : NX;    // where the east neighbour's west strip lands
      if (c.west >= 0) strip_cols_kernel<<<grid, 128, 0, s>>>(D, a, c.sendBuf[0], sendW, 0, nr, 0);
      if (c.east >= 0) strip_cols_kernel<<<grid, 128, 0, s>>>(D, a, c.sendBuf[1], NX - 2, 0, nr, 0);
      NCCL_TRY(h, g_nccl.GroupStart());
      if (c.west >= 0) NCCL_TRY(h, g_nccl.Send(c.sendBuf[0], cnt, ncclDouble, c.west, comm, s));
      if (c.east >= 0) NCCL_TRY(h, g_nccl.Send(c.sendBuf[1], cnt, ncclDouble, c.east, comm, s));
      if (c.east >= 0) NCCL_TRY(h, g_nccl.Recv(c.recvBuf[1], cnt, ncclDouble, c.east, comm, s));
      if (c.west >= 0) NCCL_TRY(h, g_nccl.Recv(c.recvBuf[0], cnt, ncclDouble, c.west, comm, s));
      NCCL_TRY(h, g_nccl.GroupEnd());
      if (c.east >= 0) strip_cols_kernel<<<grid, 128, 0, s>>>(D, a, c.recvBuf[1], recvE, 0, nr, 1);
      if (c.west >= 0) strip_cols_kernel<<<grid, 128, 0, s>>>(D, a, c.recvBuf[0], -2, 0, nr, 1);
      h->launches += 4;
   }
   if (h->oneD) return 0;
   // ---- y phase: two rows each side over the x-haloed width (corners ride along)
   if (c.py == 1) {
      if (h->globalPeriodic) {
         halo_periodic_y_kernel<<<(NX + 4 + ex + 127) / 128, 128, 0, s>>>(D, ha, ex);
         h->launches++;
      }
   } else {
      int nc = NX + 4 + ex;
      size_t cnt = (size_t)nf * nc * 2;
      int grid = (nc * 2 + 127) / 128;
      int sendS = vertices ? 1 : 0;
      int recvN = vertices ? NY + 1 : NY;
      if (c.south >= 0) strip_rows_kernel<<<grid, 128, 0, s>>>(D, a, c.sendBuf[2], sendS, -2, nc, 0);
      if (c.north >= 0) strip_rows_kernel<<<grid, 128, 0, s>>>(D, a, c.sendBuf[3], NY - 2, -2, nc, 0);
      NCCL_TRY(h, g_nccl.GroupStart());
      if (c.south >= 0) NCCL_TRY(h, g_nccl.Send(c.sendBuf[2], cnt, ncclDouble, c.south, comm, s));
      if (c.north >= 0) NCCL_TRY(h, g_nccl.Send(c.sendBuf[3], cnt, ncclDouble, c.north, comm, s));
      if (c.north >= 0) NCCL_TRY(h, g_nccl.Recv(c.recvBuf[3], cnt, ncclDouble, c.north, comm, s));
      if (c.south >= 0) NCCL_TRY(h, g_nccl.Recv(c.recvBuf[2], cnt, ncclDouble, c.south, comm, s));
      NCCL_TRY(h, g_nccl.GroupEnd());
      if (c.north >= 0) strip_rows_kernel<<<grid, 128, 0, s>>>(D, a, c.recvBuf[3], recvN, -2, nc, 1);
      if (c.south >= 0) strip_rows_kernel<<<grid, 128, 0, s>>>(D, a, c.recvBuf[2], -2, -2, nc, 1);
      h->launches += 4;
   }
   CUDA_TRY(h, cudaGetLastError());
   return 0;
}

// global minimum of the unit-CFL step of slot k (TimeStepper.f90:155-169): exact and order-free,
// so every rank takes bit-identical dt decisions.
static int allreduceCfl(kgpu_handle *h, int slot) {
   if (!h->comm.active) return 0;
   double *p = reinterpret_cast<double *>(&h->d_ctrl->cflBits[slot]);
   NCCL_TRY(h, g_nccl.AllReduce(p, p, 1, ncclDouble, ncclMin, (ncclComm_t)h->comm.nccl, h->stream));
   return 0;
}

// A state that went non-finite on one rank must stop every rank (a rank returning alone would leave the others
// waiting in the next exchange): the flag is max-reduced before the host reads the control block.
static int allreduceNonfinite(kgpu_handle *h) {
   if (!h->comm.active) return 0;
   NCCL_TRY(h, g_nccl.AllReduce(&h->d_ctrl->nonfinite, &h->d_ctrl->nonfinite, 1, ncclInt, ncclMax, (ncclComm_t)h->comm.nccl, h->stream));
   return 0;
}

// the morphodynamic refine flag and the length of the redistribution list (adjacent ints of the control
// block): every rank must take the same decision (TimeStepper.f90:709-773), so reduce with max into the
// g* pair -- the local list length is still needed afterwards
static int allreduceMorphoFlags(kgpu_handle *h) {
   if (!h->comm.active) return 0;
   static_assert(offsetof(Ctrl, nRedist) == offsetof(Ctrl, refineMorpho) + sizeof(int), "flags must be adjacent");
   static_assert(offsetof(Ctrl, gRedistMax) == offsetof(Ctrl, gRefine) + sizeof(int), "flags must be adjacent");
   NCCL_TRY(h, g_nccl.AllReduce(&h->d_ctrl->refineMorpho, &h->d_ctrl->gRefine, 2, ncclInt, ncclMax, (ncclComm_t)h->comm.nccl, h->stream));
   return 0;
}
