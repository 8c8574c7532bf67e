// kgpu_morpho_host.inl -- Strang remainder M(2 dt) . H(dt) of one step
// (TimeStepper.f90:194-255, 532-781); included by kestrel_gpu.cu after hydraulicTimeStepper.
//
// Buffers on entry (first H done): S[ib] = result of H1, S[ia] = its momenta before the
// implicit correction.  M stages alternate between S[ic] and S[id] (w and Hnpsi planes only;
// the momenta do not change over M).  The second H runs q0' = S[ic] -> result S[id].

__global__ void ctrl_morpho_reset_kernel(Ctrl *c) {
   c->refineMorpho = 0;
   c->nRedist = 0;
}

static int fillHaloPlanes(kgpu_handle *h, double *const *planes, int n, bool vertices) {
   if (h->comm.active) return exchangeHalo(h, planes, n, vertices, h->stream);   // n <= 8 (StripArgs)
   if (!h->periodic) return 0;
   HaloArgs a;
   a.nf = n;
   for (int d = 0; d < n; d++) a.f[d] = planes[d];
   int ex = vertices ? 1 : 0;
   if (!h->oneD) {
      halo_periodic_x_kernel<<<(h->NY + ex + 127) / 128, 128, 0, h->stream>>>(h->D, a, ex);
      halo_periodic_y_kernel<<<(h->NX + 4 + ex + 127) / 128, 128, 0, h->stream>>>(h->D, a, ex);
      h->launches += 2;
   } else {
      halo_periodic_x_kernel<<<1, 32, 0, h->stream>>>(h->D, a, ex);
      h->launches += 1;
   }
   return 0;
}

// RedistributeGrid across ranks (see kgpu_morpho.cuh): gather the lists and the patches, replay the global
// walk identically on every rank, scatter the canonical values into the local planes.  The host only
// sorts indices and builds slot tables -- the same bookkeeping role it has on a single device.
static int redistributeAcrossRanks(kgpu_handle *h, int nLocal, int M, int R1, int MA) {
   kgpu_comm &c = h->comm;
   const bool gather = c.active;   // false: single device driven through the same walk (KGPU_REDIST_GLOBAL=1, test hook)
   const int R = gather ? c.size : 1;
   ncclComm_t comm = (ncclComm_t)c.nccl;
   const size_t tot = (size_t)R * M;
   RedistGlobalBufs &B = h->rg;
   if (B.cap < tot) {
      cudaFree(B.dEntries); cudaFree(B.dPatch); cudaFree(B.dVslot); cudaFree(B.dCslot); cudaFree(B.dVkey); cudaFree(B.dVbase);
      cudaFree(B.dCkey); cudaFree(B.dCbase);
      B.cap = tot + tot / 2 + 64;
      CUDA_TRY(h, cudaMalloc(&B.dEntries, sizeof(RedistEntry) * B.cap));
      CUDA_TRY(h, cudaMalloc(&B.dPatch, sizeof(double) * RP_DOUBLES * B.cap));
      CUDA_TRY(h, cudaMalloc(&B.dVslot, sizeof(int) * RP_V * B.cap));
      CUDA_TRY(h, cudaMalloc(&B.dCslot, sizeof(int) * RP_C * B.cap));
      CUDA_TRY(h, cudaMalloc(&B.dVkey, sizeof(int) * 2 * RP_V * B.cap));
      CUDA_TRY(h, cudaMalloc(&B.dVbase, sizeof(int) * RP_V * B.cap));
      CUDA_TRY(h, cudaMalloc(&B.dCkey, sizeof(int) * 2 * RP_C * B.cap));
      CUDA_TRY(h, cudaMalloc(&B.dCbase, sizeof(int) * RP_C * B.cap));
   }
   if (B.sendCap < M) {
      cudaFree(B.dSend);
      B.sendCap = M + M / 2 + 64;
      CUDA_TRY(h, cudaMalloc(&B.dSend, sizeof(double) * RP_DOUBLES * B.sendCap));
   }
   if (!B.dCounts) CUDA_TRY(h, cudaMalloc(&B.dCounts, sizeof(int) * 64));
   if (R > 64) { h->err = "redistribution across more than 64 ranks"; return KGPU_ERR_UNSUPPORTED; }
   // 1. list lengths, lists (M entries per rank, the tail beyond a rank's own length is padding) and patches
   if (gather) {
      NCCL_TRY(h, g_nccl.AllGather(&h->d_ctrl->nRedist, B.dCounts, 1, ncclInt, comm, h->stream));
      NCCL_TRY(h, g_nccl.AllGather(h->d_redist, B.dEntries, sizeof(RedistEntry) * (size_t)M, ncclChar, comm, h->stream));
   } else {
      CUDA_TRY(h, cudaMemcpyAsync(B.dCounts, &h->d_ctrl->nRedist, sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
      CUDA_TRY(h, cudaMemcpyAsync(B.dEntries, h->d_redist, sizeof(RedistEntry) * (size_t)M, cudaMemcpyDeviceToDevice, h->stream));
   }
   if (nLocal > 0) {
      RedistPackArgs pa;
      pa.b0v = h->b0v; pa.bt0 = h->btv[h->bt0]; pa.bt3 = h->btv[h->bt3];
      pa.w0 = h->S[R1][QW]; pa.hpsi0 = h->S[R1][QHPSI]; pa.w3 = h->S[MA][QW]; pa.hpsi3 = h->S[MA][QHPSI];
      pa.list = h->d_redist; pa.n = nLocal; pa.tileMask = h->d_tileMask; pa.allActive = h->allActive() ? 1 : 0;
      int nthr = nLocal * RP_DOUBLES;
      redist_pack_kernel<<<(nthr + 255) / 256, 256, 0, h->stream>>>(h->D, pa, B.dSend);
      h->launches++;
   }
   if (gather) NCCL_TRY(h, g_nccl.AllGather(B.dSend, B.dPatch, (size_t)RP_DOUBLES * M, ncclDouble, comm, h->stream));
   else CUDA_TRY(h, cudaMemcpyAsync(B.dPatch, B.dSend, sizeof(double) * RP_DOUBLES * M, cudaMemcpyDeviceToDevice, h->stream));
   std::vector<RedistEntry> all(tot);
   std::vector<int> counts(R);
   CUDA_TRY(h, cudaMemcpyAsync(all.data(), B.dEntries, sizeof(RedistEntry) * tot, cudaMemcpyDeviceToHost, h->stream));
   CUDA_TRY(h, cudaMemcpyAsync(counts.data(), B.dCounts, sizeof(int) * R, cudaMemcpyDeviceToHost, h->stream));
   CUDA_TRY(h, cudaStreamSynchronize(h->stream));
   // 2. + 3. the global walk order and the slot tables (kgpu_redist_tables.hpp; identical on every rank)
   std::vector<double> ex(tot);
   std::vector<int> li(tot), lj(tot);
   for (size_t k = 0; k < tot; k++) { ex[k] = all[k].excess; li[k] = all[k].i; lj[k] = all[k].j; }
   RedistGeometry geo{R, M, gather ? c.px : 1, h->NX, h->NY, h->nX, h->nY, h->gnXt, h->gnYt, h->oneD ? 1 : 0};
   RedistTables T;
   buildRedistTables(geo, counts.data(), ex.data(), li.data(), lj.data(), T);
   const int n = (int)T.patch.size();
   const int NXg = h->nX * h->gnXt, NYg = h->oneD ? 1 : h->nY * h->gnYt;
   const std::vector<int> &vslot = T.vslot, &cslot = T.cslot, &vkey = T.vkey, &vbase = T.vbase, &ckey = T.ckey, &cbase = T.cbase;
   CUDA_TRY(h, cudaMemcpyAsync(B.dVslot, vslot.data(), sizeof(int) * vslot.size(), cudaMemcpyHostToDevice, h->stream));
   CUDA_TRY(h, cudaMemcpyAsync(B.dCslot, cslot.data(), sizeof(int) * cslot.size(), cudaMemcpyHostToDevice, h->stream));
   CUDA_TRY(h, cudaMemcpyAsync(B.dVkey, vkey.data(), sizeof(int) * vkey.size(), cudaMemcpyHostToDevice, h->stream));
   CUDA_TRY(h, cudaMemcpyAsync(B.dVbase, vbase.data(), sizeof(int) * vbase.size(), cudaMemcpyHostToDevice, h->stream));
   CUDA_TRY(h, cudaMemcpyAsync(B.dCkey, ckey.data(), sizeof(int) * ckey.size(), cudaMemcpyHostToDevice, h->stream));
   CUDA_TRY(h, cudaMemcpyAsync(B.dCbase, cbase.data(), sizeof(int) * cbase.size(), cudaMemcpyHostToDevice, h->stream));
   // 4. the walk (one thread, as on a single device) and the scatter
   RedistGlobalArgs ga;
   ga.G = B.dPatch; ga.vslot = B.dVslot; ga.cslot = B.dCslot; ga.n = n; ga.ctrl = h->d_ctrl;
   redist_global_kernel<<<1, 32, 0, h->stream>>>(h->D, ga);
   RedistScatterArgs sa;
   sa.G = B.dPatch; sa.vkey = B.dVkey; sa.vbase = B.dVbase; sa.ckey = B.dCkey; sa.cbase = B.dCbase;
   sa.nv = (int)vbase.size(); sa.nc = (int)cbase.size();
   sa.bt3 = h->btv[h->bt3]; sa.w3 = h->S[MA][QW]; sa.hpsi3 = h->S[MA][QHPSI];
   sa.gx0 = gather ? c.rx * h->NX : 0; sa.gy0 = gather ? c.ry * h->NY : 0; sa.NXg = NXg; sa.NYg = NYg;
   redist_scatter_kernel<<<(sa.nv + sa.nc + 255) / 256, 256, 0, h->stream>>>(h->D, sa);
   h->launches += 2;
   CUDA_TRY(h, cudaGetLastError());
   CUDA_TRY(h, cudaStreamSynchronize(h->stream));   // the host tables go out of scope
   return 0;
}

template <bool ONED>
static int morphoStageT(kgpu_handle *h, const MorphoArgs &a) {
   constexpr int BX = ONED ? BX1 : BX2, BY = ONED ? BY1 : BY2;
   if (!h->comm.active && h->morphoFusion > 0) {
      // measured alternatives (slower than the three kernels below on B200): the stage bed and the cell
      // update fused (morpho_stage_kernel), E - D from its plane (level 1) or evaluated in the same launch (level 2); then the
      // periodic images of the new bed and of the six centre planes the next stage reads
      if (h->morphoFusion == 2) {
         if (ONED) morpho_stage_kernel<BX1, BY1, true, false><<<h->nBlocks, NTHREADS, 0, h->stream>>>(h->D, a, h->d_blockList);
         else morpho_stage_kernel<BX2, MORPHO_STAGE_BY, false, false><<<h->nBlocksM, NTHREADS, 0, h->stream>>>(h->D, a, h->d_blockListM);
         h->launches++;
      } else {   // E - D once per cell into its plane, then the bed and the cell update fused
         morpho_emd_kernel<BX, BY><<<h->nBlocks, NTHREADS, 0, h->stream>>>(h->D, a);
         if (ONED) morpho_stage_kernel<BX1, BY1, true, true><<<h->nBlocks, NTHREADS, 0, h->stream>>>(h->D, a, h->d_blockList);
         else morpho_stage_kernel<BX2, MORPHO_STAGE_BY, false, true><<<h->nBlocksM, NTHREADS, 0, h->stream>>>(h->D, a, h->d_blockListM);
         h->launches += 2;
      }
      double *pl[1] = {a.btn};
      int rc = fillHaloPlanes(h, pl, 1, true);
      if (rc) return rc;
      double *pc[6] = {a.wn, a.hpsin, a.nBt, a.nBx, a.nBy, a.nHn};
      rc = fillHaloPlanes(h, pc, 6, false);
      CUDA_TRY(h, cudaGetLastError());
      return rc;
   }
   // (a rank of a decomposed run may hold no active block: it still takes part in every exchange)
   if (h->nBlocks > 0) morpho_emd_kernel<BX, BY><<<h->nBlocks, NTHREADS, 0, h->stream>>>(h->D, a);
   if (h->comm.active) {   // the bed kernel reads E - D of the cells across the block edge (single device: by wrapped index)
      double *pe[1] = {a.EmD};
      int rce = exchangeHalo(h, pe, 1, false, h->stream);
      if (rce) return rce;
   }
   dim3 gv((h->NX + 1 + 127) / 128, h->oneD ? 1 : h->NY + 1);
   morpho_bed_kernel<<<gv, 128, 0, h->stream>>>(h->D, a);
   double *pl[1] = {a.btn};
   int rc = fillHaloPlanes(h, pl, 1, true);
   if (rc) return rc;
   if (h->nBlocks > 0) morpho_cell_kernel<BX, BY><<<h->nBlocks, NTHREADS, 0, h->stream>>>(h->D, a);
   double *pc[5] = {a.wn, a.hpsin, a.nBx, a.nBy, a.nHn};   // the next stage reads slopes and depths of halo cells
   rc = fillHaloPlanes(h, pc, 5, false);
   h->launches += 3;
   CUDA_TRY(h, cudaGetLastError());
   return rc;
}

static int strangRemainder(kgpu_handle *h, double t0, double &dt_hydro, bool &again) {
   again = false;
   int rc = 0;
   const int R1 = h->ib, PRE = h->ia, MA = h->ic, MB = h->id;
   const int allAct = h->allActive() ? 1 : 0;
   int BX = h->oneD ? BX1 : BX2, BY = h->oneD ? BY1 : BY2;
   if (h->nBlocks == 0 && !h->comm.active) return 0;
   const bool some = h->nBlocks > 0;   // false: a rank without active blocks, which only takes part in the collectives
   // (the halo of the H1 result was produced together with it)
   // velocities frozen over M = those of H1's 4th RHS evaluation (pre-correction momenta)
   // the hydraulic topography planes are those of bt0 here (H1 just ran with them)
   if (h->topoBtIdx != h->bt0 && (rc = computeTopo(h, h->bt0))) return rc;
   if (!some) {}
   else if (h->oneD)
      morpho_prepare_kernel<BX1, BY1><<<h->nBlocks, NTHREADS, 0, h->stream>>>(h->D, h->S[R1][QW], h->S[R1][QHPSI], h->S[PRE][QHU], h->S[PRE][QHV],
                                                                            h->topo.b0c, h->topo.btc, h->topo.bxc, h->topo.byc, h->Um, h->Vm,
                                                                            h->mcHn0, h->d_tileMask, h->d_blockList, allAct);
   else
      morpho_prepare_kernel<BX2, BY2><<<h->nBlocks, NTHREADS, 0, h->stream>>>(h->D, h->S[R1][QW], h->S[R1][QHPSI], h->S[PRE][QHU], h->S[PRE][QHV],
                                                                            h->topo.b0c, h->topo.btc, h->topo.bxc, h->topo.byc, h->Um, h->Vm,
                                                                            h->mcHn0, h->d_tileMask, h->d_blockList, allAct);
   // (the fused stage kernel evaluates E - D in the cells around the tile: the frozen velocities need their images too)
   { double *ph[3] = {h->mcHn0, h->Um, h->Vm}; if ((rc = fillHaloPlanes(h, ph, 3, false))) return rc; }
   h->launches++;
   double dt_morpho = 2.0 * dt_hydro;
   MorphoArgs a;
   a.w0 = h->S[R1][QW]; a.hpsi0 = h->S[R1][QHPSI]; a.U = h->Um; a.V = h->Vm; a.b0v = h->b0v; a.bt0 = h->btv[h->bt0];
   a.EmD = h->EmD; a.tileMask = h->d_tileMask; a.blockList = h->d_blockList; a.ctrl = h->d_ctrl; a.allActive = allAct;
   a.dtMorpho = dt_morpho;
   a.b0c = h->topo.b0c; a.zBt = h->topo.btc; a.zBx = h->topo.bxc; a.zBy = h->topo.byc; a.zGam = h->topo.gamc;
   auto centreIn = [&](const double *bt, const double *bx, const double *by, const double *hn) { a.cBt = bt; a.cBx = bx; a.cBy = by; a.cHn = hn; };
   auto centreOut = [&](int set) { a.nBt = h->mc[set][0]; a.nBx = h->mc[set][1]; a.nBy = h->mc[set][2]; a.nHn = h->mc[set][3]; };
   // stage 1 (TimeStepper.f90:566-610)
   centreIn(h->topo.btc, h->topo.bxc, h->topo.byc, h->mcHn0); centreOut(0);
   a.w = h->S[R1][QW]; a.hpsi = h->S[R1][QHPSI]; a.btk = h->btv[h->bt0]; a.btn = h->btv[h->bt1];
   a.wn = h->S[MA][QW]; a.hpsin = h->S[MA][QHPSI]; a.a0 = 0.0; a.a1 = 1.0;
   if ((rc = h->oneD ? morphoStageT<true>(h, a) : morphoStageT<false>(h, a))) return rc;
   // stage 2 (:612-655)
   centreIn(h->mc[0][0], h->mc[0][1], h->mc[0][2], h->mc[0][3]); centreOut(1);
   a.w = h->S[MA][QW]; a.hpsi = h->S[MA][QHPSI]; a.btk = h->btv[h->bt1]; a.btn = h->btv[h->bt2];
   a.wn = h->S[MB][QW]; a.hpsin = h->S[MB][QHPSI]; a.a0 = 0.75; a.a1 = 0.25;
   if ((rc = h->oneD ? morphoStageT<true>(h, a) : morphoStageT<false>(h, a))) return rc;
   // stage 3 (:657-699)
   centreIn(h->mc[1][0], h->mc[1][1], h->mc[1][2], h->mc[1][3]); centreOut(0);
   a.w = h->S[MB][QW]; a.hpsi = h->S[MB][QHPSI]; a.btk = h->btv[h->bt2]; a.btn = h->btv[h->bt3];
   a.wn = h->S[MA][QW]; a.hpsin = h->S[MA][QHPSI]; a.a0 = 1.0 / 3.0; a.a1 = 2.0 / 3.0;
   if ((rc = h->oneD ? morphoStageT<true>(h, a) : morphoStageT<false>(h, a))) return rc;
   // checks (:709-753)
   ctrl_morpho_reset_kernel<<<1, 1, 0, h->stream>>>(h->d_ctrl);
   CheckArgs c;
   c.w0 = h->S[R1][QW]; c.hpsi0 = h->S[R1][QHPSI]; c.w3 = h->S[MA][QW]; c.b0v = h->b0v; c.bt0 = h->btv[h->bt0]; c.bt3 = h->btv[h->bt3];
   c.b0c = h->topo.b0c; c.zBt = h->topo.btc; c.zGam = h->topo.gamc; c.c3Bt = h->mc[0][0]; c.c3Bx = h->mc[0][1]; c.c3By = h->mc[0][2];
   c.tileMask = h->d_tileMask; c.blockList = h->d_blockList; c.ctrl = h->d_ctrl; c.list = h->d_redist; c.listCap = h->redistCap; c.allActive = allAct;
   if (!some) {}
   else if (h->oneD) morpho_check_kernel<BX1, BY1><<<h->nBlocks, NTHREADS, 0, h->stream>>>(h->D, c);
   else morpho_check_kernel<BX2, BY2><<<h->nBlocks, NTHREADS, 0, h->stream>>>(h->D, c);
   h->launches += 2;
   if ((rc = allreduceMorphoFlags(h))) return rc;
   if ((rc = readCtrl(h))) return rc;
   // RedistributeGrid's list is unbounded in the reference (a linked list, Redistribute.f90:69-101): when it
   // outgrows the device buffer the buffer is enlarged and the (cheap, idempotent) check pass repeated
   {
      const int need = h->comm.active ? h->h_ctrl->gRedistMax : h->h_ctrl->nRedist;
      const bool ref0 = h->comm.active ? h->h_ctrl->gRefine != 0 : h->h_ctrl->refineMorpho != 0;
      if (!ref0 && need > h->redistCap) {
         cudaFree(h->d_redist); cudaFreeHost(h->h_redist);
         h->d_redist = nullptr; h->h_redist = nullptr;
         h->redistCap = need + need / 2 + 1024;
         CUDA_TRY(h, cudaMalloc(&h->d_redist, sizeof(RedistEntry) * (size_t)h->redistCap));
         CUDA_TRY(h, cudaMallocHost(&h->h_redist, sizeof(RedistEntry) * (size_t)h->redistCap));
         h->nRedistGrows++;
         ctrl_morpho_reset_kernel<<<1, 1, 0, h->stream>>>(h->d_ctrl);
         c.list = h->d_redist; c.listCap = h->redistCap;
         if (!some) {}
         else if (h->oneD) morpho_check_kernel<BX1, BY1><<<h->nBlocks, NTHREADS, 0, h->stream>>>(h->D, c);
         else morpho_check_kernel<BX2, BY2><<<h->nBlocks, NTHREADS, 0, h->stream>>>(h->D, c);
         h->launches += 2;
         if ((rc = allreduceMorphoFlags(h))) return rc;
         if ((rc = readCtrl(h))) return rc;
      }
   }
   bool refine = h->h_ctrl->refineMorpho != 0;
   int nRed = h->h_ctrl->nRedist;
   if (h->comm.active) {   // every rank takes the decisions of the whole domain
      const int nLocal = nRed;
      refine = h->h_ctrl->gRefine != 0;
      nRed = h->h_ctrl->gRedistMax;
      if (!refine && nRed > 0) {
         h->nRedistCells += nLocal;
         if ((rc = redistributeAcrossRanks(h, nLocal, nRed, R1, MA))) return rc;
         if ((rc = readCtrl(h))) return rc;
         refine = h->h_ctrl->refineMorpho != 0;   // the walk is replicated: every rank sets the same flag
      }
      nRed = 0;   // done (or refining): skip the single-device walk below
   } else if (!refine && nRed > 0 && h->periodic && allAct && h->debugGlobalWalk) {
      // test probe (kgpu_debug_global_walk): a single periodic device takes the walk the decomposed runs use
      h->nRedistCells += nRed;
      if ((rc = redistributeAcrossRanks(h, nRed, nRed, R1, MA))) return rc;
      if ((rc = readCtrl(h))) return rc;
      refine = h->h_ctrl->refineMorpho != 0;
      nRed = 0;
   }
   if (!refine && nRed > 0) {
      h->nRedistCells += nRed;
      // sorted ascending by excess, ties in scan order: active-tile order, then j, then i (Redistribute.f90:69-101)
      CUDA_TRY(h, cudaMemcpyAsync(h->h_redist, h->d_redist, sizeof(RedistEntry) * nRed, cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      int nX = h->nX, nY = h->nY, nXt = h->nXt;
      std::stable_sort(h->h_redist, h->h_redist + nRed, [&](const RedistEntry &x, const RedistEntry &y) {
         if (x.excess != y.excess) return x.excess < y.excess;
         int tx_ = (x.i / nX) + (x.j / nY) * nXt, ty_ = (y.i / nX) + (y.j / nY) * nXt;
         if (tx_ != ty_) return tx_ < ty_;
         if (x.j != y.j) return x.j < y.j;
         return x.i < y.i;
      });
      CUDA_TRY(h, cudaMemcpyAsync(h->d_redist, h->h_redist, sizeof(RedistEntry) * nRed, cudaMemcpyHostToDevice, h->stream));
      RedistArgs r;
      r.w0 = h->S[R1][QW]; r.hpsi0 = h->S[R1][QHPSI]; r.w3 = h->S[MA][QW]; r.hpsi3 = h->S[MA][QHPSI];
      r.b0v = h->b0v; r.bt0 = h->btv[h->bt0]; r.bt3 = h->btv[h->bt3]; r.tileMask = h->d_tileMask; r.list = h->d_redist; r.n = nRed;
      r.allActive = allAct; r.ctrl = h->d_ctrl;
      if (h->debugSequentialWalk) {
         morpho_redistribute_kernel<<<1, 32, 0, h->stream>>>(h->D, r);
         h->launches++;
      } else {
         if (!h->d_rankMap) {   // INT_MAX-like everywhere (0x7f7f7f7f); every walk leaves it that way
            CUDA_TRY(h, cudaMalloc(&h->d_rankMap, sizeof(int) * (h->fieldElems + 1)));
            CUDA_TRY(h, cudaMemsetAsync(h->d_rankMap, 0x7f, sizeof(int) * h->fieldElems, h->stream));
         }
         RedistWaveArgs wv;
         wv.R = r; wv.rankMap = h->d_rankMap; wv.ticket = h->d_rankMap + h->fieldElems;
         redist_rank_kernel<<<(nRed + 255) / 256, 256, 0, h->stream>>>(h->D, h->d_redist, nRed, wv.rankMap, wv.ticket);
         morpho_redistribute_wave_kernel<<<(nRed + 127) / 128, 128, 0, h->stream>>>(h->D, wv);
         h->launches += 2;
      }
      if ((rc = readCtrl(h))) return rc;
      refine = h->h_ctrl->refineMorpho != 0;
   }
   if (refine) {
      // newdt = dt_morpho/2, dt_hydro = newdt/2 (TimeStepper.f90:762-773, 203-214)
      dt_morpho = 0.5 * dt_morpho;
      dt_hydro = 0.5 * dt_morpho;
      h->nrefines++;
      again = true;
      return 0;
   }
   // state after M = (w3, rhoHnu, rhoHnv, Hnpsi3): momenta are those of H1
   size_t fb = h->fieldElems * sizeof(double);
   CUDA_TRY(h, cudaMemcpyAsync(h->S[MA][QHU], h->S[R1][QHU], fb, cudaMemcpyDeviceToDevice, h->stream));
   CUDA_TRY(h, cudaMemcpyAsync(h->S[MA][QHV], h->S[R1][QHV], fb, cudaMemcpyDeviceToDevice, h->stream));
   // second hydraulic operator (TimeStepper.f90:217-254); grid%t = t0 + dt_hydro
   double tNow = t0 + dt_hydro;
   if ((rc = fillHaloCells(h, MA))) return rc;  // redistribution may have touched cells after the stage halo
   if ((rc = computeTopo(h, h->bt3))) return rc;  // the bed moved: new cell / face topography for the second H
   if ((rc = firstRHS(h, MA, h->bt3, tNow, 0.0, 0))) return rc;
   h->e0Valid = false;
   if ((rc = readCtrl(h))) return rc;
   double advised = h->h_ctrl->dtAdvised;
   if (advised < dt_hydro) {
      if ((dt_hydro - advised) / dt_hydro < 0.1) dt_hydro = 0.9 * dt_hydro;
      else dt_hydro = advised;
      h->nrefines++;
      again = true;
      return 0;
   }
   ctrl_set_dt_kernel<<<1, 1, 0, h->stream>>>(h->d_ctrl, dt_hydro);
   h->launches++;
   if ((rc = hydraulicTimeStepper(h, MA, h->ia, MB, h->bt3))) return rc;
   if (h->h_ctrl->nonfinite) { h->err = "non-finite state"; return KGPU_ERR_DT; }
   if (h->h_ctrl->failed) {
      dt_hydro = h->h_ctrl->dtNew;
      h->nrefines++;
      again = true;
      return 0;
   }
   return 0;
}
