// kgpu_dyn_host.inl -- dynamic tiles in a decomposed run (non-periodic domains); included by kestrel_gpu.cu.
//
// UpdateTiles.f90 keeps ONE ordered list of active tiles and one of ghost tiles for the whole domain, and
// CheckIfNearBoundaries (TimeStepper.f90:924-1150) walks the first while it grows (quirk Q3).  A decomposed run
// therefore keeps the table replicated (kgpu_tile_table.hpp): every rank holds it for the WHOLE tile grid, the four
// flag bits of the active tiles are combined with one ncclAllReduce(max) per step (each tile's bits come from its
// owner, the others contribute 0), and every rank replays the same mutation sequence.  The device work the replay asks
// for -- heights, ghost data, activation -- is executed, in the table's order, by every rank whose STORAGE the tile
// touches: the owner, and the neighbours that hold two of its cell columns / rows (three of its vertex columns / rows)
// in their halo, each kernel clipped to the rank's own storage (cellInStorage / vertexInStorage, kgpu_tiles.cuh).
// After a replay a rank's planes are therefore the exact restriction of the single-device planes to its block plus
// halo -- including the shared seams, whose value depends on the order in which the tiles on both sides got their
// heights (EqualiseTopographicBoundaryData) -- without any exchange; the time step then keeps the halo current
// exactly as in the periodic all-active case (exchangeHalo: domain edges have no neighbour and keep their static
// ghost data).  kgpu_upload_tile is collective in this mode: every rank passes every tile, in the same order.
// The local tstate / activeList / ghostList / loaded arrays the rest of the library reads are projections of the
// table onto the rank's block.

static bool dynTouches(const kgpu_handle *h, int g0, int &ltx, int &lty) {
   ltx = g0 % h->gnXt - h->gtx0;
   lty = g0 / h->gnXt - h->gty0;
   if (ltx < -1 || ltx > h->nXt) return false;
   return h->oneD ? lty == 0 : (lty >= -1 && lty <= h->nYt);
}
static bool dynOwns(const kgpu_handle *h, int ltx, int lty) { return ltx >= 0 && ltx < h->nXt && lty >= 0 && lty < h->nYt; }

// local projections of the replicated table; every rank raises masksDirty together (refreshMasks is followed by an
// exchange all ranks must take part in)
static void dynProject(kgpu_handle *h) {
   const kgpu::TileTable &T = h->gt;
   for (int t0 = 0; t0 < h->nTiles; t0++) {
      int g0 = globalTileId(h, t0) - 1;
      h->tstate[t0] = T.tstate[g0];
      h->loaded[t0] = T.loaded[g0];
      h->hasSource[t0] = (char)h->gSource[g0];
   }
   h->activeList.clear();
   for (int id : T.activeList) {   // ascending global ids; the local ids of a block ascend with them
      int t0 = localTile0(h, id);
      if (t0 >= 0) h->activeList.push_back(t0 + 1);
   }
   h->ghostList.clear();
   for (int id : T.ghostList) {
      int t0 = localTile0(h, id);
      if (t0 >= 0) h->ghostList.push_back(t0 + 1);
   }
   h->ntilesAdded = T.ntilesAdded;
   h->masksDirty = true;
}

static int defaultTileXY(kgpu_handle *h, int tx, int ty, int kind);

// the device side of loadHeights for tile (ltx, lty) in local tile coordinates (possibly in the ring)
static int dynLoadHeights(kgpu_handle *h, int g0, int ltx, int lty, int mask, const double *given) {
   int nX = h->nX, nY = h->nY;
   size_t nv = (size_t)(nX + 1) * (nY + 1);
   double *hb = h->h_stage;
   const bool fromRaster = !given && h->raster.elev != nullptr;
   const bool onDevice = !given && (fromRaster || h->topogFn.func >= 0);
   if (given) {
      std::memcpy(hb, given, sizeof(double) * (size_t)(nX + 1) * (h->oneD ? 1 : nY + 1));
   } else if (!onDevice) {
      if (!h->P.heights) { h->err = "no heights callback registered and no b0_vertices given"; return KGPU_ERR_ARG; }
      if (h->P.heights(h->P.heights_ctx, g0 + 1, hb) != 0) { h->err = "heights callback failed"; return KGPU_ERR_ARG; }
   }
   dim3 grid((nX + 1 + 127) / 128, h->oneD ? 1 : nY + 1);
   const int gi = g0 % h->gnXt + 1, gj = g0 / h->gnXt + 1;
   if (onDevice) {
      if (fromRaster) tile_raster_kernel<<<grid, 128, 0, h->stream>>>(h->D, h->raster, gi, gj, h->d_stage);
      else tile_topog_kernel<<<grid, 128, 0, h->stream>>>(h->D, h->topogFn, gi, gj, h->d_stage);
      h->launches++;
   } else CUDA_TRY(h, cudaMemcpyAsync(h->d_stage, hb, nv * sizeof(double), cudaMemcpyHostToDevice, h->stream));
   tile_vertices_kernel<<<grid, 128, 0, h->stream>>>(h->D, h->b0v, h->d_stage, ltx, lty, 1, mask);
   h->launches++;
   CUDA_TRY(h, cudaStreamSynchronize(h->stream));  // h_stage is reused
   h->topoBtIdx = -1;
   return 0;
}

// Execute the device work the table has recorded since the last call.  `given` = heights handed to kgpu_upload_tile
// for tile givenTile0 (the one TOP_LOAD_HEIGHTS of this batch that does not go through the callback).
static int dynExecOps(kgpu_handle *h, int givenTile0, const double *given) {
   kgpu::TileTable &T = h->gt;
   const bool any = h->opsDone < T.ops.size();
   for (; h->opsDone < T.ops.size(); h->opsDone++) {
      const kgpu::TileOp op = T.ops[h->opsDone];
      int ltx, lty, rc = 0;
      if (!dynTouches(h, op.tile0, ltx, lty)) continue;
      switch (op.kind) {
         case kgpu::TOP_LOAD_HEIGHTS:
            rc = dynLoadHeights(h, op.tile0, ltx, lty, op.mask, op.tile0 == givenTile0 ? given : nullptr);
            break;
         case kgpu::TOP_GHOST_DATA:
            rc = defaultTileXY(h, ltx, lty, (h->P.bcs == KGPU_BC_DIRICHLET && T.onDomainEdge(op.tile0)) ? 1 : 0);
            break;
         case kgpu::TOP_ACTIVATE_FRESH: rc = defaultTileXY(h, ltx, lty, 0); break;
         default: rc = defaultTileXY(h, ltx, lty, 2); break;   // promoted ghost: w = b0
      }
      if (rc) return rc;
   }
   if (any) dynProject(h);
   return 0;
}

static int dynHaltCheck(kgpu_handle *h, bool ok) {
   if (ok) return 0;
   if (h->gt.haltViolation) {
      h->err = "tried to add a tile outside the domain (Boundary Conditions = halt)";
      return KGPU_ERR_HALT_BC;
   }
   h->err = "ghost tile out of bounds (UpdateTiles.f90:423)";
   return KGPU_ERR_ARG;
}

// CheckIfNearBoundaries across ranks: own flags -> allreduce -> replay -> device work
static int dynCheckIfNearBoundaries(kgpu_handle *h) {
   kgpu::TileTable &T = h->gt;
   if (T.activeList.empty()) return 0;
   const int nG = T.nTiles();
   std::vector<int> flags(nG, 0);
   if (h->firstScan) {
      for (int id : T.activeList) flags[id - 1] = h->gSeed[id - 1];
   } else {
      int nAct = (int)h->activeList.size();
      std::vector<int> tl(std::max(nAct, 1));
      for (int k = 0; k < nAct; k++) tl[k] = h->activeList[k] - 1;
      if (nAct > 0) {
         CUDA_TRY(h, cudaMemcpyAsync(h->d_tileList, tl.data(), nAct * sizeof(int), cudaMemcpyHostToDevice, h->stream));
         tile_flags_kernel<<<nAct, 128, 0, h->stream>>>(h->D, h->sp(h->i0), h->b0v, h->morpho ? h->btv[h->bt0] : nullptr,
                                                        h->d_tileList, h->P.TileBuffer, h->d_flags);
         h->launches++;
         CUDA_TRY(h, cudaMemcpyAsync(h->h_flags, h->d_flags, nAct * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
         CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      }
      for (int k = 0; k < nAct; k++) h->h_gflags[globalTileId(h, tl[k]) - 1] = h->h_flags[k];
      // every tile's bits come from its owner; the other ranks hold 0 there
      CUDA_TRY(h, cudaMemcpyAsync(h->d_gflags, h->h_gflags, nG * sizeof(int), cudaMemcpyHostToDevice, h->stream));
      NCCL_TRY(h, g_nccl.AllReduce(h->d_gflags, h->d_gflags, nG, ncclInt, ncclMax, (ncclComm_t)h->comm.nccl, h->stream));
      CUDA_TRY(h, cudaMemcpyAsync(h->h_gflags, h->d_gflags, nG * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      for (int g = 0; g < nG; g++) { flags[g] = h->h_gflags[g]; h->h_gflags[g] = 0; }
   }
   h->firstScan = false;
   bool ok = T.replay(flags.data(), h->nX, h->nY, h->P.TileBuffer);
   int rc = dynHaltCheck(h, ok);
   if (rc) return rc;
   return dynExecOps(h, -1, nullptr);
}

// kgpu_upload_tile, collective form: tile_id is any tile of the whole grid; every rank applies the table mutation and
// the part of the device work that lands in its storage
static int dynUploadTile(kgpu_handle *h, int32_t tile_id, const double *u13, const double *b0_vertices, const double *bt_vertices,
                         const double *maxima, const double *tfirst, int32_t contains_source) {
   kgpu::TileTable &T = h->gt;
   const int g0 = tile_id - 1;
   if (g0 < 0 || g0 >= T.nTiles()) { h->err = "tile id out of range"; return KGPU_ERR_ARG; }
   int rc;
   if (b0_vertices) {
      T.loadHeights(g0, true);
      if ((rc = dynExecOps(h, g0, b0_vertices))) return rc;
   }
   bool ok = T.addTile(g0, false, false);
   if ((rc = dynHaltCheck(h, ok))) return rc;
   if ((rc = dynExecOps(h, -1, nullptr))) return rc;
   if (T.tstate[g0] != 2) { h->err = "tile lies on the domain edge and cannot be active"; return KGPU_ERR_ARG; }
   h->gSource[g0] = contains_source ? 1 : 0;
   int nX = h->nX, nY = h->nY;
   size_t ncell = (size_t)nX * nY;
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
   h->gSeed[g0] = f;
   h->havePre = false;
   dynProject(h);
   int ltx, lty;
   if (!dynTouches(h, g0, ltx, lty)) return KGPU_OK;
   if (bt_vertices && h->morpho) {
      size_t nv = (size_t)(nX + 1) * (h->oneD ? 1 : nY + 1);
      std::memcpy(h->h_stage, bt_vertices, nv * sizeof(double));
      CUDA_TRY(h, cudaMemcpyAsync(h->d_stage, h->h_stage, nv * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      dim3 gridv((nX + 1 + 127) / 128, h->oneD ? 1 : nY + 1);
      tile_vertices_kernel<<<gridv, 128, 0, h->stream>>>(h->D, h->btv[h->bt0], h->d_stage, ltx, lty, 1, 7);
      h->launches++;
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
   }
   std::memcpy(h->h_stage, u13, ncell * 13 * sizeof(double));
   if (maxima) std::memcpy(h->h_stage + ncell * 13, maxima, ncell * 10 * sizeof(double));
   if (tfirst) std::memcpy(h->h_stage + ncell * 23, tfirst, ncell * sizeof(double));
   CUDA_TRY(h, cudaMemcpyAsync(h->d_stage, h->h_stage, ncell * 24 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
   dim3 grid((nX + 127) / 128, nY);
   import_tile_kernel<<<grid, 128, 0, h->stream>>>(h->D, h->sp(h->i0), h->mp(), h->d_stage, ltx, lty, maxima ? 1 : 0, tfirst ? 1 : 0);
   h->launches++;
   CUDA_TRY(h, cudaGetLastError());
   CUDA_TRY(h, cudaStreamSynchronize(h->stream));
   return KGPU_OK;
}
