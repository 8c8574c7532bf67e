/*
 * kestrel_gpu_debug.h -- test probes of libkestrel_gpu.  NOT part of the drop-in boundary: nothing a host
 * program needs is declared here, and the reference has no counterpart for any of it.  The parity tests use
 * these entry points to compare single pieces of the path (one RHS evaluation, the redistribution walk, the
 * host bookkeeping of the decomposed runs) with the oracle.
 */
#ifndef KESTREL_GPU_DEBUG_H
#define KESTREL_GPU_DEBUG_H

#include "kestrel_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Diagnostics: ONE evaluation of CalculateHydraulicRHS (src/HydraulicRHS.f90:64-137) on the
   current state, without advancing it: ddtExplicit E4[(4, NY, NX)], ddtImplicit I[(NY, NX)]
   (either may be NULL) and the advised time step of ComputeAdvisedTimeStep (:141-176) for
   `substep`.  Used by the parity tests to compare a single RHS with the oracle's. */
int kgpu_debug_rhs(kgpu_handle *h, int32_t substep, double *E4, double *I, double *dt);

/* Test probes (no device needed) of the replicated tile table (kestrel_b200/csrc/kgpu_tile_table.hpp), the
 * groundwork for dynamic tile activation across ranks: AddTile / AddGhostTiles (src/UpdateTiles.f90:56-78,
 * 389-481) and the CheckIfNearBoundaries replay (src/TimeStepper.f90:924-1150) on global tile indices.
 * Tile ids are 1-based as in the reference; flags = 4 bits per global tile (N, S, E, W).  add / replay
 * return 0, or KGPU_ERR_HALT_BC when a tile outside a `halt` domain was requested; lists() fills ascending
 * active ids, ghost ids in creation order and the number of device operations requested so far.        */
typedef struct kgpu_tiletable kgpu_tiletable;
kgpu_tiletable *kgpu_debug_tiletable_new(int32_t nXtiles, int32_t nYtiles, int32_t periodic, int32_t isOneD, int32_t halt_bc);
void kgpu_debug_tiletable_free(kgpu_tiletable *t);
int kgpu_debug_tiletable_add(kgpu_tiletable *t, int32_t tile_id);
int kgpu_debug_tiletable_replay(kgpu_tiletable *t, const int32_t *flags, int32_t nXpertile, int32_t nYpertile, int32_t tile_buffer);
int kgpu_debug_tiletable_lists(const kgpu_tiletable *t, int32_t *n_active, int32_t *active, int32_t *n_ghost, int32_t *ghost,
                               int64_t *n_added, int32_t *n_ops);


/* RedistributeGrid (src/Redistribute.f90:203-247) walks one sorted list sequentially.  The library runs the
 * same walk as a dependency-ordered wave (cells more than two apart commute); `on` = 1 makes it walk the list
 * with ONE thread exactly as the reference does, so that a test can require both to agree bit for bit. */
int kgpu_debug_sequential_walk(kgpu_handle *h, int32_t on);
/* `on` = 1 drives a single periodic device through the walk the decomposed runs use (patches gathered from
 * every rank + canonical slots, kgpu_morpho.cuh) instead of the in-place one. */
int kgpu_debug_global_walk(kgpu_handle *h, int32_t on);
/* How a single device runs a morphodynamic Runge-Kutta stage: level 0 = three kernels (E - D, bed, cells: the default, and
 * what the decomposed runs use), 1 = E - D, then morpho_stage_kernel (bed and cells fused), 2 = morpho_stage_kernel alone.
 * The fused forms are measured alternatives (slower on B200); a test requires all three to agree bit for bit. */
int kgpu_debug_morpho_fusion(kgpu_handle *h, int32_t level);
/* Shrink the device buffer of RedistributeGrid's list to `entries` so that a small test overflows it and
 * exercises the enlargement path (the library's default holds 65 536 entries). */
int kgpu_debug_redist_capacity(kgpu_handle *h, int32_t entries);

/* Test probe (no device needed): the host bookkeeping of RedistributeGrid across ranks -- global walk order
 * (src/Redistribute.f90:69-101 on global indices) and canonical patch slots.  geometry10 = {ranks, slots per
 * rank, ranks per row, NX, NY of one block, nXpertile, nYpertile, nXtiles, nYtiles (whole domain), isOneD};
 * the lists are rank-major with `slots per rank` entries each, counts[r] of them valid, LOCAL cell indices.
 * Outputs sized for sum(counts) entries: patch[n], vslot[n*16], cslot[n*9]; n_unique2 = {vertices, cells}. */
int kgpu_debug_redist_tables(const int32_t *geometry10, const int32_t *counts, const double *excess,
                             const int32_t *li, const int32_t *lj, int32_t *n_out, int32_t *patch,
                             int32_t *vslot, int32_t *cslot, int32_t *n_unique2);


#ifdef __cplusplus
}
#endif
#endif /* KESTREL_GPU_DEBUG_H */
