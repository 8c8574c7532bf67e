/*
 * kestrel_gpu.h -- C-ABI of libkestrel_gpu: the B200 replacement for the body of
 * Kestrel's serial time-integration loop.
 *
 * The reference has no FFI around its compute loop; the seam this ABI cuts is
 *     timestepper_module::IntegrateTo(RunParams, tend, grid)   src/TimeStepper.f90:116
 * called from Run (src/TimeStepper.f90:102) between output calls that read
 * grid%tileContainer(k)%u / %b0 / %bt / maxima / %tfirst and grid%activeTiles
 * (src/Output.f90:845-916, 1206-1257).  The Fortran host keeps its input parser,
 * closure selection, DEM handling and writers; it binds these symbols through
 * ISO_C_BINDING in the style it already uses for GDAL/PROJ
 * (src/GeoTiffRead.f90:43-108, src/utm.f90:53-104).  INTEGRATION.md shows the
 * Fortran interface module.
 *
 * Conventions
 *  - plain C, no C++/torch types; every entry point returns an int status.
 *  - all reals are IEEE binary64 (wp = c_double, src/SetPrecision.f90:36).
 *  - tile ids are the reference's 1-based TileID = grid_i + (grid_j-1)*nXtiles
 *    (src/Grid.f90:260-268).
 *  - per-tile arrays use Fortran order exactly as the host owns them:
 *      u13      : u(d,i,j)  d=1..13 fastest, then i=1..nXpertile, then j   (src/Grid.f90:83)
 *      vertices : b0(i,j), bt(i,j)  i=1..nXpertile+1 fastest, j=1..nYpertile+1
 *      maxima   : 5 blocks [Hnmax, umax, emax, dmax, psimax], each (i,j,k) with
 *                 k=1 value, k=2 time (src/Grid.f90:96-100)
 *      tfirst   : (i,j)
 *  - host buffers are borrowed for the duration of a call only.
 *  - one host thread per handle; the library is not re-entrant per handle.
 */
#ifndef KESTREL_GPU_H
#define KESTREL_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ---------------------------------------------------------- */
enum {
   KGPU_OK = 0,
   KGPU_ERR_ARG = 1,        /* bad argument                                         */
   KGPU_ERR_CUDA = 2,       /* CUDA / NCCL failure                                  */
   KGPU_ERR_HALT_BC = 3,    /* flow reached the domain edge with bcs = halt:
                               AddTile's FatalErrorMessage, src/UpdateTiles.f90:63-65 */
   KGPU_ERR_DT = 4,         /* time step underflow / non-finite state               */
   KGPU_ERR_UNSUPPORTED = 5
};

/* ---- closure enums: the strings RunSet already stores (src/RunSettings.f90:254-268),
 *      i.e. which procedure pointer Params_Set / Solver_Set bound ----------------- */
enum { KGPU_BC_HALT = 0, KGPU_BC_PERIODIC = 1, KGPU_BC_DIRICHLET = 2, KGPU_BC_SPONGE = 3 };
/* src/SolverSettings.f90:102-126, src/Limiters.f90:83-184 */
enum { KGPU_LIM_MINMOD1 = 0, KGPU_LIM_MINMOD2 = 1, KGPU_LIM_NONE = 2, KGPU_LIM_VANALBADA = 3, KGPU_LIM_WENO = 4 };
/* src/Parameters.f90:268-297, src/Closures.f90:365-558 */
enum { KGPU_DRAG_CHEZY = 0, KGPU_DRAG_COULOMB = 1, KGPU_DRAG_VOELLMY = 2, KGPU_DRAG_POULIQUEN = 3,
       KGPU_DRAG_EDWARDS2019 = 4, KGPU_DRAG_VARIABLE = 5, KGPU_DRAG_MANNING = 6 };
/* src/Parameters.f90:359-387, src/Closures.f90:566-675 */
enum { KGPU_ERO_OFF = 0, KGPU_ERO_SIMPLE = 1, KGPU_ERO_FLUID = 2, KGPU_ERO_GRANULAR = 3, KGPU_ERO_MIXED = 4 };
/* src/Parameters.f90:208-224, src/Closures.f90:320-356 */
enum { KGPU_DEP_NONE = 0, KGPU_DEP_SIMPLE = 1, KGPU_DEP_SPEARMAN_MANNING = 2 };
/* src/Parameters.f90:404-420, src/Closures.f90:684-732 */
enum { KGPU_EROTRANS_SMOOTH = 0, KGPU_EROTRANS_STEP = 1, KGPU_EROTRANS_OFF = 2 };
/* src/Parameters.f90:226-242, src/Closures.f90:744-791 */
enum { KGPU_DAMP_NONE = 0, KGPU_DAMP_TANH = 1, KGPU_DAMP_RAT3 = 2 };
/* src/Parameters.f90:175-206, src/Closures.f90:797-915 */
enum { KGPU_SWITCH_TANH = 0, KGPU_SWITCH_RAT3 = 1, KGPU_SWITCH_COS = 2, KGPU_SWITCH_LINEAR = 3,
       KGPU_SWITCH_EQUAL = 4, KGPU_SWITCH_ZERO = 5, KGPU_SWITCH_ONE = 6, KGPU_SWITCH_STEP = 7 };

/* Flux source: mirror of type Sources (src/RunSettings.f90:101-109). */
typedef struct kgpu_source {
   double x, y, radius;
   int32_t num_cells_in_src;   /* NumCellsInSrc, counted by LoadSourceConditions (src/SetSources.f90:227,372) */
   int32_t n_series;           /* nFluxSeries */
   const double *time;         /* [n_series] */
   const double *flux;         /* [n_series] */
   const double *psi;          /* [n_series] */
} kgpu_source;

/* Topography callback: wraps GetHeights (src/dem.f90:360) so DEM / GDAL stays on
 * the host.  Must fill b0 at the (nXpertile+1)*(nYpertile+1) vertices of tile
 * `tile_id`, i fastest.  Called when a tile is activated or becomes a ghost tile
 * (src/UpdateTiles.f90:171, 466).  Return 0 on success. */
typedef int (*kgpu_heights_fn)(void *ctx, int32_t tile_id, double *b0_vertices);

/* POD mirror of the RunSet fields the path reads (src/RunSettings.f90:161-340). */
typedef struct kgpu_params {
   int32_t struct_bytes;       /* = sizeof(kgpu_params), checked by kgpu_create */

   /* Domain (src/DomainSettings.f90:48-228) */
   int32_t nXpertile, nYpertile, nXtiles, nYtiles;
   int32_t isOneD;
   double deltaX, deltaY;
   double xSize, ySize;
   int32_t bcs;                /* KGPU_BC_* */
   int32_t _pad0;
   double bcsHnval, bcsuval, bcsvval, bcspsival;

   /* Parameters (src/Parameters.f90:82-648) */
   int32_t geometric_factors;
   int32_t MorphodynamicsOn;
   double g, rhow, rhos, gred;
   double ChezyCo, ManningCo, CoulombCo;
   double PouliquenMinSlope, PouliquenMaxSlope, PouliquenIntermediateSlope, PouliquenBeta;
   double Edwards2019betastar, Edwards2019kappa, Edwards2019Gamma;
   double VoellmySwitchRate, VoellmySwitchValue;
   double EroRate, EroRateGranular, CriticalShields, EroDepth, EroCriticalHeight;
   double BedPorosity, maxPack, SolidDiameter, ws0, nsettling, EddyViscosity;

   /* Solver (src/SolverSettings.f90:59-268) */
   double heightThreshold;
   double cfl, diffusiveTimeScale, maxdt;
   double tstart;              /* grid%t on entry to the first IntegrateTo (src/Restart.f90:79) */
   int32_t TileBuffer;
   int32_t SpongeLayer;
   double SpongeStrength;

   /* closures */
   int32_t limiter, drag, erosion, deposition, erosion_transition, morpho_damp, fswitch;

   /* flux sources */
   int32_t n_sources;
   const kgpu_source *sources;

   /* topography */
   kgpu_heights_fn heights;
   void *heights_ctx;

   /* library options (no reference counterpart) */
   int32_t device;             /* CUDA device ordinal; -1 = current device */
   int32_t arithmetic;         /* 0 = faithful (no FMA contraction, reference operation order);
                                  1 = contracted (FMA + shared reciprocals), validated to 1e-10 */
   /* 2-D block decomposition of the tile grid over comm_size = comm_px*comm_py processes
      (one per GPU).  comm_size <= 1: single device.  Rank r owns tile columns
      [r%px * nXtiles/px, ...) and tile rows [r/px * nYtiles/py, ...).                     */
   int32_t comm_rank, comm_size, comm_px, comm_py;
} kgpu_params;

typedef struct kgpu_handle kgpu_handle;

/* Statistics of one kgpu_integrate_to call. */
typedef struct kgpu_step_info {
   double t;                   /* grid%t on return                                  */
   double dt_last;             /* dt_hydro of the last completed step               */
   int64_t nsteps;             /* completed passes of the do-while at TimeStepper.f90:151 */
   int64_t nrefines;           /* rolled-back attempts (TimeStepper.f90:183,203,224,245) */
   int64_t ntiles_added;       /* tiles activated by CheckIfNearBoundaries          */
} kgpu_step_info;

/* ---- lifecycle ------------------------------------------------------------- */

/* After LoadSourceConditions (src/main.f90:126).  Copies params and source tables. */
int kgpu_create(const kgpu_params *p, kgpu_handle **h);
int kgpu_destroy(kgpu_handle *h);
const char *kgpu_last_error(const kgpu_handle *h);

/* ---- state in --------------------------------------------------------------- */

/* Upload one active tile (host: grid%tileContainer(tile_id), TileOn).  Marks it
 * active exactly like AddToActiveTiles (src/UpdateTiles.f90:81); its inactive
 * neighbours become ghost tiles (AddGhostTiles, :389) and their heights are
 * requested through the callback.
 *   u13           (13,nX,nY)  -- components 1..4 (w, rhoHnu, rhoHnv, Hnpsi) are state;
 *                               5 (Hn) seeds the first tile-activation scan
 *                               (src/TimeStepper.f90:982); 10..13 are ignored
 *                               (recomputed from the vertices).
 *   b0_vertices   (nX+1,nY+1) or NULL to use the callback
 *   bt_vertices   (nX+1,nY+1) or NULL for zero
 *   maxima        5*(nX,nY,2) or NULL for zero;  tfirst (nX,nY) or NULL for -1
 *   contains_source: TileType%containsSource (src/SetSources.f90:226,371)       */
int kgpu_upload_tile(kgpu_handle *h, int32_t tile_id, const double *u13,
                     const double *b0_vertices, const double *bt_vertices,
                     const double *maxima, const double *tfirst, int32_t contains_source);

/* ---- the path --------------------------------------------------------------- */

/* Body of IntegrateTo (src/TimeStepper.f90:116-277): runs the do-while loop to tend,
 * including tile activation, dt selection, Strang splitting and rollback.
 * max_steps > 0 stops after that many completed steps (benchmarking; the
 * reference has no such limit); 0 = run to tend. */
int kgpu_integrate_to(kgpu_handle *h, double tend, int64_t max_steps, kgpu_step_info *info);

/* ---- state out --------------------------------------------------------------- */

/* grid%activeTiles%List in the reference's order (ascending, src/utilities.f90:260).
 * ids may be NULL to query n only.
 * Tile ids everywhere in this interface (upload, download, the two lists, the heights
 * callback) number the WHOLE tile grid, 1-based with tx fastest, as the reference's
 * tileContainer index does.  A decomposed handle (comm_size > 1) lists and downloads
 * only the tiles of its own block (kgpu_comm_block) and answers KGPU_ERR_ARG for
 * another rank's tile; see kgpu_comm_attach for uploads. */
int kgpu_active_tiles(kgpu_handle *h, int32_t *n, int32_t *ids);
int kgpu_ghost_tiles(kgpu_handle *h, int32_t *n, int32_t *ids);

/* Fill exactly what the writers read (src/Output.f90:845-916, 1206-1257).  Any
 * pointer may be NULL.  u13(6:7) carry the values desingularised before the final
 * implicit momentum correction, as in the reference (src/TimeStepper.f90:501-517). */
int kgpu_download_tile(kgpu_handle *h, int32_t tile_id, double *u13,
                       double *b0_vertices, double *bt_vertices,
                       double *maxima, double *tfirst);

/* ---- bulk variants for all-tiles-active domains (same data, one call) --------
 * Flat arrays over the whole domain, i fastest: cells NX*NY with NX = nXtiles*nXpertile,
 * vertices (NX+1)*(NY+1).  q4 = [w | rhoHnu | rhoHnv | Hnpsi] planes.  Every tile
 * becomes active (requires bcs = periodic, src/UpdateTiles.f90:61-69).          */
int kgpu_upload_domain(kgpu_handle *h, const double *q4, const double *b0_vertices,
                       const double *bt_vertices);
int kgpu_download_domain(kgpu_handle *h, double *q4, double *bt_vertices);

/* ---- asynchronous output gather --------------------------
 * Replaces the blocking read of grid%tileContainer before OutputSolutionData
 * (src/TimeStepper.f90:106-109, src/Output.f90:170-227) when the host wants to keep
 * integrating while the output interval is written.  kgpu_output_begin takes a device-side
 * snapshot of the current state (4 planes, + the bed with morphodynamics; a device copy at HBM
 * speed) and starts its transfer to the host buffers on a copy stream; it returns at once and
 * kgpu_integrate_to may be called again immediately.  kgpu_output_wait blocks until the
 * buffers hold the snapshot.  Same layout as kgpu_download_domain; the buffers should be
 * page-locked (cudaHostRegister / pinned allocation) for the copy to overlap the time steps,
 * and must stay valid and untouched until kgpu_output_wait returns.  One output may be in
 * flight per handle: a second kgpu_output_begin first waits for the previous one.        */
int kgpu_output_begin(kgpu_handle *h, double *q4, double *bt_vertices);
int kgpu_output_wait(kgpu_handle *h);

/* ---- analytic topography on the device --------------------
 * For `Topog Type = Function` inputs (src/TopogFuncs.f90) the library can evaluate the heights of a
 * tile itself when the tile is activated, instead of calling the heights callback and uploading
 * the result (src/dem.f90:360-415 GetHeights stays the path for rasters).  Coordinates are those of
 * src/Grid.f90:339-353 and src/UpdateTiles.f90:288-325; params as given in `Topog params`.  The
 * algebraic functions reproduce the host's values bit for bit, the transcendental ones to the
 * rounding of the device's libm.  Call after kgpu_create, before the first upload / tile activation;
 * func < 0 returns to the callback.                                                            */
enum { KGPU_TOPOG_FLAT = 0, KGPU_TOPOG_XSLOPE, KGPU_TOPOG_YSLOPE, KGPU_TOPOG_XYSLOPE, KGPU_TOPOG_XSINSLOPE,
       KGPU_TOPOG_XYSINSLOPE, KGPU_TOPOG_XHUMP, KGPU_TOPOG_XTANH, KGPU_TOPOG_XPARAB, KGPU_TOPOG_XYPARAB,
       KGPU_TOPOG_XBISLOPE, KGPU_TOPOG_X2SLOPES, KGPU_TOPOG_USGS, KGPU_TOPOG_FLUME, KGPU_TOPOG_CHANNEL_POWERLAW,
       KGPU_TOPOG_CHANNEL_TRAPEZIUM, KGPU_TOPOG_XTRISLOPE };
int kgpu_set_topography_function(kgpu_handle *h, int32_t func, const double *params, int32_t nparams);

/* ---- DEM ingest on the device -----------------------------------
 * For raster topographies (`Topog Type = DEM / raster / SRTM`) the host keeps GDAL: it reads the raster section that
 * covers the domain once (GetRasterData, src/dem.f90:193-257) and hands it over; the library then resamples the heights
 * of every tile it activates itself -- TileHeightData (src/dem.f90:260-356): the mean of four bicubic interpolations
 * (Bicubic_r, src/Interp2d.f90:214-292) at the vertex +- half a cell -- instead of calling the heights callback per tile.
 *   elev      : nx * ny pixel values, x fastest: Elev(i, j) of GetRasterData, nodata already replaced
 *   origin_*, pixel_* : rOX, rOY, rdX, rdY of that section (pixel_h is negative for north-up rasters)
 *   centre_e, centre_n: RunParams%centerUTM
 * Copies the raster to the device; func < 0 of kgpu_set_topography_function returns to the callback. */
int kgpu_set_topography_raster(kgpu_handle *h, const double *elev, int32_t nx, int32_t ny, double origin_x, double origin_y,
                               double pixel_w, double pixel_h, double centre_e, double centre_n);

/* ---- initial conditions on the device ------------------------
 * LoadSourceConditions (src/SetSources.f90:47-392) inside the library: the caps and cubes of the input file are
 * rasterised on the device instead of on the host followed by one kgpu_upload_tile per tile.  Every tile with a
 * cell centre inside a cap (R2 <= R^2), a cube or a flux-source disc is switched on through the library's
 * AddTile (heights from the callback or from kgpu_set_topography_function) in the reference's tile order; the
 * shapes are then added cell by cell in the reference's order (caps, then cubes) and operation order, including
 * its quirks (1-D flat caps, 1-D parabolic momentum, psimax in 1-D: src/SetSources.f90:132-144); containsSource
 * and NumCellsInSrc (:367-372) are set from the handle's source table, the first tile-activation scan is seeded
 * from the rasterised depth (src/TimeStepper.f90:982).  With bcs = halt a shape that reaches an edge tile returns
 * KGPU_ERR_HALT_BC (src/UpdateTiles.f90:63-65).  num_cells_in_src (n_sources entries) may be NULL.
 * Call once, after kgpu_create (and kgpu_set_topography_function), instead of the kgpu_upload_tile calls. */
enum { KGPU_SHAPE_FLAT = 0, KGPU_SHAPE_PARA = 1, KGPU_SHAPE_LEVEL = 2 };
typedef struct kgpu_cap {      /* type Caps, src/RunSettings.f90:56-66 */
   double x, y, radius, height, u, v, psi;
   int32_t shape;              /* KGPU_SHAPE_* */
   int32_t _pad;
} kgpu_cap;
typedef struct kgpu_cube {     /* type Cubes, src/RunSettings.f90:78-88 */
   double x, y, length, width, height, u, v, psi;
   int32_t shape;              /* KGPU_SHAPE_FLAT or KGPU_SHAPE_LEVEL */
   int32_t _pad;
} kgpu_cube;
int kgpu_load_source_conditions(kgpu_handle *h, const kgpu_cap *caps, int32_t ncaps, const kgpu_cube *cubes,
                                int32_t ncubes, int32_t *num_cells_in_src);

/* ---- multi-GPU (one process per GPU; 2-D block decomposition of the tile grid) */

/* Bytes of an opaque communicator id (ncclUniqueId). */
int kgpu_comm_id_bytes(void);
/* Rank 0 creates the id; the host broadcasts it to all ranks by its own means. */
int kgpu_comm_create_id(void *id_out);
/* Call between kgpu_create (with comm_* set in the params) and the first upload; collective
 * over all ranks.  With a communicator attached, kgpu_upload_domain / kgpu_download_domain
 * take the rank's own block (local NX x NY cells, (NX+1) x (NY+1) vertices); 2-cell halos of
 * the four primary fields move by ncclSend/ncclRecv after every stage, overlapped with the
 * interior of the stage kernel, and one ncclAllReduce(min) per dt decision keeps every rank
 * on the same time step.  The morphodynamic operator exchanges E - D, the stage beds and its
 * centre planes the same way, max-reduces its refine flags, and replays RedistributeGrid
 * identically on every rank over all-gathered patches.
 * Two modes, chosen by the boundary conditions:
 *  - periodic: every tile active, state moved in bulk with kgpu_upload_domain /
 *    kgpu_download_domain (kgpu_upload_tile takes the rank's own tiles);
 *  - halt / dirichlet: dynamic tiles as on one device (src/UpdateTiles.f90,
 *    CheckIfNearBoundaries src/TimeStepper.f90:924-1150): the tile table is replicated, so
 *    kgpu_upload_tile is COLLECTIVE -- every rank passes every initial tile with the same
 *    arguments in the same order -- and the heights callback may be asked for tiles of the
 *    neighbouring ranks that reach into this rank's halo.  kgpu_active_tiles,
 *    kgpu_ghost_tiles and kgpu_download_tile cover the rank's own block; step counters and
 *    tiles added are those of the whole domain.  Tiles need at least 3 cells a side.
 *    kgpu_load_source_conditions is single-device only.
 * In both modes the decomposed run is bit-identical to the single-device run. */
int kgpu_comm_attach(kgpu_handle *h, const void *id);
/* Tile block owned by this handle (0-based global tile coordinates). */
int kgpu_comm_block(kgpu_handle *h, int32_t *tx0, int32_t *ty0, int32_t *ntx, int32_t *nty);

/* ---- introspection ------------------------------------------------------------ */
/* Kernel launches issued by this handle since creation (bench.py's gpu_launches). */
int64_t kgpu_launch_count(const kgpu_handle *h);
/* Device time (ms) accumulated in the fused RHS kernels since the last reset, and
 * the number of launches, measured with CUDA events on the launching stream. */
int kgpu_rhs_timing(kgpu_handle *h, double *ms, int64_t *launches, int32_t reset);
/* Morphodynamic bookkeeping since creation: cells handed to RedistributeGrid (src/Redistribute.f90:203-247)
 * and how often its list outgrew the device buffer and was enlarged (the reference's list is unbounded). */
int kgpu_morpho_stats(const kgpu_handle *h, int64_t *redistributed_cells, int64_t *list_enlargements);
/* Stream the handle launches on (cudaStream_t as void*), for external event timing. */
void *kgpu_stream(kgpu_handle *h);
const char *kgpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* KESTREL_GPU_H */
