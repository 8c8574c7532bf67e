! kestrel_gpu.f90 -- ISO_C_BINDING interface to libkestrel_gpu (include/kestrel_gpu.h).
!
! The Fortran side of the drop-in boundary: one interface block per public entry point of
! include/kestrel_gpu.h, the bind(C) mirrors of its three structs, and its enums as named constants.
! Written in the style the host already uses for its C dependencies (src/GeoTiffRead.f90:43-108,
! src/utm.f90:53-104).  Fortran 2008; no compiler exists in the build image of this repository, so the
! file is checked mechanically instead: tests/test_fortran_binding.py parses every interface below and the
! C header and requires names, argument counts, by-value / by-reference passing and C types to agree, and
! the derived types to list the header's struct members in the same order with the same types.
!
! Add this file and IntegrateTo_gpu.f90 to kestrel_SOURCES in src/Makefile.am:16-23 (after Grid.f90,
! before TimeStepper.f90) and link with -lkestrel_gpu.
module kestrel_gpu

   use, intrinsic :: iso_c_binding

   implicit none


   ! ---- status codes
   integer(c_int), parameter :: KGPU_OK = 0
   integer(c_int), parameter :: KGPU_ERR_ARG = 1
   integer(c_int), parameter :: KGPU_ERR_CUDA = 2
   integer(c_int), parameter :: KGPU_ERR_HALT_BC = 3      ! AddTile's FatalErrorMessage, src/UpdateTiles.f90:63-65
   integer(c_int), parameter :: KGPU_ERR_DT = 4
   integer(c_int), parameter :: KGPU_ERR_UNSUPPORTED = 5

   ! ---- closure enums: which procedure pointer Params_Set / Solver_Set bound (src/RunSettings.f90:254-268)
   integer(c_int32_t), parameter :: KGPU_BC_HALT = 0, KGPU_BC_PERIODIC = 1, KGPU_BC_DIRICHLET = 2, KGPU_BC_SPONGE = 3
   integer(c_int32_t), parameter :: KGPU_LIM_MINMOD1 = 0, KGPU_LIM_MINMOD2 = 1, KGPU_LIM_NONE = 2, &
                                    KGPU_LIM_VANALBADA = 3, KGPU_LIM_WENO = 4
   integer(c_int32_t), parameter :: KGPU_DRAG_CHEZY = 0, KGPU_DRAG_COULOMB = 1, KGPU_DRAG_VOELLMY = 2, &
                                    KGPU_DRAG_POULIQUEN = 3, KGPU_DRAG_EDWARDS2019 = 4, KGPU_DRAG_VARIABLE = 5, &
                                    KGPU_DRAG_MANNING = 6
   integer(c_int32_t), parameter :: KGPU_ERO_OFF = 0, KGPU_ERO_SIMPLE = 1, KGPU_ERO_FLUID = 2, &
                                    KGPU_ERO_GRANULAR = 3, KGPU_ERO_MIXED = 4
   integer(c_int32_t), parameter :: KGPU_DEP_NONE = 0, KGPU_DEP_SIMPLE = 1, KGPU_DEP_SPEARMAN_MANNING = 2
   integer(c_int32_t), parameter :: KGPU_EROTRANS_SMOOTH = 0, KGPU_EROTRANS_STEP = 1, KGPU_EROTRANS_OFF = 2
   integer(c_int32_t), parameter :: KGPU_DAMP_NONE = 0, KGPU_DAMP_TANH = 1, KGPU_DAMP_RAT3 = 2
   integer(c_int32_t), parameter :: KGPU_SWITCH_TANH = 0, KGPU_SWITCH_RAT3 = 1, KGPU_SWITCH_COS = 2, &
                                    KGPU_SWITCH_LINEAR = 3, KGPU_SWITCH_EQUAL = 4, KGPU_SWITCH_ZERO = 5, &
                                    KGPU_SWITCH_ONE = 6, KGPU_SWITCH_STEP = 7
   integer(c_int32_t), parameter :: KGPU_TOPOG_FLAT = 0, KGPU_TOPOG_XSLOPE = 1, KGPU_TOPOG_YSLOPE = 2, &
                                    KGPU_TOPOG_XYSLOPE = 3, KGPU_TOPOG_XSINSLOPE = 4, KGPU_TOPOG_XYSINSLOPE = 5, &
                                    KGPU_TOPOG_XHUMP = 6, KGPU_TOPOG_XTANH = 7, KGPU_TOPOG_XPARAB = 8, &
                                    KGPU_TOPOG_XYPARAB = 9, KGPU_TOPOG_XBISLOPE = 10, KGPU_TOPOG_X2SLOPES = 11, &
                                    KGPU_TOPOG_USGS = 12, KGPU_TOPOG_FLUME = 13, KGPU_TOPOG_CHANNEL_POWERLAW = 14, &
                                    KGPU_TOPOG_CHANNEL_TRAPEZIUM = 15, KGPU_TOPOG_XTRISLOPE = 16

   ! ---- struct kgpu_source: mirror of type Sources (src/RunSettings.f90:101-109)
   type, bind(C) :: kgpu_source
      real(c_double) :: x, y, radius
      integer(c_int32_t) :: num_cells_in_src
      integer(c_int32_t) :: n_series
      type(c_ptr) :: time
      type(c_ptr) :: flux
      type(c_ptr) :: psi
   end type kgpu_source

   ! ---- struct kgpu_cap / kgpu_cube: types Caps and Cubes (src/RunSettings.f90:56-88) for kgpu_load_source_conditions
   integer(c_int32_t), parameter :: KGPU_SHAPE_FLAT = 0, KGPU_SHAPE_PARA = 1, KGPU_SHAPE_LEVEL = 2
   type, bind(C) :: kgpu_cap
      real(c_double) :: x, y, radius, height, u, v, psi
      integer(c_int32_t) :: shape
      integer(c_int32_t) :: pad
   end type kgpu_cap
   type, bind(C) :: kgpu_cube
      real(c_double) :: x, y, length, width, height, u, v, psi
      integer(c_int32_t) :: shape
      integer(c_int32_t) :: pad
   end type kgpu_cube

   ! ---- struct kgpu_params: the RunSet fields the path reads (src/RunSettings.f90:161-340)
   type, bind(C) :: kgpu_params
      integer(c_int32_t) :: struct_bytes
      integer(c_int32_t) :: nXpertile, nYpertile, nXtiles, nYtiles
      integer(c_int32_t) :: isOneD
      real(c_double) :: deltaX, deltaY
      real(c_double) :: xSize, ySize
      integer(c_int32_t) :: bcs
      integer(c_int32_t) :: pad0
      real(c_double) :: bcsHnval, bcsuval, bcsvval, bcspsival
      integer(c_int32_t) :: geometric_factors
      integer(c_int32_t) :: MorphodynamicsOn
      real(c_double) :: g, rhow, rhos, gred
      real(c_double) :: ChezyCo, ManningCo, CoulombCo
      real(c_double) :: PouliquenMinSlope, PouliquenMaxSlope, PouliquenIntermediateSlope, PouliquenBeta
      real(c_double) :: Edwards2019betastar, Edwards2019kappa, Edwards2019Gamma
      real(c_double) :: VoellmySwitchRate, VoellmySwitchValue
      real(c_double) :: EroRate, EroRateGranular, CriticalShields, EroDepth, EroCriticalHeight
      real(c_double) :: BedPorosity, maxPack, SolidDiameter, ws0, nsettling, EddyViscosity
      real(c_double) :: heightThreshold
      real(c_double) :: cfl, diffusiveTimeScale, maxdt
      real(c_double) :: tstart
      integer(c_int32_t) :: TileBuffer
      integer(c_int32_t) :: SpongeLayer
      real(c_double) :: SpongeStrength
      integer(c_int32_t) :: limiter, drag, erosion, deposition, erosion_transition, morpho_damp, fswitch
      integer(c_int32_t) :: n_sources
      type(c_ptr) :: sources
      type(c_funptr) :: heights
      type(c_ptr) :: heights_ctx
      integer(c_int32_t) :: device
      integer(c_int32_t) :: arithmetic
      integer(c_int32_t) :: comm_rank, comm_size, comm_px, comm_py
   end type kgpu_params

   ! ---- struct kgpu_step_info
   type, bind(C) :: kgpu_step_info
      real(c_double) :: t
      real(c_double) :: dt_last
      integer(c_int64_t) :: nsteps
      integer(c_int64_t) :: nrefines
      integer(c_int64_t) :: ntiles_added
   end type kgpu_step_info

   ! ---- the heights callback: int (*kgpu_heights_fn)(void *ctx, int32_t tile_id, double *b0_vertices)
   abstract interface
      function kgpu_heights_fn(ctx, tile_id, b0_vertices) bind(C) result(rc)
         import :: c_int, c_int32_t, c_double, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int32_t), value :: tile_id
         real(c_double), intent(out) :: b0_vertices(*)
         integer(c_int) :: rc
      end function kgpu_heights_fn
   end interface

   interface

      ! int kgpu_create(const kgpu_params *p, kgpu_handle **h);
      function kgpu_create(p, h) bind(C, name="kgpu_create") result(rc)
         import :: c_int, c_ptr, kgpu_params
         type(kgpu_params), intent(in) :: p
         type(c_ptr), intent(out) :: h
         integer(c_int) :: rc
      end function kgpu_create

      ! int kgpu_destroy(kgpu_handle *h);
      function kgpu_destroy(h) bind(C, name="kgpu_destroy") result(rc)
         import :: c_int, c_ptr
         type(c_ptr), value :: h
         integer(c_int) :: rc
      end function kgpu_destroy

      ! const char *kgpu_last_error(const kgpu_handle *h);
      function kgpu_last_error(h) bind(C, name="kgpu_last_error") result(msg)
         import :: c_ptr
         type(c_ptr), value :: h
         type(c_ptr) :: msg
      end function kgpu_last_error

      ! int kgpu_upload_tile(kgpu_handle *h, int32_t tile_id, const double *u13, const double *b0_vertices,
      !                      const double *bt_vertices, const double *maxima, const double *tfirst,
      !                      int32_t contains_source);
      function kgpu_upload_tile(h, tile_id, u13, b0_vertices, bt_vertices, maxima, tfirst, contains_source) &
            bind(C, name="kgpu_upload_tile") result(rc)
         import :: c_int, c_int32_t, c_ptr
         type(c_ptr), value :: h
         integer(c_int32_t), value :: tile_id
         type(c_ptr), value :: u13
         type(c_ptr), value :: b0_vertices
         type(c_ptr), value :: bt_vertices
         type(c_ptr), value :: maxima
         type(c_ptr), value :: tfirst
         integer(c_int32_t), value :: contains_source
         integer(c_int) :: rc
      end function kgpu_upload_tile

      ! int kgpu_integrate_to(kgpu_handle *h, double tend, int64_t max_steps, kgpu_step_info *info);
      function kgpu_integrate_to(h, tend, max_steps, info) bind(C, name="kgpu_integrate_to") result(rc)
         import :: c_int, c_int64_t, c_double, c_ptr, kgpu_step_info
         type(c_ptr), value :: h
         real(c_double), value :: tend
         integer(c_int64_t), value :: max_steps
         type(kgpu_step_info), intent(out) :: info
         integer(c_int) :: rc
      end function kgpu_integrate_to

      ! int kgpu_active_tiles(kgpu_handle *h, int32_t *n, int32_t *ids);
      function kgpu_active_tiles(h, n, ids) bind(C, name="kgpu_active_tiles") result(rc)
         import :: c_int, c_int32_t, c_ptr
         type(c_ptr), value :: h
         integer(c_int32_t), intent(out) :: n
         type(c_ptr), value :: ids
         integer(c_int) :: rc
      end function kgpu_active_tiles

      ! int kgpu_ghost_tiles(kgpu_handle *h, int32_t *n, int32_t *ids);
      function kgpu_ghost_tiles(h, n, ids) bind(C, name="kgpu_ghost_tiles") result(rc)
         import :: c_int, c_int32_t, c_ptr
         type(c_ptr), value :: h
         integer(c_int32_t), intent(out) :: n
         type(c_ptr), value :: ids
         integer(c_int) :: rc
      end function kgpu_ghost_tiles

      ! int kgpu_download_tile(kgpu_handle *h, int32_t tile_id, double *u13, double *b0_vertices,
      !                        double *bt_vertices, double *maxima, double *tfirst);
      function kgpu_download_tile(h, tile_id, u13, b0_vertices, bt_vertices, maxima, tfirst) &
            bind(C, name="kgpu_download_tile") result(rc)
         import :: c_int, c_int32_t, c_ptr
         type(c_ptr), value :: h
         integer(c_int32_t), value :: tile_id
         type(c_ptr), value :: u13
         type(c_ptr), value :: b0_vertices
         type(c_ptr), value :: bt_vertices
         type(c_ptr), value :: maxima
         type(c_ptr), value :: tfirst
         integer(c_int) :: rc
      end function kgpu_download_tile

      ! int kgpu_upload_domain(kgpu_handle *h, const double *q4, const double *b0_vertices, const double *bt_vertices);
      function kgpu_upload_domain(h, q4, b0_vertices, bt_vertices) bind(C, name="kgpu_upload_domain") result(rc)
         import :: c_int, c_ptr
         type(c_ptr), value :: h
         type(c_ptr), value :: q4
         type(c_ptr), value :: b0_vertices
         type(c_ptr), value :: bt_vertices
         integer(c_int) :: rc
      end function kgpu_upload_domain

      ! int kgpu_download_domain(kgpu_handle *h, double *q4, double *bt_vertices);
      function kgpu_download_domain(h, q4, bt_vertices) bind(C, name="kgpu_download_domain") result(rc)
         import :: c_int, c_ptr
         type(c_ptr), value :: h
         type(c_ptr), value :: q4
         type(c_ptr), value :: bt_vertices
         integer(c_int) :: rc
      end function kgpu_download_domain

      ! int kgpu_output_begin(kgpu_handle *h, double *q4, double *bt_vertices);
      function kgpu_output_begin(h, q4, bt_vertices) bind(C, name="kgpu_output_begin") result(rc)
         import :: c_int, c_ptr
         type(c_ptr), value :: h
         type(c_ptr), value :: q4
         type(c_ptr), value :: bt_vertices
         integer(c_int) :: rc
      end function kgpu_output_begin

      ! int kgpu_output_wait(kgpu_handle *h);
      function kgpu_output_wait(h) bind(C, name="kgpu_output_wait") result(rc)
         import :: c_int, c_ptr
         type(c_ptr), value :: h
         integer(c_int) :: rc
      end function kgpu_output_wait

      ! int kgpu_set_topography_function(kgpu_handle *h, int32_t func, const double *params, int32_t nparams);
      function kgpu_set_topography_function(h, func, params, nparams) &
            bind(C, name="kgpu_set_topography_function") result(rc)
         import :: c_int, c_int32_t, c_ptr
         type(c_ptr), value :: h
         integer(c_int32_t), value :: func
         type(c_ptr), value :: params
         integer(c_int32_t), value :: nparams
         integer(c_int) :: rc
      end function kgpu_set_topography_function

      ! int kgpu_set_topography_raster(kgpu_handle *h, const double *elev, int32_t nx, int32_t ny, double origin_x,
      !                                double origin_y, double pixel_w, double pixel_h, double centre_e, double centre_n);
      function kgpu_set_topography_raster(h, elev, nx, ny, origin_x, origin_y, pixel_w, pixel_h, centre_e, centre_n) &
            bind(C, name="kgpu_set_topography_raster") result(rc)
         import :: c_int, c_int32_t, c_double, c_ptr
         type(c_ptr), value :: h
         type(c_ptr), value :: elev
         integer(c_int32_t), value :: nx
         integer(c_int32_t), value :: ny
         real(c_double), value :: origin_x
         real(c_double), value :: origin_y
         real(c_double), value :: pixel_w
         real(c_double), value :: pixel_h
         real(c_double), value :: centre_e
         real(c_double), value :: centre_n
         integer(c_int) :: rc
      end function kgpu_set_topography_raster

      ! int kgpu_load_source_conditions(kgpu_handle *h, const kgpu_cap *caps, int32_t ncaps, const kgpu_cube *cubes,
      !                                 int32_t ncubes, int32_t *num_cells_in_src);
      function kgpu_load_source_conditions(h, caps, ncaps, cubes, ncubes, num_cells_in_src) &
            bind(C, name="kgpu_load_source_conditions") result(rc)
         import :: c_int, c_int32_t, c_ptr
         type(c_ptr), value :: h
         type(c_ptr), value :: caps
         integer(c_int32_t), value :: ncaps
         type(c_ptr), value :: cubes
         integer(c_int32_t), value :: ncubes
         type(c_ptr), value :: num_cells_in_src
         integer(c_int) :: rc
      end function kgpu_load_source_conditions

      ! int kgpu_comm_id_bytes(void);
      function kgpu_comm_id_bytes() bind(C, name="kgpu_comm_id_bytes") result(nbytes)
         import :: c_int
         integer(c_int) :: nbytes
      end function kgpu_comm_id_bytes

      ! int kgpu_comm_create_id(void *id_out);
      function kgpu_comm_create_id(id_out) bind(C, name="kgpu_comm_create_id") result(rc)
         import :: c_int, c_ptr
         type(c_ptr), value :: id_out
         integer(c_int) :: rc
      end function kgpu_comm_create_id

      ! int kgpu_comm_attach(kgpu_handle *h, const void *id);
      function kgpu_comm_attach(h, id) bind(C, name="kgpu_comm_attach") result(rc)
         import :: c_int, c_ptr
         type(c_ptr), value :: h
         type(c_ptr), value :: id
         integer(c_int) :: rc
      end function kgpu_comm_attach

      ! int kgpu_comm_block(kgpu_handle *h, int32_t *tx0, int32_t *ty0, int32_t *ntx, int32_t *nty);
      function kgpu_comm_block(h, tx0, ty0, ntx, nty) bind(C, name="kgpu_comm_block") result(rc)
         import :: c_int, c_int32_t, c_ptr
         type(c_ptr), value :: h
         integer(c_int32_t), intent(out) :: tx0
         integer(c_int32_t), intent(out) :: ty0
         integer(c_int32_t), intent(out) :: ntx
         integer(c_int32_t), intent(out) :: nty
         integer(c_int) :: rc
      end function kgpu_comm_block

      ! int64_t kgpu_launch_count(const kgpu_handle *h);
      function kgpu_launch_count(h) bind(C, name="kgpu_launch_count") result(n)
         import :: c_int64_t, c_ptr
         type(c_ptr), value :: h
         integer(c_int64_t) :: n
      end function kgpu_launch_count

      ! int kgpu_rhs_timing(kgpu_handle *h, double *ms, int64_t *launches, int32_t reset);
      function kgpu_rhs_timing(h, ms, launches, reset) bind(C, name="kgpu_rhs_timing") result(rc)
         import :: c_int, c_int32_t, c_int64_t, c_double, c_ptr
         type(c_ptr), value :: h
         real(c_double), intent(out) :: ms
         integer(c_int64_t), intent(out) :: launches
         integer(c_int32_t), value :: reset
         integer(c_int) :: rc
      end function kgpu_rhs_timing

      ! int kgpu_morpho_stats(const kgpu_handle *h, int64_t *redistributed_cells, int64_t *list_enlargements);
      function kgpu_morpho_stats(h, redistributed_cells, list_enlargements) bind(C, name="kgpu_morpho_stats") result(rc)
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: h
         integer(c_int64_t), intent(out) :: redistributed_cells
         integer(c_int64_t), intent(out) :: list_enlargements
         integer(c_int) :: rc
      end function kgpu_morpho_stats

      ! void *kgpu_stream(kgpu_handle *h);
      function kgpu_stream(h) bind(C, name="kgpu_stream") result(stream)
         import :: c_ptr
         type(c_ptr), value :: h
         type(c_ptr) :: stream
      end function kgpu_stream

      ! const char *kgpu_version(void);
      function kgpu_version() bind(C, name="kgpu_version") result(msg)
         import :: c_ptr
         type(c_ptr) :: msg
      end function kgpu_version

   end interface

contains

   ! NUL-terminated C string -> Fortran string (for kgpu_last_error / kgpu_version).
   function kgpu_c_string(cstr) result(str)
      type(c_ptr), intent(in) :: cstr
      character(len=:), allocatable :: str
      character(kind=c_char), pointer :: chars(:)
      integer :: n

      str = ""
      if (.not. c_associated(cstr)) return
      call c_f_pointer(cstr, chars, [4096])
      n = 0
      do while (n < 4096)
         if (chars(n + 1) == c_null_char) exit
         n = n + 1
      end do
      allocate (character(len=n) :: str)
      if (n > 0) str = transfer(chars(1:n), str)
   end function kgpu_c_string

end module kestrel_gpu
