! IntegrateTo_gpu.f90 -- the host side of the drop-in: IntegrateTo with its body on the GPU.
!
! Replaces the body of timestepper_module::IntegrateTo (src/TimeStepper.f90:116-277).  Run
! (src/TimeStepper.f90:72-113) keeps calling
!     call IntegrateTo(RunParams, timestep, grid)
! between its output calls; to switch a build over, `use kestrel_gpu_host, only: IntegrateTo => IntegrateToGPU`
! in timestepper_module and delete (or rename) the original subroutine.  Everything else of the host -- input
! parser, closure selection, LoadSourceConditions, DEM handling, the txt / NetCDF / KML writers, restart -- is
! untouched and keeps reading grid%tileContainer exactly as before (src/Output.f90:845-916, 1206-1257).
!
! What crosses the boundary (include/kestrel_gpu.h; module kestrel_gpu is its Fortran mirror):
!   first call : kgpu_create(MakeParams(RunParams)) and one kgpu_upload_tile per tile of grid%activeTiles
!                (u, b0, bt, the five maxima, tfirst, containsSource) -- the state LoadSourceConditions or
!                LoadInitialCondition (restart) left in grid%tileContainer;
!   every call : kgpu_integrate_to(tend) = the whole do-while loop of src/TimeStepper.f90:151-275, tile
!                activation included; then the active set and every active tile are read back, tiles the flow
!                switched on meanwhile being allocated through the host's own AddTile (src/UpdateTiles.f90:56);
!   topography : the library asks for the heights of a tile it activates through a callback that wraps
!                GetHeights (src/dem.f90:360), so rasters, SRTM and GDAL stay on the host.
!
! Fortran 2008.  There is no Fortran compiler in the image this repository is built in; the bind(C) signatures are
! cross-checked against the C header by tests/test_fortran_binding.py, and the same call sequence is exercised
! end to end by the C++ host driver (kestrel_b200/host_cpp) and the ctypes binding (kestrel_b200/capi.py).
module kestrel_gpu_host

   use, intrinsic :: iso_c_binding
   use set_precision_module, only: wp
   use messages_module, only: FatalErrorMessage, WarningMessage
   use grid_module, only: GridType, TileType
   use runsettings_module, only: RunSet
   use update_tiles_module, only: AddTile, AllocateTile
   use kestrel_gpu

   implicit none

   private
   public :: IntegrateToGPU
   public :: FinaliseGPU

   ! the library handle lives as long as the run
   type(c_ptr), save :: handle = c_null_ptr

   ! the heights callback reaches the host's data through these (set on every entry to IntegrateToGPU)
   type(RunSet), pointer, save :: cbRunParams => null()
   type(GridType), pointer, save :: cbGrid => null()

   ! flux-source tables handed to kgpu_create: the library copies them during the call, these only have to
   ! outlive it, but keeping them at module scope avoids any doubt about the lifetime of c_loc targets
   type(kgpu_source), allocatable, target, save :: srcTable(:)
   real(c_double), allocatable, target, save :: srcTime(:, :), srcFlux(:, :), srcPsi(:, :)

contains

   ! ------------------------------------------------------------------------------------------------
   ! Integrate the solution data in the numerical grid from t = grid%t to tend.
   ! Same interface as the routine it replaces (src/TimeStepper.f90:116).
   subroutine IntegrateToGPU(RunParams, tend, grid)

      implicit none

      type(RunSet), target, intent(inout) :: RunParams
      real(kind=wp), intent(in) :: tend
      type(GridType), target, intent(inout) :: grid

      type(kgpu_params) :: params
      type(kgpu_step_info) :: info
      integer(c_int) :: rc
      integer(c_int32_t) :: nActive
      integer(c_int32_t), allocatable, target :: ids(:)
      real(c_double), allocatable, target :: maxbuf(:)
      integer :: tt, k, nX, nY

      if (tend <= grid%t) then
         call WarningMessage("End-time Tend is less than initial time")
         return
      end if

      cbRunParams => RunParams
      cbGrid => grid
      nX = RunParams%nXpertile
      nY = RunParams%nYpertile
      allocate (maxbuf(10 * nX * nY))

      ! ---- first call: create the handle and upload the tiles the host has switched on
      if (.not. c_associated(handle)) then
         call MakeParams(RunParams, grid, params)
         rc = kgpu_create(params, handle)
         if (rc /= KGPU_OK) call FatalErrorMessage("kgpu_create failed: no usable CUDA device " // &
                                                   "(libkestrel_gpu has no CPU path)")
         do tt = 1, grid%activeTiles%size
            k = grid%activeTiles%List(tt)
            call PackMaxima(grid%tileContainer(k), nX, nY, maxbuf)
            call UploadTile(grid%tileContainer(k), k, maxbuf)
         end do
      end if

      ! ---- the whole do-while loop of src/TimeStepper.f90:151-275, on the GPU
      rc = kgpu_integrate_to(handle, real(tend, c_double), 0_c_int64_t, info)
      if (rc == KGPU_ERR_HALT_BC) then
         call FatalErrorMessage("Error: tried to add a tile outside the domain." // &
                                new_line('A') // " Try increasing the domain size or repositioning its location.")
      else if (rc /= KGPU_OK) then
         call FatalErrorMessage("kgpu_integrate_to: " // kgpu_c_string(kgpu_last_error(handle)))
      end if
      grid%t = real(info%t, wp)
      grid%dt = real(info%dt_last, wp)

      ! ---- read back what the writers read: the active set, then every active tile
      rc = kgpu_active_tiles(handle, nActive, c_null_ptr)
      allocate (ids(max(1, int(nActive))))
      rc = kgpu_active_tiles(handle, nActive, c_loc(ids))
      do tt = 1, int(nActive)
         k = int(ids(tt))
         ! a tile the flow switched on meanwhile: AddToActiveTiles + AllocateTile + ActivateTile + ghost ring,
         ! exactly what CheckIfNearBoundaries would have called (src/TimeStepper.f90:924-946)
         if (.not. grid%tileContainer(k)%TileOn) call AddTile(grid, k, RunParams)
         call DownloadTile(grid%tileContainer(k), k, maxbuf)
         call UnpackMaxima(maxbuf, nX, nY, grid%tileContainer(k))
      end do

      deallocate (ids, maxbuf)

   end subroutine IntegrateToGPU

   ! Release the device (call once after Run, src/main.f90).
   subroutine FinaliseGPU()
      integer(c_int) :: rc
      if (c_associated(handle)) rc = kgpu_destroy(handle)
      handle = c_null_ptr
   end subroutine FinaliseGPU

   ! ------------------------------------------------------------------------------------------------
   ! One tile to / from the device.  The dummy is a TARGET so that c_loc of its allocatable components is valid.
   subroutine UploadTile(tile, k, maxbuf)
      type(TileType), target, intent(in) :: tile
      integer, intent(in) :: k
      real(c_double), target, contiguous, intent(in) :: maxbuf(:)
      integer(c_int) :: rc
      integer(c_int32_t) :: hasSource

      hasSource = 0_c_int32_t
      if (tile%containsSource) hasSource = 1_c_int32_t
      rc = kgpu_upload_tile(handle, int(k, c_int32_t), c_loc(tile%u), c_loc(tile%b0), c_loc(tile%bt), &
                            c_loc(maxbuf), c_loc(tile%tfirst), hasSource)
      if (rc /= KGPU_OK) call FatalErrorMessage("kgpu_upload_tile: " // kgpu_c_string(kgpu_last_error(handle)))
   end subroutine UploadTile

   subroutine DownloadTile(tile, k, maxbuf)
      type(TileType), target, intent(inout) :: tile
      integer, intent(in) :: k
      real(c_double), target, contiguous, intent(inout) :: maxbuf(:)
      integer(c_int) :: rc

      rc = kgpu_download_tile(handle, int(k, c_int32_t), c_loc(tile%u), c_loc(tile%b0), c_loc(tile%bt), &
                              c_loc(maxbuf), c_loc(tile%tfirst))
      if (rc /= KGPU_OK) call FatalErrorMessage("kgpu_download_tile: " // kgpu_c_string(kgpu_last_error(handle)))
   end subroutine DownloadTile

   ! The five running maxima as one buffer: [Hnmax | umax | emax | dmax | psimax], each (nX, nY, 2) with the value
   ! plane first and the time-of-maximum plane second -- the arrays' own Fortran order (src/Grid.f90:96-100).
   subroutine PackMaxima(tile, nX, nY, buf)
      type(TileType), intent(in) :: tile
      integer, intent(in) :: nX, nY
      real(c_double), intent(out) :: buf(:)
      integer :: n

      n = 2 * nX * nY
      buf(1:n) = reshape(tile%Hnmax, [n])
      buf(n + 1:2 * n) = reshape(tile%umax, [n])
      buf(2 * n + 1:3 * n) = reshape(tile%emax, [n])
      buf(3 * n + 1:4 * n) = reshape(tile%dmax, [n])
      buf(4 * n + 1:5 * n) = reshape(tile%psimax, [n])
   end subroutine PackMaxima

   subroutine UnpackMaxima(buf, nX, nY, tile)
      real(c_double), intent(in) :: buf(:)
      integer, intent(in) :: nX, nY
      type(TileType), intent(inout) :: tile
      integer :: n

      n = 2 * nX * nY
      tile%Hnmax = reshape(buf(1:n), [nX, nY, 2])
      tile%umax = reshape(buf(n + 1:2 * n), [nX, nY, 2])
      tile%emax = reshape(buf(2 * n + 1:3 * n), [nX, nY, 2])
      tile%dmax = reshape(buf(3 * n + 1:4 * n), [nX, nY, 2])
      tile%psimax = reshape(buf(4 * n + 1:5 * n), [nX, nY, 2])
   end subroutine UnpackMaxima

   ! ------------------------------------------------------------------------------------------------
   ! kgpu_params from RunSet (src/RunSettings.f90:161-340).  Closure procedure pointers
   ! (src/Closures.f90:67-151, src/Limiters.f90:68-75) become enums, selected from the strings the settings
   ! modules store beside the pointers (src/Parameters.f90:175-420, src/SolverSettings.f90:102-126).
   subroutine MakeParams(RunParams, grid, p)
      type(RunSet), intent(in) :: RunParams
      type(GridType), intent(in) :: grid
      type(kgpu_params), intent(out) :: p
      integer :: kk, n, nMax

      p%struct_bytes = int(c_sizeof(p), c_int32_t)

      ! Domain
      p%nXpertile = int(RunParams%nXpertile, c_int32_t)
      p%nYpertile = int(RunParams%nYpertile, c_int32_t)
      p%nXtiles = int(RunParams%nXtiles, c_int32_t)
      p%nYtiles = int(RunParams%nYtiles, c_int32_t)
      p%isOneD = LogicalToC(RunParams%isOneD)
      p%deltaX = real(RunParams%deltaX, c_double)
      p%deltaY = real(RunParams%deltaY, c_double)
      p%xSize = real(RunParams%xSize, c_double)
      p%ySize = real(RunParams%ySize, c_double)
      select case (RunParams%bcs%s)
         case ('halt')
            p%bcs = KGPU_BC_HALT
         case ('periodic')
            p%bcs = KGPU_BC_PERIODIC
         case ('dirichlet')
            p%bcs = KGPU_BC_DIRICHLET
         case ('sponge')
            p%bcs = KGPU_BC_SPONGE
         case default
            call FatalErrorMessage("kestrel_gpu: unknown boundary condition " // RunParams%bcs%s)
      end select
      p%pad0 = 0_c_int32_t
      p%bcsHnval = real(RunParams%bcsHnval, c_double)
      p%bcsuval = real(RunParams%bcsuval, c_double)
      p%bcsvval = real(RunParams%bcsvval, c_double)
      p%bcspsival = real(RunParams%bcspsival, c_double)

      ! Parameters
      p%geometric_factors = LogicalToC(RunParams%geometric_factors)
      p%MorphodynamicsOn = LogicalToC(RunParams%MorphodynamicsOn)
      p%g = real(RunParams%g, c_double)
      p%rhow = real(RunParams%rhow, c_double)
      p%rhos = real(RunParams%rhos, c_double)
      p%gred = real(RunParams%gred, c_double)
      p%ChezyCo = real(RunParams%ChezyCo, c_double)
      p%ManningCo = real(RunParams%ManningCo, c_double)
      p%CoulombCo = real(RunParams%CoulombCo, c_double)
      p%PouliquenMinSlope = real(RunParams%PouliquenMinSlope, c_double)
      p%PouliquenMaxSlope = real(RunParams%PouliquenMaxSlope, c_double)
      p%PouliquenIntermediateSlope = real(RunParams%PouliquenIntermediateSlope, c_double)
      p%PouliquenBeta = real(RunParams%PouliquenBeta, c_double)
      p%Edwards2019betastar = real(RunParams%Edwards2019betastar, c_double)
      p%Edwards2019kappa = real(RunParams%Edwards2019kappa, c_double)
      p%Edwards2019Gamma = real(RunParams%Edwards2019Gamma, c_double)
      p%VoellmySwitchRate = real(RunParams%VoellmySwitchRate, c_double)
      p%VoellmySwitchValue = real(RunParams%VoellmySwitchValue, c_double)
      p%EroRate = real(RunParams%EroRate, c_double)
      p%EroRateGranular = real(RunParams%EroRateGranular, c_double)
      p%CriticalShields = real(RunParams%CriticalShields, c_double)
      p%EroDepth = real(RunParams%EroDepth, c_double)
      p%EroCriticalHeight = real(RunParams%EroCriticalHeight, c_double)
      p%BedPorosity = real(RunParams%BedPorosity, c_double)
      p%maxPack = real(RunParams%maxPack, c_double)
      p%SolidDiameter = real(RunParams%SolidDiameter, c_double)
      p%ws0 = real(RunParams%ws0, c_double)
      p%nsettling = real(RunParams%nsettling, c_double)
      p%EddyViscosity = real(RunParams%EddyViscosity, c_double)

      ! Solver
      p%heightThreshold = real(RunParams%heightThreshold, c_double)
      p%cfl = real(RunParams%cfl, c_double)
      p%diffusiveTimeScale = real(RunParams%diffusiveTimeScale, c_double)
      p%maxdt = real(RunParams%maxdt, c_double)
      p%tstart = real(grid%t, c_double)      ! grid%t on entry to the first IntegrateTo (restart: src/Restart.f90:79)
      p%TileBuffer = int(RunParams%TileBuffer, c_int32_t)
      p%SpongeLayer = LogicalToC(RunParams%SpongeLayer)
      p%SpongeStrength = real(RunParams%SpongeStrength, c_double)

      ! closures
      select case (RunParams%limiter%s)
         case ('MinMod1')
            p%limiter = KGPU_LIM_MINMOD1
         case ('MinMod2')
            p%limiter = KGPU_LIM_MINMOD2
         case ('None')
            p%limiter = KGPU_LIM_NONE
         case ('van Albada')
            p%limiter = KGPU_LIM_VANALBADA
         case ('Weno')
            p%limiter = KGPU_LIM_WENO
         case default
            call FatalErrorMessage("kestrel_gpu: unknown limiter " // RunParams%limiter%s)
      end select
      select case (RunParams%DragChoice%s)
         case ('Chezy')
            p%drag = KGPU_DRAG_CHEZY
         case ('Coulomb')
            p%drag = KGPU_DRAG_COULOMB
         case ('Voellmy')
            p%drag = KGPU_DRAG_VOELLMY
         case ('Pouliquen')
            p%drag = KGPU_DRAG_POULIQUEN
         case ('Edwards2019')
            p%drag = KGPU_DRAG_EDWARDS2019
         case ('Variable')
            p%drag = KGPU_DRAG_VARIABLE
         case ('Manning')
            p%drag = KGPU_DRAG_MANNING
         case default
            call FatalErrorMessage("kestrel_gpu: unknown drag " // RunParams%DragChoice%s)
      end select
      select case (RunParams%ErosionChoice%s)
         case ('Off')
            p%erosion = KGPU_ERO_OFF
         case ('Simple')
            p%erosion = KGPU_ERO_SIMPLE
         case ('Fluid')
            p%erosion = KGPU_ERO_FLUID
         case ('Granular')
            p%erosion = KGPU_ERO_GRANULAR
         case ('Mixed')
            p%erosion = KGPU_ERO_MIXED
         case default
            call FatalErrorMessage("kestrel_gpu: unknown erosion " // RunParams%ErosionChoice%s)
      end select
      select case (RunParams%DepositionChoice%s)
         case ('None')
            p%deposition = KGPU_DEP_NONE
         case ('Simple')
            p%deposition = KGPU_DEP_SIMPLE
         case ('Spearman Manning')
            p%deposition = KGPU_DEP_SPEARMAN_MANNING
         case default
            call FatalErrorMessage("kestrel_gpu: unknown deposition " // RunParams%DepositionChoice%s)
      end select
      select case (RunParams%ErosionTransition%s)
         case ('smooth')
            p%erosion_transition = KGPU_EROTRANS_SMOOTH
         case ('step')
            p%erosion_transition = KGPU_EROTRANS_STEP
         case ('off')
            p%erosion_transition = KGPU_EROTRANS_OFF
         case default
            call FatalErrorMessage("kestrel_gpu: unknown erosion transition " // RunParams%ErosionTransition%s)
      end select
      select case (RunParams%MorphoDamp%s)
         case ('None')
            p%morpho_damp = KGPU_DAMP_NONE
         case ('tanh')
            p%morpho_damp = KGPU_DAMP_TANH
         case ('rat3')
            p%morpho_damp = KGPU_DAMP_RAT3
         case default
            call FatalErrorMessage("kestrel_gpu: unknown morphodynamic damping " // RunParams%MorphoDamp%s)
      end select
      select case (RunParams%fswitch%s)
         case ('tanh')
            p%fswitch = KGPU_SWITCH_TANH
         case ('rat3')
            p%fswitch = KGPU_SWITCH_RAT3
         case ('cos')
            p%fswitch = KGPU_SWITCH_COS
         case ('linear')
            p%fswitch = KGPU_SWITCH_LINEAR
         case ('equal')
            p%fswitch = KGPU_SWITCH_EQUAL
         case ('zero')
            p%fswitch = KGPU_SWITCH_ZERO
         case ('one')
            p%fswitch = KGPU_SWITCH_ONE
         case ('step')
            p%fswitch = KGPU_SWITCH_STEP
         case default
            call FatalErrorMessage("kestrel_gpu: unknown switch function " // RunParams%fswitch%s)
      end select

      ! flux sources (type Sources, src/RunSettings.f90:101-109): series copied into rectangular c_double tables
      p%n_sources = 0_c_int32_t
      p%sources = c_null_ptr
      if (RunParams%set_Sources .and. RunParams%nSources > 0) then
         n = RunParams%nSources
         nMax = 1
         do kk = 1, n
            nMax = max(nMax, RunParams%FluxSources(kk)%nFluxSeries)
         end do
         if (allocated(srcTable)) deallocate (srcTable, srcTime, srcFlux, srcPsi)
         allocate (srcTable(n), srcTime(nMax, n), srcFlux(nMax, n), srcPsi(nMax, n))
         srcTime = 0.0_c_double
         srcFlux = 0.0_c_double
         srcPsi = 0.0_c_double
         do kk = 1, n
            call FillSource(RunParams, kk)
         end do
         p%n_sources = int(n, c_int32_t)
         p%sources = c_loc(srcTable)
      end if

      ! topography: the library calls back into GetHeights when it activates a tile
      p%heights = c_funloc(HeightsCallback)
      p%heights_ctx = c_null_ptr

      ! library options
      p%device = -1_c_int32_t        ! current CUDA device
      p%arithmetic = 0_c_int32_t     ! faithful: reference operation order, no FMA contraction
      p%comm_rank = 0_c_int32_t
      p%comm_size = 1_c_int32_t
      p%comm_px = 1_c_int32_t
      p%comm_py = 1_c_int32_t

   end subroutine MakeParams

   subroutine FillSource(RunParams, kk)
      type(RunSet), intent(in) :: RunParams
      integer, intent(in) :: kk
      integer :: m

      m = RunParams%FluxSources(kk)%nFluxSeries
      srcTime(1:m, kk) = real(RunParams%FluxSources(kk)%time(1:m), c_double)
      srcFlux(1:m, kk) = real(RunParams%FluxSources(kk)%flux(1:m), c_double)
      srcPsi(1:m, kk) = real(RunParams%FluxSources(kk)%psi(1:m), c_double)
      srcTable(kk)%x = real(RunParams%FluxSources(kk)%x, c_double)
      srcTable(kk)%y = real(RunParams%FluxSources(kk)%y, c_double)
      srcTable(kk)%radius = real(RunParams%FluxSources(kk)%radius, c_double)
      srcTable(kk)%num_cells_in_src = int(RunParams%FluxSources(kk)%NumCellsInSrc, c_int32_t)
      srcTable(kk)%n_series = int(m, c_int32_t)
      srcTable(kk)%time = c_loc(srcTime(1, kk))
      srcTable(kk)%flux = c_loc(srcFlux(1, kk))
      srcTable(kk)%psi = c_loc(srcPsi(1, kk))
   end subroutine FillSource

   pure function LogicalToC(flag) result(v)
      logical, intent(in) :: flag
      integer(c_int32_t) :: v
      v = 0_c_int32_t
      if (flag) v = 1_c_int32_t
   end function LogicalToC

   ! ------------------------------------------------------------------------------------------------
   ! int (*kgpu_heights_fn)(void *ctx, int32_t tile_id, double *b0_vertices): b0 at the (nX+1) x (nY+1) vertices of
   ! tile `tile_id`, i fastest.  The library calls it for tiles it activates and for their ghost ring
   ! (src/UpdateTiles.f90:171, 466).  The host-side arrays of such a tile may not exist yet; they are then allocated
   ! and filled through the host's own AllocateTile -> GetHeights (src/dem.f90:360), see AllocateAndLoadHeights.
   function HeightsCallback(ctx, tile_id, b0_vertices) bind(C) result(rc)
      type(c_ptr), value :: ctx
      integer(c_int32_t), value :: tile_id
      real(c_double), intent(out) :: b0_vertices(*)
      integer(c_int) :: rc
      integer :: k, n, i, j, nXv, nYv

      rc = 1_c_int
      if (.not. associated(cbGrid) .or. .not. associated(cbRunParams)) return
      k = int(tile_id)
      if (k < 1 .or. k > cbGrid%nTiles) return
      if (.not. allocated(cbGrid%tileContainer(k)%b0)) call AllocateAndLoadHeights(cbRunParams, cbGrid, k)
      nXv = cbRunParams%nXpertile + 1
      nYv = cbRunParams%nYpertile + 1
      if (cbRunParams%isOneD) nYv = 1
      n = 0
      do j = 1, nYv
         do i = 1, nXv
            n = n + 1
            b0_vertices(n) = real(cbGrid%tileContainer(k)%b0(i, j), c_double)
         end do
      end do
      rc = 0_c_int
   end function HeightsCallback

   ! Heights of a tile the host has not allocated yet (a ghost tile, or a tile the library is switching on inside
   ! kgpu_integrate_to).  AllocateTile (src/UpdateTiles.f90:120-219, public) allocates the tile's arrays, sets its
   ! neighbour table and coordinates and calls GetHeights -- the reference's own way to obtain a tile's heights; every
   ! allocation in it is guarded by `if (.not. allocated(...))`, so the AddTile that follows for tiles that do become
   ! active (IntegrateToGPU) finds them in place.  AllocateTile also flags the tile as on; that is undone here, the
   ! active list is the library's to decide.
   subroutine AllocateAndLoadHeights(RunParams, grid, k)
      type(RunSet), intent(in) :: RunParams
      type(GridType), target, intent(inout) :: grid
      integer, intent(in) :: k

      call AllocateTile(RunParams, grid, k)
      grid%tileContainer(k)%TileOn = .false.
   end subroutine AllocateAndLoadHeights

end module kestrel_gpu_host
