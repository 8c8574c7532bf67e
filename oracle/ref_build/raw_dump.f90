! raw_dump.f90 -- raw fp64 dump of what the reference's writers read (test infrastructure, oracle/ref_build).
!
! The txt writer keeps 10 significant digits (src/Output.f90:858, quirk Q7), too few to pin an oracle to.
! DumpRawState writes, for output index n, the file <OutDir>/raw_NNNNNN.bin (unformatted stream):
!    int32 nActive, nX, nY, isOneD;  real64 t
!    per active tile, in grid%activeTiles%List order:
!       int32 tileID
!       real64 u(13, nX, nY)            (src/Grid.f90:83)
!       real64 b0(nX+1, nY+1), bt(nX+1, nY+1)
!       real64 Hnmax, umax, emax, dmax, psimax (nX, nY, 2 each);  tfirst(nX, nY)
! It is called from a patched COPY of Run (patch_timestepper.sed) right after each OutputSolutionData.
module raw_dump_module

   use, intrinsic :: iso_fortran_env, only: int32, real64
   use set_precision_module, only: wp
   use grid_module, only: GridType
   use runsettings_module, only: RunSet

   implicit none

   private
   public :: DumpRawState

contains

   subroutine DumpRawState(RunParams, n, grid)
      type(RunSet), intent(in) :: RunParams
      integer, intent(in) :: n
      type(GridType), intent(in) :: grid

      character(len=6) :: idx
      character(len=:), allocatable :: fname
      integer :: unit, tt, k, oneD

      write (idx, '(i6.6)') n
      fname = RunParams%out_path%s // "raw_" // idx // ".bin"
      oneD = 0
      if (RunParams%isOneD) oneD = 1
      open (newunit=unit, file=fname, access='stream', form='unformatted', status='replace')
      write (unit) int(grid%activeTiles%size, int32), int(RunParams%nXpertile, int32), int(RunParams%nYpertile, int32), int(oneD, int32)
      write (unit) real(grid%t, real64)
      do tt = 1, grid%activeTiles%size
         k = grid%activeTiles%List(tt)
         write (unit) int(k, int32)
         write (unit) real(grid%tileContainer(k)%u, real64)
         write (unit) real(grid%tileContainer(k)%b0, real64)
         write (unit) real(grid%tileContainer(k)%bt, real64)
         write (unit) real(grid%tileContainer(k)%Hnmax, real64)
         write (unit) real(grid%tileContainer(k)%umax, real64)
         write (unit) real(grid%tileContainer(k)%emax, real64)
         write (unit) real(grid%tileContainer(k)%dmax, real64)
         write (unit) real(grid%tileContainer(k)%psimax, real64)
         write (unit) real(grid%tileContainer(k)%tfirst, real64)
      end do
      close (unit)
   end subroutine DumpRawState

end module raw_dump_module
