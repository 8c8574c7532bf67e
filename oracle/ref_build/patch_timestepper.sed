# Applied by build_ref.sh to a COPY of src/TimeStepper.f90 inside oracle/_ref/ (never to the reference tree).
# 1. make the dump module visible in timestepper_module (after the last `use` of the module header, :59)
/^   use utilities_module, only: Int2String/a\   use raw_dump_module, only: DumpRawState
# 2. dump after each OutputSolutionData of Run (src/TimeStepper.f90:87 and :106)
/^ *call OutputSolutionData(RunParams, RunParams%CurrentOut, grid)/a\         call DumpRawState(RunParams, RunParams%CurrentOut, grid)
