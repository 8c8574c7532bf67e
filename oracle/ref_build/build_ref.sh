#!/bin/bash
# build_ref.sh -- stubbed gfortran build of the reference for Topog Type = Function inputs (test infrastructure).
#
# usage: oracle/ref_build/build_ref.sh [contract]
#   no argument : oracle/_ref/kestrel_ref      -O2 -ffp-contract=off (every operation individually rounded: the oracle's arithmetic)
#   contract    : oracle/_ref/kestrel_ref_fma  -O2 -march=native, the reference's own release flags (src/Makefile.am:4-8, quirk Q12)
# Sources are compiled where they lie ($KESTREL_SRC, default /root/reference/src); objects, .mod files, the patched
# copy of TimeStepper.f90 and the binaries go to oracle/_ref/ only.  Needs gfortran, gcc and g++; no GDAL, PROJ,
# NetCDF or autotools.  Flags: src/Makefile.am:25; file order: src/Makefile.am:16-23.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${KESTREL_SRC:-/root/reference/src}"
OUT="$HERE/../_ref"
MODE="${1:-parity}"
command -v gfortran > /dev/null || { echo "build_ref.sh: gfortran not found (the recipe cannot run in this image)" >&2; exit 3; }
[ -f "$SRC/TimeStepper.f90" ] || { echo "build_ref.sh: reference sources not found under $SRC" >&2; exit 4; }
if [ "$MODE" = contract ]; then OPT="-O2 -falign-loops=16 -march=native -mtune=native"; BIN=kestrel_ref_fma; OBJ="$OUT/obj_fma"
else OPT="-O2 -ffp-contract=off"; BIN=kestrel_ref; OBJ="$OUT/obj"; fi
mkdir -p "$OBJ"
FC="gfortran -cpp -ffree-line-length-0 -fno-range-check $OPT -J$OBJ -I$OBJ"
# the patched copy of Run (raw dump after every output)
sed -f "$HERE/patch_timestepper.sed" "$SRC/TimeStepper.f90" > "$OBJ/TimeStepper_dump.f90"
grep -q "call DumpRawState" "$OBJ/TimeStepper_dump.f90" || { echo "build_ref.sh: the TimeStepper.f90 patch did not apply" >&2; exit 5; }
gcc -O2 -c "$HERE/gdal_proj_stubs.c" -o "$OBJ/gdal_proj_stubs.o"
g++ -O2 -c "$SRC/cversion.cpp" -o "$OBJ/cversion.o"
ORDER_A="SetPrecision Messages varStringClass utilities Interp2d utm RunSettings GeoTiffRead DomainSettings InitConds Closures TopogFuncs
         Parameters Limiters SolverSettings OutputSettings TopogSettings Input Grid Equations HydraulicRHS MorphodynamicRHS Redistribute
         dem UpdateTiles SetSources NetCDFUtils Output Restart"
OBJS="$OBJ/gdal_proj_stubs.o $OBJ/cversion.o"
for f in $ORDER_A; do $FC -c "$SRC/$f.f90" -o "$OBJ/$f.o"; OBJS="$OBJS $OBJ/$f.o"; done
$FC -c "$HERE/raw_dump.f90" -o "$OBJ/raw_dump.o"
$FC -c "$OBJ/TimeStepper_dump.f90" -o "$OBJ/TimeStepper.o"
$FC -c "$SRC/version.f90" -o "$OBJ/version.o"
$FC -c "$SRC/main.f90" -o "$OBJ/main.o"
gfortran $OPT -o "$OUT/$BIN" $OBJS "$OBJ/raw_dump.o" "$OBJ/TimeStepper.o" "$OBJ/version.o" "$OBJ/main.o" -lstdc++
echo "$OUT/$BIN"
