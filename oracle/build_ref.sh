#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds oracle/_ref/: the reference's own C++ solver stack
# (PoissonOp / MGSolver / LevelHybridSolver / Chombo containers), compiled UNMODIFIED from the
# sources where they lie under /root/reference (serial, no Python, no MPI), linked against
# oracle/fort_leaves.cpp (our C++ restatement of the Chombo-Fortran leaf kernels + dgtsv; this
# image has no gfortran / LAPACK) and oracle/ref_driver.cpp (a small driver standing in for
# AMRNSLevel::projectCorrect).  Outputs go ONLY into oracle/_ref/ (git-ignored).
#
# Usage: oracle/build_ref.sh [2|3]       (space dimension; default 3)
set -euo pipefail
DIM="${1:-3}"
REF="${SOMAR_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref/d$DIM"
GEN="$OUT/gen"
OBJ="$OUT/obj"
JOBS="${JOBS:-$(nproc)}"
if [ ! -d "$REF/src" ]; then echo "reference tree not found at $REF -- nothing to build"; exit 0; fi
mkdir -p "$GEN" "$OBJ"

# 1. ChF -> _F.H prototypes with the vendored perl preprocessor (site_scons/site_init.py:58-66).
for f in $(find "$REF/src" -name '*.ChF'); do
  b="$(basename "$f" .ChF)"
  if [ ! -s "$GEN/${b}_F.H" ]; then
    (cd "$GEN" && perl -I "$REF/compileUtils/chfpp" "$REF/compileUtils/chfpp/uber.pl" -f "$f" -p /dev/null -c "$GEN/${b}_F.H" -D"$DIM")
  fi
done

# 2. include path = every directory under src (PyGlue excluded), as site_init.py:285-296 does.
INC="-I$GEN -I$HERE"
for d in $(find "$REF/src" -type d | grep -v PyGlue); do INC="$INC -I$d"; done
DEFS="-DCH_SPACEDIM=$DIM -DCH_Linux -DCH_USE_64 -DCH_USE_DOUBLE -DCH_FORT_UNDERSCORE -DCH_LANG_CC -DCH_NTIMER -DNDEBUG -DCH_USE_COMPLEX"
CXXFLAGS="-std=c++17 -O3 -funroll-loops -ffp-contract=off -fPIC -w $DEFS"

# 3. compile the reference sources (object names flattened).
# Grade5 (the Navier-Stokes physics) is out of scope: only the parameter singleton the solvers read.
SRCS=$(find "$REF/src" -name '*.cpp' | grep -v PyGlue | grep -v Grade5_SOMAR;
       ls "$REF"/src/Grade5_SOMAR/{ProblemContext,ProjectorParameters,RHSParameters}.cpp)
echo "$SRCS" | xargs -P "$JOBS" -I{} bash -c '
  s="{}"; o="'"$OBJ"'/$(basename "$s" .cpp).o"
  if [ ! -s "$o" ] || [ "$s" -nt "$o" ]; then g++ '"$CXXFLAGS $INC"' -c "$s" -o "$o" || { echo "FAILED: $s"; exit 255; }; fi'

# 4. our pieces.
g++ $CXXFLAGS $INC -c "$HERE/fort_leaves.cpp" -o "$OBJ/_fort_leaves.o"
gcc -O3 -ffp-contract=off -fPIC -c "$HERE/dgtsv.c" -o "$OBJ/_dgtsv.o"
g++ $CXXFLAGS $INC -c "$HERE/ref_driver.cpp" -o "$OBJ/_ref_driver.o"

# 5. link; any Fortran leaf the driver never reaches becomes an aborting stub.
OBJS=$(for s in $SRCS; do echo "$OBJ/$(basename "$s" .cpp).o"; done)
link() { g++ -o "$OUT/somar_ref" $OBJS "$OBJ/_fort_leaves.o" "$OBJ/_dgtsv.o" "$OBJ/_ref_driver.o" $1 -lpthread 2>&1; }
if ! link "" > "$OUT/link1.log"; then
  grep -o "undefined reference to \`[A-Za-z0-9_]*'" "$OUT/link1.log" | sed "s/.*\`//; s/'//" | sort -u > "$OUT/undefined.txt"
  { echo '#include <stdio.h>'; echo '#include <stdlib.h>';
    while read -r s; do echo "void $s(void){fprintf(stderr,\"oracle/_ref: unreached Fortran leaf $s was called\\n\");abort();}"; done < "$OUT/undefined.txt"; } > "$GEN/fort_stubs.c"
  gcc -fPIC -c "$GEN/fort_stubs.c" -o "$OBJ/_fort_stubs.o"
  link "$OBJ/_fort_stubs.o" > "$OUT/link2.log" || { cat "$OUT/link2.log" | head -50; exit 1; }
fi
echo "built $OUT/somar_ref"

# 6. the same driver with SOMAR's `using PoissonOp / LevelProjSolver` lines swapped for the B200 drop-ins of integration/
#    (compiled against the reference's headers, linked to libsomar_b200.so): somar_ref_b200.  Needs the product library.
ROOT="$(cd "$HERE/.." && pwd)"
LIBDIR="$ROOT/somar_b200/lib"
if [ -f "$LIBDIR/libsomar_b200.so" ]; then
  SHIMINC="$INC -I$ROOT/include -I$ROOT/integration"
  g++ $CXXFLAGS $SHIMINC -c "$ROOT/integration/B200PoissonOp.cpp" -o "$OBJ/_b200_poissonop.o"
  g++ $CXXFLAGS $SHIMINC -c "$ROOT/integration/B200LevelHybridSolver.cpp" -o "$OBJ/_b200_hybrid.o"
  g++ $CXXFLAGS $SHIMINC -DSB_SHIM -c "$HERE/ref_driver.cpp" -o "$OBJ/_ref_driver_b200.o"
  STUBS=""; [ -f "$OBJ/_fort_stubs.o" ] && STUBS="$OBJ/_fort_stubs.o"
  g++ -o "$OUT/somar_ref_b200" $OBJS "$OBJ/_fort_leaves.o" "$OBJ/_dgtsv.o" "$OBJ/_ref_driver_b200.o" "$OBJ/_b200_poissonop.o" \
      "$OBJ/_b200_hybrid.o" $STUBS -L"$LIBDIR" -lsomar_b200 -Wl,-rpath,'$ORIGIN/../../../somar_b200/lib' -lpthread > "$OUT/link_b200.log" 2>&1 \
      || { head -40 "$OUT/link_b200.log"; exit 1; }
  echo "built $OUT/somar_ref_b200"
else
  echo "libsomar_b200.so not built yet: skipping somar_ref_b200"
fi
