#!/usr/bin/env bash
# Builds everything that needs the reference sources, from where they lie (default /root/reference), into oracle/_ref/
# (git-ignored, shipped to the GPU box as built files).  The reference's own CMake is NOT used: it hard-requires MPI
# (CMakeLists.txt:171) which this image lacks; the few sources on the path compile directly.
#   1. libmurbref*.so   reference Bodies + cpu+naive/optim/simd/omp behind oracle/ref_wrap.cpp   (oracle pin, CPU baseline)
#   2. murb_b200        reference main.cpp patched with the gpu+b200 branch + reference CPU and CUDA variants + the glue
#   3. murb-test-b200   Catch2 runner: the reference's OWN test bodies (test_SimulationNBody.cpp, test_CUDABodies.cpp) with
#                       only the class under test re-targeted at gpu+b200 by oracle/patch_test.py, the reference's
#                       test_SimulationHistory.cu unchanged, and tests/catch2/test_B200.cpp (what the reference lacks)
set -u
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")/.." && pwd)
OUT=$HERE/oracle/_ref
PKG=$HERE/nbody-eurohpc_b200
[ -d "$REF/src" ] || { echo "build_ref: no reference at $REF (nothing to do)"; exit 0; }
mkdir -p "$OUT"
INC="-I$REF/src/common -I$REF/src/murb -I$REF/src/murb/implem -I$REF/lib/MIPP/src"
CORE="$REF/src/common/core/Bodies.cpp $REF/src/common/core/BodiesAllocator.cpp $REF/src/common/core/SimulationNBodyInterface.cpp $REF/src/common/utils/Perf.cpp"
CPUIMPL="$REF/src/murb/implem/SimulationNBodyNaive.cpp $REF/src/murb/implem/SimulationNBodyOptim.cpp $REF/src/murb/implem/SimulationNBodySIMD.cpp $REF/src/murb/implem/SimulationNBodyOpenMP.cpp"
REFFLAGS="-std=c++20 -O3 -ffast-math -fopenmp"   # the reference's flags (CMakeLists.txt:128-131), OpenMP on
rc=0
# rebuild an object only when its source (or the glue / C-ABI headers) is newer
stale() { [ ! -e "$1" ] || [ "$2" -nt "$1" ] || [ "$PKG/glue/SimulationNBodyB200.hpp" -nt "$1" ] || [ "$HERE/include/b200nb.h" -nt "$1" ]; }

echo "== 1. libmurbref (as shipped: no -march => SSE2 MIPP) and ISA-specific builds"
for v in "" "_v3:-march=x86-64-v3" "_v4:-march=x86-64-v4"; do
  suffix=${v%%:*}; march=${v#*:}; [ "$v" = "" ] && march=""
  so="$OUT/libmurbref$suffix.so"
  if [ ! -e "$so" ] || [ "$HERE/oracle/ref_wrap.cpp" -nt "$so" ]; then
    g++ $REFFLAGS $march -fPIC -shared $INC "$HERE/oracle/ref_wrap.cpp" $CORE $CPUIMPL -o "$so" || rc=1
  fi
done

echo "== 2. murb_b200 (patched CLI)"
HAVE_MPI=0; command -v mpicxx >/dev/null 2>&1 && HAVE_MPI=1
python3 "$HERE/oracle/patch_main.py" "$REF/src/murb/main.cpp" "$OUT/main_b200.cpp" $HAVE_MPI || rc=1
OBJ=$OUT/obj; mkdir -p "$OBJ"
CUDAINC="-I/usr/local/cuda/include"
NVFLAGS="-std=c++17 -O3 --use_fast_math -gencode arch=compute_100a,code=sm_100a -DUSE_CUDA $INC"   # CMakeLists.txt:134-141 + arch
objs=""
for f in CUDABodies SimulationHistoryGPU; do
  if stale "$OBJ/$f.o" "$REF/src/common/core/$f.cu"; then nvcc $NVFLAGS -c "$REF/src/common/core/$f.cu" -o "$OBJ/$f.o" 2>/dev/null || rc=1; fi; objs="$objs $OBJ/$f.o"
done
for f in SimulationNBodyCUDATile SimulationNBodyCUDATileFullDevice SimulationNBodyCUDATileFullDevice200k SimulationNBodyCUDAPropertyTracking SimulationNBodyCUDALeapfrog SimulationNBodyHetero; do
  if stale "$OBJ/$f.o" "$REF/src/murb/implem/$f.cu"; then nvcc $NVFLAGS -Xcompiler -fopenmp -c "$REF/src/murb/implem/$f.cu" -o "$OBJ/$f.o" 2>/dev/null || rc=1; fi; objs="$objs $OBJ/$f.o"
done
HOSTSRC="$CORE $REF/src/common/ogl/SpheresVisuNo.cpp $REF/src/common/core/SimulationHistory.cpp $REF/src/common/core/HistoryTrackingInterface.cpp $REF/src/common/utils/ArgumentsReader.cpp $CPUIMPL $REF/src/murb/implem/SimulationNBodyNop.cpp"
[ $HAVE_MPI = 1 ] && HOSTSRC="$HOSTSRC $REF/src/murb/implem/SimulationNBodyMultiNode.cpp"
HOSTOBJS=""
for f in $HOSTSRC "$PKG/glue/SimulationNBodyB200.cpp"; do
  o="$OBJ/$(basename "${f%.cpp}").host.o"
  if stale "$o" "$f"; then g++ $REFFLAGS -DUSE_CUDA $INC $CUDAINC -I"$PKG/glue" -I"$HERE/include" -c "$f" -o "$o" || rc=1; fi
  HOSTOBJS="$HOSTOBJS $o"
done
g++ $REFFLAGS -DUSE_CUDA $INC $CUDAINC -I"$PKG/glue" -I"$HERE/include" -c "$OUT/main_b200.cpp" -o "$OBJ/main_b200.o" || rc=1
LINK="-L$PKG/b200nb -lb200nb -Wl,-rpath,\$ORIGIN/../../nbody-eurohpc_b200/b200nb -L/usr/local/cuda/lib64 -lcudart -fopenmp"
g++ "$OBJ/main_b200.o" $HOSTOBJS $objs $LINK -o "$OUT/murb_b200" || rc=1

echo "== 3. murb-test-b200 (Catch2)"
TESTINC="$INC $CUDAINC -I$PKG/glue -I$HERE/include -I$REF/lib/Catch2/include"
python3 "$HERE/oracle/patch_test.py" "$REF" "$OUT" || rc=1
g++ $REFFLAGS -DUSE_CUDA $TESTINC -c "$HERE/tests/catch2/test_B200.cpp" -o "$OBJ/test_B200.o" || rc=1
for t in test_SimulationNBody_b200 test_CUDABodies_b200; do   # the reference's test bodies, class under test re-targeted
  g++ $REFFLAGS -DUSE_CUDA $TESTINC -c "$OUT/$t.cpp" -o "$OBJ/$t.o" || rc=1
done
if stale "$OBJ/test_SimulationHistory.o" "$REF/src/test/implem/test_SimulationHistory.cu"; then   # unchanged
  nvcc $NVFLAGS -I"$REF/lib/Catch2/include" -c "$REF/src/test/implem/test_SimulationHistory.cu" -o "$OBJ/test_SimulationHistory.o" 2>/dev/null || rc=1
fi
g++ $REFFLAGS -I"$REF/lib/Catch2/include" -c "$REF/src/test/main.cpp" -o "$OBJ/test_main.o" || rc=1
TESTOBJS=""
for o in $HOSTOBJS; do case "$o" in *ArgumentsReader*|*Nop.host.o|*SpheresVisuNo*) ;; *) TESTOBJS="$TESTOBJS $o";; esac; done
g++ "$OBJ/test_main.o" "$OBJ/test_B200.o" "$OBJ/test_SimulationNBody_b200.o" "$OBJ/test_CUDABodies_b200.o" \
    "$OBJ/test_SimulationHistory.o" $TESTOBJS $objs $LINK -o "$OUT/murb-test-b200" || rc=1
ls -la "$OUT" | grep -v obj
[ $rc = 0 ] && touch "$OUT/.stamp"
exit $rc
