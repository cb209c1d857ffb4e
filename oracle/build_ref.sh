#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Compiles the reference's own CPU processor from the sources where they
# lie under /root/reference into oracle/_ref/libac_ref.so (git-ignored, travels with gpurun).
# The reference's CMake build cannot run offline (every dependency is FetchContent), so this is
# the manual recipe of exactly the hot-path translation units; flags mirror
# core/CMakeLists.txt:51-66,193-211,292-326.  No reference source is copied into the repo.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
R="${AC_REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
GEN="$OUT/gen"
if [ ! -d "$R/core/src" ]; then echo "reference tree not found at $R (nothing to do)"; exit 0; fi
mkdir -p "$GEN/AC/Core/Internal/Model/Param" "$GEN/src" "$OUT/obj"

# stand-in 1: the CMake-generated export header (core/CMakeLists.txt:356-360)
printf '#pragma once\n#define AC_CORE_EXPORT\n#define AC_CORE_NO_EXPORT\n' > "$GEN/ACCoreExport.hpp"
# stand-in 2: half.hpp (Christian Rau's half 2.2, normally fetched) -- any copy on this image
HALF_DIR=""
for d in /opt/prime-rl/.venv/lib/python3.12/site-packages/tilelang/src/tl_templates/cpp; do
  [ -f "$d/half.hpp" ] && HALF_DIR="$d"
done
HALF_DEF=""; HALF_INC=""
if [ -n "$HALF_DIR" ]; then HALF_DEF="-DAC_CORE_WITH_HALF"; HALF_INC="-I$HALF_DIR"; fi
# stand-in 3: ARNet.p (missing blob) -> seeded synthetic arrays shared with the product
gcc -O1 "$HERE/gen_arnet_standin.c" -o "$OUT/gen_arnet_standin"
"$OUT/gen_arnet_standin" > "$GEN/AC/Core/Internal/Model/Param/ARNet.p"
# per-ISA backend translation units (cmake/GenCPUProcessorBackend.cmake:1-14)
gen_backend() { # name suffix header
  printf '#include "AC/Core/Internal/Processor/CPU/%s"\n#define BACKEND_NAME %s\n#define LAYER_SUFFIX %s\n#include "AC/Core/Internal/Processor/CPU/Backend.hpp"\n' "$3" "$1" "$2" > "$GEN/src/$1.cpp"
}
gen_backend Generic generic Generic.hpp
gen_backend SSE sse X86/SSE.hpp
gen_backend AVX avx X86/AVX.hpp
gen_backend FMA fma X86/AVX.hpp
gen_backend AVX512 avx512 X86/AVX512.hpp

INC="-I$GEN -I$R/core/include -I$R/core/internal -I$R/util/misc/include -I$R/util/threads/include -I$R/util/parallel/include $HALF_INC"
DEF="-DAC_CORE_WITH_SSE -DAC_CORE_WITH_AVX -DAC_CORE_WITH_FMA -DAC_CORE_WITH_AVX512 $HALF_DEF \
 -DAC_CORE_HAVE_STD_ALIGNED_ALLOC -DAC_CORE_MALLOC_ALIGN=64 -DAC_CORE_PARAM_ALIGN=64 -DAC_CORE_STRIDE_ALIGN=4 -DAC_DEP_PARALLEL_OPENMP"
CXX="g++ -std=c++17 -O2 -fPIC -fopenmp $INC $DEF -c"
O="$OUT/obj"
pids=()
for f in Alloc Image ImageProcess Model; do $CXX "$R/core/src/$f.cpp" -o "$O/$f.o" & pids+=($!); done
$CXX "$R/core/src/processor/Processor.cpp" -o "$O/Processor.o" & pids+=($!)
$CXX "$R/core/src/processor/cpu/CPUProcessor.cpp" -o "$O/CPUProcessor.o" & pids+=($!)
$CXX "$GEN/src/Generic.cpp" -o "$O/bGeneric.o" & pids+=($!)
$CXX -msse "$GEN/src/SSE.cpp" -o "$O/bSSE.o" & pids+=($!)
$CXX -mavx "$GEN/src/AVX.cpp" -o "$O/bAVX.o" & pids+=($!)
$CXX -mavx -mfma "$GEN/src/FMA.cpp" -o "$O/bFMA.o" & pids+=($!)
$CXX -mavx512f -mfma "$GEN/src/AVX512.cpp" -o "$O/bAVX512.o" & pids+=($!)
$CXX "$HERE/ref_shim.cpp" -o "$O/ref_shim.o" & pids+=($!)
gcc -std=c11 -O2 -fPIC -fopenmp -mfma -ffp-contract=off -c "$HERE/ac_oracle.c" -o "$O/ac_oracle.o" & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
g++ -shared -fopenmp -o "$OUT/libac_ref.so" "$O"/*.o -lm
rm -f "$OUT/gen_arnet_standin"
echo "built $OUT/libac_ref.so"
