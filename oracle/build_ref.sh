#!/bin/bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference (andreyypopov/integrator2) from the
# sources where they lie under /root/reference into oracle/_ref/ (git-ignored, shipped to the
# GPU box by gpurun).  Nothing from /root/reference is copied into the tracked tree.
#   oracle/_ref/integrator2test3D   the reference CLI (tests/integrator3D/main.cu)
#   oracle/_ref/ref_dump            reference library + oracle/ref_dump.cu (binary dump harness)
#   oracle/_ref/ref_dump_nofma      same, compiled with -fmad=false (the reference's own
#                                   rounding-noise floor: reference vs reference)
#   oracle/_ref/examples/           the reference's example meshes (input data)
# The reference's CMake targets sm_61/70/75 (cmake/functions.cmake:9); we compile the same
# translation units directly with nvcc for sm_100 (-rdc=true as CUDA_SEPARABLE_COMPILATION ON,
# CMakeLists.txt:63), -O2 like its default RelWithDebInfo.
set -e
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
    echo "build_ref: $REF not present, keeping prebuilt files in $OUT" >&2
    exit 0
fi
mkdir -p "$OUT/obj" "$OUT/obj_nofma" "$OUT/examples"
SRCS="src/Mesh3d.cu src/NumericalIntegrator3d.cu src/common/cuda_helper.cu src/common/cuda_math.cu src/evaluators/evaluator3d.cu src/evaluators/evaluatorJ3DK.cu"
NVCC=${NVCC:-nvcc}
FLAGS="-arch=sm_100 -rdc=true -O2 -std=c++17 -w -I$REF/src"
build_variant() { # objdir extra_flags
    local objdir=$1; shift
    local objs=""
    for s in $SRCS; do
        o="$objdir/$(basename ${s%.cu}).o"
        if [ ! -f "$o" ] || [ "$REF/$s" -nt "$o" ]; then
            $NVCC $FLAGS "$@" -dc "$REF/$s" -o "$o" &
        fi
        objs="$objs $o"
    done
    wait
    echo $objs
}
OBJS=$(build_variant "$OUT/obj")
$NVCC $FLAGS -dc "$REF/tests/integrator3D/main.cu" -o "$OUT/obj/main.o"
$NVCC -arch=sm_100 -rdc=true $OBJS "$OUT/obj/main.o" -o "$OUT/integrator2test3D"
$NVCC $FLAGS -dc "$HERE/ref_dump.cu" -o "$OUT/obj/ref_dump.o"
$NVCC -arch=sm_100 -rdc=true $OBJS "$OUT/obj/ref_dump.o" -o "$OUT/ref_dump"
OBJS2=$(build_variant "$OUT/obj_nofma" -fmad=false)
$NVCC $FLAGS -fmad=false -dc "$HERE/ref_dump.cu" -o "$OUT/obj_nofma/ref_dump.o"
$NVCC -arch=sm_100 -rdc=true $OBJS2 "$OUT/obj_nofma/ref_dump.o" -o "$OUT/ref_dump_nofma"
cp -u "$REF"/examples/*.dat "$OUT/examples/" 2>/dev/null || true
for f in 0012e2 13bad 1x1x1_extrafine MeshScreen ellipsoid2000; do cp -u "$REF/examples/$f" "$OUT/examples/" 2>/dev/null || true; done
echo "build_ref: done -> $OUT"
