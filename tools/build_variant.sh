#!/bin/bash
# Builds build/variants/<name>/libtiledmm_b200.so = the library at HEAD with ONE kernel source taken from another revision, for A/B runs of the same
# tool binary on the GPU box (LD_LIBRARY_PATH=build/variants/<name> ./build/devtest ...; the binaries carry a RUNPATH, so the variable wins).
#   tools/build_variant.sh f64old 9a0f944 gemm_f64.cu        (run here, before the gpurun call: the variant travels with the snapshot)
set -e
NAME=$1; REV=$2; FILE=$3
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/tiled-mm_b200/csrc
make -C "$SRC" -j8 > /dev/null
mkdir -p "$ROOT/build/variants/$NAME"
git -C "$ROOT" show "$REV:tiled-mm_b200/csrc/$FILE" > "$SRC/_variant_$FILE"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -I"$ROOT/include" -DTMM_HAVE_ZGEMM_DMMA \
     -c "$SRC/_variant_$FILE" -o "$ROOT/build/variants/$NAME/variant.o"
rm -f "$SRC/_variant_$FILE"
OBJS=$(ls "$SRC"/obj/*.o | grep -v "/${FILE%.*}.o$")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$ROOT/build/variants/$NAME/libtiledmm_b200.so" $OBJS "$ROOT/build/variants/$NAME/variant.o" -ldl -lrt
echo "built build/variants/$NAME/libtiledmm_b200.so ($FILE from $REV)"
