#!/bin/bash
# Builds libcbfrrt.so for sm_100a, in-tree (cbf-inc-rrt_b200/cbfrrt/_lib/).  -fmad=false: the deterministic-math
# contract forbids implicit contraction; every FMA in the sources is an explicit fmaf().  The translation units
# (C-ABI, host front end, one per kernel family and robot) are compiled in parallel.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../cbfrrt/_lib"
OBJ="$OUT/obj"
mkdir -p "$OBJ"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -diag-suppress 20091 ${EXTRA_NVCC_FLAGS:-}"
pids=()
for src in "$HERE"/abi.cu "$HERE"/host_api.cu "$HERE"/k_*.cu; do
  name=$(basename "$src" .cu)
  "$NVCC" $FLAGS -c -o "$OBJ/$name.o" "$src" ${PTXAS_V:+-Xptxas -v} &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libcbfrrt.so" "$OBJ"/*.o -lcudart
rm -rf "$OBJ"
echo "built $OUT/libcbfrrt.so"
