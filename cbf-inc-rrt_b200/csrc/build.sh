#!/bin/bash
# Builds libcbfrrt.so for sm_100a, in-tree (cbf-inc-rrt_b200/cbfrrt/_lib/).  -fmad=false: the deterministic-math
# contract forbids implicit contraction; every FMA in the sources is an explicit fmaf().
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../cbfrrt/_lib"
mkdir -p "$OUT"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -diag-suppress 20091"
"$NVCC" $FLAGS -dc -o "$OUT/kernels.o" "$HERE/kernels.cu" ${PTXAS_V:+-Xptxas -v} &
"$NVCC" $FLAGS -dc -o "$OUT/host_api.o" "$HERE/host_api.cu" &
wait
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libcbfrrt.so" "$OUT/kernels.o" "$OUT/host_api.o" -lcudart
rm -f "$OUT/kernels.o" "$OUT/host_api.o"
echo "built $OUT/libcbfrrt.so"
