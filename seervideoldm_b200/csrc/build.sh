#!/bin/bash
# Build libseer_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="${SEER_B200_OUT:-$HERE/../libseer_b200.so}"      # SEER_B200_OUT / SEER_B200_BUILD / EXTRA_NVCC_FLAGS: A/B builds of kernel variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I$ROOT/include -I$HERE"
BUILD="${SEER_B200_BUILD:-$HERE/build}"
mkdir -p "$BUILD"
pids=()
for f in gemm_tc attention attention_tc attention_tc80 norm elementwise fp32_path capi; do
  if [ ! -f "$BUILD/$f.o" ] || [ "$HERE/$f.cu" -nt "$BUILD/$f.o" ] || [ "$HERE/common.cuh" -nt "$BUILD/$f.o" ] || [ "$HERE/gemm_epilogue.cuh" -nt "$BUILD/$f.o" ] || [ "$ROOT/include/seer_b200.h" -nt "$BUILD/$f.o" ]; then
    $NVCC $FLAGS $EXTRA_NVCC_FLAGS ${PTXAS_V:+-Xptxas -v} -c "$HERE/$f.cu" -o "$BUILD/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$BUILD"/*.o -lcudart
echo "built $OUT"
