#!/bin/bash
# Build libseer_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$HERE/../libseer_b200.so"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I$ROOT/include -I$HERE"
mkdir -p "$HERE/build"
pids=()
for f in gemm_tc attention attention_tc attention_tc80 norm elementwise fp32_path capi; do
  if [ ! -f "$HERE/build/$f.o" ] || [ "$HERE/$f.cu" -nt "$HERE/build/$f.o" ] || [ "$HERE/common.cuh" -nt "$HERE/build/$f.o" ] || [ "$HERE/gemm_epilogue.cuh" -nt "$HERE/build/$f.o" ] || [ "$ROOT/include/seer_b200.h" -nt "$HERE/build/$f.o" ]; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c "$HERE/$f.cu" -o "$HERE/build/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$HERE"/build/*.o -lcudart
echo "built $OUT"
