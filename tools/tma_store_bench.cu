// Micro-benchmark: how fast can one SM push small TMA stores (the GEMM epilogue pattern)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/tma_store_bench tools/tma_store_bench.cu -lcuda
// Every CTA (one per SM, `nwarps` warps) loops: write a box to smem, fence.proxy.async, TMA-store it, commit, wait<pend>.
// Variants: box rows (32 / 128), row bytes (64 / 128), commit every `cevery` stores.  Prints cycles per store per SM and GB/s.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int PEND>
__device__ __forceinline__ void wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PEND) : "memory"); }

__global__ void __launch_bounds__(256, 1)
store_kernel(const __grid_constant__ CUtensorMap tm, int iters, int box_rows, int row_bytes, int cevery, int pend, int rows_per_cta,
             long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int box_bytes = box_rows * row_bytes;
  const int R = 4;
  uint8_t* ring = smem + (size_t)warp * R * box_bytes;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint8_t* slot = ring + (i % R) * box_bytes;
    if (lane == 0) {
      switch (pend) {
        case 0: wait_read<0>(); break;
        case 1: wait_read<1>(); break;
        case 2: wait_read<2>(); break;
        default: wait_read<3>(); break;
      }
    }
    __syncwarp();
    // fill: lane writes 16-byte pieces
    for (int o = lane * 16; o < box_bytes; o += 32 * 16) *reinterpret_cast<uint4*>(slot + o) = make_uint4(i, lane, warp, o);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      // distinct destination per (cta, warp, iteration): rows advance, column block = warp
      const int row = blockIdx.x * rows_per_cta + (i * box_rows) % rows_per_cta;
      const int col = warp * (row_bytes / 4);
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                       reinterpret_cast<uint64_t>(&tm)),
                   "r"(smem_u32(slot)), "r"(col), "r"(row)
                   : "memory");
      if ((i + 1) % cevery == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles_out[blockIdx.x] = clock64() - t0;
  (void)nwarps;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const int rows_per_cta = 8192;
  const size_t cols = 1024;  // floats per row (4 KB rows)
  float* buf;
  cudaMalloc(&buf, (size_t)nsm * rows_per_cta * cols * 4);
  long long* cyc;
  cudaMalloc(&cyc, nsm * sizeof(long long));
  cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("%6s %6s %6s %6s %6s | %12s %10s\n", "warps", "rows", "rowB", "cevery", "pend", "cyc/store/SM", "GB/s");
  const int cfgs[][5] = {
      // warps, box_rows, row_bytes, commit_every, pending
      {4, 32, 128, 1, 1}, {4, 32, 128, 1, 3}, {4, 32, 128, 2, 1}, {4, 32, 128, 4, 0}, {8, 32, 128, 1, 3}, {8, 32, 64, 1, 3},
      {4, 32, 64, 1, 3},  {1, 32, 128, 1, 3}, {1, 128, 128, 1, 3}, {2, 128, 128, 1, 3}, {4, 128, 128, 1, 3}, {1, 128, 64, 1, 3},
      {1, 256, 128, 1, 3}, {2, 64, 128, 1, 3}, {4, 16, 128, 1, 3}, {4, 8, 128, 1, 3},
  };
  for (auto& c : cfgs) {
    const int warps = c[0], box_rows = c[1], row_bytes = c[2], cevery = c[3], pend = c[4];
    CUtensorMap tm;
    cuuint64_t dims[2] = {cols, (cuuint64_t)nsm * rows_per_cta};
    cuuint64_t strides[1] = {cols * 4};
    cuuint32_t box[2] = {(cuuint32_t)(row_bytes / 4), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
    const int iters = 2000;
    const int smem = warps * 4 * box_rows * row_bytes + 1024;
    if (smem > 200 * 1024) { printf("skip (smem)\n"); continue; }
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      store_kernel<<<nsm, warps * 32, smem>>>(tm, iters, box_rows, row_bytes, cevery, pend, rows_per_cta, cyc);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 1) {
        long long h[256];
        cudaMemcpy(h, cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < nsm; ++i) avg += h[i];
        avg /= nsm;
        const double stores_per_sm = (double)iters * warps;
        const double bytes = (double)nsm * stores_per_sm * box_rows * row_bytes;
        printf("%6d %6d %6d %6d %6d | %12.1f %10.1f\n", warps, box_rows, row_bytes, cevery, pend, avg / stores_per_sm,
               bytes / (ms * 1e-3) / 1e9);
      }
    }
  }
  return 0;
}
