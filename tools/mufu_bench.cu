// MUFU.EX2 throughput probe (the roofline of the d = 40 attention kernel): ex2.approx.ftz.f32 per clock per SM as a
// function of resident warps per scheduler, alone and mixed with the companions of the softmax loop
// (FFMA2 scale, FADD2 row sum, F2FP bf16 pack, STS.128 of P).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench tools/mufu_bench.cu && ./mufu_bench
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t f2_pack(float lo, float hi) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(f2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) { f2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) { f2_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }

// MODE 0: MUFU + FADD      1: FFMA + MUFU + FADD      2: FFMA2 + 2 MUFU + FADD2     3: mode 2 + F2FP      4: mode 3 + STS.128 / 8 values
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  __shared__ uint4 sbuf[1024];
  float x[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) x[i] = -0.001f * (threadIdx.x + i);
  float acc = 0.f;
  f2_t sum2[2] = {f2_pack(0.f, 0.f), f2_pack(0.f, 0.f)};
  uint32_t keep = 0;
  const f2_t a2 = f2_pack(0.999f, 0.999f), b2 = f2_pack(-0.0001f, -0.0001f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE <= 1) {
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        float e;
        if (MODE == 1) x[i] = fmaf(x[i], 0.999f, -0.0001f);
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x[i]));
        if (MODE == 0) x[i] = e - 1.0001f;
        if (MODE == 1) acc += e;
      }
    } else {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint32_t o[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int i = 8 * g + 2 * kk;
          float x0, x1, e0, e1;
          f2_unpack(f2_fma(f2_pack(x[i], x[i + 1]), a2, b2), x0, x1);
          x[i] = x0; x[i + 1] = x1;
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(x0));
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(x1));
          sum2[kk & 1] = f2_add(sum2[kk & 1], f2_pack(e0, e1));
          if (MODE >= 3) o[kk] = pack_bf16(e0, e1);
        }
        if (MODE == 3) keep ^= o[0] ^ o[1] ^ o[2] ^ o[3];
        if (MODE == 4) sbuf[(threadIdx.x + g * 32) & 1023] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  long long t1 = clock64();
  float s = acc, s0, s1;
  f2_unpack(f2_add(sum2[0], sum2[1]), s0, s1);
  s += s0 + s1 + __uint_as_float(keep & 0xff) + __uint_as_float(sbuf[threadIdx.x & 1023].x & 0xff);
#pragma unroll
  for (int i = 0; i < 64; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, float* out, long long* cyc) {
  const int iters = 500;
  for (int warps = 4; warps <= 16; warps *= 2) {
    const int threads = warps * 32;
    for (int rep = 0; rep < 2; ++rep) {
      k<MODE><<<148, threads>>>(out, iters, cyc);
      cudaDeviceSynchronize();
    }
    const double n = (double)iters * 64 * threads;
    printf("mode %d (%-36s) warps/scheduler %d: %6.2f ex2 / clk / SM\n", MODE, name, warps / 4, n / (double)*cyc);
  }
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMallocManaged(&cyc, 8);
  run<0>("MUFU + FADD", out, cyc);
  run<1>("FFMA + MUFU + FADD", out, cyc);
  run<2>("FFMA2 + 2 MUFU + FADD2", out, cyc);
  run<3>("FFMA2 + 2 MUFU + FADD2 + F2FP", out, cyc);
  run<4>("... + STS.128 per 8", out, cyc);
  return 0;
}
