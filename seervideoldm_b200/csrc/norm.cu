// GroupNorm (statistics pooled over ALL frames of a clip — SURVEY F6) and LayerNorm for the channels-last
// token layout [B, T = F*h*w, C].  Memory-bound kernels: coalesced 64/128-bit accesses, warp-shuffle /
// shared-memory reductions, deterministic two-level partial sums (no float atomics across CTAs).
//
// Reference semantics:
//   GroupNorm(32, C, eps) on the 5-D tensor (b, c, f, h, w): /root/reference/seer/models/resnet.py:179,197,
//   attention.py:109,133, unet_3d_condition.py:368  — statistics over (C/32, F, H, W) per sample, fp32.
//   LayerNorm(C), eps 1e-5: attention.py:198-200, 275-277.
//
// The GroupNorm input may be the virtual channel concatenation of two tensors (skip connections,
// unet_3d_blocks.py:596,712): x = cat([x1 (C1 ch), x2 (C2 ch)], channel) is never materialised in fp32.
#include "common.cuh"
#include "seer_b200.h"

namespace seer {

constexpr int GN_GROUPS = 32;
constexpr int GN_THREADS = 256;
constexpr int GN_MAX_SLOTS = 8;  // float2 columns per thread: C/2 <= 256*8 -> C <= 4096

__device__ __forceinline__ float2 load_cat2(const float* x1, int C1, const float* x2, int C2, size_t row, int c) {
  // c is even; C1 is even, so a float2 never straddles the two sources
  return c < C1 ? *reinterpret_cast<const float2*>(x1 + row * C1 + c)
                : *reinterpret_cast<const float2*>(x2 + row * C2 + (c - C1));
}

// partial[b][chunk][g][2] = (sum, sumsq) over this chunk's tokens and group g's channels
__global__ void __launch_bounds__(GN_THREADS) gn_partial_kernel(const float* __restrict__ x1, int C1,
                                                                const float* __restrict__ x2, int C2, int T, int tpc,
                                                                float* __restrict__ partial) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  const int C = C1 + C2;
  const int cpg = C / GN_GROUPS;
  const int ncol = C / 2;
  const int b = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
  const int t0 = chunk * tpc, t1 = min(T, t0 + tpc);
  __shared__ float2 s_col[GN_THREADS * GN_MAX_SLOTS];  // per-column (sum, sumsq): deterministic, no float atomics
  float sum[GN_MAX_SLOTS], sq[GN_MAX_SLOTS];
#pragma unroll
  for (int s = 0; s < GN_MAX_SLOTS; ++s) { sum[s] = 0.f; sq[s] = 0.f; }
  for (int t = t0; t < t1; ++t) {
    const size_t row = (size_t)b * T + t;
#pragma unroll
    for (int s = 0; s < GN_MAX_SLOTS; ++s) {
      const int col = threadIdx.x + s * GN_THREADS;
      if (col < ncol) {
        float2 v = load_cat2(x1, C1, x2, C2, row, col * 2);
        sum[s] += v.x + v.y;
        sq[s] += v.x * v.x + v.y * v.y;
      }
    }
  }
#pragma unroll
  for (int s = 0; s < GN_MAX_SLOTS; ++s) {
    const int col = threadIdx.x + s * GN_THREADS;
    if (col < ncol) s_col[col] = make_float2(sum[s], sq[s]);
  }
  __syncthreads();
  // 8 lanes per group, fixed summation order (cpg is even, so a float2 column never straddles two groups)
  {
    const int g = threadIdx.x >> 3, sub = threadIdx.x & 7;
    const int c0 = g * (cpg / 2), c1 = c0 + cpg / 2;
    float a = 0.f, q = 0.f;
    for (int c = c0 + sub; c < c1; c += 8) { a += s_col[c].x; q += s_col[c].y; }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (sub == 0) {
      float* dst = partial + (((size_t)b * nchunks + chunk) * GN_GROUPS + g) * 2;
      dst[0] = a;
      dst[1] = q;
    }
  }
}

// One warp per (b, g): combine chunk partials in double, then emit per-(b, c) affine: y = x*scale + shift.
__global__ void gn_finalize_kernel(const float* __restrict__ partial, int nchunks, int C, int T, float eps,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_rstd) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  const int b = blockIdx.x;
  const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cpg = C / GN_GROUPS;
  double s = 0.0, q = 0.0;
  for (int c = lane; c < nchunks; c += 32) {
    const float* src = partial + (((size_t)b * nchunks + c) * GN_GROUPS + g) * 2;
    s += (double)src[0];
    q += (double)src[1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const double n = (double)cpg * (double)T;
  const double mean = s / n;
  double var = q / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float meanf = (float)mean;
  if (lane == 0 && mean_rstd) {
    mean_rstd[((size_t)b * GN_GROUPS + g) * 2] = meanf;
    mean_rstd[((size_t)b * GN_GROUPS + g) * 2 + 1] = rstd;
  }
  for (int c = g * cpg + lane; c < (g + 1) * cpg; c += 32) {
    const float sc = rstd * gamma[c];
    scale[(size_t)b * C + c] = sc;
    shift[(size_t)b * C + c] = beta[c] - meanf * sc;
  }
}

// GroupNorm statistics from the per-(32-row slab, channel) partial sums the producing GEMM's epilogue emitted
// (SeerGemmDesc::col_stats): one block per (group, sample) sums its channels over the sample's T/32 slabs in double
// and emits the same per-(b, c) affine as gn_finalize_kernel.  The input may be the virtual concat of two tensors,
// each with its own statistics buffer.
__global__ void __launch_bounds__(256) gn_finalize_cols_kernel(const float2* __restrict__ st1, int C1,
                                                               const float2* __restrict__ st2, int C2, int T, float eps,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               float* __restrict__ scale, float* __restrict__ shift) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  const int C = C1 + C2;
  const int cpg = C / GN_GROUPS;
  const int g = blockIdx.x, b = blockIdx.y;
  const int slabs = T / 32;
  // thread = (slab stripe, channel of the group): fixed channel, slabs strided by the number of stripes, four independent
  // loads in flight per thread and no division in the loop (the kernel is latency-bound: 20 MB of statistics at most).
  // Per-thread partials in compensated fp32 (Kahan) — the fp64 pipe of this part runs at 1/64 rate; threads are combined in
  // double below.
  float sf = 0.f, qf = 0.f, sc_ = 0.f, qc_ = 0.f;
  auto kahan = [&](float2 v) {
    const float ys = __fsub_rn(v.x, sc_), ts = __fadd_rn(sf, ys);
    sc_ = __fsub_rn(__fsub_rn(ts, sf), ys);
    sf = ts;
    const float yq = __fsub_rn(v.y, qc_), tq = __fadd_rn(qf, yq);
    qc_ = __fsub_rn(__fsub_rn(tq, qf), yq);
    qf = tq;
  };
  const int stripes = blockDim.x / cpg;
  const int cl = threadIdx.x % cpg, st = threadIdx.x / cpg;
  if (st < stripes) {
    const int c = g * cpg + cl;
    const bool first = c < C1;
    const float2* base = first ? st1 + c : st2 + (c - C1);
    const size_t ld = first ? C1 : C2;
    const size_t slab0 = (size_t)b * slabs;
    int sl = st;
    for (; sl + 3 * stripes < slabs; sl += 4 * stripes) {
      const float2 v0 = __ldg(base + (slab0 + sl) * ld);
      const float2 v1 = __ldg(base + (slab0 + sl + stripes) * ld);
      const float2 v2 = __ldg(base + (slab0 + sl + 2 * stripes) * ld);
      const float2 v3 = __ldg(base + (slab0 + sl + 3 * stripes) * ld);
      kahan(v0); kahan(v1); kahan(v2); kahan(v3);
    }
    for (; sl < slabs; sl += stripes) kahan(__ldg(base + (slab0 + sl) * ld));
  }
  double s = (double)sf - (double)sc_, q = (double)qf - (double)qc_;
  __shared__ double sh[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = q; }
  __syncthreads();
  s = 0.0; q = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) { s += sh[0][w]; q += sh[1][w]; }     // fixed order: deterministic
  const double n = (double)cpg * (double)T;
  const double mean = s / n;
  double var = q / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float meanf = (float)mean;
  for (int c = g * cpg + threadIdx.x; c < (g + 1) * cpg; c += blockDim.x) {
    const float sc = rstd * gamma[c];
    scale[(size_t)b * C + c] = sc;
    shift[(size_t)b * C + c] = beta[c] - meanf * sc;
  }
}

// y = act(x * scale[b, c] + shift[b, c]).  Optional second output: raw bf16 copy of x (the un-normalised concat that feeds
// the fused ResNet 1x1 shortcut GEMM).  HBM-bound (4 B in, 2 B out per element): a thread owns ONE 8-channel chunk for
// its whole life — the 16 scale / shift values sit in registers, the inner loop is two 16-byte loads, 8 FMAs (+ SiLU) and
// one 16-byte store per row with no index arithmetic, unrolled by 4 rows so eight loads are in flight per thread.
// Block = rps rows x (C/8) chunks (consecutive threads = consecutive chunks of a row: 32-byte pieces of one contiguous row),
// it walks `rows_per_block` rows of ONE sample; grid = (blocks per sample, B).
template <bool OUT_F32, bool IN_BF16>
__global__ void __launch_bounds__(512) gn_apply_kernel(const void* __restrict__ x1v, int C1, const void* __restrict__ x2v,
                                                       int C2, int T, int rows_per_block, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, int silu, void* __restrict__ y,
                                                       __nv_bfloat16* __restrict__ raw) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  const int C = C1 + C2;
  const int c8n = C / 8;
  const int rps = blockDim.x / c8n;                 // rows per step
  const int chunk = threadIdx.x % c8n, rl = threadIdx.x / c8n;
  const int c = chunk * 8;
  const int b = blockIdx.y;
  const int r_begin = blockIdx.x * rows_per_block;
  const int r_end = min(T, r_begin + rows_per_block);
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + (size_t)b * C + c));
  const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale + (size_t)b * C + c + 4));
  const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift + (size_t)b * C + c));
  const float4 h1 = __ldg(reinterpret_cast<const float4*>(shift + (size_t)b * C + c + 4));
  const bool first = c < C1;
  const int ldx = first ? C1 : C2;
  // IN_BF16: both sources are bf16 tensors (conv1's output; or, with the bf16 residual stream, a block output and its skip partner)
  const float* src = IN_BF16 ? nullptr
                             : (first ? reinterpret_cast<const float*>(x1v) + c : reinterpret_cast<const float*>(x2v) + (c - C1)) +
                                   ((size_t)b * T + r_begin + rl) * ldx;
  const __nv_bfloat16* src16 = IN_BF16 ? (first ? reinterpret_cast<const __nv_bfloat16*>(x1v) + c
                                                : reinterpret_cast<const __nv_bfloat16*>(x2v) + (c - C1)) + ((size_t)b * T + r_begin + rl) * ldx
                                       : nullptr;
  const size_t out_off = ((size_t)b * T + r_begin + rl) * C + c;
  const size_t in_step = (size_t)rps * ldx, out_step = (size_t)rps * C;

  auto emit = [&](const float4& a0, const float4& a1, size_t off) {
    float v[8] = {fmaf(a0.x, s0.x, h0.x), fmaf(a0.y, s0.y, h0.y), fmaf(a0.z, s0.z, h0.z), fmaf(a0.w, s0.w, h0.w),
                  fmaf(a1.x, s1.x, h1.x), fmaf(a1.y, s1.y, h1.y), fmaf(a1.z, s1.z, h1.z), fmaf(a1.w, s1.w, h1.w)};
    if (silu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = OUT_F32 ? silu_f(v[j]) : silu_fast(v[j]);    // fp32 output = the fp32-parity path: exact form
    }
    if (OUT_F32) {
      float* dst = reinterpret_cast<float*>(y) + off;
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      uint4 o;
      o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(y) + off) = o;
    }
    if (raw) {
      uint4 o;
      o.x = pack_bf16(a0.x, a0.y); o.y = pack_bf16(a0.z, a0.w); o.z = pack_bf16(a1.x, a1.y); o.w = pack_bf16(a1.z, a1.w);
      *reinterpret_cast<uint4*>(raw + off) = o;
    }
  };
  auto emit16 = [&](const uint4& a, size_t off) {
    const float2 p0 = unpack_bf16(a.x), p1 = unpack_bf16(a.y), p2 = unpack_bf16(a.z), p3 = unpack_bf16(a.w);
    emit(make_float4(p0.x, p0.y, p1.x, p1.y), make_float4(p2.x, p2.y, p3.x, p3.y), off);
  };

  if (rl >= rps) return;                            // threads beyond rps * c8n (block size rounded up to a warp multiple)
  int r = r_begin + rl;
  size_t oo = out_off;
  if constexpr (IN_BF16) {
    for (; r + 3 * rps < r_end; r += 4 * rps) {
      uint4 a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = __ldcs(reinterpret_cast<const uint4*>(src16 + u * in_step));
#pragma unroll
      for (int u = 0; u < 4; ++u) emit16(a[u], oo + u * out_step);
      src16 += 4 * in_step;
      oo += 4 * out_step;
    }
    for (; r < r_end; r += rps) {
      emit16(__ldcs(reinterpret_cast<const uint4*>(src16)), oo);
      src16 += in_step;
      oo += out_step;
    }
  } else {
    for (; r + 3 * rps < r_end; r += 4 * rps) {
      float4 a0[4], a1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a0[u] = __ldcs(reinterpret_cast<const float4*>(src + u * in_step));
        a1[u] = __ldcs(reinterpret_cast<const float4*>(src + u * in_step + 4));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) emit(a0[u], a1[u], oo + u * out_step);
      src += 4 * in_step;
      oo += 4 * out_step;
    }
    for (; r < r_end; r += rps) {
      const float4 a0 = __ldcs(reinterpret_cast<const float4*>(src));
      const float4 a1 = __ldcs(reinterpret_cast<const float4*>(src + 4));
      emit(a0, a1, oo);
      src += in_step;
      oo += out_step;
    }
  }
}

// launch geometry of gn_apply_kernel: block = rps rows x C/8 chunks (rounded up to a warp multiple), blocks per sample
static inline void gn_apply_geometry(int T, int C, int B, int& threads, int& rows_per_block, int& blocks_per_sample) {
  const int c8n = C / 8;
  const int rps = c8n >= 256 ? 1 : 256 / c8n;
  threads = ((rps * c8n + 31) / 32) * 32;
  int want = (148 * 12 + B - 1) / B;                  // ~12 blocks per SM in total
  if (want < 1) want = 1;
  rows_per_block = (T + want - 1) / want;
  if (rows_per_block < 4 * rps) rows_per_block = 4 * rps;
  rows_per_block = ((rows_per_block + rps - 1) / rps) * rps;
  blocks_per_sample = (T + rows_per_block - 1) / rows_per_block;
}

// LayerNorm over the last dim, one warp per row, two-pass in registers (exact mean/variance), bf16 out.
constexpr int LN_MAX_V4 = 10;  // C <= 1280
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int M, int C, int ldx,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, __nv_bfloat16* __restrict__ y, int ldy) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int nv = C / 4;
  const float4* src = reinterpret_cast<const float4*>(x + (size_t)row * ldx);
  float4 v[LN_MAX_V4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i) {
    const int j = lane + i * 32;
    if (j < nv) {
      v[i] = src[j];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i) {
    const int j = lane + i * 32;
    if (j < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  uint2* dst = reinterpret_cast<uint2*>(y + (size_t)row * ldy);
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i) {
    const int j = lane + i * 32;
    if (j < nv) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + j);
      const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + j);
      uint2 o;
      o.x = pack_bf16((v[i].x - mean) * rstd * g.x + be.x, (v[i].y - mean) * rstd * g.y + be.y);
      o.y = pack_bf16((v[i].z - mean) * rstd * g.z + be.z, (v[i].w - mean) * rstd * g.w + be.w);
      dst[j] = o;
    }
  }
}

}  // namespace seer

using namespace seer;

extern "C" int seer_b200_groupnorm_tokens_per_chunk(int T) { return T <= 4096 ? 16 : 32; }

extern "C" int seer_b200_groupnorm_workspace_floats(int B, int T) {
  const int tpc = seer_b200_groupnorm_tokens_per_chunk(T);
  return B * ceil_div(T, tpc) * GN_GROUPS * 2;
}

extern "C" int seer_b200_groupnorm(const float* x1, int C1, const float* x2, int C2, int B, int T, const float* gamma,
                                   const float* beta, float eps, int silu, float* workspace, float* scale_shift, void* y,
                                   int y_is_f32, void* raw_bf16, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int C = C1 + C2;
  SEER_CHECK_ARG(x1 && gamma && beta && workspace && scale_shift && y && B > 0 && T > 0);
  SEER_CHECK_ARG(C1 % 8 == 0 && C2 % 8 == 0 && (C2 == 0 || x2) && C % (2 * GN_GROUPS) == 0);
  SEER_CHECK_ARG(C / 2 <= GN_THREADS * GN_MAX_SLOTS && C / 8 <= 512);
  const int tpc = seer_b200_groupnorm_tokens_per_chunk(T);
  const int nchunks = ceil_div(T, tpc);
  float* scale = scale_shift;
  float* shift = scale_shift + (size_t)B * C;
  { cudaError_t le__ = launch_pdl(gn_partial_kernel, dim3(nchunks, B), GN_THREADS, 0, stream, x1, C1, x2, C2, T, tpc, workspace); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  { cudaError_t le__ = launch_pdl(gn_finalize_kernel, B, GN_GROUPS * 32, 0, stream, workspace, nchunks, C, T, eps, gamma, beta, scale, shift, nullptr); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  int ga_threads, ga_rows, ga_blocks;
  gn_apply_geometry(T, C, B, ga_threads, ga_rows, ga_blocks);
  if (y_is_f32)
    { cudaError_t le__ = launch_pdl(gn_apply_kernel<true, false>, dim3(ga_blocks, B), ga_threads, 0, stream, (const void*)x1, C1, (const void*)x2, C2, T, ga_rows, scale, shift, silu, y, (__nv_bfloat16*)raw_bf16); if (le__ != cudaSuccess) return (int)le__; }
  else
    { cudaError_t le__ = launch_pdl(gn_apply_kernel<false, false>, dim3(ga_blocks, B), ga_threads, 0, stream, (const void*)x1, C1, (const void*)x2, C2, T, ga_rows, scale, shift, silu, y, (__nv_bfloat16*)raw_bf16); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_groupnorm_from_stats_ex(const void* x1, int x1_is_bf16, int C1, const float* stats1, const void* x2, int C2,
                                                 const float* stats2, int B, int T, const float* gamma, const float* beta, float eps,
                                                 int silu, float* scale_shift, void* y, int y_is_f32, void* raw_bf16, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int C = C1 + C2;
  SEER_CHECK_ARG(x1 && stats1 && gamma && beta && scale_shift && y && B > 0 && T > 0 && T % 32 == 0);
  SEER_CHECK_ARG(C1 % 8 == 0 && C2 % 8 == 0 && (C2 == 0 || (x2 && stats2)) && C % (2 * GN_GROUPS) == 0 && C / 8 <= 512);
  SEER_CHECK_ARG(!x1_is_bf16 || !y_is_f32);
  float* scale = scale_shift;
  float* shift = scale_shift + (size_t)B * C;
  { cudaError_t le__ = launch_pdl(gn_finalize_cols_kernel, dim3(GN_GROUPS, B), 256, 0, stream, (const float2*)stats1, C1, (const float2*)stats2, C2, T, eps,
                                                                  gamma, beta, scale, shift); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  int ga_threads, ga_rows, ga_blocks;
  gn_apply_geometry(T, C, B, ga_threads, ga_rows, ga_blocks);
  cudaError_t le__;
  if (x1_is_bf16)
    le__ = launch_pdl(gn_apply_kernel<false, true>, dim3(ga_blocks, B), ga_threads, 0, stream, x1, C1, x2, C2, T, ga_rows, scale, shift, silu, y, (__nv_bfloat16*)raw_bf16);
  else if (y_is_f32)
    le__ = launch_pdl(gn_apply_kernel<true, false>, dim3(ga_blocks, B), ga_threads, 0, stream, x1, C1, x2, C2, T, ga_rows, scale, shift, silu, y, (__nv_bfloat16*)raw_bf16);
  else
    le__ = launch_pdl(gn_apply_kernel<false, false>, dim3(ga_blocks, B), ga_threads, 0, stream, x1, C1, x2, C2, T, ga_rows, scale, shift, silu, y, (__nv_bfloat16*)raw_bf16);
  if (le__ != cudaSuccess) return (int)le__;
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_groupnorm_from_stats(const float* x1, int C1, const float* stats1, const float* x2, int C2,
                                              const float* stats2, int B, int T, const float* gamma, const float* beta, float eps,
                                              int silu, float* scale_shift, void* y, int y_is_f32, void* raw_bf16, void* stream_) {
  return seer_b200_groupnorm_from_stats_ex(x1, 0, C1, stats1, x2, C2, stats2, B, T, gamma, beta, eps, silu, scale_shift, y, y_is_f32,
                                           raw_bf16, stream_);
}

extern "C" int seer_b200_layernorm(const float* x, int M, int C, int ldx, const float* gamma, const float* beta, float eps,
                                   void* y_bf16, int ldy, void* stream) {
  SEER_CHECK_ARG(x && gamma && beta && y_bf16 && M > 0);
  SEER_CHECK_ARG(C % 4 == 0 && C <= LN_MAX_V4 * 128 && ldx % 4 == 0 && ldy % 4 == 0);
  { cudaError_t le__ = launch_pdl(layernorm_kernel, ceil_div(M, 8), 256, 0, (cudaStream_t)stream, x, M, C, ldx, gamma, beta, eps, (__nv_bfloat16*)y_bf16, ldy); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}
