// fp32-parity path (north-star: per-step noise prediction within rel-L2 1e-4 of the reference's fp32 PyTorch result).
//
// The tensor cores have no fp32 input mode, so the fp32 path runs every contraction of the UNet on the SAME tcgen05
// GEMM / implicit-GEMM conv kernel (gemm_tc.cu) with error-compensated bf16 operands:
//
//     a = a_hi + a_lo (+ O(2^-17 |a|)),  a_hi = bf16(a), a_lo = bf16(a - a_hi);   same for the weights w
//     a . w  ~=  a_hi . w_hi  +  a_hi . w_lo  +  a_lo . w_hi          (the dropped a_lo . w_lo term is O(2^-18))
//
// realised as ONE GEMM over a 3x longer K: A' = [a_hi | a_hi | a_lo] (written by split3_kernel below),
// W' = [w_hi | w_lo | w_hi] (packing.split3_weight), fp32 accumulation in TMEM.  Everything between the GEMMs stays
// fp32 in HBM and uses the small SIMT kernels of this file (exact erf GELU, accurate exp2f softmax) — this path is a
// parity mode, its speed is irrelevant; the bf16 path (fused epilogues, tcgen05 attention) is the product.
//
// Reference operators: nn.LayerNorm attention.py:198-200; GEGLU attention.py:791-793; RoPE attention.py:649-651 (and the
// frame-axis RoPE of the FSText temporal blocks, attention.py:529-530); softmax attention attention.py:622-630.
#include "common.cuh"
#include "seer_b200.h"

namespace seer {

static inline unsigned f32_grid(size_t total, int threads) {
  size_t b = (total + threads - 1) / threads;
  const size_t cap = 148 * 16;
  return (unsigned)(b < cap ? (b ? b : 1) : cap);
}

// ---------------------------------------------------------------------------------------------------
// split3: x fp32 [rows_in, C] (ldx) -> out bf16 [rows_out, ldo]:
//   out[r, col0 + c] = hi, out[r, Ctot + col0 + c] = hi, out[r, 2 Ctot + col0 + c] = lo
// (col0 / Ctot place one part of a channel concat; up_H > 0: rows are pixels of [n_img, 2 up_H, 2 up_W] images and
// read the nearest-neighbour source pixel of [n_img, up_H, up_W] — Upsample3D's F.interpolate, resnet.py:52)
// ---------------------------------------------------------------------------------------------------
__global__ void split3_kernel(const float* __restrict__ x, int ldx, int C, __nv_bfloat16* __restrict__ out, int ldo, int Ctot,
                              int col0, size_t rows_out, int up_H, int up_W) {
  const int c4n = C / 4;
  const size_t total = rows_out * c4n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / c4n;
    const int c = (int)(i - r * c4n) * 4;
    size_t rin = r;
    if (up_H > 0) {
      const int ox = (int)(r % (2 * up_W));
      const size_t t = r / (2 * up_W);
      const int oy = (int)(t % (2 * up_H));
      const size_t img = t / (2 * up_H);
      rin = (img * up_H + oy / 2) * up_W + ox / 2;
    }
    const float4 a = *reinterpret_cast<const float4*>(x + rin * ldx + c);
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
    const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
    uint2 hi, lo;
    hi.x = *reinterpret_cast<const uint32_t*>(&h0);
    hi.y = *reinterpret_cast<const uint32_t*>(&h1);
    lo.x = pack_bf16(a.x - f0.x, a.y - f0.y);
    lo.y = pack_bf16(a.z - f1.x, a.w - f1.y);
    __nv_bfloat16* o = out + r * ldo + col0 + c;
    *reinterpret_cast<uint2*>(o) = hi;
    *reinterpret_cast<uint2*>(o + Ctot) = hi;
    *reinterpret_cast<uint2*>(o + 2 * Ctot) = lo;
  }
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm, fp32 in / fp32 out, one warp per row, two-pass in registers
// ---------------------------------------------------------------------------------------------------
constexpr int LN32_MAX_V4 = 10;  // C <= 1280
__global__ void __launch_bounds__(256) layernorm_f32_kernel(const float* __restrict__ x, int M, int C, int ldx,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float eps, float* __restrict__ y, int ldy) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int nv = C / 4;
  const float4* src = reinterpret_cast<const float4*>(x + (size_t)row * ldx);
  float4 v[LN32_MAX_V4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN32_MAX_V4; ++i) {
    const int j = lane + i * 32;
    if (j < nv) {
      v[i] = src[j];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN32_MAX_V4; ++i) {
    const int j = lane + i * 32;
    if (j < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
  float4* dst = reinterpret_cast<float4*>(y + (size_t)row * ldy);
#pragma unroll
  for (int i = 0; i < LN32_MAX_V4; ++i) {
    const int j = lane + i * 32;
    if (j < nv) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + j);
      const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + j);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + be.x;
      o.y = (v[i].y - mean) * rstd * g.y + be.y;
      o.z = (v[i].z - mean) * rstd * g.z + be.z;
      o.w = (v[i].w - mean) * rstd * g.w + be.w;
      dst[j] = o;
    }
  }
}

// GEGLU: h [M, 2I] fp32 -> out [M, I] = h[:, :I] * gelu_erf(h[:, I:])   (attention.py:791-793, exact erff)
__global__ void geglu_f32_kernel(const float* __restrict__ h, int ldh, float* __restrict__ out, int ldo, size_t M, int I) {
  const int i4n = I / 4;
  const size_t total = M * i4n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / i4n;
    const int c = (int)(i - r * i4n) * 4;
    const float4 a = *reinterpret_cast<const float4*>(h + r * ldh + c);
    const float4 g = *reinterpret_cast<const float4*>(h + r * ldh + I + c);
    float4 o;
    o.x = a.x * gelu_erf_f(g.x); o.y = a.y * gelu_erf_f(g.y); o.z = a.z * gelu_erf_f(g.z); o.w = a.w * gelu_erf_f(g.w);
    *reinterpret_cast<float4*>(out + r * ldo + c) = o;
  }
}

// RoPE (interleaved pairs over the first 2*half channels of every head) in place on the Q and K column blocks;
// position = (row / pos_div) % pos_mod.  T = float (fp32 path) or __nv_bfloat16 (FSText frame-axis RoPE, bf16 path).
template <typename T>
__global__ void rope_ex_kernel(T* __restrict__ qk, int ld, int M, int pos_div, int pos_mod, int heads, int head_dim, int q_col,
                               int k_col, const float* __restrict__ freqs, int half) {
  const size_t total = (size_t)M * half;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / half);
    const int j = (int)(i - (size_t)row * half);
    const int pos = (row / pos_div) % pos_mod;
    const float ang = (float)pos * __ldg(freqs + j);
    float sn, cs;
    sincosf(ang, &sn, &cs);
    T* base = qk + (size_t)row * ld + 2 * j;
    for (int h = 0; h < heads; ++h) {
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        T* ptr = base + (w ? k_col : q_col) + h * head_dim;
        if constexpr (sizeof(T) == 4) {
          const float2 x = *reinterpret_cast<float2*>(ptr);
          *reinterpret_cast<float2*>(ptr) = make_float2(x.x * cs - x.y * sn, x.y * cs + x.x * sn);
        } else {
          const float2 x = unpack_bf16(*reinterpret_cast<uint32_t*>(ptr));
          *reinterpret_cast<uint32_t*>(ptr) = pack_bf16(x.x * cs - x.y * sn, x.y * cs + x.x * sn);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// fp32 softmax attention, one warp per query, 32 keys per step (lane = key for Q K^T, lane = channel for P V).
// Same row-index functions (head split / window partition / frame axis) as the bf16 kernels (attention.cu).
// ---------------------------------------------------------------------------------------------------
struct Attn32Params {
  const float* q; int ldq;
  const float* k; int ldk;
  const float* v; int ldv;
  float* o; int ldo;
  int windowed;      // 0: rows = outer * L + s (SPATIAL / CROSS);  1: SCTA / FRAME geometry below
  int heads, Lq, Lk, causal;
  int F, H, W, ws, nwx, nwin;
  float scale_log2;
};

__device__ __forceinline__ int attn32_row(const Attn32Params& p, int b, int win, int s) {
  if (p.ws == 0) return b * p.F * p.H * p.W + s;
  const int ws2 = p.ws * p.ws;
  const int f = s / ws2;
  const int r = s - f * ws2;
  const int iy = r / p.ws, ix = r - iy * p.ws;
  const int wy = win / p.nwx, wx = win - wy * p.nwx;
  return ((b * p.F + f) * p.H + wy * p.ws + iy) * p.W + wx * p.ws + ix;
}

constexpr int A32_WARPS = 16;
constexpr int A32_KT = 32;

template <int D>
__global__ void __launch_bounds__(A32_WARPS * 32) attention_f32_kernel(const Attn32Params p) {
  constexpr int LDS = D + 4;                 // 16-byte aligned rows, conflict-free float4 reads with lane = row
  constexpr int NV = (D + 31) / 32;
  constexpr int D4 = D / 4;
  extern __shared__ __align__(16) float smem_f[];
  float* sQ = smem_f;                        // [A32_WARPS][D]
  float* sK = sQ + A32_WARPS * D;            // [32][LDS]
  float* sV = sK + A32_KT * LDS;             // [32][LDS]
  __shared__ int kv_rows[A32_KT];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * A32_WARPS;
  const int qi = q0 + warp;
  const int head = blockIdx.y % p.heads;
  const int outer = blockIdx.y / p.heads;
  int b = 0, win = 0;
  if (p.windowed) { b = outer / p.nwin; win = outer - b * p.nwin; }
  const int col0 = head * D;
  const int qrow = qi < p.Lq ? (p.windowed ? attn32_row(p, b, win, qi) : outer * p.Lq + qi) : -1;

  for (int c = lane; c < D; c += 32) sQ[warp * D + c] = qrow >= 0 ? p.q[(size_t)qrow * p.ldq + col0 + c] : 0.f;

  int n_tiles = ceil_div(p.Lk, A32_KT);
  if (p.causal) n_tiles = min(n_tiles, ceil_div(min(q0 + A32_WARPS, p.Lq), A32_KT));
  float m = -INFINITY, l = 0.f;
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;

  for (int t = 0; t < n_tiles; ++t) {
    __syncthreads();                         // previous tile consumed (and sQ written, first iteration)
    if (tid < A32_KT) {
      const int s = t * A32_KT + tid;
      kv_rows[tid] = s < p.Lk ? (p.windowed ? attn32_row(p, b, win, s) : outer * p.Lk + s) : -1;
    }
    __syncthreads();
    for (int i = tid; i < A32_KT * D4; i += A32_WARPS * 32) {
      const int r = i / D4, c = (i - r * D4) * 4;
      const int g = kv_rows[r];
      float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
      if (g >= 0) {
        kk = *reinterpret_cast<const float4*>(p.k + (size_t)g * p.ldk + col0 + c);
        vv = *reinterpret_cast<const float4*>(p.v + (size_t)g * p.ldv + col0 + c);
      }
      *reinterpret_cast<float4*>(sK + r * LDS + c) = kk;
      *reinterpret_cast<float4*>(sV + r * LDS + c) = vv;
    }
    __syncthreads();
    if (qrow < 0 || (p.causal && t * A32_KT > qi)) continue;
    const int kj = t * A32_KT + lane;
    float dot = 0.f;
    const float4* qp = reinterpret_cast<const float4*>(sQ + warp * D);
    const float4* kp = reinterpret_cast<const float4*>(sK + lane * LDS);
#pragma unroll 10
    for (int c = 0; c < D4; ++c) {
      const float4 a = qp[c], kk = kp[c];
      dot = fmaf(a.x, kk.x, dot); dot = fmaf(a.y, kk.y, dot); dot = fmaf(a.z, kk.z, dot); dot = fmaf(a.w, kk.w, dot);
    }
    const bool valid = kj < p.Lk && (!p.causal || kj <= qi);
    const float s = valid ? dot * p.scale_log2 : -INFINITY;
    const float m_new = fmaxf(m, warp_max(s));          // finite: key 0 is visible to every query in tile 0
    const float pj = valid ? exp2f(s - m_new) : 0.f;
    const float corr = exp2f(m - m_new);                // first tile: exp2(-inf) = 0
    l = l * corr + warp_sum(pj);
    m = m_new;
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] *= corr;
    for (int j = 0; j < A32_KT; ++j) {
      const float pjj = __shfl_sync(0xffffffffu, pj, j);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int dim = lane + 32 * i;
        if (dim < D) acc[i] = fmaf(pjj, sV[j * LDS + dim], acc[i]);
      }
    }
  }
  if (qrow >= 0) {
    const float inv = 1.0f / l;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int dim = lane + 32 * i;
      if (dim < D) p.o[(size_t)qrow * p.ldo + col0 + dim] = acc[i] * inv;
    }
  }
}

template <int D>
static int launch_attention_f32(const Attn32Params& p, int n_problems, cudaStream_t stream) {
  constexpr int SMEM = (A32_WARPS * D + 2 * A32_KT * (D + 4)) * 4;
  static SmemAttrOnce smem_attr_attr_done;
  { cudaError_t e = smem_attr_attr_done.ensure(attention_f32_kernel<D>, SMEM); if (e != cudaSuccess) return (int)e; }
  dim3 grid(ceil_div(p.Lq, A32_WARPS), n_problems);
  attention_f32_kernel<D><<<grid, A32_WARPS * 32, SMEM, stream>>>(p);
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

}  // namespace seer

using namespace seer;

extern "C" int seer_b200_split3_bf16(const float* x, int ldx, long long rows_in, int C, void* out, int ldo, int Ctot, int col0,
                                     int up_n_img, int up_H, int up_W, void* stream) {
  SEER_CHECK_ARG(x && out && rows_in > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 && Ctot % 4 == 0 && col0 % 4 == 0);
  SEER_CHECK_ARG(col0 + C <= Ctot && 3 * Ctot <= ldo);
  size_t rows_out = (size_t)rows_in;
  if (up_H > 0) {
    SEER_CHECK_ARG(up_n_img > 0 && up_W > 0 && (long long)up_n_img * up_H * up_W == rows_in);
    rows_out = (size_t)rows_in * 4;
  }
  split3_kernel<<<f32_grid(rows_out * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, C, (__nv_bfloat16*)out, ldo, Ctot, col0,
                                                                                    rows_out, up_H > 0 ? up_H : 0, up_W);
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_layernorm_f32(const float* x, int M, int C, int ldx, const float* gamma, const float* beta, float eps,
                                       float* y, int ldy, void* stream) {
  SEER_CHECK_ARG(x && gamma && beta && y && M > 0);
  SEER_CHECK_ARG(C % 4 == 0 && C <= LN32_MAX_V4 * 128 && ldx % 4 == 0 && ldy % 4 == 0);
  layernorm_f32_kernel<<<ceil_div(M, 8), 256, 0, (cudaStream_t)stream>>>(x, M, C, ldx, gamma, beta, eps, y, ldy);
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_geglu_f32(const float* h, int ldh, float* out, int ldo, long long M, int inner, void* stream) {
  SEER_CHECK_ARG(h && out && M > 0 && inner > 0 && inner % 4 == 0 && ldh % 4 == 0 && ldo % 4 == 0);
  geglu_f32_kernel<<<f32_grid((size_t)M * (inner / 4), 256), 256, 0, (cudaStream_t)stream>>>(h, ldh, out, ldo, (size_t)M, inner);
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_rope_ex(void* qk, int is_f32, int ld, int M, int pos_div, int pos_mod, int heads, int head_dim, int q_col,
                                 int k_col, const float* freqs, int n_freqs, void* stream) {
  SEER_CHECK_ARG(qk && freqs && M > 0 && pos_div > 0 && pos_mod > 0);
  SEER_CHECK_ARG(2 * n_freqs <= head_dim && ld % 2 == 0 && q_col % 2 == 0 && k_col % 2 == 0 && head_dim % 2 == 0);
  const unsigned grid = f32_grid((size_t)M * n_freqs, 256);
  if (is_f32)
    rope_ex_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((float*)qk, ld, M, pos_div, pos_mod, heads, head_dim, q_col, k_col,
                                                                 freqs, n_freqs);
  else
    rope_ex_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)qk, ld, M, pos_div, pos_mod, heads, head_dim,
                                                                         q_col, k_col, freqs, n_freqs);
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_attention_f32(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* o, int ldo,
                                       int mode, int heads, int head_dim, int n_outer, int Lq, int Lk, int F, int H, int W,
                                       void* stream) {
  SEER_CHECK_ARG(q && k && v && o && heads > 0 && n_outer > 0);
  SEER_CHECK_ARG(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && head_dim % 4 == 0);
  SEER_CHECK_ARG(((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0));
  Attn32Params p{};
  p.q = q; p.ldq = ldq; p.k = k; p.ldk = ldk; p.v = v; p.ldv = ldv; p.o = o; p.ldo = ldo;
  p.heads = heads;
  p.scale_log2 = (1.0f / sqrtf((float)head_dim)) * 1.4426950408889634f;
  int n_problems;
  if (mode == SEER_ATTN_SCTA) {
    SEER_CHECK_ARG(F > 0 && H > 0 && W > 0);
    p.windowed = 1; p.causal = 1;
    p.F = F; p.H = H; p.W = W;
    p.ws = (H <= 4) ? 0 : ((H / 8) >= 4 ? 8 : 4);          // window rule of attention.py:30-33,661-668
    if (p.ws == 0) { p.nwx = 1; p.nwin = 1; p.Lq = p.Lk = F * H * W; }
    else {
      SEER_CHECK_ARG(H % p.ws == 0 && W % p.ws == 0);
      p.nwx = W / p.ws; p.nwin = (H / p.ws) * p.nwx; p.Lq = p.Lk = F * p.ws * p.ws;
    }
    n_problems = n_outer * p.nwin * heads;
  } else if (mode == SEER_ATTN_FRAME) {
    // sequence = the F frames of token l (H = tokens per frame) of clip b, causal: a 1x1 "window" per token
    SEER_CHECK_ARG(F > 0 && H > 0);
    p.windowed = 1; p.causal = 1;
    p.F = F; p.H = H; p.W = 1; p.ws = 1; p.nwx = 1; p.nwin = H; p.Lq = p.Lk = F;
    n_problems = n_outer * p.nwin * heads;
  } else if (mode == SEER_ATTN_SPATIAL || mode == SEER_ATTN_CROSS) {
    SEER_CHECK_ARG(Lq > 0 && Lk > 0);
    p.windowed = 0; p.causal = 0; p.Lq = Lq; p.Lk = Lk; p.nwin = 1; p.nwx = 1;
    n_problems = n_outer * heads;
  } else {
    return SEER_EINVAL;
  }
  if (n_problems > 65535) return SEER_EUNSUPPORTED;
  switch (head_dim) {
    case 40: return launch_attention_f32<40>(p, n_problems, (cudaStream_t)stream);
    case 80: return launch_attention_f32<80>(p, n_problems, (cudaStream_t)stream);
    case 96: return launch_attention_f32<96>(p, n_problems, (cudaStream_t)stream);
    case 160: return launch_attention_f32<160>(p, n_problems, (cudaStream_t)stream);
    default: return SEER_EUNSUPPORTED;
  }
}
