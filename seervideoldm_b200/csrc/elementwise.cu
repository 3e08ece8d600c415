// Memory-bound pieces of the Seer denoising step: RoPE, timestep embedding + small-batch linears, the 4-channel
// boundary convs (kept in fp32 — SURVEY F11), nearest-2x upsample, stride-2 im2col, dtype casts and the fused
// CFG-combine + DDIM update.  Coalesced 128-bit accesses, warp-shuffle reductions.
#include "common.cuh"
#include "seer_b200.h"

namespace seer {

// ---------------------------------------------------------------------------------------------------
// cast fp32 -> bf16 (context embeddings, once per clip)
// ---------------------------------------------------------------------------------------------------
__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n8) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = *reinterpret_cast<const float4*>(x + i * 8);
    const float4 b = *reinterpret_cast<const float4*>(x + i * 8 + 4);
    uint4 o;
    o.x = pack_bf16(a.x, a.y); o.y = pack_bf16(a.z, a.w); o.z = pack_bf16(b.x, b.y); o.w = pack_bf16(b.z, b.w);
    *reinterpret_cast<uint4*>(y + i * 8) = o;
  }
}

// ---------------------------------------------------------------------------------------------------
// RoPE on the Q and K column blocks of a token-major projection buffer [M, ld] (bf16), in place.
// Reference: rotary-embedding-torch 0.1.5 rotate_queries_or_keys at /root/reference/seer/models/attention.py:649-651:
// position = flat token index f*h*w + y*w + x inside the clip; interleaved pairs (2j, 2j+1), j < rot/2,
// angle = pos * freqs[j]; channels >= rot of each head pass through.
// One thread per (token, pair j): sincos once, applied to all heads of Q and K.
// ---------------------------------------------------------------------------------------------------
__global__ void rope_kernel(__nv_bfloat16* __restrict__ qk, int ld, int M, int tokens_per_clip, int heads, int head_dim,
                            int q_col, int k_col, const float* __restrict__ freqs, int half) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  const size_t total = (size_t)M * half;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / half);
    const int j = (int)(i - (size_t)row * half);
    const int pos = row % tokens_per_clip;
    const float ang = (float)pos * __ldg(freqs + j);
    float sn, cs;
    sincosf(ang, &sn, &cs);
    __nv_bfloat16* base = qk + (size_t)row * ld + 2 * j;
    // all loads of a batch of heads first, then the stores: the read-modify-write chain per (head, q/k) serialised 16 global
    // round trips per thread (ncu: 63 us at 42 % DRAM throughput, issue slots 20 % busy — latency-bound)
    constexpr int HB = 8;
    for (int h0 = 0; h0 < heads; h0 += HB) {
      uint32_t raw[HB][2];
#pragma unroll
      for (int hh = 0; hh < HB; ++hh)
#pragma unroll
        for (int w = 0; w < 2; ++w)
          if (h0 + hh < heads) raw[hh][w] = *reinterpret_cast<const uint32_t*>(base + (w ? k_col : q_col) + (h0 + hh) * head_dim);
#pragma unroll
      for (int hh = 0; hh < HB; ++hh)
#pragma unroll
        for (int w = 0; w < 2; ++w)
          if (h0 + hh < heads) {
            const float2 x = unpack_bf16(raw[hh][w]);
            *reinterpret_cast<uint32_t*>(base + (w ? k_col : q_col) + (h0 + hh) * head_dim) =
                pack_bf16(x.x * cs - x.y * sn, x.y * cs + x.x * sn);
          }
    }
  }
}

// (cos, sin) table for the RoPE fused into the q/k/v projection's epilogue: tab[pos][j] = half2(sincosf(pos * freqs[j]))
__global__ void rope_table_kernel(const float* __restrict__ freqs, int half, int T, __half2* __restrict__ tab) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * half) return;
  const int pos = i / half, j = i - pos * half;
  float sn, cs;
  sincosf((float)pos * __ldg(freqs + j), &sn, &cs);
  tab[i] = __floats2half2_rn(cs, sn);
}

// RoPE from the fp16 (cos, sin) table (seer_b200_rope_table), in place on the Q and K column blocks of a bf16 [M, ld] buffer:
// one thread = 8 channels (4 rotary pairs, one 16-byte load / store) of one (row, q|k, head); consecutive threads walk a row's
// heads, so a warp touches a few contiguous 16-byte pieces per row and the row's 64 table bytes are shared through L1.  The
// stand-alone pass of the 320-channel level (the q/k/v projection there is epilogue-bound, the fused form costs more).
__global__ void rope_tab_kernel(__nv_bfloat16* __restrict__ qk, int ld, size_t M, int T, int heads, int head_dim, int q_col, int k_col,
                                const __half2* __restrict__ tab) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  // thread -> (row within the block's row group, piece s of the row): no 64-bit division per element (the former flat index
  // cost two of them per 16 bytes and made this pass ALU-bound); the position inside the clip advances incrementally
  const int per_row = 2 * heads * 4;
  const int rpb = blockDim.x / per_row;                      // rows per block pass (host guarantees >= 1)
  if ((int)threadIdx.x >= rpb * per_row) return;
  const int rl = threadIdx.x / per_row, s = threadIdx.x - rl * per_row;
  const int quad = s & 3, head = (s >> 2) % heads, sec = (s >> 2) / heads;
  const int col = (sec ? k_col : q_col) + head * head_dim + 8 * quad;
  const size_t row_step = (size_t)gridDim.x * rpb;
  const int pos_step = (int)(row_step % (size_t)T);
  size_t row = (size_t)blockIdx.x * rpb + rl;
  int pos = (int)(row % (size_t)T);
  for (; row < M; row += row_step) {
    const uint4 tq = __ldg(reinterpret_cast<const uint4*>(tab + (size_t)pos * 16 + 4 * quad));
    uint4* p = reinterpret_cast<uint4*>(qk + row * ld + col);
    uint4 v = *p;
    const uint32_t tw[4] = {tq.x, tq.y, tq.z, tq.w};
    uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 cs = __half22float2(*reinterpret_cast<const __half2*>(&tw[k]));
      const float2 x = unpack_bf16(vw[k]);
      vw[k] = pack_bf16(fmaf(x.x, cs.x, -(x.y * cs.y)), fmaf(x.y, cs.x, x.x * cs.y));
    }
    *p = make_uint4(vw[0], vw[1], vw[2], vw[3]);
    pos += pos_step;
    if (pos >= T) pos -= T;
  }
}

// ---------------------------------------------------------------------------------------------------
// Timestep embedding (diffusers 0.10.2 Timesteps, flip_sin_to_cos): out[b] = [cos(t*f_i) | sin(t*f_i)],
// f_i = exp(-ln(10000) * i / (half - shift)).   Call site unet_3d_condition.py:307.
// ---------------------------------------------------------------------------------------------------
__global__ void timestep_embed_kernel(const float* __restrict__ t, float* __restrict__ out, int B, int dim, float shift,
                                      int flip) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, k = i - b * half;
  const float f = expf(-logf(10000.0f) * (float)k / ((float)half - shift));
  const float a = t[b] * f;
  float sn, cs;
  sincosf(a, &sn, &cs);
  float* o = out + (size_t)b * dim;
  if (flip) { o[k] = cs; o[half + k] = sn; } else { o[k] = sn; o[half + k] = cs; }
}

// ---------------------------------------------------------------------------------------------------
// Small-batch linear: out[b, n] = dot(act(in[b, :]), W[n, :]) + bias[n] (+ add[n]);  fp32 (exact: the fp32-parity path uses it
// too).  Used for time_embedding.linear_{1,2} and all 22 time_emb_proj at once (their weights are concatenated along n:
// 19520 x 1280 fp32 = 100 MB, the only large operand).   resnet.py:190-192, unet_3d_condition.py:308.
// Block = 8 warps x 4 output columns, up to 16 batch rows (the benchmark's CFG batch: the weights are read exactly once).
// The activation rows are staged in shared memory per 512-wide K tile (SiLU applied once there), so one weight float4 from
// HBM meets 16 broadcast-free LDS.128 reused by the warp's 4 columns: 20 loads per 256 FMAs — the former one-column-per-warp
// form issued 17 loads per 64 FMAs and ran at 0.8 TB/s (144 us for the 100 MB panel).
// ---------------------------------------------------------------------------------------------------
constexpr int SL_ROWS = 16;    // batch rows per block pass
constexpr int SL_NPW = 4;      // output columns per warp
constexpr int SL_KT = 512;     // K tile (floats) staged in shared memory: 16 x 512 x 4 B = 32 KB
__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ in, int ldi, const float* __restrict__ W,
                                                           const float* __restrict__ bias, const float* __restrict__ add,
                                                           float* __restrict__ out, int ldo, int B, int N, int K,
                                                           int silu_in, int silu_out) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  __shared__ float4 xs[SL_ROWS][SL_KT / 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = (blockIdx.x * 8 + warp) * SL_NPW;
  const int b0 = blockIdx.y * SL_ROWS;
  float acc[SL_NPW][SL_ROWS];
#pragma unroll
  for (int i = 0; i < SL_NPW; ++i)
#pragma unroll
    for (int r = 0; r < SL_ROWS; ++r) acc[i][r] = 0.f;
  const int n4 = K / 4;
  for (int kt4 = 0; kt4 < n4; kt4 += SL_KT / 4) {
    __syncthreads();                        // the previous tile has been consumed
    for (int idx = threadIdx.x; idx < SL_ROWS * (SL_KT / 4); idx += 256) {
      const int r = idx / (SL_KT / 4), k4 = idx - r * (SL_KT / 4);
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b0 + r < B && kt4 + k4 < n4) {
        x = *reinterpret_cast<const float4*>(in + (size_t)(b0 + r) * ldi + (size_t)(kt4 + k4) * 4);
        if (silu_in) { x.x = silu_f(x.x); x.y = silu_f(x.y); x.z = silu_f(x.z); x.w = silu_f(x.w); }
      }
      xs[r][k4] = x;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SL_KT / 128; ++j) {
      const int k4 = lane + 32 * j;
      float4 w[SL_NPW];
#pragma unroll
      for (int i = 0; i < SL_NPW; ++i)
        w[i] = (n0 + i < N && kt4 + k4 < n4) ? __ldg(reinterpret_cast<const float4*>(W + (size_t)(n0 + i) * K) + kt4 + k4)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < SL_ROWS; ++r) {
        const float4 x = xs[r][k4];
#pragma unroll
        for (int i = 0; i < SL_NPW; ++i) acc[i][r] += (x.x * w[i].x + x.y * w[i].y) + (x.z * w[i].z + x.w * w[i].w);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < SL_NPW; ++i) {
#pragma unroll
    for (int r = 0; r < SL_ROWS; ++r) {
      const float v = warp_sum(acc[i][r]);
      if (lane == 0 && b0 + r < B && n0 + i < N) {
        float o = v + (bias ? bias[n0 + i] : 0.f) + (add ? add[n0 + i] : 0.f);
        if (silu_out) o = silu_f(o);
        out[(size_t)(b0 + r) * ldo + n0 + i] = o;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// conv_in: 3x3, Cin = 4 -> Cout, fp32, input in the reference's (B, Cin, F, H, W) layout, output token-major
// [B*F*H*W, Cout] fp32.   unet_3d_condition.py:94,311.
// ---------------------------------------------------------------------------------------------------
constexpr int CI_PIX = 64;     // pixels per block: the 36 weights a thread keeps in registers are fetched once per 64 pixels
template <bool OUT_BF16>
__global__ void conv_in_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                               void* __restrict__ outv, int B, int Cin, int F, int H, int W, int Cout, float2* __restrict__ col_stats) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  extern __shared__ __align__(16) float s_in[];  // [CI_PIX][Cin*9]
  constexpr int K = 36;  // Cin == 4 (checked on the host)
  const int HW = H * W;
  const size_t npix = (size_t)B * F * HW;
  const size_t p0 = (size_t)blockIdx.x * CI_PIX;
  for (int i = threadIdx.x; i < CI_PIX * K; i += blockDim.x) {
    const int pi = i / K, k = i - pi * K;
    const size_t pix = p0 + pi;
    float v = 0.f;
    if (pix < npix) {
      const int c = k / 9, tap = k - c * 9, ky = tap / 3, kx = tap - ky * 3;
      const int bf = (int)(pix / HW), rem = (int)(pix - (size_t)bf * HW);
      const int b = bf / F, f = bf - b * F;
      const int yy = rem / W + ky - 1, xx = rem % W + kx - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = x[((((size_t)b * Cin + c) * F + f) * H + yy) * W + xx];
    }
    s_in[i] = v;
  }
  __syncthreads();
  for (int co = threadIdx.x; co < Cout; co += blockDim.x) {
    float wr[K];
#pragma unroll
    for (int k = 0; k < K; ++k) wr[k] = w[(size_t)co * K + k];  // weight (Cout, Cin, 3, 3) is already [co][c*9 + tap]
    const float bz = bias[co];
    float ssum = 0.f, ssq = 0.f;             // per-(32-pixel slab, channel) sums for the GroupNorms that consume this tensor
    for (int pi = 0; pi < CI_PIX; ++pi) {
      if (p0 + pi >= npix) break;
      float acc = bz;
      const float4* sp = reinterpret_cast<const float4*>(s_in + pi * K);     // broadcast reads, 16 bytes each
#pragma unroll
      for (int k4 = 0; k4 < K / 4; ++k4) {
        const float4 v = sp[k4];
        acc += v.x * wr[4 * k4];
        acc += v.y * wr[4 * k4 + 1];
        acc += v.z * wr[4 * k4 + 2];
        acc += v.w * wr[4 * k4 + 3];
      }
      if (OUT_BF16) reinterpret_cast<__nv_bfloat16*>(outv)[(p0 + pi) * Cout + co] = __float2bfloat16_rn(acc);
      else reinterpret_cast<float*>(outv)[(p0 + pi) * Cout + co] = acc;
      ssum += acc;                           // statistics of the fp32 values, whatever dtype is stored
      ssq = fmaf(acc, acc, ssq);
      if (col_stats && ((pi & 31) == 31)) {
        col_stats[((p0 + pi) >> 5) * Cout + co] = make_float2(ssum, ssq);     // same layout as SeerGemmDesc::col_stats
        ssum = ssq = 0.f;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// conv_in on the tensor cores: im2col of the 4-channel latent (B, 4, F, H, W) fp32 -> bf16 [B*F*H*W, 64], column k = c * 9 + tap
// for k < 36 (the order of the (Cout, Cin, 3, 3) weight), zeros above — one 64-wide k-block of the tcgen05 GEMM, whose epilogue
// then emits the bias, the bf16 / fp32 stream tensor and the GroupNorm column sums like every other producer.  The fp32 SIMT
// kernel above is FP32-pipe bound at ~380 us for 8 clips; this pass writes 17 MB and the GEMM streams the output once.
// One thread = one pixel x 8 columns (one 16-byte store).
// ---------------------------------------------------------------------------------------------------
__global__ void conv_in_im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ a, int B, int F, int H, int W) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  const int HW = H * W;
  const size_t total = (size_t)B * F * HW * 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(i & 7);
    const size_t pix = i >> 3;
    const int bf = (int)(pix / HW), rem = (int)(pix - (size_t)bf * HW);
    const int b = bf / F, f = bf - b * F;
    const int y = rem / W, xx0 = rem - y * W;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = g * 8 + j;
      v[j] = 0.f;
      if (k < 36) {
        const int c = k / 9, tap = k - c * 9, ky = tap / 3, kx = tap - ky * 3;
        const int yy = y + ky - 1, xx = xx0 + kx - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v[j] = x[((((size_t)b * 4 + c) * F + f) * H + yy) * W + xx];
      }
    }
    uint4 o;
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(a + pix * 64 + g * 8) = o;
  }
}

// ---------------------------------------------------------------------------------------------------
// conv_out: 3x3, Cin -> <= 4, fp32, input token-major [B*F*H*W, Cin] fp32 (already GroupNorm+SiLU'd), output in the
// reference's (B, Cout, F, H, W) layout.   unet_3d_condition.py:205,370.  Weights pre-packed as wp[co][tap][Cin].
// One block = one output image row.  Per 64-channel chunk the three input rows (and the chunk's weights) are staged in
// shared memory once — every input pixel is read from L2 by the 3 row-blocks that need it instead of by 9 taps x 1 warp —
// and a half-warp (16 lanes = the chunk's 16 channel quads) accumulates 4 adjacent output pixels x 4 output channels,
// so a weight quad fetched from smem feeds 4 pixels.  Deterministic: fixed summation order, no atomics.
// ---------------------------------------------------------------------------------------------------
constexpr int CO_CK = 64;        // channels per chunk
constexpr int CO_PIX = 4;        // output pixels per half-warp pass
constexpr int CO_THREADS = 128;
constexpr int CO_SEG = 64;       // output pixels of one image row per block (wider rows are split into segments: blockIdx.y)
__global__ void __launch_bounds__(CO_THREADS) conv_out_kernel(const float* __restrict__ x, const float* __restrict__ wp,
                                                              const float* __restrict__ bias, float* __restrict__ out, int B,
                                                              int Cin, int F, int H, int W, int Cout) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  extern __shared__ __align__(16) float co_smem[];
  const int seg0 = blockIdx.y * CO_SEG;
  const int segw = min(CO_SEG, W - seg0);
  const int sw_ = segw + 2;                  // staged pixels per row: the segment plus a one-pixel halo on each side
  float* sx = co_smem;                       // [3][segw + 2][CO_CK]
  float* sw = sx + 3 * (CO_SEG + 2) * CO_CK; // [9][4][CO_CK]
  const int bf = blockIdx.x / H, y = blockIdx.x - bf * H;
  const int tid = threadIdx.x;
  const int hw = tid >> 4, l = tid & 15;     // half-warp index (0..7), channel quad
  const int n_groups = (segw + CO_PIX - 1) / CO_PIX;
  constexpr int MAX_G = 2;                   // pixel groups per half-warp (segment <= 64 pixels)
  float acc[MAX_G][CO_PIX][4];
#pragma unroll
  for (int g = 0; g < MAX_G; ++g)
#pragma unroll
    for (int p = 0; p < CO_PIX; ++p) acc[g][p][0] = acc[g][p][1] = acc[g][p][2] = acc[g][p][3] = 0.f;

  for (int c0 = 0; c0 < Cin; c0 += CO_CK) {
    __syncthreads();                         // previous chunk consumed
    for (int i = tid; i < 3 * sw_ * (CO_CK / 4); i += CO_THREADS) {
      const int q = i % (CO_CK / 4), px = (i / (CO_CK / 4)) % sw_, rr = i / ((CO_CK / 4) * sw_);
      const int yy = y + rr - 1, xx = seg0 + px - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(reinterpret_cast<const float4*>(x + (((size_t)bf * H + yy) * W + xx) * Cin + c0) + q);
      reinterpret_cast<float4*>(sx)[(rr * sw_ + px) * (CO_CK / 4) + q] = v;
    }
    for (int i = tid; i < 9 * 4 * (CO_CK / 4); i += CO_THREADS) {
      const int q = i % (CO_CK / 4), co = (i / (CO_CK / 4)) % 4, tap = i / ((CO_CK / 4) * 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (co < Cout) v = __ldg(reinterpret_cast<const float4*>(wp + ((size_t)co * 9 + tap) * Cin + c0) + q);
      reinterpret_cast<float4*>(sw)[i] = v;
    }
    __syncthreads();
#pragma unroll
    for (int g = 0; g < MAX_G; ++g) {
      const int grp = hw + g * 8;
      if (grp >= n_groups) break;
      const int x0 = grp * CO_PIX;
      for (int tap = 0; tap < 9; ++tap) {
        const int rr = tap / 3, dx = tap % 3;          // staged pixel of output pixel x and tap dx: x + dx (halo offset folded in)
        float4 wv[4];
#pragma unroll
        for (int co = 0; co < 4; ++co) wv[co] = reinterpret_cast<const float4*>(sw)[(tap * 4 + co) * (CO_CK / 4) + l];
#pragma unroll
        for (int p = 0; p < CO_PIX; ++p) {
          const int xs = x0 + p + dx;
          if (xs >= sw_) continue;                     // ragged last pixel group of the segment
          const float4 v = reinterpret_cast<const float4*>(sx)[(rr * sw_ + xs) * (CO_CK / 4) + l];
#pragma unroll
          for (int co = 0; co < 4; ++co)
            acc[g][p][co] += (v.x * wv[co].x + v.y * wv[co].y) + (v.z * wv[co].z + v.w * wv[co].w);
        }
      }
    }
  }
  const int b = bf / F, f = bf - b * F;
#pragma unroll
  for (int g = 0; g < MAX_G; ++g) {
    const int grp = hw + g * 8;
    if (grp >= n_groups) break;
#pragma unroll
    for (int p = 0; p < CO_PIX; ++p) {
#pragma unroll
      for (int co = 0; co < 4; ++co) {
        float v = acc[g][p][co];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);    // within the 16-lane half
        const int xq = grp * CO_PIX + p;
        if (l == 0 && co < Cout && xq < segw) out[((((size_t)b * Cout + co) * F + f) * H + y) * W + seg0 + xq] = v + bias[co];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Row softmax for the VAE's single-head d = 512 attention (diffusers 0.10.2 AttentionBlock, call site
// utils/ddim_sampling_utils.py:39 through vae.decode): P[r, :] = softmax(scale * S[r, :]) as bf16, one warp per row.
// The scores come from a tcgen05 GEMM (Q K^T) and P feeds another one (P V): at d = 512 the two contractions dominate.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ S, int lds_, long long rows, int L, float scale,
                                                           __nv_bfloat16* __restrict__ P, int ldp) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(S + row * lds_);
  const int n4 = L / 4;
  const float sl2 = scale * 1.4426950408889634f;
  float mx = -INFINITY;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = src[i];
    mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = src[i];
    sum += (exp2f((v.x - mx) * sl2) + exp2f((v.y - mx) * sl2)) + (exp2f((v.z - mx) * sl2) + exp2f((v.w - mx) * sl2));
  }
  const float inv = 1.0f / warp_sum(sum);
  uint2* dst = reinterpret_cast<uint2*>(P + row * ldp);
  for (int i = lane; i < n4; i += 32) {
    const float4 v = src[i];
    uint2 o;
    o.x = pack_bf16(exp2f((v.x - mx) * sl2) * inv, exp2f((v.y - mx) * sl2) * inv);
    o.y = pack_bf16(exp2f((v.z - mx) * sl2) * inv, exp2f((v.w - mx) * sl2) * inv);
    dst[i] = o;
  }
}

// ---------------------------------------------------------------------------------------------------
// Token-major [B*F*HW, ld] fp32 (first C columns) -> the reference's (B, C, F, H, W) layout: the 4-channel epsilon that leaves the
// UNet when conv_out runs as a tensor-core implicit GEMM with its output channels zero-padded to one 64-column tile.
// ---------------------------------------------------------------------------------------------------
__global__ void tokens_to_nchw_kernel(const float* __restrict__ x, int ld, float* __restrict__ out, int B, int C, int F, int HW) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  const size_t total = (size_t)B * C * F * HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    size_t r = i / HW;
    const int f = (int)(r % F); r /= F;
    const int c = (int)(r % C);
    const int b = (int)(r / C);
    out[i] = x[(((size_t)b * F + f) * HW + p) * ld + c];
  }
}

// ---------------------------------------------------------------------------------------------------
// nearest 2x upsample, fp32 [n_img, H, W, C] -> bf16 [n_img, 2H, 2W, C]   (resnet.py:52, conv input operand)
// ---------------------------------------------------------------------------------------------------
__global__ void upsample2x_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int n_img, int H, int W, int C) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  const int c8n = C / 8;
  const size_t total = (size_t)n_img * 4 * H * W * c8n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8n) * 8;
    size_t r = i / c8n;
    const int ox = (int)(r % (2 * W)); r /= (2 * W);
    const int oy = (int)(r % (2 * H));
    const int img = (int)(r / (2 * H));
    const float* src = x + (((size_t)img * H + oy / 2) * W + ox / 2) * C + c;
    const float4 a = *reinterpret_cast<const float4*>(src);
    const float4 b = *reinterpret_cast<const float4*>(src + 4);
    uint4 o;
    o.x = pack_bf16(a.x, a.y); o.y = pack_bf16(a.z, a.w); o.z = pack_bf16(b.x, b.y); o.w = pack_bf16(b.z, b.w);
    *reinterpret_cast<uint4*>(y + i * 8) = o;
  }
}

// ---------------------------------------------------------------------------------------------------
// im2col for pad-1 3x3 convs, stride 1 or 2: [n_img, H, W, C] (fp32 or bf16) -> bf16 [n_img*Ho*Wo, 9*C],
// K order = [C/64][tap][64] (matches the packed conv weight, packing.pack_conv3x3).  Stride 2: Downsample3D (resnet.py:95-104).
// ---------------------------------------------------------------------------------------------------
template <bool IN_BF16>
__global__ void im2col3x3_kernel(const void* __restrict__ xin, __nv_bfloat16* __restrict__ y, int n_img, int H, int W, int C,
                                 int stride) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  const int nchunk = C / 64;
  const int Ho = H / stride, Wo = W / stride;
  const size_t total = (size_t)n_img * Ho * Wo * nchunk * 72;     // 72 = 9 taps x 8 groups of 8 channels
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int g8 = (int)(r % 8); r /= 8;
    const int tap = (int)(r % 9); r /= 9;
    const int chunk = (int)(r % nchunk); r /= nchunk;
    const int c = chunk * 64 + g8 * 8;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho);
    const int img = (int)(r / Ho);
    const int yy = stride * oy + tap / 3 - 1, xx = stride * ox + tap % 3 - 1;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const size_t off = (((size_t)img * H + yy) * W + xx) * C + c;
      if (IN_BF16) {
        o = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(xin) + off);
      } else {
        const float* src = reinterpret_cast<const float*>(xin) + off;
        const float4 a = *reinterpret_cast<const float4*>(src);
        const float4 b = *reinterpret_cast<const float4*>(src + 4);
        o.x = pack_bf16(a.x, a.y); o.y = pack_bf16(a.z, a.w); o.z = pack_bf16(b.x, b.y); o.w = pack_bf16(b.z, b.w);
      }
    }
    *reinterpret_cast<uint4*>(y + i * 8) = o;
  }
}

// ---------------------------------------------------------------------------------------------------
// CFG combine + DDIM update (eta = 0), fp32, in the reference's exact operation order with explicitly
// rounded (non-contracted) arithmetic so the result is bit-identical to the PyTorch expressions at
// /root/reference/ldm/models/diffusion/ddim_video.py:209-211, 229-237:
//   e       = e_u + s * (e_c - e_u)                      (frames >= cond_f only)
//   pred_x0 = (x - sqrt(1 - a_t) * e) / sqrt(a_t)
//   x_prev  = sqrt(a_prev) * pred_x0 + sqrt(1 - a_prev) * e
// eps is (2b | b, C, F, H, W) with F = cond_f + F2; x / x_prev / pred_x0 are (b, C, F2, H, W).
// ---------------------------------------------------------------------------------------------------
__global__ void cfg_ddim_kernel(const float* __restrict__ eps, const float* __restrict__ x, float* __restrict__ x_prev,
                                float* __restrict__ pred_x0, int b, int C, int F2, int cond_f, int HW, int use_cfg, float scale,
                                float sqrt_one_minus_at, float sqrt_at, float sqrt_a_prev, float dir_coef) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  const size_t total = (size_t)b * C * F2 * HW;
  const int F = F2 + cond_f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    size_t r = i / HW;
    const int f = (int)(r % F2); r /= F2;
    const int c = (int)(r % C);
    const int bi = (int)(r / C);
    const size_t eo = (((size_t)bi * C + c) * F + (f + cond_f)) * HW + p;
    float e;
    if (use_cfg) {
      const float eu = eps[eo];
      const float ec = eps[eo + (size_t)b * C * F * HW];
      e = __fadd_rn(eu, __fmul_rn(scale, __fsub_rn(ec, eu)));
    } else {
      e = eps[eo];
    }
    const float xv = x[i];
    const float p0 = __fdiv_rn(__fsub_rn(xv, __fmul_rn(sqrt_one_minus_at, e)), sqrt_at);
    const float xp = __fadd_rn(__fmul_rn(sqrt_a_prev, p0), __fmul_rn(dir_coef, e));
    pred_x0[i] = p0;
    x_prev[i] = xp;
  }
}

// ---------------------------------------------------------------------------------------------------
// CFG-branch split over NVLink peer memory (parallel.py): the two ranks of a pair each hold ONE branch of
// the noise prediction.  One kernel per rank and step does the exchange and the update:
//   phase 1  every CTA PUSHES its part of this rank's branch (frames >= cond_f only: (b, C, F2, HW) fp32) into the
//            partner's receive slot with plain stores through the peer mapping (posted writes over NVLink);
//   signal   fence.sys per thread, one device-scope counter per CTA; the last CTA to finish releases `seq` into the
//            partner's flag word (st.release.sys);
//   wait     thread 0 of every CTA acquires this rank's own flag (ld.acquire.sys) until it reaches `seq`;
//   phase 2  the update of cfg_ddim_kernel, bit for bit: e_u / e_c come from the local tensor and the receive slot
//            (ld.global.cg: the slot is written by the peer, never through this SM's L1).
// Both ranks compute the same x_prev / pred_x0 redundantly, so nothing flows back.  The host alternates two receive
// slots / flag words by step parity: the partner overwrites slot s two steps later, after it has seen this rank's flag
// of the step in between, which this rank only sends after its reads of slot s have completed (stream order).
// Every CTA spins, so the grid must be co-resident: the launcher caps it at one CTA per SM.  The wait is bounded
// (~20 s of %globaltimer, then trap): a lost partner fails the launch instead of hanging the GPU.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(256)
cfg_ddim_p2p_kernel(const float* __restrict__ eps_local, int branch, float* __restrict__ peer_recv, const float* local_recv,
                    unsigned* peer_flag, const unsigned* local_flag, unsigned* counter, unsigned seq,
                    const float* __restrict__ x, float* __restrict__ x_prev, float* __restrict__ pred_x0, int b, int C, int F2,
                    int cond_f, int HW, float scale, float sqrt_one_minus_at, float sqrt_at, float sqrt_a_prev, float dir_coef) {
  pdl_wait();
  const size_t total = (size_t)b * C * F2 * HW;
  const size_t fstride = (size_t)F2 * HW, estride = (size_t)(F2 + cond_f) * HW, eoff = (size_t)cond_f * HW;
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
  for (size_t i = i0; i < total; i += step) {
    const size_t bc = i / fstride;
    peer_recv[i] = eps_local[bc * estride + eoff + (i - bc * fstride)];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();                         // cumulative over the CTA's pushes (ordered before it by the barrier)
    if (atomicAdd(counter, 1u) == gridDim.x - 1) {
      atomicExch(counter, 0u);                      // next launch on this stream starts from zero
      __threadfence_system();
      st_release_sys(peer_flag, seq);
    }
    const unsigned long long t0 = globaltimer_ns();
    while ((int)(ld_acquire_sys(local_flag) - seq) < 0) {
      __nanosleep(64);
      if (globaltimer_ns() - t0 > 20000000000ull) __trap();
    }
  }
  __syncthreads();
  for (size_t i = i0; i < total; i += step) {
    const size_t bc = i / fstride;
    const float mine = eps_local[bc * estride + eoff + (i - bc * fstride)];
    const float theirs = __ldcg(local_recv + i);
    const float eu = branch == 0 ? mine : theirs;
    const float ec = branch == 0 ? theirs : mine;
    const float e = __fadd_rn(eu, __fmul_rn(scale, __fsub_rn(ec, eu)));
    const float xv = x[i];
    const float p0 = __fdiv_rn(__fsub_rn(xv, __fmul_rn(sqrt_one_minus_at, e)), sqrt_at);
    const float xp = __fadd_rn(__fmul_rn(sqrt_a_prev, p0), __fmul_rn(dir_coef, e));
    pred_x0[i] = p0;
    x_prev[i] = xp;
  }
}

static inline int grid_for(size_t n, int threads) {
  size_t b = (n + threads - 1) / threads;
  const size_t cap = (size_t)148 * 32;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace seer

using namespace seer;

extern "C" int seer_b200_cast_f32_to_bf16(const float* x, void* y, long long n, void* stream) {
  SEER_CHECK_ARG(x && y && n > 0 && n % 8 == 0);
  { cudaError_t le__ = launch_pdl(cast_bf16_kernel, grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream, x, (__nv_bfloat16*)y, (size_t)n / 8); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_rope_inplace(void* qk_bf16, int ld, int M, int tokens_per_clip, int heads, int head_dim, int q_col,
                                      int k_col, const float* freqs, int n_freqs, void* stream) {
  SEER_CHECK_ARG(qk_bf16 && freqs && M > 0 && tokens_per_clip > 0 && M % tokens_per_clip == 0);
  SEER_CHECK_ARG(2 * n_freqs <= head_dim && ld % 2 == 0 && q_col % 2 == 0 && k_col % 2 == 0 && head_dim % 2 == 0);
  { cudaError_t le__ = launch_pdl(rope_kernel, grid_for((size_t)M * n_freqs, 256), 256, 0, (cudaStream_t)stream, (__nv_bfloat16*)qk_bf16, ld, M, tokens_per_clip,
                                                                                   heads, head_dim, q_col, k_col, freqs, n_freqs); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_rope_table(const float* freqs, int n_freqs, int T, void* out, void* stream) {
  SEER_CHECK_ARG(freqs && out && n_freqs > 0 && T > 0);
  { cudaError_t le__ = launch_pdl(rope_table_kernel, ceil_div(T * n_freqs, 256), 256, 0, (cudaStream_t)stream, freqs, n_freqs, T, (__half2*)out); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_rope_apply_table(void* qk_bf16, int ld, long long M, int tokens_per_clip, int heads, int head_dim, int q_col,
                                          int k_col, const void* tab, void* stream) {
  SEER_CHECK_ARG(qk_bf16 && tab && M > 0 && tokens_per_clip > 0 && heads > 0 && head_dim >= 32);
  SEER_CHECK_ARG(ld % 8 == 0 && q_col % 8 == 0 && k_col % 8 == 0 && head_dim % 8 == 0 && ((uintptr_t)qk_bf16 % 16) == 0);
  const int per_row = 2 * heads * 4;
  SEER_CHECK_ARG(per_row <= 1024);
  const int threads = per_row <= 256 ? (256 / per_row) * per_row : per_row;
  const size_t blocks = ((size_t)M + threads / per_row - 1) / (threads / per_row);
  const size_t cap = (size_t)148 * 32;
  { cudaError_t le__ = launch_pdl(rope_tab_kernel, (unsigned)(blocks < cap ? blocks : cap), threads, 0, (cudaStream_t)stream, (__nv_bfloat16*)qk_bf16, ld, (size_t)M,
                                  tokens_per_clip, heads, head_dim, q_col, k_col, (const __half2*)tab); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_timestep_embedding(const float* t, float* out, int B, int dim, float shift, int flip_sin_to_cos,
                                            void* stream) {
  SEER_CHECK_ARG(t && out && B > 0 && dim > 0 && dim % 2 == 0);
  { cudaError_t le__ = launch_pdl(timestep_embed_kernel, ceil_div(B * dim / 2, 128), 128, 0, (cudaStream_t)stream, t, out, B, dim, shift, flip_sin_to_cos); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_small_linear(const float* in, int ldi, const float* W, const float* bias, const float* add, float* out,
                                      int ldo, int B, int N, int K, int silu_in, int silu_out, void* stream) {
  SEER_CHECK_ARG(in && W && out && B > 0 && N > 0 && K > 0 && K % 4 == 0 && ldi % 4 == 0);
  dim3 grid(ceil_div(N, 8 * SL_NPW), ceil_div(B, SL_ROWS));
  { cudaError_t le__ = launch_pdl(small_linear_kernel, grid, 256, 0, (cudaStream_t)stream, in, ldi, W, bias, add, out, ldo, B, N, K, silu_in, silu_out); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_conv_in(const float* x, const float* w, const float* bias, float* out, int B, int Cin, int F, int H,
                                 int W, int Cout, void* stream) {
  return seer_b200_conv_in_stats(x, w, bias, out, nullptr, B, Cin, F, H, W, Cout, stream);
}

extern "C" int seer_b200_conv_in_ex(const float* x, const float* w, const float* bias, void* out, int out_is_bf16, float* col_stats,
                                    int B, int Cin, int F, int H, int W, int Cout, void* stream) {
  SEER_CHECK_ARG(x && w && bias && out && Cin == 4);
  const size_t npix = (size_t)B * F * H * W;
  SEER_CHECK_ARG(!col_stats || npix % 32 == 0);
  const int threads = Cout >= 320 ? 320 : ((Cout + 31) / 32) * 32;
  const unsigned grid = (unsigned)((npix + CI_PIX - 1) / CI_PIX);
  const size_t smem = CI_PIX * Cin * 9 * sizeof(float);
  cudaError_t le__;
  if (out_is_bf16)
    le__ = launch_pdl(conv_in_kernel<true>, grid, threads, smem, (cudaStream_t)stream, x, w, bias, out, B, Cin, F, H, W, Cout, (float2*)col_stats);
  else
    le__ = launch_pdl(conv_in_kernel<false>, grid, threads, smem, (cudaStream_t)stream, x, w, bias, out, B, Cin, F, H, W, Cout, (float2*)col_stats);
  if (le__ != cudaSuccess) return (int)le__;
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_conv_in_im2col(const float* x, void* a_bf16, int B, int Cin, int F, int H, int W, void* stream) {
  SEER_CHECK_ARG(x && a_bf16 && Cin == 4 && B > 0 && F > 0 && H > 0 && W > 0);
  const size_t total = (size_t)B * F * H * W * 8;
  { cudaError_t le__ = launch_pdl(conv_in_im2col_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, x, (__nv_bfloat16*)a_bf16, B, F, H, W); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_conv_in_stats(const float* x, const float* w, const float* bias, float* out, float* col_stats, int B, int Cin,
                                       int F, int H, int W, int Cout, void* stream) {
  return seer_b200_conv_in_ex(x, w, bias, out, 0, col_stats, B, Cin, F, H, W, Cout, stream);
}

extern "C" int seer_b200_conv_out(const float* x, const float* w_packed, const float* bias, float* out, int B, int Cin, int F,
                                  int H, int W, int Cout, void* stream) {
  SEER_CHECK_ARG(x && w_packed && bias && out && Cout <= 4 && Cin % CO_CK == 0 && W >= 1);
  const size_t smem = (size_t)(3 * (CO_SEG + 2) * CO_CK + 9 * 4 * CO_CK) * sizeof(float);
  static SmemAttrOnce smem_attr;
  if (smem > 48 * 1024) { cudaError_t e = smem_attr.ensure(conv_out_kernel, (int)smem); if (e != cudaSuccess) return (int)e; }
  { cudaError_t le__ = launch_pdl(conv_out_kernel, dim3((unsigned)(B * F * H), (unsigned)ceil_div(W, CO_SEG)), CO_THREADS, smem, (cudaStream_t)stream, x, w_packed, bias, out, B, Cin, F, H, W, Cout); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_tokens_to_nchw(const float* x, int ld, float* out, int B, int C, int F, int HW, void* stream) {
  SEER_CHECK_ARG(x && out && B > 0 && C > 0 && F > 0 && HW > 0 && ld >= C);
  { cudaError_t le__ = launch_pdl(tokens_to_nchw_kernel, grid_for((size_t)B * C * F * HW, 256), 256, 0, (cudaStream_t)stream, x, ld, out, B, C, F, HW); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_softmax_rows(const float* S, int lds, long long rows, int L, float scale, void* P_bf16, int ldp, void* stream) {
  SEER_CHECK_ARG(S && P_bf16 && rows > 0 && L > 0 && L % 4 == 0 && lds % 4 == 0 && ldp % 4 == 0);
  { cudaError_t le__ = launch_pdl(softmax_rows_kernel, (unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream, S, lds, rows, L, scale, (__nv_bfloat16*)P_bf16, ldp); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_upsample2x_to_bf16(const float* x, void* y, int n_img, int H, int W, int C, void* stream) {
  SEER_CHECK_ARG(x && y && C % 8 == 0);
  const size_t total = (size_t)n_img * 4 * H * W * (C / 8);
  { cudaError_t le__ = launch_pdl(upsample2x_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, x, (__nv_bfloat16*)y, n_img, H, W, C); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_im2col3x3_to_bf16(const void* x, int in_is_bf16, void* y, int n_img, int H, int W, int C, int stride,
                                           void* stream) {
  SEER_CHECK_ARG(x && y && C % 64 == 0 && (stride == 1 || stride == 2) && H % stride == 0 && W % stride == 0);
  const size_t total = (size_t)n_img * (H / stride) * (W / stride) * 9 * (C / 8);
  if (in_is_bf16)
    { cudaError_t le__ = launch_pdl(im2col3x3_kernel<true>, grid_for(total, 256), 256, 0, (cudaStream_t)stream, x, (__nv_bfloat16*)y, n_img, H, W, C, stride); if (le__ != cudaSuccess) return (int)le__; }
  else
    { cudaError_t le__ = launch_pdl(im2col3x3_kernel<false>, grid_for(total, 256), 256, 0, (cudaStream_t)stream, x, (__nv_bfloat16*)y, n_img, H, W, C, stride); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_cfg_ddim_update_p2p(const float* eps_local, int branch, float* peer_recv, const float* local_recv,
                                             unsigned* peer_flag, const unsigned* local_flag, unsigned* counter, unsigned seq,
                                             const float* x, float* x_prev, float* pred_x0, int b, int C, int F2, int cond_f, int HW,
                                             float scale, float sqrt_one_minus_at, float sqrt_at, float sqrt_a_prev, float dir_coef,
                                             void* stream) {
  SEER_CHECK_ARG(eps_local && peer_recv && local_recv && peer_flag && local_flag && counter && x && x_prev && pred_x0);
  SEER_CHECK_ARG((branch == 0 || branch == 1) && b > 0 && C > 0 && F2 > 0 && HW > 0 && cond_f >= 0 && seq != 0);
  const size_t total = (size_t)b * C * F2 * HW;
  const size_t blocks = (total + 255) / 256;
  const int grid = (int)(blocks < 148 ? blocks : 148);      // every CTA spins on the partner's flag: keep the grid co-resident
  { cudaError_t le__ = launch_pdl(cfg_ddim_p2p_kernel, grid, 256, 0, (cudaStream_t)stream, eps_local, branch, peer_recv, local_recv, peer_flag,
                                  local_flag, counter, seq, x, x_prev, pred_x0, b, C, F2, cond_f, HW, scale, sqrt_one_minus_at, sqrt_at,
                                  sqrt_a_prev, dir_coef); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

extern "C" int seer_b200_cfg_ddim_update(const float* eps, const float* x, float* x_prev, float* pred_x0, int b, int C, int F2,
                                         int cond_f, int HW, int use_cfg, float scale, float sqrt_one_minus_at, float sqrt_at,
                                         float sqrt_a_prev, float dir_coef, void* stream) {
  SEER_CHECK_ARG(eps && x && x_prev && pred_x0 && b > 0 && C > 0 && F2 > 0 && HW > 0 && cond_f >= 0);
  const size_t total = (size_t)b * C * F2 * HW;
  { cudaError_t le__ = launch_pdl(cfg_ddim_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, eps, x, x_prev, pred_x0, b, C, F2, cond_f, HW, use_cfg,
                                                                          scale, sqrt_one_minus_at, sqrt_at, sqrt_a_prev, dir_coef); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}
