// Fused softmax attention for the three attention shapes of the Seer UNet, reading Q/K/V straight out of the
// token-major projection buffers (no head split / window partition copies) and writing O token-major:
//
//   mode SEER_ATTN_SPATIAL : per-frame self attention, sequence = the h*w tokens of one frame
//                            (/root/reference/seer/models/attention.py:310-311, 512-554)
//   mode SEER_ATTN_CROSS   : per-frame cross attention against that frame's Lk (=77) text tokens
//                            (attention.py:313-322)
//   mode SEER_ATTN_SCTA    : spatial-causal temporal attention — sequence = (frame, y, x) inside one ws x ws
//                            window (or the whole clip when ws == 0), lower-triangular causal
//                            (attention.py:632-703; window order :42-53)
//
// The reference's reshape_heads_to_batch_dim / window_partition / window_reverse (attention.py:492-504, 42-69)
// are pure index permutations; here they are the row-index functions q_row()/kv_row() below, applied while
// gathering 16-byte chunks with cp.async, so they cost no memory traffic.
//
// v1 math path: flash-attention-2 style online softmax with mma.sync.m16n8k16 (bf16 in, fp32 accumulate),
// 64 queries x 64 keys per step, K/V double-buffered through cp.async.  (The tcgen05 port of this kernel is
// the next optimisation step; the GEMM/conv kernels carry 94 % of the FLOPs.)
#include "common.cuh"
#include "seer_b200.h"

namespace seer {

struct AttnParams {
  const __nv_bfloat16* q; int ldq;
  const __nv_bfloat16* k; int ldk;
  const __nv_bfloat16* v; int ldv;
  __nv_bfloat16* o; int ldo;
  int mode;
  int heads;
  int Lq, Lk;       // sequence lengths per problem
  int n_outer;      // spatial/cross: number of frames (b*f); scta: batch b
  // scta geometry
  int F, H, W, ws, nwx, nwin;
  float scale_log2; // d^-0.5 * log2(e)
  int causal;
};

__device__ __forceinline__ int scta_row(const AttnParams& p, int b, int win, int s) {
  if (p.ws == 0) return b * p.F * p.H * p.W + s;
  const int ws2 = p.ws * p.ws;
  const int f = s / ws2;
  const int r = s - f * ws2;
  const int iy = r / p.ws, ix = r - iy * p.ws;
  const int wy = win / p.nwx, wx = win - wy * p.nwx;
  return ((b * p.F + f) * p.H + wy * p.ws + iy) * p.W + wx * p.ws + ix;
}

constexpr int ATT_BM = 64;
constexpr int ATT_BN = 64;
constexpr int ATT_THREADS = 128;

template <int D>
struct AttnCfg {
  static constexpr int DP = (D + 15) / 16 * 16;  // K-dim of QK^T padded to the MMA k=16
  static constexpr int LDS = DP + 8;             // smem row stride (elements): +16 B keeps ldmatrix conflict-free
  static constexpr int CHUNKS = D / 8;           // 16-byte chunks of real data per row
  static constexpr int PCHUNKS = DP / 8;
  static constexpr int TILE_ELEMS = 64 * LDS;
  static constexpr int SMEM_BYTES = 5 * TILE_ELEMS * 2;  // Q + 2x(K,V)
};

template <int D>
__device__ __forceinline__ void load_tile(__nv_bfloat16* s, const __nv_bfloat16* g, int ld, int col0, const int* rows,
                                          int tid) {
  using C = AttnCfg<D>;
  // 64 rows x PCHUNKS chunks; padded chunks and invalid rows are zero-filled.
  for (int i = tid; i < 64 * C::PCHUNKS; i += ATT_THREADS) {
    const int r = i / C::PCHUNKS, c = i - r * C::PCHUNKS;
    const int grow = rows[r];
    const bool ok = (grow >= 0) && (c < C::CHUNKS);
    const __nv_bfloat16* src = ok ? g + (size_t)grow * ld + col0 + c * 8 : g;
    cp_async_16(s + r * C::LDS + c * 8, src, ok);
  }
}

template <int D>
__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(const AttnParams p) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  using C = AttnCfg<D>;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sK = sQ + C::TILE_ELEMS;           // 2 buffers
  __nv_bfloat16* sV = sK + 2 * C::TILE_ELEMS;       // 2 buffers
  __shared__ int q_rows[64];
  __shared__ int kv_rows[2][64];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * ATT_BM;
  const int head = blockIdx.y % p.heads;
  int outer = blockIdx.y / p.heads;          // spatial/cross: frame index; scta: b * nwin + win
  int b = 0, win = 0;
  if (p.mode == SEER_ATTN_SCTA) { b = outer / p.nwin; win = outer - b * p.nwin; }
  const int col0 = head * D;

  auto qrow = [&](int s) -> int {
    if (s >= p.Lq) return -1;
    return p.mode == SEER_ATTN_SCTA ? scta_row(p, b, win, s) : outer * p.Lq + s;
  };
  auto kvrow = [&](int s) -> int {
    if (s >= p.Lk) return -1;
    return p.mode == SEER_ATTN_SCTA ? scta_row(p, b, win, s) : outer * p.Lk + s;
  };

  if (tid < 64) q_rows[tid] = qrow(q0 + tid);
  int n_kv_tiles = ceil_div(p.Lk, ATT_BN);
  if (p.causal) n_kv_tiles = min(n_kv_tiles, ceil_div(q0 + ATT_BM, ATT_BN));
  if (tid >= 64) kv_rows[0][tid - 64] = kvrow(tid - 64);
  __syncthreads();
  load_tile<D>(sQ, p.q, p.ldq, col0, q_rows, tid);
  load_tile<D>(sK, p.k, p.ldk, col0, kv_rows[0], tid);
  load_tile<D>(sV, p.v, p.ldv, col0, kv_rows[0], tid);
  cp_async_commit();

  float o_acc[C::DP / 8][4];
#pragma unroll
  for (int i = 0; i < C::DP / 8; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  const int qi0 = q0 + warp * 16 + (lane >> 2);  // this thread's two query rows: qi0, qi0 + 8

  for (int t = 0; t < n_kv_tiles; ++t) {
    const int buf = t & 1;
    // prefetch next K/V tile
    if (t + 1 < n_kv_tiles) {
      if (tid < 64) kv_rows[buf ^ 1][tid] = kvrow((t + 1) * ATT_BN + tid);
    }
    __syncthreads();  // kv_rows[next] visible; everyone finished reading buffers [buf^1] from iteration t-1
    if (t + 1 < n_kv_tiles) {
      load_tile<D>(sK + (buf ^ 1) * C::TILE_ELEMS, p.k, p.ldk, col0, kv_rows[buf ^ 1], tid);
      load_tile<D>(sV + (buf ^ 1) * C::TILE_ELEMS, p.v, p.ldv, col0, kv_rows[buf ^ 1], tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();  // tile t (and Q) landed for all threads

    const __nv_bfloat16* kt = sK + buf * C::TILE_ELEMS;
    const __nv_bfloat16* vt = sV + buf * C::TILE_ELEMS;

    // ---- S = Q K^T (16 x 64 per warp) ----
    float s_acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s_acc[i][0] = s_acc[i][1] = s_acc[i][2] = s_acc[i][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < C::DP / 16; ++kk) {
      uint32_t a[4];
      ldmatrix_x4(a, sQ + (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * C::LDS + kk * 16 + (lane >> 4) * 8);
#pragma unroll
      for (int nn = 0; nn < 4; ++nn) {  // pairs of 8-key n-tiles
        uint32_t bfr[4];
        ldmatrix_x4(bfr, kt + (nn * 16 + (lane & 7) + (lane >> 4) * 8) * C::LDS + kk * 16 + ((lane >> 3) & 1) * 8);
        mma_bf16_16816(s_acc[2 * nn], a, bfr[0], bfr[1]);
        mma_bf16_16816(s_acc[2 * nn + 1], a, bfr[2], bfr[3]);
      }
    }

    // ---- mask + online softmax ----
    const int kv0 = t * ATT_BN;
    const bool need_mask = (kv0 + ATT_BN > p.Lk) || (p.causal && (kv0 + ATT_BN - 1 > q0 + warp * 16));
    if (need_mask) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int kj = kv0 + i * 8 + (lane & 3) * 2 + (e & 1);
          const int qi = qi0 + (e >> 1) * 8;
          if (kj >= p.Lk || (p.causal && kj > qi)) s_acc[i][e] = -INFINITY;
        }
      }
    }
    float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mx[0] = fmaxf(mx[0], fmaxf(s_acc[i][0], s_acc[i][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s_acc[i][2], s_acc[i][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      // every row sees key 0 in its first tile (causal rows included), so mx is finite from t = 0 on
      corr[r] = exp2f((m_run[r] - mx[r]) * p.scale_log2);
      msc[r] = mx[r] * p.scale_log2;
      m_run[r] = mx[r];
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pfrag[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float p0 = exp2f(s_acc[i][0] * p.scale_log2 - msc[0]);
      const float p1 = exp2f(s_acc[i][1] * p.scale_log2 - msc[0]);
      const float p2 = exp2f(s_acc[i][2] * p.scale_log2 - msc[1]);
      const float p3 = exp2f(s_acc[i][3] * p.scale_log2 - msc[1]);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pfrag[i][0] = pack_bf16(p0, p1);
      pfrag[i][1] = pack_bf16(p2, p3);
    }
    l_run[0] = l_run[0] * corr[0] + rs[0];
    l_run[1] = l_run[1] * corr[1] + rs[1];
#pragma unroll
    for (int i = 0; i < C::DP / 8; ++i) {
      o_acc[i][0] *= corr[0]; o_acc[i][1] *= corr[0];
      o_acc[i][2] *= corr[1]; o_acc[i][3] *= corr[1];
    }

    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {  // 16 keys per step
      uint32_t a[4] = {pfrag[2 * kk][0], pfrag[2 * kk][1], pfrag[2 * kk + 1][0], pfrag[2 * kk + 1][1]};
#pragma unroll
      for (int nn = 0; nn < C::DP / 16; ++nn) {  // pairs of 8-wide d n-tiles
        uint32_t bfr[4];
        ldmatrix_x4_trans(bfr, vt + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * C::LDS + nn * 16 + (lane >> 4) * 8);
        mma_bf16_16816(o_acc[2 * nn], a, bfr[0], bfr[1]);
        mma_bf16_16816(o_acc[2 * nn + 1], a, bfr[2], bfr[3]);
      }
    }
  }

  // ---- finalize: O /= l, write token-major ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int srow = warp * 16 + (lane >> 2) + r * 8;
    const int grow = q_rows[srow];
    if (grow < 0) continue;
    const float inv = 1.0f / l_run[r];
    __nv_bfloat16* dst = p.o + (size_t)grow * p.ldo + col0;
#pragma unroll
    for (int i = 0; i < C::DP / 8; ++i) {
      const int c = i * 8 + (lane & 3) * 2;
      if (c < D) {
        *reinterpret_cast<uint32_t*>(dst + c) = pack_bf16(o_acc[i][2 * r] * inv, o_acc[i][2 * r + 1] * inv);
      }
    }
  }
}

template <int D>
static int launch_attention(const AttnParams& p, int n_problems, cudaStream_t stream) {
  using C = AttnCfg<D>;
  static SmemAttrOnce smem_attr_attr_done;
  { cudaError_t e = smem_attr_attr_done.ensure(attention_kernel<D>, C::SMEM_BYTES); if (e != cudaSuccess) return (int)e; }
  dim3 grid(ceil_div(p.Lq, ATT_BM), n_problems);
  { cudaError_t le__ = launch_pdl(attention_kernel<D>, grid, ATT_THREADS, C::SMEM_BYTES, stream, p); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

// Debug/parity export of the SCTA row permutation (bit-exact check against the reference's window_partition).
__global__ void scta_rows_kernel(AttnParams p, int B, int* out) {
  const int L = p.Lq;
  const int total = B * p.nwin * L;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int s = i % L;
    const int win = (i / L) % p.nwin;
    const int b = i / (L * p.nwin);
    out[i] = scta_row(p, b, win, s);
  }
}

static void fill_scta_geometry(AttnParams& p, int F, int H, int W) {
  p.F = F; p.H = H; p.W = W;
  // window rule of attention.py:30-33,661-668: decided from h only
  p.ws = (H <= 4) ? 0 : ((H / 8) >= 4 ? 8 : 4);
  if (p.ws == 0) { p.nwx = 1; p.nwin = 1; p.Lq = p.Lk = F * H * W; }
  else { p.nwx = W / p.ws; p.nwin = (H / p.ws) * p.nwx; p.Lq = p.Lk = F * p.ws * p.ws; }
}

}  // namespace seer

using namespace seer;

// mma.sync path: every head dim / geometry.  The d = 40 level-0 problems are routed to the tcgen05 kernel in
// attention_tc.cu by the C entry point defined there.
int seer::attention_mma_launch(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo,
                               int mode, int heads, int head_dim, int n_outer, int Lq, int Lk, int F, int H, int W,
                               void* stream) {
  SEER_CHECK_ARG(q && k && v && o && heads > 0 && n_outer > 0);
  SEER_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0);
  AttnParams p{};
  p.q = (const __nv_bfloat16*)q; p.ldq = ldq;
  p.k = (const __nv_bfloat16*)k; p.ldk = ldk;
  p.v = (const __nv_bfloat16*)v; p.ldv = ldv;
  p.o = (__nv_bfloat16*)o; p.ldo = ldo;
  p.mode = mode; p.heads = heads; p.n_outer = n_outer;
  p.scale_log2 = (1.0f / sqrtf((float)head_dim)) * 1.4426950408889634f;
  int n_problems;
  if (mode == SEER_ATTN_SCTA) {
    SEER_CHECK_ARG(F > 0 && H > 0 && W > 0);
    fill_scta_geometry(p, F, H, W);
    if (p.ws) SEER_CHECK_ARG(H % p.ws == 0 && W % p.ws == 0);
    p.causal = 1;
    n_problems = n_outer * p.nwin * heads;
  } else if (mode == SEER_ATTN_FRAME) {
    // causal attention over the F frames of token l (H = tokens per frame) of clip b = SCTA geometry with a 1x1 window
    // per token: row = (b*F + f)*H + l  (FSText temporal blocks, attention.py:393-396)
    SEER_CHECK_ARG(F > 0 && H > 0);
    p.mode = SEER_ATTN_SCTA;
    p.F = F; p.H = H; p.W = 1; p.ws = 1; p.nwx = 1; p.nwin = H; p.Lq = p.Lk = F;
    p.causal = 1;
    n_problems = n_outer * p.nwin * heads;
  } else if (mode == SEER_ATTN_SPATIAL || mode == SEER_ATTN_CROSS) {
    SEER_CHECK_ARG(Lq > 0 && Lk > 0);
    p.Lq = Lq; p.Lk = Lk; p.causal = 0;
    p.nwin = 1; p.nwx = 1;
    n_problems = n_outer * heads;
  } else {
    return SEER_EINVAL;
  }
  SEER_CHECK_ARG(n_problems <= 65535 * 32);
  if (n_problems > 65535) return SEER_EUNSUPPORTED;
  switch (head_dim) {
    case 40: return launch_attention<40>(p, n_problems, (cudaStream_t)stream);
    case 80: return launch_attention<80>(p, n_problems, (cudaStream_t)stream);
    case 96: return launch_attention<96>(p, n_problems, (cudaStream_t)stream);
    case 160: return launch_attention<160>(p, n_problems, (cudaStream_t)stream);
    default: return SEER_EUNSUPPORTED;
  }
}

extern "C" int seer_b200_scta_row_index(int B, int F, int H, int W, int* out_dev, int* out_nwin, int* out_L, void* stream) {
  SEER_CHECK_ARG(B > 0 && F > 0 && H > 0 && W > 0);
  AttnParams p{};
  fill_scta_geometry(p, F, H, W);
  if (out_nwin) *out_nwin = p.nwin;
  if (out_L) *out_L = p.Lq;
  if (out_dev) {
    scta_rows_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(p, B, out_dev);
    SEER_LAUNCH_CHECK();
  }
  return SEER_OK;
}
