// tcgen05 flash attention for the d = 40 level of the Seer UNet (32x32 latents: 84 % of the attention-core FLOPs), and
// the C entry point seer_b200_attention that routes every other shape to the mma.sync kernel in attention.cu.
//
//   mode SPATIAL : per-frame self attention, L = h*w tokens            (reference attention.py:310-311, 512-554)
//   mode CROSS   : per-frame cross attention against Lk (= 77) text tokens                     (attention.py:313-322)
//   mode SCTA    : spatial-causal temporal attention over one 8x8 window of every frame, sequence order
//                  (frame, wy, wx), lower-triangular causal          (attention.py:632-703, window order :42-53)
//
// One CTA = 128 queries of one (problem, head); two CTAs share an SM (80 KB smem, 256 TMEM columns each) so one CTA's
// softmax overlaps the other's MMAs.  Per 128-key tile:
//   warp 0  TMA: K and V tiles [128 rows x 64 columns] straight from the token-major projection buffers.  The head
//           split (column offset head*40) and the window partition (a 4-D box {64 ch, 8, 8, 2 frames} of the
//           [B*F, H, W, C] view) are TMA coordinates — no gather, no copies.  Columns 40..63 of a box belong to the next
//           head; Q's are zeroed in smem once, so they contribute nothing to Q K^T, and the matching O columns are dropped.
//   warp 1  tcgen05.mma: S[128x128] = Q K^T into TMEM; then O[128x64] += P V (P from smem, V as MN-major B operand),
//           accumulated IN TMEM across all key tiles.
//   warps 2-5  softmax, thread = query row = TMEM lane: the 128 scores of the row are pulled into registers with four
//           back-to-back tcgen05.ld (one wait), S is handed back to the MMA warp at once (so Q K^T of tile j+1 runs under
//           the softmax of tile j), then max / exp2 / sum / bf16 P -> smem in the UMMA K-major swizzled layout.
//           The running maximum is LAZY: it is raised (and O in TMEM rescaled by a tcgen05.ld / st round trip) only when
//           some row of the warp would otherwise produce P > 2^8 — with the stale maximum P / l stay exact because both
//           use the same reference, so the per-tile O correction of flash attention leaves the critical path.
// The kernel is MUFU-bound by design (128x128 exp2 per tile = 1024 SM cycles vs 512 tensor cycles).
#include "common.cuh"
#include "seer_b200.h"

namespace seer {

// Share of the exp2 evaluated on the FMA pipe instead of MUFU (unmasked tiles only): SEER_ATTN_POLY of every 4 score pairs.
// 2^x = 2^n * p(f), n = round(x), f = x - n in [-0.5, 0.5], p = degree-3 minimax (max relative error 7.5e-5, far below the
// bf16 rounding of P, 2^-9); n comes out of the magic-number add (1.5 * 2^23), its low bits are shifted into the exponent.
// Measured on B200 (profiles/r2_attn_poly_ab.txt, spatial / SCTA at d = 40, L = 1024): share 0 -> 859 / 615 us, 1/4 -> 856 / 645,
// 2/4 -> 874 / 677, 3/4 -> 993 / 747.  The XU pipe is ~60 % busy: the softmax warps are bound by their dependent chain and
// issue slots (two warps per scheduler at the 168-register cap), not by MUFU throughput, so moving work to the FMA pipe
// only adds instructions.  Kept as a build-time switch (-DSEER_ATTN_POLY=n), off by default.
#ifndef SEER_ATTN_POLY
#define SEER_ATTN_POLY 0
#endif
__device__ __forceinline__ void exp2_poly_pair(f2_t X, float& e0, float& e1) {
  float x0, x1;
  f2_unpack(X, x0, x1);
  const f2_t xc = f2_pack(fmaxf(x0, -126.f), fmaxf(x1, -126.f));       // below 2^-126 the exponent add would wrap
  const f2_t rr = f2_add(xc, f2_pack(12582912.f, 12582912.f));
  const f2_t tt = f2_add(rr, f2_pack(-12582912.f, -12582912.f));
  const f2_t fr = f2_fma(tt, f2_pack(-1.f, -1.f), xc);
  f2_t pp = f2_fma(f2_pack(0.05517132207751274f, 0.05517132207751274f), fr, f2_pack(0.24261054396629333f, 0.24261054396629333f));
  pp = f2_fma(pp, fr, f2_pack(0.6932609677314758f, 0.6932609677314758f));
  pp = f2_fma(pp, fr, f2_pack(0.9999281167984009f, 0.9999281167984009f));
  float r0, r1, p0, p1;
  f2_unpack(rr, r0, r1);
  f2_unpack(pp, p0, p1);
  e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(r0) << 23));
  e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(r1) << 23));
}

constexpr int AT_BM = 128;     // queries per CTA
constexpr int AT_BN = 128;     // keys per tile
constexpr int AT_DP = 64;      // padded head dim (one 128-byte swizzle atom)
constexpr int AT_TILE = AT_BM * AT_DP * 2;           // 16 KB: Q / K / V tile
constexpr int AT_P_BYTES = AT_BM * AT_BN * 2;        // 32 KB: P tile (two 64-key atoms)
constexpr int AT_SMEM = 3 * AT_TILE + AT_P_BYTES + 256 + 1024;
constexpr int AT_THREADS = 192;

struct AttnTcParams {
  __nv_bfloat16* o;
  int ldo;
  int mode, heads;
  int Lq, Lk;
  int F, H, W, nwx, nwin;      // SCTA geometry (window side fixed at 8)
  float scale_log2;            // d^-0.5 * log2(e)
  int causal;
  int look_ahead;              // persistent kernel: issue the next item's first S before the current item's last P V
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn at_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

template <int D>
__global__ void __launch_bounds__(AT_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  pdl_wait();   // PDL secondary only: multi-wave grids must not hand their SMs to the successor early
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AT_TILE;
  uint8_t* sV = sK + AT_TILE;
  uint8_t* sP = sV + AT_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + AT_P_BYTES);
  uint64_t* q_full = bars + 0;     // TMA: Q landed
  uint64_t* q_ready = bars + 1;    // softmax warps zeroed Q's pad columns (count 4)
  uint64_t* k_full = bars + 2;
  uint64_t* k_empty = bars + 3;    // S MMA has read K
  uint64_t* v_full = bars + 4;
  uint64_t* v_empty = bars + 5;    // P V MMA has read V (and P)
  uint64_t* s_full = bars + 6;     // S in TMEM
  uint64_t* s_free = bars + 7;     // softmax finished reading S (count 4)
  uint64_t* p_full = bars + 8;     // P in smem (count 4)
  uint64_t* o_full = bars + 9;     // P V of tile j has completed (O updated, P and V consumed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x;
  const int head = blockIdx.y % p.heads;
  const int outer = blockIdx.y / p.heads;       // spatial/cross: frame; scta: b * nwin + win
  const int col0 = head * D;
  int b = 0, wy = 0, wx = 0;
  if (p.mode == SEER_ATTN_SCTA) {
    b = outer / p.nwin;
    const int win = outer - b * p.nwin;
    wy = win / p.nwx;
    wx = win - wy * p.nwx;
  }
  int n_kv = ceil_div(p.Lk, AT_BN);
  if (p.causal) n_kv = min(n_kv, qt + 1);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_ready, 4);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;           // 128 columns
  const uint32_t tO = tmem_base + 128;     // 64 columns

  if (warp == 0) {
    // ===================== TMA producer =====================
    auto load_tile = [&](uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int tile) {
      if (p.mode == SEER_ATTN_SCTA) tma_load_4d(dst, tm, bar, col0, wx * 8, wy * 8, b * p.F + tile * 2);
      else tma_load_2d(dst, tm, bar, col0, outer * (tm == &tmQ ? p.Lq : p.Lk) + tile * AT_BN);
    };
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, AT_TILE);
      load_tile(sQ, &tmQ, q_full, qt);
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(k_empty, (j & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(k_full, AT_TILE);
        load_tile(sK, &tmK, k_full, j);
      }
      __syncwarp();
      mbar_wait(v_empty, (j & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(v_full, AT_TILE);
        load_tile(sV, &tmV, v_full, j);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(AT_BM, AT_BN);          // S = Q K^T : N = 128 keys, K = 64 (padded d)
    constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(AT_BM, AT_DP);      // O = P V   : N = 64 (padded d), K = 128 keys, V MN-major
    const uint64_t q_desc = umma_desc_sw128(smem_u32(sQ));
    const uint64_t k_desc = umma_desc_sw128(smem_u32(sK));
    const uint64_t v_desc = umma_desc_sw128_mn(smem_u32(sV));
    const uint64_t p_desc0 = umma_desc_sw128(smem_u32(sP));
    const uint64_t p_desc1 = umma_desc_sw128(smem_u32(sP + AT_BM * 128));
    mbar_wait(q_ready, 0);
    // S(j+1) = Q K(j+1)^T is issued BEFORE the P V of tile j: it only needs the softmax warps to have pulled S(j) into
    // registers (s_free), so it runs under the exp pass of tile j and S is ready when the softmax warps come back.
    auto issue_s = [&](int j) {
      mbar_wait(k_full, j & 1);
      mbar_wait(s_free, (j & 1) ^ 1);          // softmax has drained S of tile j-1
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < AT_DP / 16; ++k)
          umma_bf16(tS, q_desc + (uint64_t)(k * 2), k_desc + (uint64_t)(k * 2), idesc_s, k != 0);
        umma_commit(s_full);
        umma_commit(k_empty);
      }
      __syncwarp();
    };
    issue_s(0);
    for (int j = 0; j < n_kv; ++j) {
      const uint32_t ph = j & 1;
      if (j + 1 < n_kv) issue_s(j + 1);
      mbar_wait(v_full, ph);
      mbar_wait(p_full, ph);                   // P of tile j in smem; any rescale of O by the softmax warps has retired
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < AT_BN / 16; ++k) {
          // A: 16 keys = 32 B inside the 64-key atom (two atoms 16 KB apart); B: 16 keys = two 8-row groups = 2048 B
          const uint64_t a = (k < 4 ? p_desc0 : p_desc1) + (uint64_t)((k & 3) * 2);
          umma_bf16(tO, a, v_desc + (uint64_t)(k * 128), idesc_o, (j | k) != 0);
        }
        umma_commit(o_full);
        umma_commit(v_empty);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax warps (thread = query row) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;               // row inside the tile = TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    // zero Q's pad columns D..63 (they hold the next head's channels), then hand Q to the MMA warp
    mbar_wait(q_full, 0);
#pragma unroll
    for (int c = D / 8; c < 8; ++c) sts128u(sQ + r * 128 + ((c ^ (r & 7)) << 4), make_uint4(0, 0, 0, 0));
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(q_ready);

    float m_run = -INFINITY, l_run = 0.f;      // m_run: the (lazily raised) reference maximum of this row
    const float sl2 = p.scale_log2;
    const int qi = qt * AT_BM + r;             // query index in the sequence

    for (int j = 0; j < n_kv; ++j) {
      const uint32_t ph = j & 1;
      const int kv0 = j * AT_BN;
      const bool need_mask = (kv0 + AT_BN > p.Lk) || (p.causal && j == qt);
      const int k_hi = min(p.Lk - kv0, p.causal && j == qt ? r + 1 : AT_BN);   // keys [0, k_hi) of this tile are visible
      mbar_wait(s_full, ph);
      tc_fence_after();
      // ---- the whole score row into registers, S handed back to the MMA warp ----
      uint32_t v[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32(tS + lane_addr + c * 32, v[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      // ---- row maximum of this tile (4 independent chains) ----
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (need_mask) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (c * 32 + k < k_hi) mx4[c] = fmaxf(mx4[c], __uint_as_float(v[c][k]));
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int k = 0; k < 32; ++k) mx4[c] = fmaxf(mx4[c], __uint_as_float(v[c][k]));
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      if (j == 0) {
        m_run = mx;                            // O is overwritten (not accumulated) by the first P V
      } else {
        // P V of tile j-1 must have completed before P is overwritten below (and before O may be rescaled)
        mbar_wait(o_full, ph ^ 1);
        // raise the reference maximum only if some row of this warp would exceed 2^8 with the stale one
        if (__any_sync(0xffffffffu, (mx - m_run) * sl2 > 8.0f)) {
          tc_fence_after();
          const float m_new = fmaxf(m_run, mx);
          const float corr = exp2f((m_run - m_new) * sl2);
          l_run *= corr;
          m_run = m_new;
#pragma unroll
          for (int c = 0; c < 3; ++c) {        // columns 0..47 hold the D = 40 live output channels
            uint32_t o[16];
            tmem_ld_32x16(tO + lane_addr + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
            tmem_st_32x16(tO + lane_addr + c * 16, o);
          }
          tmem_st_wait();
          tc_fence_before();
        }
      }
      // ---- p = exp2((s - m_run) * scale), row sum, bf16 P into the swizzled K-major A tile ----
      const float msc = m_run * sl2;
      float sum = 0.f;
      if (need_mask) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint8_t* prow = sP + (c >> 1) * (AT_BM * 128) + r * 128;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float e[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[k]) : "f"(fmaf(__uint_as_float(v[c][8 * g + k]), sl2, -msc)));
              if (c * 32 + 8 * g + k >= k_hi) e[k] = 0.f;
            }
            sum += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
            uint4 o;
            o.x = pack_bf16(e[0], e[1]);
            o.y = pack_bf16(e[2], e[3]);
            o.z = pack_bf16(e[4], e[5]);
            o.w = pack_bf16(e[6], e[7]);
            sts128u(prow + ((((c & 1) * 4 + g) ^ (r & 7)) << 4), o);
          }
        }
      } else {
        // unmasked tiles (all but the diagonal / tail tile): packed fp32x2 scale and sum, one issue slot per two scores
        const f2_t sl22 = f2_pack(sl2, sl2), nmsc2 = f2_pack(-msc, -msc);
        f2_t sum2[2] = {f2_pack(0.f, 0.f), f2_pack(0.f, 0.f)};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint8_t* prow = sP + (c >> 1) * (AT_BM * 128) + r * 128;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float x0, x1, e0, e1;
              f2_unpack(f2_fma(f2_pack_u(v[c][8 * g + 2 * k], v[c][8 * g + 2 * k + 1]), sl22, nmsc2), x0, x1);
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(x0));
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(x1));
              sum2[k & 1] = f2_add(sum2[k & 1], f2_pack(e0, e1));
              o[k] = pack_bf16(e0, e1);
            }
            sts128u(prow + ((((c & 1) * 4 + g) ^ (r & 7)) << 4), make_uint4(o[0], o[1], o[2], o[3]));
          }
        }
        float s0, s1;
        f2_unpack(f2_add(sum2[0], sum2[1]), s0, s1);
        sum = s0 + s1;
      }
      l_run += sum;
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }

    // ---- O (accumulated in TMEM) after the last P V ----
    mbar_wait(o_full, (n_kv - 1) & 1);
    tc_fence_after();
    float o_acc[48];
    {
      uint32_t v0[32], v1[16];
      tmem_ld_32x32(tO + lane_addr, v0);
      tmem_ld_32x16(tO + lane_addr + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[i] = __uint_as_float(v0[i]);
#pragma unroll
      for (int i = 0; i < 16; ++i) o_acc[32 + i] = __uint_as_float(v1[i]);
    }

    // ---- finalize: O / l -> bf16, token-major row ----
    size_t grow;
    if (p.mode == SEER_ATTN_SCTA) {
      const int f = qt * 2 + (r >> 6), iy = (r >> 3) & 7, ix = r & 7;
      grow = ((size_t)(b * p.F + f) * p.H + wy * 8 + iy) * p.W + wx * 8 + ix;
    } else {
      grow = (size_t)outer * p.Lq + qi;
    }
    if (qi < p.Lq) {
      const float inv = 1.0f / l_run;
      __nv_bfloat16* dst = p.o + grow * p.ldo + col0;
#pragma unroll
      for (int g = 0; g < D / 8; ++g) {
        uint4 o;
        o.x = pack_bf16(o_acc[8 * g] * inv, o_acc[8 * g + 1] * inv);
        o.y = pack_bf16(o_acc[8 * g + 2] * inv, o_acc[8 * g + 3] * inv);
        o.z = pack_bf16(o_acc[8 * g + 4] * inv, o_acc[8 * g + 5] * inv);
        o.w = pack_bf16(o_acc[8 * g + 6] * inv, o_acc[8 * g + 7] * inv);
        *reinterpret_cast<uint4*>(dst + 8 * g) = o;
      }
    }
  }

  tc_fence_before();
  __syncwarp();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// 2-D map over a token-major buffer [rows, ld] restricted to its first `cols` columns; box {64, 128}
static int at_map_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld) {
  EncodeTiledFn enc = at_encode_fn();
  if (!enc) return SEER_ENODRIVER;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {AT_DP, AT_BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SEER_OK : SEER_EINVAL;
}
// 4-D map [n_frames, H, W, cols] (row pitch ld); box {64 ch, 8, 8, 2 frames} = one 8x8 window of two frames
static int at_map_4d(CUtensorMap* tm, const void* base, uint64_t n_frames, uint64_t H, uint64_t W, uint64_t cols, uint64_t ld) {
  EncodeTiledFn enc = at_encode_fn();
  if (!enc) return SEER_ENODRIVER;
  cuuint64_t dims[4] = {cols, W, H, n_frames};
  cuuint64_t strides[3] = {ld * 2, W * ld * 2, H * W * ld * 2};
  cuuint32_t box[4] = {AT_DP, 8, 8, 2};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SEER_OK : SEER_EINVAL;
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent variant (default): 2 CTAs per SM loop over (problem, query tile) work items with ONE prologue (barrier init,
// TMEM allocation) per CTA, and the tile pipeline runs across item boundaries: Q of the next item is fetched as soon as the
// last Q K^T of the current one has been issued, its S(0) is computed while the softmax warps normalise and store O.
// With 8 (spatial), 1..8 (SCTA) or 1 (cross) key tiles per item the per-CTA prologue was ~20 % of the softmax warps' time
// (ncu, profiles/r1_attention_tc.summary.txt; the L = 4096 problems of config 5 reach 77 % of the MUFU roofline with the
// same tile loop, the L = 1024 ones 64 %).
// The head split uses tensor maps over the [rows, heads, d] view with a 64-wide box: columns d..63 of every tile are
// out-of-bounds of the head and arrive as ZEROS, so nothing has to be zeroed in shared memory.
// Item order: non-causal = query tiles of a problem adjacent (K/V stay in L2); causal = longest items (highest query tile)
// first, round-robin over the CTAs, so every CTA receives the same mix of lengths.
// ---------------------------------------------------------------------------------------------------------------------
// PT: P is handed to the P V MMA through TENSOR MEMORY (tcgen05.st by the softmax warps, tcgen05.mma with a TMEM A operand):
// no swizzled smem tile, no generic -> async proxy fence on the softmax warps' critical chain (measured ~250 clk per tile).
template <int D, bool PT>
__global__ void __launch_bounds__(AT_THREADS, 2)
attention_tc_persist_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                            const __grid_constant__ CUtensorMap tmV, const AttnTcParams p, const int n_problems,
                            const int nq_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AT_TILE;
  uint8_t* sV = sK + AT_TILE;
  uint8_t* sP = sV + AT_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + AT_P_BYTES);
  uint64_t* q_full = bars + 0;     // TMA: Q of item n landed
  uint64_t* q_empty = bars + 1;    // every Q K^T of item n has completed: Q may be overwritten
  uint64_t* k_full = bars + 2;
  uint64_t* k_empty = bars + 3;    // S MMA has read K
  uint64_t* v_full = bars + 4;
  uint64_t* v_empty = bars + 5;    // P V MMA has read V (and P)
  uint64_t* s_full = bars + 6;     // S in TMEM
  uint64_t* s_free = bars + 7;     // softmax pulled S into registers (count 4)
  uint64_t* p_full = bars + 8;     // P in smem (count 4)
  uint64_t* o_full = bars + 9;     // P V of a tile has completed
  uint64_t* o_free = bars + 10;    // softmax read the finished O of item n (count 4): P V(0) of item n+1 may overwrite it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = n_problems * nq_tiles;
  const int n_kv_full = ceil_div(p.Lk, AT_BN);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    mbar_init(o_free, 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;           // 128 columns
  const uint32_t tO = tmem_base + 128;     // 64 columns
  const uint32_t tP = tmem_base + 192;     // 64 columns: P as bf16 pairs (PT)
  pdl_wait();                              // PDL secondary: the prologue above overlapped the previous kernel's tail

  // work item -> (problem, query tile)
  auto decode = [&](int idx, int& prob, int& qt) {
    if (p.causal) { qt = nq_tiles - 1 - idx / n_problems; prob = idx - (idx / n_problems) * n_problems; }
    else { prob = idx / nq_tiles; qt = idx - prob * nq_tiles; }
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t it = 0, n = 0;
    for (int idx = blockIdx.x; idx < total; idx += gridDim.x, ++n) {
      int prob, qt;
      decode(idx, prob, qt);
      const int head = prob % p.heads, outer = prob / p.heads;
      int b = 0, wy = 0, wx = 0;
      if (p.mode == SEER_ATTN_SCTA) {
        b = outer / p.nwin;
        const int win = outer - b * p.nwin;
        wy = win / p.nwx;
        wx = win - wy * p.nwx;
      }
      const int n_kv = p.causal ? min(n_kv_full, qt + 1) : n_kv_full;
      // Q: [rows, heads, d] view, columns d..63 arrive as zeros (so K's and V's pad columns — the next head's channels,
      // fetched through the plain token-major maps with full 128-byte rows — contribute nothing to Q K^T; the matching
      // O columns are never stored)
      auto load_tile = [&](uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int tile, int L) {
        if (p.mode == SEER_ATTN_SCTA) tma_load_4d(dst, tm, bar, head * D, wx * 8, wy * 8, b * p.F + tile * 2);
        else tma_load_2d(dst, tm, bar, head * D, outer * L + tile * AT_BN);
      };
      mbar_wait(q_empty, (n & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, AT_TILE);
        if (p.mode == SEER_ATTN_SCTA) tma_load_5d(sQ, &tmQ, q_full, 0, head, wx * 8, wy * 8, b * p.F + qt * 2);
        else tma_load_3d(sQ, &tmQ, q_full, 0, head, outer * p.Lq + qt * AT_BN);
      }
      __syncwarp();
      for (int j = 0; j < n_kv; ++j, ++it) {
        mbar_wait(k_empty, (it & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(k_full, AT_TILE);
          load_tile(sK, &tmK, k_full, j, p.Lk);
        }
        __syncwarp();
        mbar_wait(v_empty, (it & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(v_full, AT_TILE);
          load_tile(sV, &tmV, v_full, j, p.Lk);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(AT_BM, AT_BN);          // S = Q K^T : N = 128 keys, K = 64 (padded d)
    constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(AT_BM, AT_DP);      // O = P V   : N = 64 (padded d), K = 128 keys, V MN-major
    const uint64_t q_desc = umma_desc_sw128(smem_u32(sQ));
    const uint64_t k_desc = umma_desc_sw128(smem_u32(sK));
    const uint64_t v_desc = umma_desc_sw128_mn(smem_u32(sV));
    const uint64_t p_desc0 = umma_desc_sw128(smem_u32(sP));
    const uint64_t p_desc1 = umma_desc_sw128(smem_u32(sP + AT_BM * 128));
    uint32_t it = 0, n = 0;
    // S of tile `t` (global tile counter); `last` = it is the last Q K^T of its item -> Q may be refilled afterwards
    auto issue_s = [&](uint32_t t, bool last) {
      mbar_wait(k_full, t & 1);
      mbar_wait(s_free, (t & 1) ^ 1);          // softmax has drained S of tile t-1
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < AT_DP / 16; ++k)
          umma_bf16(tS, q_desc + (uint64_t)(k * 2), k_desc + (uint64_t)(k * 2), idesc_s, k != 0);
        umma_commit(s_full);
        umma_commit(k_empty);
        if (last) umma_commit(q_empty);
      }
      __syncwarp();
    };
    // Optional (p.look_ahead, an A/B switch that did not pay): S of the first tile of the NEXT item issued before P V of the current
    // item's last tile, as S(j+1) precedes P V(j) inside an item — Q / K of the next item arrive as soon as the last S of this one
    // has read the buffers, and the softmax warps hand S back (s_free) at the start of their pass
    const bool look_ahead = p.look_ahead != 0;
    bool s_issued = false;
    for (int idx = blockIdx.x; idx < total; idx += gridDim.x, ++n) {
      int prob, qt;
      decode(idx, prob, qt);
      const int n_kv = p.causal ? min(n_kv_full, qt + 1) : n_kv_full;
      if (!s_issued) {
        mbar_wait(q_full, n & 1);
        issue_s(it, n_kv == 1);
      }
      s_issued = false;
      for (int j = 0; j < n_kv; ++j, ++it) {
        const uint32_t ph = it & 1;
        if (j + 1 < n_kv) {
          issue_s(it + 1, j + 2 == n_kv);
        } else if (look_ahead && idx + (int)gridDim.x < total) {
          int prob2, qt2;
          decode(idx + (int)gridDim.x, prob2, qt2);
          const int n_kv2 = p.causal ? min(n_kv_full, qt2 + 1) : n_kv_full;
          mbar_wait(q_full, (n + 1) & 1);
          issue_s(it + 1, n_kv2 == 1);
          s_issued = true;
        }
        mbar_wait(v_full, ph);
        mbar_wait(p_full, ph);                 // P of this tile in smem; any rescale of O by the softmax warps has retired
        if (j == 0) mbar_wait(o_free, (n & 1) ^ 1);   // the previous item's O has been read out of TMEM
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < AT_BN / 16; ++k) {
            if (PT) {
              umma_bf16_ts(tO, tP + (uint32_t)(k * 8), v_desc + (uint64_t)(k * 128), idesc_o, (j | k) != 0);   // 16 keys = 8 columns
            } else {
              const uint64_t a = (k < 4 ? p_desc0 : p_desc1) + (uint64_t)((k & 3) * 2);
              umma_bf16(tO, a, v_desc + (uint64_t)(k * 128), idesc_o, (j | k) != 0);
            }
          }
          umma_commit(o_full);
          umma_commit(v_empty);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== softmax warps (thread = query row) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;               // row inside the tile = TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float sl2 = p.scale_log2;
    uint32_t it = 0, n = 0;
    for (int idx = blockIdx.x; idx < total; idx += gridDim.x, ++n) {
      int prob, qt;
      decode(idx, prob, qt);
      const int head = prob % p.heads, outer = prob / p.heads;
      const int n_kv = p.causal ? min(n_kv_full, qt + 1) : n_kv_full;
      const int qi = qt * AT_BM + r;           // query index in the sequence
      float m_run = -INFINITY, l_run = 0.f;    // m_run: the (lazily raised) reference maximum of this row

      for (int j = 0; j < n_kv; ++j, ++it) {
        const uint32_t ph = it & 1;
        const int kv0 = j * AT_BN;
        const bool need_mask = (kv0 + AT_BN > p.Lk) || (p.causal && j == qt);
        const int k_hi = min(p.Lk - kv0, p.causal && j == qt ? r + 1 : AT_BN);   // keys [0, k_hi) of this tile are visible
        mbar_wait(s_full, ph);
        tc_fence_after();
        uint32_t v[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_32x32(tS + lane_addr + c * 32, v[c]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (need_mask) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (c * 32 + k < k_hi) mx4[c] = fmaxf(mx4[c], __uint_as_float(v[c][k]));
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int k = 0; k < 32; ++k) mx4[c] = fmaxf(mx4[c], __uint_as_float(v[c][k]));
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        if (j == 0) {
          m_run = mx;                          // O is overwritten (not accumulated) by the first P V of the item; P is free:
                                               // the previous item's epilogue waited for its last P V
        } else {
          mbar_wait(o_full, ph ^ 1);           // P V of the previous tile done: P may be overwritten, O may be rescaled
          if (__any_sync(0xffffffffu, (mx - m_run) * sl2 > 8.0f)) {
            tc_fence_after();
            const float m_new = fmaxf(m_run, mx);
            const float corr = exp2f((m_run - m_new) * sl2);
            l_run *= corr;
            m_run = m_new;
#pragma unroll
            for (int c = 0; c < 3; ++c) {      // columns 0..47 hold the D = 40 live output channels
              uint32_t o[16];
              tmem_ld_32x16(tO + lane_addr + c * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
              tmem_st_32x16(tO + lane_addr + c * 16, o);
            }
            tmem_st_wait();
            tc_fence_before();
          }
        }
        const float msc = m_run * sl2;
        float sum = 0.f;
        if (need_mask) {
          const int k_vis = __reduce_max_sync(0xffffffffu, k_hi);   // no row of this warp sees keys >= k_vis
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint8_t* prow = sP + (c >> 1) * (AT_BM * 128) + r * 128;
            uint32_t pk[16];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (c * 32 + 8 * g >= k_vis) {                         // warp-uniform: P = 0, no exp2
                if (PT) { pk[4 * g] = pk[4 * g + 1] = pk[4 * g + 2] = pk[4 * g + 3] = 0u; }
                else sts128u(prow + ((((c & 1) * 4 + g) ^ (r & 7)) << 4), make_uint4(0, 0, 0, 0));
                continue;
              }
              float e[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[k]) : "f"(fmaf(__uint_as_float(v[c][8 * g + k]), sl2, -msc)));
                if (c * 32 + 8 * g + k >= k_hi) e[k] = 0.f;
              }
              sum += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
              uint4 o;
              o.x = pack_bf16(e[0], e[1]);
              o.y = pack_bf16(e[2], e[3]);
              o.z = pack_bf16(e[4], e[5]);
              o.w = pack_bf16(e[6], e[7]);
              if (PT) { pk[4 * g] = o.x; pk[4 * g + 1] = o.y; pk[4 * g + 2] = o.z; pk[4 * g + 3] = o.w; }
              else sts128u(prow + ((((c & 1) * 4 + g) ^ (r & 7)) << 4), o);
            }
            if (PT) tmem_st_32x16(tP + lane_addr + c * 16, pk);
          }
        } else {
          const f2_t sl22 = f2_pack(sl2, sl2), nmsc2 = f2_pack(-msc, -msc);
          f2_t sum2[2] = {f2_pack(0.f, 0.f), f2_pack(0.f, 0.f)};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint8_t* prow = sP + (c >> 1) * (AT_BM * 128) + r * 128;
            uint32_t pk[16];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t o[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float x0, x1, e0, e1;
                const f2_t xs = f2_fma(f2_pack_u(v[c][8 * g + 2 * k], v[c][8 * g + 2 * k + 1]), sl22, nmsc2);
                if (k >= 4 - SEER_ATTN_POLY) {                       // compile-time: this pair goes through the FMA pipe
                  exp2_poly_pair(xs, e0, e1);
                } else {
                  f2_unpack(xs, x0, x1);
                  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(x0));
                  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(x1));
                }
                sum2[k & 1] = f2_add(sum2[k & 1], f2_pack(e0, e1));
                o[k] = pack_bf16(e0, e1);
              }
              if (PT) { pk[4 * g] = o[0]; pk[4 * g + 1] = o[1]; pk[4 * g + 2] = o[2]; pk[4 * g + 3] = o[3]; }
              else sts128u(prow + ((((c & 1) * 4 + g) ^ (r & 7)) << 4), make_uint4(o[0], o[1], o[2], o[3]));
            }
            if (PT) tmem_st_32x16(tP + lane_addr + c * 16, pk);
          }
          float s0, s1;
          f2_unpack(f2_add(sum2[0], sum2[1]), s0, s1);
          sum = s0 + s1;
        }
        l_run += sum;
        if (PT) tmem_st_wait(); else fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }

      // ---- item epilogue: O (accumulated in TMEM) after the last P V, handed back to the MMA warp at once ----
      mbar_wait(o_full, (it - 1) & 1);
      tc_fence_after();
      uint32_t v0[32], v1[8];
      tmem_ld_32x32(tO + lane_addr, v0);
      tmem_ld_32x8(tO + lane_addr + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
      size_t grow;
      if (p.mode == SEER_ATTN_SCTA) {
        const int b = outer / p.nwin, win = outer - b * p.nwin;
        const int wy = win / p.nwx, wx = win - wy * p.nwx;
        const int f = qt * 2 + (r >> 6), iy = (r >> 3) & 7, ix = r & 7;
        grow = ((size_t)(b * p.F + f) * p.H + wy * 8 + iy) * p.W + wx * 8 + ix;
      } else {
        grow = (size_t)outer * p.Lq + qi;
      }
      if (qi < p.Lq) {
        const float inv = 1.0f / l_run;
        __nv_bfloat16* dst = p.o + grow * p.ldo + head * D;
#pragma unroll
        for (int g = 0; g < D / 8; ++g) {
          const uint32_t* src = g < 4 ? &v0[8 * g] : &v1[8 * (g - 4)];
          uint4 o;
          o.x = pack_bf16(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv);
          o.y = pack_bf16(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv);
          o.z = pack_bf16(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv);
          o.w = pack_bf16(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv);
          *reinterpret_cast<uint4*>(dst + 8 * g) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncwarp();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// head-split maps with zero fill beyond the head: [rows, heads, d] view, box {64, 1, 128}
static int at_map_heads_3d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t heads, uint64_t d, uint64_t ld) {
  EncodeTiledFn enc = at_encode_fn();
  if (!enc) return SEER_ENODRIVER;
  cuuint64_t dims[3] = {d, heads, rows};
  cuuint64_t strides[2] = {d * 2, ld * 2};
  cuuint32_t box[3] = {AT_DP, 1, AT_BN};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SEER_OK : SEER_EUNSUPPORTED;
}
// [n_frames, H, W, heads, d] view; box {64, 1, 8, 8, 2 frames} = one 8x8 window of two frames of one head
static int at_map_heads_5d(CUtensorMap* tm, const void* base, uint64_t n_frames, uint64_t H, uint64_t W, uint64_t heads, uint64_t d,
                           uint64_t ld) {
  EncodeTiledFn enc = at_encode_fn();
  if (!enc) return SEER_ENODRIVER;
  cuuint64_t dims[5] = {d, heads, W, H, n_frames};
  cuuint64_t strides[4] = {d * 2, ld * 2, W * ld * 2, H * W * ld * 2};
  cuuint32_t box[5] = {AT_DP, 1, 8, 8, 2};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SEER_OK : SEER_EUNSUPPORTED;
}

static int env_flag(const char* name, int dflt) { return env_cached(name, dflt); }

}  // namespace seer

using namespace seer;

extern "C" int seer_b200_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo,
                                   int mode, int heads, int head_dim, int n_outer, int Lq, int Lk, int F, int H, int W,
                                   void* stream) {
  SEER_CHECK_ARG(q && k && v && o && heads > 0 && n_outer > 0);
  if ((head_dim == 80 || (head_dim == 160 && env_flag("SEER_ATTN_TC160", 1)) || (head_dim == 40 && env_flag("SEER_ATTN_D40_BN64", 0))) &&
      mode != SEER_ATTN_FRAME && env_flag("SEER_ATTN_TC80", 1)) {
    const int rc80 = attention_tc80_launch(q, ldq, k, ldk, v, ldv, o, ldo, mode, heads, head_dim, n_outer, Lq, Lk, F, H, W, stream);
    if (rc80 != SEER_EUNSUPPORTED) return rc80;
  }
  // tcgen05 path: d = 40, 128-row query tiles; SCTA needs 8x8 windows (H/8 >= 4) and an even frame count
  bool tc = head_dim == 40 && env_flag("SEER_ATTN_TC", 1) && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 &&
            ((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0) && ((uintptr_t)o % 16 == 0);
  AttnTcParams p{};
  int n_problems = 0, nq_tiles = 0;
  if (tc) {
    if (mode == SEER_ATTN_SCTA) {
      tc = F > 0 && H >= 32 && H % 8 == 0 && W % 8 == 0 && F % 2 == 0;
      if (tc) {
        p.F = F; p.H = H; p.W = W; p.nwx = W / 8; p.nwin = (H / 8) * p.nwx;
        p.Lq = p.Lk = F * 64; p.causal = 1;
        n_problems = n_outer * p.nwin * heads;
        nq_tiles = p.Lq / AT_BM;
      }
    } else if (mode == SEER_ATTN_SPATIAL || mode == SEER_ATTN_CROSS) {
      tc = Lq > 0 && Lk > 0 && Lq % AT_BM == 0;
      if (tc) {
        p.Lq = Lq; p.Lk = Lk; p.causal = 0; p.nwin = 1; p.nwx = 1;
        n_problems = n_outer * heads;
        nq_tiles = Lq / AT_BM;
      }
    } else if (mode == SEER_ATTN_FRAME) {
      tc = false;
    } else {
      return SEER_EINVAL;
    }
    if (n_problems > 65535) tc = false;
  }
  if (!tc) {
    debug_note_attention("attention_kernel<D> mma.sync (legacy path)");
    return attention_mma_launch(q, ldq, k, ldk, v, ldv, o, ldo, mode, heads, head_dim, n_outer, Lq, Lk, F, H, W, stream);
  }

  p.o = (__nv_bfloat16*)o; p.ldo = ldo;
  p.mode = mode; p.heads = heads;
  p.scale_log2 = (1.0f / sqrtf((float)head_dim)) * 1.4426950408889634f;
  const uint64_t C = (uint64_t)heads * head_dim;
  CUtensorMap tq, tk, tv;
  int rc;
  if (env_flag("SEER_ATTN_PERSIST", 1)) {
    // persistent kernel, zero-filling head-split maps
    int ok;
    if (mode == SEER_ATTN_SCTA) {
      const uint64_t nf = (uint64_t)n_outer * F;
      ok = at_map_heads_5d(&tq, q, nf, H, W, heads, head_dim, ldq) == SEER_OK && at_map_4d(&tk, k, nf, H, W, C, ldk) == SEER_OK &&
           at_map_4d(&tv, v, nf, H, W, C, ldv) == SEER_OK;
    } else {
      ok = at_map_heads_3d(&tq, q, (uint64_t)n_outer * Lq, heads, head_dim, ldq) == SEER_OK &&
           at_map_2d(&tk, k, (uint64_t)n_outer * Lk, C, ldk) == SEER_OK && at_map_2d(&tv, v, (uint64_t)n_outer * Lk, C, ldv) == SEER_OK;
    }
    if (ok) {
      // (measured, profiles/r2_attn_bench_p_tmem.txt: 870 vs 854 us spatial, 655 vs 616 us SCTA — the TMEM store + wait costs more
      //  than the smem store + proxy fence it replaces at 168 registers; kept as an A/B switch, off by default)
      const bool pt = env_flag("SEER_ATTN_P_TMEM", 0) != 0;
      // (measured, profiles/r2_attn_lookahead_ab.txt: cross 241.8 -> 239.1 us, spatial 857.7 -> 877.1 us, SCTA unchanged — the first S
      //  of an item is not what its softmax warps wait for; off)
      p.look_ahead = env_flag("SEER_ATTN_LOOKAHEAD", 0);
      auto kern = pt ? attention_tc_persist_kernel<40, true> : attention_tc_persist_kernel<40, false>;
      static SmemAttrOnce smem_attr_attr_done_p[2];
  { cudaError_t e = smem_attr_attr_done_p[pt].ensure(kern, AT_SMEM); if (e != cudaSuccess) return (int)e; }
      int dev = 0, nsm = 148;
      if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0)
        nsm = 148;
      const long total = (long)n_problems * nq_tiles;
      const int grid = (int)(total < 2L * nsm ? total : 2L * nsm);
      cudaError_t le = launch_pdl(kern, dim3(grid), dim3(AT_THREADS), (size_t)AT_SMEM, (cudaStream_t)stream, tq, tk,
                                  tv, p, n_problems, nq_tiles);
      if (le != cudaSuccess) return (int)le;
      SEER_LAUNCH_CHECK();
      debug_note_attention(pt ? "attention_tc_persist_kernel<40> tcgen05 (P in TMEM)" : "attention_tc_persist_kernel<40> tcgen05");
      return SEER_OK;
    }
    static bool warned = false;
    if (!warned) {
      warned = true;
      fprintf(stderr, "[seer_b200] attention: head-split tensor maps rejected by the driver, using the one-tile-per-CTA kernel\n");
    }
  }
  if (mode == SEER_ATTN_SCTA) {
    const uint64_t nf = (uint64_t)n_outer * F;
    if ((rc = at_map_4d(&tq, q, nf, H, W, C, ldq))) return rc;
    if ((rc = at_map_4d(&tk, k, nf, H, W, C, ldk))) return rc;
    if ((rc = at_map_4d(&tv, v, nf, H, W, C, ldv))) return rc;
  } else {
    if ((rc = at_map_2d(&tq, q, (uint64_t)n_outer * Lq, C, ldq))) return rc;
    if ((rc = at_map_2d(&tk, k, (uint64_t)n_outer * Lk, C, ldk))) return rc;
    if ((rc = at_map_2d(&tv, v, (uint64_t)n_outer * Lk, C, ldv))) return rc;
  }
  static SmemAttrOnce smem_attr_attr_done;
  { cudaError_t e = smem_attr_attr_done.ensure(attention_tc_kernel<40>, AT_SMEM); if (e != cudaSuccess) return (int)e; }
  dim3 grid(nq_tiles, n_problems);
  { cudaError_t le__ = launch_pdl(attention_tc_kernel<40>, grid, AT_THREADS, AT_SMEM, (cudaStream_t)stream, tq, tk, tv, p); if (le__ != cudaSuccess) return (int)le__; }
  SEER_LAUNCH_CHECK();
  debug_note_attention("attention_tc_kernel<40> tcgen05 (one tile per CTA)");
  return SEER_OK;
}
