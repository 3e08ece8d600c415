// tcgen05 flash attention for the d = 80 and d = 160 levels of the Seer UNet (16x16 / 8x8 / 4x4 latents at the bench shape):
// the persistent structure of attention_tc.cu (2 CTAs per SM looping over (problem, query tile) items, S / O in TMEM, lazily
// raised maximum, packed fp32x2 softmax) with the tile geometry a two-atom head needs:
//
//   * d = 80 = 64 + 16 channels = TWO 128-byte swizzle atoms per row.  Q K^T issues exactly 5 k-steps of 16 (4 in atom 0,
//     1 in atom 1), so the 48 pad columns of atom 1 (the next head's channels) are never read: no zero fill, no smem zeroing.
//   * 128 queries x 64 keys per tile (Q 32 KB, K 16 KB, V 16 KB, P 16 KB = 80 KB, two CTAs per SM as for d = 40).
//   * O = P V with V as an MN-major B operand of N = 80: two N atoms 8 KB apart (descriptor LBO), 4 k-steps of 16 keys.
//   * SCTA windows are 4x4 (H = 16) or 8x8 (H = 32): the window partition is a 4-D TMA box {64 ch, ws, ws, frames}.
//
// Reference: seer/models/attention.py:310-322, 512-554 (spatial / cross), :632-703 (SCTA), window order :42-53.
// d = 160 = 64 + 64 + 32: three atoms, 10 exact k-steps, Q 48 KB + K 24 KB + V 24 KB + P 16 KB = 112 KB — two CTAs still share
// an SM because the 1024-byte alignment slack is cut to the 896 bytes a 128-byte-aligned dynamic smem window can need.
// Selected by seer_b200_attention for head_dim 80 / 160 (SEER_ATTN_TC80=0 falls back to the mma.sync kernel).
#include "common.cuh"
#include "seer_b200.h"

namespace seer {

constexpr int A8_BM = 128;               // queries per tile
constexpr int A8_BN = 64;                // keys per tile
constexpr int A8_P_BYTES = A8_BM * 128;          // one atom: 64 keys
constexpr int A8_SLACK = 896;                    // dynamic smem starts 128-byte aligned: 1024-alignment costs <= 896 B
constexpr int A8_THREADS = 192;
template <int D>
struct A8Cfg {
  static constexpr int NA = (D + 63) / 64;               // 64-channel swizzle atoms per row
  static constexpr int Q_BYTES = NA * A8_BM * 128;
  static constexpr int KV_BYTES = NA * A8_BN * 128;
  // P is single-buffered.  (Double-buffering it lets the softmax warps run two tiles ahead of the P V MMA, and then p_full —
  // one barrier, waited on by parity — can complete twice before the MMA warp looks: the MMA warp waits for a phase that has
  // already been overtaken and the CTA deadlocks.  Seen on 2-GPU runs; it bought nothing either: 119.5 vs 117 us at d = 80.)
  static constexpr int NP = 1;
  static constexpr int SMEM = Q_BYTES + 2 * KV_BYTES + NP * A8_P_BYTES + 128 + A8_SLACK;
  static constexpr int DP = (D + 15) / 16 * 16;         // contraction extent of Q K^T in whole k-steps (d = 40 -> 48)
  static constexpr bool QZ = (D % 16) != 0;             // Q needs zero-filled pad columns (fetched through a [rows, heads, d] map)
  static constexpr int DO = DP;                         // N of the P V MMA / O columns in TMEM
  static constexpr int TMEM_COLS = (64 + DO) <= 128 ? 128 : 256;
  static constexpr int MIN_CTAS = 2;
  static constexpr int MAXREG = 65536 / (MIN_CTAS * A8_THREADS) / 8 * 8;   // 168
  // S of tile t+1 pulled into a second register set in the middle of the exp2 pass of tile t (hides the tcgen05.ld latency
  // and hands S back to the MMA warp earlier); needs 64 more registers: only where the epilogue does not need them
  static constexpr bool PRE = D <= 48;
};

struct Attn80Params {
  __nv_bfloat16* o;
  int ldo;
  int mode, heads;
  int Lq, Lk;
  int F, H, W, ws, nwx, nwin;  // SCTA geometry
  float scale_log2;
  int causal;
};

typedef CUresult (*EncodeTiledFn80)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn80 a8_encode_fn() {
  static EncodeTiledFn80 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn80>(ptr);
  }
  return fn;
}

// MN-major SW128 B operand spanning two 64-element N atoms `lbo_bytes` apart; 8-row K groups 1024 B apart
__device__ __forceinline__ uint64_t a8_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int D>
__global__ void __launch_bounds__(A8_THREADS) __maxnreg__(A8Cfg<D>::MAXREG)
attention_tc80_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const Attn80Params p, const int n_problems, const int nq_tiles) {
  extern __shared__ uint8_t smem_raw[];
  using C = A8Cfg<D>;
  constexpr int A8_D = D, A8_Q_BYTES = C::Q_BYTES, A8_KV_BYTES = C::KV_BYTES;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (threadIdx.x == 0 && (size_t)(smem - smem_raw) > (size_t)A8_SLACK) {
    printf("[seer_b200] attention_tc80: dynamic shared memory base is not 128-byte aligned\n");
    __trap();
  }
  uint8_t* sQ = smem;                        // [NA atoms][128 rows][128 B]
  uint8_t* sK = sQ + A8_Q_BYTES;             // [NA atoms][64 rows][128 B]
  uint8_t* sV = sK + A8_KV_BYTES;
  uint8_t* sP = sV + A8_KV_BYTES;            // [128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + C::NP * A8_P_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;
  uint64_t* k_empty = bars + 3;
  uint64_t* v_full = bars + 4;
  uint64_t* v_empty = bars + 5;
  uint64_t* s_full = bars + 6;
  uint64_t* s_free = bars + 7;     // count 4
  uint64_t* p_full = bars + 8;     // count 4
  uint64_t* o_full = bars + 9;     // [2] P V of tile t has completed: barrier t & 1 (a wait two tiles back stays an unambiguous parity wait)
  uint64_t* o_free = bars + 11;    // count 4
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = n_problems * nq_tiles;
  const int n_kv_full = ceil_div(p.Lk, A8_BN);
  const int tpf = p.ws * p.ws;             // window tokens per frame (SCTA)

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
    mbar_init(&o_full[0], 1);
    mbar_init(&o_full[1], 1);
    mbar_init(o_free, 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;           // 64 columns
  const uint32_t tO = tmem_base + 64;      // D columns
  pdl_wait();

  auto decode = [&](int idx, int& prob, int& qt) {
    if (p.causal) { qt = nq_tiles - 1 - idx / n_problems; prob = idx - (idx / n_problems) * n_problems; }
    else { prob = idx / nq_tiles; qt = idx - prob * nq_tiles; }
  };
  // key tiles a query tile needs: causal rows of tile qt see keys <= qt*128 + 127
  auto kv_tiles = [&](int qt) { return p.causal ? min(n_kv_full, ceil_div(min((qt + 1) * A8_BM, p.Lk), A8_BN)) : n_kv_full; };

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t it = 0, n = 0;
    for (int idx = blockIdx.x; idx < total; idx += gridDim.x, ++n) {
      int prob, qt;
      decode(idx, prob, qt);
      const int head = prob % p.heads, outer = prob / p.heads;
      const int col0 = head * A8_D;
      int b = 0, wy = 0, wx = 0;
      if (p.mode == SEER_ATTN_SCTA) {
        b = outer / p.nwin;
        const int win = outer - b * p.nwin;
        wy = win / p.nwx;
        wx = win - wy * p.nwx;
      }
      const int n_kv = kv_tiles(qt);
      // one tile = two 64-column atoms; `rows` tile rows, `atom_bytes` bytes per atom
      auto load_tile = [&](uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int row0, int L, int atom_bytes) {
#pragma unroll
        for (int a = 0; a < C::NA; ++a) {
          if (p.mode == SEER_ATTN_SCTA) tma_load_4d(dst + a * atom_bytes, tm, bar, col0 + 64 * a, wx * p.ws, wy * p.ws, b * p.F + row0 / tpf);
          else tma_load_2d(dst + a * atom_bytes, tm, bar, col0 + 64 * a, outer * L + row0);
        }
      };
      mbar_wait(q_empty, (n & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, A8_Q_BYTES);
        if constexpr (C::QZ) {
          // d not a multiple of 16: Q through the [rows, heads, d] view, columns d..63 arrive as zeros, so the garbage
          // K reads in the padded k-step contributes nothing (K / V keep full-row maps: partial-extent boxes are ~3x slower)
          if (p.mode == SEER_ATTN_SCTA) tma_load_5d(sQ, &tmQ, q_full, 0, head, wx * p.ws, wy * p.ws, b * p.F + (qt * A8_BM) / tpf);
          else tma_load_3d(sQ, &tmQ, q_full, 0, head, outer * p.Lq + qt * A8_BM);
        } else {
          load_tile(sQ, &tmQ, q_full, qt * A8_BM, p.Lq, A8_BM * 128);
        }
      }
      __syncwarp();
      for (int j = 0; j < n_kv; ++j, ++it) {
        mbar_wait(k_empty, (it & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(k_full, A8_KV_BYTES);
          load_tile(sK, &tmK, k_full, j * A8_BN, p.Lk, A8_BN * 128);
        }
        __syncwarp();
        mbar_wait(v_empty, (it & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(v_full, A8_KV_BYTES);
          load_tile(sV, &tmV, v_full, j * A8_BN, p.Lk, A8_BN * 128);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(A8_BM, A8_BN);          // S = Q K^T : N = 64 keys, K = D
    constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(A8_BM, C::DO);      // O = P V   : N = D (padded to 16), K = 64 keys, V MN-major
    const uint64_t q_desc0 = umma_desc_sw128(smem_u32(sQ));              // atom a: + a * 128 rows * 128 B (in 16-byte units)
    const uint64_t k_desc0 = umma_desc_sw128(smem_u32(sK));              // atom a: + a * 64 rows * 128 B
    const uint64_t v_desc = a8_desc_mn(smem_u32(sV), A8_BN * 128);
    const uint64_t p_desc = umma_desc_sw128(smem_u32(sP));
    uint32_t it = 0, n = 0;
    auto issue_s = [&](uint32_t t, bool last) {
      mbar_wait(k_full, t & 1);
      mbar_wait(s_free, (t & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int a = 0; a < C::NA; ++a) {
          constexpr int full = 4;
          const int steps = (C::DP - 64 * a) >= 64 ? full : (C::DP - 64 * a) / 16;   // the last atom holds DP % 64 live channels
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < steps)
              umma_bf16(tS, q_desc0 + (uint64_t)(a * (A8_BM * 128 / 16) + k * 2), k_desc0 + (uint64_t)(a * (A8_BN * 128 / 16) + k * 2),
                        idesc_s, (a | k) != 0);
        }
        umma_commit(s_full);
        umma_commit(k_empty);
        if (last) umma_commit(q_empty);
      }
      __syncwarp();
    };
    for (int idx = blockIdx.x; idx < total; idx += gridDim.x, ++n) {
      int prob, qt;
      decode(idx, prob, qt);
      const int n_kv = kv_tiles(qt);
      mbar_wait(q_full, n & 1);
      issue_s(it, n_kv == 1);
      for (int j = 0; j < n_kv; ++j, ++it) {
        const uint32_t ph = it & 1;
        if (j + 1 < n_kv) issue_s(it + 1, j + 2 == n_kv);
        mbar_wait(v_full, ph);
        mbar_wait(p_full, ph);
        if (j == 0) mbar_wait(o_free, (n & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < A8_BN / 16; ++k)    // A: 16 keys = 32 B inside P's atom; B: 16 keys = two 8-row groups = 2048 B
            umma_bf16(tO, p_desc + (uint64_t)((C::NP == 2 ? ph : 0) * (A8_P_BYTES >> 4) + k * 2), v_desc + (uint64_t)(k * 128), idesc_o,
                      (j | k) != 0);
          umma_commit(&o_full[ph]);
          umma_commit(v_empty);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== softmax warps (thread = query row) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float sl2 = p.scale_log2;
    uint32_t it = 0, n = 0;
    for (int idx = blockIdx.x; idx < total; idx += gridDim.x, ++n) {
      int prob, qt;
      decode(idx, prob, qt);
      const int head = prob % p.heads, outer = prob / p.heads;
      const int n_kv = kv_tiles(qt);
      const int qi = qt * A8_BM + r;
      float m_run = -INFINITY, l_run = 0.f;

      bool have_next = false;                  // v_next already holds S of the coming tile (and S was handed back)
      uint32_t v_next[2][32];
      for (int j = 0; j < n_kv; ++j, ++it) {
        const uint32_t ph = it & 1;
        const int kv0 = j * A8_BN;
        const bool need_mask = (kv0 + A8_BN > p.Lk) || (p.causal && kv0 + A8_BN - 1 > qt * A8_BM);
        const int k_hi = min(p.Lk - kv0, p.causal ? qi - kv0 + 1 : A8_BN);   // keys [0, k_hi) visible (may be <= 0)
        uint32_t v[2][32];
        if (C::PRE && have_next) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int k = 0; k < 32; ++k) v[c][k] = v_next[c][k];
          have_next = false;
        } else {
          mbar_wait(s_full, ph);
          tc_fence_after();
          tmem_ld_32x32(tS + lane_addr, v[0]);
          tmem_ld_32x32(tS + lane_addr + 32, v[1]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free);
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (need_mask) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (c * 32 + k < k_hi) mx4[2 * c + (k & 1)] = fmaxf(mx4[2 * c + (k & 1)], __uint_as_float(v[c][k]));
        } else {
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int k = 0; k < 32; ++k) mx4[2 * c + (k & 1)] = fmaxf(mx4[2 * c + (k & 1)], __uint_as_float(v[c][k]));
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // the P buffer this tile writes was last read by P V of tile it-NP
        if (C::NP == 2) { if (it >= 2) mbar_wait(&o_full[ph], ((it >> 1) - 1) & 1); }
        else if (it >= 1) mbar_wait(&o_full[ph ^ 1], ((it - 1) >> 1) & 1);
        if (j == 0) {
          m_run = mx;                          // tile 0 always holds key 0, visible to every valid row
        } else if (__any_sync(0xffffffffu, (mx - m_run) * sl2 > 8.0f)) {
          mbar_wait(&o_full[ph ^ 1], ((it - 1) >> 1) & 1);     // O may only be rescaled once P V of the previous tile is done
          tc_fence_after();
          const float m_new = fmaxf(m_run, mx);
          const float corr = exp2f((m_run - m_new) * sl2);
          l_run *= corr;
          m_run = m_new;
#pragma unroll
          for (int c = 0; c < C::DO / 16; ++c) {
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
        const float msc = m_run * sl2;
        uint8_t* const prow = sP + (C::NP == 2 ? ph : 0) * A8_P_BYTES + r * 128;
        float sum = 0.f;
        bool pre_issued = false;
        if (need_mask) {
          const int k_vis = __reduce_max_sync(0xffffffffu, k_hi);   // no row of this warp sees keys >= k_vis
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (c * 32 + 8 * g >= k_vis) {                         // warp-uniform: P = 0, no exp2
                sts128u(prow + (((c * 4 + g) ^ (r & 7)) << 4), make_uint4(0, 0, 0, 0));
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
              sts128u(prow + (((c * 4 + g) ^ (r & 7)) << 4), o);
            }
          }
        } else {
          const f2_t sl22 = f2_pack(sl2, sl2), nmsc2 = f2_pack(-msc, -msc);
          f2_t sum2[2] = {f2_pack(0.f, 0.f), f2_pack(0.f, 0.f)};
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (C::PRE && c == 1 && j + 1 < n_kv) {
              // half of the exp2 pass is done: if S of the next tile is already in TMEM, start pulling it into v_next now
              const bool ready = __shfl_sync(0xffffffffu, (int)mbar_test_wait(s_full, ph ^ 1), 0) != 0;
              if (ready) {
                tc_fence_after();
                tmem_ld_32x32(tS + lane_addr, v_next[0]);
                tmem_ld_32x32(tS + lane_addr + 32, v_next[1]);
                pre_issued = true;
              }
            }
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
              sts128u(prow + (((c * 4 + g) ^ (r & 7)) << 4), make_uint4(o[0], o[1], o[2], o[3]));
            }
          }
          float s0, s1;
          f2_unpack(f2_add(sum2[0], sum2[1]), s0, s1);
          sum = s0 + s1;
        }
        l_run += sum;
        if (C::PRE && pre_issued) {                // S of the next tile has landed in v_next: hand S back to the MMA warp
          tmem_ld_wait_regs(v_next[0]);
          reg_fence32(v_next[1]);
          have_next = true;
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (C::PRE && pre_issued) mbar_arrive(s_free);
          mbar_arrive(p_full);
        }
      }

      // ---- item epilogue: O out of TMEM in 80-column pieces (register budget), handed back to the MMA warp after the last ----
      mbar_wait(&o_full[(it - 1) & 1], ((it - 1) >> 1) & 1);
      tc_fence_after();
      size_t grow = 0;
      if (p.mode == SEER_ATTN_SCTA) {
        const int b = outer / p.nwin, win = outer - b * p.nwin;
        const int wy = win / p.nwx, wx = win - wy * p.nwx;
        const int f = qi / tpf, rem = qi - f * tpf;
        const int iy = rem / p.ws, ix = rem - iy * p.ws;
        grow = ((size_t)(b * p.F + f) * p.H + wy * p.ws + iy) * p.W + wx * p.ws + ix;
      } else {
        grow = (size_t)outer * p.Lq + qi;
      }
      const float inv = 1.0f / l_run;
      __nv_bfloat16* dst = p.o + grow * p.ldo + head * A8_D;
      constexpr int PW = D % 80 == 0 ? 80 : D;          // columns per piece (D = 40: one piece of 40)
#pragma unroll
      for (int piece = 0; piece < D / PW; ++piece) {
        uint32_t v0[32], v1[32], v2[16];
        tmem_ld_32x32(tO + lane_addr + piece * PW, v0);
        if constexpr (PW == 80) {
          tmem_ld_32x32(tO + lane_addr + piece * PW + 32, v1);
          tmem_ld_32x16(tO + lane_addr + piece * PW + 64, v2);
        } else {
          tmem_ld_32x16(tO + lane_addr + piece * PW + 32, v2);      // D = 40: columns 32..47 (40..47 are padding)
        }
        tmem_ld_wait();
        if (piece == D / PW - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_free);
        }
        if (qi < p.Lq) {
#pragma unroll
          for (int g = 0; g < PW / 8; ++g) {
            const uint32_t* src;
            if constexpr (PW == 80) src = g < 4 ? &v0[8 * g] : (g < 8 ? &v1[8 * (g - 4)] : &v2[8 * (g - 8)]);
            else src = g < 4 ? &v0[8 * g] : &v2[8 * (g - 4)];
            uint4 o;
            o.x = pack_bf16(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv);
            o.y = pack_bf16(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv);
            o.z = pack_bf16(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv);
            o.w = pack_bf16(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv);
            *reinterpret_cast<uint4*>(dst + piece * PW + 8 * g) = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncwarp();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// 2-D map over a token-major buffer [rows, ld] restricted to its first `cols` columns; box {64, box_rows}
static int a8_map_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn80 enc = a8_encode_fn();
  if (!enc) return SEER_ENODRIVER;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SEER_OK : SEER_EINVAL;
}
// 4-D map [n_frames, H, W, cols]; box {64 ch, ws, ws, box_frames}
static int a8_map_4d(CUtensorMap* tm, const void* base, uint64_t n_frames, uint64_t H, uint64_t W, uint64_t cols, uint64_t ld,
                     uint32_t ws, uint32_t box_frames) {
  EncodeTiledFn80 enc = a8_encode_fn();
  if (!enc) return SEER_ENODRIVER;
  cuuint64_t dims[4] = {cols, W, H, n_frames};
  cuuint64_t strides[3] = {ld * 2, W * ld * 2, H * W * ld * 2};
  cuuint32_t box[4] = {64, ws, ws, box_frames};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SEER_OK : SEER_EINVAL;
}

// head-split Q maps with zero fill beyond the head ([rows, heads, d] / [frames, H, W, heads, d] views), for d % 16 != 0
static int a8_map_heads_3d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t heads, uint64_t d, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn80 enc = a8_encode_fn();
  if (!enc) return SEER_ENODRIVER;
  cuuint64_t dims[3] = {d, heads, rows};
  cuuint64_t strides[2] = {d * 2, ld * 2};
  cuuint32_t box[3] = {64, 1, box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SEER_OK : SEER_EUNSUPPORTED;
}
static int a8_map_heads_5d(CUtensorMap* tm, const void* base, uint64_t n_frames, uint64_t H, uint64_t W, uint64_t heads, uint64_t d,
                           uint64_t ld, uint32_t ws, uint32_t box_frames) {
  EncodeTiledFn80 enc = a8_encode_fn();
  if (!enc) return SEER_ENODRIVER;
  cuuint64_t dims[5] = {d, heads, W, H, n_frames};
  cuuint64_t strides[4] = {d * 2, ld * 2, W * ld * 2, H * W * ld * 2};
  cuuint32_t box[5] = {64, 1, ws, ws, box_frames};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SEER_OK : SEER_EUNSUPPORTED;
}

// Returns SEER_EUNSUPPORTED when the geometry is not covered (caller falls back to the mma.sync kernel).
template <int D>
static int a8_launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const Attn80Params& p, int n_problems,
                     int nq_tiles, cudaStream_t stream) {
  static SmemAttrOnce smem_attr_attr_done;
  { cudaError_t e = smem_attr_attr_done.ensure(attention_tc80_kernel<D>, A8Cfg<D>::SMEM); if (e != cudaSuccess) return (int)e; }
  int dev = 0, nsm = 148;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0)
    nsm = 148;
  const long total = (long)n_problems * nq_tiles;
  const long resident = (long)A8Cfg<D>::MIN_CTAS * nsm;
  const int grid = (int)(total < resident ? total : resident);
  cudaError_t le = launch_pdl(attention_tc80_kernel<D>, dim3(grid), dim3(A8_THREADS), (size_t)A8Cfg<D>::SMEM, stream, tq, tk, tv, p,
                              n_problems, nq_tiles);
  if (le != cudaSuccess) return (int)le;
  SEER_LAUNCH_CHECK();
  debug_note_attention(D == 40 ? "attention_tc80_kernel<40> tcgen05" : (D == 80 ? "attention_tc80_kernel<80> tcgen05" : "attention_tc80_kernel<160> tcgen05"));
  return SEER_OK;
}

int attention_tc80_launch(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int mode,
                          int heads, int head_dim, int n_outer, int Lq, int Lk, int F, int H, int W, void* stream) {
  if (head_dim != 40 && head_dim != 80 && head_dim != 160) return SEER_EUNSUPPORTED;
  const int A8_D = head_dim;
  if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8) return SEER_EUNSUPPORTED;
  if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o) % 16) return SEER_EUNSUPPORTED;
  Attn80Params p{};
  int n_problems, nq_tiles;
  if (mode == SEER_ATTN_SCTA && H <= 4) {
    // no windows at the 4x4 level (attention.py:661): one causal sequence of F*H*W plain token rows per clip
    if (F <= 0 || H <= 0 || W <= 0) return SEER_EUNSUPPORTED;
    mode = SEER_ATTN_SPATIAL;
    Lq = Lk = F * H * W;
    p.Lq = Lq; p.Lk = Lk; p.causal = 1; p.nwin = 1; p.nwx = 1; p.ws = 1;
    n_problems = n_outer * heads;
  } else if (mode == SEER_ATTN_SCTA) {
    if (F <= 0 || W <= 4) return SEER_EUNSUPPORTED;
    const int ws = (H / 8) >= 4 ? 8 : 4;                 // window rule of attention.py:30-33,661-668
    const int tpf = ws * ws;
    if (H % ws || W % ws) return SEER_EUNSUPPORTED;
    if ((F * tpf) % A8_BN) return SEER_EUNSUPPORTED;     // whole key tiles; a ragged last QUERY tile is fine
    p.F = F; p.H = H; p.W = W; p.ws = ws; p.nwx = W / ws; p.nwin = (H / ws) * p.nwx;
    p.Lq = p.Lk = F * tpf; p.causal = 1;
    n_problems = n_outer * p.nwin * heads;
  } else if (mode == SEER_ATTN_SPATIAL || mode == SEER_ATTN_CROSS) {
    if (Lq <= 0 || Lk <= 0) return SEER_EUNSUPPORTED;
    p.Lq = Lq; p.Lk = Lk; p.causal = 0; p.nwin = 1; p.nwx = 1; p.ws = 1;
    n_problems = n_outer * heads;
  } else {
    return SEER_EUNSUPPORTED;
  }
  nq_tiles = ceil_div(p.Lq, A8_BM);
  p.o = (__nv_bfloat16*)o; p.ldo = ldo;
  p.mode = mode; p.heads = heads;
  p.scale_log2 = (1.0f / sqrtf((float)A8_D)) * 1.4426950408889634f;
  const uint64_t C = (uint64_t)heads * A8_D;
  CUtensorMap tq, tk, tv;
  int rc;
  if (mode == SEER_ATTN_SCTA) {
    const uint64_t nf = (uint64_t)n_outer * F;
    const uint32_t tpf = p.ws * p.ws;
    if (head_dim % 16) { if ((rc = a8_map_heads_5d(&tq, q, nf, H, W, heads, head_dim, ldq, p.ws, A8_BM / tpf))) return rc; }
    else if ((rc = a8_map_4d(&tq, q, nf, H, W, C, ldq, p.ws, A8_BM / tpf))) return rc;
    if ((rc = a8_map_4d(&tk, k, nf, H, W, C, ldk, p.ws, A8_BN / tpf))) return rc;
    if ((rc = a8_map_4d(&tv, v, nf, H, W, C, ldv, p.ws, A8_BN / tpf))) return rc;
  } else {
    if (head_dim % 16) { if ((rc = a8_map_heads_3d(&tq, q, (uint64_t)n_outer * Lq, heads, head_dim, ldq, A8_BM))) return rc; }
    else if ((rc = a8_map_2d(&tq, q, (uint64_t)n_outer * Lq, C, ldq, A8_BM))) return rc;
    if ((rc = a8_map_2d(&tk, k, (uint64_t)n_outer * Lk, C, ldk, A8_BN))) return rc;
    if ((rc = a8_map_2d(&tv, v, (uint64_t)n_outer * Lk, C, ldv, A8_BN))) return rc;
  }
  if (head_dim == 40) return a8_launch<40>(tq, tk, tv, p, n_problems, nq_tiles, (cudaStream_t)stream);
  return head_dim == 80 ? a8_launch<80>(tq, tk, tv, p, n_problems, nq_tiles, (cudaStream_t)stream)
                        : a8_launch<160>(tq, tk, tv, p, n_problems, nq_tiles, (cudaStream_t)stream);
}

}  // namespace seer
