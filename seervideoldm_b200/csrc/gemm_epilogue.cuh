// Epilogue warps of the tcgen05 GEMM (gemm_tc.cu): TMEM accumulator -> (folded LayerNorm | bias | residual | GEGLU) ->
// fp32 / bf16 outputs + GroupNorm column statistics + LayerNorm row statistics.
//
// The K <= 640 launches of the 32x32 and 16x16 levels (q/k/v/out projections, proj_in/out, GEGLU) have a main loop of
// only 5-10 k-blocks per tile, so the epilogue — not the tensor pipe — sets their speed (ncu, profiles/r1: 2300 cycles
// per 32-column chunk, issue slots 30 % busy, 223 SASS instructions per chunk of which most were runtime option tests).
// Hence:
//   * the option set is a TEMPLATE parameter (SPEC): the combinations the UNet issues compile to straight-line code,
//     everything else runs the same body with runtime flags (SPEC = -1);
//   * packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2, sm_100) for the affine, residual, GELU and statistics math;
//   * the TMEM load of the NEXT chunk (also across a tile boundary, when that accumulator is already complete) is issued
//     as soon as the current chunk's registers are consumed, so its latency hides behind the staging / store phase;
//   * the per-tile bias / LayerNorm column-sum slices and per-row LayerNorm partial sums of the NEXT tile are prefetched
//     into registers while the current tile is processed (they used to cost an exposed L2 round trip per tile);
//   * no integer division in the tile loop.
#pragma once
#include "common.cuh"

namespace seer {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;
constexpr int MAX_STAGES = 8;
constexpr int MAX_EPI_WARPS = 8;
constexpr int MAX_RING = 8;
constexpr int GEMM_MAX_THREADS = 64 + 32 * MAX_EPI_WARPS;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int BAR_BYTES = 1024;
constexpr int EVEC_FLOATS = 256;                       // per-warp staging of the tile's bias / LN column-sum slices (BN <= 256) ...
constexpr int EVEC_FLOATS_320 = 320;                   // ... and for the 320-wide tiles (GemmParams::evec_floats says which)
constexpr int EVEC_BYTES_PER_WARP = 2 * EVEC_FLOATS * 4;
constexpr int EVEC_BYTES_PER_WARP_320 = 2 * EVEC_FLOATS_320 * 4;
constexpr int BIAS_ONE_ROW = 0x7fffffff;               // GemmParams::bias_div value meaning "a single bias row"

// epilogue option bits (template parameter SPEC of the epilogue; -1 = decide at run time from GemmParams)
enum : int { EF_LN = 1, EF_GEGLU = 2, EF_RES32 = 4, EF_OUT32 = 8, EF_OUT16 = 16, EF_CSTAT = 32, EF_RSTAT = 64, EF_RES16 = 128,
             EF_ROPE = 256 };
// the combinations one UNet evaluation issues (all with a bias vector)
constexpr int EK_PIN = EF_OUT32 | EF_OUT16 | EF_RSTAT;               // proj_in: fp32 + bf16 token stream, LN row sums
constexpr int EK_QKV = EF_LN | EF_OUT16;                              // LN-folded q / qkv projections
constexpr int EK_ATTN_OUT = EF_RES32 | EF_OUT32 | EF_OUT16 | EF_RSTAT;  // to_out + residual
constexpr int EK_FF1 = EF_LN | EF_GEGLU | EF_OUT16;                   // LN-folded GEGLU projection
constexpr int EK_FF1_PLAIN = EF_GEGLU | EF_OUT16;                     // GEGLU without the fold
constexpr int EK_FF2 = EF_RES32 | EF_OUT16;                           // FF out + residual -> bf16
constexpr int EK_POUT = EF_RES32 | EF_OUT32 | EF_CSTAT;               // proj_out / conv2 + residual, GroupNorm column sums
constexpr int EK_CONV = EF_OUT32 | EF_CSTAT;                          // conv1 / conv2+shortcut / resample convs
constexpr int EK_BF16 = EF_OUT16;                                     // plain bf16 projection
// bf16 token stream inside a transformer block (the reference's own autocast dtype, attention.py:231-248,308-327): the
// residual is read and written as bf16, fp32 only at the block boundary (proj_out + the block input)
constexpr int EK_PIN16 = EF_OUT16 | EF_RSTAT;                         // proj_in: bf16 token stream + LN row sums
constexpr int EK_ATTN_OUT16 = EF_RES16 | EF_OUT16 | EF_RSTAT;         // to_out + bf16 residual
constexpr int EK_FF2_16 = EF_RES16 | EF_OUT16;                        // FF out + bf16 residual
constexpr int EK_CONV16 = EF_OUT16 | EF_CSTAT;                        // conv1: bf16 out (only GroupNorm 2 reads it) + column sums
// bf16 residual stream BETWEEN blocks (SeerUNet.residual_stream = "bf16"): block outputs are stored once, as bf16, next to the
// GroupNorm column sums of the fp32 values; the block input is read back as a bf16 residual
constexpr int EK_POUT16 = EF_RES16 | EF_OUT16 | EF_CSTAT;             // proj_out / conv2 + bf16 residual -> bf16 + column sums
// SCTA's q/k/v projection: LayerNorm fold + rotary embedding of the Q and K heads (attention.py:649-651) applied to the fp32
// accumulators before the single bf16 rounding — the separate read-modify-write RoPE pass over [M, 2C] disappears
constexpr int EK_QKV_ROPE = EF_LN | EF_OUT16 | EF_ROPE;
constexpr int ROPE_PAIRS = 16;                                        // rotary dim 32 = 16 (cos, sin) pairs per row (rotary_emb dim=32)
constexpr int ROPE_BYTES_PER_WARP = 2 * 32 * ROPE_PAIRS * 4;          // double-buffered [32 rows][16 x half2 (cos, sin)] per epilogue warp

struct GemmParams {
  int M, N;            // N = accumulator columns (GEGLU: twice the output columns)
  int mode;            // 0 plain, 1 conv3x3
  int kb_main;         // k-blocks (of 64) from the main source
  int kb_total;        // + k-blocks from the tail source
  int cblk;            // conv: Cin / 64
  int H, W;            // conv image geometry
  int tiles_n, num_tiles;
  int stages, nepi, ring, slot_bytes;
  int ntaps, taps_w, taps_h, off_x, off_y, cstride;   // conv taps: tap t reads pixel (s*y + t / taps_w + off_y, s*x + t % taps_w + off_x)
  int up_phase;        // 0: output row = GEMM row; 1 + (2 py + px): rows are the (py, px) phase of a nearest-2x upsampled image
  int up_wshift;       //    (low-res width = 1 << up_wshift): out row = ((m >> ws) << (ws + 2)) + py * 2W + 2 (m & (W - 1)) + px
  int l2_prefetch;     // > 0: the producer prefetches the A rows / residual tile of the tile this many iterations ahead into L2
  int evec_floats;     // floats per staged per-tile vector (bias / LN column sums) and warp: EVEC_FLOATS or EVEC_FLOATS_320
  int bstat;           // 1: the whole Wt panel of this CTA's (fixed) n-block is resident in smem; only A is streamed
  int epi_spec;        // EK_* combination compiled as a specialisation, or -1 (generic runtime-flag epilogue)
  const float* bias;
  int ldb;
  int bias_div;
  int res_mode;        // 0 none, 1 fp32, 2 bf16
  float* out_f32;      // fp32 output (or null), leading dim ldo_f32 elements
  int ldo_f32;
  __nv_bfloat16* out_bf16;
  int ldo_bf16;
  int geglu;
  float* col_stats;
  float* row_stats_out;
  const float* row_stats_in;
  int row_parts_in;
  float ln_inv_dim, ln_eps;
  const float* ln_colsum;
  // rotary embedding in the epilogue (EK_QKV_ROPE): rope_tab[pos][16] = half2 (cos, sin)(pos * freq_j), pos = row % rope_T;
  // output columns [0, rope_cols) are heads of width rope_d whose first 32 channels rotate as interleaved pairs.  (cos, sin)
  // in fp16: 11 significant bits in [-1, 1] — a quarter of the bf16 rounding of the rotated output, half the smem / L2 bytes
  const __half2* rope_tab;
  int rope_T, rope_cols, rope_d;
};

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA PAIR (cluster of 2, tcgen05 cta_group::2) per 256 x BN tile — each
// CTA stages its own 128 A rows and only HALF of the Wt rows, so the L2 -> SM operand traffic per FLOP drops by
// 64*(128+BN)/BN -> 64*(128+BN/2)/BN bytes per MMA cycle (the measured limiter of the 1-CTA kernel, profiles/).
// BN = 320 (N = 320 / 640 launches with a long K: the level-0 / level-1 3x3 convs and FF-out GEMMs): the whole 512-column
// TMEM holds ONE 256 x 320 pair accumulator — each A stage is multiplied by two N = 160 UMMAs, so the A bytes a CTA pulls from
// L2 per FLOP halve against 160-wide tiles (the measured limiter of those launches: tensor pipe 63 %, L2 -> SM 47 B/clk/SM).
// The price is a single-buffered accumulator (the epilogue of tile i no longer overlaps the main loop of tile i + 1), which
// is why the planner only picks it when the main loop is long (K >= 1280).
template <int BN, int CG>
struct GemmCfg {
  static constexpr int B_BYTES = (BN / CG) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NBUF = BN <= 256 ? 2 : 1;                          // TMEM accumulator buffers
  static constexpr int NMMA = BN <= 256 ? 1 : 2;                          // UMMAs per k-step, N = BN / NMMA each
  static constexpr int TBUF = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);   // TMEM column stride between the 2 buffers
  static constexpr int TMEM_COLS = BN <= 256 ? 2 * TBUF : 512;
};

// byte offset of 16-byte chunk `j` of row `r` inside a TMA-swizzled box whose rows are 128 B / 64 B wide
__device__ __forceinline__ int sw128(int r, int j) { return r * 128 + ((j ^ (r & 7)) << 4); }
__device__ __forceinline__ int sw64(int r, int j) { return r * 64 + ((j ^ ((r >> 1) & 3)) << 4); }

__device__ __forceinline__ uint32_t f2_to_bf16x2(f2_t v) {
  float lo, hi;
  f2_unpack(v, lo, hi);
  return pack_bf16(lo, hi);
}
// 16 bytes of shared memory as two packed pairs
__device__ __forceinline__ void lds128_f2(const void* p, f2_t& a, f2_t& b) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void sts128_f2(void* p, f2_t a, f2_t b) {
  asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(smem_u32(p)), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ldg128_f2(const float* p, f2_t& a, f2_t& b) {
  asm volatile("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
}

// erf-GELU of a pair through the clamp-free tanh form gelu(x) = hx + hx tanh(x (0.79788456 + 0.03567741 x^2)), hx = x / 2:
// 1.8e-4 rel-L2 from the exact erf form on N(0, 1.5) gates (max abs 4.7e-4), an order of magnitude below the bf16 rounding of the
// stored product.  Its cubic argument is monotone, so — unlike the quintic erf fit of common.cuh's gelu_erf_tanhfit, whose
// leading coefficient turns negative beyond |z| ~ 7.6 — it needs no clamp: 5 packed fp32x2 instructions + 2 MUFU.TANH per
// pair.  Measured against the quintic in the same build (profiles/r2_gemm_probe.txt): level-0 GEGLU launch 468 -> 447 us.
__device__ __forceinline__ f2_t f2_gelu_erf(f2_t x) {
  const f2_t x2 = f2_mul(x, x);
  const f2_t pz = f2_fma(x2, f2_pack(0.0356774081f, 0.0356774081f), f2_pack(0.7978845608f, 0.7978845608f));
  float a, b, ta, tb;
  f2_unpack(f2_mul(x, pz), a, b);
  asm("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(a));
  asm("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(b));
  const f2_t hx = f2_mul(x, f2_pack(0.5f, 0.5f));
  return f2_fma(hx, f2_pack(ta, tb), hx);
}

// tcgen05.wait::ld that also names the destination registers of the outstanding load, so no use of them can be
// scheduled above the wait (the load is issued a whole chunk earlier than it is consumed)
// f[k] (pairs of one 32-column accumulator chunk) = rstd * (acc - mean * colsum) + bias | acc + bias | acc.
// BSRC: 0 none, 1 shared memory (staged slice), 2 global row pointer (rows of the warp span two bias rows — rare)
template <bool LN, int BSRC>
__device__ __forceinline__ void epi_affine(f2_t (&f)[16], const uint32_t (&v)[32], const float* cs, const float* bsm, const float* bgl,
                                           f2_t rstd2, f2_t nmean2) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    f2_t x0 = f2_pack_u(v[4 * k], v[4 * k + 1]), x1 = f2_pack_u(v[4 * k + 2], v[4 * k + 3]);
    f2_t b0 = 0, b1 = 0;
    if (BSRC == 1) lds128_f2(bsm + 4 * k, b0, b1);
    if (BSRC == 2) ldg128_f2(bgl + 4 * k, b0, b1);
    if (LN) {
      f2_t c0, c1;
      lds128_f2(cs + 4 * k, c0, c1);
      x0 = f2_fma(rstd2, f2_fma(nmean2, c0, x0), b0);
      x1 = f2_fma(rstd2, f2_fma(nmean2, c1, x1), b1);
    } else if (BSRC != 0) {
      x0 = f2_add(x0, b0);
      x1 = f2_add(x1, b1);
    }
    f[2 * k] = x0;
    f[2 * k + 1] = x1;
  }
}
__device__ __forceinline__ void epi_affine_dispatch(f2_t (&f)[16], const uint32_t (&v)[32], bool ln, bool bias, bool bias_smem,
                                                    const float* cs, const float* bsm, const float* bgl, f2_t rstd2, f2_t nmean2) {
  if (ln) {
    if (!bias) epi_affine<true, 0>(f, v, cs, bsm, bgl, rstd2, nmean2);
    else if (bias_smem) epi_affine<true, 1>(f, v, cs, bsm, bgl, rstd2, nmean2);
    else epi_affine<true, 2>(f, v, cs, bsm, bgl, rstd2, nmean2);
  } else if (bias) {
    if (bias_smem) epi_affine<false, 1>(f, v, cs, bsm, bgl, rstd2, nmean2);
    else epi_affine<false, 2>(f, v, cs, bsm, bgl, rstd2, nmean2);
  } else {
    epi_affine<false, 0>(f, v, cs, bsm, bgl, rstd2, nmean2);
  }
}

// One epilogue warp's whole persistent loop.  `warp` = warp index in the CTA (epilogue warps are 2 .. 2 + nepi).
template <int BN, int CG, int SPEC>
__device__ __forceinline__ void gemm_epilogue_warp(const GemmParams& p, const CUtensorMap* tmRes, uint8_t* ring_base,
                                                   uint64_t* tmem_full_bar, uint64_t* tmem_empty_bar, uint64_t* res_full_bar,
                                                   float* evec_base, uint32_t tmem_base, int warp, int lane, int rank, int unit,
                                                   int nunits) {
  using C = GemmCfg<BN, CG>;
  constexpr bool S = SPEC >= 0;
  constexpr bool ROPE = S && (SPEC & EF_ROPE) != 0;        // only as a specialisation (the host rejects anything else)
  constexpr bool GEGLU = S && (SPEC & EF_GEGLU) != 0;      // the host routes every GEGLU launch to a specialisation
  const bool f_ln = S ? (SPEC & EF_LN) != 0 : p.row_stats_in != nullptr;
  const int res_mode = S ? ((SPEC & EF_RES32) ? 1 : ((SPEC & EF_RES16) ? 2 : 0)) : p.res_mode;
  const bool f_o32 = S ? (SPEC & EF_OUT32) != 0 : p.out_f32 != nullptr;
  const bool f_o16 = S ? (SPEC & EF_OUT16) != 0 : p.out_bf16 != nullptr;
  const bool f_cst = S ? (SPEC & EF_CSTAT) != 0 : p.col_stats != nullptr;
  const bool f_rst = S ? (SPEC & EF_RSTAT) != 0 : p.row_stats_out != nullptr;
  const bool f_bias = S ? true : p.bias != nullptr;

  const int ew = warp - 2;
  const int q = warp & 3;                  // TMEM lane quarter this warp may access
  const int nhalf = p.nepi >> 2;           // 1 or 2 warps per quarter; they interleave the column chunks
  const int half = ew >> 2;
  constexpr int CW = GEGLU ? 64 : 32;      // accumulator columns per chunk (always 32 output columns)
  constexpr int NCH = BN / CW;
  constexpr int NV = BN / 32;              // bias / column-sum floats per lane and tile
  const int my_nch = (NCH - half + nhalf - 1) / nhalf;
  const int my_tiles = (p.num_tiles - unit + nunits - 1) / nunits;
  const int total = my_tiles * my_nch;
  const uint32_t tempty0 = CG == 2 ? mapa_shared(smem_u32(tmem_empty_bar), 0) : 0;   // the leader's tmem_empty barriers
  // Ring of R smem slots per warp.  A slot first receives the TMA-prefetched residual chunk (32 rows x 32 cols),
  // then stages the output chunk for the coalesced copy-out; it is free again as soon as the warp has read it back,
  // so the residual for step g + R - 1 can be requested at the top of step g.
  const int R = p.ring, P = p.ring - 1;
  uint8_t* ring = ring_base + (size_t)ew * R * p.slot_bytes;
  uint64_t* rfull = res_full_bar + ew * MAX_RING;
  const uint32_t res_bytes = res_mode == 1 ? 4096u : 2048u;
  float* vb = evec_base + ew * (2 * p.evec_floats);
  float* vc = vb + p.evec_floats;
  // RoPE: the (cos, sin) rows of this warp's 32 token rows live in a double-buffered per-warp smem area (64 B per row, 16-byte
  // quads XOR-swizzled by the row: the per-lane 16-byte reads are bank-conflict free).  The 32 rows are consecutive positions of
  // one clip (rope_T % 32 == 0), i.e. ONE contiguous 2 KB block of the table: it is fetched with fully coalesced cp.async (a
  // lane copies pieces of other lanes' rows) a WHOLE TILE ahead — measured on the way here: per-lane 16-byte pieces of 32
  // different rows cost 32 L1 wavefronts per instruction and doubled the kernel's time; a single buffer refilled at the tile
  // switch exposed the full L2 / DRAM round trip (~2 us) on every tile.
  uint8_t* rope_buf = reinterpret_cast<uint8_t*>(evec_base + MAX_EPI_WARPS * 2 * p.evec_floats) + ew * ROPE_BYTES_PER_WARP;
  auto rope_prefetch = [&](int t_mb, int bufi) {
    const int r0 = (t_mb * CG + rank) * BM + q * 32;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.rope_tab + (size_t)(r0 % p.rope_T) * ROPE_PAIRS);
    uint8_t* dst = rope_buf + bufi * (ROPE_BYTES_PER_WARP / 2);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = i * 32 + lane;           // 16-byte piece k of the block = quad (k & 3) of row (k >> 2)
      cp_async_16(dst + sw64(k >> 2, k & 3), src + k * 16, true);
    }
    cp_async_commit();
  };

  // tile -> (m block, n block), advanced incrementally (tile += nunits)
  const int dm = nunits / p.tiles_n, dn = nunits - dm * p.tiles_n;
  int mb = unit / p.tiles_n, nb = unit - mb * p.tiles_n;

  // residual prefetch cursor (tile / chunk of step g + P)
  int pf_step = 0, pf_slot = 0, pf_mb = mb, pf_nb = nb, pf_j = 0;
  auto issue_res = [&]() {                 // lane 0: TMA-prefetch the residual chunk of step pf_step
    mbar_arrive_expect_tx(&rfull[pf_slot], res_bytes);
    tma_load_2d(ring + pf_slot * p.slot_bytes, tmRes, &rfull[pf_slot], pf_nb * BN + (half + pf_j * nhalf) * 32,
                (pf_mb * CG + rank) * BM + q * 32);
  };
  auto advance_pf = [&]() {
    ++pf_step;
    if (++pf_slot == R) pf_slot = 0;
    if (++pf_j == my_nch) {
      pf_j = 0;
      pf_mb += dm;
      pf_nb += dn;
      if (pf_nb >= p.tiles_n) { pf_nb -= p.tiles_n; ++pf_mb; }
    }
  };
  if (res_mode) {
    for (int st = 0; st < P && st < total; ++st) {
      if (lane == 0) issue_res();
      advance_pf();
    }
    __syncwarp();
  }

  // ---- per-tile vectors, prefetched one tile ahead into registers ----
  constexpr int PF_PARTS = 4;              // LayerNorm partial row sums prefetched (the rest, if any, are read in place)
  float nbv[NV], ncv[NV];
  float2 nrs[PF_PARTS];
  bool n_bias_smem = false;
  const float* n_bias_row = nullptr;       // this lane's bias row when the warp's rows span two bias rows
  auto prefetch_vecs = [&](int t_mb, int t_nb) {
    const int row0 = (t_mb * CG + rank) * BM + q * 32;
    const int n0 = t_nb * BN;
    if (f_bias) {
      int brow = 0;
      n_bias_smem = row0 < p.M;
      if (p.bias_div != BIAS_ONE_ROW) {
        brow = row0 / p.bias_div;
        const int rlast = min(row0 + 31, p.M - 1);
        n_bias_smem = n_bias_smem && (brow == rlast / p.bias_div);
        n_bias_row = p.bias + (size_t)(min(row0 + lane, p.M - 1) / p.bias_div) * p.ldb;
      } else {
        n_bias_row = p.bias;
      }
      if (n_bias_smem) {
        const float* src = p.bias + (size_t)brow * p.ldb + n0 + lane;
#pragma unroll
        for (int i = 0; i < NV; ++i) nbv[i] = __ldg(src + 32 * i);
      }
    }
    if (f_ln) {
      const float* src = p.ln_colsum + n0 + lane;
#pragma unroll
      for (int i = 0; i < NV; ++i) ncv[i] = __ldg(src + 32 * i);
      const int row = min(row0 + lane, p.M - 1);
#pragma unroll
      for (int i = 0; i < PF_PARTS; ++i)
        nrs[i] = i < p.row_parts_in ? __ldg(reinterpret_cast<const float2*>(p.row_stats_in) + (size_t)i * p.M + row)
                                    : make_float2(0.f, 0.f);
    }
  };
  prefetch_vecs(mb, nb);
  if constexpr (ROPE) rope_prefetch(mb, 0);

  auto tmem_addr = [&](int buf, int c) {
    return tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * C::TBUF + c * CW);
  };
  uint32_t v[32], vg[GEGLU ? 32 : 1];
  auto issue_ld = [&](uint32_t taddr) {
    tmem_ld_32x32(taddr, v);
    if constexpr (GEGLU) tmem_ld_32x32(taddr + 32, reinterpret_cast<uint32_t(&)[32]>(vg));
  };
  auto wait_ld = [&]() {
    tmem_ld_wait_regs(v);
    if constexpr (GEGLU) reg_fence32(reinterpret_cast<uint32_t(&)[32]>(vg));
  };
  auto release_acc = [&](int b) {          // this warp has read its last chunk of accumulator buffer b: hand it back
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (CG == 2) mbar_arrive_cluster(tempty0 + (uint32_t)b * 8u); else mbar_arrive(&tmem_empty_bar[b]);
    }
  };
  // `have`: v (and vg) already hold the accumulator chunk of the current step.  The load is issued one step early —
  // right after the previous chunk's registers were consumed — and completed (tcgen05.wait::ld) at the END of that
  // step, behind its staging / store phase, so no asynchronously written register is live across the loop edge.
  bool have = false;
  // output row of GEMM row m: identity, or the (py, px) phase rows of the nearest-2x upsampled image (Upsample3D folded
  // into four 2x2-tap convs on the low-res image, resnet.py:52)
  const int up_ws = p.up_wshift, up_mask = (1 << p.up_wshift) - 1;
  const int up_off = p.up_phase ? (((p.up_phase - 1) >> 1) << (p.up_wshift + 1)) + ((p.up_phase - 1) & 1) : 0;
  auto out_row = [&](int m) -> size_t {
    return p.up_phase ? (size_t)(((m >> up_ws) << (up_ws + 2)) + up_off + ((m & up_mask) << 1)) : (size_t)m;
  };

  int g = 0, it = 0, slot_i = 0;
  uint32_t slot_par = 0;                   // parity of the residual barrier of slot `slot_i` = (g / R) & 1
  for (int tile = unit; tile < p.num_tiles; tile += nunits, ++it) {
    const int m0 = (mb * CG + rank) * BM, n0 = nb * BN;
    const int buf = C::NBUF == 2 ? (it & 1) : 0;
    const uint32_t buf_par = C::NBUF == 2 ? (((uint32_t)it >> 1) & 1) : ((uint32_t)it & 1);   // phase parity of this buffer's barriers
    const int row0 = m0 + q * 32;
    const int row = row0 + lane;
    const bool row_ok = row < p.M;
    // warp-uniform: the fast copy-out form applies.  Not in the GEGLU instantiation: the second code path costs it 72 B more
    // spills and 35 % of its speed (537 -> 728 us at level 0)
#ifdef SEER_NO_FAST_COPYOUT
    const bool plain_rows = false;
#else
    const bool plain_rows = !GEGLU && row0 + 32 <= p.M && p.up_phase == 0;
#endif
    // ---- this tile's vectors: registers -> shared memory; LayerNorm mean / rstd of this lane's row ----
    const bool bias_smem = n_bias_smem;
    const float* bias_row = n_bias_row;
    __syncwarp();                          // every lane is done with the previous tile's slices
    if (f_bias && bias_smem) {
#pragma unroll
      for (int i = 0; i < NV; ++i) sts32(vb + 32 * i + lane, nbv[i]);
    }
    float mean = 0.f, rstd = 1.f;
    if (f_ln) {
#pragma unroll
      for (int i = 0; i < NV; ++i) sts32(vc + 32 * i + lane, ncv[i]);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < PF_PARTS; ++i) { s1 += nrs[i].x; s2 += nrs[i].y; }
      for (int i = PF_PARTS; i < p.row_parts_in; ++i) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p.row_stats_in) + (size_t)i * p.M + min(row, p.M - 1));
        s1 += t.x;
        s2 += t.y;
      }
      mean = s1 * p.ln_inv_dim;
      rstd = rsqrtf(fmaxf(s2 * p.ln_inv_dim - mean * mean, 0.f) + p.ln_eps);
    }
    __syncwarp();
    const f2_t rstd2 = f2_pack(rstd, rstd), nmean2 = f2_pack(-mean, -mean);
    // next tile's coordinates + prefetch of its vectors (in flight during this tile's chunks)
    int mb_n = mb + dm, nb_n = nb + dn;
    if (nb_n >= p.tiles_n) { nb_n -= p.tiles_n; ++mb_n; }
    const bool has_next = tile + nunits < p.num_tiles;
    if (has_next) prefetch_vecs(mb_n, nb_n);
    // channel (inside its head) of the first column of this warp's first chunk; advanced incrementally per chunk
    int rope_within = 0;
    if constexpr (ROPE) {
      rope_within = (n0 + half * 32) % p.rope_d;
    }
    const uint8_t* rope_cur = rope_buf + (it & 1) * (ROPE_BYTES_PER_WARP / 2);
    if constexpr (ROPE) {
      // the other buffer was last read during the previous tile (program order + the __syncwarp below cover every lane):
      // refill it for the NEXT tile now, then make sure THIS tile's block (issued one tile ago) has landed
      __syncwarp();
      if (has_next) { rope_prefetch(mb_n, (it & 1) ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
      __syncwarp();                        // pieces were copied by other lanes
    }

    f2_t rs2 = 0, rq2 = 0;                 // packed (sum, sumsq) accumulators of this lane's row (bit pattern 0 = +0.f pair)

    for (int j = 0; j < my_nch; ++j, ++g) {
      const int c = half + j * nhalf;
      uint8_t* slot = ring + slot_i * p.slot_bytes;
      if (res_mode && pf_step < total) {   // slot (g+P)%R = (g-1)%R was fully consumed in the previous step
        if (lane == 0) issue_res();
        advance_pf();
        __syncwarp();
      }
      // RoPE: the chunk's (cos, sin) quads are requested BEFORE the accumulator wait / affine, so the shared-memory latency is
      // off the chunk's dependent chain (the epilogue warps are latency-bound: two warps per scheduler)
      uint4 rope_q[4] = {};
      bool rope_on[4] = {false, false, false, false};
      if constexpr (ROPE) {
        if (n0 + c * 32 < p.rope_cols) {   // warp-uniform
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) { // 8-column groups = 4 rotary pairs = one 16-byte quad of half2 (cos, sin)
            int w = rope_within + 8 * g8;  // channel of the group's first column inside its head (a multiple of 8)
            if (w >= p.rope_d) w -= p.rope_d;
            rope_on[g8] = w < 2 * ROPE_PAIRS;      // warp-uniform: channels >= 32 of a head pass through
            if (rope_on[g8]) rope_q[g8] = lds128u(rope_cur + sw64(lane, w >> 3));
          }
        }
      }
      if (!have) {                         // first chunk of a tile whose accumulator was not complete one step ago
        mbar_wait(&tmem_full_bar[buf], buf_par);
        tc_fence_after();
        issue_ld(tmem_addr(buf, c));
        wait_ld();
        if (j == my_nch - 1) release_acc(buf);
      }
      have = false;
      f2_t f[16];
      if constexpr (!GEGLU) {
        epi_affine_dispatch(f, v, f_ln, f_bias, bias_smem, vc + c * 32, vb + c * 32, bias_row + n0 + c * 32, rstd2, nmean2);
        if constexpr (ROPE) {
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            if (rope_on[g8]) {
              const uint32_t qw[4] = {rope_q[g8].x, rope_q[g8].y, rope_q[g8].z, rope_q[g8].w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 cs = __half22float2(*reinterpret_cast<const __half2*>(&qw[i]));     // (cos, sin) of pair i
                float x0, x1;
                f2_unpack(f[4 * g8 + i], x0, x1);
                f[4 * g8 + i] = f2_pack(fmaf(x0, cs.x, -(x1 * cs.y)), fmaf(x1, cs.x, x0 * cs.y));
              }
            }
          }
          rope_within += 32 * nhalf;                          // next chunk of this warp
          while (rope_within >= p.rope_d) rope_within -= p.rope_d;
        }
      } else {
        // GEGLU: accumulator columns [c*64, +32) are "value", [c*64+32, +64) the matching "gate" (bias is staged)
        f2_t fg[16];
        if (f_ln) {
          epi_affine<true, 1>(f, v, vc + c * 64, vb + c * 64, nullptr, rstd2, nmean2);
          epi_affine<true, 1>(fg, reinterpret_cast<const uint32_t(&)[32]>(vg), vc + c * 64 + 32, vb + c * 64 + 32, nullptr, rstd2, nmean2);
        } else {
          epi_affine<false, 1>(f, v, vc + c * 64, vb + c * 64, nullptr, rstd2, nmean2);
          epi_affine<false, 1>(fg, reinterpret_cast<const uint32_t(&)[32]>(vg), vc + c * 64 + 32, vb + c * 64 + 32, nullptr, rstd2, nmean2);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) f[k] = f2_mul(f[k], f2_gelu_erf(fg[k]));
      }
      // ---- the accumulator registers are consumed: start the TMEM load of the next chunk now ----
      bool issued = false;
      int rel_buf = -1;
      if (j + 1 < my_nch) {
        issue_ld(tmem_addr(buf, c + nhalf));
        issued = true;
        if (j + 2 == my_nch) rel_buf = buf;
      } else if (C::NBUF == 2 && has_next) {
        const int nbuf = buf ^ 1;
        if (__all_sync(0xffffffffu, mbar_try_wait(&tmem_full_bar[nbuf], (((uint32_t)it + 1u) >> 1) & 1))) {
          tc_fence_after();
          issue_ld(tmem_addr(nbuf, half));
          issued = true;
          if (my_nch == 1) rel_buf = nbuf;
        }
      }
      // ---- residual (prefetched into the slot by TMA) ----
      if (res_mode) {
        mbar_wait(&rfull[slot_i], slot_par);
        if (res_mode == 1) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            f2_t t0, t1;
            lds128_f2(slot + sw128(lane, k), t0, t1);
            f[2 * k] = f2_add(f[2 * k], t0);
            f[2 * k + 1] = f2_add(f[2 * k + 1], t1);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint4 t = lds128u(slot + sw64(lane, k));
            const float2 a = unpack_bf16(t.x), b = unpack_bf16(t.y), cc = unpack_bf16(t.z), d = unpack_bf16(t.w);
            f[4 * k] = f2_add(f[4 * k], f2_pack(a.x, a.y));
            f[4 * k + 1] = f2_add(f[4 * k + 1], f2_pack(b.x, b.y));
            f[4 * k + 2] = f2_add(f[4 * k + 2], f2_pack(cc.x, cc.y));
            f[4 * k + 3] = f2_add(f[4 * k + 3], f2_pack(d.x, d.y));
          }
        }
      }
      if (f_cst || f_rst) {
        if (!row_ok) {
#pragma unroll
          for (int k = 0; k < 16; ++k) f[k] = 0;
        }
        if (f_rst) {
#pragma unroll
          for (int k = 0; k < 16; ++k) { rs2 = f2_add(rs2, f[k]); rq2 = f2_fma(f[k], f[k], rq2); }
        }
      }
      // ---- outputs: stage the chunk in the slot (thread = row, XOR-swizzled 16-byte pieces: conflict-free), then
      // copy it out with coalesced 128-bit global stores (each warp store covers 4 full 128-byte lines).  This stays
      // in the generic proxy: a TMA store would need fence.proxy.async + a bulk-group round trip per chunk, measured
      // at ~1000 cycles of serial latency per warp (tools/tma_store_bench.cu).
      const int ocol = (GEGLU ? (n0 >> 1) : n0) + c * 32;
      __syncwarp();                          // every lane has read its residual row: the slot may be overwritten
      // bf16-only output with column sums: the sums are taken from the staged bf16 tile below (the statistics of exactly the
      // tensor the consuming GroupNorm reads) — no second, fp32 staging round trip on the warp's dependent chain
      const bool cst16 = f_cst && f_o16 && !f_o32;
      if (f_o32 || (f_cst && !cst16)) {
#pragma unroll
        for (int k = 0; k < 8; ++k) sts128_f2(slot + sw128(lane, k), f[2 * k], f[2 * k + 1]);
        __syncwarp();
        if (f_o32) {
          if (plain_rows) {
            // whole 32-row group inside M, output row = GEMM row: one pointer per chunk, constant row stride, no per-store
            // bounds test / row mapping (the general form below costs ~15 instructions per store, this one 3)
            float* d0 = p.out_f32 + (size_t)(row0 + (lane >> 3)) * p.ldo_f32 + ocol + (lane & 7) * 4;
            const size_t rstep = (size_t)4 * p.ldo_f32;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<float4*>(d0 + i * rstep) = lds128(slot + sw128(4 * i + (lane >> 3), lane & 7));
          } else {
            float* dst = p.out_f32 + ocol + (lane & 7) * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = 4 * i + (lane >> 3);
              const float4 t = lds128(slot + sw128(r, lane & 7));
              if (row0 + r < p.M) *reinterpret_cast<float4*>(dst + out_row(row0 + r) * p.ldo_f32) = t;
            }
          }
        }
        if (f_cst) {
          // lane = column: (sum, sumsq) over this warp's 32 rows, read back from the staged fp32 tile
          const uint8_t* src = slot + (lane & 3) * 4;
          float cs = 0.f, cq = 0.f;
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const float t = lds32(src + sw128(r, lane >> 2));
            cs += t;
            cq = fmaf(t, t, cq);
          }
          if (row0 < p.M) {
            const size_t slab = p.up_phase ? (((size_t)(row0 >> 5) << 2) + (size_t)(p.up_phase - 1)) : (size_t)(row0 >> 5);
            reinterpret_cast<float2*>(p.col_stats)[slab * p.N + ocol + lane] = make_float2(cs, cq);
          }
        }
        __syncwarp();
      }
      if (f_o16) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint4 o;
          o.x = f2_to_bf16x2(f[4 * k]);
          o.y = f2_to_bf16x2(f[4 * k + 1]);
          o.z = f2_to_bf16x2(f[4 * k + 2]);
          o.w = f2_to_bf16x2(f[4 * k + 3]);
          sts128u(slot + sw64(lane, k), o);
        }
        __syncwarp();
        if (plain_rows) {
          __nv_bfloat16* d0 = p.out_bf16 + (size_t)(row0 + (lane >> 2)) * p.ldo_bf16 + ocol + (lane & 3) * 8;
          const size_t rstep = (size_t)8 * p.ldo_bf16;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(d0 + i * rstep) = lds128u(slot + sw64(8 * i + (lane >> 2), lane & 3));
        } else {
          __nv_bfloat16* dst = p.out_bf16 + ocol + (lane & 3) * 8;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = 8 * i + (lane >> 2);
            const uint4 t = lds128u(slot + sw64(r, lane & 3));
            if (row0 + r < p.M) *reinterpret_cast<uint4*>(dst + out_row(row0 + r) * p.ldo_bf16) = t;
          }
        }
        if (cst16) {
          // lane = (column pair cp, row parity h): (sum, sumsq) of columns 2cp, 2cp+1 over rows h, h+2, ... (adjacent rows sit 64 B
          // apart: the two half-warps hit disjoint banks), halves combined with one shuffle; rows >= M were zeroed above
          const int cp = lane & 15, h = lane >> 4;
          float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int r = 2 * i + h;
            const float2 t = unpack_bf16(lds32u(slot + sw64(r, cp >> 2) + (cp & 3) * 4));
            s0 += t.x; q0 = fmaf(t.x, t.x, q0);
            s1 += t.y; q1 = fmaf(t.y, t.y, q1);
          }
          s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
          q0 += __shfl_xor_sync(0xffffffffu, q0, 16);
          s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
          q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
          if (h == 0 && row0 < p.M) {
            const size_t slab = p.up_phase ? (((size_t)(row0 >> 5) << 2) + (size_t)(p.up_phase - 1)) : (size_t)(row0 >> 5);
            reinterpret_cast<float4*>(p.col_stats)[(slab * p.N + ocol) / 2 + cp] = make_float4(s0, q0, s1, q1);
          }
        }
        __syncwarp();
      }
      if (++slot_i == R) { slot_i = 0; slot_par ^= 1; }
      if (issued) {                        // the next chunk has landed behind the store phase
        wait_ld();
        have = true;
        if (rel_buf >= 0) release_acc(rel_buf);
      }
    }
    if (f_rst && row_ok) {
      float a, b, cq0, cq1;
      f2_unpack(rs2, a, b);
      f2_unpack(rq2, cq0, cq1);
      reinterpret_cast<float2*>(p.row_stats_out)[(size_t)(nb * nhalf + half) * p.M + row] = make_float2(a + b, cq0 + cq1);
    }
    mb = mb_n;
    nb = nb_n;
  }
}

}  // namespace seer
