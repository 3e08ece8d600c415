// tcgen05 / TMEM / TMA persistent tile GEMM for sm_100a (v2) — the contraction engine behind every Linear, 1x1 conv
// and (as an implicit GEMM with shifted 4-D TMA boxes) every frame-wise 3x3 conv of the Seer UNet.
//
//   acc[M, N] = A[M, K] * Wt[N, K]^T          A, Wt bf16 K-major, fp32 accumulation in TMEM
//   out       = epilogue(acc)                  LN-fold / bias / residual / GEGLU / fp32 + bf16 outputs / norm statistics
//
// Replaces (on B200) the cuBLAS / cuDNN calls behind nn.Linear and InflatedConv3d in the reference:
//   /root/reference/seer/models/resnet.py:8-16 (InflatedConv3d), attention.py:484-489 (to_q/k/v/out),
//   attention.py:781-793 (GEGLU proj), attention.py:742 (FF out); and absorbs the statistics passes of
//   nn.GroupNorm (resnet.py:179,197) and nn.LayerNorm (attention.py:198-200) into the producing / consuming GEMM.
//
// Structure (one persistent CTA — or cta_group::2 CTA pair, 256-row tiles — per SM, static round-robin over output tiles,
// n fastest so the CTAs that share an A row-block run at the same time and A is read from HBM once):
//   warp 0      TMA producer: A (2-D box; for the convs shifted 4-D boxes, OOB zero fill = padding; a traversal stride of 2
//               for Downsample3D; 2x2 tap sets for the four phases of Upsample3D) and Wt tiles into a `stages`-deep
//               SWIZZLE_128B smem ring, mbarrier full/empty.  B-stationary mode (K <= 320) keeps the Wt panel resident.
//   warp 1      TMEM allocator + tcgen05.mma issuer (one lane): (128 | 256) x BN x 16 UMMAs into one of TWO TMEM accumulator
//               buffers, so the epilogue of tile i overlaps the main loop of tile i+1 (BN = 320: one 512-column buffer, two
//               N = 160 UMMAs per A stage — half the A bytes per FLOP for the long-K N = 320 / 640 convs).
//   warps 2..   epilogue warps (4 or 8).  Each owns one TMEM lane quarter (32 rows) and a private ring of smem slots: the
//               fp32 / bf16 residual chunk (32 rows x 32 cols) is PREFETCHED into the slot by TMA one to three chunks ahead,
//               the warp adds accumulator + bias (+ LayerNorm fold, GEGLU, RoPE) in registers (thread = row), stages the
//               chunk in the slot (XOR-swizzled 16-byte pieces, bank-conflict free) and copies it out with coalesced
//               128-bit LSU stores — TMA stores were measured slower here (fence.proxy.async + bulk-group round trip per
//               chunk, tools/tma_store_bench.cu), as were direct per-row stores (thread = row, four 16-byte
//               stores: qkv 220 -> 288 us, profiles/r2_gemm_probe.txt).
//
// Epilogue options: see SeerGemmDesc in include/seer_b200.h; the combinations the UNet issues are compiled as straight-line
// specialisations (gemm_epilogue.cuh EK_*), grouped into three kernel instantiations per tile shape (GRP below).
#include "gemm_epilogue.cuh"
#include "seer_b200.h"

#include <stdlib.h>

namespace seer {

// GRP: which epilogue specialisations this instantiation carries — 0: everything except the two register-heavy families,
// 1: the GEGLU kinds, 2: the RoPE kind.  One kernel holding all of them let the heaviest path dictate the register
// allocation (and the spills) of every other one (measured: adding the RoPE kind slowed the GEGLU launches by 12 %).
template <int BN, int CG, int GRP>
__global__ void __launch_bounds__(GEMM_MAX_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmRes, const GemmParams p) {
  using C = GemmCfg<BN, CG>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // B-stationary mode (small K): [Wt panel: kb_total x B_BYTES][A stages: stages x A_BYTES]; else [stages x (A | B)]
  const int stage_stride = p.bstat ? A_BYTES : C::STAGE_BYTES;
  uint8_t* stage_base = smem + (p.bstat ? p.kb_total * C::B_BYTES : 0);
  uint8_t* ring_base = stage_base + p.stages * stage_stride;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring_base + p.nepi * p.ring * p.slot_bytes);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + MAX_STAGES;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;         // [2]
  uint64_t* res_full_bar = tmem_empty_bar + 2;          // [MAX_EPI_WARPS][MAX_RING]
  uint64_t* bpanel_bar = res_full_bar + MAX_EPI_WARPS * MAX_RING;   // [1] B-stationary: the Wt panel has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bpanel_bar + 1);
  float* evec_base = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTA pair bookkeeping: `unit` = this CTA (CG 1) or this pair (CG 2) in the persistent tile loop
  const int rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int unit = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nunits = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.kb_total > p.kb_main) tma_prefetch_desc(&tmA2);
    if (p.res_mode) tma_prefetch_desc(&tmRes);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], p.nepi * CG);     // CG 2: the epilogue warps of BOTH CTAs arrive on the leader's
    }
    for (int i = 0; i < MAX_EPI_WARPS * MAX_RING; ++i) mbar_init(&res_full_bar[i], 1);
    mbar_init(bpanel_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) { tmem_alloc_cg2(tmem_slot, C::TMEM_COLS); tmem_relinquish_cg2(); }
    else { tmem_alloc(tmem_slot, C::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncwarp();                                            // reconverge (barrier.cluster is .aligned)
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // barrier inits visible (cluster-wide) before any arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above overlapped the previous kernel's tail; from here on global memory is touched
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer (whole warp loops — uniform control flow; one elected lane issues) ==========
    {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t full0 = CG == 2 ? mapa_shared(smem_u32(full_bar), 0) : 0;   // the pair leader's full barriers
      if (p.bstat) {
        // this CTA's n-block never changes (grid is a multiple of tiles_n): fetch its whole Wt panel once
        const int nb0 = unit % p.tiles_n;
        if (elect_one()) {
          if (CG == 1) {
            mbar_arrive_expect_tx(bpanel_bar, p.kb_total * C::B_BYTES);
            for (int kb = 0; kb < p.kb_total; ++kb) tma_load_2d(smem + kb * C::B_BYTES, &tmB, bpanel_bar, kb * BK, nb0 * BN);
          } else {
            if (rank == 0) mbar_arrive_expect_tx(bpanel_bar, 2 * p.kb_total * C::B_BYTES);
            const uint32_t bb = mapa_shared(smem_u32(bpanel_bar), 0);
            for (int kb = 0; kb < p.kb_total; ++kb)
              tma_load_2d_cg2(smem + kb * C::B_BYTES, &tmB, bb, kb * BK, nb0 * BN + rank * (BN / CG));
          }
        }
        __syncwarp();
      }
      const uint32_t stage_tx = p.bstat ? A_BYTES : C::STAGE_BYTES;
      for (int tile = unit; tile < p.num_tiles; tile += nunits) {
        const int mb = tile / p.tiles_n, nb = tile - mb * p.tiles_n;
        const int m0 = (mb * CG + rank) * BM, n0 = nb * BN + rank * (BN / CG);
        int img0 = 0, y0 = 0, x0 = 0;
        if (p.mode == 1) {
          const int hw = p.H * p.W;
          img0 = m0 / hw;
          const int rem = m0 - img0 * hw;
          y0 = rem / p.W;
          x0 = rem - y0 * p.W;               // 0 unless the image is wider than a 128-pixel tile (W > 128: row segments)
        }
        if (p.l2_prefetch) {
          // streaming (small-K) launch: pull the operand rows and the residual tile of a later tile of this CTA into L2 now
          const int t2 = tile + p.l2_prefetch * nunits;
          if (t2 < p.num_tiles && elect_one()) {
            const int mb2 = t2 / p.tiles_n, nb2 = t2 - mb2 * p.tiles_n;
            const int m2 = (mb2 * CG + rank) * BM;
            for (int kb = 0; kb < p.kb_main; ++kb) tma_prefetch_2d(&tmA, kb * BK, m2);
            for (int kb = p.kb_main; kb < p.kb_total; ++kb) tma_prefetch_2d(&tmA2, (kb - p.kb_main) * BK, m2);
            if (p.res_mode) {
              const int cpb = p.res_mode == 1 ? 32 : 32;      // residual map boxes are 32 x 32 elements
              for (int ry = 0; ry < BM / 32; ++ry)
                for (int cx = 0; cx < BN / 32; ++cx) tma_prefetch_2d(&tmRes, nb2 * BN + cx * cpb, m2 + ry * 32);
            }
          }
          __syncwarp();
        }
        // conv tap cursor, advanced incrementally (no division by the run-time tap count on the producer's critical path)
        int t_kx = 0, t_ky = 0, t_c0 = 0;
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sA = stage_base + s * stage_stride;
          uint8_t* sB = sA + A_BYTES;
          if (elect_one()) {
          if (CG == 1) {
            mbar_arrive_expect_tx(&full_bar[s], stage_tx);
            if (kb < p.kb_main) {
              if (p.mode == 0) {
                tma_load_2d(sA, &tmA, &full_bar[s], kb * BK, m0);
              } else {
                // K order [Cin/64][tap][64]: the 9 shifted views of one 64-channel slab are fetched back to back, so
                // the slab (and its halo) is served from L2 while it is hot instead of being re-read 9 x Cin/64 k-blocks apart
                tma_load_4d(sA, &tmA, &full_bar[s], t_c0, p.cstride * x0 + t_kx + p.off_x, p.cstride * y0 + t_ky + p.off_y, img0);
              }
            } else {
              tma_load_2d(sA, &tmA2, &full_bar[s], (kb - p.kb_main) * BK, m0);
            }
            if (!p.bstat) {
#pragma unroll
              for (int h = 0; h < C::NMMA; ++h)
                tma_load_2d(sB + h * (C::B_BYTES / C::NMMA), &tmB, &full_bar[s], kb * BK, nb * BN + h * (BN / C::NMMA));
            }
          } else {
            // both CTAs of the pair fill their own stage; all bytes are counted on the LEADER's full barrier
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * stage_tx);
            const uint32_t fb = full0 + (uint32_t)s * 8u;
            if (kb < p.kb_main) {
              if (p.mode == 0) {
                tma_load_2d_cg2(sA, &tmA, fb, kb * BK, m0);
              } else {
                // K order [Cin/64][tap][64]: the 9 shifted views of one 64-channel slab are fetched back to back, so
                // the slab (and its halo) is served from L2 while it is hot instead of being re-read 9 x Cin/64 k-blocks apart
                tma_load_4d_cg2(sA, &tmA, fb, t_c0, p.cstride * x0 + t_kx + p.off_x, p.cstride * y0 + t_ky + p.off_y, img0);
              }
            } else {
              tma_load_2d_cg2(sA, &tmA2, fb, (kb - p.kb_main) * BK, m0);
            }
            if (!p.bstat) {
              // NMMA = 2: UMMA h multiplies accumulator columns [h BN/2, (h+1) BN/2), whose Wt rows are split over the pair
#pragma unroll
              for (int h = 0; h < C::NMMA; ++h)
                tma_load_2d_cg2(sB + h * (C::B_BYTES / C::NMMA), &tmB, fb, kb * BK,
                                nb * BN + h * (BN / C::NMMA) + rank * (BN / C::NMMA / CG));
            }
          }
          }
          __syncwarp();
          if (++t_kx == p.taps_w) { t_kx = 0; if (++t_ky == p.taps_h) { t_ky = 0; t_c0 += BK; } }
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // ===== whole warp loops (uniform control flow, descriptors in uniform registers); one elected lane issues =====
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM * CG, BN / C::NMMA);
      constexpr uint64_t b_half = (uint64_t)((C::B_BYTES / C::NMMA) >> 4);     // descriptor step to the second UMMA's Wt rows
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      if (p.bstat) mbar_wait(bpanel_bar, 0);
      for (int tile = unit; tile < p.num_tiles; tile += nunits, ++it) {
        const int buf = C::NBUF == 2 ? (it & 1) : 0;
        const uint32_t buf_par = C::NBUF == 2 ? (((uint32_t)it >> 1) & 1) : ((uint32_t)it & 1);
        mbar_wait(&tmem_empty_bar[buf], buf_par ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * C::TBUF);
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(stage_base + s * stage_stride);
          const uint64_t a_desc = umma_desc_sw128(a_addr);
          const uint64_t b_desc = umma_desc_sw128(p.bstat ? smem_u32(smem + kb * C::B_BYTES) : a_addr + A_BYTES);
          if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 B inside the 128 B swizzle atom: +2 in the (addr >> 4) field
#pragma unroll
            for (int h = 0; h < C::NMMA; ++h) {
              const uint32_t dt = d_tmem + (uint32_t)(h * (BN / C::NMMA));
              const uint64_t bd = b_desc + (uint64_t)(k * 2) + (uint64_t)h * b_half;
              if (CG == 2) umma_bf16_cg2(dt, a_desc + (uint64_t)(k * 2), bd, idesc, (kb | k) != 0);
              else umma_bf16(dt, a_desc + (uint64_t)(k * 2), bd, idesc, (kb | k) != 0);
            }
          }
          // frees this smem stage (in both CTAs of a pair) once the MMAs above have read it
          if (CG == 2) umma_commit_cg2(&empty_bar[s], 3); else umma_commit(&empty_bar[s]);
          if (kb == p.kb_total - 1) {          // accumulator complete
            if (CG == 2) umma_commit_cg2(&tmem_full_bar[buf], 3); else umma_commit(&tmem_full_bar[buf]);
          }
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp < 2 + p.nepi) {
    // ===================== epilogue warps (gemm_epilogue.cuh) =====================
#define SEER_EPI(SPEC)                                                                                                  \
  gemm_epilogue_warp<BN, CG, SPEC>(p, &tmRes, ring_base, tmem_full_bar, tmem_empty_bar, res_full_bar, evec_base, tmem_base, \
                                   warp, lane, rank, unit, nunits)
    if constexpr (GRP == 1) {
      if constexpr (BN == 128 || BN == 256) {
        if (p.epi_spec == EK_FF1) SEER_EPI(EK_FF1); else SEER_EPI(EK_FF1_PLAIN);
      }
    } else if constexpr (GRP == 2) {
      if constexpr (BN != 320) SEER_EPI(EK_QKV_ROPE);
    } else {
      switch (p.epi_spec) {
        case EK_PIN: SEER_EPI(EK_PIN); break;
        case EK_QKV: SEER_EPI(EK_QKV); break;
        case EK_ATTN_OUT: SEER_EPI(EK_ATTN_OUT); break;
        case EK_FF2: SEER_EPI(EK_FF2); break;
        case EK_POUT: SEER_EPI(EK_POUT); break;
        case EK_CONV: SEER_EPI(EK_CONV); break;
        case EK_BF16: SEER_EPI(EK_BF16); break;
        case EK_PIN16: SEER_EPI(EK_PIN16); break;
        case EK_ATTN_OUT16: SEER_EPI(EK_ATTN_OUT16); break;
        case EK_FF2_16: SEER_EPI(EK_FF2_16); break;
        case EK_CONV16: SEER_EPI(EK_CONV16); break;
        case EK_POUT16: SEER_EPI(EK_POUT16); break;
        default: SEER_EPI(-1); break;
      }
    }
#undef SEER_EPI
  }

  tc_fence_before();
  __syncwarp();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // pair: the peer may still arrive on / read this CTA's smem
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_cg2(tmem_base, C::TMEM_COLS); else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps, launch plan
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
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

// 2-D row-major [rows, cols] (cols contiguous, leading dim ld elements), box = {box_cols, box_rows}.
static int make_map_2d(CUtensorMap* tm, CUtensorMapDataType dt, int esize, const void* base, uint64_t rows, uint64_t cols,
                       uint64_t ld_elems, uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle sw) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return SEER_ENODRIVER;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * (uint64_t)esize};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[seer_b200] cuTensorMapEncodeTiled(2d) failed: %d (rows=%llu cols=%llu ld=%llu box=%ux%u esize=%d)\n", (int)r,
            (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_cols, box_rows, esize);
    return SEER_EINVAL;
  }
  return SEER_OK;
}
static int make_map_bf16_k64(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  return make_map_2d(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rows, cols, ld, 64, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}

// 4-D bf16 activation [n_img, H, W, C] (C contiguous), box = {64, bw, bh, bn}.
// `stride` > 1: every stride-th pixel of a (bw*stride) x (bh*stride) bounding box (TMA traversal stride) — the stride-2
// Downsample3D conv reads its taps straight from the full-resolution image.
static int make_map_4d(CUtensorMap* tm, const void* base, uint64_t n_img, uint64_t H, uint64_t W, uint64_t C, uint32_t bw,
                       uint32_t bh, uint32_t bn, uint32_t stride = 1) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return SEER_ENODRIVER;
  cuuint64_t dims[4] = {C, W, H, n_img};
  cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
  cuuint32_t box[4] = {64, bw * stride, bh * stride, bn};
  cuuint32_t estr[4] = {1, stride, stride, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[seer_b200] cuTensorMapEncodeTiled(4d) failed: %d\n", (int)r);
    return SEER_EINVAL;
  }
  return SEER_OK;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

struct Plan {
  int bn, cg, stages, nepi, ring, slot_bytes, bstat, tiles_n, num_tiles, grid, smem_bytes;
};

static int env_int(const char* name, int dflt) { return env_cached(name, dflt); }

static inline int conv_ntaps(const SeerGemmDesc& d) {
  return (d.conv_taps_w > 0 && d.conv_taps_h > 0) ? d.conv_taps_w * d.conv_taps_h : 9;
}

static int make_plan(const SeerGemmDesc& d, Plan& pl) {
  const int N = d.N;
  // CTA pairs (cta_group::2, 256-row tiles) for tensor-bound launches: they halve the Wt bytes each SM pulls from L2
  // per FLOP (measured +15..25 % on K >= 1280 GEMMs); HBM-bound launches (small K, fp32 residual + output) run a
  // little better as independent 128-row CTAs.  Crude roofline estimate with the measured peaks (profiles/).
  const double Kd = d.X ? (double)conv_ntaps(d) * d.Cin + d.K2 : (double)d.K1 + d.K2;
  const double n_out = d.geglu ? N / 2 : N;
  const double bytes = (double)d.M * (d.X ? d.Cin + d.K2 : Kd) * 2 + (double)N * Kd * 2 +
                       (double)d.M * n_out * ((d.out_f32 ? 4 : 0) + (d.out_bf16 ? 2 : 0) + (d.residual ? (d.residual_bf16 ? 2 : 4) : 0));
  const double t_mem = bytes / 5.5e12, t_mma = 2.0 * d.M * N * Kd / 1.6e15;
  // (measured again after the epilogue rewrite, profiles/r1_gemm_sweep_cg.txt: pairs are now faster or equal on every shape of
  // the UNet, HBM-bound ones included — L1 proj_in 88 -> 70 us, L0 FF-out 270 -> 245 us, L2 attention-out 72 -> 66 us — so the
  // roofline estimate no longer selects single CTAs; t_mma / t_mem are kept for the planner's other decisions)
  int cg = d.M > BM ? 2 : 1;
  if (env_int("SEER_GEMM_CG", 0) == 3) cg = (d.M > BM && t_mma > t_mem) ? 2 : 1;     // A/B hook: the former roofline rule
  // B-stationary candidates (K <= 320, measured: proj_in 168 -> 118 us, qkv 412 -> 286 us at M = 262144; K = 640 and the
  // epilogue-bound GEGLU launches ran slower with it): a pair halves the resident panel
  const bool bstat_ok = d.M > BM && Kd <= 320.0 && !d.geglu && env_int("SEER_GEMM_BSTAT", 1);
  if (bstat_ok) cg = 2;
  const int fcg = env_int("SEER_GEMM_CG", 0);      // tuning hook
  if (fcg == 1 || fcg == 2) cg = fcg;
  pl.cg = cg;
  const int tiles_m = ceil_div(d.M, BM * cg);
  const int nsm = num_sms() / cg;                  // scheduling units: CTAs or CTA pairs
  // Tile width by a small time model (clocks per scheduling unit), calibrated on tools/gemm_bench.py --sweep at the UNet's shapes
  // (profiles/r2_gemm_sweep_320.txt, r2_gemm_sweep_L2L3.txt).  Per tile: main loop = k-blocks x 2 BN clocks / e(BN), where
  // e(BN) is the operand-delivery efficiency of a BN-wide pair tile (narrow tiles re-read A from L2 once per n-block: 128-wide
  // tiles measured ~1000, 160 ~1250, 256 ~1500, 320 ~1600+ TF/s on the same long-K conv); epilogue = chunks per warp x a
  // per-chunk latency; the two overlap (two TMEM accumulators) except for BN = 320 (single buffer).  The launch as a whole
  // cannot beat its HBM stream.  The former rule (rounds x (BN + 24)) treated a column as equally expensive at every width and
  // chose 160-wide tiles for the N = 1280 convs / FF-out of the 8x8 level (measured 392 vs 323 us) and for the 4x4 level.
  static const int cand_plain[] = {160, 256, 192, 320, 128, 64};      // ties (HBM-bound launches) go to the earlier entry
  static const int cand_geglu[] = {256, 128};
  const int* cand = d.geglu ? cand_geglu : cand_plain;
  const int ncand = d.geglu ? 2 : 6;
  // 320-wide pair tiles: one single-buffered 512-column accumulator, two N = 160 UMMAs per A stage (long main loops only)
  const bool allow320 = cg == 2 && Kd >= 1280.0 && !d.rope_tab && env_int("SEER_GEMM_BN320", 1);
  const double kb = Kd / BK;
  const bool of32 = d.out_f32 != nullptr;
  const double epi_clk = d.geglu ? 1300.0 : ((of32 && d.residual && !d.residual_bf16) ? 1000.0 : (of32 ? 800.0 : 600.0));
  const int nhalf_guess = (d.geglu || !of32 || Kd <= 1280.0) ? 2 : 1;
  const double mem_clk = bytes / 5.5e12 * 1.9e9;
  int best = 0;
  double best_cost = 1e30;
  const int forced = env_int("SEER_GEMM_BN", 0);   // tuning hook
  const bool old_model = env_int("SEER_GEMM_OLDPLAN", 0) != 0;      // A/B hook: the round-1 cost rule
  for (int i = 0; i < ncand; ++i) {
    const int bn = cand[i];
    if (N % bn) continue;
    if (bn == 320 && !(allow320 || (forced == 320 && cg == 2))) continue;
    if (forced && bn != forced) continue;
    const long tiles = (long)tiles_m * (N / bn);
    const long rounds = (tiles + nsm - 1) / nsm;
    double cost;
    if (old_model) {
      if (bn == 320 && !(Kd >= 2880.0 && N % 256 != 0)) continue;
      cost = (double)rounds * (bn + 24);
    } else {
      const double eff = bn >= 320 ? 1.0 : (bn >= 256 ? 0.95 : (bn >= 192 ? 0.85 : (bn >= 160 ? 0.76 : (bn >= 128 ? 0.62 : 0.40))));
      const double main_clk = kb * 2.0 * bn / eff;
      const int chunks = bn / (d.geglu ? 64 : 32);
      const double epi = (double)((chunks + nhalf_guess - 1) / nhalf_guess) * epi_clk;
      const double tile = bn == 320 ? (main_clk + epi) * 1.03 : (main_clk > epi ? main_clk : epi);
      cost = (double)rounds * tile;
      if (cost < mem_clk) cost = mem_clk;
    }
    if (cost < best_cost * 0.999) { best_cost = cost; best = bn; }
  }
  if (!best) {
    for (int i = 0; i < ncand && !best; ++i)
      if (N % cand[i] == 0 && cand[i] != 320) best = cand[i];
    if (!best) return SEER_EUNSUPPORTED;
  }
  // tuning hook: 320-wide (single-accumulator, non-resident-panel) tiles on the K <= 640 linear launches whose N allows it
  if (env_int("SEER_GEMM_320_SMALLK", 0) && cg == 2 && !d.X && Kd <= 640.0 && N % 320 == 0 && !d.rope_tab && !d.geglu && !forced) best = 320;
  if (d.row_stats_out && !forced) {
    // LayerNorm row-statistic producers: the number and the column extent of the per-row partial sums must not depend on M,
    // or a clip evaluated alone and inside a batch would sum its (sum, sumsq) in different orders (one fp32 ulp in mean / rstd,
    // a handful of bf16 outputs rounding the other way: profiles/r1_batch_dependence_probe.txt).  The tile width is therefore
    // a function of N alone here, and the launch always runs 8 epilogue warps (two partials per tile).
    // (256 for the 1280-channel levels, 160 for 320 / 640: what the time model picks at the benchmark's M; measured
    // profiles/r2_gemm_sweep_L2L3.txt: proj_in at M = 16384 52 vs 58 us)
    best = (N % 256 == 0 && N >= 1024) ? 256 : (N % 160 == 0 ? 160 : (N % 128 == 0 ? 128 : 64));
    if (env_int("SEER_GEMM_320_SMALLK", 0) && cg == 2 && !d.X && Kd <= 640.0 && N % 320 == 0) best = 320;
  }
  pl.bn = best;
  pl.tiles_n = N / best;
  pl.num_tiles = tiles_m * pl.tiles_n;
  pl.grid = (pl.num_tiles < nsm ? pl.num_tiles : nsm) * cg;
  // slot: residual chunk and staged output chunk share the bytes (fp32: 32 x 128 B, bf16: 32 x 64 B)
  const bool of = d.out_f32 != nullptr;
  const int rm = d.residual ? (d.residual_bf16 ? 2 : 1) : 0;
  pl.slot_bytes = (of || rm == 1) ? 4096 : 2048;      // bf16-only kinds stage bf16 (their column sums are read from that tile)
  // 8 epilogue warps (two per scheduler, so one warp's dependent-issue latency hides behind the other's) unless a long
  // main loop (big K) hides the epilogue anyway and the smem is better spent on operand stages
  // (4 warps + the extra operand stages for the K >= 1280 FF-out launches: level 0 229 vs 237 us, but the evaluation as a whole
  //  measured 0.5 ms slower in two A/B repetitions — not adopted; 4 vs 8 everywhere: profiles/r2_breakdown_nepi4.txt / nepi8.txt)
  pl.nepi = env_int("SEER_GEMM_NEPI", (d.geglu || !of || Kd <= 1280.0) ? 8 : 4);
  // 320-wide tiles are single-buffered: the epilogue is exposed, so it gets all 8 warps (and a short residual ring)
  if (best == 320) pl.nepi = env_int("SEER_GEMM_NEPI320", 8);
  const bool fixed_nepi = best == 320 || (d.row_stats_out && best >= 64 && !forced);
  if (d.row_stats_out && !forced) pl.nepi = best >= 64 ? 8 : 4;
  if (pl.nepi != 4 && pl.nepi != 8) pl.nepi = 4;
  if (best / (d.geglu ? 64 : 32) < 2) pl.nepi = 4;   // every epilogue warp needs at least one chunk
  // B-stationary: the CTA keeps the Wt panel of ONE n-block resident (needs grid % tiles_n == 0) and streams only A —
  // for K <= 640 the per-tile re-read of the panel from L2 (all SMs hammering the same ~100 KB) dominated the operand
  // traffic and capped these launches at ~25 B/cycle/SM (profiles/)
  const int Ktot_i = (int)Kd;
  const int panel_bytes = (Ktot_i / BK) * (best / cg) * BK * 2;
  const int units_bs = (nsm / pl.tiles_n) * pl.tiles_n;
  // (320-wide tiles issue two UMMAs per stage from a split Wt stage: the resident-panel layout does not provide that — a forced
  //  SEER_GEMM_BN=320 on a K <= 320 launch used to deadlock here, caught by the bounded mbarrier wait)
  pl.bstat = bstat_ok && best != 320 && panel_bytes <= 120 * 1024 && units_bs > 0 && pl.num_tiles >= 2 * units_bs;
  if (pl.bstat) pl.grid = units_bs * cg;
  const int stage_bytes = pl.bstat ? A_BYTES : A_BYTES + (best / cg) * BK * 2;
  const int rope_bytes = d.rope_tab ? MAX_EPI_WARPS * ROPE_BYTES_PER_WARP : 0;
  const int evec_bytes = MAX_EPI_WARPS * (best == 320 ? EVEC_BYTES_PER_WARP_320 : EVEC_BYTES_PER_WARP);
  const int avail = SMEM_LIMIT - 1024 /*alignment slack*/ - BAR_BYTES - evec_bytes - rope_bytes - (pl.bstat ? panel_bytes : 0);
  // (epilogue warps, ring depth, operand stages) by score: operand stages matter most (up to 5), then 8 epilogue warps
  // for the latency-bound bf16-only / GEGLU epilogues, then ring depth
  const int want_nepi = pl.nepi;
  const int want_ring = rm ? (best == 320 ? 2 : env_int("SEER_EPI_RING", 4)) : 1;    // no residual: the slot is only a staging buffer
  int best_score = -1;
  for (int ne = want_nepi; ne >= (fixed_nepi ? want_nepi : 4); ne -= 4) {
    for (int rg = want_ring; rg >= (rm ? 2 : 1); --rg) {
      int st = (avail - ne * rg * pl.slot_bytes) / stage_bytes;
      if (st < 2) continue;
      // (7 / 8 stages on the long-K launches measured no gain: 82.3 ms per evaluation either way; the streaming small-K
      //  launches measured the same with 6 and 8: profiles/r2_gemm_probe.txt)
      const int st_cap = (!d.X && Kd <= 640.0) ? env_int("SEER_GEMM_STAGES_SMALLK", 6) : 6;
      if (st > st_cap) st = st_cap;
      const int score = (st > 5 ? 5 : st) * 100 + (ne == 8 ? 30 : 0) + rg * 5 + (st > 5 ? st - 5 : 0);
      if (score > best_score) { best_score = score; pl.nepi = ne; pl.ring = rg; pl.stages = st; }
    }
  }
  if (best_score < 0) return SEER_EUNSUPPORTED;
  if (pl.stages < 2) return SEER_EUNSUPPORTED;
  if (pl.stages > MAX_STAGES) pl.stages = MAX_STAGES;
  const int fs = env_int("SEER_GEMM_STAGES", 0);
  if (fs >= 2 && fs <= pl.stages) pl.stages = fs;
  pl.smem_bytes = (pl.bstat ? panel_bytes : 0) + pl.stages * stage_bytes + pl.nepi * pl.ring * pl.slot_bytes + BAR_BYTES +
                  evec_bytes + rope_bytes + 1024;
  // > half of the SM's shared memory, so two CTAs (2 x TMEM_COLS could exceed 512 columns) never share an SM
  if (pl.smem_bytes < 120 * 1024) pl.smem_bytes = 120 * 1024;
  return SEER_OK;
}

template <int BN, int CG, int GRP>
static int launch_gemm_grp(const CUtensorMap* maps, const GemmParams& p, const Plan& pl, cudaStream_t stream) {
  static SmemAttrOnce smem_attr;
  { cudaError_t e = smem_attr.ensure(gemm_tc_kernel<BN, CG, GRP>, SMEM_LIMIT); if (e != cudaSuccess) return (int)e; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(pl.grid);
  cfg.blockDim = dim3(64 + 32 * pl.nepi);
  cfg.dynamicSmemBytes = pl.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, CG, GRP>, maps[0], maps[1], maps[2], maps[3], p);
  if (e != cudaSuccess) return (int)e;
  debug_note_gemm(BN, CG, pl.stages, pl.nepi, pl.ring, pl.bstat, p.epi_spec, p.mode);
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

template <int BN, int CG>
static int launch_gemm(const CUtensorMap* maps, const GemmParams& p, const Plan& pl, cudaStream_t stream) {
  if (p.epi_spec == EK_FF1 || p.epi_spec == EK_FF1_PLAIN) {
    if constexpr (BN == 128 || BN == 256) return launch_gemm_grp<BN, CG, 1>(maps, p, pl, stream);
    else return SEER_EUNSUPPORTED;
  }
  if (p.epi_spec == EK_QKV_ROPE) {
    if constexpr (BN != 320) return launch_gemm_grp<BN, CG, 2>(maps, p, pl, stream);
    else return SEER_EUNSUPPORTED;
  }
  return launch_gemm_grp<BN, CG, 0>(maps, p, pl, stream);
}

template <int BN>
static int launch_gemm_cg(const CUtensorMap* maps, const GemmParams& p, const Plan& pl, cudaStream_t stream) {
  return pl.cg == 2 ? launch_gemm<BN, 2>(maps, p, pl, stream) : launch_gemm<BN, 1>(maps, p, pl, stream);
}

static int check_desc(const SeerGemmDesc& d) {
  SEER_CHECK_ARG(d.Wt && d.M > 0 && d.N > 0 && (d.A || d.X));
  SEER_CHECK_ARG(d.out_f32 || d.out_bf16);
  SEER_CHECK_ARG(d.K2 % 64 == 0 && (d.K2 == 0 || (d.A2 && d.lda2 % 8 == 0)));
  if (d.X) {
    SEER_CHECK_ARG(d.n_img > 0 && d.H > 0 && d.W > 0 && d.Cin > 0 && d.Cin % 64 == 0);
    const int cs = d.conv_stride > 1 ? d.conv_stride : 1;
    SEER_CHECK_ARG(cs <= 2 && d.H % cs == 0 && d.W % cs == 0);
    SEER_CHECK_ARG(d.M == d.n_img * (d.H / cs) * (d.W / cs));
    SEER_CHECK_ARG((d.conv_taps_w > 0) == (d.conv_taps_h > 0) && d.conv_taps_w <= 3 && d.conv_taps_h <= 3);
    SEER_CHECK_ARG(d.out_up_phase >= 0 && d.out_up_phase <= 4);
    if (d.out_up_phase) SEER_CHECK_ARG(cs == 1 && (d.W & (d.W - 1)) == 0 && !d.residual && !d.row_stats_out && d.M % 32 == 0);
  } else {
    SEER_CHECK_ARG(d.out_up_phase == 0);
    SEER_CHECK_ARG(d.K1 > 0 && d.K1 % 64 == 0 && d.lda % 8 == 0);
  }
  SEER_CHECK_ARG(!d.out_f32 || d.ldo_f32 % 4 == 0);
  SEER_CHECK_ARG(!d.out_bf16 || d.ldo_bf16 % 8 == 0);
  SEER_CHECK_ARG(!d.residual || (d.residual_bf16 ? d.ldr % 8 == 0 : d.ldr % 4 == 0));
  SEER_CHECK_ARG(d.ldb <= 0 || d.ldb % 4 == 0);
  if (d.geglu) SEER_CHECK_ARG(d.bias_div <= 0 && d.N % 128 == 0 && d.out_bf16 && !d.out_f32 && d.bias && !d.residual && !d.col_stats && !d.row_stats_out);
  SEER_CHECK_ARG(!d.col_stats || d.out_f32 || d.out_bf16);
  if (d.row_stats_in) SEER_CHECK_ARG(d.row_parts_in > 0 && d.ln_colsum && !d.X);
  if (d.rope_tab)
    SEER_CHECK_ARG(!d.X && !d.geglu && d.out_bf16 && !d.out_f32 && !d.residual && !d.col_stats && !d.row_stats_out && d.row_stats_in && d.bias &&
                   d.rope_T > 0 && d.rope_T % 32 == 0 && d.M % d.rope_T == 0 && d.rope_d >= 32 && d.rope_d % 8 == 0 &&
                   d.rope_cols > 0 && d.rope_cols % 32 == 0 && d.rope_cols <= d.N && d.rope_cols % d.rope_d == 0);
  return SEER_OK;
}

}  // namespace seer

using namespace seer;

extern "C" int seer_b200_gemm_row_parts(const SeerGemmDesc* desc) {
  if (!desc) return SEER_EINVAL;
  SeerGemmDesc d = *desc;
  if (!d.row_stats_out) d.row_stats_out = reinterpret_cast<float*>(sizeof(float));   // plan as the launch that WILL write them
  Plan pl{};
  int rc = make_plan(d, pl);
  if (rc) return rc;
  return pl.tiles_n * (pl.nepi >> 2);
}

extern "C" int seer_b200_gemm_ex(const SeerGemmDesc* desc, void* stream) {
  if (!desc) return SEER_EINVAL;
  const SeerGemmDesc& d = *desc;
  int rc = check_desc(d);
  if (rc) return rc;
  Plan pl{};
  if ((rc = make_plan(d, pl))) return rc;

  GemmParams p{};
  p.M = d.M; p.N = d.N;
  p.tiles_n = pl.tiles_n; p.num_tiles = pl.num_tiles;
  {
    // epilogue specialisation: the option combinations of the UNet's hot launches are compiled as straight-line code
    const int flags = (d.row_stats_in ? EF_LN : 0) | (d.geglu ? EF_GEGLU : 0) | ((d.residual && !d.residual_bf16) ? EF_RES32 : 0) |
                      ((d.residual && d.residual_bf16) ? EF_RES16 : 0) |
                      (d.out_f32 ? EF_OUT32 : 0) | (d.out_bf16 ? EF_OUT16 : 0) | (d.col_stats ? EF_CSTAT : 0) |
                      (d.row_stats_out ? EF_RSTAT : 0) | (d.rope_tab ? EF_ROPE : 0);
    static const int kinds[] = {EK_PIN, EK_QKV, EK_ATTN_OUT, EK_FF1, EK_FF1_PLAIN, EK_FF2, EK_POUT, EK_CONV, EK_BF16,
                                EK_PIN16, EK_ATTN_OUT16, EK_FF2_16, EK_CONV16, EK_POUT16, EK_QKV_ROPE};
    p.epi_spec = -1;
    const bool generic_forced = env_int("SEER_GEMM_GENERIC", 0) != 0 && !d.geglu;    // A/B hook
    if (d.bias && !generic_forced)
      for (int k : kinds)
        if (k == flags) p.epi_spec = k;
    if ((d.geglu || d.rope_tab) && p.epi_spec < 0) return SEER_EUNSUPPORTED;
  }
  p.bstat = pl.bstat;
  p.evec_floats = pl.bn == 320 ? EVEC_FLOATS_320 : EVEC_FLOATS;
  // L2 prefetch distance (tiles of this CTA) for the streaming launches (plain GEMM with K <= 640).  OFF by default: measured
  // slower at every distance (profiles/r2_gemm_probe.txt: proj_out 161 -> 239 us, to_out 125 -> 136 us at distance 2) — these
  // launches are not short of bytes in flight
  p.l2_prefetch = (!d.X && d.K1 + d.K2 <= 640) ? env_int("SEER_GEMM_L2PF", 0) : 0;
  p.stages = pl.stages; p.nepi = pl.nepi; p.ring = pl.ring; p.slot_bytes = pl.slot_bytes;
  p.bias = d.bias; p.ldb = d.ldb > 0 ? d.ldb : d.N; p.bias_div = d.bias_div > 0 ? d.bias_div : BIAS_ONE_ROW;
  p.res_mode = d.residual ? (d.residual_bf16 ? 2 : 1) : 0;
  p.out_f32 = (float*)d.out_f32; p.ldo_f32 = d.ldo_f32;
  p.out_bf16 = (__nv_bfloat16*)d.out_bf16; p.ldo_bf16 = d.ldo_bf16;
  p.geglu = d.geglu ? 1 : 0;
  p.col_stats = d.col_stats; p.row_stats_out = d.row_stats_out;
  p.row_stats_in = d.row_stats_in; p.row_parts_in = d.row_parts_in; p.ln_eps = d.ln_eps; p.ln_colsum = d.ln_colsum;
  p.rope_tab = reinterpret_cast<const __half2*>(d.rope_tab); p.rope_T = d.rope_T; p.rope_cols = d.rope_cols; p.rope_d = d.rope_d;

  CUtensorMap maps[4];
  int Ktot;
  if (d.X) {
    // A 128-pixel M tile must be a whole number of image rows (or of images): W | 128 and the tile never
    // straddles an image boundary mid-row.
    const int cs = d.conv_stride > 1 ? d.conv_stride : 1;
    const int W = d.W / cs, H = d.H / cs;          // OUTPUT geometry: an M tile is whole output rows
    int bw, bh, bn_img;
    if (W > 128) {
      // images wider than one 128-pixel tile (the VAE decoder's 256x256 level): a tile is a 128-pixel row segment
      if (W % 128 != 0) return SEER_EUNSUPPORTED;
      bw = 128; bh = 1; bn_img = 1;
    } else {
      if (128 % W != 0) return SEER_EUNSUPPORTED;
      const int rows = 128 / W;
      bw = W;
      if (rows <= H) {
        if (H % rows != 0) return SEER_EUNSUPPORTED;
        bh = rows; bn_img = 1;
      } else {
        if (rows % H != 0) return SEER_EUNSUPPORTED;
        bh = H; bn_img = rows / H;
      }
    }
    p.mode = 1;
    p.cblk = d.Cin / 64;
    p.ntaps = conv_ntaps(d);
    p.taps_w = d.conv_taps_w > 0 ? d.conv_taps_w : 3;
    p.taps_h = d.conv_taps_w > 0 ? d.conv_taps_h : 3;
    p.off_x = d.conv_taps_w > 0 ? d.conv_off_x : -1;
    p.off_y = d.conv_taps_w > 0 ? d.conv_off_y : -1;
    p.cstride = cs;
    p.kb_main = p.ntaps * p.cblk;
    p.H = H; p.W = W;
    Ktot = p.ntaps * d.Cin + d.K2;
    if (d.out_up_phase) {
      p.up_phase = d.out_up_phase;
      p.up_wshift = 0;
      while ((1 << p.up_wshift) < W) ++p.up_wshift;
    }
    if ((rc = make_map_4d(&maps[0], d.X, d.n_img, d.H, d.W, d.Cin, bw, bh, bn_img, cs))) return rc;
  } else {
    p.mode = 0;
    p.cblk = 1; p.H = 1; p.W = 1;
    p.ntaps = 9; p.taps_w = 3; p.taps_h = 3; p.off_x = p.off_y = -1; p.cstride = 1;
    p.kb_main = d.K1 / 64;
    Ktot = d.K1 + d.K2;
    if ((rc = make_map_bf16_k64(&maps[0], d.A, d.M, d.K1, d.lda, BM))) return rc;
  }
  p.kb_total = Ktot / 64;
  p.ln_inv_dim = 1.0f / (float)(d.X ? 1 : d.K1);
  if (d.K2) { if ((rc = make_map_bf16_k64(&maps[1], d.A2, d.M, d.K2, d.lda2, BM))) return rc; } else maps[1] = maps[0];
  if ((rc = make_map_bf16_k64(&maps[2], d.Wt, d.N, Ktot, Ktot, pl.bn / pl.cg / (pl.bn > 256 ? 2 : 1)))) return rc;
  const int n_out = d.geglu ? d.N / 2 : d.N;
  maps[3] = maps[0];
  if (d.residual) {
    if (d.residual_bf16)
      rc = make_map_2d(&maps[3], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d.residual, d.M, n_out, d.ldr, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    else
      rc = make_map_2d(&maps[3], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d.residual, d.M, n_out, d.ldr, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  switch (pl.bn) {
    case 64: return launch_gemm_cg<64>(maps, p, pl, st);
    case 128: return launch_gemm_cg<128>(maps, p, pl, st);
    case 160: return launch_gemm_cg<160>(maps, p, pl, st);
    case 192: return launch_gemm_cg<192>(maps, p, pl, st);
    case 256: return launch_gemm_cg<256>(maps, p, pl, st);
    case 320: return pl.cg == 2 ? launch_gemm<320, 2>(maps, p, pl, st) : SEER_EUNSUPPORTED;
    default: return SEER_EUNSUPPORTED;
  }
}

// ---- the two original entry points, now thin wrappers over the descriptor call ----
extern "C" int seer_b200_gemm_bf16(const void* A, int lda, int K1, const void* A2, int lda2, int K2, const void* Wt, int M, int N,
                                   const float* bias, int ldb, int bias_div, const float* residual, int ldr, void* out,
                                   int ldo, int flags, void* stream) {
  SeerGemmDesc d{};
  d.A = A; d.lda = lda; d.K1 = K1;
  d.A2 = A2; d.lda2 = lda2; d.K2 = K2;
  d.Wt = Wt; d.M = M; d.N = N;
  d.bias = bias; d.ldb = ldb; d.bias_div = bias_div;
  d.residual = residual; d.ldr = ldr;
  if (flags & SEER_GEMM_OUT_BF16) { d.out_bf16 = out; d.ldo_bf16 = ldo; } else { d.out_f32 = out; d.ldo_f32 = ldo; }
  d.geglu = (flags & SEER_GEMM_GEGLU) ? 1 : 0;
  return seer_b200_gemm_ex(&d, stream);
}

extern "C" int seer_b200_conv3x3_bf16(const void* X, int n_img, int H, int W, int Cin, const void* A2, int lda2, int K2,
                                      const void* Wt, int Cout, const float* bias, int ldb, int bias_div, const float* residual,
                                      int ldr, void* out, int ldo, int flags, void* stream) {
  if (flags & SEER_GEMM_GEGLU) return SEER_EINVAL;
  SeerGemmDesc d{};
  d.X = X; d.n_img = n_img; d.H = H; d.W = W; d.Cin = Cin;
  d.A2 = A2; d.lda2 = lda2; d.K2 = K2;
  d.Wt = Wt; d.M = n_img * H * W; d.N = Cout;
  d.bias = bias; d.ldb = ldb; d.bias_div = bias_div;
  d.residual = residual; d.ldr = ldr;
  if (flags & SEER_GEMM_OUT_BF16) { d.out_bf16 = out; d.ldo_bf16 = ldo; } else { d.out_f32 = out; d.ldo_f32 = ldo; }
  return seer_b200_gemm_ex(&d, stream);
}
