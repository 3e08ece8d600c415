// tcgen05 / TMEM / TMA tile GEMM for sm_100a — the contraction engine behind every Linear, 1x1 conv and
// (as an implicit GEMM with shifted 4-D TMA boxes) every frame-wise 3x3 conv of the Seer UNet.
//
//   out[M, N] = epilogue( A[M, K] * Wt[N, K]^T )          A, Wt bf16 K-major, fp32 accumulate in TMEM
//
// Replaces (on B200) the cuBLAS / cuDNN calls behind nn.Linear and InflatedConv3d in the reference:
//   /root/reference/seer/models/resnet.py:8-16 (InflatedConv3d), attention.py:484-489 (to_q/k/v/out),
//   attention.py:781-793 (GEGLU proj), attention.py:742 (FF out).
//
// Tile: 128 (M) x BN (N) x 64 (K) per pipeline stage, SWIZZLE_128B operand tiles written by TMA and read by
// tcgen05.mma through shared-memory descriptors.  Warp roles: warp 0 = TMA producer (one elected lane),
// warp 1 = TMEM allocator + MMA issuer (one lane), warps 2..5 = epilogue (one TMEM lane quarter each).
// Two CTAs are co-resident per SM (<= 110 KB smem, <= 256 TMEM columns each) so one CTA's epilogue overlaps
// the other's main loop.
//
// A-operand sources (K-blocks are consumed in this order):
//   mode 0: 2-D row-major A[M, K1]                              (Linear / 1x1 conv)
//   mode 1: 4-D activation X[n_img, H, W, Cin] read as 9 shifted boxes (tap-major K = 9*Cin), zero padding
//           comes from TMA out-of-bounds fill                     (3x3 conv, stride 1, pad 1)
//   tail  : optional second 2-D source A2[M, K2] appended to K   (fused ResNet 1x1 shortcut / skip concat)
//
// Epilogue: + bias[(row / bias_div), col]  (+ fp32 residual[row, col]) -> fp32 or bf16;  or GEGLU:
// value/gate column blocks of 32 are interleaved in Wt so out[:, j] = (a + ba) * gelu_erf(g + bg) -> bf16.
#include "common.cuh"
#include "seer_b200.h"

namespace seer {

struct GemmParams {
  int M, N;
  int mode;        // 0 plain, 1 conv3x3
  int kb_main;     // k-blocks (of 64) from the main source
  int kb_total;    // + k-blocks from the tail source
  int cblk;        // conv: Cin / 64
  int H, W;        // conv image geometry
  const float* bias;
  int ldb;         // bias row stride (elements)
  int bias_div;    // rows per bias row (>= M for a plain bias vector)
  const float* residual;
  int ldr;
  void* out;
  int ldo;
  int out_bf16;
  int geglu;
};

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 192;

template <int BN>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN <= 128) ? 3 : (BN <= 160 ? 3 : 4);
  static constexpr int CTAS_PER_SM = (BN <= 160) ? 2 : 1;
  static constexpr int TMEM_COLS = (BN <= 128) ? 128 : 256;
  static constexpr int BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, GemmSmem<BN>::CTAS_PER_SM)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using S = GemmSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + S::STAGES;
  uint64_t* tmem_full_bar = empty_bar + S::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.kb_total > p.kb_main) tma_prefetch_desc(&tmA2);
    for (int s = 0; s < S::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, S::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int img0 = 0, y0 = 0;
      if (p.mode == 1) {
        const int hw = p.H * p.W;
        img0 = m0 / hw;
        y0 = (m0 % hw) / p.W;
      }
      for (int kb = 0; kb < p.kb_total; ++kb) {
        const int s = kb % S::STAGES;
        const uint32_t ph = (kb / S::STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sA = smem + s * S::STAGE_BYTES;
        uint8_t* sB = sA + S::A_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], S::STAGE_BYTES);
        if (kb < p.kb_main) {
          if (p.mode == 0) {
            tma_load_2d(sA, &tmA, &full_bar[s], kb * BK, m0);
          } else {
            const int tap = kb / p.cblk;
            const int c0 = (kb - tap * p.cblk) * BK;
            const int ky = tap / 3, kx = tap - ky * 3;
            tma_load_4d(sA, &tmA, &full_bar[s], c0, kx - 1, y0 + ky - 1, img0);
          }
        } else {
          tma_load_2d(sA, &tmA2, &full_bar[s], (kb - p.kb_main) * BK, m0);
        }
        tma_load_2d(sB, &tmB, &full_bar[s], kb * BK, n0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      for (int kb = 0; kb < p.kb_total; ++kb) {
        const int s = kb % S::STAGES;
        const uint32_t ph = (kb / S::STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
        const uint32_t b_addr = a_addr + S::A_BYTES;
        const uint64_t a_desc = umma_desc_sw128(a_addr);
        const uint64_t b_desc = umma_desc_sw128(b_addr);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 bf16 = 32 B inside the 128 B swizzle atom: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[s]);  // frees this smem stage once the MMAs above have read it
      }
      umma_commit(tmem_full_bar);    // accumulator complete
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < p.M;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const int brow = row_ok ? (row / p.bias_div) : 0;
    const float* bias_row = p.bias ? p.bias + (size_t)brow * p.ldb : nullptr;

    if (!p.geglu) {
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr_row + c, v);
        tmem_ld_wait();
        if (row_ok) {
          const int col0 = n0 + c;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (bias_row) {
            const float4* b4 = reinterpret_cast<const float4*>(bias_row + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 b = __ldg(b4 + j);
              f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
            }
          }
          if (p.residual) {
            const float4* r4 = reinterpret_cast<const float4*>(p.residual + (size_t)row * p.ldr + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 r = r4[j];
              f[4 * j] += r.x; f[4 * j + 1] += r.y; f[4 * j + 2] += r.z; f[4 * j + 3] += r.w;
            }
          }
          if (p.out_bf16) {
            uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack_bf16(f[8 * j], f[8 * j + 1]);
              o.y = pack_bf16(f[8 * j + 2], f[8 * j + 3]);
              o.z = pack_bf16(f[8 * j + 4], f[8 * j + 5]);
              o.w = pack_bf16(f[8 * j + 6], f[8 * j + 7]);
              o4[j] = o;
            }
          } else {
            float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) o4[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          }
        }
      }
    } else {
      // GEGLU: columns [c, c+32) are "value", [c+32, c+64) the matching "gate" (weights packed that way).
#pragma unroll 1
      for (int c = 0; c < BN; c += 64) {
        uint32_t va[32], vg[32];
        tmem_ld_32x32(taddr_row + c, va);
        tmem_ld_32x32(taddr_row + c + 32, vg);
        tmem_ld_wait();
        if (row_ok) {
          const int col0 = n0 + c;
          float o[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float a = __uint_as_float(va[j]) + __ldg(bias_row + col0 + j);
            float g = __uint_as_float(vg[j]) + __ldg(bias_row + col0 + 32 + j);
            o[j] = a * gelu_erf_f(g);
          }
          uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + (col0 >> 1));
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = pack_bf16(o[8 * j], o[8 * j + 1]);
            u.y = pack_bf16(o[8 * j + 2], o[8 * j + 3]);
            u.z = pack_bf16(o[8 * j + 4], o[8 * j + 5]);
            u.w = pack_bf16(o[8 * j + 6], o[8 * j + 7]);
            o4[j] = u;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, S::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
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

// 2-D bf16 row-major [rows, cols] (cols contiguous), box = {64 cols, box_rows}, 128 B swizzle.
static int make_map_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return SEER_ENODRIVER;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[seer_b200] cuTensorMapEncodeTiled(2d) failed: %d (rows=%llu cols=%llu ld=%llu box_rows=%u)\n", (int)r,
            (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows);
    return SEER_EINVAL;
  }
  return SEER_OK;
}

// 4-D bf16 activation [n_img, H, W, C] (C contiguous), box = {64, bw, bh, bn}.
static int make_map_4d(CUtensorMap* tm, const void* base, uint64_t n_img, uint64_t H, uint64_t W, uint64_t C, uint32_t bw,
                       uint32_t bh, uint32_t bn) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return SEER_ENODRIVER;
  cuuint64_t dims[4] = {C, W, H, n_img};
  cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
  cuuint32_t box[4] = {64, bw, bh, bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[seer_b200] cuTensorMapEncodeTiled(4d) failed: %d\n", (int)r);
    return SEER_EINVAL;
  }
  return SEER_OK;
}

template <int BN>
static int launch_gemm(const CUtensorMap& tA, const CUtensorMap& tA2, const CUtensorMap& tB, const GemmParams& p,
                       cudaStream_t stream) {
  using S = GemmSmem<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  dim3 grid(p.N / BN, ceil_div(p.M, BM));
  gemm_tc_kernel<BN><<<grid, GEMM_THREADS, S::BYTES, stream>>>(tA, tA2, tB, p);
  SEER_LAUNCH_CHECK();
  return SEER_OK;
}

static int pick_bn(int N, int geglu) {
  if (geglu) return (N % 128 == 0) ? 128 : 0;
  if (N % 160 == 0) return 160;
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  return 0;
}

static int dispatch(int bn, const CUtensorMap& tA, const CUtensorMap& tA2, const CUtensorMap& tB, const GemmParams& p,
                    cudaStream_t stream) {
  switch (bn) {
    case 64: return launch_gemm<64>(tA, tA2, tB, p, stream);
    case 128: return launch_gemm<128>(tA, tA2, tB, p, stream);
    case 160: return launch_gemm<160>(tA, tA2, tB, p, stream);
    default: return SEER_EUNSUPPORTED;
  }
}

static int fill_epilogue(GemmParams& p, int M, int N, const float* bias, int ldb, int bias_div, const float* residual, int ldr,
                         void* out, int ldo, int flags) {
  p.M = M; p.N = N;
  p.bias = bias; p.ldb = ldb > 0 ? ldb : N; p.bias_div = bias_div > 0 ? bias_div : 0x7fffffff;
  p.residual = residual; p.ldr = ldr;
  p.out = out; p.ldo = ldo;
  p.out_bf16 = (flags & SEER_GEMM_OUT_BF16) ? 1 : 0;
  p.geglu = (flags & SEER_GEMM_GEGLU) ? 1 : 0;
  if (p.geglu && (!p.out_bf16 || !bias || residual)) return SEER_EINVAL;
  return SEER_OK;
}

}  // namespace seer

using namespace seer;

extern "C" int seer_b200_gemm_bf16(const void* A, int lda, int K1, const void* A2, int lda2, int K2, const void* Wt, int M, int N,
                                   const float* bias, int ldb, int bias_div, const float* residual, int ldr, void* out,
                                   int ldo, int flags, void* stream) {
  SEER_CHECK_ARG(A && Wt && out && M > 0 && N > 0 && K1 > 0);
  SEER_CHECK_ARG(K1 % 64 == 0 && K2 % 64 == 0 && lda % 8 == 0 && (K2 == 0 || (A2 && lda2 % 8 == 0)));
  SEER_CHECK_ARG(ldo % 8 == 0 && (residual == nullptr || ldr % 4 == 0) && (ldb <= 0 || ldb % 4 == 0));
  const int geglu = (flags & SEER_GEMM_GEGLU) ? 1 : 0;
  const int bn = pick_bn(N, geglu);
  if (!bn) return SEER_EUNSUPPORTED;
  GemmParams p{};
  int rc = fill_epilogue(p, M, N, bias, ldb, bias_div, residual, ldr, out, ldo, flags);
  if (rc) return rc;
  p.mode = 0;
  p.kb_main = K1 / 64;
  p.kb_total = (K1 + K2) / 64;
  p.cblk = 1; p.H = 1; p.W = 1;
  CUtensorMap tA, tA2, tB;
  if ((rc = make_map_2d(&tA, A, M, K1, lda, BM))) return rc;
  if (K2) { if ((rc = make_map_2d(&tA2, A2, M, K2, lda2, BM))) return rc; } else tA2 = tA;
  if ((rc = make_map_2d(&tB, Wt, N, K1 + K2, K1 + K2, bn))) return rc;
  return dispatch(bn, tA, tA2, tB, p, (cudaStream_t)stream);
}

extern "C" int seer_b200_conv3x3_bf16(const void* X, int n_img, int H, int W, int Cin, const void* A2, int lda2, int K2,
                                      const void* Wt, int Cout, const float* bias, int ldb, int bias_div, const float* residual,
                                      int ldr, void* out, int ldo, int flags, void* stream) {
  SEER_CHECK_ARG(X && Wt && out && n_img > 0 && H > 0 && W > 0);
  SEER_CHECK_ARG(Cin % 64 == 0 && K2 % 64 == 0 && (K2 == 0 || (A2 && lda2 % 8 == 0)));
  SEER_CHECK_ARG(ldo % 8 == 0 && (residual == nullptr || ldr % 4 == 0) && (ldb <= 0 || ldb % 4 == 0));
  SEER_CHECK_ARG(!(flags & SEER_GEMM_GEGLU));
  // A 128-pixel M tile must be a whole number of image rows (or of images): W | 128 and the tile never
  // straddles an image boundary mid-row.
  if (W > 128 || 128 % W != 0) return SEER_EUNSUPPORTED;
  int rows = 128 / W, bh, bn_img;
  if (rows <= H) {
    if (H % rows != 0) return SEER_EUNSUPPORTED;
    bh = rows; bn_img = 1;
  } else {
    if (rows % H != 0) return SEER_EUNSUPPORTED;
    bh = H; bn_img = rows / H;
  }
  const int M = n_img * H * W;
  const int bn = pick_bn(Cout, 0);
  if (!bn) return SEER_EUNSUPPORTED;
  GemmParams p{};
  int rc = fill_epilogue(p, M, Cout, bias, ldb, bias_div, residual, ldr, out, ldo, flags);
  if (rc) return rc;
  p.mode = 1;
  p.cblk = Cin / 64;
  p.kb_main = 9 * p.cblk;
  p.kb_total = p.kb_main + K2 / 64;
  p.H = H; p.W = W;
  CUtensorMap tA, tA2, tB;
  if ((rc = make_map_4d(&tA, X, n_img, H, W, Cin, W, bh, bn_img))) return rc;
  if (K2) { if ((rc = make_map_2d(&tA2, A2, M, K2, lda2, BM))) return rc; } else tA2 = tA;
  const int Kt = 9 * Cin + K2;
  if ((rc = make_map_2d(&tB, Wt, Cout, Kt, Kt, bn))) return rc;
  return dispatch(bn, tA, tA2, tB, p, (cudaStream_t)stream);
}
