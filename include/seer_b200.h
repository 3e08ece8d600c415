/* seer_b200.h — C ABI of libseer_b200.so: the sm_100a kernels behind Seer's DDIM+CFG denoising step.
 *
 * The reference (seervideodiffusion/SeerVideoLDM) has no native layer and no FFI: its hot path is PyTorch
 * modules calling cuDNN / cuBLAS / xformers.  Each entry point below therefore cites the reference *operator*
 * it replaces (paths relative to /root/reference).  Conventions:
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers unless noted;
 *   - every function returns int: 0 = ok, >0 = cudaError_t of the failed launch, <0 = argument/support error
 *     (SEER_B200_EINVAL -1, SEER_B200_EUNSUPPORTED -2, SEER_B200_ENODRIVER -3); nothing throws across the ABI;
 *   - no hidden allocation and no host synchronisation: outputs and workspaces are caller-owned, launches go to
 *     the caller's `stream` (a cudaStream_t passed as void*), so every call is CUDA-graph capturable;
 *   - activations are channels-last "token-major": row = ((b*F + f)*H + y)*W + x, columns = channels.
 */
#ifndef SEER_B200_H
#define SEER_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SEER_B200_EINVAL (-1)
#define SEER_B200_EUNSUPPORTED (-2)
#define SEER_B200_ENODRIVER (-3)

/* flags for the GEMM / conv epilogue */
#define SEER_GEMM_OUT_BF16 1 /* store bf16 (default fp32) */
#define SEER_GEMM_GEGLU 2    /* out[:, j] = (a + ba) * gelu_erf(g + bg); Wt/bias rows packed in value/gate blocks of 32 */

/* attention modes */
#define SEER_ATTN_SPATIAL 0
#define SEER_ATTN_CROSS 1
#define SEER_ATTN_SCTA 2
#define SEER_ATTN_FRAME 3 /* causal attention along the frame axis, one sequence per (clip, token): FSText temporal blocks */

const char* seer_b200_version(void);
/* Debug hooks: a description of the kernel the dispatcher chose for the calling thread's last seer_b200_attention /
 * seer_b200_gemm_ex launch ("attention_tc_persist_kernel<40> tcgen05", "attention_kernel<D> mma.sync (legacy path)",
 * "gemm_tc_kernel<160,2> tcgen05 stages=5 ..."); tests assert the tcgen05 paths run for every benchmark shape. */
const char* seer_b200_debug_last_attention(void);
const char* seer_b200_debug_last_gemm(void);
/* The SEER_* tuning switches are read from the environment once per process; tuning tools override (clear = 0) or reset to the
 * built-in default (clear = 1) a switch inside the running process with this. */
void seer_b200_debug_setenv(const char* name, int value, int clear);

/* out[M,N] = A[M,K1] (|| A2[M,K2]) * Wt[N,K1+K2]^T + bias[(row/bias_div), :] (+ residual), tcgen05/TMEM/TMA.
 * Replaces nn.Linear / 1x1 InflatedConv3d: seer/models/attention.py:484-489 (to_q/k/v/out), :111,126 (proj_in/out),
 * :783,742 (GEGLU proj, FF out), resnet.py:172 (conv_shortcut).  A, A2, Wt bf16; bias/residual fp32; K1,K2 % 64 == 0;
 * N % 64 == 0 (GEGLU: N % 128 == 0, out has N/2 columns).  bias is [rows, ldb] (ldb <= 0: ldb = N) and row
 * (row / bias_div) is used; bias_div <= 0: one bias row (per-sample bias: bias_div = tokens per sample). */
int seer_b200_gemm_bf16(const void* A, int lda, int K1, const void* A2, int lda2, int K2, const void* Wt, int M, int N,
                        const float* bias, int ldb, int bias_div, const float* residual, int ldr, void* out, int ldo,
                        int flags, void* stream);

/* Full descriptor of one tcgen05 GEMM / implicit-GEMM conv launch (superset of the two entry points around it):
 *
 *   acc[M,N]  = [A | A2] * Wt^T                       A: plain A[M,K1] bf16, or (X != NULL) the 3x3/pad-1 im2col of
 *                                                      X[n_img,H,W,Cin] bf16 (K1 = 9*Cin, K order [Cin/64][ky][kx][64])
 *   v         = ln ? rstd_r * (acc - mean_r * ln_colsum[n]) : acc        (LayerNorm folded into the GEMM: Wt holds
 *                                                      W*gamma, bias holds beta*W^T (+b); mean/rstd of row r come from
 *                                                      row_stats_in = `row_parts_in` partial (sum, sumsq) pairs)
 *   v        += bias[(r / bias_div), n] + residual[r, n]                 (residual fp32 or bf16)
 *   geglu: v[:, j] = value * gelu_erf(gate)            (Wt / bias / ln_colsum rows packed value/gate in blocks of 32)
 *   outputs   : out_f32 and/or out_bf16 (either may be NULL, not both), staged in shared memory and written with
 *               coalesced 128-bit stores;
 *   col_stats : optional [ceil(M/32)][N][2] fp32 (sum, sumsq) of v per 32-row slab and column (feeds GroupNorm);
 *   row_stats_out : optional [parts][M][2] fp32 partial (sum, sumsq) of v per row (feeds the next folded LayerNorm);
 *                   parts = seer_b200_gemm_row_parts(desc).
 * Requirements: K1, K2 % 64 == 0; N % 64 == 0 (geglu: % 128); lda/lda2/ldo_bf16 % 8 == 0; ldr/ldo_f32 % 4 == 0
 * (ldr % 8 for a bf16 residual); all bases 16-byte aligned; conv: Cin % 64 == 0, W | 128, (128/W) | H or H | (128/W).
 * col_stats is computed from the fp32 values when an fp32 output exists, from the stored bf16 values of a bf16-only output (the
 * statistics of exactly the tensor the consuming GroupNorm reads). */
typedef struct SeerGemmDesc {
  const void* A; int lda; int K1;
  const void* X; int n_img, H, W, Cin;
  const void* A2; int lda2; int K2;
  const void* Wt; int M, N;
  const float* bias; int ldb; int bias_div;
  const void* residual; int ldr; int residual_bf16;
  void* out_f32; int ldo_f32;
  void* out_bf16; int ldo_bf16;
  int geglu;
  float* col_stats;
  float* row_stats_out;
  const float* row_stats_in; int row_parts_in; float ln_eps; const float* ln_colsum;
  /* conv variants (X != NULL; all 0 = the 3x3 / stride-1 / pad-1 conv):
   *   conv_stride 2: Downsample3D (resnet.py:95-104) read straight from X through a strided TMA box; M = n_img*(H/2)*(W/2);
   *   conv_taps_w x conv_taps_h taps, tap t reading pixel (s*y + t / taps_w + conv_off_y, s*x + t % taps_w + conv_off_x),
   *     K1 = taps*Cin in the order [Cin/64][tap][64] (0 x 0 = 3 x 3 taps with offsets -1);
   *   out_up_phase 1 + (2 py + px): GEMM row m = low-res pixel (n, y, x) is stored at the row of pixel (n, 2y+py, 2x+px) of the
   *     [n_img, 2H, 2W] output and its col_stats slab at 4 (m / 32) + phase — Upsample3D (resnet.py:47-61: nearest 2x, then
   *     conv3x3) as four 2x2-tap convs on the low-res image with pre-summed weights (W must be a power of two). */
  int conv_stride; int conv_taps_w, conv_taps_h, conv_off_x, conv_off_y; int out_up_phase;
  /* rotary embedding fused into the epilogue (rope_tab != NULL; needs the LayerNorm fold, a bias and a bf16-only output):
   * output columns [0, rope_cols) are heads of width rope_d whose first 32 channels are rotated as interleaved pairs
   * (2j, 2j+1) by the angle of row position (r % rope_T): rope_tab[pos][16][2] = fp16 (cos, sin) from seer_b200_rope_table.
   * Replaces rotary_emb.rotate_queries_or_keys on q and k, attention.py:649-651, applied before the bf16 rounding. */
  const void* rope_tab; int rope_T; int rope_cols; int rope_d;
} SeerGemmDesc;
int seer_b200_gemm_ex(const SeerGemmDesc* desc, void* stream);
/* sizeof(SeerGemmDesc) as compiled into the library (binding sanity check) */
int seer_b200_gemm_desc_size(void);
/* number of per-row partial (sum, sumsq) pairs seer_b200_gemm_ex writes to row_stats_out for this descriptor (<= 0: error) */
int seer_b200_gemm_row_parts(const SeerGemmDesc* desc);

/* Frame-wise 3x3 conv, stride 1, pad 1, as implicit GEMM over 9 shifted TMA boxes of X[n_img,H,W,Cin] (bf16),
 * optional fused 1x1 tail A2[M,K2] (ResNet shortcut).  Wt[Cout, 9*Cin + K2] with K order [Cin/64][ky][kx][64] then tail.
 * Replaces InflatedConv3d(k=3): seer/models/resnet.py:8-16,147,155 and Upsample3D's conv :39.
 * Requires Cin % 64 == 0, W | 128, and (128/W) | H or H | (128/W). */
int seer_b200_conv3x3_bf16(const void* X, int n_img, int H, int W, int Cin, const void* A2, int lda2, int K2, const void* Wt,
                           int Cout, const float* bias, int ldb, int bias_div, const float* residual, int ldr, void* out,
                           int ldo, int flags, void* stream);

/* GroupNorm(32 groups) over (C/32, T) per sample on the virtual concat [x1 | x2] (fp32, [B*T, C1] and [B*T, C2]),
 * then optional SiLU; y is bf16 (or fp32 if y_is_f32) [B*T, C1+C2]; raw_bf16 (optional) receives the un-normalised
 * concat as bf16.  workspace: seer_b200_groupnorm_workspace_floats(B,T) floats; scale_shift: 2*B*(C1+C2) floats.
 * Replaces torch.nn.GroupNorm on the 5-D tensor: resnet.py:179-180,197-198; attention.py:133; unet_3d_condition.py:368-369. */
int seer_b200_groupnorm_workspace_floats(int B, int T);
int seer_b200_groupnorm(const float* x1, int C1, const float* x2, int C2, int B, int T, const float* gamma, const float* beta,
                        float eps, int silu, float* workspace, float* scale_shift, void* y, int y_is_f32, void* raw_bf16,
                        void* stream);

/* Same GroupNorm, but the statistics come from the per-(32-row slab, channel) partial sums a seer_b200_gemm_ex launch
 * emitted while producing x1 / x2 (SeerGemmDesc::col_stats, [B*T/32][Ci][2]): no statistics pass over the activation.
 * Requires T % 32 == 0. */
int seer_b200_groupnorm_from_stats(const float* x1, int C1, const float* stats1, const float* x2, int C2, const float* stats2,
                                   int B, int T, const float* gamma, const float* beta, float eps, int silu, float* scale_shift,
                                   void* y, int y_is_f32, void* raw_bf16, void* stream);

/* As above with bf16 sources (x1_is_bf16: x1 AND x2 are bf16 tensors — conv1's output, which only GroupNorm 2 of the ResNet
 * block reads, or block outputs / skip partners of the bf16 residual stream); needs a bf16 y. */
int seer_b200_groupnorm_from_stats_ex(const void* x1, int x1_is_bf16, int C1, const float* stats1, const void* x2, int C2,
                                      const float* stats2, int B, int T, const float* gamma, const float* beta, float eps, int silu,
                                      float* scale_shift, void* y, int y_is_f32, void* raw_bf16, void* stream);

/* LayerNorm over the last dim (fp32 in, bf16 out).  Replaces nn.LayerNorm: attention.py:198-200,237,244,311,322-323. */
int seer_b200_layernorm(const float* x, int M, int C, int ldx, const float* gamma, const float* beta, float eps, void* y_bf16,
                        int ldy, void* stream);

/* softmax(Q K^T / sqrt(d)) V over token-major bf16 buffers; row gathers implement the reference's head split and window
 * partition.  SPATIAL/CROSS: n_outer = b*f frames, sequences Lq/Lk rows per frame.  SCTA: n_outer = b, geometry (F,H,W),
 * window rule and causal order of attention.py:632-703 (Lq/Lk ignored).  FRAME: n_outer = clips, F = frames, H = tokens
 * per frame (W ignored): causal over the F frames of each (clip, token) — the FSText temporal blocks, attention.py:393-396,
 * 521-524.  head_dim in {40, 80, 96, 160}.
 * Replaces xformers.ops.memory_efficient_attention at attention.py:622-630. */
int seer_b200_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int mode,
                        int heads, int head_dim, int n_outer, int Lq, int Lk, int F, int H, int W, void* stream);

/* Writes the SCTA gather permutation used by seer_b200_attention: out_dev[(b*nwin + win)*L + s] = token row.
 * out_nwin / out_L are HOST pointers.  Parity hook for window_partition (attention.py:42-53). */
int seer_b200_scta_row_index(int B, int F, int H, int W, int* out_dev, int* out_nwin, int* out_L, void* stream);

/* RoPE (interleaved pairs, first 2*n_freqs channels of every head) in place on the Q and K column blocks of a bf16
 * buffer [M, ld]; position = row % tokens_per_clip.  Replaces rotary_emb.rotate_queries_or_keys, attention.py:649-651. */
int seer_b200_rope_inplace(void* qk_bf16, int ld, int M, int tokens_per_clip, int heads, int head_dim, int q_col, int k_col,
                           const float* freqs, int n_freqs, void* stream);

/* out[pos][j] = fp16 pair (cos, sin)(pos * freqs[j]) for pos < T, j < n_freqs (sincosf of the fp32 product — the same angles
 * seer_b200_rope_inplace uses — rounded to fp16: 11 significant bits in [-1, 1]): the table SeerGemmDesc::rope_tab points at.
 * out: T * n_freqs * 2 halves. */
int seer_b200_rope_table(const float* freqs, int n_freqs, int T, void* out, void* stream);

/* RoPE in place on the Q and K column blocks of a bf16 [M, ld] buffer from that table (16 frequencies = rotary dim 32; position =
 * row % tokens_per_clip): vectorised stand-alone pass, attention.py:649-651. */
int seer_b200_rope_apply_table(void* qk_bf16, int ld, long long M, int tokens_per_clip, int heads, int head_dim, int q_col, int k_col,
                               const void* tab, void* stream);

/* diffusers Timesteps(dim, flip_sin_to_cos, shift): t[B] (fp32) -> out[B, dim].  unet_3d_condition.py:307. */
int seer_b200_timestep_embedding(const float* t, float* out, int B, int dim, float shift, int flip_sin_to_cos, void* stream);

/* out[b,n] = silu_out?( dot(silu_in?(in[b,:]), W[n,:]) + bias[n] + add[n] ), fp32, B small.
 * time_embedding.linear_{1,2} (unet_3d_condition.py:308) and time_emb_proj (resnet.py:190-192). */
int seer_b200_small_linear(const float* in, int ldi, const float* W, const float* bias, const float* add, float* out, int ldo,
                           int B, int N, int K, int silu_in, int silu_out, void* stream);

/* conv_in (3x3, 4 -> Cout, fp32): x (B,4,F,H,W) -> out [B*F*H*W, Cout].  unet_3d_condition.py:94,311. */
int seer_b200_conv_in(const float* x, const float* w, const float* bias, float* out, int B, int Cin, int F, int H, int W,
                      int Cout, void* stream);
/* conv_in that also emits col_stats [B*F*H*W/32][Cout][2] (sum, sumsq per 32-row slab and channel, the layout of
 * SeerGemmDesc::col_stats) so the GroupNorms consuming its output (first ResNet norm1, last up-block skip concat) skip their
 * statistics pass; col_stats may be NULL.  Needs B*F*H*W % 32 == 0 when given. */
int seer_b200_conv_in_stats(const float* x, const float* w, const float* bias, float* out, float* col_stats, int B, int Cin, int F,
                            int H, int W, int Cout, void* stream);
/* The same with the output stored as bf16 (out_is_bf16: the bf16 residual stream between blocks); the statistics are taken from
 * the fp32 values either way. */
int seer_b200_conv_in_ex(const float* x, const float* w, const float* bias, void* out, int out_is_bf16, float* col_stats, int B,
                         int Cin, int F, int H, int W, int Cout, void* stream);
/* conv_in as a tensor-core GEMM: im2col of the 4-channel latent, x (B,4,F,H,W) fp32 -> a_bf16 [B*F*H*W, 64] (column c*9 + tap for the
 * 36 real taps, zeros above); multiply by the (Cout, 36 -> 64 zero-padded) bf16 weight with seer_b200_gemm_ex. */
int seer_b200_conv_in_im2col(const float* x, void* a_bf16, int B, int Cin, int F, int H, int W, void* stream);
/* conv_out (3x3, Cin -> <=4, fp32): x [B*F*H*W, Cin] -> out (B,Cout,F,H,W); w_packed[co][tap][Cin].  :205,370.  Any W (rows are
 * processed in 64-pixel segments); also the VAE decoder's 128 -> 3 output conv at 256x256. */
int seer_b200_conv_out(const float* x, const float* w_packed, const float* bias, float* out, int B, int Cin, int F, int H,
                       int W, int Cout, void* stream);

/* P[r, :] = softmax(scale * S[r, :]) (fp32 in, bf16 out), one row per warp.  The row softmax of the VAE's single-head d = 512
 * attention (diffusers 0.10.2 AttentionBlock behind vae.decode / vae.encode, utils/ddim_sampling_utils.py:39, inference.py:186):
 * scores and P V run as tcgen05 GEMMs around it. */
int seer_b200_softmax_rows(const float* S, int lds, long long rows, int L, float scale, void* P_bf16, int ldp, void* stream);

/* Token-major x [B*F*HW, ld] fp32 (first C columns) -> out (B, C, F, H, W) fp32: the layout change after conv_out when it runs
 * as a tensor-core implicit GEMM with zero-padded output channels (unet_3d_condition.py:370 returns (B, 4, F, H, W)). */
int seer_b200_tokens_to_nchw(const float* x, int ld, float* out, int B, int C, int F, int HW, void* stream);

/* nearest 2x upsample fp32 [n,H,W,C] -> bf16 [n,2H,2W,C] (resnet.py:52). */
int seer_b200_upsample2x_to_bf16(const float* x, void* y, int n_img, int H, int W, int C, void* stream);
/* pad-1 3x3 im2col, stride 1 or 2: [n,H,W,C] (fp32, or bf16 if in_is_bf16) -> bf16 [n*(H/s)*(W/s), 9*C], K order [C/64][tap][64]; C % 64 == 0.
 * Stride 2 = Downsample3D (resnet.py:95-104); stride 1 = fallback for image sizes the TMA-box conv does not tile. */
int seer_b200_im2col3x3_to_bf16(const void* x, int in_is_bf16, void* y, int n_img, int H, int W, int C, int stride, void* stream);
int seer_b200_cast_f32_to_bf16(const float* x, void* y, long long n, void* stream);

/* e = e_u + s (e_c - e_u) on frames >= cond_f; pred_x0 = (x - c1 e)/c2; x_prev = c3 pred_x0 + c4 e — fp32, bit-identical
 * to ldm/models/diffusion/ddim_video.py:209-211,229-237 (eta = 0).  eps: (2b|b, C, cond_f+F2, H, W). */
int seer_b200_cfg_ddim_update(const float* eps, const float* x, float* x_prev, float* pred_x0, int b, int C, int F2, int cond_f,
                              int HW, int use_cfg, float scale, float sqrt_one_minus_at, float sqrt_at, float sqrt_a_prev,
                              float dir_coef, void* stream);

/* The same update with the two CFG branches on two GPUs of one NVLink domain (CFG-branch split, the reference has no
 * counterpart: its `[uc; c]` batch at ddim_video.py:199-207 runs on one device).  This rank holds branch `branch`
 * (0 = unconditional, 1 = conditional) in eps_local (b, C, cond_f+F2, H, W).  One launch pushes the frames >= cond_f of
 * eps_local into `peer_recv` (the PARTNER's receive slot, b*C*F2*HW floats, mapped into this process: CUDA VMM / IPC peer
 * mapping), releases `seq` into `peer_flag` (partner's flag word), waits until `local_flag` (this rank's flag word, written
 * by the partner) reaches `seq`, then combines eps_local with `local_recv` (this rank's receive slot).  `counter` is a
 * zero-initialised device word private to this rank.  The caller alternates two (slot, flag) pairs by step parity and
 * increases `seq` (!= 0) every step; both ranks must launch with the same arguments apart from branch / pointers.
 * Bit-identical to seer_b200_cfg_ddim_update on the concatenated `[e_u; e_c]`.  Waits at most ~20 s, then traps. */
int seer_b200_cfg_ddim_update_p2p(const float* eps_local, int branch, float* peer_recv, const float* local_recv,
                                  unsigned* peer_flag, const unsigned* local_flag, unsigned* counter, unsigned seq,
                                  const float* x, float* x_prev, float* pred_x0, int b, int C, int F2, int cond_f, int HW,
                                  float scale, float sqrt_one_minus_at, float sqrt_at, float sqrt_a_prev, float dir_coef,
                                  void* stream);

/* ---- fp32-parity path (rel-L2 <= 1e-4 against the reference's fp32 PyTorch forward) --------------------------------
 * Contractions run on seer_b200_gemm_ex with error-compensated bf16 operands: A' = [a_hi | a_hi | a_lo] (this split),
 * W' = [w_hi | w_lo | w_hi] (host packing), so A'.W' = a_hi.w_hi + a_hi.w_lo + a_lo.w_hi with fp32 accumulation. */

/* x fp32 [rows_in, C] (ldx) -> out bf16 [rows_out, ldo]: out[r, col0+c] = out[r, Ctot+col0+c] = bf16(x), out[r, 2*Ctot+col0+c]
 * = bf16(x - bf16(x)).  col0 / Ctot place one part of a channel concat (torch.cat at unet_3d_blocks.py:596,712).
 * up_H > 0: rows are pixels of [up_n_img, up_H, up_W] images and the output is the nearest 2x upsampling
 * [up_n_img, 2 up_H, 2 up_W] (resnet.py:52), rows_out = 4 rows_in. */
int seer_b200_split3_bf16(const float* x, int ldx, long long rows_in, int C, void* out, int ldo, int Ctot, int col0, int up_n_img,
                          int up_H, int up_W, void* stream);
/* LayerNorm fp32 -> fp32 (attention.py:198-200). */
int seer_b200_layernorm_f32(const float* x, int M, int C, int ldx, const float* gamma, const float* beta, float eps, float* y,
                            int ldy, void* stream);
/* out[M, inner] = h[:, :inner] * gelu_erf(h[:, inner:])  (GEGLU, attention.py:791-793), exact erff. */
int seer_b200_geglu_f32(const float* h, int ldh, float* out, int ldo, long long M, int inner, void* stream);
/* RoPE in place on the Q / K column blocks of a fp32 (is_f32) or bf16 buffer; position = (row / pos_div) % pos_mod
 * (UNet SCTA: pos_div 1, pos_mod F*h*w, attention.py:649-651; FSText frame axis: pos_div 77, pos_mod F, :529-530). */
int seer_b200_rope_ex(void* qk, int is_f32, int ld, int M, int pos_div, int pos_mod, int heads, int head_dim, int q_col, int k_col,
                      const float* freqs, int n_freqs, void* stream);
/* fp32 softmax attention, same modes / row gathers as seer_b200_attention plus SEER_ATTN_FRAME (n_outer = clips,
 * F = frames, H = tokens per frame); head_dim in {40, 80, 96, 160}. */
int seer_b200_attention_f32(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* o, int ldo, int mode,
                            int heads, int head_dim, int n_outer, int Lq, int Lk, int F, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEER_B200_H */
