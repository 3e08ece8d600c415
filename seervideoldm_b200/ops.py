"""Python-level API of the seer_b200 kernels: output allocation and shape bookkeeping over `torch.ops.seer_b200.*`
(torch_ops.py — the C ABI of include/seer_b200.h registered as a torch custom-op library).

PyTorch is plumbing here: device memory, the current stream, the op registry.  The ops validate dtype / shape / device
and raise Python exceptions (mirroring the reference's plain-exception convention, e.g. unet_3d_blocks.py:57,72); every
one launches hand-written sm_100a kernels from libseer_b200.so on the current stream of the tensors' device; there is no
PyTorch or CPU fallback (a CPU tensor fails in the dispatcher, a missing library raises on import of the first op).
`LAUNCHES` counts kernel launches (for bench.py's `gpu_launches`).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib, torch_ops  # noqa: F401  (torch_ops registers the seer_b200 op library)

_ops = torch.ops.seer_b200

GEMM_OUT_BF16 = 1
GEMM_GEGLU = 2
ATTN_SPATIAL, ATTN_CROSS, ATTN_SCTA, ATTN_FRAME = 0, 1, 2, 3

LAUNCHES = 0           # kernels launched through this module (graph replays are counted by the caller)
PROFILE = None         # bench.py sets this to a list: (name, algorithmic flops, start event, end event) per GEMM/conv launch


class _Timed:
    """CUDA events on the launching stream around one kernel launch (only while ops.PROFILE is a list).  Records
    (name, executed FLOPs, start event, end event, algorithmic bytes = every operand / output tensor counted once, the
    library's description of the GEMM kernel the launch dispatched to)."""

    def __init__(self, name: str, flops: float, nbytes: float = 0.0):
        self.name, self.flops, self.nbytes = name, flops, nbytes

    def __enter__(self):
        if PROFILE is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record()
            detail = last_gemm_kernel() if (self.name.startswith("gemm") or self.name.startswith("conv")) and self.flops else ""
            PROFILE.append((self.name, self.flops, self.e0, self.e1, self.nbytes, detail))
        return False


def _timed_op(fn):
    """Decorator: CUDA-event timing of a (non-GEMM) op while ops.PROFILE is a list (bench / tools only)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*a, **k):
        if PROFILE is None:
            return fn(*a, **k)
        name = fn.__name__
        if name == "groupnorm":          # shape and input dtype in the label: the per-level passes differ by 16x in bytes
            x2 = a[1] if len(a) > 1 else k.get("x2")
            name = (f"groupnorm M={a[0].shape[0]} C={a[0].shape[1] + (x2.shape[1] if x2 is not None else 0)} "
                    f"{'bf16' if a[0].dtype == torch.bfloat16 else 'f32'}->{'f32' if k.get('out_dtype') == torch.float32 else 'bf16'}"
                    f"{' +raw' if k.get('want_raw') else ''}")
        with _Timed(name, 0.0):
            return fn(*a, **k)
    return wrapper


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def _cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (seer_b200 has no CPU path)")


class GemmOut:
    """Outputs of one `gemm_ex` launch: `out` (fp32 or bf16 primary), `out16` (extra bf16 copy when both were asked),
    `col_stats` ([ceil(M/32), N, 2] per-slab channel (sum, sumsq) for GroupNorm), `row_stats` ([parts, M, 2] per-row
    partial (sum, sumsq) for a LayerNorm folded into the next GEMM)."""
    __slots__ = ("out", "out16", "col_stats", "row_stats")

    def __init__(self, out, out16=None, col_stats=None, row_stats=None):
        self.out, self.out16, self.col_stats, self.row_stats = out, out16, col_stats, row_stats


def gemm_ex(a: Optional[torch.Tensor], wt: torch.Tensor, *, x_img: Optional[torch.Tensor] = None,
            a2: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None, bias_div: int = 0,
            residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
            out_dtype: torch.dtype = torch.float32, also_bf16=False, geglu: bool = False, col_stats=False,
            row_stats: bool = False, ln: Optional[Tuple[torch.Tensor, torch.Tensor, float]] = None, conv_stride: int = 1,
            conv_taps: Optional[Tuple[int, int, int, int]] = None, up_phase: int = 0,
            rope: Optional[Tuple[torch.Tensor, int, int]] = None) -> Optional[GemmOut]:
    """One tcgen05 GEMM / implicit-GEMM conv launch (SeerGemmDesc, include/seer_b200.h; torch.ops.seer_b200.gemm_ex).

    acc = [a | a2] @ wt.T   (a:[M,K1] bf16, or x_img:[n_img,H,W,Cin] bf16 for the 3x3/pad-1 conv; a2:[M,K2] bf16)
    v   = LN-fold(acc) + bias[(row // bias_div)] + residual;  geglu: value * gelu(gate)
    `ln` = (row_stats [parts, M, 2] from the producer, colsum [N] fp32, eps) folds LayerNorm(a) into this GEMM (wt must
    hold W*gamma and bias beta@W.T (+b)).  `also_bf16` (True or a tensor) adds a bf16 copy next to an fp32 `out`.
    residual may be fp32 or bf16.  `col_stats` True allocates the statistics, a tensor is filled in place.
    Conv variants (x_img only): `conv_stride` 2 = Downsample3D read through a strided TMA box; `conv_taps` =
    (taps_w, taps_h, off_x, off_y) replaces the 3x3 / offset -1 tap set; `up_phase` 1 + (2 py + px) scatters the rows to the
    (py, px) phase of the 2x upsampled output (out / col_stats then have 4x the GEMM's rows).
    `rope` = (table [T, 16, 2] from rope_table(), rotated columns, head dim): rotary embedding of the leading `cols` output
    columns (heads of width `head dim`, first 32 channels of each) by the row's position r % T, fused into the epilogue (needs
    `ln`, a bias and a bf16 output).
    Returns None when an x_img geometry is outside the TMA-box tiling (the caller falls back to im2col + GEMM)."""
    _cuda(wt, "wt")
    if wt.dim() != 2:
        raise ValueError("wt must be [N, K]")
    dev = wt.device
    if x_img is not None:
        if x_img.dim() != 4:
            raise ValueError("x_img must be [n_img,H,W,Cin]")
        n_img, H, W, Cin = x_img.shape
        ntaps = conv_taps[0] * conv_taps[1] if conv_taps is not None else 9
        if conv_stride not in (1, 2) or H % conv_stride or W % conv_stride:
            raise ValueError("conv_stride must be 1 or 2 and divide H and W")
        M, K1 = n_img * (H // conv_stride) * (W // conv_stride), ntaps * Cin
    else:
        if a is None or a.dim() != 2:
            raise ValueError("a must be [M, K]")
        M, K1 = a.shape
    N = wt.shape[0]
    K2 = a2.shape[1] if a2 is not None else 0
    n_out = N // 2 if geglu else N
    M_out = 4 * M if up_phase else M          # an upsample-phase launch fills a quarter of the rows of its output
    if up_phase and (out is None or x_img is None):
        raise ValueError("up_phase needs x_img and a caller-owned `out` shared by the four phase launches")
    if out is None:
        out = torch.empty((M, n_out), device=dev, dtype=torch.bfloat16 if geglu else out_dtype)
    if out.dtype not in (torch.bfloat16, torch.float32):
        raise TypeError("out must be bf16 or fp32")
    want16 = also_bf16 is not False and also_bf16 is not None
    out16 = None
    if out.dtype == torch.float32:
        o32, o16 = out, None
        if want16:
            o16 = out16 = torch.empty((M_out, n_out), device=dev, dtype=torch.bfloat16) if also_bf16 is True else also_bf16
    else:
        if want16:
            raise ValueError("also_bf16 needs an fp32 primary output")
        o32, o16 = None, out
    cs = None
    if col_stats is not False and col_stats is not None:
        cs = torch.empty(((M_out + 31) // 32, N, 2), device=dev, dtype=torch.float32) if col_stats is True else col_stats
    rs_in, colsum, eps = ln if ln is not None else (None, None, 0.0)
    taps = list(conv_taps) if conv_taps is not None else []
    rope_tab, rope_cols, rope_d = rope if rope is not None else (None, 0, 0)
    args = [a if x_img is None else None, x_img, a2, wt, bias, int(bias_div), residual, o32, o16, bool(geglu), cs, None, rs_in,
            float(eps), colsum, int(conv_stride), taps, int(up_phase), rope_tab, (rope_tab.shape[0] if rope is not None else 0),
            int(rope_cols), int(rope_d)]
    rs = None
    if row_stats:
        rs = torch.empty((_ops.gemm_row_parts(*args), M, 2), device=dev, dtype=torch.float32)
        args[11] = rs
    if x_img is None:
        name = f"gemm M={M} N={N} K={K1 + K2}"
    elif conv_stride != 1 or conv_taps is not None:
        name = f"conv{'2x2up' if up_phase else '3x3s2'} M={M} N={N} K={K1 + K2}"
    else:
        name = f"conv3x3 M={M} N={N} K={K1 + K2}"
    nbytes = 0.0
    if PROFILE is not None:          # algorithmic bytes of this launch: operands, weights, outputs, statistics — each once
        src = x_img if x_img is not None else a
        nbytes = sum(t.numel() * t.element_size() for t in (src, a2, wt, o32, o16, residual, cs, rs, rs_in) if t is not None)
        if up_phase:                 # a phase launch fills a quarter of the shared output / statistics tensors
            nbytes -= 0.75 * sum(t.numel() * t.element_size() for t in (o32, o16, cs) if t is not None)
    with _Timed(name, 2.0 * M * N * (K1 + K2), nbytes):
        rc = _ops.gemm_ex(*args)
    if rc != 0:
        return None          # geometry not tileable by TMA boxes: caller falls back to explicit im2col
    _count()
    return GemmOut(out, out16, cs, rs)


def gemm(a: torch.Tensor, wt: torch.Tensor, *, a2: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
         bias_div: int = 0, residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
         out_dtype: torch.dtype = torch.float32, geglu: bool = False) -> torch.Tensor:
    """out = [a | a2] @ wt.T + bias (+ residual).  a:[M,K1] bf16, a2:[M,K2] bf16, wt:[N,K1+K2] bf16."""
    return gemm_ex(a, wt, a2=a2, bias=bias, bias_div=bias_div, residual=residual, out=out, out_dtype=out_dtype, geglu=geglu).out


def conv3x3_ex(x: torch.Tensor, wt: torch.Tensor, **kw) -> GemmOut:
    """Frame-wise 3x3 conv (stride 1, pad 1).  x:[n_img,H,W,Cin] bf16 contiguous; wt:[Cout, 9*Cin (+K2)] bf16 with
    K order [Cin/64][ky][kx][64] (packing.pack_conv3x3); result rows [n_img*H*W, Cout].  Falls back to im2col + GEMM (still seer_b200 kernels) for
    image sizes the TMA-box tiling does not cover."""
    r = gemm_ex(None, wt, x_img=x, **kw)
    if r is None:
        cols = im2col3x3(x, stride=1)
        a2 = kw.pop("a2", None)
        a_full = cols if a2 is None else torch.cat([cols, a2], dim=1)
        r = gemm_ex(a_full, wt, **kw)
    return r


def conv3x3(x: torch.Tensor, wt: torch.Tensor, *, a2: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
            bias_div: int = 0, residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
            out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    return conv3x3_ex(x, wt, a2=a2, bias=bias, bias_div=bias_div, residual=residual, out=out, out_dtype=out_dtype).out


@_timed_op
def groupnorm(x1: torch.Tensor, x2: Optional[torch.Tensor], B: int, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
              silu: bool, out_dtype: torch.dtype = torch.bfloat16, want_raw: bool = False,
              stats1: Optional[torch.Tensor] = None, stats2: Optional[torch.Tensor] = None):
    """GroupNorm(32) over (C/32, T) per sample on the virtual concat [x1 | x2] (token-major fp32) (+SiLU).
    With `stats1` (and `stats2` when x2 is given) — the col_stats a gemm_ex launch emitted while producing the
    tensor — the statistics pass over the activation is skipped; x1 / x2 may then be bf16 tensors (a ResNet block's conv1
    output; block outputs of the bf16 residual stream)."""
    _cuda(x1, "x1")
    if x1.dim() != 2 or x1.shape[0] % B:
        raise ValueError("x1 must be contiguous [B*T, C1]")
    x1_bf16 = x1.dtype == torch.bfloat16
    M, C1 = x1.shape
    C2 = x2.shape[1] if x2 is not None else 0
    T = M // B
    C = C1 + C2
    use_stats = stats1 is not None and (x2 is None or stats2 is not None) and T % 32 == 0
    if x2 is not None and x2.dtype != x1.dtype:
        raise ValueError("the two halves of a concat GroupNorm input must share one dtype")
    if x1_bf16 and (not use_stats or out_dtype != torch.bfloat16):
        raise ValueError("a bf16 GroupNorm input needs producer statistics (T % 32 == 0) and a bf16 output")
    ss = torch.empty(2 * B * C, device=x1.device, dtype=torch.float32)
    y = torch.empty((M, C), device=x1.device, dtype=out_dtype)
    raw = torch.empty((M, C), device=x1.device, dtype=torch.bfloat16) if want_raw else None
    if use_stats:
        _ops.groupnorm_from_stats(x1, stats1, x2, stats2 if x2 is not None else None, B, gamma, beta, float(eps), bool(silu), ss, y, raw)
        _count(2)
    else:
        ws = torch.empty(_lib.lib().seer_b200_groupnorm_workspace_floats(B, T), device=x1.device, dtype=torch.float32)
        _ops.groupnorm(x1, x2, B, gamma, beta, float(eps), bool(silu), ws, ss, y, raw)
        _count(3)
    return (y, raw) if want_raw else y


@_timed_op
def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(x, "x")
    if out is None:
        out = torch.empty(tuple(x.shape), device=x.device, dtype=torch.bfloat16)
    _ops.layernorm(x, gamma, beta, float(eps), out)
    _count()
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, mode: int, heads: int, n_outer: int, Lq: int = 0,
              Lk: int = 0, F: int = 0, H: int = 0, W: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q/k/v: 2-D token-major views [rows, heads*d] (column slices of wider buffers are fine), all bf16 (product path:
    tcgen05 / mma.sync kernels) or all fp32 (fp32-parity path: SIMT kernel, accurate exp2f)."""
    _cuda(q, "q")
    dt = q.dtype
    if dt not in (torch.bfloat16, torch.float32):
        raise TypeError(f"attention: expected bf16 or fp32 q/k/v, got {dt}")
    C = q.shape[1]
    d = C // heads
    if out is None:
        out = torch.empty((q.shape[0], C), device=q.device, dtype=dt)
    if mode == ATTN_FRAME:
        name, flops = f"attention frame d={d} L={F} x{n_outer * H * heads}", 4.0 * F * (F + 1) / 2 * d * n_outer * H * heads
    elif mode == ATTN_SCTA:
        ws = 0 if H <= 4 else (8 if H // 8 >= 4 else 4)
        L = F * (ws * ws if ws else H * W)
        nprob = n_outer * heads * ((H // ws) * (W // ws) if ws else 1)
        name, flops = f"attention scta d={d} L={L} x{nprob}", 4.0 * L * (L + 1) / 2 * d * nprob
    else:
        name = f"attention {'spatial' if mode == ATTN_SPATIAL else 'cross'} d={d} Lq={Lq} Lk={Lk} x{n_outer * heads}"
        flops = 4.0 * Lq * Lk * d * n_outer * heads
    with _Timed(name, flops):
        _ops.attention(q, k, v, out, int(mode), int(heads), int(n_outer), int(Lq), int(Lk), int(F), int(H), int(W))
    _count()
    return out


def scta_row_index(B: int, F: int, H: int, W: int, device="cuda") -> torch.Tensor:
    """(B, n_windows, L) int32 gather permutation used by the SCTA kernel (parity hook)."""
    import ctypes
    nwin, L = ctypes.c_int(0), ctypes.c_int(0)
    _lib.check(_lib.lib().seer_b200_scta_row_index(B, F, H, W, None, ctypes.byref(nwin), ctypes.byref(L), None), "scta_row_index")
    out = torch.empty((B, nwin.value, L.value), device=device, dtype=torch.int32)
    _ops.scta_row_index(out, B, F, H, W)
    return out


@_timed_op
def softmax_rows(S: torch.Tensor, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bf16 softmax(scale * S) along the last dim of an fp32 [rows, L] matrix (VAE attention block)."""
    _cuda(S, "S")
    if out is None:
        out = torch.empty(tuple(S.shape), device=S.device, dtype=torch.bfloat16)
    _ops.softmax_rows(S, float(scale), out)
    _count()
    return out


def rope_table(freqs: torch.Tensor, T: int) -> torch.Tensor:
    """[T, n_freqs, 2] fp16 (cos, sin)(pos * freqs[j]) — the table the q/k/v GEMM's fused rotary epilogue reads."""
    _cuda(freqs, "freqs")
    out = torch.empty((T, freqs.numel(), 2), device=freqs.device, dtype=torch.float16)
    _ops.rope_table(freqs, out)
    _count()
    return out


@_timed_op
def rope_inplace(qkv: torch.Tensor, tokens_per_clip: int, heads: int, head_dim: int, q_col: int, k_col: int,
                 freqs: torch.Tensor, tab: Optional[torch.Tensor] = None) -> None:
    """RoPE on the Q / K column blocks of a bf16 [M, ld] buffer; position = row % tokens_per_clip (attention.py:649-651).
    With `tab` (rope_table(freqs, tokens_per_clip)) the vectorised table kernel runs, else sincosf per (token, pair)."""
    _cuda(qkv, "qkv")
    if qkv.dtype != torch.bfloat16:
        raise TypeError("rope_inplace: bf16 buffer expected")
    if qkv.shape[0] % tokens_per_clip:
        raise ValueError("rope_inplace: rows must be a multiple of tokens_per_clip")
    if tab is not None and head_dim % 8 == 0 and q_col % 8 == 0 and k_col % 8 == 0 and qkv.stride(0) % 8 == 0:
        _ops.rope_apply_table(qkv, int(tokens_per_clip), int(heads), int(head_dim), int(q_col), int(k_col), tab)
    else:
        _ops.rope_inplace(qkv, int(tokens_per_clip), int(heads), int(head_dim), int(q_col), int(k_col), freqs)
    _count()


@_timed_op
def timestep_embedding(t: torch.Tensor, dim: int, shift: float, flip_sin_to_cos: bool) -> torch.Tensor:
    _cuda(t, "t")
    out = torch.empty((t.numel(), dim), device=t.device, dtype=torch.float32)
    _ops.timestep_embedding(t, out, float(shift), bool(flip_sin_to_cos))
    _count()
    return out


@_timed_op
def small_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], add: Optional[torch.Tensor] = None,
                 silu_in: bool = False, silu_out: bool = False) -> torch.Tensor:
    _cuda(x, "x")
    out = torch.empty((x.shape[0], w.shape[0]), device=x.device, dtype=torch.float32)
    _ops.small_linear(x, w, bias, add, out, bool(silu_in), bool(silu_out))
    _count()
    return out


@_timed_op
def conv_in(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, col_stats: bool = False, out_dtype: torch.dtype = torch.float32):
    """x:(B,4,F,H,W) fp32 -> [B*F*H*W, Cout] fp32 (or bf16).  col_stats=True also returns the [M/32, Cout, 2] per-slab channel
    (sum, sumsq of the fp32 values) the consuming GroupNorms need (None when M % 32 != 0): returns (out, stats)."""
    _cuda(x, "x")
    if x.dim() != 5:
        raise ValueError("x must be contiguous (B,C,F,H,W)")
    B, Cin, F, H, W = x.shape
    Cout = w.shape[0]
    M = B * F * H * W
    out = torch.empty((M, Cout), device=x.device, dtype=out_dtype)
    st = torch.empty((M // 32, Cout, 2), device=x.device, dtype=torch.float32) if (col_stats and M % 32 == 0) else None
    _ops.conv_in(x, w, bias, out, st)
    _count()
    return (out, st) if col_stats else out


@_timed_op
def conv_in_im2col(x: torch.Tensor) -> torch.Tensor:
    """x:(B,4,F,H,W) fp32 -> bf16 [B*F*H*W, 64]: the 36 taps of the 4-channel 3x3 conv (column c*9 + tap), zero-padded to one
    k-block; `gemm_ex(a, w16)` with the (Cout, 64) zero-padded bf16 weight is conv_in on the tensor cores."""
    _cuda(x, "x")
    B, Cin, F, H, W = x.shape
    out = torch.empty((B * F * H * W, 64), device=x.device, dtype=torch.bfloat16)
    _ops.conv_in_im2col(x, out)
    _count()
    return out


@_timed_op
def conv_out(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, B: int, F: int, H: int, W: int) -> torch.Tensor:
    """x:[B*F*H*W, Cin] fp32 -> (B,Cout,F,H,W) fp32."""
    _cuda(x, "x")
    out = torch.empty((B, w_packed.shape[0], F, H, W), device=x.device, dtype=torch.float32)
    _ops.conv_out(x, w_packed, bias, out)
    _count()
    return out


@_timed_op
def tokens_to_nchw(x: torch.Tensor, B: int, C: int, F: int, H: int, W: int) -> torch.Tensor:
    """x:[B*F*H*W, >= C] fp32 token-major -> (B, C, F, H, W) fp32 (the first C columns)."""
    _cuda(x, "x")
    out = torch.empty((B, C, F, H, W), device=x.device, dtype=torch.float32)
    _ops.tokens_to_nchw(x, out)
    _count()
    return out


@_timed_op
def upsample2x(x: torch.Tensor, n_img: int, H: int, W: int) -> torch.Tensor:
    """x:[n_img*H*W, C] fp32 -> [n_img, 2H, 2W, C] bf16."""
    _cuda(x, "x")
    y = torch.empty((n_img, 2 * H, 2 * W, x.shape[1]), device=x.device, dtype=torch.bfloat16)
    _ops.upsample2x(x, y)
    _count()
    return y


@_timed_op
def im2col3x3(x: torch.Tensor, stride: int) -> torch.Tensor:
    """x:[n_img,H,W,C] fp32 or bf16 -> [n_img*(H/s)*(W/s), 9*C] bf16."""
    _cuda(x, "x")
    if x.dim() != 4:
        raise ValueError("x must be a contiguous CUDA [n_img,H,W,C] tensor")
    n_img, H, W, C = x.shape
    y = torch.empty((n_img * (H // stride) * (W // stride), 9 * C), device=x.device, dtype=torch.bfloat16)
    _ops.im2col3x3(x, y, int(stride))
    _count()
    return y


@_timed_op
def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    _cuda(x, "x")
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _ops.cast_bf16(x, y)
    _count()
    return y


@_timed_op
def cfg_ddim_update(eps: torch.Tensor, x: torch.Tensor, cond_f: int, use_cfg: bool, scale: float, sqrt_one_minus_at: float,
                    sqrt_at: float, sqrt_a_prev: float, dir_coef: float, x_prev: Optional[torch.Tensor] = None,
                    pred_x0: Optional[torch.Tensor] = None):
    """eps:(2b|b,C,cond_f+F2,H,W) fp32, x:(b,C,F2,H,W) fp32 -> (x_prev, pred_x0)."""
    _cuda(eps, "eps")
    x_prev = torch.empty_like(x) if x_prev is None else x_prev
    pred_x0 = torch.empty_like(x) if pred_x0 is None else pred_x0
    _ops.cfg_ddim_update(eps, x, x_prev, pred_x0, int(cond_f), bool(use_cfg), float(scale), float(sqrt_one_minus_at), float(sqrt_at),
                         float(sqrt_a_prev), float(dir_coef))
    _count()
    return x_prev, pred_x0


@_timed_op
def cfg_ddim_update_p2p(eps_local: torch.Tensor, branch: int, peer_recv: torch.Tensor, local_recv: torch.Tensor,
                        peer_flag: torch.Tensor, local_flag: torch.Tensor, counter: torch.Tensor, seq: int, x: torch.Tensor,
                        cond_f: int, scale: float, sqrt_one_minus_at: float, sqrt_at: float, sqrt_a_prev: float, dir_coef: float):
    """CFG-branch split: eps_local (b,C,cond_f+F2,H,W) is this rank's branch; the partner's arrives through `local_recv`
    (NVLink peer memory, see parallel.CfgPeerExchange) inside the same kernel -> (x_prev, pred_x0)."""
    _cuda(eps_local, "eps_local")
    x_prev, pred_x0 = torch.empty_like(x), torch.empty_like(x)
    _ops.cfg_ddim_update_p2p(eps_local, int(branch), peer_recv, local_recv, peer_flag, local_flag, counter, int(seq), x, x_prev,
                             pred_x0, int(cond_f), float(scale), float(sqrt_one_minus_at), float(sqrt_at), float(sqrt_a_prev),
                             float(dir_coef))
    _count()
    return x_prev, pred_x0


# ---------------------------------------------------------------------------------------------------------------------
# fp32-parity path (csrc/fp32_path.cu): error-compensated bf16 operands for the tcgen05 GEMM, fp32 everywhere else
# ---------------------------------------------------------------------------------------------------------------------
@_timed_op
def split3(x: torch.Tensor, out: Optional[torch.Tensor] = None, col0: int = 0, ctot: Optional[int] = None,
           upsample: Optional[Tuple[int, int, int]] = None) -> torch.Tensor:
    """x fp32 [M, C] -> bf16 [M', 3*ctot] = [hi | hi | lo] (this part at columns col0.. of each third).
    `upsample` = (n_img, H, W): nearest 2x upsampling of the [n_img, H, W] pixel rows while splitting (M' = 4 M)."""
    _cuda(x, "x")
    M, C = x.shape
    ctot = C if ctot is None else ctot
    rows_out = 4 * M if upsample else M
    if out is None:
        out = torch.empty((rows_out, 3 * ctot), device=x.device, dtype=torch.bfloat16)
    n_img, H, W = upsample if upsample else (0, 0, 0)
    _ops.split3(x, out, int(ctot), int(col0), int(n_img), int(H), int(W))
    _count()
    return out


@_timed_op
def layernorm_f32(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    _cuda(x, "x")
    out = torch.empty(tuple(x.shape), device=x.device, dtype=torch.float32)
    _ops.layernorm(x, gamma, beta, float(eps), out)
    _count()
    return out


@_timed_op
def geglu_f32(h: torch.Tensor) -> torch.Tensor:
    """h fp32 [M, 2I] -> [M, I] = h[:, :I] * gelu_erf(h[:, I:])."""
    _cuda(h, "h")
    out = torch.empty((h.shape[0], h.shape[1] // 2), device=h.device, dtype=torch.float32)
    _ops.geglu_f32(h, out)
    _count()
    return out


@_timed_op
def rope_ex(qk: torch.Tensor, pos_div: int, pos_mod: int, heads: int, head_dim: int, q_col: int, k_col: int,
            freqs: torch.Tensor) -> None:
    """RoPE in place on a bf16 or fp32 [M, ld] buffer; position = (row // pos_div) % pos_mod."""
    _cuda(qk, "qk")
    _ops.rope(qk, int(pos_div), int(pos_mod), int(heads), int(head_dim), int(q_col), int(k_col), freqs)
    _count()


def last_attention_kernel() -> str:
    """Which kernel seer_b200_attention dispatched on this thread's last call (debug hook, include/seer_b200.h)."""
    return _lib.lib().seer_b200_debug_last_attention().decode()


def last_gemm_kernel() -> str:
    return _lib.lib().seer_b200_debug_last_gemm().decode()


def set_tuning(name: str, value: Optional[int]) -> None:
    """Override (or with None: reset to the built-in default) one of the library's SEER_* tuning switches in this process."""
    _lib.lib().seer_b200_debug_setenv(name.encode(), int(value or 0), int(value is None))
