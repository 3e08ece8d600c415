"""Torch-tensor wrappers over the C ABI (include/seer_b200.h).

PyTorch is plumbing here: device memory, the current stream, dtype/shape checks that raise Python
exceptions (mirroring the reference's plain-exception convention, e.g. unet_3d_blocks.py:57,72).
Every op launches hand-written sm_100a kernels from libseer_b200.so on `torch.cuda.current_stream()`;
there is no PyTorch fallback.  `LAUNCHES` counts kernel launches (for bench.py's `gpu_launches`).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib

GEMM_OUT_BF16 = 1
GEMM_GEGLU = 2
ATTN_SPATIAL, ATTN_CROSS, ATTN_SCTA, ATTN_FRAME = 0, 1, 2, 3

LAUNCHES = 0           # kernels launched through this module (graph replays are counted by the caller)
PROFILE = None         # bench.py sets this to a list: (name, algorithmic flops, start event, end event) per GEMM/conv launch


class _Timed:
    """CUDA events on the launching stream around one kernel launch (only while ops.PROFILE is a list)."""

    def __init__(self, name: str, flops: float):
        self.name, self.flops = name, flops

    def __enter__(self):
        if PROFILE is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.append((self.name, self.flops, self.e0, self.e1))
        return False


def _timed_op(fn):
    """Decorator: CUDA-event timing of a (non-GEMM) op while ops.PROFILE is a list (bench / tools only)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*a, **k):
        if PROFILE is None:
            return fn(*a, **k)
        with _Timed(fn.__name__, 0.0):
            return fn(*a, **k)
    return wrapper


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]) -> Optional[ctypes.c_void_p]:
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _req(t: torch.Tensor, dtype: torch.dtype, name: str, ndim: Optional[int] = None) -> None:
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (seer_b200 has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name}: expected {ndim} dims, got shape {tuple(t.shape)}")
    if t.stride(-1) != 1:
        raise ValueError(f"{name}: last dim must be contiguous")


def _rows(t: torch.Tensor) -> Tuple[int, int]:
    """(rows, leading-dim) of a 2-D row-major view (possibly a column slice of a wider buffer)."""
    return t.shape[0], t.stride(0)


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
            out_dtype: torch.dtype = torch.float32, also_bf16=False, geglu: bool = False, col_stats: bool = False,
            row_stats: bool = False, ln: Optional[Tuple[torch.Tensor, torch.Tensor, float]] = None) -> GemmOut:
    """One tcgen05 GEMM / implicit-GEMM conv launch (SeerGemmDesc, include/seer_b200.h).

    acc = [a | a2] @ wt.T   (a:[M,K1] bf16, or x_img:[n_img,H,W,Cin] bf16 for the 3x3/pad-1 conv; a2:[M,K2] bf16)
    v   = LN-fold(acc) + bias[(row // bias_div)] + residual;  geglu: value * gelu(gate)
    `ln` = (row_stats [parts, M, 2] from the producer, colsum [N] fp32, eps) folds LayerNorm(a) into this GEMM (wt must
    hold W*gamma and bias beta@W.T (+b)).  `also_bf16` (True or a tensor) adds a bf16 copy next to an fp32 `out`.
    residual may be fp32 or bf16."""
    _req(wt, torch.bfloat16, "wt", 2)
    if not wt.is_contiguous():
        raise ValueError("wt must be contiguous")
    d = _lib.GemmDesc()
    keep = []
    if x_img is not None:
        _req(x_img, torch.bfloat16, "x_img", 4)
        if not x_img.is_contiguous():
            raise ValueError("x_img must be contiguous [n_img,H,W,Cin]")
        n_img, H, W, Cin = x_img.shape
        M, K1 = n_img * H * W, 9 * Cin
        d.X, d.n_img, d.H, d.W, d.Cin = x_img.data_ptr(), n_img, H, W, Cin
        dev = x_img.device
    else:
        _req(a, torch.bfloat16, "a", 2)
        M, K1 = a.shape
        d.A, d.lda, d.K1 = a.data_ptr(), a.stride(0), K1
        dev = a.device
    N = wt.shape[0]
    K2 = 0
    if a2 is not None:
        _req(a2, torch.bfloat16, "a2", 2)
        K2 = a2.shape[1]
        if a2.shape[0] != M:
            raise ValueError("a2 rows != a rows")
        d.A2, d.lda2, d.K2 = a2.data_ptr(), a2.stride(0), K2
    if wt.shape[1] != K1 + K2:
        raise ValueError(f"wt must be [N, {K1 + K2}], got {tuple(wt.shape)}")
    d.Wt, d.M, d.N = wt.data_ptr(), M, N
    n_out = N // 2 if geglu else N
    if out is None:
        out = torch.empty((M, n_out), device=dev, dtype=torch.bfloat16 if geglu else out_dtype)
    if out.dim() != 2 or out.shape[0] != M or out.shape[1] != n_out:
        raise ValueError(f"out shape {tuple(out.shape)} != ({M}, {n_out})")
    if out.dtype not in (torch.bfloat16, torch.float32):
        raise TypeError("out must be bf16 or fp32")
    _req(out, out.dtype, "out", 2)
    out16 = None
    if out.dtype == torch.float32:
        d.out_f32, d.ldo_f32 = out.data_ptr(), out.stride(0)
        if also_bf16 is not False and also_bf16 is not None:
            out16 = torch.empty((M, n_out), device=dev, dtype=torch.bfloat16) if also_bf16 is True else also_bf16
            _req(out16, torch.bfloat16, "also_bf16", 2)
            if out16.shape != out.shape:
                raise ValueError("also_bf16 shape != out shape")
            d.out_bf16, d.ldo_bf16 = out16.data_ptr(), out16.stride(0)
    else:
        if also_bf16 is not False and also_bf16 is not None:
            raise ValueError("also_bf16 needs an fp32 primary output")
        d.out_bf16, d.ldo_bf16 = out.data_ptr(), out.stride(0)
    if bias is not None:
        _req(bias, torch.float32, "bias")
        d.bias, d.ldb, d.bias_div = bias.data_ptr(), (bias.stride(0) if bias.dim() == 2 else N), bias_div
    if residual is not None:
        if residual.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("residual must be fp32 or bf16")
        _req(residual, residual.dtype, "residual", 2)
        if residual.shape[0] != M or residual.shape[1] != n_out:
            raise ValueError("residual shape != out shape")
        d.residual, d.ldr, d.residual_bf16 = residual.data_ptr(), residual.stride(0), int(residual.dtype == torch.bfloat16)
    d.geglu = int(geglu)
    cs = rs = None
    if col_stats:
        cs = torch.empty(((M + 31) // 32, N, 2), device=dev, dtype=torch.float32)
        d.col_stats = cs.data_ptr()
    if ln is not None:
        st, colsum, eps = ln
        _req(st, torch.float32, "ln row_stats", 3); _req(colsum, torch.float32, "ln colsum", 1)
        if st.shape[1] != M or st.shape[2] != 2 or not st.is_contiguous() or colsum.numel() != N:
            raise ValueError("ln: row_stats must be [parts, M, 2] contiguous and colsum [N]")
        d.row_stats_in, d.row_parts_in, d.ln_eps, d.ln_colsum = st.data_ptr(), st.shape[0], float(eps), colsum.data_ptr()
    L = _lib.lib()
    if row_stats:
        parts = L.seer_b200_gemm_row_parts(ctypes.byref(d))
        if parts <= 0:
            _lib.check(parts if parts < 0 else -2, "gemm_row_parts")
        rs = torch.empty((parts, M, 2), device=dev, dtype=torch.float32)
        d.row_stats_out = rs.data_ptr()
    name = f"{'conv3x3' if x_img is not None else 'gemm'} M={M} N={N} K={K1 + K2}"
    with _Timed(name, 2.0 * M * N * (K1 + K2)):
        rc = L.seer_b200_gemm_ex(ctypes.byref(d), _stream())
    if rc == -2 and x_img is not None:
        return None          # geometry not tileable by TMA boxes: caller falls back to explicit im2col
    _lib.check(rc, name)
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
    tensor — the statistics pass over the activation is skipped."""
    _req(x1, torch.float32, "x1", 2)
    M, C1 = x1.shape
    C2 = 0
    if x2 is not None:
        _req(x2, torch.float32, "x2", 2)
        C2 = x2.shape[1]
        if x2.shape[0] != M or not x2.is_contiguous():
            raise ValueError("x2 must be contiguous with the same rows as x1")
    if not x1.is_contiguous() or M % B:
        raise ValueError("x1 must be contiguous [B*T, C1]")
    T = M // B
    C = C1 + C2
    _req(gamma, torch.float32, "gamma", 1); _req(beta, torch.float32, "beta", 1)
    if gamma.numel() != C or beta.numel() != C:
        raise ValueError("gamma/beta size != C1 + C2")
    L = _lib.lib()
    ss = torch.empty(2 * B * C, device=x1.device, dtype=torch.float32)
    y = torch.empty((M, C), device=x1.device, dtype=out_dtype)
    raw = torch.empty((M, C), device=x1.device, dtype=torch.bfloat16) if want_raw else None
    use_stats = stats1 is not None and (x2 is None or stats2 is not None) and T % 32 == 0
    if use_stats:
        for nm, st, Ci in (("stats1", stats1, C1), ("stats2", stats2, C2)):
            if st is not None:
                _req(st, torch.float32, nm, 3)
                if tuple(st.shape) != (M // 32, Ci, 2) or not st.is_contiguous():
                    raise ValueError(f"{nm} must be contiguous [{M // 32}, {Ci}, 2]")
        rc = L.seer_b200_groupnorm_from_stats(_p(x1), C1, _p(stats1), _p(x2), C2, _p(stats2) if x2 is not None else None, B, T,
                                              _p(gamma), _p(beta), float(eps), int(silu), _p(ss), _p(y),
                                              int(out_dtype == torch.float32), _p(raw), _stream())
        _lib.check(rc, f"groupnorm_from_stats(B={B},T={T},C={C1}+{C2})")
        _count(2)
    else:
        ws = torch.empty(L.seer_b200_groupnorm_workspace_floats(B, T), device=x1.device, dtype=torch.float32)
        rc = L.seer_b200_groupnorm(_p(x1), C1, _p(x2), C2, B, T, _p(gamma), _p(beta), float(eps), int(silu), _p(ws), _p(ss),
                                   _p(y), int(out_dtype == torch.float32), _p(raw), _stream())
        _lib.check(rc, f"groupnorm(B={B},T={T},C={C1}+{C2})")
        _count(3)
    return (y, raw) if want_raw else y


@_timed_op
def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(x, torch.float32, "x", 2)
    M, C = x.shape
    if out is None:
        out = torch.empty((M, C), device=x.device, dtype=torch.bfloat16)
    rc = _lib.lib().seer_b200_layernorm(_p(x), M, C, x.stride(0), _p(gamma), _p(beta), float(eps), _p(out), out.stride(0), _stream())
    _lib.check(rc, f"layernorm(M={M},C={C})")
    _count()
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, mode: int, heads: int, n_outer: int, Lq: int = 0,
              Lk: int = 0, F: int = 0, H: int = 0, W: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q/k/v: 2-D token-major views [rows, heads*d] (column slices of wider buffers are fine), all bf16 (product path:
    tcgen05 / mma.sync kernels) or all fp32 (fp32-parity path: SIMT kernel, accurate exp2f)."""
    dt = q.dtype
    if dt not in (torch.bfloat16, torch.float32):
        raise TypeError(f"attention: expected bf16 or fp32 q/k/v, got {dt}")
    for n, t in (("q", q), ("k", k), ("v", v)):
        _req(t, dt, n, 2)
    C = q.shape[1]
    d = C // heads
    if out is None:
        out = torch.empty((q.shape[0], C), device=q.device, dtype=dt)
    _req(out, dt, "out", 2)
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
    fn = _lib.lib().seer_b200_attention if dt == torch.bfloat16 else _lib.lib().seer_b200_attention_f32
    with _Timed(name, flops):
        rc = fn(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out), out.stride(0),
                mode, heads, d, n_outer, Lq, Lk, F, H, W, _stream())
    _lib.check(rc, f"attention(mode={mode},heads={heads},d={d},outer={n_outer},Lq={Lq},Lk={Lk},F={F},H={H},W={W})")
    _count()
    return out


def scta_row_index(B: int, F: int, H: int, W: int, device="cuda") -> torch.Tensor:
    """(B, n_windows, L) int32 gather permutation used by the SCTA kernel (parity hook)."""
    nwin, L = ctypes.c_int(0), ctypes.c_int(0)
    lib = _lib.lib()
    _lib.check(lib.seer_b200_scta_row_index(B, F, H, W, None, ctypes.byref(nwin), ctypes.byref(L), None), "scta_row_index")
    out = torch.empty((B, nwin.value, L.value), device=device, dtype=torch.int32)
    _lib.check(lib.seer_b200_scta_row_index(B, F, H, W, _p(out), ctypes.byref(nwin), ctypes.byref(L), _stream()), "scta_row_index")
    return out


@_timed_op
def rope_inplace(qkv: torch.Tensor, tokens_per_clip: int, heads: int, head_dim: int, q_col: int, k_col: int,
                 freqs: torch.Tensor) -> None:
    _req(qkv, torch.bfloat16, "qkv", 2); _req(freqs, torch.float32, "freqs", 1)
    rc = _lib.lib().seer_b200_rope_inplace(_p(qkv), qkv.stride(0), qkv.shape[0], tokens_per_clip, heads, head_dim, q_col, k_col,
                                           _p(freqs), freqs.numel(), _stream())
    _lib.check(rc, "rope_inplace")
    _count()


@_timed_op
def timestep_embedding(t: torch.Tensor, dim: int, shift: float, flip_sin_to_cos: bool) -> torch.Tensor:
    _req(t, torch.float32, "t", 1)
    out = torch.empty((t.numel(), dim), device=t.device, dtype=torch.float32)
    rc = _lib.lib().seer_b200_timestep_embedding(_p(t), _p(out), t.numel(), dim, float(shift), int(flip_sin_to_cos), _stream())
    _lib.check(rc, "timestep_embedding")
    _count()
    return out


@_timed_op
def small_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], add: Optional[torch.Tensor] = None,
                 silu_in: bool = False, silu_out: bool = False) -> torch.Tensor:
    _req(x, torch.float32, "x", 2); _req(w, torch.float32, "w", 2)
    B, K = x.shape
    N = w.shape[0]
    out = torch.empty((B, N), device=x.device, dtype=torch.float32)
    rc = _lib.lib().seer_b200_small_linear(_p(x), x.stride(0), _p(w), _p(bias), _p(add), _p(out), N, B, N, K, int(silu_in),
                                           int(silu_out), _stream())
    _lib.check(rc, "small_linear")
    _count()
    return out


@_timed_op
def conv_in(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, col_stats: bool = False):
    """x:(B,4,F,H,W) fp32 -> [B*F*H*W, Cout] fp32.  col_stats=True also returns the [M/32, Cout, 2] per-slab channel
    (sum, sumsq) the consuming GroupNorms need (None when M % 32 != 0): returns (out, stats)."""
    _req(x, torch.float32, "x", 5)
    if not x.is_contiguous():
        raise ValueError("x must be contiguous (B,C,F,H,W)")
    B, Cin, F, H, W = x.shape
    Cout = w.shape[0]
    M = B * F * H * W
    out = torch.empty((M, Cout), device=x.device, dtype=torch.float32)
    st = torch.empty((M // 32, Cout, 2), device=x.device, dtype=torch.float32) if (col_stats and M % 32 == 0) else None
    rc = _lib.lib().seer_b200_conv_in_stats(_p(x), _p(w), _p(bias), _p(out), _p(st), B, Cin, F, H, W, Cout, _stream())
    _lib.check(rc, "conv_in")
    _count()
    return (out, st) if col_stats else out


@_timed_op
def conv_out(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, B: int, F: int, H: int, W: int) -> torch.Tensor:
    """x:[B*F*H*W, Cin] fp32 -> (B,Cout,F,H,W) fp32."""
    _req(x, torch.float32, "x", 2)
    Cin = x.shape[1]
    Cout = w_packed.shape[0]
    out = torch.empty((B, Cout, F, H, W), device=x.device, dtype=torch.float32)
    rc = _lib.lib().seer_b200_conv_out(_p(x), _p(w_packed), _p(bias), _p(out), B, Cin, F, H, W, Cout, _stream())
    _lib.check(rc, "conv_out")
    _count()
    return out


@_timed_op
def upsample2x(x: torch.Tensor, n_img: int, H: int, W: int) -> torch.Tensor:
    """x:[n_img*H*W, C] fp32 -> [n_img, 2H, 2W, C] bf16."""
    _req(x, torch.float32, "x", 2)
    C = x.shape[1]
    y = torch.empty((n_img, 2 * H, 2 * W, C), device=x.device, dtype=torch.bfloat16)
    rc = _lib.lib().seer_b200_upsample2x_to_bf16(_p(x), _p(y), n_img, H, W, C, _stream())
    _lib.check(rc, "upsample2x")
    _count()
    return y


@_timed_op
def im2col3x3(x: torch.Tensor, stride: int) -> torch.Tensor:
    """x:[n_img,H,W,C] fp32 or bf16 -> [n_img*(H/s)*(W/s), 9*C] bf16."""
    if x.dim() != 4 or not x.is_contiguous() or not x.is_cuda:
        raise ValueError("x must be a contiguous CUDA [n_img,H,W,C] tensor")
    n_img, H, W, C = x.shape
    y = torch.empty((n_img * (H // stride) * (W // stride), 9 * C), device=x.device, dtype=torch.bfloat16)
    rc = _lib.lib().seer_b200_im2col3x3_to_bf16(_p(x), int(x.dtype == torch.bfloat16), _p(y), n_img, H, W, C, stride, _stream())
    _lib.check(rc, "im2col3x3")
    _count()
    return y


@_timed_op
def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    _req(x, torch.float32, "x")
    if not x.is_contiguous():
        raise ValueError("x must be contiguous")
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    rc = _lib.lib().seer_b200_cast_f32_to_bf16(_p(x), _p(y), x.numel(), _stream())
    _lib.check(rc, "cast_f32_to_bf16")
    _count()
    return y


@_timed_op
def cfg_ddim_update(eps: torch.Tensor, x: torch.Tensor, cond_f: int, use_cfg: bool, scale: float, sqrt_one_minus_at: float,
                    sqrt_at: float, sqrt_a_prev: float, dir_coef: float, x_prev: Optional[torch.Tensor] = None,
                    pred_x0: Optional[torch.Tensor] = None):
    """eps:(2b|b,C,cond_f+F2,H,W) fp32, x:(b,C,F2,H,W) fp32 -> (x_prev, pred_x0)."""
    _req(eps, torch.float32, "eps", 5); _req(x, torch.float32, "x", 5)
    if not (eps.is_contiguous() and x.is_contiguous()):
        raise ValueError("eps and x must be contiguous")
    b, C, F2, H, W = x.shape
    if eps.shape != ((2 * b if use_cfg else b), C, F2 + cond_f, H, W):
        raise ValueError(f"eps shape {tuple(eps.shape)} inconsistent with x {tuple(x.shape)}")
    x_prev = torch.empty_like(x) if x_prev is None else x_prev
    pred_x0 = torch.empty_like(x) if pred_x0 is None else pred_x0
    rc = _lib.lib().seer_b200_cfg_ddim_update(_p(eps), _p(x), _p(x_prev), _p(pred_x0), b, C, F2, cond_f, H * W, int(use_cfg),
                                              float(scale), float(sqrt_one_minus_at), float(sqrt_at), float(sqrt_a_prev),
                                              float(dir_coef), _stream())
    _lib.check(rc, "cfg_ddim_update")
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
    _req(x, torch.float32, "x", 2)
    M, C = x.shape
    ctot = C if ctot is None else ctot
    rows_out = 4 * M if upsample else M
    if out is None:
        out = torch.empty((rows_out, 3 * ctot), device=x.device, dtype=torch.bfloat16)
    _req(out, torch.bfloat16, "out", 2)
    if out.shape[0] != rows_out or out.shape[1] != 3 * ctot:
        raise ValueError(f"split3: out shape {tuple(out.shape)} != ({rows_out}, {3 * ctot})")
    n_img, H, W = upsample if upsample else (0, 0, 0)
    rc = _lib.lib().seer_b200_split3_bf16(_p(x), x.stride(0), M, C, _p(out), out.stride(0), ctot, col0, n_img, H, W, _stream())
    _lib.check(rc, f"split3(M={M},C={C})")
    _count()
    return out


@_timed_op
def layernorm_f32(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    _req(x, torch.float32, "x", 2)
    M, C = x.shape
    out = torch.empty((M, C), device=x.device, dtype=torch.float32)
    rc = _lib.lib().seer_b200_layernorm_f32(_p(x), M, C, x.stride(0), _p(gamma), _p(beta), float(eps), _p(out), out.stride(0), _stream())
    _lib.check(rc, f"layernorm_f32(M={M},C={C})")
    _count()
    return out


@_timed_op
def geglu_f32(h: torch.Tensor) -> torch.Tensor:
    """h fp32 [M, 2I] -> [M, I] = h[:, :I] * gelu_erf(h[:, I:])."""
    _req(h, torch.float32, "h", 2)
    M, two_i = h.shape
    out = torch.empty((M, two_i // 2), device=h.device, dtype=torch.float32)
    rc = _lib.lib().seer_b200_geglu_f32(_p(h), h.stride(0), _p(out), out.stride(0), M, two_i // 2, _stream())
    _lib.check(rc, "geglu_f32")
    _count()
    return out


@_timed_op
def rope_ex(qk: torch.Tensor, pos_div: int, pos_mod: int, heads: int, head_dim: int, q_col: int, k_col: int,
            freqs: torch.Tensor) -> None:
    """RoPE in place on a bf16 or fp32 [M, ld] buffer; position = (row // pos_div) % pos_mod."""
    if qk.dtype not in (torch.bfloat16, torch.float32):
        raise TypeError("rope_ex: bf16 or fp32 buffer expected")
    _req(qk, qk.dtype, "qk", 2); _req(freqs, torch.float32, "freqs", 1)
    rc = _lib.lib().seer_b200_rope_ex(_p(qk), int(qk.dtype == torch.float32), qk.stride(0), qk.shape[0], pos_div, pos_mod, heads,
                                      head_dim, q_col, k_col, _p(freqs), freqs.numel(), _stream())
    _lib.check(rc, "rope_ex")
    _count()
