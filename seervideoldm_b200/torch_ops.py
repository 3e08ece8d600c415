"""`torch.ops.seer_b200.*` — the C ABI of libseer_b200.so (include/seer_b200.h) registered as a torch custom-op library.

SURVEY §8(b): "one shared library exporting extern "C" launchers ... wrapped by TORCH_LIBRARY(seer_b200, m) ops that
check dtype / shape / contiguity / device and raise Python exceptions — mirroring the reference's convention (plain
Python exceptions: ValueError unet_3d_blocks.py:57,72; assert resnet.py:48,96)".  The library itself stays torch-free
(plain pointers and sizes, C99 host test in tests/test_host_cpu.py); this module is the binding a PyTorch caller sees:

  * every op is defined with a schema (mutated outputs / workspaces are caller-owned `Tensor(a!)` arguments — the ABI
    allocates nothing), implemented for the CUDA dispatch key only (a CPU tensor fails in the dispatcher: there is no
    CPU path), and has a fake (meta) kernel so FakeTensor / torch.compile tracing see the op without launching;
  * the implementation validates its tensors, makes the tensors' device current, and launches on that device's current
    stream (graph-capturable: no host sync, no allocation);
  * ops return the ABI's status only where the caller branches on it (`-2` = geometry outside the kernel's tiling, the
    caller falls back to another seer_b200 kernel); every other nonzero status raises.

`seervideoldm_b200.ops` holds the Python-level API (output allocation, shape bookkeeping) and calls these ops.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch

from . import _lib

NS = "seer_b200"
_library = torch.library.Library(NS, "DEF")
OP_NAMES: List[str] = []

bf16, f32 = torch.bfloat16, torch.float32


# --------------------------------------------------------------------------------------------------------------------
# argument checking helpers (TORCH_CHECK equivalents: raise TypeError / ValueError)
# --------------------------------------------------------------------------------------------------------------------
def _chk(t: Optional[torch.Tensor], name: str, dtype=None, ndim: Optional[int] = None, contiguous: bool = False,
         optional: bool = False, dev: Optional[torch.device] = None) -> None:
    if t is None:
        if optional:
            return
        raise ValueError(f"{name}: tensor required")
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (seer_b200 has no CPU path)")
    if dev is not None and t.device != dev:
        raise ValueError(f"{name}: on {t.device}, expected {dev} (all tensors of one call must share a device)")
    if dtype is not None:
        ok = t.dtype in dtype if isinstance(dtype, (tuple, list)) else t.dtype == dtype
        if not ok:
            raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name}: expected {ndim} dims, got shape {tuple(t.shape)}")
    if t.dim() and t.stride(-1) != 1:
        raise ValueError(f"{name}: last dim must be contiguous")
    if contiguous and not t.is_contiguous():
        raise ValueError(f"{name}: must be contiguous")
    if t.data_ptr() % 16:
        raise ValueError(f"{name}: base address must be 16-byte aligned")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class _Dev:
    """Make `t.device` current for the launch and hand out its current stream."""

    def __init__(self, t: torch.Tensor):
        self.guard = torch.cuda.device(t.device)

    def __enter__(self):
        self.guard.__enter__()
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def __exit__(self, *exc):
        return self.guard.__exit__(*exc)


def _define(schema: str, impl, fake=None) -> None:
    name = schema.split("(", 1)[0].strip()
    _library.define(schema)
    _library.impl(name, impl, "CUDA")
    torch.library.register_fake(f"{NS}::{name}", fake if fake is not None else (lambda *a, **k: None))
    OP_NAMES.append(name)


# --------------------------------------------------------------------------------------------------------------------
# GEMM / implicit-GEMM conv (SeerGemmDesc)
# --------------------------------------------------------------------------------------------------------------------
_GEMM_ARGS = ("Tensor? a, Tensor? x_img, Tensor? a2, Tensor wt, Tensor? bias, int bias_div, Tensor? residual, "
              "Tensor(a!)? out_f32, Tensor(b!)? out_bf16, bool geglu, Tensor(c!)? col_stats, Tensor(d!)? row_stats_out, "
              "Tensor? row_stats_in, float ln_eps, Tensor? ln_colsum, int conv_stride, int[] conv_taps, int up_phase, "
              "Tensor? rope_tab, int rope_T, int rope_cols, int rope_d")


def _gemm_desc(a, x_img, a2, wt, bias, bias_div, residual, out_f32, out_bf16, geglu, col_stats, row_stats_out, row_stats_in,
               ln_eps, ln_colsum, conv_stride, conv_taps, up_phase, rope_tab, rope_T, rope_cols, rope_d) -> "_lib.GemmDesc":
    dev = wt.device
    _chk(wt, "wt", bf16, 2, contiguous=True)
    d = _lib.GemmDesc()
    if x_img is not None:
        _chk(x_img, "x_img", bf16, 4, contiguous=True, dev=dev)
        n_img, H, W, Cin = x_img.shape
        ntaps = 9
        if len(conv_taps):
            if len(conv_taps) != 4:
                raise ValueError("conv_taps = (taps_w, taps_h, off_x, off_y)")
            d.conv_taps_w, d.conv_taps_h, d.conv_off_x, d.conv_off_y = (int(v) for v in conv_taps)
            ntaps = d.conv_taps_w * d.conv_taps_h
        if conv_stride not in (1, 2) or H % conv_stride or W % conv_stride:
            raise ValueError("conv_stride must be 1 or 2 and divide H and W")
        d.conv_stride, d.out_up_phase = conv_stride, up_phase
        M, K1 = n_img * (H // conv_stride) * (W // conv_stride), ntaps * Cin
        d.X, d.n_img, d.H, d.W, d.Cin = x_img.data_ptr(), n_img, H, W, Cin
    else:
        _chk(a, "a", bf16, 2, dev=dev)
        if len(conv_taps) or conv_stride != 1 or up_phase:
            raise ValueError("conv_stride / conv_taps / up_phase need x_img")
        M, K1 = a.shape
        d.A, d.lda, d.K1 = a.data_ptr(), a.stride(0), K1
    N = wt.shape[0]
    K2 = 0
    if a2 is not None:
        _chk(a2, "a2", bf16, 2, dev=dev)
        K2 = a2.shape[1]
        if a2.shape[0] != M:
            raise ValueError("a2 rows != a rows")
        d.A2, d.lda2, d.K2 = a2.data_ptr(), a2.stride(0), K2
    if wt.shape[1] != K1 + K2:
        raise ValueError(f"wt must be [N, {K1 + K2}], got {tuple(wt.shape)}")
    d.Wt, d.M, d.N = wt.data_ptr(), M, N
    n_out = N // 2 if geglu else N
    M_out = 4 * M if up_phase else M
    if out_f32 is None and out_bf16 is None:
        raise ValueError("out_f32 or out_bf16 required")
    for name, o, dt in (("out_f32", out_f32, f32), ("out_bf16", out_bf16, bf16)):
        if o is not None:
            _chk(o, name, dt, 2, dev=dev)
            if tuple(o.shape) != (M_out, n_out):
                raise ValueError(f"{name} shape {tuple(o.shape)} != ({M_out}, {n_out})")
    if out_f32 is not None:
        d.out_f32, d.ldo_f32 = out_f32.data_ptr(), out_f32.stride(0)
    if out_bf16 is not None:
        d.out_bf16, d.ldo_bf16 = out_bf16.data_ptr(), out_bf16.stride(0)
    if bias is not None:
        _chk(bias, "bias", f32, dev=dev)
        if bias.shape[-1] != N:
            raise ValueError("bias must have N columns")
        d.bias, d.ldb, d.bias_div = bias.data_ptr(), (bias.stride(0) if bias.dim() == 2 else N), bias_div
    if residual is not None:
        _chk(residual, "residual", (f32, bf16), 2, dev=dev)
        if tuple(residual.shape) != (M, n_out):
            raise ValueError("residual shape != out shape")
        d.residual, d.ldr, d.residual_bf16 = residual.data_ptr(), residual.stride(0), int(residual.dtype == bf16)
    d.geglu = int(geglu)
    if col_stats is not None:
        _chk(col_stats, "col_stats", f32, 3, contiguous=True, dev=dev)
        if tuple(col_stats.shape) != ((M_out + 31) // 32, N, 2):
            raise ValueError(f"col_stats must be [{(M_out + 31) // 32}, {N}, 2]")
        d.col_stats = col_stats.data_ptr()
    if row_stats_in is not None:
        _chk(row_stats_in, "row_stats_in", f32, 3, contiguous=True, dev=dev)
        _chk(ln_colsum, "ln_colsum", f32, 1, dev=dev)
        if row_stats_in.shape[1] != M or row_stats_in.shape[2] != 2 or ln_colsum.numel() != N:
            raise ValueError("ln: row_stats_in must be [parts, M, 2] and ln_colsum [N]")
        d.row_stats_in, d.row_parts_in, d.ln_eps, d.ln_colsum = (row_stats_in.data_ptr(), row_stats_in.shape[0], float(ln_eps),
                                                                 ln_colsum.data_ptr())
    if row_stats_out is not None:
        _chk(row_stats_out, "row_stats_out", f32, 3, contiguous=True, dev=dev)
        if row_stats_out.shape[1] != M or row_stats_out.shape[2] != 2:
            raise ValueError("row_stats_out must be [parts, M, 2]")
        d.row_stats_out = row_stats_out.data_ptr()
    if rope_tab is not None:
        _chk(rope_tab, "rope_tab", torch.float16, 3, contiguous=True, dev=dev)
        if tuple(rope_tab.shape) != (rope_T, 16, 2):
            raise ValueError(f"rope_tab must be [rope_T = {rope_T}, 16, 2] (seer_b200.rope_table)")
        d.rope_tab, d.rope_T, d.rope_cols, d.rope_d = rope_tab.data_ptr(), rope_T, rope_cols, rope_d
    return d


def _gemm_ex(*args) -> int:
    d = _gemm_desc(*args)
    wt, row_stats_out = args[3], args[11]
    with _Dev(wt) as stream:
        L = _lib.lib()
        if row_stats_out is not None:
            want = L.seer_b200_gemm_row_parts(ctypes.byref(d))
            if want != row_stats_out.shape[0]:
                raise ValueError(f"row_stats_out needs {want} partials per row (torch.ops.seer_b200.gemm_row_parts), got {row_stats_out.shape[0]}")
        rc = L.seer_b200_gemm_ex(ctypes.byref(d), stream)
    if rc == -2 and d.X:
        return rc              # geometry outside the TMA-box tiling: the caller falls back to im2col + GEMM
    _lib.check(rc, f"gemm_ex(M={d.M},N={d.N})")
    return 0


def _gemm_row_parts(*args) -> int:
    d = _gemm_desc(*args)
    with _Dev(args[3]):
        parts = _lib.lib().seer_b200_gemm_row_parts(ctypes.byref(d))
    if parts <= 0:
        _lib.check(parts if parts < 0 else -2, "gemm_row_parts")
    return parts


_define(f"gemm_ex({_GEMM_ARGS}) -> int", _gemm_ex, lambda *a, **k: 0)
_define(f"gemm_row_parts({_GEMM_ARGS}) -> int", _gemm_row_parts, lambda *a, **k: 1)


# --------------------------------------------------------------------------------------------------------------------
# norms
# --------------------------------------------------------------------------------------------------------------------
def _gn_common(x1, x2, B, gamma, beta, scale_shift, y, raw):
    dev = x1.device
    _chk(x1, "x1", (f32, bf16), 2, contiguous=True)
    _chk(x2, "x2", x1.dtype, 2, contiguous=True, optional=True, dev=dev)      # the two halves of a concat share one dtype
    M, C1 = x1.shape
    C2 = x2.shape[1] if x2 is not None else 0
    if M % B or (x2 is not None and x2.shape[0] != M):
        raise ValueError("x1 / x2 must be [B*T, C] with the same rows")
    C = C1 + C2
    _chk(gamma, "gamma", f32, 1, dev=dev); _chk(beta, "beta", f32, 1, dev=dev)
    if gamma.numel() != C or beta.numel() != C:
        raise ValueError("gamma/beta size != C1 + C2")
    _chk(scale_shift, "scale_shift", f32, 1, dev=dev)
    if scale_shift.numel() < 2 * B * C:
        raise ValueError("scale_shift needs 2*B*C floats")
    _chk(y, "y", (f32, bf16), 2, contiguous=True, dev=dev)
    _chk(raw, "raw", bf16, 2, contiguous=True, optional=True, dev=dev)
    if tuple(y.shape) != (M, C) or (raw is not None and tuple(raw.shape) != (M, C)):
        raise ValueError("y / raw must be [B*T, C1+C2]")
    return M // B, C1, C2


def _groupnorm(x1, x2, B, gamma, beta, eps, silu, workspace, scale_shift, y, raw) -> None:
    T, C1, C2 = _gn_common(x1, x2, B, gamma, beta, scale_shift, y, raw)
    _chk(x1, "x1", f32)
    _chk(workspace, "workspace", f32, 1, dev=x1.device)
    L = _lib.lib()
    if workspace.numel() < L.seer_b200_groupnorm_workspace_floats(B, T):
        raise ValueError("groupnorm workspace too small (seer_b200_groupnorm_workspace_floats)")
    with _Dev(x1) as stream:
        rc = L.seer_b200_groupnorm(_p(x1), C1, _p(x2), C2, B, T, _p(gamma), _p(beta), float(eps), int(silu), _p(workspace),
                                   _p(scale_shift), _p(y), int(y.dtype == f32), _p(raw), stream)
    _lib.check(rc, f"groupnorm(B={B},T={T},C={C1}+{C2})")


def _groupnorm_from_stats(x1, stats1, x2, stats2, B, gamma, beta, eps, silu, scale_shift, y, raw) -> None:
    T, C1, C2 = _gn_common(x1, x2, B, gamma, beta, scale_shift, y, raw)
    M = B * T
    for nm, st, Ci in (("stats1", stats1, C1), ("stats2", stats2, C2)):
        if Ci:
            _chk(st, nm, f32, 3, contiguous=True, dev=x1.device)
            if tuple(st.shape) != (M // 32, Ci, 2):
                raise ValueError(f"{nm} must be [{M // 32}, {Ci}, 2]")
    if T % 32:
        raise ValueError("groupnorm_from_stats needs T % 32 == 0")
    with _Dev(x1) as stream:
        rc = _lib.lib().seer_b200_groupnorm_from_stats_ex(_p(x1), int(x1.dtype == bf16), C1, _p(stats1), _p(x2), C2,
                                                          _p(stats2) if x2 is not None else None, B, T, _p(gamma), _p(beta),
                                                          float(eps), int(silu), _p(scale_shift), _p(y), int(y.dtype == f32),
                                                          _p(raw), stream)
    _lib.check(rc, f"groupnorm_from_stats(B={B},T={T},C={C1}+{C2})")


def _layernorm(x, gamma, beta, eps, out) -> None:
    _chk(x, "x", f32, 2); _chk(out, "out", (bf16, f32), 2, dev=x.device)
    M, C = x.shape
    _chk(gamma, "gamma", f32, 1, dev=x.device); _chk(beta, "beta", f32, 1, dev=x.device)
    if gamma.numel() != C or beta.numel() != C or tuple(out.shape) != (M, C):
        raise ValueError("layernorm: gamma/beta [C], out [M, C]")
    fn = _lib.lib().seer_b200_layernorm if out.dtype == bf16 else _lib.lib().seer_b200_layernorm_f32
    with _Dev(x) as stream:
        rc = fn(_p(x), M, C, x.stride(0), _p(gamma), _p(beta), float(eps), _p(out), out.stride(0), stream)
    _lib.check(rc, f"layernorm(M={M},C={C})")


_define("groupnorm(Tensor x1, Tensor? x2, int B, Tensor gamma, Tensor beta, float eps, bool silu, Tensor(a!) workspace, "
        "Tensor(b!) scale_shift, Tensor(c!) y, Tensor(d!)? raw) -> ()", _groupnorm)
_define("groupnorm_from_stats(Tensor x1, Tensor stats1, Tensor? x2, Tensor? stats2, int B, Tensor gamma, Tensor beta, float eps, "
        "bool silu, Tensor(a!) scale_shift, Tensor(b!) y, Tensor(c!)? raw) -> ()", _groupnorm_from_stats)
_define("layernorm(Tensor x, Tensor gamma, Tensor beta, float eps, Tensor(a!) out) -> ()", _layernorm)


# --------------------------------------------------------------------------------------------------------------------
# attention
# --------------------------------------------------------------------------------------------------------------------
def _attention(q, k, v, out, mode, heads, n_outer, Lq, Lk, F, H, W) -> None:
    dt = q.dtype
    if dt not in (bf16, f32):
        raise TypeError(f"attention: expected bf16 or fp32 q/k/v, got {dt}")
    for n, t in (("q", q), ("k", k), ("v", v), ("out", out)):
        _chk(t, n, dt, 2, dev=q.device)
    C = q.shape[1]
    if C % heads or k.shape[1] != C or v.shape[1] != C or tuple(out.shape) != tuple(q.shape):
        raise ValueError("attention: q/k/v/out must be [rows, heads*d]")
    d = C // heads
    fn = _lib.lib().seer_b200_attention if dt == bf16 else _lib.lib().seer_b200_attention_f32
    with _Dev(q) as stream:
        rc = fn(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out), out.stride(0), mode, heads, d, n_outer,
                Lq, Lk, F, H, W, stream)
    _lib.check(rc, f"attention(mode={mode},heads={heads},d={d},outer={n_outer},Lq={Lq},Lk={Lk},F={F},H={H},W={W})")


def _scta_row_index(out, B, F, H, W) -> None:
    _chk(out, "out", torch.int32, 3, contiguous=True)
    nwin, L = ctypes.c_int(0), ctypes.c_int(0)
    lib = _lib.lib()
    _lib.check(lib.seer_b200_scta_row_index(B, F, H, W, None, ctypes.byref(nwin), ctypes.byref(L), None), "scta_row_index")
    if tuple(out.shape) != (B, nwin.value, L.value):
        raise ValueError(f"scta_row_index: out must be [{B}, {nwin.value}, {L.value}]")
    with _Dev(out) as stream:
        _lib.check(lib.seer_b200_scta_row_index(B, F, H, W, _p(out), ctypes.byref(nwin), ctypes.byref(L), stream), "scta_row_index")


_define("attention(Tensor q, Tensor k, Tensor v, Tensor(a!) out, int mode, int heads, int n_outer, int Lq, int Lk, int F, int H, "
        "int W) -> ()", _attention)
_define("scta_row_index(Tensor(a!) out, int B, int F, int H, int W) -> ()", _scta_row_index)


# --------------------------------------------------------------------------------------------------------------------
# elementwise / boundary convs / sampler update
# --------------------------------------------------------------------------------------------------------------------
def _rope(qk, pos_div, pos_mod, heads, head_dim, q_col, k_col, freqs) -> None:
    _chk(qk, "qk", (bf16, f32), 2); _chk(freqs, "freqs", f32, 1, dev=qk.device)
    with _Dev(qk) as stream:
        rc = _lib.lib().seer_b200_rope_ex(_p(qk), int(qk.dtype == f32), qk.stride(0), qk.shape[0], pos_div, pos_mod, heads, head_dim,
                                          q_col, k_col, _p(freqs), freqs.numel(), stream)
    _lib.check(rc, "rope")


def _timestep_embedding(t, out, shift, flip_sin_to_cos) -> None:
    _chk(t, "t", f32, 1); _chk(out, "out", f32, 2, contiguous=True, dev=t.device)
    if out.shape[0] != t.numel():
        raise ValueError("timestep_embedding: out must be [B, dim]")
    with _Dev(t) as stream:
        rc = _lib.lib().seer_b200_timestep_embedding(_p(t), _p(out), t.numel(), out.shape[1], float(shift), int(flip_sin_to_cos), stream)
    _lib.check(rc, "timestep_embedding")


def _small_linear(x, w, bias, add, out, silu_in, silu_out) -> None:
    _chk(x, "x", f32, 2); _chk(w, "w", f32, 2, contiguous=True, dev=x.device); _chk(out, "out", f32, 2, dev=x.device)
    _chk(bias, "bias", f32, 1, optional=True, dev=x.device); _chk(add, "add", f32, 1, optional=True, dev=x.device)
    B, K = x.shape
    N = w.shape[0]
    if w.shape[1] != K or tuple(out.shape) != (B, N):
        raise ValueError("small_linear: w [N, K], out [B, N]")
    with _Dev(x) as stream:
        rc = _lib.lib().seer_b200_small_linear(_p(x), x.stride(0), _p(w), _p(bias), _p(add), _p(out), out.stride(0), B, N, K,
                                               int(silu_in), int(silu_out), stream)
    _lib.check(rc, "small_linear")


def _conv_in(x, w, bias, out, col_stats) -> None:
    _chk(x, "x", f32, 5, contiguous=True)
    B, Cin, F, H, W = x.shape
    _chk(w, "w", f32, 2, contiguous=True, dev=x.device); _chk(bias, "bias", f32, 1, dev=x.device)
    Cout = w.shape[0]
    M = B * F * H * W
    _chk(out, "out", (f32, bf16), 2, contiguous=True, dev=x.device)
    _chk(col_stats, "col_stats", f32, 3, contiguous=True, optional=True, dev=x.device)
    if tuple(out.shape) != (M, Cout) or w.shape[1] != Cin * 9 or (col_stats is not None and tuple(col_stats.shape) != (M // 32, Cout, 2)):
        raise ValueError("conv_in: w [Cout, Cin*9], out [B*F*H*W, Cout], col_stats [M/32, Cout, 2]")
    with _Dev(x) as stream:
        rc = _lib.lib().seer_b200_conv_in_ex(_p(x), _p(w), _p(bias), _p(out), int(out.dtype == bf16), _p(col_stats), B, Cin, F, H, W,
                                             Cout, stream)
    _lib.check(rc, "conv_in")


def _conv_in_im2col(x, out) -> None:
    _chk(x, "x", f32, 5, contiguous=True)
    B, Cin, F, H, W = x.shape
    _chk(out, "out", bf16, 2, contiguous=True, dev=x.device)
    if Cin != 4 or tuple(out.shape) != (B * F * H * W, 64):
        raise ValueError("conv_in_im2col: x (B, 4, F, H, W), out [B*F*H*W, 64]")
    with _Dev(x) as stream:
        rc = _lib.lib().seer_b200_conv_in_im2col(_p(x), _p(out), B, Cin, F, H, W, stream)
    _lib.check(rc, "conv_in_im2col")


def _conv_out(x, w_packed, bias, out) -> None:
    _chk(x, "x", f32, 2, contiguous=True); _chk(out, "out", f32, 5, contiguous=True, dev=x.device)
    _chk(w_packed, "w_packed", f32, 3, contiguous=True, dev=x.device); _chk(bias, "bias", f32, 1, dev=x.device)
    B, Cout, F, H, W = out.shape
    Cin = x.shape[1]
    if x.shape[0] != B * F * H * W or tuple(w_packed.shape) != (Cout, 9, Cin):
        raise ValueError("conv_out: x [B*F*H*W, Cin], w_packed [Cout, 9, Cin], out (B, Cout, F, H, W)")
    with _Dev(x) as stream:
        rc = _lib.lib().seer_b200_conv_out(_p(x), _p(w_packed), _p(bias), _p(out), B, Cin, F, H, W, Cout, stream)
    _lib.check(rc, "conv_out")


def _upsample2x(x, out) -> None:
    _chk(x, "x", f32, 2, contiguous=True); _chk(out, "out", bf16, 4, contiguous=True, dev=x.device)
    n_img, H2, W2, C = out.shape
    if H2 % 2 or W2 % 2 or tuple(x.shape) != (n_img * (H2 // 2) * (W2 // 2), C):
        raise ValueError("upsample2x: x [n*H*W, C], out [n, 2H, 2W, C]")
    with _Dev(x) as stream:
        rc = _lib.lib().seer_b200_upsample2x_to_bf16(_p(x), _p(out), n_img, H2 // 2, W2 // 2, C, stream)
    _lib.check(rc, "upsample2x")


def _im2col3x3(x, out, stride) -> None:
    _chk(x, "x", (f32, bf16), 4, contiguous=True); _chk(out, "out", bf16, 2, contiguous=True, dev=x.device)
    n_img, H, W, C = x.shape
    if stride not in (1, 2) or tuple(out.shape) != (n_img * (H // stride) * (W // stride), 9 * C):
        raise ValueError("im2col3x3: out must be [n*(H/s)*(W/s), 9*C]")
    with _Dev(x) as stream:
        rc = _lib.lib().seer_b200_im2col3x3_to_bf16(_p(x), int(x.dtype == bf16), _p(out), n_img, H, W, C, stride, stream)
    _lib.check(rc, "im2col3x3")


def _cast_bf16(x, out) -> None:
    _chk(x, "x", f32, contiguous=True); _chk(out, "out", bf16, contiguous=True, dev=x.device)
    if out.numel() != x.numel():
        raise ValueError("cast_bf16: size mismatch")
    with _Dev(x) as stream:
        rc = _lib.lib().seer_b200_cast_f32_to_bf16(_p(x), _p(out), x.numel(), stream)
    _lib.check(rc, "cast_f32_to_bf16")


def _cfg_ddim_update(eps, x, x_prev, pred_x0, cond_f, use_cfg, scale, sqrt_one_minus_at, sqrt_at, sqrt_a_prev, dir_coef) -> None:
    _chk(eps, "eps", f32, 5, contiguous=True)
    for n, t in (("x", x), ("x_prev", x_prev), ("pred_x0", pred_x0)):
        _chk(t, n, f32, 5, contiguous=True, dev=eps.device)
    b, C, F2, H, W = x.shape
    if tuple(eps.shape) != ((2 * b if use_cfg else b), C, F2 + cond_f, H, W) or x_prev.shape != x.shape or pred_x0.shape != x.shape:
        raise ValueError(f"eps shape {tuple(eps.shape)} inconsistent with x {tuple(x.shape)}")
    with _Dev(eps) as stream:
        rc = _lib.lib().seer_b200_cfg_ddim_update(_p(eps), _p(x), _p(x_prev), _p(pred_x0), b, C, F2, cond_f, H * W, int(use_cfg),
                                                  float(scale), float(sqrt_one_minus_at), float(sqrt_at), float(sqrt_a_prev),
                                                  float(dir_coef), stream)
    _lib.check(rc, "cfg_ddim_update")


def _cfg_ddim_update_p2p(eps_local, branch, peer_recv, local_recv, peer_flag, local_flag, counter, seq, x, x_prev, pred_x0, cond_f,
                         scale, sqrt_one_minus_at, sqrt_at, sqrt_a_prev, dir_coef) -> None:
    _chk(eps_local, "eps_local", f32, 5, contiguous=True)
    for n, t in (("x", x), ("x_prev", x_prev), ("pred_x0", pred_x0)):
        _chk(t, n, f32, 5, contiguous=True, dev=eps_local.device)
    b, C, F2, H, W = x.shape
    if tuple(eps_local.shape) != (b, C, F2 + cond_f, H, W) or x_prev.shape != x.shape or pred_x0.shape != x.shape:
        raise ValueError(f"eps_local shape {tuple(eps_local.shape)} inconsistent with x {tuple(x.shape)}")
    # peer_recv / peer_flag live on the PARTNER's device (mapped here through symmetric memory): no device check on them
    for n, t, dt in (("peer_recv", peer_recv, f32), ("local_recv", local_recv, f32)):
        if not (t.is_cuda and t.dtype == dt and t.is_contiguous() and t.numel() >= x.numel()):
            raise ValueError(f"{n}: contiguous CUDA {dt} tensor of at least {x.numel()} elements expected")
    for n, t in (("peer_flag", peer_flag), ("local_flag", local_flag), ("counter", counter)):
        if not (t.is_cuda and t.dtype == torch.int32 and t.numel() >= 1):
            raise ValueError(f"{n}: CUDA int32 word expected")
    if branch not in (0, 1) or not (0 < seq < 2 ** 31):
        raise ValueError("branch must be 0 or 1 and 0 < seq < 2**31")
    with _Dev(eps_local) as stream:
        rc = _lib.lib().seer_b200_cfg_ddim_update_p2p(_p(eps_local), int(branch), _p(peer_recv), _p(local_recv), _p(peer_flag),
                                                      _p(local_flag), _p(counter), int(seq), _p(x), _p(x_prev), _p(pred_x0), b, C, F2,
                                                      cond_f, H * W, float(scale), float(sqrt_one_minus_at), float(sqrt_at),
                                                      float(sqrt_a_prev), float(dir_coef), stream)
    _lib.check(rc, "cfg_ddim_update_p2p")


def _split3(x, out, ctot, col0, up_n_img, up_H, up_W) -> None:
    _chk(x, "x", f32, 2); _chk(out, "out", bf16, 2, dev=x.device)
    M, C = x.shape
    rows_out = 4 * M if up_H else M
    if out.shape[0] != rows_out or out.shape[1] != 3 * ctot:
        raise ValueError(f"split3: out shape {tuple(out.shape)} != ({rows_out}, {3 * ctot})")
    with _Dev(x) as stream:
        rc = _lib.lib().seer_b200_split3_bf16(_p(x), x.stride(0), M, C, _p(out), out.stride(0), ctot, col0, up_n_img, up_H, up_W, stream)
    _lib.check(rc, f"split3(M={M},C={C})")


def _geglu_f32(h, out) -> None:
    _chk(h, "h", f32, 2); _chk(out, "out", f32, 2, dev=h.device)
    M, two_i = h.shape
    if two_i % 2 or tuple(out.shape) != (M, two_i // 2):
        raise ValueError("geglu_f32: out must be [M, inner]")
    with _Dev(h) as stream:
        rc = _lib.lib().seer_b200_geglu_f32(_p(h), h.stride(0), _p(out), out.stride(0), M, two_i // 2, stream)
    _lib.check(rc, "geglu_f32")


def _rope_table(freqs, out) -> None:
    _chk(freqs, "freqs", f32, 1); _chk(out, "out", torch.float16, 3, contiguous=True, dev=freqs.device)
    if out.shape[1] != freqs.numel() or out.shape[2] != 2:
        raise ValueError("rope_table: out must be [T, n_freqs, 2]")
    with _Dev(freqs) as stream:
        rc = _lib.lib().seer_b200_rope_table(_p(freqs), freqs.numel(), out.shape[0], _p(out), stream)
    _lib.check(rc, "rope_table")


def _softmax_rows(S, scale, out) -> None:
    _chk(S, "S", f32, 2); _chk(out, "out", bf16, 2, dev=S.device)
    if tuple(out.shape) != tuple(S.shape):
        raise ValueError("softmax_rows: out must have the shape of S")
    with _Dev(S) as stream:
        rc = _lib.lib().seer_b200_softmax_rows(_p(S), S.stride(0), S.shape[0], S.shape[1], float(scale), _p(out), out.stride(0), stream)
    _lib.check(rc, "softmax_rows")


def _tokens_to_nchw(x, out) -> None:
    _chk(x, "x", f32, 2); _chk(out, "out", f32, 5, contiguous=True, dev=x.device)
    B, C, F, H, W = out.shape
    if x.shape[0] != B * F * H * W or x.shape[1] < C:
        raise ValueError("tokens_to_nchw: x must be [B*F*H*W, >= C]")
    with _Dev(x) as stream:
        rc = _lib.lib().seer_b200_tokens_to_nchw(_p(x), x.stride(0), _p(out), B, C, F, H * W, stream)
    _lib.check(rc, "tokens_to_nchw")


_define("tokens_to_nchw(Tensor x, Tensor(a!) out) -> ()", _tokens_to_nchw)
_define("softmax_rows(Tensor S, float scale, Tensor(a!) out) -> ()", _softmax_rows)
def _rope_apply_table(qk, tokens_per_clip, heads, head_dim, q_col, k_col, tab) -> None:
    _chk(qk, "qk", bf16, 2); _chk(tab, "tab", torch.float16, 3, contiguous=True, dev=qk.device)
    if tuple(tab.shape) != (tokens_per_clip, 16, 2) or qk.shape[0] % tokens_per_clip:
        raise ValueError("rope_apply_table: tab must be [tokens_per_clip, 16, 2] and rows a multiple of tokens_per_clip")
    with _Dev(qk) as stream:
        rc = _lib.lib().seer_b200_rope_apply_table(_p(qk), qk.stride(0), qk.shape[0], tokens_per_clip, heads, head_dim, q_col, k_col,
                                                   _p(tab), stream)
    _lib.check(rc, "rope_apply_table")


def _rope_inplace(qk, tokens_per_clip, heads, head_dim, q_col, k_col, freqs) -> None:
    _chk(qk, "qk", bf16, 2); _chk(freqs, "freqs", f32, 1, dev=qk.device)
    if qk.shape[0] % tokens_per_clip:
        raise ValueError("rope_inplace: rows must be a multiple of tokens_per_clip")
    with _Dev(qk) as stream:
        rc = _lib.lib().seer_b200_rope_inplace(_p(qk), qk.stride(0), qk.shape[0], tokens_per_clip, heads, head_dim, q_col, k_col,
                                               _p(freqs), freqs.numel(), stream)
    _lib.check(rc, "rope_inplace")


_define("rope_apply_table(Tensor(a!) qk, int tokens_per_clip, int heads, int head_dim, int q_col, int k_col, Tensor tab) -> ()", _rope_apply_table)
_define("rope_inplace(Tensor(a!) qk, int tokens_per_clip, int heads, int head_dim, int q_col, int k_col, Tensor freqs) -> ()", _rope_inplace)
_define("rope_table(Tensor freqs, Tensor(a!) out) -> ()", _rope_table)
_define("rope(Tensor(a!) qk, int pos_div, int pos_mod, int heads, int head_dim, int q_col, int k_col, Tensor freqs) -> ()", _rope)
_define("timestep_embedding(Tensor t, Tensor(a!) out, float shift, bool flip_sin_to_cos) -> ()", _timestep_embedding)
_define("small_linear(Tensor x, Tensor w, Tensor? bias, Tensor? add, Tensor(a!) out, bool silu_in, bool silu_out) -> ()", _small_linear)
_define("conv_in(Tensor x, Tensor w, Tensor bias, Tensor(a!) out, Tensor(b!)? col_stats) -> ()", _conv_in)
_define("conv_in_im2col(Tensor x, Tensor(a!) out) -> ()", _conv_in_im2col)
_define("conv_out(Tensor x, Tensor w_packed, Tensor bias, Tensor(a!) out) -> ()", _conv_out)
_define("upsample2x(Tensor x, Tensor(a!) out) -> ()", _upsample2x)
_define("im2col3x3(Tensor x, Tensor(a!) out, int stride) -> ()", _im2col3x3)
_define("cast_bf16(Tensor x, Tensor(a!) out) -> ()", _cast_bf16)
_define("cfg_ddim_update(Tensor eps, Tensor x, Tensor(a!) x_prev, Tensor(b!) pred_x0, int cond_f, bool use_cfg, float scale, "
        "float sqrt_one_minus_at, float sqrt_at, float sqrt_a_prev, float dir_coef) -> ()", _cfg_ddim_update)
_define("cfg_ddim_update_p2p(Tensor eps_local, int branch, Tensor(a!) peer_recv, Tensor local_recv, Tensor(b!) peer_flag, "
        "Tensor local_flag, Tensor(c!) counter, int seq, Tensor x, Tensor(d!) x_prev, Tensor(e!) pred_x0, int cond_f, float scale, "
        "float sqrt_one_minus_at, float sqrt_at, float sqrt_a_prev, float dir_coef) -> ()", _cfg_ddim_update_p2p)
_define("split3(Tensor x, Tensor(a!) out, int ctot, int col0, int up_n_img, int up_H, int up_W) -> ()", _split3)
_define("geglu_f32(Tensor h, Tensor(a!) out) -> ()", _geglu_f32)
