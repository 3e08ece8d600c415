"""ctypes loader for libseer_b200.so (the C-ABI kernel library, include/seer_b200.h).

The product path has NO CPU or PyTorch fallback: if the shared library is missing or a symbol is
absent, importing the ops raises.  `build()` compiles it in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SEER_B200_LIB") or os.path.join(_HERE, "libseer_b200.so")     # SEER_B200_LIB: an A/B build of the kernels
BUILD_SCRIPT = os.path.join(_HERE, "csrc", "build.sh")

_c = ctypes
_vp, _i, _f, _ll = _c.c_void_p, _c.c_int, _c.c_float, _c.c_longlong
_ip = _c.POINTER(_c.c_int)
_u32 = _c.c_uint32



class GemmDesc(ctypes.Structure):
    """Mirror of `SeerGemmDesc` (include/seer_b200.h) — field order and types must match the C struct."""
    _fields_ = [
        ("A", _vp), ("lda", _i), ("K1", _i),
        ("X", _vp), ("n_img", _i), ("H", _i), ("W", _i), ("Cin", _i),
        ("A2", _vp), ("lda2", _i), ("K2", _i),
        ("Wt", _vp), ("M", _i), ("N", _i),
        ("bias", _vp), ("ldb", _i), ("bias_div", _i),
        ("residual", _vp), ("ldr", _i), ("residual_bf16", _i),
        ("out_f32", _vp), ("ldo_f32", _i),
        ("out_bf16", _vp), ("ldo_bf16", _i),
        ("geglu", _i),
        ("col_stats", _vp),
        ("row_stats_out", _vp),
        ("row_stats_in", _vp), ("row_parts_in", _i), ("ln_eps", _f), ("ln_colsum", _vp),
        ("conv_stride", _i), ("conv_taps_w", _i), ("conv_taps_h", _i), ("conv_off_x", _i), ("conv_off_y", _i),
        ("out_up_phase", _i),
        ("rope_tab", _vp), ("rope_T", _i), ("rope_cols", _i), ("rope_d", _i),
    ]


_dp = _c.POINTER(GemmDesc)

# name -> (restype, argtypes): must list every symbol include/seer_b200.h declares
SIGNATURES = {
    "seer_b200_gemm_ex": (_i, [_dp, _vp]),
    "seer_b200_gemm_row_parts": (_i, [_dp]),
    "seer_b200_gemm_desc_size": (_i, []),
    "seer_b200_groupnorm_from_stats": (_i, [_vp, _i, _vp, _vp, _i, _vp, _i, _i, _vp, _vp, _f, _i, _vp, _vp, _i, _vp, _vp]),
    "seer_b200_groupnorm_from_stats_ex": (_i, [_vp, _i, _i, _vp, _vp, _i, _vp, _i, _i, _vp, _vp, _f, _i, _vp, _vp, _i, _vp, _vp]),
    "seer_b200_version": (_c.c_char_p, []),
    "seer_b200_debug_last_attention": (_c.c_char_p, []),
    "seer_b200_debug_last_gemm": (_c.c_char_p, []),
    "seer_b200_debug_setenv": (None, [_c.c_char_p, _i, _i]),
    "seer_b200_gemm_bf16": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp, _i, _vp, _i, _i, _vp]),
    "seer_b200_conv3x3_bf16": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _i, _vp, _i, _i, _vp, _i, _vp, _i, _i, _vp]),
    "seer_b200_groupnorm_workspace_floats": (_i, [_i, _i]),
    "seer_b200_groupnorm": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _f, _i, _vp, _vp, _vp, _i, _vp, _vp]),
    "seer_b200_layernorm": (_i, [_vp, _i, _i, _i, _vp, _vp, _f, _vp, _i, _vp]),
    "seer_b200_attention": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "seer_b200_scta_row_index": (_i, [_i, _i, _i, _i, _vp, _ip, _ip, _vp]),
    "seer_b200_rope_inplace": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "seer_b200_rope_table": (_i, [_vp, _i, _i, _vp, _vp]),
    "seer_b200_rope_apply_table": (_i, [_vp, _i, _ll, _i, _i, _i, _i, _i, _vp, _vp]),
    "seer_b200_timestep_embedding": (_i, [_vp, _vp, _i, _i, _f, _i, _vp]),
    "seer_b200_small_linear": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "seer_b200_conv_in": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "seer_b200_conv_in_stats": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "seer_b200_conv_in_ex": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "seer_b200_conv_in_im2col": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "seer_b200_conv_out": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "seer_b200_tokens_to_nchw": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp]),
    "seer_b200_softmax_rows": (_i, [_vp, _i, _ll, _i, _f, _vp, _i, _vp]),
    "seer_b200_upsample2x_to_bf16": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "seer_b200_im2col3x3_to_bf16": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp]),
    "seer_b200_cast_f32_to_bf16": (_i, [_vp, _vp, _ll, _vp]),
    "seer_b200_cfg_ddim_update": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _f, _f, _f, _f, _vp]),
    "seer_b200_cfg_ddim_update_p2p": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _u32, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _f, _f, _f, _vp]),
    # fp32-parity path
    "seer_b200_split3_bf16": (_i, [_vp, _i, _ll, _i, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "seer_b200_layernorm_f32": (_i, [_vp, _i, _i, _i, _vp, _vp, _f, _vp, _i, _vp]),
    "seer_b200_geglu_f32": (_i, [_vp, _i, _vp, _i, _ll, _i, _vp]),
    "seer_b200_rope_ex": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "seer_b200_attention_f32": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
}

_lib = None


class SeerB200Error(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into seervideoldm_b200/libseer_b200.so."""
    r = subprocess.run(["bash", BUILD_SCRIPT], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise SeerB200Error(f"nvcc build failed (exit {r.returncode})")
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SeerB200Error(
                f"{LIB_PATH} not found — run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU/PyTorch fallback for the seer_b200 kernels)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)       # AttributeError if the symbol is missing: fail loudly
            fn.restype = res
            fn.argtypes = args
        if handle.seer_b200_gemm_desc_size() != ctypes.sizeof(GemmDesc):
            raise SeerB200Error("SeerGemmDesc layout mismatch between libseer_b200.so and _lib.GemmDesc — rebuild the library")
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc > 0:
        raise SeerB200Error(f"{what}: CUDA error {rc}")
    names = {-1: "invalid argument", -2: "unsupported shape", -3: "CUDA driver entry point unavailable"}
    exc = ValueError if rc in (-1, -2) else SeerB200Error
    raise exc(f"{what}: {names.get(rc, rc)}")
