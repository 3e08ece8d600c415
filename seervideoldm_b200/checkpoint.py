"""Checkpoint / wire formats on either side of the hot path.

What the reference's scripts load, and what this module reads without diffusers / accelerate / safetensors installed:

  * `SeerUNet.from_pretrained(sd15_dir, subfolder="unet", ...)`          (/root/reference/inference.py:82-87)
        diffusers 0.10.2 `ModelMixin.from_pretrained`: `config.json` + `diffusion_pytorch_model.bin`
        (newer hubs: `.safetensors`, `.fp16.` variants).  The SD-1.5 file is a 2-D UNet: its 686 entries are the
        conv / spatial-transformer part of the 1006-entry Seer schema with identical names and shapes
        (`InflatedConv3d` is a `Conv2d` subclass, unet_3d_condition.py:41-49); the 320 `temporal_attentions.*` entries
        are absent and keep their constructor initialisation, unexpected keys are reported, a shape mismatch raises —
        the behaviour of diffusers' loader (missing/unexpected keys warn, mismatched sizes raise).
  * `torch.load(".../pytorch_model.bin")`, `".../pytorch_model_1.bin"`    (inference.py:123-127)
        accelerate's `save_state` files of the fine-tuned SeerUNet / FSTextTransformer, loaded with strict=True.

Tensors are upcast to fp32 on load (parameters stay fp32); packing into the kernels' bf16 layouts happens
lazily on the first forward (`SeerUNet._pack`).  The safetensors reader/writer below is a restatement of the published
format (8-byte little-endian header length, JSON header {name: {dtype, shape, data_offsets}}, raw little-endian data);
`tests/test_checkpoint_cpu.py` cross-checks it against the `safetensors` package when that is importable.
"""
from __future__ import annotations

import inspect
import json
import os
import struct
import warnings
from collections import OrderedDict
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch

_ST_DTYPES = {
    "F64": torch.float64, "F32": torch.float32, "F16": torch.float16, "BF16": torch.bfloat16,
    "I64": torch.int64, "I32": torch.int32, "I16": torch.int16, "I8": torch.int8, "U8": torch.uint8, "BOOL": torch.bool,
}
_ST_NAMES = {v: k for k, v in _ST_DTYPES.items()}
_MAX_HEADER = 100 * 1024 * 1024          # the format's own limit

UNET_WEIGHT_NAMES = ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.bin",
                     "diffusion_pytorch_model.fp16.safetensors", "diffusion_pytorch_model.fp16.bin")
ACCELERATE_UNET, ACCELERATE_FSTEXT = "pytorch_model.bin", "pytorch_model_1.bin"      # inference.py:123,126


# ---------------------------------------------------------------------------------------------------------------------
# safetensors
# ---------------------------------------------------------------------------------------------------------------------
def read_safetensors(path: str) -> "OrderedDict[str, torch.Tensor]":
    """Read every tensor of a .safetensors file (returned in file-offset order, detached from the file)."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        head = f.read(8)
        if len(head) != 8:
            raise ValueError(f"{path}: not a safetensors file (shorter than its 8-byte header length)")
        (n,) = struct.unpack("<Q", head)
        if n > _MAX_HEADER or 8 + n > size:
            raise ValueError(f"{path}: invalid safetensors header length {n}")
        try:
            header = json.loads(f.read(n).decode("utf-8"))
        except (UnicodeDecodeError, json.JSONDecodeError) as e:
            raise ValueError(f"{path}: safetensors header is not valid JSON: {e}") from None
        data = np.fromfile(f, dtype=np.uint8)
    header.pop("__metadata__", None)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, info in sorted(header.items(), key=lambda kv: kv[1]["data_offsets"][0]):
        if info["dtype"] not in _ST_DTYPES:
            raise ValueError(f"{path}: tensor {name!r} has unsupported dtype {info['dtype']}")
        dt = _ST_DTYPES[info["dtype"]]
        shape = tuple(int(s) for s in info["shape"])
        lo, hi = (int(v) for v in info["data_offsets"])
        want = int(np.prod(shape, dtype=np.int64)) * torch.empty((), dtype=dt).element_size()
        if not (0 <= lo <= hi <= data.size) or hi - lo != want:
            raise ValueError(f"{path}: tensor {name!r} spans bytes [{lo}, {hi}) but its shape {shape} needs {want}")
        raw = torch.from_numpy(data[lo:hi].copy())
        out[name] = raw.view(torch.uint8).view(dt).reshape(shape) if want else torch.empty(shape, dtype=dt)
    return out


def write_safetensors(path: str, tensors: Dict[str, torch.Tensor], metadata: Optional[Dict[str, str]] = None) -> None:
    """Write a .safetensors file (names sorted, offsets contiguous, header padded with spaces to 8 bytes)."""
    header: dict = {}
    if metadata:
        header["__metadata__"] = {str(k): str(v) for k, v in metadata.items()}
    blobs: List[bytes] = []
    off = 0
    for name in sorted(tensors):
        t = tensors[name].detach().cpu().contiguous()
        if t.dtype not in _ST_NAMES:
            raise ValueError(f"tensor {name!r}: dtype {t.dtype} has no safetensors encoding")
        b = t.reshape(-1).view(torch.uint8).numpy().tobytes() if t.numel() else b""
        header[name] = {"dtype": _ST_NAMES[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + len(b)]}
        blobs.append(b)
        off += len(b)
    hj = json.dumps(header, separators=(",", ":")).encode("utf-8")
    hj += b" " * (-len(hj) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for b in blobs:
            f.write(b)


# ---------------------------------------------------------------------------------------------------------------------
# state dicts
# ---------------------------------------------------------------------------------------------------------------------
def read_state_dict(path: str) -> "OrderedDict[str, torch.Tensor]":
    """A flat {key: CPU tensor} dict from `.safetensors`, or a `torch.save`d `.bin` / `.pt` / `.ckpt` (optionally
    wrapped as {"state_dict": ...}, as LDM checkpoints are); a leading `module.` (DDP) prefix is stripped."""
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    if path.endswith(".safetensors"):
        sd = read_safetensors(path)
    else:
        sd = torch.load(path, map_location="cpu", weights_only=True)
        if isinstance(sd, dict) and "state_dict" in sd and isinstance(sd["state_dict"], dict):
            sd = sd["state_dict"]
        if not isinstance(sd, dict) or not all(isinstance(v, torch.Tensor) for v in sd.values()):
            raise ValueError(f"{path}: expected a flat state dict of tensors")
    if sd and all(k.startswith("module.") for k in sd):
        sd = OrderedDict((k[len("module."):], v) for k, v in sd.items())
    return OrderedDict(sd)


def _to_param_dtype(t: torch.Tensor) -> torch.Tensor:
    return t.float() if t.is_floating_point() else t


def load_into(module: torch.nn.Module, state_dict: Dict[str, torch.Tensor], strict: bool = True,
              allow_missing_prefixes: Iterable[str] = ()) -> Tuple[List[str], List[str]]:
    """`module.load_state_dict` with the loader semantics the reference relies on.

    strict=True  — accelerate checkpoints (inference.py:124,127): every key must be present, none extra.
    strict=False — diffusers' `from_pretrained`: missing keys keep their initialisation and unexpected keys are
                   dropped, both returned; a key whose shape differs ALWAYS raises (diffusers raises too unless
                   `ignore_mismatched_sizes`), and so does a missing key outside `allow_missing_prefixes` when that
                   is given — a truncated file must not load silently.
    Floating-point tensors are upcast to fp32 (fp16 / bf16 hub variants)."""
    own = module.state_dict()
    bad = [f"{k}: checkpoint {tuple(v.shape)} vs model {tuple(own[k].shape)}"
           for k, v in state_dict.items() if k in own and tuple(v.shape) != tuple(own[k].shape)]
    if bad:
        raise RuntimeError("size mismatch for " + "; ".join(bad[:8]) + (f" (+{len(bad) - 8} more)" if len(bad) > 8 else ""))
    sd = OrderedDict((k, _to_param_dtype(v)) for k, v in state_dict.items())
    if strict:
        module.load_state_dict(sd, strict=True)
        return [], []
    missing = [k for k in own if k not in sd]
    unexpected = [k for k in sd if k not in own]
    allow = tuple(allow_missing_prefixes)
    if allow:
        stray = [k for k in missing if not any(p in k for p in allow)]
        if stray:
            raise RuntimeError(f"checkpoint lacks {len(stray)} non-temporal keys, e.g. {stray[:4]} — truncated file?")
    module.load_state_dict(OrderedDict((k, v) for k, v in sd.items() if k in own), strict=False)
    return missing, unexpected


# ---------------------------------------------------------------------------------------------------------------------
# diffusers-style directories
# ---------------------------------------------------------------------------------------------------------------------
def _model_dir(path: str, subfolder: Optional[str]) -> str:
    d = os.path.join(path, subfolder) if subfolder else path
    if not os.path.isdir(d):
        raise OSError(f"{d} is not a directory — seer_b200 loads local checkpoints only (there is no hub access)")
    return d


def find_weights(model_dir: str) -> str:
    for name in UNET_WEIGHT_NAMES:
        p = os.path.join(model_dir, name)
        if os.path.isfile(p):
            return p
    raise OSError(f"no weights file in {model_dir} (looked for {', '.join(UNET_WEIGHT_NAMES)})")


def unet_from_pretrained(cls, pretrained_model_name_or_path: str, subfolder: Optional[str] = None, **kwargs):
    """Body of `SeerUNet.from_pretrained` (diffusers 0.10.2 `ModelMixin.from_pretrained` as the reference calls it,
    inference.py:82-87).  Hub-only arguments (`revision`, `low_cpu_mem_usage`, `cache_dir`, `torch_dtype`, ...) are
    accepted and ignored; `output_loading_info=True` returns `(model, info)` like diffusers."""
    output_loading_info = bool(kwargs.pop("output_loading_info", False))
    d = _model_dir(pretrained_model_name_or_path, subfolder)
    cfg_path = os.path.join(d, "config.json")
    if not os.path.isfile(cfg_path):
        raise OSError(f"{cfg_path} not found")
    with open(cfg_path, "r", encoding="utf-8") as f:
        raw = json.load(f)
    accepted = set(inspect.signature(cls.__init__).parameters) - {"self"}
    # `_class_name` etc. and arguments of newer UNet2DConditionModel versions the reference's class does not have are
    # dropped, exactly what diffusers' `extract_init_dict` does for a class with a narrower signature
    init = {k: (tuple(v) if isinstance(v, list) else v) for k, v in raw.items() if k in accepted}
    model = cls(**init)
    missing, unexpected = load_into(model, read_state_dict(find_weights(d)), strict=False,
                                    allow_missing_prefixes=("temporal_attentions.",))
    if missing:
        warnings.warn(f"{len(missing)} temporal-attention entries are not in the checkpoint and keep their initialisation "
                      f"(a 2-D Stable Diffusion UNet was inflated); load the fine-tuned pytorch_model.bin next", stacklevel=3)
    if unexpected:
        warnings.warn(f"{len(unexpected)} checkpoint entries were not used, e.g. {unexpected[:4]}", stacklevel=3)
    model.eval()
    if output_loading_info:
        return model, {"missing_keys": missing, "unexpected_keys": unexpected, "mismatched_keys": [], "error_msgs": []}
    return model


def unet_save_pretrained(model, save_directory: str, safe_serialization: bool = False) -> str:
    """`config.json` + `diffusion_pytorch_model.{bin,safetensors}` in diffusers' layout; returns the weights path."""
    os.makedirs(save_directory, exist_ok=True)
    cfg = {"_class_name": type(model).__name__, "_diffusers_version": "0.10.2"}
    for k, v in vars(model.config).items():
        cfg[k] = list(v) if isinstance(v, tuple) else v
    with open(os.path.join(save_directory, "config.json"), "w", encoding="utf-8") as f:
        json.dump(cfg, f, indent=2, sort_keys=True)
    sd = OrderedDict((k, v.detach().cpu()) for k, v in model.state_dict().items())
    if safe_serialization:
        path = os.path.join(save_directory, UNET_WEIGHT_NAMES[0])
        write_safetensors(path, sd, metadata={"format": "pt"})
    else:
        path = os.path.join(save_directory, UNET_WEIGHT_NAMES[1])
        torch.save(sd, path)
    return path


def load_seer_checkpoint(load_path: str, sunet=None, fstext_model=None) -> None:
    """The reference's checkpoint restore (inference.py:119-128): `pytorch_model.bin` -> SeerUNet and
    `pytorch_model_1.bin` -> FSTextTransformer, both strict."""
    if sunet is not None:
        load_into(sunet, read_state_dict(os.path.join(load_path, ACCELERATE_UNET)), strict=True)
    if fstext_model is not None:
        load_into(fstext_model, read_state_dict(os.path.join(load_path, ACCELERATE_FSTEXT)), strict=True)
