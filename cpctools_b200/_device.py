"""Host<->device plumbing: torch is used for device memory, streams and pinned staging only.

Inputs may be ``numpy.ndarray`` (any contiguity) or CUDA ``torch.Tensor``; results come back in
kind (numpy in -> numpy out, CUDA tensor in -> CUDA tensor out), like SURVEY.md section 8(b) asks.
There is deliberately NO CPU compute path: without a CUDA device every entry point raises.
"""
from __future__ import annotations

import numpy as np

from . import _lib

try:  # torch is a hard requirement of the compute path, but the host-only helpers import without it
    import torch
except Exception:  # pragma: no cover
    torch = None


def require_cuda():
    if torch is None:
        raise _lib.SoapB200Error("PyTorch is required for device memory (cpctools_b200 has no CPU fallback)")
    if not torch.cuda.is_available():
        raise _lib.SoapB200Error("no CUDA device: cpctools_b200 runs on B200 (sm_100a) only, there is no CPU fallback")
    _lib.load()
    return torch


def is_tensor(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


def stream_ptr() -> int:
    return int(torch.cuda.current_stream().cuda_stream)


def dtype_code(t) -> int:
    if t.dtype == torch.float64:
        return _lib.F64
    if t.dtype == torch.float32:
        return _lib.F32
    raise TypeError(f"unsupported floating dtype {t.dtype}")


def to_device(x, dtype=None):
    """numpy / tensor -> contiguous CUDA tensor (H2D through pinned memory for numpy)."""
    th = require_cuda()
    if is_tensor(x):
        t = x if x.is_cuda else x.cuda(non_blocking=True)
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.contiguous()
    arr = np.asarray(x)
    if arr.dtype == np.float16:
        arr = arr.astype(np.float32)
    if not arr.flags.c_contiguous or not arr.flags.aligned or arr.dtype.byteorder not in "=|<":
        arr = np.ascontiguousarray(arr.astype(arr.dtype.newbyteorder("="), copy=False))
    if not arr.flags.writeable:
        arr = arr.copy()
    host = th.from_numpy(arr)
    if host.numel() >= (1 << 16):
        pinned = th.empty(host.shape, dtype=host.dtype, pin_memory=True)
        pinned.copy_(host)
        t = pinned.to("cuda", non_blocking=True)
    else:
        t = host.to("cuda")
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t


def to_host(t) -> np.ndarray:
    return t.detach().cpu().numpy()


def like_input(result, original):
    """Return ``result`` (CUDA tensor) in the container type of ``original``: CUDA tensor -> CUDA
    tensor (stays on the device), host tensor -> host tensor, anything else -> numpy."""
    if is_tensor(original):
        return result if original.is_cuda else result.cpu()
    return to_host(result)


def as_float_tensor(x):
    """Device tensor in float32/float64; integers are promoted to float64 like numpy's true divide."""
    t = to_device(x)
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    return t
