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


class _HostStager:
    """Pageable host memory -> HBM through two fixed pinned buffers: a piece is copied into a pinned buffer
    by a few host threads (numpy releases the GIL) while the previous piece is on its way over the link.
    Pinning the caller's whole array instead costs seconds for a 12 GB trajectory (``cudaHostAlloc``)."""

    CHUNK = 256 << 20
    MIN_BYTES = 64 << 20

    def __init__(self):
        import os
        import threading

        self.lock = threading.Lock()
        self.bufs = [None, None]
        self.events = [None, None]
        self.stream = None
        self.pool = None
        self.threads = max(1, min(8, (os.cpu_count() or 2) // 2))

    def upload(self, arr: np.ndarray):
        th = torch
        flat = arr.reshape(-1).view(np.uint8)
        n = int(flat.size)
        with self.lock:
            if self.stream is None:
                from concurrent.futures import ThreadPoolExecutor

                self.stream = th.cuda.Stream()
                self.pool = ThreadPoolExecutor(self.threads)
            main = th.cuda.current_stream()
            dev = th.empty(n, dtype=th.uint8, device="cuda")
            self.stream.wait_stream(main)  # the block may be recycled memory with work pending on `main`
            k = 0
            for off in range(0, n, self.CHUNK):
                m = min(self.CHUNK, n - off)
                if self.bufs[k] is None:
                    self.bufs[k] = th.empty(self.CHUNK, dtype=th.uint8, pin_memory=True)
                if self.events[k] is not None:
                    self.events[k].synchronize()  # the copy that last read this pinned buffer
                dst = self.bufs[k].numpy()
                cuts = [m * i // self.threads for i in range(self.threads + 1)]
                list(self.pool.map(lambda ab: np.copyto(dst[ab[0] : ab[1]], flat[off + ab[0] : off + ab[1]]),
                                   zip(cuts[:-1], cuts[1:])))
                with th.cuda.stream(self.stream):
                    dev[off : off + m].copy_(self.bufs[k][:m], non_blocking=True)
                    ev = th.cuda.Event()
                    ev.record(self.stream)
                self.events[k] = ev
                k ^= 1
            main.wait_stream(self.stream)
            dev.record_stream(self.stream)
        return dev


    def upload_rows(self, arr: np.ndarray, lo: int, hi: int):
        """Frames ``f`` of ``arr[f, lo:hi, :]`` (each a contiguous run of the host array) -> one flat device buffer."""
        th = torch
        F, A, D = arr.shape
        row = (hi - lo) * D * arr.dtype.itemsize
        n = F * row
        src = arr.reshape(F, A * D).view(np.uint8)[:, lo * D * arr.dtype.itemsize : hi * D * arr.dtype.itemsize]
        per = max(1, self.CHUNK // max(1, row))  # frames per pinned chunk
        with self.lock:
            if self.stream is None:
                from concurrent.futures import ThreadPoolExecutor

                self.stream = th.cuda.Stream()
                self.pool = ThreadPoolExecutor(self.threads)
            main = th.cuda.current_stream()
            dev = th.empty(n, dtype=th.uint8, device="cuda")
            self.stream.wait_stream(main)
            k = 0
            for f0 in range(0, F, per):
                f1 = min(F, f0 + per)
                m = (f1 - f0) * row
                need = max(self.CHUNK, m)
                if self.bufs[k] is None or self.bufs[k].numel() < need:
                    if self.events[k] is not None:
                        self.events[k].synchronize()
                    self.bufs[k] = th.empty(need, dtype=th.uint8, pin_memory=True)
                if self.events[k] is not None:
                    self.events[k].synchronize()
                dst = self.bufs[k].numpy()[:m].reshape(f1 - f0, row)
                cuts = [f0 + (f1 - f0) * i // self.threads for i in range(self.threads + 1)]
                list(self.pool.map(lambda ab: np.copyto(dst[ab[0] - f0 : ab[1] - f0], src[ab[0] : ab[1]]), zip(cuts[:-1], cuts[1:])))
                with th.cuda.stream(self.stream):
                    dev[f0 * row : f0 * row + m].copy_(self.bufs[k][:m], non_blocking=True)
                    ev = th.cuda.Event()
                    ev.record(self.stream)
                self.events[k] = ev
                k ^= 1
            main.wait_stream(self.stream)
            dev.record_stream(self.stream)
        return dev

    def download(self, t) -> np.ndarray:
        """HBM -> fresh pageable numpy array, the mirror image of :meth:`upload` (a plain ``.cpu()`` into
        untouched pageable memory runs at a fraction of the link rate: one thread takes all the page faults)."""
        th = torch
        t = t.detach().contiguous()
        out = np.empty(tuple(t.shape), dtype=th.empty(0, dtype=t.dtype).numpy().dtype)
        flat = out.reshape(-1).view(np.uint8)
        src = t.reshape(-1).view(th.uint8)
        n = int(flat.size)
        with self.lock:
            if self.stream is None:
                from concurrent.futures import ThreadPoolExecutor

                self.stream = th.cuda.Stream()
                self.pool = ThreadPoolExecutor(self.threads)
            self.stream.wait_stream(th.cuda.current_stream())
            for k in range(2):
                if self.bufs[k] is None:
                    self.bufs[k] = th.empty(self.CHUNK, dtype=th.uint8, pin_memory=True)
                if self.events[k] is not None:
                    self.events[k].synchronize()
                    self.events[k] = None
            offs = list(range(0, n, self.CHUNK))
            done = [None, None]

            def fetch(i):
                off = offs[i]
                m = min(self.CHUNK, n - off)
                with th.cuda.stream(self.stream):
                    self.bufs[i & 1][:m].copy_(src[off : off + m], non_blocking=True)
                    ev = th.cuda.Event()
                    ev.record(self.stream)
                done[i & 1] = ev

            if offs:
                fetch(0)
            for i, off in enumerate(offs):
                m = min(self.CHUNK, n - off)
                done[i & 1].synchronize()
                if i + 1 < len(offs):
                    fetch(i + 1)  # the other pinned buffer: its previous contents were drained last iteration
                buf = self.bufs[i & 1].numpy()
                cuts = [m * j // self.threads for j in range(self.threads + 1)]
                list(self.pool.map(lambda ab: np.copyto(flat[off + ab[0] : off + ab[1]], buf[ab[0] : ab[1]]),
                                   zip(cuts[:-1], cuts[1:])))
            src.record_stream(self.stream)
        return out


_stagers = {}
_stagers_lock = None


def _get_stager() -> "_HostStager":
    """One stager (two pinned buffers + a copy stream) per device: several devices are fed from several host
    threads at once (``analysis`` fans a host trajectory out over all visible GPUs)."""
    global _stagers_lock
    import threading

    if _stagers_lock is None:
        _stagers_lock = threading.Lock()
    dev = torch.cuda.current_device()
    with _stagers_lock:
        st = _stagers.get(dev)
        if st is None:
            st = _stagers[dev] = _HostStager()
        return st


def _staged_upload(arr: np.ndarray, tdtype):
    return _get_stager().upload(arr).view(tdtype).view(*arr.shape)


def upload_atom_range(arr: np.ndarray, lo: int, hi: int):
    """``arr[:, lo:hi, :]`` of a C-contiguous host array ``[F, A, D]`` -> contiguous CUDA tensor ``[F, hi-lo, D]``
    through the pinned buffers (the strided rows are gathered by the stager's host threads; numpy's own
    ``ascontiguousarray`` would do it on one thread at a few GB/s)."""
    th = require_cuda()
    F, A, D = arr.shape
    tdtype = th.from_numpy(np.empty(0, dtype=arr.dtype)).dtype
    return _get_stager().upload_rows(arr, lo, hi).view(tdtype).view(F, hi - lo, D)


def to_device(x, dtype=None):
    """numpy / tensor -> contiguous CUDA tensor (H2D through pinned memory for numpy)."""
    th = require_cuda()
    if is_tensor(x):
        if (not x.is_cuda and not x.is_pinned() and x.is_contiguous() and x.numel() * x.element_size() >= _HostStager.MIN_BYTES
                and x.dtype in (th.float32, th.float64, th.int32, th.int64)):
            t = _staged_upload(x.numpy(), x.dtype)
        else:
            t = x if x.is_cuda else x.cuda(non_blocking=True)
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.contiguous()
    arr = np.asarray(x)
    if arr.dtype == np.float16:
        arr = arr.astype(np.float32)
    if not arr.flags.c_contiguous or not arr.flags.aligned or arr.dtype.byteorder not in "=|<":
        arr = np.ascontiguousarray(arr.astype(arr.dtype.newbyteorder("="), copy=False))
    if arr.nbytes >= _HostStager.MIN_BYTES and arr.dtype.kind in "fiu" and arr.dtype.itemsize in (4, 8):
        t = _staged_upload(arr, th.from_numpy(np.empty(0, dtype=arr.dtype)).dtype)
        return t if dtype is None or t.dtype == dtype else t.to(dtype)
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
    if t.is_cuda and t.numel() * t.element_size() >= _HostStager.MIN_BYTES and t.dtype in (
            torch.float32, torch.float64, torch.int32, torch.int64):
        with torch.cuda.device(t.device):
            return _get_stager().download(t)
    return t.detach().cpu().numpy()


def resolve_devices(devices=None, nbytes: int = 0):
    """Device indices a HOST input is fanned out over.  ``devices``: ``None`` (policy below), ``"all"``, an int
    or a sequence of ints.  Policy for ``None``: the environment variable ``SOAP_B200_DEVICES`` (``all`` or a comma
    separated list) if set; the current device alone inside a ``torch.distributed`` job (one process per GPU
    already) or for small inputs; otherwise every visible GPU."""
    import os

    th = require_cuda()
    cur = th.cuda.current_device()
    if devices is None:
        env = os.environ.get("SOAP_B200_DEVICES", "").strip()
        if env:
            devices = env
        elif int(os.environ.get("WORLD_SIZE", "1")) > 1 or (th.distributed.is_available() and th.distributed.is_initialized()):
            return [cur]
        elif nbytes < (1 << 30):
            return [cur]
        else:
            devices = "all"
    if isinstance(devices, str):
        if devices.lower() == "all":
            return list(range(th.cuda.device_count()))
        devices = [int(d) for d in devices.split(",") if d.strip()]
    if isinstance(devices, int):
        devices = [devices]
    devices = [int(d) for d in devices]
    if not devices or any(d < 0 or d >= th.cuda.device_count() for d in devices):
        raise ValueError(f"devices={devices}: this process sees {th.cuda.device_count()} CUDA device(s)")
    return devices


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
