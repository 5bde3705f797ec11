"""Chunk streaming: HDF5-like dataset -> pinned host ring -> HBM -> kernels.

The reference reads one HDF5 frame-chunk at a time with a one-frame halo
(``src/SOAPify/analysis.py:186-201``) or 100 frames at a time (``classify.py:140-158``).  Here a
chunk is read into one of two pinned buffers, copied to HBM on a side stream while the previous
block is being processed, and the halo (the last ``max(windows)`` frames) never leaves the
device.  Any object with ``.shape``, ``.attrs``, ``.chunks``/``.iter_chunks()`` and slicing works
(an ``h5py.Dataset`` or ``oracle.fake_dataset.FakeSOAPDataset``).
"""
from __future__ import annotations

import numpy as np

from . import _device, _lib
from .utils import getExpansionPlan

#: expanded bytes one device block may occupy; bigger HDF5 chunks are cut into several blocks
MAX_BLOCK_BYTES = 8 << 30


def frame_ranges(dataset, max_frames: int = None):
    """Frame ranges in dataset order: the HDF5 chunk boundaries on axis 0 (``iter_chunks``), cut so
    that no range exceeds ``max_frames``."""
    n_frames = int(dataset.shape[0])
    bounds = set()
    if hasattr(dataset, "iter_chunks"):
        try:
            for c in dataset.iter_chunks():
                first = c[0] if isinstance(c, tuple) else c
                bounds.add((int(first.start), int(first.stop)))
        except TypeError:  # contiguous (unchunked) h5py dataset
            bounds = set()
    if not bounds:
        step = int(dataset.chunks[0]) if getattr(dataset, "chunks", None) else n_frames
        bounds = {(lo, min(n_frames, lo + step)) for lo in range(0, n_frames, max(step, 1))}
    out = []
    for lo, hi in sorted(bounds):
        step = hi - lo if not max_frames else max(1, min(hi - lo, int(max_frames)))
        out.extend((s, min(hi, s + step)) for s in range(lo, hi, step))
    return out


class _Uploader:
    """Two pinned staging buffers + a copy stream: ``put(array)`` returns a device tensor whose H2D
    copy is ordered before later work on the current stream."""

    def __init__(self, th):
        self.th = th
        self.copy_stream = th.cuda.Stream()
        self.pinned = [None, None]
        self.busy = [None, None]
        self.turn = 0
        import os

        self.fill_threads = max(1, min(8, (os.cpu_count() or 2) // 2))

    def _stage(self, k: int, shape, dtype):
        """pinned staging tensor ``k`` viewed as ``shape`` (waits for the copy that last used it)"""
        th = self.th
        n = int(np.prod(shape))
        if self.busy[k] is not None:
            self.busy[k].synchronize()
        if self.pinned[k] is None or self.pinned[k].numel() < n or self.pinned[k].dtype != dtype:
            self.pinned[k] = th.empty(n, dtype=dtype, pin_memory=True)
        return self.pinned[k][:n].view(*shape)

    def _launch(self, k: int, stage):
        th = self.th
        with th.cuda.stream(self.copy_stream):
            dev = stage.to("cuda", non_blocking=True)
            done = th.cuda.Event()
            done.record(self.copy_stream)
        self.busy[k] = done
        th.cuda.current_stream().wait_event(done)
        dev.record_stream(th.cuda.current_stream())
        return dev

    def put_frames(self, dataset, lo: int, hi: int):
        """Frames ``[lo, hi)`` of a chunked dataset -> device tensor, read STRAIGHT into the pinned staging
        buffer: ``dataset.read_direct`` when the dataset has it (h5py does: no intermediate array), plain
        slicing otherwise; the frame range is split over a few host threads (numpy releases the GIL while
        it copies; h5py serialises internally, which is harmless)."""
        th = self.th
        dt = np.dtype(getattr(dataset, "dtype", np.float64))
        if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
            host = np.asarray(dataset[lo:hi]).astype(np.float64)
            return self.put(host)
        k = self.turn
        self.turn ^= 1
        shape = (hi - lo,) + tuple(int(v) for v in dataset.shape[1:])
        stage = self._stage(k, shape, th.float32 if dt == np.dtype(np.float32) else th.float64)
        dst = stage.numpy()
        direct = getattr(dataset, "read_direct", None)

        def fill(a, b):
            if direct is not None:
                direct(dst, np.s_[lo + a : lo + b], np.s_[a:b])
            else:
                dst[a:b] = dataset[lo + a : lo + b]

        n = hi - lo
        parts = max(1, min(self.fill_threads, n))
        if parts == 1:
            fill(0, n)
        else:
            from concurrent.futures import ThreadPoolExecutor

            cuts = [n * i // parts for i in range(parts + 1)]
            with ThreadPoolExecutor(parts) as pool:
                list(pool.map(lambda ab: fill(*ab), zip(cuts[:-1], cuts[1:])))
        return self._launch(k, stage)

    def put(self, host: np.ndarray):
        th = self.th
        k = self.turn
        self.turn ^= 1
        src = th.from_numpy(np.ascontiguousarray(host))
        if self.busy[k] is not None:
            self.busy[k].synchronize()  # the copy that last used this pinned buffer
        if self.pinned[k] is None or self.pinned[k].numel() < src.numel() or self.pinned[k].dtype != src.dtype:
            self.pinned[k] = th.empty(src.numel(), dtype=src.dtype, pin_memory=True)
        stage = self.pinned[k][: src.numel()].view(src.shape)
        stage.copy_(src)
        with th.cuda.stream(self.copy_stream):
            dev = stage.to("cuda", non_blocking=True)
            done = th.cuda.Event()
            done.record(self.copy_stream)
        self.busy[k] = done
        th.cuda.current_stream().wait_event(done)
        dev.record_stream(th.cuda.current_stream())
        return dev


def stream_time_soap(dataset, settings: dict, windows, euclid: bool = True, kind: int = _lib.KIND_SIMPLE,
                     power: int = 1, max_block_bytes: int = None):
    """timeSOAP of a chunked compressed dataset for every window in ``windows`` (true lag semantics).

    ``euclid=True``: the arithmetic of ``timeSOAPsimple`` on ``normalizeArray(fillSOAPVectorFromdscribe(chunk))``
    (analysis.py:192-200, :136) WITHOUT materialising the expanded vectors: one pass over the compressed block
    for the norms of the expanded vectors (multiplicity-weighted), then ``|x/|x| - y/|y||`` summed with the
    multiplicities by the time-lag kernel (``SOAP_KIND_EUCLID_NORMALIZED``): 2 x Dc elements read per vector
    instead of Dc + 3 x Df.  ``euclid=False``: distances of ``kind`` straight from the compressed block
    (expansion folded into multiplicities).  Returns CUDA float64 tensors ``[F-w, A]``.
    """
    from .analysis import timelag_device

    th = _device.require_cuda()
    lib = _lib.load()
    F, A, Dc = (int(s) for s in dataset.shape)
    plan = getExpansionPlan(settings["lMax"], settings["nMax"], settings["atomTypes"], settings["atomicSlices"])
    if Dc != plan.d_compressed:
        raise ValueError("fillSOAPVectorFromdscribe: the given soap vector do not have the expected dimensions")
    windows = [int(w) for w in windows]
    maxw = max(windows)
    width = Dc
    budget = max_block_bytes or MAX_BLOCK_BYTES
    max_frames = max(maxw + 1, budget // max(1, A * width * 8) - maxw)
    outs = [th.empty((F - w, A), dtype=th.float64, device="cuda") for w in windows]
    up = _Uploader(th)
    tail = None  # last <= maxw processed frames, on the device
    for lo, hi in frame_ranges(dataset, max_frames):
        raw = up.put_frames(dataset, lo, hi)
        h = 0 if tail is None else int(tail.shape[0])
        block = th.empty((h + hi - lo, A, width), dtype=raw.dtype, device="cuda")
        if h:
            block[:h].copy_(tail)
        block[h:].copy_(raw)
        nb = h + hi - lo
        use = [w for w in windows if w < nb]
        if use:
            res = timelag_device(block, use, _lib.KIND_EUCLID_NORMALIZED if euclid else kind, power,
                                 mult=plan.multiplicity_device(block.dtype))
            for w, tb in zip(use, res):
                # block-local frame g pairs with g-w; only frames of THIS range (g >= h) are new
                g0 = max(h, w)
                if g0 < nb:
                    dst0 = (lo - h) + g0 - w
                    outs[windows.index(w)][dst0 : dst0 + nb - g0] = tb[g0 - w :]
        keep = min(maxw, nb)
        tail = block[nb - keep :].clone() if keep < nb else block
    return outs
