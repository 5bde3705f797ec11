"""Mirror of the timeSOAP family of ``SOAPify.analysis`` (``src/SOAPify/analysis.py:13-203``).

``timeSOAP`` / ``timeSOAPsimple`` / ``getTimeSOAPSimple`` keep the reference's signatures, shapes,
dtypes and error messages; the work is done by the time-lag kernel of ``libsoapb200.so``
(``csrc/timelag.cu``).  Additions that have no reference counterpart are marked as such:
``timeSOAPsweep`` (several windows in one pass), ``timeSOAPfromCompressed`` (expansion and
normalisation folded into the distance) and the north-star alias ``tempoSOAP``.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _device, _lib
from .distances import resolve_distance, simpleSOAPdistance
from .utils import getExpansionPlan, getSOAPSettings, orderByZ, _quippyContainerSlices

__all__ = [
    "timeSOAP",
    "tempoSOAP",
    "timeSOAPsimple",
    "getTimeSOAPSimple",
    "timeSOAPsweep",
    "timeSOAPfromCompressed",
]


def _validate(n_frames: int, window: int, stride):
    # analysis.py:47-52 (messages pinned by the reference's tests/test_analysis.py:41,46)
    if stride is None:
        stride = window
    if stride > window:
        raise ValueError("the window must be bigger than the stride")
    if window >= n_frames or stride >= n_frames:
        raise ValueError("stride and window must be smaller than simulation lenght")
    if window < 1:
        raise ValueError("the window must be at least 1")


def timelag_device(X, windows, kind: int, power: int = 1, mult=None):
    """Run the time-lag kernel on a CUDA tensor ``X[F, A, D]``; returns one ``[F-w, A]`` float64
    CUDA tensor per window (all windows of a group of <= 4 share ONE pass over ``X``)."""
    th = _device.require_cuda()
    lib = _lib.load()
    if X.dim() != 3:
        raise ValueError("a SOAP trajectory must have shape (nFrames, nAtoms, SOAPlength)")
    F, A, D = (int(s) for s in X.shape)
    windows = [int(w) for w in windows]
    es = X.element_size()
    tma_rows = (D * es) % 16 == 0 or ((A * D * es) % 16 == 0 and X.data_ptr() % 16 == 0)
    if not tma_rows and F * A > 4096:
        # the TMA engine fetches rows of any length as long as a FRAME is a whole number of 16-byte units (an atom's
        # window then starts at the unit boundary below its row, foreign elements masked).  Otherwise the kernel
        # would fall back to element-wise cp.async staging (~0.15 of the HBM peak): when memory allows, pad the
        # rows once on the device instead (one extra read + write pass)
        unit = 16 // es
        Dp = (D + unit - 1) // unit * unit
        free, _ = th.cuda.mem_get_info()
        if F * A * Dp * es * 1.05 < free:
            Xp = th.zeros((F, A, Dp), dtype=X.dtype, device=X.device)
            Xp[..., :D] = X
            mp = th.zeros(Dp, dtype=X.dtype, device=X.device)
            mp[:D] = 1 if mult is None else mult
            X, mult, D = Xp, mp, Dp
    results = {}
    todo = sorted(set(windows))
    for g in range(0, len(todo), _lib.MAX_WINDOWS):
        group = todo[g : g + _lib.MAX_WINDOWS]
        nw = len(group)
        sizes = [(F - w) * A for w in group]
        offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
        t = th.empty(int(sum(sizes)), dtype=th.float64, device=X.device)
        c_w = (ctypes.c_int * nw)(*group)
        c_off = (ctypes.c_int64 * nw)(*[int(o) for o in offsets])
        code = _device.dtype_code(X)
        nbytes = lib.soap_timelag_scratch_bytes(F, A, D, code, c_w, nw)
        if nbytes == 0:
            raise _lib.SoapB200Error(f"timeSOAP: window {max(group)} does not fit the shared-memory ring")
        scratch = th.empty(nbytes // 8, dtype=th.float64, device=X.device)
        rc = lib.soap_timelag(
            X.data_ptr(), None if mult is None else mult.data_ptr(), t.data_ptr(), c_off, F, A, D, c_w, nw,
            kind, int(power), code, scratch.data_ptr(), _device.stream_ptr(),
        )
        _lib.check(rc, "soap_timelag")
        for w, o, n in zip(group, offsets, sizes):
            results[w] = t[int(o) : int(o) + n].view(F - w, A)
    return [results[w] for w in windows]


def diff_transposed_device(t):
    """``numpy.diff(t.T, axis=-1)`` on the device: ``[rows, A] -> [A, rows-1]`` (analysis.py:67)."""
    th = _device.require_cuda()
    rows, A = (int(s) for s in t.shape)
    dt = th.empty((A, max(rows - 1, 0)), dtype=th.float64, device=t.device)
    if rows > 1 and A > 0:
        rc = _lib.load().soap_diff_transposed(t.data_ptr(), dt.data_ptr(), rows, A, _device.stream_ptr())
        _lib.check(rc, "soap_diff_transposed")
    return dt


def _atom_tiles(x, bytes_per_atom: int):
    """Atom ranges for HOST inputs that do not fit in HBM: every atom's time series is independent,
    so a ``[F, A, D]`` trajectory is processed in atom tiles sized to ~40 % of the free device memory
    (input tile + partial sums + results), results concatenated along the atom axis."""
    n_atoms = int(x.shape[1])
    if _device.is_tensor(x) and x.is_cuda:
        return [(0, n_atoms)]
    th = _device.require_cuda()
    free, _ = th.cuda.mem_get_info()
    per_tile = max(1, int(0.4 * free) // max(1, bytes_per_atom))
    if per_tile >= n_atoms:
        return [(0, n_atoms)]
    return [(lo, min(n_atoms, lo + per_tile)) for lo in range(0, n_atoms, per_tile)]


def _run_windows(X, windows, kind, power, mult, simple):
    """One device tile: the time-lag kernel, or (``simple``) the row stitching of ``timeSOAPsimple``.
    ``mult`` may be a callable returning the multiplicities ON THE CURRENT DEVICE."""
    if simple:
        return [_simple_rows(X, windows[0])]
    return timelag_device(X, windows, kind, power, mult=mult() if callable(mult) else mult)


def _join_atoms(parts_per_window, like, returnDiff):
    """Concatenate per-tile results along the atom axis (``t``: axis 1, ``dt``: axis 0)."""
    th = _device.torch
    res = []
    for parts in parts_per_window:
        cat_t = (lambda xs, d: th.cat(xs, dim=d)) if _device.is_tensor(like) else (lambda xs, d: np.concatenate(xs, axis=d))
        if returnDiff:
            res.append((cat_t([p[0] for p in parts], 1), cat_t([p[1] for p in parts], 0)))
        else:
            res.append(cat_t(parts, 1))
    return res


def _host_tile(x, lo, hi):
    """Atoms ``lo:hi`` of a host trajectory as a contiguous CUDA tensor (strided rows gathered by host threads)."""
    if (isinstance(x, np.ndarray) and x.ndim == 3 and x.flags.c_contiguous and x.dtype in (np.float32, np.float64)
            and x.dtype.byteorder in "=|<" and (hi - lo) * x.shape[0] * x.shape[2] * x.itemsize >= _device._HostStager.MIN_BYTES):
        return _device.upload_atom_range(x, lo, hi)
    return _device.as_float_tensor(x[:, lo:hi, :])


def _tiled(x, windows, kind, power, mult, returnDiff, simple: bool = False, devices=None, atoms=None):
    """Run the time-lag kernel over atom tiles of ``x``; returns per window ``t`` (and ``dt``) in the
    container type of ``x``.  ``simple``: ``windows == [w]`` and the rows are those of ``timeSOAPsimple``.

    HOST inputs are fanned out over ``devices`` (see :func:`_device.resolve_devices`: by default every visible
    GPU for inputs of a GB or more outside ``torch.distributed`` jobs): contiguous atom ranges, one host thread
    and one upload stream per device, so that a notebook call uses every PCIe link of the box.  ``atoms`` restricts
    the work to an atom range of ``x`` (what a device worker is handed)."""
    th = _device.require_cuda()
    F, A_all, D = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
    itemsize = 4 if str(getattr(x, "dtype", "")).endswith("float32") else 8
    on_host = not (_device.is_tensor(x) and x.is_cuda)
    lo0, hi0 = atoms if atoms is not None else (0, A_all)
    if on_host and atoms is None:
        devs = _device.resolve_devices(devices, F * A_all * D * itemsize)
        if len(devs) > 1 and A_all >= len(devs):
            return _fan_out(x, windows, kind, power, mult, returnDiff, simple, devs)
        if devs[0] != th.cuda.current_device():
            with th.cuda.device(devs[0]):
                return _tiled(x, windows, kind, power, mult, returnDiff, simple, devices=devs, atoms=(0, A_all))
    per_atom = F * D * itemsize + F * 8 * (len(windows) + 1) * 8
    if _pinned_host_tensor(x) and F * (hi0 - lo0) * D * itemsize >= PIPELINE_MIN_BYTES:
        return _tiled_pipelined(x, windows, kind, power, mult, returnDiff, per_atom, simple, (lo0, hi0))
    if not on_host:
        return [_finish(t, x, returnDiff) for t in _run_windows(_device.as_float_tensor(x), windows, kind, power, mult, simple)]
    tiles = [(lo0 + lo, lo0 + hi) for lo, hi in _atom_tiles(x[:, lo0:hi0], per_atom)]
    if tiles == [(0, A_all)]:
        X = _device.as_float_tensor(x)
        return [_finish(t, x, returnDiff) for t in _run_windows(X, windows, kind, power, mult, simple)]
    outs = [[] for _ in windows]
    for lo, hi in tiles:
        X = _host_tile(x, lo, hi)
        for k, t in enumerate(_run_windows(X, windows, kind, power, mult, simple)):
            outs[k].append(_finish(t, x, returnDiff))
        del X
    return _join_atoms(outs, x, returnDiff)


def _fan_out(x, windows, kind, power, mult, returnDiff, simple, devs):
    """One host thread per device, each taking a contiguous atom range of the host trajectory through the
    single-device path (its own pinned staging buffers, copy stream and kernels); results joined on the host."""
    from concurrent.futures import ThreadPoolExecutor

    from .sharding import shard_range

    th = _device.require_cuda()
    A = int(x.shape[1])
    ranges = [shard_range(A, i, len(devs)) for i in range(len(devs))]

    def work(i):
        r = ranges[i]
        if r.stop == r.start:
            return None
        with th.cuda.device(devs[i]):
            res = _tiled(x, windows, kind, power, mult, returnDiff, simple, devices=[devs[i]], atoms=(r.start, r.stop))
            th.cuda.synchronize()
        return res

    with ThreadPoolExecutor(len(devs)) as pool:
        parts = [p for p in pool.map(work, range(len(devs))) if p is not None]
    return _join_atoms([[p[k] for p in parts] for k in range(len(windows))], x, returnDiff)


#: pinned host trajectories at least this large are uploaded in atom tiles while the previous tile computes
PIPELINE_MIN_BYTES = 256 << 20
#: target size of one uploaded atom tile
PIPELINE_TILE_BYTES = 1 << 30


def _pinned_host_tensor(x) -> bool:
    return (_device.is_tensor(x) and not x.is_cuda and x.dim() == 3 and x.is_contiguous() and x.is_pinned()
            and x.dtype in (_device.torch.float32, _device.torch.float64))


def _tiled_pipelined(x, windows, kind, power, mult, returnDiff, per_atom, simple: bool = False, atoms=None):
    """Pinned host trajectory -> host results with the link kept busy: atom tiles ``x[:, lo:hi, :]`` are
    uploaded on a copy stream (strided ``cudaMemcpy2DAsync`` into one of two device buffers) while the
    previous tile is processed.  The step then costs the H2D time of ``x`` plus the tail of the last tile and
    one staged download of the (small) results."""
    th = _device.require_cuda()
    lib = _lib.load()
    F, A_all, D = (int(v) for v in x.shape)
    a0, a1 = atoms if atoms is not None else (0, A_all)
    A = a1 - a0  # the atoms this call works on (results are [F-w, A] / [A, F-w-1])
    es = x.element_size()
    free, _ = th.cuda.mem_get_info()
    n_tiles = max(2, -(-F * A * D * es // PIPELINE_TILE_BYTES))
    while n_tiles < A and 2 * -(-A // n_tiles) * per_atom > 0.5 * free:
        n_tiles *= 2
    tile = -(-A // min(n_tiles, A))
    ranges = [(lo, min(A, lo + tile)) for lo in range(0, A, tile)]
    bufs = [th.empty((F, tile, D), dtype=x.dtype, device="cuda") for _ in range(2)]
    copy_stream = th.cuda.Stream()
    main = th.cuda.current_stream()
    copied = [th.cuda.Event() for _ in ranges]
    done = [th.cuda.Event() for _ in ranges]
    # results are assembled on the device in ONE flat buffer (tiny next to the input) and come back through the
    # pinned staging buffers with a threaded copy: per-tile copies into fresh pageable tensors take their page
    # faults on one thread and made the step time vary by 10 %
    sizes = [(F - w) * A for w in windows] + ([A * (F - w - 1) for w in windows] if returnDiff else [])
    res = th.empty(int(sum(sizes)), dtype=th.float64, device="cuda")
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    nw = len(windows)
    t_dev = [res[int(offs[k]) : int(offs[k + 1])].view(F - w, A) for k, w in enumerate(windows)]
    dt_dev = [res[int(offs[nw + k]) : int(offs[nw + k + 1])].view(A, F - w - 1) for k, w in enumerate(windows)] if returnDiff else None

    def upload(i):
        lo, hi = ranges[i]
        with th.cuda.stream(copy_stream):
            if i >= 2:
                copy_stream.wait_event(done[i - 2])  # the tile that last used this buffer has been consumed
            rc = lib.soap_copy_rows(
                bufs[i & 1].data_ptr(), (hi - lo) * D * es, x.data_ptr() + (a0 + lo) * D * es, A_all * D * es, (hi - lo) * D * es, F, 1,
                int(copy_stream.cuda_stream),
            )
            _lib.check(rc, "soap_copy_rows")
            copied[i].record(copy_stream)

    upload(0)
    for i, (lo, hi) in enumerate(ranges):
        if i + 1 < len(ranges):
            upload(i + 1)
        main.wait_event(copied[i])
        X = bufs[i & 1][:, : hi - lo, :] if hi - lo == tile else bufs[i & 1].view(-1)[: F * (hi - lo) * D].view(F, hi - lo, D)
        ts = _run_windows(X, windows, kind, power, mult, simple)
        for k in range(nw):
            t_dev[k][:, lo:hi] = ts[k]
            if returnDiff:
                dt_dev[k][lo:hi] = diff_transposed_device(ts[k])
        done[i].record(main)
    host = th.from_numpy(_device.to_host(res))
    t_host = [host[int(offs[k]) : int(offs[k + 1])].view(F - w, A) for k, w in enumerate(windows)]
    if returnDiff:
        dt_host = [host[int(offs[nw + k]) : int(offs[nw + k + 1])].view(A, F - w - 1) for k, w in enumerate(windows)]
        return list(zip(t_host, dt_host))
    return t_host


def _finish(t, like, returnDiff):
    if returnDiff:
        return _device.like_input(t, like), _device.like_input(diff_transposed_device(t), like)
    return _device.like_input(t, like)


def timeSOAP(
    SOAPTrajectory,
    window: int = 1,
    stride: int = None,
    backward: bool = False,
    returnDiff: bool = True,
    distanceFunction: callable = simpleSOAPdistance,
    devices=None,
):
    """performs the 'timeSOAP' analysis on the given SOAP trajectory (analysis.py:13-69).

    ``devices`` (extension, host inputs only): GPUs the atoms are fanned out over -- ``None`` = every visible GPU
    for trajectories of a GB or more (one process, one upload stream per device), ``"all"``, an index or a list.

    ``t[f-window, a] = distanceFunction(X[f, a], X[f-window, a])``; returns ``t`` (float64,
    ``(frames-window, natoms)``) and, if ``returnDiff``, ``numpy.diff(t.T, axis=-1)``.  ``stride`` and
    ``backward`` are validated but unused, as in the reference.  ``distanceFunction`` must be one of
    the built-in distances of :mod:`cpctools_b200.distances` (dispatched by identity).
    """
    _validate(int(SOAPTrajectory.shape[0]), window, stride)
    kind, power = resolve_distance(distanceFunction)
    if len(SOAPTrajectory.shape) != 3:
        raise ValueError("a SOAP trajectory must have shape (nFrames, nAtoms, SOAPlength)")
    return _tiled(SOAPTrajectory, [int(window)], kind, power, None, returnDiff, devices=devices)[0]


def tempoSOAP(*args, **kwargs):
    """Historical name of :func:`timeSOAP` (reference ``Changelog.md:51``); north-star alias."""
    return timeSOAP(*args, **kwargs)


def timeSOAPsweep(SOAPTrajectory, windows=(1, 5, 10, 50), returnDiff: bool = True, distanceFunction: callable = simpleSOAPdistance,
                  devices=None):
    """``[timeSOAP(X, window=w, ...) for w in windows]`` with ONE pass over ``X`` per group of four
    windows (no reference counterpart: the reference would be called once per window)."""
    n_frames = int(SOAPTrajectory.shape[0])
    for w in windows:
        _validate(n_frames, int(w), None)
    kind, power = resolve_distance(distanceFunction)
    if len(SOAPTrajectory.shape) != 3:
        raise ValueError("a SOAP trajectory must have shape (nFrames, nAtoms, SOAPlength)")
    return _tiled(SOAPTrajectory, [int(w) for w in windows], kind, power, None, returnDiff, devices=devices)


def timeSOAPfromCompressed(
    soapFromdscribe,
    lMax: int,
    nMax: int,
    atomTypes: list = None,
    atomicSlices: dict = None,
    windows=(1,),
    returnDiff: bool = True,
    distanceFunction: callable = simpleSOAPdistance,
    quippy: bool = False,
    devices=None,
):
    """``timeSOAP(normalizeArray(fillSOAPVectorFromdscribe(x, ...)), window=w)`` for every ``w`` in
    ``windows`` WITHOUT materialising the expanded vectors: the dot products of the expanded
    vectors are multiplicity-weighted dot products of the compressed ones, and the cosine-type
    distances do not depend on the normalisation (``SOAPdistanceNormalized`` divides by the norms,
    zero norm -> 1, exactly as ``normalizeArray`` would have).  Reads ``Dc`` instead of ``Df``
    elements per vector.  ``quippy=True`` takes RAW quippy rows (``[..., Dc+1]``).
    """
    if quippy:
        species = orderByZ([str(s) for s in atomTypes])
        plan = getExpansionPlan(lMax, nMax, species, atomicSlices or _quippyContainerSlices(lMax, nMax, species), quippy=True)
    else:
        plan = getExpansionPlan(lMax, nMax, atomTypes, atomicSlices)
    if soapFromdscribe.shape[-1] != plan.d_compressed:
        raise ValueError("fillSOAPVectorFromdscribe: the given soap vector do not have the expected dimensions")
    n_frames = int(soapFromdscribe.shape[0])
    for w in windows:
        _validate(n_frames, int(w), None)
    kind, power = resolve_distance(distanceFunction)
    if kind == _lib.KIND_NORMALIZED:
        kind = _lib.KIND_NORMALIZED_FUSED
    if len(soapFromdscribe.shape) != 3:
        raise ValueError("a SOAP trajectory must have shape (nFrames, nAtoms, SOAPlength)")
    th = _device.require_cuda()
    f32 = str(soapFromdscribe.dtype).endswith("float32")
    tdtype = th.float32 if f32 else th.float64
    # evaluated where the tile runs: a host trajectory may be fanned out over several devices
    return _tiled(soapFromdscribe, [int(w) for w in windows], kind, power, lambda: plan.multiplicity_device(tdtype), returnDiff,
                  devices=devices)


def _simple_rows(X, window: int):
    """Rows of ``timeSOAPsimple`` on a CUDA tensor: the reference's ``prev`` trails by ONE frame
    (analysis.py:132-137), so row 0 is ``|X[w] - X[0]|`` and row r >= 1 is ``|X[w+r] - X[w+r-1]|``."""
    th = _device.require_cuda()
    F = int(X.shape[0])
    if window == 1:
        return timelag_device(X, [1], _lib.KIND_EUCLID)[0]
    lag1, lagw = timelag_device(X, [1, window], _lib.KIND_EUCLID)
    t = th.empty((F - window, int(X.shape[1])), dtype=th.float64, device=X.device)
    t[0] = lagw[0]
    if F - window > 1:
        t[1:] = lag1[window:]
    return t


def timeSOAPsimple(SOAPTrajectory, window: int = 1, stride: int = None, backward: bool = False, returnDiff: bool = True,
                   devices=None):
    """performs the 'timeSOAP' analysis on the given **normalized** SOAP trajectory (analysis.py:72-142).

    Euclidean distance between consecutive frames' versors.  Kept literally, including the
    reference's behaviour for ``window > 1`` (only row 0 spans ``window`` frames, see
    :func:`_simple_rows`); use :func:`timeSOAP` for true lag-``window`` distances.
    """
    _validate(int(SOAPTrajectory.shape[0]), window, stride)
    if len(SOAPTrajectory.shape) != 3:
        raise ValueError("a SOAP trajectory must have shape (nFrames, nAtoms, SOAPlength)")
    # atoms are independent for the Euclidean kind too: host trajectories larger than HBM go through the same
    # atom tiles (and the pinned upload pipeline) as timeSOAP, the rows are stitched per tile
    return _tiled(SOAPTrajectory, [int(window)], _lib.KIND_EUCLID, 1, None, returnDiff, simple=True, devices=devices)[0]


def getTimeSOAPSimple(soapDataset, window: int = 1, stride: int = None, backward: bool = False):
    """Shortcut to extract the timeSOAP from large datasets (analysis.py:145-203).

    Streams the dataset chunk by chunk (``iter_chunks()``), each chunk going pinned host buffer ->
    HBM -> fused expand+normalise kernel -> Euclidean time-lag kernel, with the last ``window``
    frames of a chunk kept ON THE DEVICE as the halo of the next one (the reference re-reads one
    frame from HDF5).  ``window == 1`` reproduces the reference; for ``window > 1`` the reference
    raises a broadcast ``ValueError`` while this returns the true
    lag-``window`` distances.  Returns ``(t[F-window, A], numpy.diff(t.T, axis=-1))`` as numpy arrays.
    """
    from .streaming import stream_time_soap

    n_frames = int(soapDataset.shape[0])
    _validate(n_frames, window, stride)
    settings = getSOAPSettings(soapDataset)
    (t,) = stream_time_soap(soapDataset, settings, [int(window)], euclid=True)
    return _device.to_host(t), _device.to_host(diff_transposed_device(t))
