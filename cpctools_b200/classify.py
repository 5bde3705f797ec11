"""Mirror of the distance-matrix / classification part of ``SOAPify.classify``
(``src/SOAPify/classify.py:17-48, 51-93, 96-182, 185-205, 250-276``).

``getDistanceBetween`` is the all-pairs boundary (``getDistanceBetween(X, X, f)``); the
``getDistancesFromRef*`` / ``applyClassification`` functions stream a chunked dataset through
expand -> normalise -> distance-to-references (-> fused argmin) on the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

import numpy as np

from . import _device, _lib
from .distances import SOAPdistanceNormalized, resolve_distance
from .streaming import _Uploader, frame_ranges
from .utils import fillSOAPVectorFromdscribe, getExpansionPlan, normalizeArray

__all__ = [
    "SOAPclassification",
    "SOAPReferences",
    "createReferencesFromTrajectory",
    "getDistanceBetween",
    "getDistancesFromRef",
    "getDistancesFromRefNormalized",
    "getDistancesFromInterval",
    "mergeReferences",
    "saveReferences",
    "getReferencesFromDataset",
    "applyClassification",
]


@dataclass
class SOAPclassification:
    """Stores the information about the SOAP classification of a trajectory (classify.py:17-29)."""

    distances: "np.ndarray[float]"  #: per frame, per atom distance from the closest reference
    references: "np.ndarray[int]"  #: per frame, per atom index of the closest reference
    legend: "list[str]"  #: the references legend


@dataclass
class SOAPReferences:
    """Stores the spectra selected for an environments dictionary (classify.py:31-48)."""

    names: "list[str]"
    spectra: "np.ndarray[np.float64]"
    lmax: int
    nmax: int

    def __len__(self) -> int:
        return len(self.names)


def createReferencesFromTrajectory(h5SOAPDataSet, addresses: dict, lmax: int, nmax: int, doNormalize=True) -> SOAPReferences:
    """Pick ``dataset[frame, atom]`` fingerprints as references (classify.py:51-93).  Expansion and
    normalisation of the picked vectors run on the GPU.  The reference compares the stored
    dimension with ``lmax*nmax*nmax`` (not ``(lmax+1)``, classify.py:85) to decide whether to expand;
    that test is kept as is."""
    names = list(addresses.keys())
    dim = h5SOAPDataSet.shape[2]
    spectra = np.empty((len(addresses), dim), dtype=h5SOAPDataSet.dtype)
    for i, key in enumerate(addresses):
        spectra[i] = h5SOAPDataSet[addresses[key][0], addresses[key][1]]
    if lmax * nmax * nmax != dim:
        spectra = fillSOAPVectorFromdscribe(spectra, lmax, nmax)
    if doNormalize:
        spectra = normalizeArray(spectra)
    return SOAPReferences(names, spectra, lmax, nmax)


def mergeReferences(*x: SOAPReferences) -> SOAPReferences:
    """Concatenate reference dictionaries (classify.py:185-205)."""
    names = []
    for i in x:
        names += i.names
        if x[0].nmax != i.nmax or x[0].lmax != i.lmax:
            raise ValueError("nmax or lmax are not the same in the two references")
    return SOAPReferences(names, np.concatenate([i.spectra for i in x]), nmax=x[0].nmax, lmax=x[0].lmax)


def saveReferences(h5position, targetDatasetName: str, refs: SOAPReferences):
    """Export a reference dictionary as a gzip-9 dataset with ``nmax`` / ``lmax`` / ``names`` attributes
    (classify.py:208-229).  ``h5position`` is an ``h5py.File`` / ``h5py.Group`` (anything with
    ``require_dataset``); host-side storage only, nothing runs on the device."""
    spectra = np.asarray(refs.spectra)
    target = h5position.require_dataset(
        targetDatasetName, shape=spectra.shape, dtype=spectra.dtype, compression="gzip", compression_opts=9
    )
    target[:] = spectra
    for key, value in (("nmax", refs.nmax), ("lmax", refs.lmax), ("names", refs.names)):
        target.attrs.create(key, value)


def getReferencesFromDataset(dataset) -> SOAPReferences:
    """Rebuild the :class:`SOAPReferences` stored by :func:`saveReferences` (classify.py:232-247)."""
    attrs = dataset.attrs
    return SOAPReferences(names=attrs["names"].tolist(), spectra=dataset[:], lmax=attrs["lmax"], nmax=attrs["nmax"])


def pairs_device(data, spectra, kind: int, power: int = 1, path: int = _lib.PAIRS_AUTO, out=None, want_argmin=False,
                 exclude_offset: int = None):
    """Distance matrix of two CUDA tensors ``[M, D]`` x ``[K, D]`` (same float dtype).

    Returns the ``[M, K]`` matrix, or ``(min float64[M], argmin int32[M])`` when ``want_argmin`` (the matrix is then
    written to ``out`` only if one is given).  Large problems take the tcgen05 path with the minimum fused into its
    epilogue; ``exclude_offset`` skips column ``i + exclude_offset`` of row ``i`` (the self distance of a row block)."""
    th = _device.require_cuda()
    lib = _lib.load()
    M, D = (int(s) for s in data.shape)
    K = int(spectra.shape[0])
    if int(spectra.shape[1]) != D:
        raise ValueError("shapes not aligned")
    code = _device.dtype_code(data)
    stream = _device.stream_ptr()
    if want_argmin:
        amin = th.empty(M, dtype=th.float64, device=data.device)
        arg = th.empty(M, dtype=th.int32, device=data.device)
        tensor = (path != _lib.PAIRS_SIMT and M >= 128 and K >= 256 and D >= 32 and M * K >= 1_000_000
                  and kind in (_lib.KIND_SIMPLE, _lib.KIND_KERNEL, _lib.KIND_NORMALIZED))
        if exclude_offset is not None and not tensor:
            raise _lib.SoapB200Error("pairs_device: exclude_offset needs a problem large enough for the tensor-core path")
        if M and K and tensor:
            nbytes = lib.soap_pairs_rowmin_scratch_bytes(M, K, D, code)
            scratch = th.empty(max(int(nbytes), 8), dtype=th.uint8, device=data.device)
            rc = lib.soap_pairs_rowmin(data.data_ptr(), spectra.data_ptr(), None if out is None else out.data_ptr(), M, K, D,
                                       K if out is None else int(out.stride(0)), kind, int(power), code,
                                       _lib.NO_EXCLUDE if exclude_offset is None else int(exclude_offset),
                                       arg.data_ptr(), amin.data_ptr(), scratch.data_ptr(), stream)
            _lib.check(rc, "soap_pairs_rowmin")
        elif M and K:
            rc = lib.soap_pairs(data.data_ptr(), spectra.data_ptr(), None if out is None else out.data_ptr(), M, K, D,
                                K if out is None else int(out.stride(0)), kind, int(power), code,
                                _lib.PAIRS_SIMT, arg.data_ptr(), amin.data_ptr(), None, stream)
            _lib.check(rc, "soap_pairs")
        return amin, arg
    if out is None:
        out = th.empty((M, K), dtype=data.dtype, device=data.device)
    if M and K:
        scratch = None
        if path != _lib.PAIRS_SIMT:
            nbytes = lib.soap_pairs_scratch_bytes(M, K, D, code, path)
            scratch = th.empty(max(int(nbytes), 8), dtype=th.uint8, device=data.device)
        rc = lib.soap_pairs(data.data_ptr(), spectra.data_ptr(), out.data_ptr(), M, K, D, int(out.stride(0)), kind,
                            int(power), code, path, None, None, None if scratch is None else scratch.data_ptr(), stream)
        _lib.check(rc, "soap_pairs")
    return out


def _pairs_fan_out(data, spectra, kind, power, devs):
    """Host arrays: row blocks of ``data`` over the devices (the vector set ``spectra`` replicated), one host
    thread per device, the blocks written into one host matrix."""
    from concurrent.futures import ThreadPoolExecutor

    from .sharding import shard_range

    th = _device.require_cuda()
    a = np.ascontiguousarray(np.asarray(data))
    b = np.ascontiguousarray(np.asarray(spectra, dtype=a.dtype))
    a2, b2 = a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)
    out = np.empty((a2.shape[0], b2.shape[0]), dtype=a.dtype if a.dtype in (np.float32, np.float64) else np.float64)

    def work(i):
        rows = shard_range(a2.shape[0], i, len(devs))
        if rows.stop == rows.start:
            return
        with th.cuda.device(devs[i]):
            da, db = _device.as_float_tensor(a2[rows]), _device.as_float_tensor(b2)
            out[rows] = _device.to_host(pairs_device(da, db.to(da.dtype), kind, power))

    with ThreadPoolExecutor(len(devs)) as pool:
        list(pool.map(work, range(len(devs))))
    return out


def getDistanceBetween(data, spectra, distanceCalculator: Callable, devices=None):
    """Array of the distances between ``data`` and ``spectra`` (classify.py:96-117).

    ``out[i, j] = distanceCalculator(data[i], spectra[j])``, shape ``(len(data), len(spectra))``,
    dtype ``data.dtype``.  ``getDistanceBetween(X, X, f)`` is the all-pairs SOAP distance matrix
    (tensor-core GEMM with the distance as its epilogue when the matrix is large).  ``devices`` (extension, host
    arrays only): row blocks are fanned out over these GPUs, by default over every visible one when the matrix
    has 1e9 entries or more."""
    kind, power = resolve_distance(distanceCalculator)
    th = _device.require_cuda()
    if not _device.is_tensor(data) and not _device.is_tensor(spectra):
        entries = int(np.shape(data)[0]) * int(np.shape(spectra)[0])
        devs = _device.resolve_devices(devices, (1 << 30) if entries >= 10 ** 9 else 0)
        if len(devs) > 1 and int(np.shape(data)[0]) >= 128 * len(devs):
            res = _pairs_fan_out(data, spectra, kind, power, devs)
            want = np.asarray(data).dtype
            return res if res.dtype == want else res.astype(want)
    a = _device.as_float_tensor(data)
    b = _device.as_float_tensor(spectra)
    if a.dtype != b.dtype:
        b = b.to(a.dtype)
    out = pairs_device(a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1), kind, power)
    if _device.is_tensor(data):
        return out
    res = _device.to_host(out)
    want = np.asarray(data).dtype
    return res if res.dtype == want else res.astype(want)  # classify.py:113: zeros(..., dtype=data.dtype)


def _prepared_blocks(SOAPTrajData, references: SOAPReferences, doNormalize: bool, frames: slice = None):
    """Yield ``(lo, hi, rows)``: the frames ``lo:hi`` of the dataset as a CUDA tensor
    ``[(hi-lo)*A, D]``, expanded (mono-species signature, classify.py:151) and normalised as asked."""
    th = _device.require_cuda()
    lib = _lib.load()
    F, A = int(SOAPTrajData.shape[0]), int(SOAPTrajData.shape[1])
    first, last = _frame_interval(frames, F)
    Dref = int(np.asarray(references.spectra).shape[-1])
    Dc = int(SOAPTrajData.shape[-1])
    convert = Dc != Dref
    plan = getExpansionPlan(references.lmax, references.nmax) if convert else None
    if convert and plan.d_compressed != Dc:
        raise ValueError("fillSOAPVectorFromdscribe: the given soap vector do not have the expected dimensions")
    budget_frames = max(1, (4 << 30) // max(1, A * max(Dc, Dref) * 8))
    up = _Uploader(th)
    for lo, hi in frame_ranges(SOAPTrajData, budget_frames):
        lo, hi = max(lo, first), min(hi, last)
        if lo >= hi:
            continue
        raw = up.put_frames(SOAPTrajData, lo, hi)
        rows = raw.reshape(-1, Dc)
        if convert or doNormalize:
            out = th.empty((rows.shape[0], Dref), dtype=rows.dtype, device="cuda")
            if doNormalize:
                rc = lib.soap_expand_normalize(rows.data_ptr(), out.data_ptr(),
                                               plan.index_device().data_ptr() if convert else None,
                                               rows.shape[0], Dc, Dref, _device.dtype_code(rows), _device.stream_ptr())
            else:
                rc = lib.soap_expand(rows.data_ptr(), out.data_ptr(), plan.index_device().data_ptr(), rows.shape[0],
                                     Dc, Dref, rows.element_size(), _device.stream_ptr())
            _lib.check(rc, "expand")
            rows = out
        yield lo, hi, rows


def _frame_interval(frames, n_frames: int):
    """(first, last) of a contiguous frame slice; a strided slice would silently change the result's shape."""
    first, last, step = (frames or slice(None)).indices(n_frames)
    if step != 1:
        raise ValueError("the frame interval must be contiguous (step 1)")
    return first, last


#: dictionaries up to this many references take the fused kernel (expansion folded into the references)
FUSED_MAX_REFS = 32


def _fused_setup(SOAPTrajData, references: SOAPReferences, kind: int):
    """Device-side constants of ``soap_classify`` for this dataset / dictionary, or ``None`` when the
    fused kernel does not apply (big dictionaries, kinds it does not implement)."""
    th = _device.require_cuda()
    spectra = np.asarray(references.spectra, dtype=np.float64)
    K, Dref = spectra.shape
    Dc = int(SOAPTrajData.shape[-1])
    if K > FUSED_MAX_REFS or kind not in (_lib.KIND_SIMPLE, _lib.KIND_KERNEL, _lib.KIND_NORMALIZED):
        return None
    if (K + 1) * Dc * 8 > 200 * 1024:
        return None
    if Dc != Dref:  # compressed dataset: mono-species table (classify.py:151)
        plan = getExpansionPlan(references.lmax, references.nmax)
        if plan.d_compressed != Dc or plan.d_full != Dref:
            raise ValueError("fillSOAPVectorFromdscribe: the given soap vector do not have the expected dimensions")
        folded = np.stack([np.bincount(plan.index, weights=r, minlength=Dc) for r in spectra])
        mult = th.from_numpy(plan.multiplicity.astype(np.float64)).cuda()
    else:
        folded, mult = spectra, None
    return {
        "refw": th.from_numpy(np.ascontiguousarray(folded)).cuda(),
        "norm2": th.from_numpy(np.einsum("kd,kd->k", spectra, spectra)).cuda(),
        "mult": mult, "K": K, "Dc": Dc,
    }


def _fused_blocks(SOAPTrajData, setup, kind, power, doNormalize, want_dist, frames: slice = None):
    """Yield ``(lo, hi, dist[(hi-lo)*A, K] | None, amin, argmin)`` per streamed block, computed by the
    fused kernel straight from the compressed rows."""
    th = _device.require_cuda()
    lib = _lib.load()
    F, A = int(SOAPTrajData.shape[0]), int(SOAPTrajData.shape[1])
    first, last = _frame_interval(frames, F)
    Dc, K = setup["Dc"], setup["K"]
    up = _Uploader(th)
    budget_frames = max(1, (4 << 30) // max(1, A * Dc * 8))
    for lo, hi in frame_ranges(SOAPTrajData, budget_frames):
        lo, hi = max(lo, first), min(hi, last)
        if lo >= hi:
            continue
        rows = up.put_frames(SOAPTrajData, lo, hi).reshape(-1, Dc)
        n = int(rows.shape[0])
        dist = th.empty((n, K), dtype=th.float64, device="cuda") if want_dist else None
        amin = th.empty(n, dtype=th.float64, device="cuda")
        arg = th.empty(n, dtype=th.int32, device="cuda")
        refw, mult, D = setup["refw"], setup["mult"], Dc
        unit = 16 // rows.element_size()
        if Dc % unit and n >= 4096:
            # soap_classify's tiled (TMA) kernel wants rows that are whole 16-byte units; its warp-per-vector
            # fallback is 5-10x slower (profiles/r02_classify_unaligned.txt).  One padding pass on the device --
            # zero columns, zero reference weights -- costs two extra passes over the rows and is still 2-4x faster.
            D = (Dc + unit - 1) // unit * unit
            padded = th.zeros((n, D), dtype=rows.dtype, device=rows.device)
            padded[:, :Dc] = rows
            rows = padded
            if "refw_pad" not in setup:
                setup["refw_pad"] = th.nn.functional.pad(refw, (0, D - Dc)).contiguous()
                setup["mult_pad"] = None if mult is None else th.nn.functional.pad(mult, (0, D - Dc)).contiguous()
            refw, mult = setup["refw_pad"], setup["mult_pad"]
        rc = lib.soap_classify(
            rows.data_ptr(), None if mult is None else mult.data_ptr(), refw.data_ptr(),
            setup["norm2"].data_ptr(), n, D, K, kind, int(power), 1 if doNormalize else 0, _device.dtype_code(rows),
            None if dist is None else dist.data_ptr(), arg.data_ptr(), amin.data_ptr(), _device.stream_ptr(),
        )
        _lib.check(rc, "soap_classify")
        yield lo, hi, dist, amin, arg


def getDistancesFromRef(SOAPTrajData, references: SOAPReferences, distanceCalculator: Callable, doNormalize: bool = False):
    """Distances between a SOAP-hdf5 trajectory and the given references (classify.py:120-160):
    ``float64[F, A, len(references)]`` (numpy)."""
    return getDistancesFromInterval(SOAPTrajData, references, distanceCalculator, doNormalize)


def getDistancesFromInterval(SOAPTrajData, references: SOAPReferences, distanceCalculator: Callable,
                             doNormalize: bool = False, interval: slice = None):
    """:func:`getDistancesFromRef` restricted to the frames ``interval`` (default: all).  North-star
    name with no reference symbol; same arithmetic as classify.py:120-160."""
    kind, power = resolve_distance(distanceCalculator)
    th = _device.require_cuda()
    F, A = int(SOAPTrajData.shape[0]), int(SOAPTrajData.shape[1])
    first, last = _frame_interval(interval, F)
    K = len(references)
    result = np.zeros((max(last - first, 0), A, K))
    setup = _fused_setup(SOAPTrajData, references, kind)
    if setup is not None:
        for lo, hi, dist, _amin, _arg in _fused_blocks(SOAPTrajData, setup, kind, power, doNormalize, True, interval):
            result[lo - first : hi - first] = _device.to_host(dist).reshape(hi - lo, A, K)
        return result
    refs = None
    for lo, hi, rows in _prepared_blocks(SOAPTrajData, references, doNormalize, interval):
        if refs is None or refs.dtype != rows.dtype:
            refs = _device.to_device(np.asarray(references.spectra), dtype=rows.dtype)
        d = pairs_device(rows, refs, kind, power)
        result[lo - first : hi - first] = _device.to_host(d.to(th.float64)).reshape(hi - lo, A, K)
    return result


def getDistancesFromRefNormalized(SOAPTrajData, references: SOAPReferences):
    """Shortcut for :func:`getDistancesFromRef` forcing normalisation (classify.py:163-182)."""
    return getDistancesFromRef(SOAPTrajData, references, SOAPdistanceNormalized, doNormalize=True)


def applyClassification(SOAPTrajData, references: SOAPReferences, distanceCalculator: Callable, doNormalize: bool = False) -> SOAPclassification:
    """Classify every atom of every frame by its closest reference (classify.py:250-276).

    The ``[F, A, K]`` distance array is never materialised: the pair kernel keeps the running
    ``amin`` / ``argmin`` per row (numpy's NaN and tie rules)."""
    kind, power = resolve_distance(distanceCalculator)
    th = _device.require_cuda()
    F, A = int(SOAPTrajData.shape[0]), int(SOAPTrajData.shape[1])
    dist = np.zeros((F, A))
    ids = np.zeros((F, A), dtype=np.int64)
    setup = _fused_setup(SOAPTrajData, references, kind)
    if setup is not None:
        for lo, hi, _dist, amin, arg in _fused_blocks(SOAPTrajData, setup, kind, power, doNormalize, False):
            dist[lo:hi] = _device.to_host(amin).reshape(hi - lo, A)
            ids[lo:hi] = _device.to_host(arg).reshape(hi - lo, A)
        return SOAPclassification(dist, ids, references.names)
    refs = None
    for lo, hi, rows in _prepared_blocks(SOAPTrajData, references, doNormalize):
        if refs is None or refs.dtype != rows.dtype:
            refs = _device.to_device(np.asarray(references.spectra), dtype=rows.dtype)
        amin, arg = pairs_device(rows, refs, kind, power, want_argmin=True)
        dist[lo:hi] = _device.to_host(amin).reshape(hi - lo, A)
        ids[lo:hi] = _device.to_host(arg).reshape(hi - lo, A)
    return SOAPclassification(dist, ids, references.names)
