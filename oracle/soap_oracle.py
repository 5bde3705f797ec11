"""numpy/scipy restatement of the reference's SOAP post-processing hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): this module is the checker
for the CUDA path and the timed CPU arm of ``bench.py``; the product package
``cpctools_b200`` never imports it.

Parity status: PINNED.  Every function below is checked (tests/test_oracle_pinned.py)
against (i) the unmodified reference imported through ``oracle/ref_shim.py`` when
``/root/reference`` exists, and (ii) the golden vectors in ``tests/golden/*.npz``
that ``oracle/gen_golden.py`` generated from that same unmodified reference.
What the reference's own tests leave unpinned (lag-``window`` > 1 numerics, see
SURVEY.md section 8c) is pinned here against the reference's ``timeSOAP`` run on
non-degenerate seeded data.

All file:line citations are relative to ``/root/reference/``.  Function names follow
the reference so that parity tests read like the reference's tests.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np
import scipy.spatial.distance as _ssd

from .elements import ATOMIC_NUMBERS

# --------------------------------------------------------------------------------------
# species helpers / layouts (src/SOAPify/utils.py:9-137)
# --------------------------------------------------------------------------------------


def orderByZ(species: Sequence[str]) -> "list[str]":
    """Species sorted by atomic number (utils.py:100-109)."""
    return sorted(species, key=ATOMIC_NUMBERS.__getitem__)


def _label(l: int, za: str, na: int, zb: str, nb: int) -> str:
    """Slot label with the heavier species first (utils.py:9-13)."""
    if ATOMIC_NUMBERS[za] < ATOMIC_NUMBERS[zb]:
        za, na, zb, nb = zb, nb, za, na
    return f"{l}_{za}{na}_{zb}{nb}"


def getdscribeSOAPMapping(lmax, nmax, species, crossover=True) -> np.ndarray:
    """Labels of the dscribe-like layout (utils.py:16-48): loops Z, Z', l, n, n' and keeps
    a slot when ``(n', Z(Z')) >= (n, Z(Z))`` in tuple order."""
    sp = orderByZ(species)
    out = []
    for za in sp:
        for zb in sp:
            if za != zb and not crossover:
                continue
            key_a, key_b = ATOMIC_NUMBERS[za], ATOMIC_NUMBERS[zb]
            for l in range(lmax + 1):
                for na in range(nmax):
                    for nb in range(nmax):
                        if (nb, key_b) >= (na, key_a):
                            out.append(_label(l, za, na, zb, nb))
    return np.array(out)


def _getRSindex(nmax, species) -> np.ndarray:
    """(radial, species) index pairs of quippy's compound index (utils.py:51-59)."""
    ns = len(species)
    rs = np.empty((2, nmax * ns), dtype=np.int32)
    rs[0] = np.tile(np.arange(nmax, dtype=np.int32), ns)
    rs[1] = np.repeat(np.arange(ns, dtype=np.int32), nmax)
    return rs


def getquippySOAPMapping(lmax, nmax, species, diagonalRadial=False) -> np.ndarray:
    """Labels of the quippy layout (utils.py:62-97): compound index ia=(species,n),
    jb<=ia, l fastest."""
    sp = orderByZ(species)
    rs = _getRSindex(nmax, sp)
    out = []
    for ia in range(rs.shape[1]):
        for jb in range(ia + 1):
            for l in range(lmax + 1):
                out.append(
                    _label(l, sp[rs[1, jb]], int(rs[0, jb]), sp[rs[1, ia]], int(rs[0, ia]))
                )
    return np.array(out)


def getAddressesQuippyLikeDscribe(lmax, nmax, species) -> np.ndarray:
    """Permutation that reorders quippy output into the dscribe-like layout
    (utils.py:112-137): ``addr[i]`` = first position of dscribe label ``i`` in the quippy
    labels; applied as ``quippyData[:, addr]`` (engine.py:260)."""
    sp = orderByZ(species)
    where = {}
    for pos, lab in enumerate(getquippySOAPMapping(lmax, nmax, sp)):
        where.setdefault(str(lab), pos)
    dscribe = getdscribeSOAPMapping(lmax, nmax, sp)
    return np.array([where[str(lab)] for lab in dscribe], dtype=int)


# --------------------------------------------------------------------------------------
# expansion (utils.py:182-324) and normalisation (utils.py:140-153)
# --------------------------------------------------------------------------------------


def _getIndexesForFillSOAPVectorFromdscribeSameSpecies(lMax, nMax) -> np.ndarray:
    """Gather table of one same-species block (utils.py:182-209): entry [l][n][n'] is the
    running id of the stored upper-triangular slot (l, min(n,n'), max(n,n'))."""
    per_l = nMax * (nMax + 1) // 2
    tri = np.zeros((nMax, nMax), dtype=int)
    iu = np.triu_indices(nMax)
    tri[iu] = np.arange(per_l)
    # (n,n') below the diagonal takes the id stored at (n',n): tri.T carries it.
    tri = np.where(np.arange(nMax)[:, None] <= np.arange(nMax)[None, :], tri, tri.T)
    table = tri[None, :, :] + per_l * np.arange(lMax + 1)[:, None, None]
    return table.reshape(-1).astype(int)


def _getIndexesForFillSOAPVectorFromdscribe(lMax, nMax, atomTypes=None, atomicSlices=None):
    """Full gather table (utils.py:212-268): pair blocks (s_i, s_j), i<=j, in the given
    species order; same species -> symmetric-fill table + slice start, cross species ->
    verbatim copy of the M-long block."""
    if atomTypes is None:
        atomTypes = [None]
    if list(atomTypes) == [None]:
        return _getIndexesForFillSOAPVectorFromdscribeSameSpecies(lMax, nMax)
    full = (lMax + 1) * nMax * nMax
    same = _getIndexesForFillSOAPVectorFromdscribeSameSpecies(lMax, nMax)
    blocks = []
    for i, sa in enumerate(atomTypes):
        for sb in atomTypes[i:]:
            start = atomicSlices[sa + sb].start
            blocks.append((same if sa == sb else np.arange(full, dtype=int)) + start)
    return np.concatenate(blocks).astype(int)


def _compressed_dim(lMax, nMax, n_species) -> int:
    upper = int((lMax + 1) * nMax * (nMax + 1) / 2)
    full = nMax * nMax * (lMax + 1)
    return upper * n_species + full * int((n_species - 1) * n_species / 2)


def fillSOAPVectorFromdscribe(soapFromdscribe, lMax, nMax, atomTypes=None, atomicSlices=None):
    """Expand the compressed power spectrum on its last axis (utils.py:271-324)."""
    if atomTypes is None:
        atomTypes = [None]
    if len(atomTypes) > 1:
        atomTypes = sorted(list(atomTypes), key=ATOMIC_NUMBERS.__getitem__)
    if soapFromdscribe.shape[-1] != _compressed_dim(lMax, nMax, len(atomTypes)):
        raise ValueError(
            "fillSOAPVectorFromdscribe: the given soap vector do not have the expected dimensions"
        )
    if soapFromdscribe.ndim > 3:
        raise ValueError("fillSOAPVectorFromdscribe: cannot convert array with len(shape) >=3")
    idx = _getIndexesForFillSOAPVectorFromdscribe(lMax, nMax, atomTypes, atomicSlices)
    return soapFromdscribe[..., idx]


def normalizeArray(x: np.ndarray) -> np.ndarray:
    """L2-normalise the last axis; all-zero vectors stay zero (utils.py:140-153)."""
    nrm = np.linalg.norm(x, axis=-1, keepdims=True)
    nrm[nrm == 0] = 1
    return x / nrm


def getSlicesFromAttrs(attrs):
    """attrs -> (species, {"AB": slice}) (utils.py:156-179)."""
    species = attrs["species"]
    slices = {}
    for a in species:
        for b in species:
            key = f"species_location_{a}-{b}"
            if key in attrs:
                slices[a + b] = slice(attrs[key][0], attrs[key][1])
    return species, slices


def getSOAPSettings(dataset) -> dict:
    """kwargs for fillSOAPVectorFromdscribe from dataset attrs (utils.py:327-353)."""
    species, slices = getSlicesFromAttrs(dataset.attrs)
    return {
        "nMax": dataset.attrs["n_max"],
        "lMax": dataset.attrs["l_max"],
        "atomTypes": species,
        "atomicSlices": slices,
    }


# --------------------------------------------------------------------------------------
# distance family (src/SOAPify/distances.py)
# --------------------------------------------------------------------------------------


def simpleKernelSoap(x, y):
    """1 - scipy cosine distance (distances.py:7-19); scipy clips the distance to [0,2]."""
    return 1 - _ssd.cosine(x, y)


def simpleSOAPdistance(x, y):
    """sqrt(2 - 2 k) with the clipped cosine kernel (distances.py:22-35)."""
    try:
        return np.sqrt(2.0 - 2.0 * simpleKernelSoap(x, y))
    except FloatingPointError:
        return 0.0


def kernelSoap(x, y, n):
    """(x.y / (|x| |y|))**n, no clipping (distances.py:38-51)."""
    return (np.dot(x, y) / (np.linalg.norm(x) * np.linalg.norm(y))) ** n


def SOAPdistance(x, y, n=1):
    """sqrt(2 - 2 kernelSoap) (distances.py:54-68); NaN whenever rounding gives k > 1."""
    try:
        return np.sqrt(2.0 - 2.0 * kernelSoap(x, y, n))
    except FloatingPointError:
        return 0.0


def SOAPdistanceNormalized(x, y):
    """sqrt(|2 - 2 x.y|) for unit vectors (distances.py:71-83)."""
    return np.sqrt(np.abs(2.0 - 2.0 * x.dot(y)))


# --------------------------------------------------------------------------------------
# timeSOAP family (src/SOAPify/analysis.py:13-203)
# --------------------------------------------------------------------------------------


def _check_window(n_frames, window, stride):
    if stride is None:
        stride = window
    if stride > window:
        raise ValueError("the window must be bigger than the stride")
    if window >= n_frames or stride >= n_frames:
        raise ValueError("stride and window must be smaller than simulation lenght")


def timeSOAP(
    SOAPTrajectory,
    window=1,
    stride=None,
    backward=False,
    returnDiff=True,
    distanceFunction: Callable = simpleSOAPdistance,
):
    """Generic per-atom lag-``window`` distance, one Python call per (frame, atom)
    (analysis.py:13-69).  Output is always float64."""
    n_frames, n_atoms = SOAPTrajectory.shape[0], SOAPTrajectory.shape[1]
    _check_window(n_frames, window, stride)
    t = np.zeros((n_frames - window, n_atoms))
    for f in range(window, n_frames):
        now, before = SOAPTrajectory[f], SOAPTrajectory[f - window]
        for a in range(n_atoms):
            t[f - window, a] = distanceFunction(now[a, :], before[a, :])
    if returnDiff:
        return t, np.diff(t.T, axis=-1)
    return t


def timeSOAPsimple(SOAPTrajectory, window=1, stride=None, backward=False, returnDiff=True):
    """Per-frame ``|X[f] - prev|`` on normalised data (analysis.py:72-142).  Faithful to the
    reference's quirk: ``prev`` starts at frame 0 and then trails by ONE frame, so for
    window > 1 only row 0 is a lag-``window`` distance."""
    n_frames, n_atoms = SOAPTrajectory.shape[0], SOAPTrajectory.shape[1]
    _check_window(n_frames, window, stride)
    t = np.zeros((n_frames - window, n_atoms))
    prev = SOAPTrajectory[0]
    for f in range(window, n_frames):
        cur = SOAPTrajectory[f]
        t[f - window] = np.linalg.norm(cur - prev, axis=1)
        prev = cur
    if returnDiff:
        return t, np.diff(t.T, axis=-1)
    return t


def getTimeSOAPSimple(soapDataset, window=1, stride=None, backward=False):
    """Stream HDF5 frame-chunks (1-frame halo) through fill -> normalise -> timeSOAPsimple
    (analysis.py:145-203).  Like the reference this only works for window == 1."""
    settings = getSOAPSettings(soapDataset)
    t = np.zeros((soapDataset.shape[0] - window, soapDataset.shape[1]))
    halo = 0
    for chunk in soapDataset.iter_chunks():
        first = chunk[0]
        src = slice(first.start - halo, first.stop, first.step)
        dst = slice(first.start - halo, first.stop - 1, first.step)
        t[dst] = timeSOAPsimple(
            normalizeArray(fillSOAPVectorFromdscribe(soapDataset[src], **settings)),
            window=window,
            stride=stride,
            backward=backward,
            returnDiff=False,
        )
        halo = 1
    return t, np.diff(t.T, axis=-1)


# --------------------------------------------------------------------------------------
# distance matrix / classification (src/SOAPify/classify.py:96-182, 250-276)
# --------------------------------------------------------------------------------------


def getDistanceBetween(data, spectra, distanceCalculator: Callable):
    """out[i, j] = f(data[i], spectra[j]); output dtype = data.dtype (classify.py:96-117)."""
    out = np.zeros((data.shape[0], spectra.shape[0]), dtype=data.dtype)
    for j in range(spectra.shape[0]):
        ref = spectra[j]
        for i in range(data.shape[0]):
            out[i, j] = distanceCalculator(data[i], ref)
    return out


def getDistancesFromRef(SOAPTrajData, references, distanceCalculator, doNormalize=False):
    """Chunked (<=100 frames) fill(mono-species signature) -> normalise -> per-frame
    distance matrix against ``references.spectra`` (classify.py:120-160)."""
    step = min(100, SOAPTrajData.chunks[0])
    n_frames, n_atoms = SOAPTrajData.shape[0], SOAPTrajData.shape[1]
    convert = SOAPTrajData.shape[-1] != references.spectra.shape[-1]
    out = np.zeros((n_frames, n_atoms, len(references.spectra)))
    for lo in range(0, n_frames, step):
        frames = SOAPTrajData[lo : min(n_frames, lo + step)]
        if convert:
            frames = fillSOAPVectorFromdscribe(frames, references.lmax, references.nmax)
        if doNormalize:
            frames = normalizeArray(frames)
        for k, frame in enumerate(frames):
            out[lo + k] = getDistanceBetween(frame, references.spectra, distanceCalculator)
    return out


def getDistancesFromRefNormalized(SOAPTrajData, references):
    """classify.py:163-182."""
    return getDistancesFromRef(SOAPTrajData, references, SOAPdistanceNormalized, doNormalize=True)


def applyClassification(SOAPTrajData, references, distanceCalculator, doNormalize=False):
    """argmin / amin of the distances over the references (classify.py:250-276).
    Returns (distances[F,A], references int[F,A], legend)."""
    d = getDistancesFromRef(SOAPTrajData, references, distanceCalculator, doNormalize)
    return np.amin(d, axis=-1), np.argmin(d, axis=-1), references.names


# --------------------------------------------------------------------------------------
# transition matrix of a classification (src/SOAPify/transitions/__init__.py:9-97)
# --------------------------------------------------------------------------------------


def transitionMatrixFromSOAPClassification(references, n_classes, stride=1, window=None):
    """Counts of (state[f-window], state[f]) over f in range(window, F, stride) and all atoms
    (transitions/__init__.py:9-48); ``references`` is the int[F, A] array of a SOAPclassification."""
    if window is None:
        window = stride
    if window < stride:
        raise ValueError("the window must be bigger than the stride")
    if window > references.shape[0]:
        raise ValueError("stride and window must be smaller than simulation lenght")
    mat = np.zeros((n_classes, n_classes), dtype=np.float64)
    for f in range(window, references.shape[0], stride):
        np.add.at(mat, (references[f - window], references[f]), 1)
    return mat


def calculateResidenceTimesFromClassification(references, n_classes, window=1, stride=None):
    """Sorted residence times per state (transitions/__init__.py:99-167).  For every atom and every
    initial frame in range(0, window, stride) the states sampled every ``window`` frames are run-length
    encoded: a run that ends contributes its length in frames, negated if it is the first run of the
    series; the run still open at the end contributes ``-(length + window)``."""
    if stride is None:
        stride = window
    if stride > window:
        raise ValueError("the window must be bigger than the stride")
    n_frames, n_atoms = references.shape
    if window > n_frames or stride > n_frames:
        raise ValueError("stride and window must be smaller than simulation lenght")
    per_state = [[] for _ in range(n_classes)]
    for atom in range(n_atoms):
        series = references[:, atom]
        for start in range(0, window, stride):
            current, elapsed, first_run = series[start], 0, True
            for frame in range(start + window, n_frames, window):
                elapsed += window
                if series[frame] != current:
                    per_state[current].append(-elapsed if first_run else elapsed)
                    first_run = False
                    current, elapsed = series[frame], 0
            per_state[current].append(-elapsed - window)
    return [np.sort(np.array(times)) for times in per_state]


def normalizeMatrixByRow(transMat):
    """Rows summing to 1, empty rows left at zero (transitions/__init__.py:51-69)."""
    out = transMat.copy()
    for row in range(out.shape[0]):
        total = np.sum(out[row, :])
        if total != 0:
            out[row, :] /= total
    return out


# --------------------------------------------------------------------------------------
# vectorised equivalents (same arithmetic per pair as the loops above, batched) -- used by
# the parity tests at sizes the Python loops cannot reach.  Each is checked against the
# loop version in tests/test_oracle_pinned.py.
# --------------------------------------------------------------------------------------

KIND_SIMPLE, KIND_KERNEL, KIND_NORMALIZED = 0, 1, 2


def _epilogue(uv, uu, vv, kind, power):
    """Pair epilogue shared by the batched oracles; mirrors the scalar functions above."""
    with np.errstate(all="ignore"):
        if kind == KIND_SIMPLE:
            cosd = np.clip(1.0 - uv / np.sqrt(uu * vv), 0.0, 2.0)
            return np.sqrt(2.0 - 2.0 * (1 - cosd))
        if kind == KIND_KERNEL:
            return np.sqrt(2.0 - 2.0 * (uv / (np.sqrt(uu) * np.sqrt(vv))) ** power)
        if kind == KIND_NORMALIZED:
            return np.sqrt(np.abs(2.0 - 2.0 * uv))
    raise ValueError(kind)


def timeSOAP_batched(X, window=1, kind=KIND_SIMPLE, power=1, returnDiff=True):
    """Batched ``timeSOAP`` for the three built-in distances (float64 accumulate)."""
    _check_window(X.shape[0], window, None)
    a = np.asarray(X[window:], dtype=np.float64)
    b = np.asarray(X[:-window], dtype=np.float64)
    uv = np.einsum("fad,fad->fa", a, b)
    uu = np.einsum("fad,fad->fa", a, a)
    vv = np.einsum("fad,fad->fa", b, b)
    t = _epilogue(uv, uu, vv, kind, power)
    if returnDiff:
        return t, np.diff(t.T, axis=-1)
    return t


def getDistanceBetween_batched(data, spectra, kind=KIND_NORMALIZED, power=1):
    """Batched distance matrix (float64 accumulate), cast to ``data.dtype`` like the loop."""
    a = np.asarray(data, dtype=np.float64)
    b = np.asarray(spectra, dtype=np.float64)
    uv = a @ b.T
    uu = np.einsum("id,id->i", a, a)[:, None]
    vv = np.einsum("jd,jd->j", b, b)[None, :]
    return _epilogue(uv, uu, vv, kind, power).astype(data.dtype, copy=False)
