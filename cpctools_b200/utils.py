"""Host-side mirror of ``SOAPify.utils`` for the post-processing hot path.

Same names, argument meaning and error behaviour as ``src/SOAPify/utils.py`` (file:line cited
per function, relative to the reference tree).  Index tables are built on the host (tiny, cached),
everything that touches the data runs in ``libsoapb200.so`` on the GPU.
"""
from __future__ import annotations

import threading
from collections import OrderedDict
from dataclasses import dataclass

import numpy as np

from . import _device, _lib
from ._elements import atomic_numbers

__all__ = [
    "orderByZ",
    "getdscribeSOAPMapping",
    "getquippySOAPMapping",
    "getAddressesQuippyLikeDscribe",
    "normalizeArray",
    "getSlicesFromAttrs",
    "fillSOAPVectorFromdscribe",
    "fillSOAPVectorFromQuip",
    "fillAndNormalize",
    "getSOAPSettings",
    "getExpansionPlan",
]


# ------------------------------------------------------------------------------------------
# layouts (utils.py:9-137) -- pure host code
# ------------------------------------------------------------------------------------------
def orderByZ(species: "list[str]") -> "list[str]":
    """Orders the list of species by their atomic number (utils.py:100-109)."""
    return sorted(species, key=lambda s: atomic_numbers[s])


def _SOAPpstr(l, Z, n, Zp, np_) -> str:
    # utils.py:9-13: heavier species first
    if atomic_numbers[Z] < atomic_numbers[Zp]:
        Z, n, Zp, np_ = Zp, np_, Z, n
    return f"{l}_{Z}{n}_{Zp}{np_}"


def getdscribeSOAPMapping(lmax: int, nmax: int, species: "list[str]", crossover: bool = True) -> np.ndarray:
    """Labels of the slots as dscribe orders them (utils.py:16-48)."""
    species = orderByZ(species)
    labels = []
    for Z in species:
        zZ = atomic_numbers[Z]
        for Zp in species:
            if not crossover and Z != Zp:
                continue
            zZp = atomic_numbers[Zp]
            for l in range(lmax + 1):
                for n in range(nmax):
                    # tuple order (n', Z') >= (n, Z): n' > n, or n' == n and Z' >= Z
                    first = n if zZp >= zZ else n + 1
                    labels.extend(_SOAPpstr(l, Z, n, Zp, m) for m in range(first, nmax))
    return np.array(labels)


def _getRSindex(nmax: int, species: "list[str]") -> np.ndarray:
    """Support function for quippy: row 0 radial index, row 1 species index (utils.py:51-59)."""
    rs = np.zeros((2, nmax * len(species)), dtype=np.int32)
    for s in range(len(species)):
        rs[0, s * nmax : (s + 1) * nmax] = np.arange(nmax)
        rs[1, s * nmax : (s + 1) * nmax] = s
    return rs


def getquippySOAPMapping(lmax: int, nmax: int, species: "list[str]", diagonalRadial: bool = False) -> np.ndarray:
    """Labels of the slots as quippy orders them (utils.py:62-97)."""
    species = orderByZ(species)
    rs = _getRSindex(nmax, species)
    labels = []
    for ia in range(rs.shape[1]):
        npr, Zp = int(rs[0, ia]), species[rs[1, ia]]
        for jb in range(ia + 1):
            n, Z = int(rs[0, jb]), species[rs[1, jb]]
            labels.extend(_SOAPpstr(l, Z, n, Zp, npr) for l in range(lmax + 1))
    return np.array(labels)


def getAddressesQuippyLikeDscribe(lmax: int, nmax: int, species: "list[str]") -> np.ndarray:
    """Indexes that reorder quippy output in the dscribe fashion (utils.py:112-137)."""
    species = orderByZ(species)
    position = {}
    for i, label in enumerate(getquippySOAPMapping(lmax, nmax, species).tolist()):
        position.setdefault(label, i)
    return np.fromiter(
        (position[label] for label in getdscribeSOAPMapping(lmax, nmax, species).tolist()), dtype=int
    )


# ------------------------------------------------------------------------------------------
# gather tables (utils.py:182-268) and the cached expansion plan
# ------------------------------------------------------------------------------------------
def _getIndexesForFillSOAPVectorFromdscribeSameSpecies(lMax: int, nMax: int) -> np.ndarray:
    """Symmetric-fill table of one same-species block (utils.py:182-209)."""
    table = np.empty((lMax + 1, nMax, nMax), dtype=int)
    slot = 0
    for l in range(lMax + 1):
        for n in range(nMax):
            run = np.arange(slot, slot + nMax - n)
            table[l, n, n:] = run
            table[l, n:, n] = run
            slot += nMax - n
    return table.reshape(-1)


def _getIndexesForFillSOAPVectorFromdscribe(lMax, nMax, atomTypes=None, atomicSlices=None) -> np.ndarray:
    """Gather table for every species pair (utils.py:212-268)."""
    if atomTypes is None:
        atomTypes = [None]
    if list(atomTypes) == [None]:
        return _getIndexesForFillSOAPVectorFromdscribeSameSpecies(lMax, nMax)
    full = (lMax + 1) * nMax * nMax
    npairs = len(atomTypes) * (len(atomTypes) + 1) // 2
    out = np.zeros(full * npairs, dtype=int)
    same = _getIndexesForFillSOAPVectorFromdscribeSameSpecies(lMax, nMax)
    pos = 0
    for i, s1 in enumerate(atomTypes):
        for s2 in atomTypes[i:]:
            start = atomicSlices[s1 + s2].start
            out[pos : pos + full] = (same if s1 == s2 else np.arange(full, dtype=int)) + start
            pos += full
    return out


@dataclass
class ExpansionPlan:
    """Everything the kernels need to expand one compressed layout (built once, cached)."""

    index: np.ndarray  #: int64[Df] gather table (bit-exact with the reference's)
    d_compressed: int  #: length the compressed vectors must have (row stride of the input)
    d_full: int
    multiplicity: np.ndarray  #: int64[d_compressed]: times each compressed slot is referenced
    _dev: dict

    def index_device(self):
        th = _device.require_cuda()
        key = ("idx", th.cuda.current_device())
        if key not in self._dev:
            self._dev[key] = th.from_numpy(self.index.astype(np.int32)).cuda()
        return self._dev[key]

    def multiplicity_device(self, dtype):
        th = _device.require_cuda()
        key = ("mult", dtype, th.cuda.current_device())
        if key not in self._dev:
            self._dev[key] = th.from_numpy(self.multiplicity.astype(np.float64)).to(dtype).cuda()
        return self._dev[key]


_PLAN_CACHE: "OrderedDict[tuple, ExpansionPlan]" = OrderedDict()
_PLAN_LOCK = threading.Lock()
_PLAN_CACHE_SIZE = 64


def _expected_compressed_dim(lMax, nMax, n_species) -> int:
    upperDiag = int((lMax + 1) * nMax * (nMax + 1) / 2)
    fullmat = nMax * nMax * (lMax + 1)
    return upperDiag * n_species + fullmat * int((n_species - 1) * n_species / 2)


def _normalise_types(atomTypes):
    if atomTypes is None:
        return [None]
    types = [None if t is None else str(t) for t in list(atomTypes)]
    if len(types) > 1:  # utils.py:304-306
        types.sort(key=lambda s: atomic_numbers[s])
    return types


def getExpansionPlan(lMax, nMax, atomTypes=None, atomicSlices=None, quippy: bool = False) -> ExpansionPlan:
    """Cached gather table (+ multiplicities) for ``fillSOAPVectorFromdscribe`` arguments.

    ``quippy=True`` composes the table with the quippy->dscribe permutation
    (``getAddressesQuippyLikeDscribe``, applied by the reference at ``engine.py:260``) so that RAW
    quippy rows (with their trailing ``covariance_sigma0`` column, ``engine.py:226``) are expanded
    by ONE gather; the input row stride is then ``Dc + 1``.
    """
    lMax, nMax = int(lMax), int(nMax)
    types = _normalise_types(atomTypes)
    slices_key = None
    if atomicSlices is not None:
        slices_key = tuple(sorted((k, int(v.start), int(v.stop)) for k, v in atomicSlices.items()))
    key = (lMax, nMax, tuple(types), slices_key, bool(quippy))
    with _PLAN_LOCK:
        plan = _PLAN_CACHE.get(key)
        if plan is not None:
            _PLAN_CACHE.move_to_end(key)
            return plan
    index = _getIndexesForFillSOAPVectorFromdscribe(lMax, nMax, types, atomicSlices)
    d_comp = _expected_compressed_dim(lMax, nMax, len(types))
    if quippy:
        species = [t for t in types if t is not None]
        addr = getAddressesQuippyLikeDscribe(lMax, nMax, species)
        if len(addr) != d_comp:
            raise ValueError("fillSOAPVectorFromQuip: species do not match the compressed dimension")
        index = addr[index]
        d_comp += 1  # raw quippy rows carry one trailing column that is never referenced
    if index.size and (index.min() < 0 or index.max() >= d_comp):
        raise ValueError("fillSOAPVectorFromdscribe: atomicSlices point outside the given soap vector")
    plan = ExpansionPlan(
        index=index,
        d_compressed=d_comp,
        d_full=int(index.size),
        multiplicity=np.bincount(index, minlength=d_comp).astype(np.int64),
        _dev={},
    )
    with _PLAN_LOCK:
        _PLAN_CACHE[key] = plan
        while len(_PLAN_CACHE) > _PLAN_CACHE_SIZE:
            _PLAN_CACHE.popitem(last=False)
    return plan


# ------------------------------------------------------------------------------------------
# device entry points
# ------------------------------------------------------------------------------------------
def _expand(x, plan: ExpansionPlan, normalize: bool, who: str):
    if len(x.shape) > 3 or len(x.shape) < 1:
        raise ValueError(f"{who}: cannot convert array with len(shape) >=3")
    if x.shape[-1] != plan.d_compressed:
        raise ValueError(f"{who}: the given soap vector do not have the expected dimensions")
    th = _device.require_cuda()
    lib = _lib.load()
    restore = None  # numpy dtypes the gather kernel does not move natively (1- and 2-byte elements)
    if not normalize and not _device.is_tensor(x):
        arr = np.asarray(x)
        if arr.dtype.kind == "u" and arr.dtype.itemsize in (4, 8):
            restore = ("view", arr.dtype)  # a gather moves bits: reinterpret as the signed type
            x = np.ascontiguousarray(arr).view(np.int32 if arr.dtype.itemsize == 4 else np.int64)
        elif arr.dtype.itemsize not in (4, 8) or arr.dtype.kind not in "fi":
            restore = ("cast", arr.dtype)
            x = arr.astype(np.float32 if arr.dtype.kind == "f" else np.int64)
    src = _device.as_float_tensor(x) if normalize else _device.to_device(x)
    if src.dtype == th.bool or src.element_size() not in (4, 8):
        src = src.to(th.int64 if not src.dtype.is_floating_point else th.float32)
    nvec = src.numel() // plan.d_compressed if plan.d_compressed else 0
    out = th.empty(src.shape[:-1] + (plan.d_full,), dtype=src.dtype, device=src.device)
    idx = plan.index_device()
    if normalize:
        rc = lib.soap_expand_normalize(
            src.data_ptr(), out.data_ptr(), idx.data_ptr(), nvec, plan.d_compressed, plan.d_full,
            _device.dtype_code(src), _device.stream_ptr(),
        )
    else:
        rc = lib.soap_expand(
            src.data_ptr(), out.data_ptr(), idx.data_ptr(), nvec, plan.d_compressed, plan.d_full,
            src.element_size(), _device.stream_ptr(),
        )
    _lib.check(rc, who)
    res = _device.like_input(out, x)
    if restore is None:
        return res
    return res.view(restore[1]) if restore[0] == "view" else res.astype(restore[1])


def fillSOAPVectorFromdscribe(soapFromdscribe, lMax: int, nMax: int, atomTypes: list = None, atomicSlices: dict = None):
    """SOAP power spectrum with the symmetric part explicitly stored (utils.py:271-324).

    ``x[..., idx]`` on the last axis of a 1/2/3-D array; dtype preserved; ``ValueError`` on a wrong
    last dimension or on more than three dimensions.  The gather is bit-exact.
    """
    plan = getExpansionPlan(lMax, nMax, atomTypes, atomicSlices)
    return _expand(soapFromdscribe, plan, False, "fillSOAPVectorFromdscribe")


def fillSOAPVectorFromQuip(soapFromQuip, lMax: int, nMax: int, atomTypes: list, atomicSlices: dict = None):
    """Expand RAW quippy output (``[..., Dc+1]``, quippy slot order, trailing covariance column).

    No reference symbol has this name: it is
    ``fillSOAPVectorFromdscribe(x[..., :Dc][..., getAddressesQuippyLikeDscribe(...)], ...)``
    (``engine.py:226,260`` followed by ``utils.py:271``) as ONE composed gather.  ``atomicSlices``
    defaults to the layout ``quippySOAPengineContainer`` builds (``engine.py:203-222``).
    """
    species = orderByZ([str(s) for s in atomTypes])
    if atomicSlices is None:
        atomicSlices = _quippyContainerSlices(lMax, nMax, species)
    plan = getExpansionPlan(lMax, nMax, species, atomicSlices, quippy=True)
    return _expand(soapFromQuip, plan, False, "fillSOAPVectorFromQuip")


def _quippyContainerSlices(lMax, nMax, species):
    """Slices of the (already permuted) quippy vector, pair blocks i<=j (engine.py:203-222)."""
    full = nMax * nMax * (lMax + 1)
    upper = ((nMax + 1) * nMax) // 2 * (lMax + 1)
    slices, pos = {}, 0
    for i, a in enumerate(species):
        for b in species[i:]:
            size = upper if a == b else full
            slices[a + b] = slice(pos, pos + size)
            pos += size
    return slices


def normalizeArray(x):
    """Normalizes the further axis of the given array (utils.py:140-153); zero vectors stay zero."""
    th = _device.require_cuda()
    lib = _lib.load()
    src = _device.as_float_tensor(x)
    if src.dim() == 0:
        raise ValueError("normalizeArray: need at least one axis")
    d = src.shape[-1]
    out = th.empty_like(src)
    nvec = src.numel() // d if d else 0
    if nvec:
        rc = lib.soap_expand_normalize(
            src.data_ptr(), out.data_ptr(), None, nvec, d, d, _device.dtype_code(src), _device.stream_ptr()
        )
        _lib.check(rc, "normalizeArray")
    return _device.like_input(out, x)


def fillAndNormalize(soapFromdscribe, lMax: int, nMax: int, atomTypes: list = None, atomicSlices: dict = None,
                     quippy: bool = False):
    """``normalizeArray(fillSOAPVectorFromdscribe(...))`` in one kernel (reads Dc, writes Df once)."""
    if quippy:
        species = orderByZ([str(s) for s in atomTypes])
        if atomicSlices is None:
            atomicSlices = _quippyContainerSlices(lMax, nMax, species)
        plan = getExpansionPlan(lMax, nMax, species, atomicSlices, quippy=True)
    else:
        plan = getExpansionPlan(lMax, nMax, atomTypes, atomicSlices)
    return _expand(soapFromdscribe, plan, True, "fillSOAPVectorFromdscribe")


# ------------------------------------------------------------------------------------------
# dataset attrs (utils.py:156-179, 327-353) -- pure host code
# ------------------------------------------------------------------------------------------
def getSlicesFromAttrs(attrs: dict) -> "tuple(list,dict)":
    """Positional slices of the pair blocks from the attributes of a SOAP dataset (utils.py:156-179)."""
    species = attrs["species"]
    slices = {}
    for s1 in species:
        for s2 in species:
            where = f"species_location_{s1}-{s2}"
            if where in attrs:
                slices[s1 + s2] = slice(attrs[where][0], attrs[where][1])
    return species, slices


def getSOAPSettings(fitsetData) -> dict:
    """Settings of the SOAP calculation, ready for :func:`fillSOAPVectorFromdscribe` (utils.py:327-353)."""
    symbols, atomicSlices = getSlicesFromAttrs(fitsetData.attrs)
    return {
        "nMax": fitsetData.attrs["n_max"],
        "lMax": fitsetData.attrs["l_max"],
        "atomTypes": symbols,
        "atomicSlices": atomicSlices,
    }
