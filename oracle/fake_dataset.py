"""In-memory stand-in for the ``h5py.Dataset`` the reference streams from.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  h5py is not installed here or on
the GPU box; ``getTimeSOAPSimple`` (analysis.py:182-190) and ``getDistancesFromRef``
(classify.py:140-149) touch only ``.shape``, ``.attrs``, ``.chunks``, ``.iter_chunks()`` and
slicing, which is what this class offers.  The attrs follow what
``saponify.py:44-69`` writes (``l_max``, ``n_max``, ``species``,
``species_location_A-B`` as [start, stop]).
"""
from __future__ import annotations

import numpy as np


class FakeSOAPDataset:
    def __init__(self, data: np.ndarray, chunk_frames: int, attrs: dict):
        self._data = np.asarray(data)
        self.shape = self._data.shape
        self.dtype = self._data.dtype
        self.attrs = dict(attrs)
        self.chunks = (int(chunk_frames),) + tuple(self.shape[1:])

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, key):
        return np.array(self._data[key])  # h5py returns a fresh ndarray

    def read_direct(self, dest, source_sel=None, dest_sel=None):
        """h5py.Dataset.read_direct: read the selection straight into ``dest`` (no intermediate array)"""
        src = self._data if source_sel is None else self._data[source_sel]
        if dest_sel is None:
            np.copyto(dest, src)
        else:
            np.copyto(dest[dest_sel], src)

    def iter_chunks(self):
        step = self.chunks[0]
        for lo in range(0, self.shape[0], step):
            hi = min(self.shape[0], lo + step)
            yield (slice(lo, hi, 1),) + tuple(slice(0, n, 1) for n in self.shape[1:])


def dscribe_attrs(l_max: int, n_max: int, species: "list[str]") -> "tuple[dict, int]":
    """attrs for a dscribe dataset whose pair blocks are stored (i<=j) in the given species
    order: same-species blocks upper-triangular, cross-species blocks full
    (tests/test_SOAPifySOAP.py:146-156 pins this layout for H/O).  Returns (attrs, Dc)."""
    upper = (l_max + 1) * n_max * (n_max + 1) // 2
    full = (l_max + 1) * n_max * n_max
    attrs = {"l_max": l_max, "n_max": n_max, "species": list(species)}
    pos = 0
    for i, a in enumerate(species):
        for b in species[i:]:
            size = upper if a == b else full
            attrs[f"species_location_{a}-{b}"] = [pos, pos + size]
            if a != b:
                attrs[f"species_location_{b}-{a}"] = [pos, pos + size]
            pos += size
    return attrs, pos


def synthetic_soap(shape, seed=12345, kind="U", dtype=np.float64):
    """Seeded synthetic compressed SOAP data.
    ``U``: uniform [0,1) (the reference tests' generator/seed, tests/test_SOAPifySOAP.py:340);
    ``W``: per-atom random walk ``X[f] = |X[f-1] + 0.02 N(0,1)|`` so that lag-1 distances
    are O(1e-2) like a real trajectory."""
    rng = np.random.default_rng(seed)
    if kind == "U":
        return rng.random(shape).astype(dtype)
    x = np.empty(shape, dtype=np.float64)
    x[0] = rng.random(shape[1:])
    for f in range(1, shape[0]):
        x[f] = np.abs(x[f - 1] + 0.02 * rng.standard_normal(shape[1:]))
    return x.astype(dtype)
