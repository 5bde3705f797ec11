"""Import the UNMODIFIED reference (``/root/reference/src/SOAPify``) in a container
that lacks ase / h5py / MDAnalysis / dscribe / quippy.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Used by
``oracle/gen_golden.py`` (to produce ``tests/golden/*.npz``) and by the
``-m "not gpu"`` oracle-pinning tests, which skip when ``/root/reference`` is
absent (it does not exist on the GPU box).

The hot-path modules touch the missing packages only for
``ase.data.atomic_numbers`` / ``chemical_symbols`` (``utils.py:5``,
``engine.py:5``) and for type annotations, so empty stub modules carrying those
two tables are enough.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

from .elements import ATOMIC_NUMBERS, CHEMICAL_SYMBOLS

REFERENCE_SRC = "/root/reference/src"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "SOAPify"))


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__dict__["__graft_stub__"] = True
        sys.modules[name] = mod
    for key, val in attrs.items():
        setattr(mod, key, val)
    return mod


def _install_stubs() -> None:
    class _Opaque:  # stands in for classes only used in annotations
        def __init__(self, *a, **k):
            raise RuntimeError("stubbed class: not available in this container")

    def _unavailable(*a, **k):
        raise RuntimeError("stubbed function: not available in this container")

    def need(name):
        try:
            importlib.import_module(name)
            return False
        except Exception:
            return True

    if need("ase"):
        ase = _stub("ase", Atoms=_Opaque)
        ase.data = _stub(
            "ase.data",
            atomic_numbers=dict(ATOMIC_NUMBERS),
            chemical_symbols=list(CHEMICAL_SYMBOLS),
        )
        ase.io = _stub("ase.io", read=_unavailable, iread=_unavailable)
        ase.build = _stub("ase.build")
    if need("h5py"):
        _stub("h5py", Dataset=_Opaque, Group=_Opaque, File=_Opaque)
    if need("MDAnalysis"):
        mda = _stub("MDAnalysis", Universe=_Opaque, AtomGroup=_Opaque)
        mda.lib = _stub("MDAnalysis.lib")
        mda.lib.NeighborSearch = _stub(
            "MDAnalysis.lib.NeighborSearch", AtomNeighborSearch=_Opaque
        )
        mda.lib.mdamath = _stub(
            "MDAnalysis.lib.mdamath",
            triclinic_vectors=_unavailable,
            triclinic_box=_unavailable,
        )


def load_reference():
    """Return the reference ``SOAPify`` package (unmodified source), or raise."""
    if not reference_available():
        raise ImportError(f"{REFERENCE_SRC}/SOAPify not present")
    _install_stubs()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module("SOAPify")
