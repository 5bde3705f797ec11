"""Mirror of ``SOAPify.distances`` (``src/SOAPify/distances.py``): the SOAP distance family.

The five functions keep the reference's names and signatures.  Called on two vectors they run
the pair kernel on the GPU (a 1 x 1 distance matrix -- correct, but meant for spot checks: the
throughput paths are ``timeSOAP`` / ``getDistanceBetween``, which recognise these functions BY
IDENTITY and run the matching fused kernel instead of calling them per pair).  Any other callable
is refused there with ``NotImplementedError``: there is no CPU fallback.
"""
from __future__ import annotations

import functools

from . import _device, _lib

__all__ = [
    "simpleKernelSoap",
    "simpleSOAPdistance",
    "kernelSoap",
    "SOAPdistance",
    "SOAPdistanceNormalized",
    "kernelSOAPdistance",
]


def _pair(x, y, kind: int, power: int = 1):
    th = _device.require_cuda()
    lib = _lib.load()
    a = _device.as_float_tensor(x).reshape(1, -1)
    b = _device.as_float_tensor(y).reshape(1, -1)
    if a.shape != b.shape:
        raise ValueError("shapes not aligned")
    if a.dtype != b.dtype:
        a, b = a.to(th.float64), b.to(th.float64)
    out = th.empty((1, 1), dtype=a.dtype, device=a.device)
    rc = lib.soap_pairs(
        a.data_ptr(), b.data_ptr(), out.data_ptr(), 1, 1, a.shape[1], 1, kind, int(power),
        _device.dtype_code(a), _lib.PAIRS_SIMT, None, None, None, _device.stream_ptr(),
    )
    _lib.check(rc, "soap_pairs")
    if _device.is_tensor(x):
        return out.reshape(())
    return _device.to_host(out)[0, 0]  # numpy scalar of the input precision, like the reference


def simpleKernelSoap(x, y):
    """a simpler SOAP Kernel than :func:`kernelSoap`, power is always 1 (distances.py:7-19)."""
    return _pair(x, y, _lib.KIND_COSCLIP)


def simpleSOAPdistance(x, y):
    """a simpler SOAP distance than :func:`SOAPdistance`, power is always 1 (distances.py:22-35)."""
    return _pair(x, y, _lib.KIND_SIMPLE)


def kernelSoap(x, y, n: int):
    """The SOAP Kernel with a variable power (distances.py:38-51)."""
    return _pair(x, y, _lib.KIND_DOT, n)


def SOAPdistance(x, y, n: int = 1):
    """the SOAP distance between two SOAP fingerprints (distances.py:54-68)."""
    return _pair(x, y, _lib.KIND_KERNEL, n)


def SOAPdistanceNormalized(x, y):
    """the SOAP distance between two normalized SOAP fingerprints (distances.py:71-83)."""
    return _pair(x, y, _lib.KIND_NORMALIZED)


def kernelSOAPdistance(x, y, n: int = 1):
    """North-star alias of :func:`SOAPdistance` (no reference symbol has this name)."""
    return SOAPdistance(x, y, n)


# ---- dispatch of ``distanceFunction`` / ``distanceCalculator`` callables -----------------------
_KINDS = {
    simpleSOAPdistance: (_lib.KIND_SIMPLE, 1),
    SOAPdistance: (_lib.KIND_KERNEL, 1),
    kernelSOAPdistance: (_lib.KIND_KERNEL, 1),
    SOAPdistanceNormalized: (_lib.KIND_NORMALIZED, 1),
    simpleKernelSoap: (_lib.KIND_COSCLIP, 1),
}


_BY_NAME = {
    "simpleSOAPdistance": (_lib.KIND_SIMPLE, 1),
    "SOAPdistance": (_lib.KIND_KERNEL, 1),
    "kernelSOAPdistance": (_lib.KIND_KERNEL, 1),
    "SOAPdistanceNormalized": (_lib.KIND_NORMALIZED, 1),
    "simpleKernelSoap": (_lib.KIND_COSCLIP, 1),
    "kernelSoap": (_lib.KIND_DOT, None),  # the power is a required argument (distances.py:38)
}


def _known(fn):
    """(kind, power) of a built-in distance: this module's functions by identity, the reference package's own
    ``SOAPify.distances.*`` (what existing user scripts import) by name + module."""
    try:
        if fn in _KINDS:
            return _KINDS[fn]
    except TypeError:
        return None
    if fn is kernelSoap:
        return _BY_NAME["kernelSoap"]
    mod = getattr(fn, "__module__", "") or ""
    if mod.endswith("SOAPify.distances") or mod.endswith("cpctools_b200.distances"):
        return _BY_NAME.get(getattr(fn, "__name__", ""))
    return None


def resolve_distance(fn) -> "tuple[int, int]":
    """(kind, power) for one of the built-in distances.  Accepted: the functions of this module and of the
    reference's ``SOAPify.distances`` (matched by module + name), wrappers made with ``functools.wraps``
    (followed through ``__wrapped__``) and ``functools.partial`` chains that only bind the power ``n``
    (``kernelSoap`` needs it, as in the reference).  Anything else raises ``NotImplementedError``."""
    power = None
    seen = 0
    base = fn
    while seen < 16:
        seen += 1
        if isinstance(base, functools.partial):
            if base.args or set(base.keywords) - {"n"}:
                raise NotImplementedError(f"unsupported partial distance {fn!r}: only the power n may be bound")
            if power is None and "n" in base.keywords:  # the outermost binding wins, like a call would
                power = int(base.keywords["n"])
            base = base.func
            continue
        if _known(base) is None and hasattr(base, "__wrapped__"):
            base = base.__wrapped__
            continue
        break
    found = _known(base)
    if found is None:
        raise NotImplementedError(
            "cpctools_b200 runs the built-in SOAP distances on the GPU (simpleSOAPdistance, SOAPdistance, "
            "functools.partial(SOAPdistance, n=k), SOAPdistanceNormalized, also when imported from SOAPify.distances "
            "or wrapped with functools.wraps); an arbitrary Python callable would need a CPU loop, which this "
            f"package does not have: {fn!r}"
        )
    kind, dflt = found
    if power is not None and kind not in (_lib.KIND_KERNEL, _lib.KIND_DOT):
        raise NotImplementedError(f"unsupported partial distance {fn!r}: this distance takes no power")
    if power is None:
        if dflt is None:
            raise TypeError("kernelSoap() missing 1 required positional argument: 'n'")
        power = dflt
    return kind, power
