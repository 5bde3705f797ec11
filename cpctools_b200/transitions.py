"""Mirror of the transition-matrix functions of ``SOAPify.transitions``
(``src/SOAPify/transitions/__init__.py:9-97``): the step after the classification.

Only the histogram of state changes is here; residence times and the
``StateTracker`` bookkeeping of ``transitions/tracker.py`` are out of scope.
"""
from __future__ import annotations

import numpy as np

from . import _device, _lib

__all__ = [
    "transitionMatrixFromSOAPClassification",
    "normalizeMatrixByRow",
    "transitionMatrixFromSOAPClassificationNormalized",
]


def transitionMatrixFromSOAPClassification(data, stride: int = 1, window: "int|None" = None) -> "np.ndarray[float]":
    """Generates the unnormalized matrix of the transitions (transitions/__init__.py:9-48).

    ``M[c_from, c_to]`` counts, over the frames ``range(window, nframes, stride)`` and all atoms, the
    atoms in class ``c_from`` at ``frame - window`` and in ``c_to`` at ``frame``.  float64, like the
    reference; the counting itself is integer and exact."""
    if window is None:
        window = stride
    if window < stride:
        raise ValueError("the window must be bigger than the stride")
    refs = data.references
    if window > refs.shape[0]:
        raise ValueError("stride and window must be smaller than simulation lenght")
    th = _device.require_cuda()
    n_frames, n_atoms = int(refs.shape[0]), int(refs.shape[1])
    n_classes = len(data.legend)
    states = _device.to_device(refs)
    if states.dtype != th.int32:
        states = states.to(th.int32)
    counts = th.empty((n_classes, n_classes), dtype=th.int64, device=states.device)
    if n_classes and n_atoms and n_frames:
        rc = _lib.load().soap_transition_counts(
            states.data_ptr(), n_frames, n_atoms, n_classes, int(window), int(stride), counts.data_ptr(), _device.stream_ptr()
        )
        _lib.check(rc, "soap_transition_counts")
    else:
        counts.zero_()
    return _device.to_host(counts).astype(np.float64)


def normalizeMatrixByRow(transMat: "np.ndarray[float]") -> "np.ndarray[float]":
    """Normalizes a transition matrix by row; empty rows stay zero (transitions/__init__.py:51-69).
    A class-by-class matrix: host arithmetic."""
    out = np.array(transMat, dtype=np.float64, copy=True)
    sums = out.sum(axis=1)
    nz = sums != 0
    out[nz] /= sums[nz, None]
    return out


def transitionMatrixFromSOAPClassificationNormalized(data, stride: int = 1, window: "int|None" = None):
    """Row-normalised transition matrix (transitions/__init__.py:72-97)."""
    return normalizeMatrixByRow(transitionMatrixFromSOAPClassification(data, stride, window))
