"""Mirror of the classification post-processing of ``SOAPify.transitions``
(``src/SOAPify/transitions/__init__.py:9-167``): the step after the classification
(SURVEY.md section 8, row f4).

The histogram of state changes and the residence times are computed on the device from the
``int[F, A]`` classification.  The event-list bookkeeping of ``transitions/tracker.py``
(``StateTracker``) is host-side Python object juggling in the reference and stays out of scope:
``calculateResidenceTimes`` / ``calculateTransitionMatrix`` refuse a ``statesTracker`` argument.
"""
from __future__ import annotations

import numpy as np

from . import _device, _lib

__all__ = [
    "transitionMatrixFromSOAPClassification",
    "normalizeMatrixByRow",
    "transitionMatrixFromSOAPClassificationNormalized",
    "calculateResidenceTimesFromClassification",
    "calculateResidenceTimes",
    "calculateTransitionMatrix",
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
    to_frames = th.arange(int(window), n_frames, int(stride), device=states.device)
    if n_atoms and to_frames.numel():
        # transMat[classFrom, classTo] is numpy indexing in the reference (transitions/__init__.py:44-47): a
        # negative state counts from the end (-1 = the last class, the "error" class of its docstring) and a
        # state outside [-n_classes, n_classes) raises IndexError.  Only the frames the loop visits matter.
        visited = states[th.unique(th.cat((to_frames, to_frames - int(window))))]
        lo, hi = int(visited.min()), int(visited.max())
        if lo < -n_classes or hi >= n_classes:
            bad = hi if hi >= n_classes else lo
            raise IndexError(f"index {bad} is out of bounds for axis 0 with size {n_classes}")
        if lo < 0:
            states = th.where(states < 0, states + n_classes, states).contiguous()
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


def calculateResidenceTimesFromClassification(data, window: int = 1, stride: "int|None" = None) -> "list[np.ndarray]":
    """Calculates the residence time for each element of the classification
    (transitions/__init__.py:99-167): an ordered (sorted) array of residence times per state, the
    first and the last run of every series stored negative.

    One device thread walks one (atom, initial frame) series; the events are sorted on the device
    by (state, time) and split per state on the host.  Integer work, bit-exact."""
    if stride is None:
        stride = window
    if stride > window:
        raise ValueError("the window must be bigger than the stride")
    refs = data.references
    if window > refs.shape[0] or stride > refs.shape[0]:
        raise ValueError("stride and window must be smaller than simulation lenght")
    th = _device.require_cuda()
    lib = _lib.load()
    n_frames, n_atoms = int(refs.shape[0]), int(refs.shape[1])
    n_classes = len(data.legend)
    if n_classes == 0:
        return []
    if n_atoms == 0:
        return [np.sort(np.array([])) for _ in range(n_classes)]
    states = _device.to_device(refs)
    if states.dtype != th.int32:
        states = states.to(th.int32)
    states = th.where(states < 0, states + n_classes, states).contiguous()  # Python's negative list index
    cap = int(lib.soap_residence_capacity(n_frames, n_atoms, int(window), int(stride)))
    ev_state = th.empty(cap, dtype=th.int32, device=states.device)
    ev_time = th.empty(cap, dtype=th.int64, device=states.device)
    n_ev = th.zeros(1, dtype=th.int64, device=states.device)
    rc = lib.soap_residence_events(
        states.data_ptr(), n_frames, n_atoms, n_classes, int(window), int(stride), ev_state.data_ptr(), ev_time.data_ptr(),
        cap, n_ev.data_ptr(), _device.stream_ptr(),
    )
    _lib.check(rc, "soap_residence_events")
    n = int(n_ev.item())
    if n > cap:
        raise _lib.SoapB200Error("soap_residence_events produced more events than its stated capacity")
    ev_state, ev_time = ev_state[:n].to(th.int64), ev_time[:n]
    if n and (int(ev_state.min()) < 0 or int(ev_state.max()) >= n_classes):
        raise IndexError("list index out of range")  # what the reference's residenceTimes[state] raises
    span = 2 * (n_frames + int(window)) + 1  # times lie in [-(F + window), F]
    key, _ = th.sort(ev_state * span + (ev_time + (n_frames + int(window))))
    key = _device.to_host(key)
    st, tm = key // span, key % span - (n_frames + int(window))
    bounds = np.searchsorted(st, np.arange(n_classes + 1))
    # numpy.sort(numpy.array([])) is an empty float64 array in the reference
    return [tm[bounds[c] : bounds[c + 1]].astype(np.int64) if bounds[c + 1] > bounds[c] else np.sort(np.array([]))
            for c in range(n_classes)]


def calculateResidenceTimes(data, statesTracker: list = None, **algokwargs) -> "list[np.ndarray]":
    """Ordered list of residence times per state (transitions/__init__.py:170-201); the
    ``statesTracker`` route of the reference (``transitions/tracker.py``) is not part of this path."""
    if statesTracker:
        raise NotImplementedError("cpctools_b200: the StateTracker route (SOAPify.transitions.tracker) is out of scope")
    return calculateResidenceTimesFromClassification(data, **algokwargs)


def calculateTransitionMatrix(data, statesTracker: list = None, **algokwargs) -> "np.ndarray":
    """Unnormalised transition matrix (transitions/__init__.py:204-240); see ``calculateResidenceTimes``."""
    if statesTracker:
        raise NotImplementedError("cpctools_b200: the StateTracker route (SOAPify.transitions.tracker) is out of scope")
    return transitionMatrixFromSOAPClassification(data, **algokwargs)
