"""Generate ``tests/golden/*.npz`` from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Run in the build container, where
``/root/reference`` exists:

    python -m oracle.gen_golden

The reference ships no on-disk fixtures: every value stored here is an
output of ``/root/reference/src/SOAPify`` imported through ``oracle/ref_shim.py`` on seeded
inputs, so the files travel to the GPU box where the reference does not exist.  Inputs are
stored next to the outputs so no RNG replay is needed to use them.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

from .fake_dataset import FakeSOAPDataset, dscribe_attrs, synthetic_soap
from .ref_shim import load_reference

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SPECIES_SETS = {1: ["H"], 2: ["O", "H"], 3: ["O", "H", "C"], 4: ["Au", "H", "C", "O"]}
NMAX = (1, 4, 6, 8)
LMAX = (0, 4, 6, 8)


def _slices_from_attrs(attrs):
    out = {}
    for key, val in attrs.items():
        if key.startswith("species_location_"):
            a, b = key[len("species_location_"):].split("-")
            out[a + b] = slice(int(val[0]), int(val[1]))
    return out


def gen_tables(S):
    out = {}
    for ns, species in SPECIES_SETS.items():
        ordered = S.orderByZ(species)
        for n in NMAX:
            for l in LMAX:
                tag = f"S{ns}_n{n}_l{l}"
                if ns == 1:
                    fill = S.utils._getIndexesForFillSOAPVectorFromdscribe(l, n)
                else:
                    attrs, _ = dscribe_attrs(l, n, ordered)
                    fill = S.utils._getIndexesForFillSOAPVectorFromdscribe(
                        l, n, ordered, _slices_from_attrs(attrs)
                    )
                out["fill_" + tag] = np.asarray(fill, dtype=np.int32)
                out["quip_" + tag] = np.asarray(
                    S.getAddressesQuippyLikeDscribe(l, n, species), dtype=np.int32
                )
    # label layouts for one small and one config-5-sized case
    for ns, n, l in ((2, 2, 1), (3, 6, 6)):
        sp = SPECIES_SETS[ns]
        out[f"labels_dscribe_S{ns}_n{n}_l{l}"] = S.getdscribeSOAPMapping(l, n, sp)
        out[f"labels_dscribe_nocross_S{ns}_n{n}_l{l}"] = S.getdscribeSOAPMapping(l, n, sp, crossover=False)
        out[f"labels_quippy_S{ns}_n{n}_l{l}"] = S.getquippySOAPMapping(l, n, sp)
    return out


def gen_fill_norm(S):
    out = {}
    rng = np.random.default_rng(12345)
    # 1 species, config-1 shape (Dc 324 -> Df 576), three dtypes and ranks
    x3 = rng.random((6, 4, 324))
    out["fill1_in_f64"] = x3
    out["fill1_out_f64"] = S.fillSOAPVectorFromdscribe(x3, 8, 8)
    out["norm1_out_f64"] = S.normalizeArray(out["fill1_out_f64"])
    x3f = x3.astype(np.float32)
    out["fill1_out_f32"] = S.fillSOAPVectorFromdscribe(x3f, 8, 8)
    out["norm1_out_f32"] = S.normalizeArray(out["fill1_out_f32"])
    xi = rng.integers(0, 10, size=(5, 324))
    out["fill1_in_i64"] = xi
    out["fill1_out_i64"] = S.fillSOAPVectorFromdscribe(xi, 8, 8)
    out["norm1_out_i64"] = S.normalizeArray(out["fill1_out_i64"])
    x1 = rng.random(324)
    out["fill1_in_1d"] = x1
    out["fill1_out_1d"] = S.fillSOAPVectorFromdscribe(x1, 8, 8)
    # 2 species H/O, config-2 shape (Dc 1224 -> Df 1728); species given unsorted
    attrs, dc = dscribe_attrs(8, 8, ["H", "O"])
    x2 = rng.random((3, 5, dc))
    x2[1, 2] = 0.0  # an all-zero vector: normalizeArray must leave it at zero
    out["fill2_in_f64"] = x2
    out["fill2_out_f64"] = S.fillSOAPVectorFromdscribe(x2, 8, 8, ["O", "H"], _slices_from_attrs(attrs))
    out["norm2_out_f64"] = S.normalizeArray(out["fill2_out_f64"])
    # 3 species quippy-style, config-5 shape (raw 1198 -> 1197 -> 1512)
    sp3 = S.orderByZ(["H", "C", "O"])
    attrs3, dc3 = dscribe_attrs(6, 6, sp3)
    raw = rng.random((4, 3, dc3 + 1))
    addr = S.getAddressesQuippyLikeDscribe(6, 6, sp3)
    permuted = raw[..., :dc3][..., addr]  # engine.py:226,260
    out["quip3_in_raw"] = raw
    out["quip3_out_f64"] = S.fillSOAPVectorFromdscribe(permuted, 6, 6, sp3, _slices_from_attrs(attrs3))
    out["quip3_norm_f64"] = S.normalizeArray(out["quip3_out_f64"])
    return out


def gen_math(S):
    """The vectors of tests/test_SOAPifySOAP.py::test_MathSOAP plus edge cases."""
    out = {}
    rng = np.random.default_rng(12345)
    case = 0
    for _ in range(5):
        size = int(rng.integers(2, 150))
        x, y = rng.random(size), rng.random(size)
        xn, yn = x / np.linalg.norm(x), y / np.linalg.norm(y)
        out[f"x{case}"], out[f"y{case}"] = x, y
        vals = [S.simpleKernelSoap(x, y), S.simpleSOAPdistance(x, y), S.SOAPdistanceNormalized(xn, yn)]
        vals += [S.kernelSoap(x, y, n) for n in range(1, 10)]
        vals += [S.SOAPdistance(x, y, n) for n in range(1, 10)]
        out[f"vals{case}"] = np.array(vals, dtype=np.float64)
        xf, yf = x.astype(np.float32), y.astype(np.float32)
        valsf = [S.simpleSOAPdistance(xf, yf), S.SOAPdistance(xf, yf, 1), S.SOAPdistance(xf, yf, 2),
                 S.SOAPdistanceNormalized(xf / np.linalg.norm(xf), yf / np.linalg.norm(yf))]
        assert all(np.asarray(v).dtype == np.float32 for v in valsf)
        out[f"vals_f32_{case}"] = np.array(valsf, dtype=np.float32)
        case += 1
    # self distances and zero vectors
    X = rng.random((120, 576))
    Xn = S.normalizeArray(X)
    out["self_in"] = X
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out["self_simple"] = np.array([S.simpleSOAPdistance(v, v) for v in X])
        out["self_kernel1"] = np.array([S.SOAPdistance(v, v) for v in X])
        out["self_kernel2"] = np.array([S.SOAPdistance(v, v, 2) for v in X])
        out["self_normalized"] = np.array([S.SOAPdistanceNormalized(v, v) for v in Xn])
        z = np.zeros(576)
        out["zero_vals"] = np.array(
            [S.simpleSOAPdistance(z, X[0]), S.SOAPdistance(z, X[0]), S.SOAPdistanceNormalized(z, Xn[0])]
        )
    return out


def gen_timesoap(S):
    out = {}
    attrs, dc = dscribe_attrs(8, 8, ["H"])
    attrs1 = {"l_max": 8, "n_max": 8, "species": ["H"], "species_location_H-H": [0, dc]}
    for dist in ("U", "W"):
        raw = synthetic_soap((14, 5, dc), seed=12345, kind=dist)
        X = S.fillSOAPVectorFromdscribe(raw, 8, 8)
        Xn = S.normalizeArray(X)
        out[f"{dist}_raw"] = raw
        for w in (1, 2, 5):
            t, dt = S.timeSOAP(Xn, window=w)
            out[f"{dist}_simple_w{w}_t"], out[f"{dist}_simple_w{w}_dt"] = t, dt
            # un-normalised input through the two self-normalising distances
            out[f"{dist}_simple_raw_w{w}_t"] = S.timeSOAP(X, window=w, returnDiff=False)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                out[f"{dist}_kernel2_w{w}_t"] = S.timeSOAP(
                    X, window=w, returnDiff=False, distanceFunction=lambda a, b: S.SOAPdistance(a, b, 2)
                )
                out[f"{dist}_kernel1_w{w}_t"] = S.timeSOAP(
                    X, window=w, returnDiff=False, distanceFunction=S.SOAPdistance
                )
            out[f"{dist}_normalized_w{w}_t"] = S.timeSOAP(
                Xn, window=w, returnDiff=False, distanceFunction=S.SOAPdistanceNormalized
            )
            ts, dts = S.timeSOAPsimple(Xn, window=w)
            out[f"{dist}_tsimple_w{w}_t"], out[f"{dist}_tsimple_w{w}_dt"] = ts, dts
        # float32 input: computed in f32, stored to f64
        out[f"{dist}_f32_simple_w1_t"] = S.timeSOAP(Xn.astype(np.float32), window=1, returnDiff=False)
    # chunked streaming shortcut on a fake dataset
    raw = synthetic_soap((47, 5, dc), seed=2024, kind="W")
    ds = FakeSOAPDataset(raw, 10, attrs1)
    t, dt = S.getTimeSOAPSimple(ds)
    out["stream_raw"], out["stream_t"], out["stream_dt"] = raw, t, dt
    # two species through the streaming shortcut
    attrs2, dc2 = dscribe_attrs(4, 4, ["H", "O"])
    raw2 = synthetic_soap((23, 4, dc2), seed=7, kind="W")
    t2, dt2 = S.getTimeSOAPSimple(FakeSOAPDataset(raw2, 6, attrs2))
    out["stream2_raw"], out["stream2_t"], out["stream2_dt"] = raw2, t2, dt2
    return out


def gen_classify(S):
    out = {}
    rng = np.random.default_rng(12345)
    data = S.normalizeArray(rng.random((37, 576)))
    spectra = S.normalizeArray(rng.random((9, 576)))
    out["data"], out["spectra"] = data, spectra
    out["dm_normalized"] = S.getDistanceBetween(data, spectra, S.SOAPdistanceNormalized)
    out["dm_simple"] = S.getDistanceBetween(data, spectra, S.simpleSOAPdistance)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out["dm_kernel2"] = S.getDistanceBetween(data, spectra, lambda a, b: S.SOAPdistance(a, b, 2))
        out["dm_self_kernel1"] = S.getDistanceBetween(data, data, S.SOAPdistance)
    out["dm_self_normalized"] = S.getDistanceBetween(data, data, S.SOAPdistanceNormalized)
    d32 = data.astype(np.float32)
    out["dm_normalized_f32"] = S.getDistanceBetween(d32, spectra.astype(np.float32), S.SOAPdistanceNormalized)
    assert out["dm_normalized_f32"].dtype == np.float32
    # chunked classification on a fake mono-species dataset (classify.py:120-182, 250-276)
    raw = synthetic_soap((130, 4, 324), seed=99, kind="W")
    ds = FakeSOAPDataset(raw, 1000, {"l_max": 8, "n_max": 8, "species": ["Au"], "species_location_Au-Au": [0, 324]})
    refs = S.createReferencesFromTrajectory(
        ds, {"a": (0, 0), "b": (100, 3), "c": (129, 1), "d": (17, 2)}, 8, 8
    )
    out["cls_raw"], out["cls_ref_spectra"] = raw, refs.spectra
    out["cls_dist"] = S.getDistancesFromRefNormalized(ds, refs)
    cl = S.applyClassification(ds, refs, S.SOAPdistanceNormalized, doNormalize=True)
    out["cls_min"], out["cls_arg"] = cl.distances, cl.references
    out["cls_dist_simple"] = S.getDistancesFromRef(ds, refs, S.simpleSOAPdistance, doNormalize=False)
    return out


def gen_transitions(S):
    """transition matrices of random classifications (transitions/__init__.py:9-97)."""
    out = {}
    rng = np.random.default_rng(12345)
    refs = rng.integers(0, 5, size=(60, 17))
    out["refs"] = refs
    data = S.SOAPclassification(np.zeros(refs.shape), refs, ["a", "b", "c", "d", "e"])
    for stride, window in ((1, None), (1, 5), (3, 3), (2, 7), (1, 59), (1, 60)):
        tag = f"s{stride}_w{window}"
        out["tm_" + tag] = S.transitionMatrixFromSOAPClassification(data, stride, window)
        out["tmn_" + tag] = S.transitionMatrixFromSOAPClassificationNormalized(data, stride, window)
    # residence times (transitions/__init__.py:99-167): one sorted array per state; a slowly changing
    # classification as well, so that long runs exist
    slow = np.repeat(rng.integers(0, 5, size=(12, 17)), 5, axis=0)
    out["refs_slow"] = slow
    for name, arr in (("", refs), ("slow_", slow)):
        cls = S.SOAPclassification(np.zeros(arr.shape), arr, ["a", "b", "c", "d", "e"])
        for window, stride in ((1, None), (3, None), (4, 2), (5, 1), (60, 7), (59, 59)):
            rts = S.calculateResidenceTimesFromClassification(cls, window=window, stride=stride)
            for state, times in enumerate(rts):
                out[f"rt_{name}w{window}_s{stride}_{state}"] = np.asarray(times)
    # a classification with unclassified (-1) entries and an "error" legend entry, as the docstring of
    # calculateTransitionMatrix suggests: numpy indexing makes -1 the LAST class (transitions/__init__.py:44-47)
    err = np.random.default_rng(777).integers(-1, 5, size=(40, 23))
    out["refs_err"] = err
    ecls = S.SOAPclassification(np.zeros(err.shape), err, ["a", "b", "c", "d", "e", "error"])
    for stride, window in ((1, None), (2, 7)):
        out[f"tm_err_s{stride}_w{window}"] = S.transitionMatrixFromSOAPClassification(ecls, stride, window)
        out[f"tmn_err_s{stride}_w{window}"] = S.transitionMatrixFromSOAPClassificationNormalized(ecls, stride, window)
    return out


def main() -> int:
    S = load_reference()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, fn in (
        ("tables", gen_tables),
        ("fill_norm", gen_fill_norm),
        ("math", gen_math),
        ("timesoap", gen_timesoap),
        ("classify", gen_classify),
        ("transitions", gen_transitions),
    ):
        data = fn(S)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **data)
        print(f"{name}: {len(data)} arrays, {os.path.getsize(path) / 1e3:.1f} kB")
    return 0


if __name__ == "__main__":
    sys.exit(main())
