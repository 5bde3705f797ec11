"""Parity of the CUDA path (through the C ABI of libsoapb200.so) with the CPU oracle and with the
golden vectors generated from the unmodified reference.  All tests need a B200: ``-m gpu``.

Tolerances (written here once, see tests/parity.py): integer / gather work bit-exact;
normalisation 1e-12 relative (fp64), 1e-6 (fp32); distances by ``assert_distance_parity``
(d^2 to 1e-12 absolute, d to 1e-12 relative where d >= 0.1; 1e-6 and d >= 0.5 for the fp32 path).
"""
import functools
import warnings

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from oracle import soap_oracle as O
from oracle.fake_dataset import FakeSOAPDataset, dscribe_attrs, synthetic_soap
from oracle.gen_golden import _slices_from_attrs
from parity import assert_distance_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import cpctools_b200

    cpctools_b200._lib.load()  # fail loudly if the extension was not built
    return cpctools_b200


def _fuzz_offset():
    """SOAP_FUZZ_SEED=n shifts the seeds of the fuzz tests (soak runs: scripts/fuzz_soak.sh); default 0 = the pinned cases."""
    import os

    return int(os.environ.get("SOAP_FUZZ_SEED", "0"))


def _kinds(S):
    return {
        "simple": (S.simpleSOAPdistance, O.KIND_SIMPLE, 1),
        "kernel1": (S.SOAPdistance, O.KIND_KERNEL, 1),
        "kernel2": (functools.partial(S.SOAPdistance, n=2), O.KIND_KERNEL, 2),
        "normalized": (S.SOAPdistanceNormalized, O.KIND_NORMALIZED, 1),
    }


# ------------------------------------------------------------------ expansion: bit-exact
def test_fill_matches_golden_bit_exact(S, golden):
    g = golden("fill_norm")
    out = S.fillSOAPVectorFromdscribe(g["fill1_in_f64"], 8, 8)
    assert out.dtype == np.float64 and out.shape == (6, 4, 576)
    assert_array_equal(out, g["fill1_out_f64"])
    out32 = S.fillSOAPVectorFromdscribe(g["fill1_in_f64"].astype(np.float32), 8, 8)
    assert out32.dtype == np.float32
    assert_array_equal(out32, g["fill1_out_f32"])
    outi = S.fillSOAPVectorFromdscribe(g["fill1_in_i64"], 8, 8)
    assert outi.dtype == g["fill1_out_i64"].dtype
    assert_array_equal(outi, g["fill1_out_i64"])
    assert_array_equal(S.fillSOAPVectorFromdscribe(g["fill1_in_1d"], 8, 8), g["fill1_out_1d"])
    attrs, _ = dscribe_attrs(8, 8, ["H", "O"])
    out2 = S.fillSOAPVectorFromdscribe(g["fill2_in_f64"], 8, 8, ["O", "H"], _slices_from_attrs(attrs))
    assert_array_equal(out2, g["fill2_out_f64"])
    # quippy raw rows (trailing column, quippy slot order) through ONE composed gather
    attrs3, _ = dscribe_attrs(6, 6, ["H", "C", "O"])
    out3 = S.fillSOAPVectorFromQuip(g["quip3_in_raw"], 6, 6, ["O", "H", "C"], _slices_from_attrs(attrs3))
    assert_array_equal(out3, g["quip3_out_f64"])
    out3b = S.fillSOAPVectorFromQuip(g["quip3_in_raw"], 6, 6, ["H", "C", "O"])  # default container slices
    assert_array_equal(out3b, g["quip3_out_f64"])


@pytest.mark.parametrize("nmax", [1, 3, 4, 8])
@pytest.mark.parametrize("lmax", [0, 3, 4, 8])
@pytest.mark.parametrize("species", [None, ["O", "H"], ["H", "C", "O"]])
def test_fill_matches_oracle_all_layouts(S, nmax, lmax, species):
    """reference tests/test_SOAPify.py:135-240 (single, array, multi-species; integer equality)
    extended to odd row lengths (unaligned rows take the non-TMA path)."""
    rng = np.random.default_rng(100 * nmax + lmax)
    if species is None:
        kw, dc = {}, (lmax + 1) * nmax * (nmax + 1) // 2
    else:
        attrs, dc = dscribe_attrs(lmax, nmax, O.orderByZ(species))
        kw = dict(atomTypes=species, atomicSlices=_slices_from_attrs(attrs))
    for shape, dtype in (((dc,), np.int64), ((7, dc), np.float64), ((5, 11, dc), np.float32), ((3, 2, dc), np.int32)):
        a = rng.integers(0, 10, size=shape).astype(dtype)
        want = O.fillSOAPVectorFromdscribe(a, lmax, nmax, **kw)
        got = S.fillSOAPVectorFromdscribe(a, lmax, nmax, **kw)
        assert got.dtype == want.dtype and got.shape == want.shape
        assert_array_equal(got, want)
    with pytest.raises(Exception):
        S.fillSOAPVectorFromdscribe(rng.integers(0, 10, size=(5, dc + 1)), lmax, nmax, **kw)
    with pytest.raises(Exception):
        S.fillSOAPVectorFromdscribe(rng.integers(0, 10, size=(3, 4, 5, dc)), lmax, nmax, **kw)


def test_fill_large_tiles_and_ragged_tail(S):
    """many tiles per CTA, a ragged last tile, and an empty input."""
    rng = np.random.default_rng(5)
    attrs, dc = dscribe_attrs(8, 8, ["H", "O"])
    sl = _slices_from_attrs(attrs)
    a = rng.random((97, 131, dc))
    assert_array_equal(S.fillSOAPVectorFromdscribe(a, 8, 8, ["H", "O"], sl), O.fillSOAPVectorFromdscribe(a, 8, 8, ["H", "O"], sl))
    empty = np.zeros((0, dc))
    assert S.fillSOAPVectorFromdscribe(empty, 8, 8, ["H", "O"], sl).shape == (0, 1728)


def test_fill_small_dtypes_and_tiny_shapes(S):
    """element sizes the gather kernel does not move natively keep their dtype; degenerate shapes."""
    rng = np.random.default_rng(2)
    for dt in (np.int16, np.uint8, np.float16, np.uint32):
        a = rng.integers(0, 100, size=(3, 324)).astype(dt)
        got = S.fillSOAPVectorFromdscribe(a, 8, 8)
        assert got.dtype == dt
        assert_array_equal(got, O.fillSOAPVectorFromdscribe(a, 8, 8))
    one = rng.random((1, 1, 1))  # nMax = 1, lMax = 0: a single slot
    assert_array_equal(S.fillSOAPVectorFromdscribe(one, 0, 1), one)
    X = O.normalizeArray(rng.random((2, 1, 3)))
    t, dt_ = S.timeSOAP(X, window=1)
    assert t.shape == (1, 1) and dt_.shape == (1, 0)
    assert_distance_parity(t, O.timeSOAP(X, window=1, returnDiff=False), "f64")
    assert S.normalizeArray(np.zeros((0, 5))).shape == (0, 5)


def test_fill_device_tensor_stays_on_device(S):
    import torch

    x = torch.rand(4, 3, 324, dtype=torch.float64, device="cuda")
    y = S.fillSOAPVectorFromdscribe(x, 8, 8)
    assert isinstance(y, torch.Tensor) and y.is_cuda and y.shape == (4, 3, 576)
    assert_array_equal(y.cpu().numpy(), O.fillSOAPVectorFromdscribe(x.cpu().numpy(), 8, 8))


# ------------------------------------------------------------------ normalisation
def test_normalize_known_answers(S):
    """reference tests/test_SOAPify.py:19-41 (exact)."""
    assert_array_equal(S.normalizeArray(np.array([2.0, 0.0])), [1.0, 0.0])
    assert_array_equal(S.normalizeArray(np.array([[2.0, 0.0, 0.0], [0.0, 0.0, 3.0]])), [[1.0, 0, 0], [0, 0, 1.0]])
    a = np.array([[[2.0, 0.0, 0.0], [0.0, 0.0, 3.0]], [[0.0, 2.0, 0.0], [0.0, 3.0, 0.0]]])
    assert_array_equal(S.normalizeArray(a), [[[1.0, 0, 0], [0, 0, 1.0]], [[0, 1.0, 0], [0, 1.0, 0]]])


def test_normalize_matches_golden(S, golden):
    g = golden("fill_norm")
    assert_allclose(S.normalizeArray(g["fill1_out_f64"]), g["norm1_out_f64"], rtol=1e-12, atol=0)
    n32 = S.normalizeArray(g["fill1_out_f32"])
    assert n32.dtype == np.float32
    assert_allclose(n32, g["norm1_out_f32"], rtol=1e-6, atol=0)
    ni = S.normalizeArray(g["fill1_out_i64"])  # integer input is promoted to float64
    assert ni.dtype == np.float64
    assert_allclose(ni, g["norm1_out_i64"], rtol=1e-12, atol=0)
    n2 = S.normalizeArray(g["fill2_out_f64"])
    assert_allclose(n2, g["norm2_out_f64"], rtol=1e-12, atol=0)
    assert not n2[1, 2].any()  # the all-zero vector stays zero
    # fused expand+normalise == the two reference steps
    attrs, _ = dscribe_attrs(8, 8, ["H", "O"])
    fused = S.fillAndNormalize(g["fill2_in_f64"], 8, 8, ["O", "H"], _slices_from_attrs(attrs))
    assert_allclose(fused, g["norm2_out_f64"], rtol=1e-12, atol=0)
    fq = S.fillAndNormalize(g["quip3_in_raw"], 6, 6, ["H", "C", "O"], quippy=True)
    assert_allclose(fq, g["quip3_norm_f64"], rtol=1e-12, atol=0)


# ------------------------------------------------------------------ scalar distances
def test_math_matches_golden(S, golden):
    """reference tests/test_SOAPifySOAP.py:339-376 (decimal 8 there; 1e-12 here)."""
    g = golden("math")
    for c in range(5):
        x, y = g[f"x{c}"], g[f"y{c}"]
        xn, yn = x / np.linalg.norm(x), y / np.linalg.norm(y)
        vals = [S.simpleKernelSoap(x, y), S.simpleSOAPdistance(x, y), S.SOAPdistanceNormalized(xn, yn)]
        vals += [S.kernelSoap(x, y, n) for n in range(1, 10)]
        vals += [S.SOAPdistance(x, y, n) for n in range(1, 10)]
        want = g[f"vals{c}"]
        assert_allclose(np.array(vals)[[0] + list(range(3, 12))], want[[0] + list(range(3, 12))], rtol=0, atol=1e-12)
        assert_distance_parity(np.array(vals)[[1, 2] + list(range(12, 21))], want[[1, 2] + list(range(12, 21))], "f64")
        xf, yf = x.astype(np.float32), y.astype(np.float32)
        vf = [S.simpleSOAPdistance(xf, yf), S.SOAPdistance(xf, yf, 1), S.SOAPdistance(xf, yf, 2),
              S.SOAPdistanceNormalized(xf / np.linalg.norm(xf), yf / np.linalg.norm(yf))]
        assert all(np.asarray(v).dtype == np.float32 for v in vf)
        assert_distance_parity(np.array(vf), g[f"vals_f32_{c}"], "f32_ref")


def test_self_and_zero_distances(S, golden):
    g = golden("math")
    X = g["self_in"]
    z = np.zeros(576)
    assert np.isnan(S.simpleSOAPdistance(z, X[0])) and np.isnan(S.SOAPdistance(z, X[0]))
    assert S.SOAPdistanceNormalized(z, X[0] / np.linalg.norm(X[0])) == np.sqrt(2.0)
    d = S.getDistanceBetween(X, X, S.simpleSOAPdistance)
    assert (np.diag(d) == 0.0).all()  # the scipy clip makes self distances exactly 0
    assert_distance_parity(np.diag(S.getDistanceBetween(X, X, S.SOAPdistance)), g["self_kernel1"], "f64")
    Xn = O.normalizeArray(X)
    assert_distance_parity(np.diag(S.getDistanceBetween(Xn, Xn, S.SOAPdistanceNormalized)), g["self_normalized"], "f64")


def test_unknown_callable_is_refused(S):
    X = np.random.default_rng(0).random((4, 3, 8))
    with pytest.raises(NotImplementedError):
        S.timeSOAP(X, distanceFunction=lambda a, b: 0.0)
    with pytest.raises(NotImplementedError):
        S.getDistanceBetween(X[0], X[1], lambda a, b: 0.0)


# ------------------------------------------------------------------ timeSOAP family
@pytest.mark.parametrize("dist", ["U", "W"])
def test_timesoap_matches_golden(S, golden, dist):
    g = golden("timesoap")
    raw = g[f"{dist}_raw"]
    X = O.fillSOAPVectorFromdscribe(raw, 8, 8)
    Xn = O.normalizeArray(X)
    K = _kinds(S)
    for w in (1, 2, 5):
        t, dt = S.timeSOAP(Xn, window=w)
        assert t.dtype == np.float64 and t.shape == (14 - w, 5) and dt.shape == (5, 13 - w)
        assert_distance_parity(t, g[f"{dist}_simple_w{w}_t"], "f64", what=f"simple w{w}")
        assert_array_equal(dt, np.diff(t.T, axis=-1))  # the diff kernel is exact on its input
        assert_distance_parity(S.timeSOAP(X, window=w, returnDiff=False), g[f"{dist}_simple_raw_w{w}_t"], "f64")
        for name in ("kernel1", "kernel2"):
            got = S.timeSOAP(X, window=w, returnDiff=False, distanceFunction=K[name][0])
            assert_distance_parity(got, g[f"{dist}_{name}_w{w}_t"], "f64", what=f"{name} w{w}")
        got = S.timeSOAP(Xn, window=w, returnDiff=False, distanceFunction=S.SOAPdistanceNormalized)
        assert_distance_parity(got, g[f"{dist}_normalized_w{w}_t"], "f64", what=f"normalized w{w}")
        ts, dts = S.timeSOAPsimple(Xn, window=w)
        assert_allclose(ts, g[f"{dist}_tsimple_w{w}_t"], rtol=1e-12, atol=1e-15)
        assert_allclose(dts, g[f"{dist}_tsimple_w{w}_dt"], rtol=0, atol=1e-13)
        # straight from the compressed data (expansion + normalisation folded in)
        for name, key in (("simple", "simple"), ("kernel2", "kernel2"), ("normalized", "normalized")):
            (got,) = S.timeSOAPfromCompressed(raw, 8, 8, windows=(w,), returnDiff=False, distanceFunction=K[name][0])
            assert_distance_parity(got, g[f"{dist}_{key}_w{w}_t"], "f64", what=f"compressed {name} w{w}")
    t32 = S.timeSOAP(Xn.astype(np.float32), window=1, returnDiff=False)
    assert t32.dtype == np.float64  # analysis.py:53
    assert_distance_parity(t32, g[f"{dist}_simple_w1_t"], "f32_ref")  # golden = fp64 arithmetic on unrounded inputs


def test_timesoap_errors(S):
    """reference tests/test_analysis.py:41,46 pin these substrings."""
    X = np.zeros((10, 2, 4))
    with pytest.raises(ValueError, match="the window must be bigger than the stride"):
        S.timeSOAP(X, window=2, stride=3)
    with pytest.raises(ValueError, match="stride and window must be smaller than simulation lenght"):
        S.timeSOAP(X, window=10)
    with pytest.raises(ValueError, match="stride and window must be smaller"):
        S.timeSOAPsimple(X, window=11)


@pytest.mark.parametrize(
    "F,A,species,dtype,windows",
    [
        (70, 37, 1, np.float64, (1, 5, 10, 50)),     # config-1 vector, the benchmark's delay sweep
        (131, 9, 2, np.float64, (1, 5, 10, 50)),     # config-2/3 vector: rows are split into pieces
        (131, 9, 2, np.float32, (1, 5, 10, 50)),
        (40, 6, 3, np.float64, (1, 7)),              # config-5 vector (Dc 1197: unaligned rows)
        (33, 3, 1, np.float64, (32,)),               # a window as long as one ring block
        (200, 2, 1, np.float64, (150,)),             # a window of many blocks
    ],
)
def test_timesoap_sweep_from_compressed_vs_oracle(S, F, A, species, dtype, windows):
    sp = {1: None, 2: ["H", "O"], 3: ["H", "C", "O"]}[species]
    l = n = 6 if species == 3 else 8
    if sp is None:
        kw, dc = {}, (l + 1) * n * (n + 1) // 2
    else:
        attrs, dc = dscribe_attrs(l, n, sp)
        kw = dict(atomTypes=sp, atomicSlices=_slices_from_attrs(attrs))
    raw = synthetic_soap((F, A, dc), seed=F + A, kind="W", dtype=dtype)
    X = O.normalizeArray(O.fillSOAPVectorFromdscribe(raw.astype(np.float64), l, n, **kw))
    path = "f64" if dtype == np.float64 else "f32_timelag"
    got = S.timeSOAPfromCompressed(raw, l, n, windows=windows, **kw)
    mat = S.timeSOAPsweep(X.astype(dtype), windows=windows)
    for w, (t, dt), (tm, dtm) in zip(windows, got, mat):
        want, wdt = O.timeSOAP_batched(X, window=w, kind=O.KIND_SIMPLE)
        assert t.shape == want.shape and dt.shape == wdt.shape
        assert_distance_parity(t, want, path, what=f"compressed w{w}")
        assert_distance_parity(tm, want, path, what=f"materialised w{w}")
        assert_array_equal(dt, np.diff(t.T, axis=-1))
        if dtype == np.float64:
            eu = S.timeSOAPsimple(X, window=w, returnDiff=False) if w == 1 else None
            if eu is not None:
                assert_allclose(eu, np.linalg.norm(X[1:] - X[:-1], axis=-1), rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize(
    "F,A,D,dtype,windows",
    [
        (5, 1, 4, np.float64, (1, 2)),        # one box is wider and taller than the whole trajectory
        (3, 2, 8, np.float32, (1,)),          # fewer frames than a block, fewer blocks than the prefetch depth
        (33, 1, 18, np.float64, (1, 32)),     # two blocks, the second almost empty
        (40, 700, 36, np.float64, (1, 3)),    # several work items per persistent CTA (700 > SMs)
        (64, 450, 250, np.float32, (1, 5, 10, 50)),
        (97, 160, 1224, np.float64, (1, 5, 10, 50)),  # rows split into pieces, pieces change inside a CTA
        (60, 640, 1224, np.float64, (1, 5, 10, 50)),  # enough atoms for the atom-major mode: pieces summed on chip
        (41, 601, 612, np.float32, (2, 40)),
        (64, 450, 1198, np.float32, (1, 5, 10, 50)),  # raw quippy fp32 rows: 8- but not 16-byte multiples, TMA + masked tail unit
        (40, 6, 6, np.float32, (1, 3)),       # the same with rows shorter than two units
        (35, 5, 7, np.float64, (2, 4)),       # unaligned rows AND frames: element-copy path, several items per CTA
        (35, 300, 7, np.float32, (1, 9)),     # rows of 28 bytes in frames of whole units: TMA, leads 0..3 floats
        (64, 448, 1197, np.float32, (1, 5, 10, 50)),  # dscribe 3 species n = l = 6 in fp32: odd rows, pieces + leads
        (50, 1000, 147, np.float32, (1, 5)),  # odd rows, atom-major mode
        (37, 6, 9, np.float64, (1, 4)),       # fp64 odd rows (72 bytes), even atom count
        (33, 12, 5, np.float32, (2,)),        # rows shorter than two units with every lead
        (33, 3, 5, np.float32, (2,)),         # ... and a frame that is NOT a whole unit: element-copy path
    ],
)
def test_timelag_persistent_stream_edge_shapes(S, F, A, D, dtype, windows, monkeypatch):
    """The timelag kernel walks (atom, piece) items as one continuous stream of 32-frame blocks per CTA;
    results must not depend on where item boundaries fall, with and without multiplicities."""
    import torch
    from cpctools_b200 import _lib
    from cpctools_b200.analysis import timelag_device

    rng = np.random.default_rng(F * 1000 + A + D)
    X = rng.random((F, A, D)).astype(dtype)
    path = "f64" if dtype == np.float64 else "f32_timelag"
    got = S.timeSOAPsweep(X, windows=windows)
    X64 = X.astype(np.float64)
    for w, (t, dt) in zip(windows, got):
        want, _ = O.timeSOAP_batched(X64, window=w, kind=O.KIND_SIMPLE)
        assert_distance_parity(t, want, path, what=f"w{w}")
        assert_array_equal(dt, np.diff(t.T, axis=-1))
    # weighted dots (multiplicities 1 / 2) == plain dots of the vectors scaled by sqrt(m)
    m = (1.0 + (rng.random(D) < 0.4)).astype(dtype)
    ts = timelag_device(torch.from_numpy(X).cuda(), windows, _lib.KIND_SIMPLE, 1, mult=torch.from_numpy(m).cuda())
    Xw = X64 * np.sqrt(m.astype(np.float64))
    for w, t in zip(windows, ts):
        want, _ = O.timeSOAP_batched(Xw, window=w, kind=O.KIND_SIMPLE)
        assert_distance_parity(t.cpu().numpy(), want, path, what=f"weighted w{w}")
    # the same call squeezed onto two CTAs (every CTA streams many items): bit-identical results
    monkeypatch.setenv("SOAP_TL_GRID", "2")
    ts2 = timelag_device(torch.from_numpy(X).cuda(), windows, _lib.KIND_SIMPLE, 1, mult=torch.from_numpy(m).cuda())
    for t, t2 in zip(ts, ts2):
        assert_array_equal(t.cpu().numpy(), t2.cpu().numpy())


@pytest.mark.parametrize(
    "F,A,D,dtype,windows",
    [
        (300, 9, 1224, np.float64, (1, 5, 10, 50)),   # the benchmark sweep: 3 register windows + 1 ring window
        (300, 9, 1224, np.float32, (1, 5, 10, 50)),   # fp32 rows start at 4 different phases of a 128-byte line
        (113, 7, 1198, np.float64, (1, 10)),          # raw quippy rows (599 units): every phase, one frame past a block
        (112, 5, 324, np.float64, (50,)),             # exactly one block, ring window only
        (225, 6, 324, np.float32, (2, 3, 7, 9)),      # four register windows
        (57, 11, 576, np.float64, (56, 4)),           # the longest window the ring keeps; unsorted windows
        (130, 4, 6, np.float64, (1, 5, 10, 50)),      # rows shorter than a line (3 units): several atoms per line
        (40, 13, 20, np.float32, (3, 39)),            # fewer frames than a block
        (260, 3, 2048, np.float64, (1, 56)),          # rows that start and end on a line (no masked units)
        (500, 160, 1224, np.float64, (1, 5, 10, 50)), # more atoms than CTAs: several atoms per CTA
    ],
)
def test_timelag_register_blocked_kernel(S, F, A, D, dtype, windows, monkeypatch):
    """The register-blocked sweep kernel (7 frames per lane, line-aligned pieces, column-pipelined ring) forced on
    for shapes of every kind: against the oracle, against the lane-per-frame kernel, and bitwise reproducible
    across grid sizes.  Neighbouring atoms' data must never leak through the masked units of a shared line."""
    import torch
    from cpctools_b200 import _lib
    from cpctools_b200.analysis import timelag_device

    rng = np.random.default_rng(F * 1000 + A + D)
    X = rng.random((F, A, D)).astype(dtype)
    m = (1.0 + (rng.random(D) < 0.4)).astype(dtype)
    path = "f64" if dtype == np.float64 else "f32_timelag"
    Xd, md = torch.from_numpy(X).cuda(), torch.from_numpy(m).cuda()
    before = _lib.launch_count()
    monkeypatch.setenv("SOAP_TL_RB", "1")
    plain = [t.cpu().numpy() for t in timelag_device(Xd, windows, _lib.KIND_SIMPLE, 1)]
    weighted = [t.cpu().numpy() for t in timelag_device(Xd, windows, _lib.KIND_SIMPLE, 1, mult=md)]
    fused = [t.cpu().numpy() for t in timelag_device(Xd, windows, _lib.KIND_NORMALIZED_FUSED, 1, mult=md)]
    monkeypatch.setenv("SOAP_TL_GRID", "3")
    weighted3 = [t.cpu().numpy() for t in timelag_device(Xd, windows, _lib.KIND_SIMPLE, 1, mult=md)]
    monkeypatch.delenv("SOAP_TL_GRID")
    assert _lib.launch_count() - before == 4 * 3  # table setup + sweep + finalize per call
    monkeypatch.setenv("SOAP_TL_RB", "0")
    old = [t.cpu().numpy() for t in timelag_device(Xd, windows, _lib.KIND_SIMPLE, 1, mult=md)]
    X64 = X.astype(np.float64)
    Xw = X64 * np.sqrt(m.astype(np.float64))
    for k, w in enumerate(windows):
        assert plain[k].shape == (F - w, A)
        assert_distance_parity(plain[k], O.timeSOAP_batched(X64, window=w, kind=O.KIND_SIMPLE)[0], path, what=f"plain w{w}")
        want = O.timeSOAP_batched(Xw, window=w, kind=O.KIND_SIMPLE)[0]
        assert_distance_parity(weighted[k], want, path, what=f"weighted w{w}")
        assert_distance_parity(weighted[k], old[k], path, what=f"vs lane-per-frame kernel w{w}")
        assert_distance_parity(fused[k], O.timeSOAP_batched(O.normalizeArray(Xw), window=w, kind=O.KIND_NORMALIZED)[0], path, what=f"fused w{w}")
        assert_array_equal(weighted[k], weighted3[k])
    # an atom full of NaN poisons only itself, although it shares its first and last line with its neighbours
    if A >= 3:
        Xn = X.copy()
        Xn[:, 1, :] = np.nan
        monkeypatch.setenv("SOAP_TL_RB", "1")
        got = timelag_device(torch.from_numpy(Xn).cuda(), windows[:1], _lib.KIND_SIMPLE, 1, mult=md)[0].cpu().numpy()
        assert np.isnan(got[:, 1]).all()
        # (a call with fewer windows may cut the rows into different pieces: same values to rounding, not bitwise)
        assert_distance_parity(np.delete(got, 1, axis=1), np.delete(weighted[0], 1, axis=1), path, what="NaN neighbour")


@pytest.mark.parametrize(
    "F,A,D,dtype,windows",
    [
        (200, 40, 1224, np.float64, (1, 5, 10, 50)),   # the bench's row: ten pieces, (atom, piece) items dealt round robin
        (130, 700, 324, np.float64, (1, 5, 10, 50)),   # atom-major mode (pieces summed on chip), a ragged last block
        (66, 9, 1224, np.float32, (5, 1, 64)),         # window 1 not first, the longest window the ring keeps, 3 windows
        (64, 3, 8, np.float64, (1, 2, 3, 4)),          # one block, rows of four units, even and odd lags side by side
        (2, 5, 40, np.float32, (1, 1, 1)),             # the smallest trajectory
        (400, 300, 36, np.float32, (1, 7, 32, 33)),    # short rows, many items per CTA
        (1000, 64, 1728, np.float64, (1, 5, 10, 50)),  # materialised vectors without multiplicities path too
    ],
)
def test_timelag_two_frames_per_lane_kernel(S, F, A, D, dtype, windows, monkeypatch):
    """The sweep kernel with a frame PAIR per lane (timelag_pair.cuh: even / odd ring halves through a 3-D tensor map,
    lag-1 partner of the odd frame from registers, 64-frame blocks) forced on: against the oracle, against the
    lane-per-frame kernel, and bitwise reproducible across grid sizes."""
    import torch
    from cpctools_b200 import _lib
    from cpctools_b200.analysis import timelag_device

    rng = np.random.default_rng(F * 1000 + A + D)
    X = rng.random((F, A, D)).astype(dtype)
    m = (1.0 + (rng.random(D) < 0.4)).astype(dtype)
    path = "f64" if dtype == np.float64 else "f32_timelag"
    Xd, md = torch.from_numpy(X).cuda(), torch.from_numpy(m).cuda()
    monkeypatch.setenv("SOAP_TL_PAIR", "1")
    plain = [t.cpu().numpy() for t in timelag_device(Xd, windows, _lib.KIND_SIMPLE, 1)]
    weighted = [t.cpu().numpy() for t in timelag_device(Xd, windows, _lib.KIND_SIMPLE, 1, mult=md)]
    monkeypatch.setenv("SOAP_TL_GRID", "3")
    weighted3 = [t.cpu().numpy() for t in timelag_device(Xd, windows, _lib.KIND_SIMPLE, 1, mult=md)]
    monkeypatch.delenv("SOAP_TL_GRID")
    monkeypatch.setenv("SOAP_TL_PAIR", "0")
    old = [t.cpu().numpy() for t in timelag_device(Xd, windows, _lib.KIND_SIMPLE, 1, mult=md)]
    X64 = X.astype(np.float64)
    Xw = X64 * np.sqrt(m.astype(np.float64))
    for k, w in enumerate(windows):
        assert plain[k].shape == (F - w, A)
        assert_distance_parity(plain[k], O.timeSOAP_batched(X64, window=w, kind=O.KIND_SIMPLE)[0], path, what=f"plain w{w}")
        assert_distance_parity(weighted[k], O.timeSOAP_batched(Xw, window=w, kind=O.KIND_SIMPLE)[0], path, what=f"weighted w{w}")
        assert_distance_parity(weighted[k], old[k], path, what=f"vs lane-per-frame kernel w{w}")
        assert_array_equal(weighted[k], weighted3[k])
    if A >= 3:  # an atom full of NaN poisons only itself
        Xn = X.copy()
        Xn[:, 1, :] = np.nan
        monkeypatch.setenv("SOAP_TL_PAIR", "1")
        got = [t.cpu().numpy() for t in timelag_device(torch.from_numpy(Xn).cuda(), windows, _lib.KIND_SIMPLE, 1, mult=md)]
        for k in range(len(windows)):
            assert np.isnan(got[k][:, 1]).all()
            assert_array_equal(np.delete(got[k], 1, axis=1), np.delete(weighted[k], 1, axis=1))


@pytest.mark.parametrize(
    "F,A,D,dtype,windows",
    [
        (130, 600, 1224, np.float64, (1, 5, 10, 50)),   # the headline sweep: five pieces per row, the last one ragged
        (60, 700, 36, np.float64, (3, 5, 10)),          # three windows, ring window 3, an odd number of blocks per item
        (77, 640, 200, np.float32, (50, 1, 10, 5)),     # one piece of 50 units, windows in any order
        (26, 600, 1728, np.float32, (2, 10, 5)),        # two blocks, the second nearly empty; four pieces
        (120, 592, 516, np.float64, (25, 5, 10, 50)),   # the largest ring window; two full columns + a ragged one
        (230, 5, 1224, np.float64, (1, 5, 10, 50)),     # forced on for a handful of atoms (bench.py's neighbour check)
    ],
)
def test_timelag_chain_kernel(S, F, A, D, dtype, windows, monkeypatch):
    """The sweep kernel with lane = units and the lag history in registers (timelag_chain.cuh: warp = frame mod 5,
    windows 5 / 10 / 50 from registers, transposing butterfly over the lanes) forced on: against the oracle, against
    the lane-per-frame kernel, bitwise reproducible across grid sizes, NaN isolation."""
    import torch
    from cpctools_b200 import _lib
    from cpctools_b200.analysis import timelag_device

    rng = np.random.default_rng(F * 1000 + A + D)
    X = rng.random((F, A, D)).astype(dtype)
    m = (1.0 + (rng.random(D) < 0.4)).astype(dtype)
    path = "f64" if dtype == np.float64 else "f32_timelag"
    Xd, md = torch.from_numpy(X).cuda(), torch.from_numpy(m).cuda()
    monkeypatch.setenv("SOAP_TL_CHAIN", "1")
    _lib.profile_enable(True)
    plain = [t.cpu().numpy() for t in timelag_device(Xd, windows, _lib.KIND_SIMPLE, 1)]
    weighted = [t.cpu().numpy() for t in timelag_device(Xd, windows, _lib.KIND_SIMPLE, 1, mult=md)]
    monkeypatch.setenv("SOAP_TL_GRID", "3")
    weighted3 = [t.cpu().numpy() for t in timelag_device(Xd, windows, _lib.KIND_SIMPLE, 1, mult=md)]
    monkeypatch.delenv("SOAP_TL_GRID")
    monkeypatch.setenv("SOAP_TL_CHAIN", "0")
    old = [t.cpu().numpy() for t in timelag_device(Xd, windows, _lib.KIND_SIMPLE, 1, mult=md)]
    _lib.profile_enable(False)
    X64 = X.astype(np.float64)
    Xw = X64 * np.sqrt(m.astype(np.float64))
    for k, w in enumerate(windows):
        assert plain[k].shape == (F - w, A)
        assert_distance_parity(plain[k], O.timeSOAP_batched(X64, window=w, kind=O.KIND_SIMPLE)[0], path, what=f"plain w{w}")
        assert_distance_parity(weighted[k], O.timeSOAP_batched(Xw, window=w, kind=O.KIND_SIMPLE)[0], path, what=f"weighted w{w}")
        assert_distance_parity(weighted[k], old[k], path, what=f"vs lane-per-frame kernel w{w}")
        assert_array_equal(weighted[k], weighted3[k])
    # an atom full of NaN poisons only itself
    Xn = X.copy()
    Xn[:, 1, :] = np.nan
    monkeypatch.setenv("SOAP_TL_CHAIN", "1")
    got = [t.cpu().numpy() for t in timelag_device(torch.from_numpy(Xn).cuda(), windows, _lib.KIND_SIMPLE, 1, mult=md)]
    for k in range(len(windows)):
        assert np.isnan(got[k][:, 1]).all()
        assert_array_equal(np.delete(got[k], 1, axis=1), np.delete(weighted[k], 1, axis=1))


@pytest.mark.parametrize("D,dtype", [(1198, np.float32), (1197, np.float32), (147, np.float32), (7, np.float32), (1197, np.float64)])
def test_timelag_rows_ending_in_half_a_unit_do_not_leak_the_next_row(S, D, dtype):
    """Rows that are not whole 16-byte units (raw quippy fp32: 1198 floats; dscribe 3 species n = l = 6: 1197) go
    through the TMA path with windows that start at the unit boundary below the row: the first and / or last unit of
    an atom carries elements of its neighbours.  A NaN / huge neighbour must not change an atom's result (and
    poisons only itself)."""
    import torch
    from cpctools_b200 import _lib
    from cpctools_b200.analysis import timelag_device

    rng = np.random.default_rng(5)
    X = rng.random((50, 8, D)).astype(dtype)
    base = [t.cpu().numpy() for t in timelag_device(torch.from_numpy(X).cuda(), (1, 7), _lib.KIND_SIMPLE, 1)]
    for w, t in zip((1, 7), base):
        assert_distance_parity(t, O.timeSOAP_batched(X.astype(np.float64), window=w, kind=O.KIND_SIMPLE)[0],
                               "f32_timelag" if dtype == np.float32 else "f64")
    Xn = X.copy()
    Xn[:, 3, :] = np.nan
    got = [t.cpu().numpy() for t in timelag_device(torch.from_numpy(Xn).cuda(), (1, 7), _lib.KIND_SIMPLE, 1)]
    for b, g in zip(base, got):
        assert np.isnan(g[:, 3]).all()
        assert_array_equal(np.delete(g, 3, axis=1), np.delete(b, 3, axis=1))


def test_host_trajectory_larger_than_budget_is_tiled_over_atoms(S, monkeypatch):
    """host inputs that do not fit the device budget are processed in atom tiles (same results)."""
    from cpctools_b200 import analysis

    rng = np.random.default_rng(8)
    X = O.normalizeArray(rng.random((20, 11, 64)))
    monkeypatch.setattr(analysis, "_atom_tiles", lambda x, per_atom: [(0, 4), (4, 8), (8, 11)])
    t, dt = S.timeSOAP(X, window=3)
    want, wdt = O.timeSOAP_batched(X, window=3)
    assert t.shape == want.shape and dt.shape == wdt.shape
    assert_distance_parity(t, want, "f64")
    assert_array_equal(dt, np.diff(t.T, axis=-1))
    res = S.timeSOAPsweep(X, windows=(1, 2), returnDiff=False)
    assert_distance_parity(res[1], O.timeSOAP_batched(X, window=2, returnDiff=False), "f64")


def test_host_inputs_fan_out_over_devices(S, monkeypatch):
    """``devices=``: a host trajectory / vector set is split into contiguous atom ranges / row blocks, one host
    thread per device (here the same device twice: the code path, not the speed-up); results are those of the
    single-device call -- numpy, pinned tensor (pipelined upload of a sub-range) and the distance matrix."""
    import torch
    from cpctools_b200 import analysis

    attrs, dc = dscribe_attrs(8, 8, ["H", "O"])
    kw = dict(atomTypes=["H", "O"], atomicSlices=_slices_from_attrs(attrs))
    raw = synthetic_soap((60, 41, dc), seed=21, kind="W")
    one = S.timeSOAPfromCompressed(raw, 8, 8, windows=(1, 7), devices=[0], **kw)
    two = S.timeSOAPfromCompressed(raw, 8, 8, windows=(1, 7), devices=[0, 0], **kw)
    for (t1, d1), (t2, d2) in zip(one, two):
        assert isinstance(t2, np.ndarray) and t2.shape == t1.shape and d2.shape == d1.shape
        assert_array_equal(t2, t1)
        assert_array_equal(d2, d1)
    monkeypatch.setattr(analysis, "PIPELINE_MIN_BYTES", 1)
    pinned = torch.from_numpy(raw).pin_memory()
    three = S.timeSOAPfromCompressed(pinned, 8, 8, windows=(1, 7), devices=[0, 0, 0], **kw)
    for (t1, d1), (t3, d3) in zip(one, three):
        assert not t3.is_cuda
        assert_array_equal(t3.numpy(), t1)
        assert_array_equal(d3.numpy(), d1)
    X = O.normalizeArray(np.random.default_rng(2).random((60, 41, 64)))
    assert_array_equal(S.timeSOAP(X, window=3, devices=[0, 0])[0], S.timeSOAP(X, window=3, devices=0)[0])
    assert_array_equal(S.timeSOAPsimple(X, window=2, devices=[0, 0])[0], S.timeSOAPsimple(X, window=2)[0])
    monkeypatch.setenv("SOAP_B200_DEVICES", "0,0")
    assert_array_equal(S.timeSOAP(X, window=3)[0], S.timeSOAP(X, window=3, devices=0)[0])
    monkeypatch.delenv("SOAP_B200_DEVICES")
    with pytest.raises(ValueError, match="CUDA device"):
        S.timeSOAP(X, window=3, devices=[0, 99])
    V = O.normalizeArray(np.random.default_rng(3).random((700, 576)))
    assert_array_equal(S.getDistanceBetween(V, V[:300], S.SOAPdistanceNormalized, devices=[0, 0]),
                       S.getDistanceBetween(V, V[:300], S.SOAPdistanceNormalized, devices=[0]))


def test_pinned_host_tensor_is_uploaded_in_overlapped_atom_tiles(S, monkeypatch):
    """a pinned host tensor takes the pipelined path (strided tile uploads on a copy stream, results
    downloaded per tile) and returns HOST tensors identical to the one-shot path."""
    import torch
    from cpctools_b200 import analysis

    attrs, dc = dscribe_attrs(8, 8, ["H", "O"])
    kw = dict(atomTypes=["H", "O"], atomicSlices=_slices_from_attrs(attrs))
    raw = synthetic_soap((70, 37, dc), seed=11, kind="W")
    want = S.timeSOAPfromCompressed(raw, 8, 8, windows=(1, 5, 50), **kw)
    pinned = torch.from_numpy(raw).pin_memory()
    monkeypatch.setattr(analysis, "PIPELINE_MIN_BYTES", 1)
    monkeypatch.setattr(analysis, "PIPELINE_TILE_BYTES", 5 * 70 * dc * 8)  # 5 atoms per tile: 8 tiles, the last of 2
    got = S.timeSOAPfromCompressed(pinned, 8, 8, windows=(1, 5, 50), **kw)
    for (t, dt), (tw, dtw) in zip(got, want):
        assert isinstance(t, torch.Tensor) and not t.is_cuda and not dt.is_cuda
        assert_array_equal(t.numpy(), tw)
        assert_array_equal(dt.numpy(), dtw)
    only_t = S.timeSOAP(torch.from_numpy(O.normalizeArray(O.fillSOAPVectorFromdscribe(raw, 8, 8, **kw))).pin_memory(),
                        window=5, returnDiff=False)
    assert_array_equal(only_t.numpy(), S.timeSOAP(O.normalizeArray(O.fillSOAPVectorFromdscribe(raw, 8, 8, **kw)), window=5, returnDiff=False))
    # a pageable host tensor keeps the one-shot path, also with host results
    t2, dt2 = S.timeSOAPfromCompressed(torch.from_numpy(raw), 8, 8, windows=(5,), **kw)[0]
    assert not t2.is_cuda
    assert_array_equal(t2.numpy(), S.timeSOAPfromCompressed(raw, 8, 8, windows=(5,), **kw)[0][0])


def test_large_pageable_arrays_are_staged_in_pieces(S, monkeypatch):
    """numpy / pageable tensors above 64 MB go to HBM through two fixed pinned buffers (no whole-array pin)."""
    import torch
    from cpctools_b200 import _device

    monkeypatch.setattr(_device._HostStager, "MIN_BYTES", 1 << 16)
    monkeypatch.setattr(_device._HostStager, "CHUNK", (1 << 20) + 24)  # several pieces, a ragged last one
    monkeypatch.setattr(_device, "_stagers", {})  # fresh stagers (pinned buffers of the patched size)
    rng = np.random.default_rng(5)
    for arr in (rng.random((37, 11, 1224)), rng.random((3, 70001)).astype(np.float32), rng.integers(0, 9, (5, 40001))):
        dev = _device.to_device(arr)
        assert dev.is_cuda and tuple(dev.shape) == arr.shape
        assert_array_equal(dev.cpu().numpy(), arr)
        ro = arr.copy()
        ro.setflags(write=False)
        assert_array_equal(_device.to_device(ro).cpu().numpy(), arr)
    t = torch.from_numpy(rng.random((64, 4099)))
    assert_array_equal(_device.to_device(t).cpu().numpy(), t.numpy())
    for shape, dt in (((37, 11, 1224), torch.float64), ((3, 70001), torch.float32), ((7, 30011), torch.int32)):
        d = (torch.rand(shape, device="cuda") * 100).to(dt)
        back = _device.to_host(d)  # staged download, several ragged pieces
        assert back.shape == shape and back.flags.writeable
        assert_array_equal(back, d.cpu().numpy())
    nc = torch.rand(5, 40000, device="cuda", dtype=torch.float64).T  # non-contiguous source
    assert_array_equal(_device.to_host(nc), nc.cpu().numpy())
    attrs, dc = dscribe_attrs(8, 8, ["H", "O"])
    kw = dict(atomTypes=["H", "O"], atomicSlices=_slices_from_attrs(attrs))
    raw = synthetic_soap((40, 9, dc), seed=2, kind="W")
    monkeypatch.setattr(_device, "_stagers", {})  # fresh stagers (pinned buffers of the patched size)
    got = S.timeSOAPfromCompressed(raw, 8, 8, windows=(1, 5), **kw)
    monkeypatch.setattr(_device._HostStager, "MIN_BYTES", 1 << 40)
    want = S.timeSOAPfromCompressed(raw, 8, 8, windows=(1, 5), **kw)
    for (t1, d1), (t2, d2) in zip(got, want):
        assert_array_equal(t1, t2)
        assert_array_equal(d1, d2)
    monkeypatch.setattr(_device, "_stagers", {})  # fresh stagers (pinned buffers of the patched size)


def test_config1_shape_against_the_loop_oracle(S):
    """BASELINE configs[0]: 309 atoms x 200 frames, 1 species (Dc 324 -> Df 576), fp64, the reference's
    own Python-loop path (oracle port) as the checker, default distance and SOAPdistanceNormalized."""
    raw = synthetic_soap((200, 309, 324), seed=12345, kind="U")
    X = O.normalizeArray(O.fillSOAPVectorFromdscribe(raw, 8, 8))
    Xg = S.normalizeArray(S.fillSOAPVectorFromdscribe(raw, 8, 8))
    assert_allclose(Xg, X, rtol=1e-12, atol=0)
    t, dt = S.timeSOAP(Xg, window=1)
    want_t, want_dt = O.timeSOAP(X, window=1)  # Python double loop + scipy cosine
    assert t.shape == (199, 309) and dt.shape == (309, 198)
    assert_distance_parity(t, want_t, "f64", what="config1 simple")
    assert_allclose(dt, want_dt, rtol=0, atol=1e-11)
    tn = S.timeSOAP(Xg, window=1, returnDiff=False, distanceFunction=S.SOAPdistanceNormalized)
    assert_distance_parity(tn, O.timeSOAP(X, window=1, returnDiff=False, distanceFunction=O.SOAPdistanceNormalized), "f64")
    (tc,) = S.timeSOAPfromCompressed(raw, 8, 8, windows=(1,), returnDiff=False)
    assert_distance_parity(tc, want_t, "f64", what="config1 fused")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_config5_quippy_three_species_fused(S, dtype):
    """BASELINE configs[4]: raw quippy rows (3 species, nMax = lMax = 6: 1198 columns incl. the trailing
    covariance column, quippy slot order) -> timeSOAP without materialising anything, fp64 and fp32
    (fp32 rows are not 16-byte multiples: cp.async staging path)."""
    sp = ["H", "C", "O"]
    attrs, dc = dscribe_attrs(6, 6, sp)
    sl = _slices_from_attrs(attrs)
    raw = synthetic_soap((45, 7, dc + 1), seed=3, kind="W", dtype=dtype)
    addr = O.getAddressesQuippyLikeDscribe(6, 6, sp)
    permuted = raw.astype(np.float64)[..., :dc][..., addr]            # engine.py:226,260
    X = O.normalizeArray(O.fillSOAPVectorFromdscribe(permuted, 6, 6, sp, sl))
    path = "f64" if dtype == np.float64 else "f32_timelag"
    got = S.timeSOAPfromCompressed(raw, 6, 6, ["O", "H", "C"], sl, windows=(1, 10), quippy=True)
    for w, (t, dt) in zip((1, 10), got):
        want, wdt = O.timeSOAP_batched(X, window=w)
        assert t.shape == want.shape and dt.shape == wdt.shape
        assert_distance_parity(t, want, path, what=f"quippy w{w}")
    full = S.fillSOAPVectorFromQuip(raw, 6, 6, sp, sl)
    assert_array_equal(full, O.fillSOAPVectorFromdscribe(raw[..., :dc][..., addr], 6, 6, sp, sl))


def test_streaming_shortcut_matches_golden(S, golden):
    g = golden("timesoap")
    attrs1 = {"l_max": 8, "n_max": 8, "species": ["H"], "species_location_H-H": [0, 324]}
    t, dt = S.getTimeSOAPSimple(FakeSOAPDataset(g["stream_raw"], 10, attrs1))
    assert t.shape == (46, 5) and dt.shape == (5, 45)
    assert_allclose(t, g["stream_t"], rtol=1e-12, atol=1e-15)
    assert_allclose(dt, g["stream_dt"], rtol=0, atol=1e-13)
    attrs2, _ = dscribe_attrs(4, 4, ["H", "O"])
    t2, dt2 = S.getTimeSOAPSimple(FakeSOAPDataset(g["stream2_raw"], 6, attrs2))
    assert_allclose(t2, g["stream2_t"], rtol=1e-12, atol=1e-15)
    # window > 1: the reference raises; here the chunks are stitched with a `window`-frame halo
    raw = g["stream_raw"]
    t5, _ = S.getTimeSOAPSimple(FakeSOAPDataset(raw, 10, attrs1), window=5)
    Xn = O.normalizeArray(O.fillSOAPVectorFromdscribe(raw, 8, 8))
    assert_allclose(t5, np.linalg.norm(Xn[5:] - Xn[:-5], axis=-1), rtol=1e-12, atol=1e-15)
    # chunk shorter than the window
    t7, _ = S.getTimeSOAPSimple(FakeSOAPDataset(raw, 3, attrs1), window=7)
    assert_allclose(t7, np.linalg.norm(Xn[7:] - Xn[:-7], axis=-1), rtol=1e-12, atol=1e-15)


# ------------------------------------------------------------------ distance matrix / classification
def test_distance_matrix_matches_golden(S, golden):
    g = golden("classify")
    data, spectra = g["data"], g["spectra"]
    K = _kinds(S)
    assert_distance_parity(S.getDistanceBetween(data, spectra, S.SOAPdistanceNormalized), g["dm_normalized"], "f64")
    assert_distance_parity(S.getDistanceBetween(data, spectra, S.simpleSOAPdistance), g["dm_simple"], "f64")
    assert_distance_parity(S.getDistanceBetween(data, spectra, K["kernel2"][0]), g["dm_kernel2"], "f64")
    assert_distance_parity(S.getDistanceBetween(data, data, S.SOAPdistance), g["dm_self_kernel1"], "f64")
    assert_distance_parity(S.getDistanceBetween(data, data, S.SOAPdistanceNormalized), g["dm_self_normalized"], "f64")
    d32 = S.getDistanceBetween(data.astype(np.float32), spectra.astype(np.float32), S.SOAPdistanceNormalized)
    assert d32.dtype == np.float32 and d32.shape == (37, 9)
    assert_distance_parity(d32, g["dm_normalized_f32"], "f32_ref")


@pytest.mark.parametrize("M,K,D", [(1, 1, 3), (65, 130, 50), (300, 7, 576), (129, 257, 1728)])
def test_distance_matrix_shapes_vs_oracle(S, M, K, D):
    rng = np.random.default_rng(M + K + D)
    a, b = rng.random((M, D)), rng.random((K, D))
    for name, (fn, kind, power) in _kinds(S).items():
        src_a, src_b = (O.normalizeArray(a), O.normalizeArray(b)) if name == "normalized" else (a, b)
        want = O.getDistanceBetween_batched(src_a, src_b, kind, power)
        assert_distance_parity(S.getDistanceBetween(src_a, src_b, fn), want, "f64", what=name)


def test_classification_matches_golden(S, golden):
    g = golden("classify")
    attrs = {"l_max": 8, "n_max": 8, "species": ["Au"], "species_location_Au-Au": [0, 324]}
    ds = FakeSOAPDataset(g["cls_raw"], 1000, attrs)
    refs = S.SOAPReferences(["a", "b", "c", "d"], g["cls_ref_spectra"], 8, 8)
    d = S.getDistancesFromRefNormalized(ds, refs)
    assert d.shape == g["cls_dist"].shape and d.dtype == np.float64
    assert_distance_parity(d, g["cls_dist"], "f64")
    cl = S.applyClassification(ds, refs, S.SOAPdistanceNormalized, doNormalize=True)
    assert cl.legend == ["a", "b", "c", "d"]
    assert_distance_parity(cl.distances, g["cls_min"], "f64")
    # argmin may only differ where two references are equidistant to rounding
    differ = cl.references != g["cls_arg"]
    if differ.any():
        srt = np.sort(g["cls_dist"], axis=-1)
        assert (np.abs(srt[..., 1] ** 2 - srt[..., 0] ** 2)[differ] <= 1e-12).all()
    assert_distance_parity(S.getDistancesFromRef(ds, refs, S.simpleSOAPdistance, doNormalize=False), g["cls_dist_simple"], "f64")
    # references built through the GPU expand/normalise match the reference-built ones
    built = S.createReferencesFromTrajectory(ds, {"a": (0, 0), "b": (100, 3), "c": (129, 1), "d": (17, 2)}, 8, 8)
    assert_allclose(built.spectra, g["cls_ref_spectra"], rtol=1e-12, atol=0)
    part = S.getDistancesFromInterval(ds, refs, S.SOAPdistanceNormalized, True, slice(20, 57))
    assert_distance_parity(part, g["cls_dist"][20:57], "f64")
    with pytest.raises(ValueError, match="contiguous"):
        S.getDistancesFromInterval(ds, refs, S.SOAPdistanceNormalized, True, slice(0, 40, 2))


def test_classification_fused_and_unfused_paths_agree(S, golden, monkeypatch):
    """the fused kernel (expansion folded into the references) against the two-kernel path and the
    oracle, for all three distance kinds, normalised or not, f64 and f32 data."""
    from cpctools_b200 import classify

    g = golden("classify")
    attrs = {"l_max": 8, "n_max": 8, "species": ["Au"], "species_location_Au-Au": [0, 324]}
    refs = S.SOAPReferences(["a", "b", "c", "d"], g["cls_ref_spectra"], 8, 8)
    orefs = type("R", (), {"names": refs.names, "spectra": refs.spectra, "lmax": 8, "nmax": 8, "__len__": lambda self: 4})()
    K = _kinds(S)
    for dtype, path in ((np.float64, "f64"), (np.float32, "f32")):
        raw = g["cls_raw"][:40].astype(dtype)
        ds = FakeSOAPDataset(raw, 16, attrs)
        for name, norm in (("normalized", True), ("simple", False), ("kernel2", False), ("simple", True)):
            fn, okind, power = K[name]
            ofn = {"normalized": O.SOAPdistanceNormalized, "simple": O.simpleSOAPdistance,
                   "kernel2": lambda a, b: O.SOAPdistance(a, b, 2)}[name]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                want = O.getDistancesFromRef(FakeSOAPDataset(raw.astype(np.float64), 16, attrs), orefs, ofn, doNormalize=norm)
            fused = S.getDistancesFromRef(ds, refs, fn, doNormalize=norm)
            assert_distance_parity(fused, want, path, what=f"fused {name} {dtype.__name__}")
            cl = S.applyClassification(ds, refs, fn, doNormalize=norm)
            assert_array_equal(cl.references, np.argmin(fused, axis=-1))
            assert_array_equal(cl.distances, np.amin(fused, axis=-1))
            monkeypatch.setattr(classify, "FUSED_MAX_REFS", 0)  # force the expand -> normalise -> pairs path
            unfused = S.getDistancesFromRef(ds, refs, fn, doNormalize=norm)
            monkeypatch.setattr(classify, "FUSED_MAX_REFS", 32)
            assert_distance_parity(unfused, want, path, what=f"unfused {name} {dtype.__name__}")
            if dtype is np.float32:
                # classify.py:113 stores every frame's distances as data.dtype: both paths return float32 values
                assert_array_equal(fused, fused.astype(np.float32).astype(np.float64))
                assert_array_equal(unfused, unfused.astype(np.float32).astype(np.float64))


@pytest.mark.parametrize("nmax,lmax,K", [(8, 8, 1), (8, 8, 9), (8, 8, 20), (8, 8, 32), (8, 8, 40), (5, 2, 6), (1, 0, 3)])
def test_classification_shapes_vs_oracle(S, nmax, lmax, K):
    """dictionary sizes across the kernel's register variants (<= 8 / 16 / 32 references), beyond them
    (two-kernel path), rows that are not 16-byte multiples (warp-per-vector kernel), several tiles."""
    rng = np.random.default_rng(nmax * 100 + lmax * 10 + K)
    dc = (lmax + 1) * nmax * (nmax + 1) // 2
    raw = synthetic_soap((23, 19, dc), seed=K, kind="W")
    attrs = {"l_max": lmax, "n_max": nmax, "species": ["Au"], "species_location_Au-Au": [0, dc]}
    ds = FakeSOAPDataset(raw, 7, attrs)
    spectra = O.normalizeArray(rng.random((K, (lmax + 1) * nmax * nmax)))
    refs = S.SOAPReferences([f"r{k}" for k in range(K)], spectra, lmax, nmax)
    orefs = type("R", (), {"names": refs.names, "spectra": spectra, "lmax": lmax, "nmax": nmax, "__len__": lambda self: K})()
    if dc == (lmax + 1) * nmax * nmax:  # nMax = 1: nothing to expand, the dataset is used as it is
        assert dc == 1
    want = O.getDistancesFromRefNormalized(ds, orefs)
    got = S.getDistancesFromRefNormalized(ds, refs)
    assert got.shape == want.shape == (23, 19, K)
    assert_distance_parity(got, want, "f64", what=f"n{nmax} l{lmax} K{K}")
    cl = S.applyClassification(ds, refs, S.SOAPdistanceNormalized, doNormalize=True)
    assert_array_equal(cl.references, np.argmin(got, axis=-1))
    assert_array_equal(cl.distances, np.amin(got, axis=-1))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_classification_pads_unaligned_rows_for_the_tiled_kernel(S, dtype):
    """Rows that are not whole 16-byte units (mono-species n = l = 6: Dc 147) would take soap_classify's slow
    warp-per-vector kernel; from 4096 vectors per block upwards the host pads them once and runs the tiled kernel.
    Both routes against the oracle, and the same arg-minima from both."""
    from cpctools_b200 import classify as C

    lmax = nmax = 6
    dc, K = (lmax + 1) * nmax * (nmax + 1) // 2, 7
    assert (dc * np.dtype(dtype).itemsize) % 16 != 0
    raw = synthetic_soap((40, 128, dc), seed=3, kind="W", dtype=dtype)       # 5120 vectors in one block
    attrs = {"l_max": lmax, "n_max": nmax, "species": ["Au"], "species_location_Au-Au": [0, dc]}
    rng = np.random.default_rng(9)
    spectra = O.normalizeArray(rng.random((K, (lmax + 1) * nmax * nmax)))
    refs = S.SOAPReferences([f"r{k}" for k in range(K)], spectra, lmax, nmax)
    orefs = type("R", (), {"names": refs.names, "spectra": spectra, "lmax": lmax, "nmax": nmax, "__len__": lambda self: K})()
    ds = FakeSOAPDataset(raw, 40, attrs)                                     # one chunk -> padded route
    small = FakeSOAPDataset(raw, 8, attrs)                                   # 1024 vectors per block -> fallback kernel
    want = O.getDistancesFromRefNormalized(FakeSOAPDataset(raw.astype(np.float64), 40, attrs), orefs)
    path = "f64" if dtype == np.float64 else "f32"
    got_pad = S.getDistancesFromRefNormalized(ds, refs)
    got_fb = S.getDistancesFromRefNormalized(small, refs)
    assert_distance_parity(got_pad, want, path, what="padded rows, tiled kernel")
    assert_distance_parity(got_fb, want, path, what="warp-per-vector kernel")
    a, b = S.applyClassification(ds, refs, S.SOAPdistanceNormalized, doNormalize=True), S.applyClassification(small, refs, S.SOAPdistanceNormalized, doNormalize=True)
    assert (a.references == b.references).mean() > 0.999  # (a near-tie may flip between two summation orders)


def test_argmin_numpy_rules(S):
    """ties -> first index; NaN -> first NaN (numpy.argmin / amin, classify.py:274-275)."""
    import torch
    from cpctools_b200.classify import pairs_device
    from cpctools_b200 import _lib

    rng = np.random.default_rng(3)
    data = rng.random((70, 40))
    refs = rng.random((9, 40))
    refs[5] = refs[2]  # exact tie
    refs[7] = 0.0      # zero reference -> NaN under the clipped cosine
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = O.getDistanceBetween_batched(data, refs, O.KIND_SIMPLE, 1)
    amin, arg = pairs_device(torch.from_numpy(data).cuda(), torch.from_numpy(refs).cuda(), _lib.KIND_SIMPLE, 1, want_argmin=True)
    assert (arg.cpu().numpy() == np.argmin(want, axis=-1)).all() and (arg.cpu().numpy() == 7).all()
    assert np.isnan(amin.cpu().numpy()).all()
    amin, arg = pairs_device(torch.from_numpy(data).cuda(), torch.from_numpy(refs[:7]).cuda(), _lib.KIND_SIMPLE, 1, want_argmin=True)
    full = S.getDistanceBetween(data, refs[:7], S.simpleSOAPdistance)
    assert_array_equal(arg.cpu().numpy(), np.argmin(full, axis=-1))
    assert_array_equal(amin.cpu().numpy(), np.amin(full, axis=-1))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind_name", ["normalized", "simple", "kernel2"])
def test_row_minima_fused_into_the_tensor_core_epilogue(S, dtype, kind_name):
    """soap_pairs(path=TENSOR, out=NULL, argmin_out, min_out): the minima kept by the tcgen05 epilogue equal
    numpy.argmin / numpy.amin of the matrix the same path writes -- ties, NaNs, ragged tile edges included --
    and soap_pairs_rowmin's exclude_offset drops the self distance of a row block."""
    import torch
    from cpctools_b200.classify import pairs_device
    from cpctools_b200 import _lib

    fn, okind, power = _kinds(S)[kind_name]
    kind = {"normalized": _lib.KIND_NORMALIZED, "simple": _lib.KIND_SIMPLE, "kernel2": _lib.KIND_KERNEL}[kind_name]
    rng = np.random.default_rng(11)
    Y = O.normalizeArray(rng.random((3001, 576))).astype(dtype)
    Y[1500] = Y[77]          # exact tie: the earlier column wins
    Y[2999] = Y[78]
    if kind_name != "normalized":
        Y[2100] = 0.0        # zero vector -> NaN under the cosine kinds: the first NaN wins
    X = Y[640 : 640 + 777].copy()
    Xd, Yd = torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda()
    full = pairs_device(Xd, Yd, kind, power, path=_lib.PAIRS_TENSOR).cpu().numpy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        amin, arg = pairs_device(Xd, Yd, kind, power, want_argmin=True)
        assert_array_equal(arg.cpu().numpy(), np.argmin(full, axis=-1))
        assert_array_equal(amin.cpu().numpy(), np.amin(full, axis=-1).astype(np.float64))
        if kind_name == "simple":  # (the unclipped kernel may already give NaN on a row's own column)
            assert (arg.cpu().numpy() == 2100).all() and np.isnan(amin.cpu().numpy()).all()
        # the matrix AND the minima in one call
        out = torch.empty((777, 3001), dtype=Xd.dtype, device="cuda")
        amin2, arg2 = pairs_device(Xd, Yd, kind, power, out=out, want_argmin=True)
        assert_array_equal(out.cpu().numpy(), full)
        assert_array_equal(arg2.cpu().numpy(), arg.cpu().numpy())
        # nearest OTHER vector: row i of the block is vector 640 + i
        if kind_name == "normalized":
            amin3, arg3 = pairs_device(Xd, Yd, kind, power, want_argmin=True, exclude_offset=640)
            masked = full.astype(np.float64).copy()
            masked[np.arange(777), 640 + np.arange(777)] = np.inf
            assert_array_equal(arg3.cpu().numpy(), np.argmin(masked, axis=-1))
            assert_array_equal(amin3.cpu().numpy(), np.amin(masked, axis=-1))
    # against the oracle by the parity rule (independent of the matrix path)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = O.getDistanceBetween_batched(X.astype(np.float64), Y.astype(np.float64), okind, power)
    if kind_name == "normalized":
        assert_distance_parity(amin.cpu().numpy(), np.amin(want, axis=-1), "f64" if dtype == np.float64 else "f32")


# ------------------------------------------------------------------ size-independent properties at scale
def test_properties_at_benchmark_row_sizes(S):
    """Config-2/3 vectors (2 species, Dc 1224 -> Df 1728) on device-generated data too large for the
    Python oracle: (i) expansion == gather by the table, checked by multiplicity-weighted sums;
    (ii) normalised rows have unit norm; (iii) the distances are invariant under per-vector
    rescaling and under the expansion itself (compressed path == materialised path);
    (iv) a trajectory that repeats with period p has t_p == 0 exactly under the clipped kernel."""
    import torch
    from cpctools_b200 import _lib
    from cpctools_b200.analysis import timelag_device

    attrs, dc = dscribe_attrs(8, 8, ["H", "O"])
    sl = _slices_from_attrs(attrs)
    F, A = 96, 600
    raw = torch.empty((F, A, dc), dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().soap_fill_uniform(raw.data_ptr(), raw.numel(), 12345, _lib.F64, None))
    torch.cuda.synchronize()
    plan = S.getExpansionPlan(8, 8, ["H", "O"], sl)
    full = S.fillSOAPVectorFromdscribe(raw, 8, 8, ["H", "O"], sl)
    mult = plan.multiplicity_device(torch.float64)
    assert torch.allclose(full.sum(-1), (raw * mult).sum(-1), rtol=1e-13, atol=0)
    assert torch.equal(full, raw[..., torch.from_numpy(plan.index).cuda()])
    nrm = S.fillAndNormalize(raw, 8, 8, ["H", "O"], sl)
    assert torch.allclose(nrm.norm(dim=-1), torch.ones(F, A, dtype=torch.float64, device="cuda"), rtol=1e-14, atol=0)
    windows = (1, 5, 10, 50)
    t_mat = timelag_device(nrm, windows, _lib.KIND_SIMPLE)
    t_cmp = timelag_device(raw, windows, _lib.KIND_SIMPLE, mult=mult)
    scale = torch.rand(F, A, 1, dtype=torch.float64, device="cuda") + 0.5
    t_scl = timelag_device(raw * scale, windows, _lib.KIND_SIMPLE, mult=mult)
    for a, b, c in zip(t_mat, t_cmp, t_scl):
        assert_distance_parity(b.cpu().numpy(), a.cpu().numpy(), "f64")
        assert_distance_parity(c.cpu().numpy(), a.cpu().numpy(), "f64")
    periodic = raw[:5].repeat(12, 1, 1)
    (t5,) = timelag_device(periodic, (5,), _lib.KIND_SIMPLE, mult=mult)
    assert (t5 == 0).all()


# ------------------------------------------------------------------ tensor-core all-pairs paths
def _lib_kind(name):
    from cpctools_b200 import _lib

    return {"simple": _lib.KIND_SIMPLE, "kernel1": _lib.KIND_KERNEL, "kernel2": _lib.KIND_KERNEL, "normalized": _lib.KIND_NORMALIZED}[name]


@pytest.mark.parametrize("path", ["int8", "native"])
@pytest.mark.parametrize("M,K,D", [(128, 256, 576), (300, 257, 576), (513, 300, 1728), (200, 333, 50)])
def test_pairs_tensor_fp64_vs_oracle(S, M, K, D, path):
    """fp64 all-pairs on tensor cores -- tcgen05 INT8 with 6 exact digit planes (default) and the
    FP64 DMMA GEMM (native) -- under the same rule as every fp64 path."""
    import torch
    from cpctools_b200 import _lib
    from cpctools_b200.classify import pairs_device

    which = _lib.PAIRS_TENSOR if path == "int8" else _lib.PAIRS_TENSOR_NATIVE
    rng = np.random.default_rng(M * 7 + K)
    a, b = rng.random((M, D)), rng.random((K, D))
    a[3] *= 1.0e-3  # rows of very different scale: every row has its own digit scale
    b[5] *= 4.0e4
    b[7, ::2] *= -1.0  # signed entries
    for name, (fn, kind, power) in _kinds(S).items():
        sa, sb = (O.normalizeArray(a), O.normalizeArray(b)) if name == "normalized" else (a, b)
        want = O.getDistanceBetween_batched(sa, sb, kind, power)
        got = pairs_device(torch.from_numpy(sa).cuda(), torch.from_numpy(sb).cuda(), _lib_kind(name), power, path=which)
        assert got.dtype == torch.float64
        assert_distance_parity(got.cpu().numpy(), want, "f64", what=f"{path} {name}")


@pytest.mark.parametrize("path", ["int8", "native"])
def test_pairs_tensor_self_distance_and_symmetry(S, path):
    import torch
    from cpctools_b200 import _lib
    from cpctools_b200.classify import pairs_device

    which = _lib.PAIRS_TENSOR if path == "int8" else _lib.PAIRS_TENSOR_NATIVE
    X = torch.nn.functional.normalize(torch.rand(700, 576, dtype=torch.float64, device="cuda"), dim=-1)
    d = pairs_device(X, X, _lib.KIND_NORMALIZED, 1, path=which)
    assert torch.equal(d, d.T)  # x.y and y.x are the same sum
    # the reference gives a few 1e-8 here, never NaN; the rule is d^2 <= 1e-12
    assert (torch.diagonal(d) ** 2 <= 1.0e-12).all() and (torch.diagonal(d) <= 2.0e-7).all()
    ds = pairs_device(X, X, _lib.KIND_SIMPLE, 1, path=which)
    assert (torch.diagonal(ds) <= 2.0e-7).all() and not torch.isnan(ds).any()


@pytest.mark.parametrize("M,K,D", [(128, 256, 576), (300, 600, 576), (517, 333, 1728), (1000, 1031, 50)])
def test_pairs_tensor_fp32_vs_oracle(S, M, K, D):
    """fp32 all-pairs through tcgen05 INT8 (3 exact digit planes) against the fp64 oracle of the same
    float32 inputs: the fp32-path rule (d^2 to 1e-6 absolute, d to 1e-6 relative where d >= 0.5)."""
    import torch
    from cpctools_b200 import _lib
    from cpctools_b200.classify import pairs_device

    rng = np.random.default_rng(M + 3 * K)
    a = rng.random((M, D)).astype(np.float32)
    b = rng.random((K, D)).astype(np.float32)
    b[2, 1::2] *= -1.0
    an = (a / np.linalg.norm(a, axis=-1, keepdims=True)).astype(np.float32)
    bn = (b / np.linalg.norm(b, axis=-1, keepdims=True)).astype(np.float32)
    for name, (fn, kind, power) in _kinds(S).items():
        sa, sb = (an, bn) if name == "normalized" else (a, b)
        want = O.getDistanceBetween_batched(sa.astype(np.float64), sb.astype(np.float64), kind, power)
        got = pairs_device(torch.from_numpy(sa).cuda(), torch.from_numpy(sb).cuda(), _lib_kind(name), power, path=_lib.PAIRS_TENSOR)
        assert got.dtype == torch.float32
        assert_distance_parity(got.cpu().numpy(), want, "f32", what=f"int8x3 {name}")


def test_pairs_tensor_many_tiles_persistent(S):
    """more tiles than SMs (persistent loop, accumulator hand-over, ragged edges), both dtypes; the
    3xTF32 native path is held to its documented 5e-5 (fp32 round-toward-zero accumulation in TMEM)."""
    import torch
    from cpctools_b200 import _lib
    from cpctools_b200.classify import pairs_device

    X = torch.nn.functional.normalize(torch.rand(2900, 576, dtype=torch.float64, device="cuda"), dim=-1)
    Y = torch.nn.functional.normalize(torch.rand(3700, 576, dtype=torch.float64, device="cuda"), dim=-1)
    want = torch.sqrt(torch.abs(2.0 - 2.0 * (X @ Y.T))).cpu().numpy()
    got64 = pairs_device(X, Y, _lib.KIND_NORMALIZED, 1, path=_lib.PAIRS_TENSOR)
    assert_distance_parity(got64.cpu().numpy(), want, "f64", what="int8x6 persistent")
    X32, Y32 = X.float(), Y.float()
    want32 = torch.sqrt(torch.abs(2.0 - 2.0 * (X32.double() @ Y32.double().T))).cpu().numpy()
    got32 = pairs_device(X32, Y32, _lib.KIND_NORMALIZED, 1, path=_lib.PAIRS_TENSOR)
    assert_distance_parity(got32.cpu().numpy(), want32, "f32", what="int8x3 persistent")
    tf32 = pairs_device(X32, Y32, _lib.KIND_NORMALIZED, 1, path=_lib.PAIRS_TENSOR_NATIVE).cpu().numpy().astype(np.float64)
    assert np.abs(tf32**2 - want32**2).max() <= 5.0e-5


@pytest.mark.parametrize("dtype", ["f32", "f64"])
@pytest.mark.parametrize("M,K,D", [(300, 256, 576), (129, 333, 64), (1000, 1777, 324), (257, 4000, 100)])
def test_pairs_cta_pairs_equal_single_ctas(S, monkeypatch, dtype, M, K, D):
    """cta_group::2 (SOAP_I8_CG=2, opt-in): two CTAs share one tcgen05.mma per pass, each staging its own rows of A
    and half of the B tile.  Same exact integer sums, same epilogue: the matrix and the fused row minima must be
    BIT-identical to the single-CTA kernel, ragged tiles included (odd 128-row tile counts leave a CTA without rows)."""
    import torch
    from cpctools_b200 import _lib
    from cpctools_b200.classify import pairs_device

    tdt = torch.float32 if dtype == "f32" else torch.float64
    g = torch.Generator(device="cuda").manual_seed(M + K)
    X = torch.nn.functional.normalize(torch.rand(M, D, dtype=tdt, device="cuda", generator=g) - 0.3, dim=-1)
    Y = torch.nn.functional.normalize(torch.rand(K, D, dtype=tdt, device="cuda", generator=g) - 0.3, dim=-1)
    monkeypatch.delenv("SOAP_I8_CG", raising=False)
    one = pairs_device(X, Y, _lib.KIND_NORMALIZED, 1, path=_lib.PAIRS_TENSOR)
    am1 = pairs_device(X, Y, _lib.KIND_NORMALIZED, 1, path=_lib.PAIRS_TENSOR, want_argmin=True)
    monkeypatch.setenv("SOAP_I8_CG", "2")
    two = pairs_device(X, Y, _lib.KIND_NORMALIZED, 1, path=_lib.PAIRS_TENSOR)
    am2 = pairs_device(X, Y, _lib.KIND_NORMALIZED, 1, path=_lib.PAIRS_TENSOR, want_argmin=True)
    torch.cuda.synchronize()
    assert torch.equal(one, two)
    for a, b in zip(am1, am2):
        assert torch.equal(a, b)
    want = torch.sqrt(torch.abs(2.0 - 2.0 * (X.double() @ Y.double().T))).cpu().numpy()
    assert_distance_parity(two.cpu().numpy(), want, dtype, what="cta pairs")


@pytest.mark.parametrize("n", [1000, 1337])
def test_symmetric_all_pairs_every_kind(S, n):
    """data is spectra: upper-triangular tiles + mirrored stores; ragged sizes, all three kinds, both dtypes."""
    rng = np.random.default_rng(n)
    X = rng.random((n, 200))
    X[5] *= 3.0e3
    X[9, ::3] *= -1.0
    Xn = O.normalizeArray(X)
    for name, (fn, kind, power) in _kinds(S).items():
        src = Xn if name == "normalized" else X
        want = O.getDistanceBetween_batched(src, src, kind, power)
        got = S.getDistanceBetween(src, src, fn)
        assert_distance_parity(got, want, "f64", what=f"sym {name}")
        assert_array_equal(got, got.T)
        s32 = src.astype(np.float32)
        got32 = S.getDistanceBetween(s32, s32, fn)
        assert_distance_parity(got32, O.getDistanceBetween_batched(s32.astype(np.float64), s32.astype(np.float64), kind, power), "f32", what=f"sym32 {name}")


def test_get_distance_between_picks_tensor_path_for_large_inputs(S):
    """the public all-pairs call (numpy in, numpy out) on a matrix large enough for the tensor path."""
    rng = np.random.default_rng(11)
    X = O.normalizeArray(rng.random((1500, 576)))
    d = S.getDistanceBetween(X, X, S.SOAPdistanceNormalized)
    assert d.shape == (1500, 1500) and d.dtype == np.float64
    assert_distance_parity(d, O.getDistanceBetween_batched(X, X, O.KIND_NORMALIZED, 1), "f64")
    d32 = S.getDistanceBetween(X.astype(np.float32), X.astype(np.float32), S.SOAPdistanceNormalized)
    assert d32.dtype == np.float32
    assert_distance_parity(d32, O.getDistanceBetween_batched(X.astype(np.float32).astype(np.float64), X.astype(np.float32).astype(np.float64), O.KIND_NORMALIZED, 1), "f32")


# ------------------------------------------------------------------ transitions (row f4): integer, bit-exact
def test_transition_matrix_matches_golden_and_oracle(S, golden):
    g = golden("transitions")
    refs = g["refs"]
    data = S.SOAPclassification(np.zeros(refs.shape), refs, ["a", "b", "c", "d", "e"])
    for stride, window in ((1, None), (1, 5), (3, 3), (2, 7), (1, 59), (1, 60)):
        tag = f"s{stride}_w{window}"
        tm = S.transitionMatrixFromSOAPClassification(data, stride, window)
        assert tm.dtype == np.float64
        assert_array_equal(tm, g["tm_" + tag])
        assert_array_equal(S.transitionMatrixFromSOAPClassificationNormalized(data, stride, window), g["tmn_" + tag])
    with pytest.raises(ValueError, match="the window must be bigger than the stride"):
        S.transitionMatrixFromSOAPClassification(data, 3, 2)
    with pytest.raises(ValueError, match="stride and window must be smaller"):
        S.transitionMatrixFromSOAPClassification(data, 1, 61)
    # many classes (global-atomic path) and a large random classification against the oracle
    rng = np.random.default_rng(1)
    big = rng.integers(0, 120, size=(300, 2000))
    bd = S.SOAPclassification(np.zeros(big.shape), big, [str(i) for i in range(120)])
    assert_array_equal(S.transitionMatrixFromSOAPClassification(bd, 2, 4), O.transitionMatrixFromSOAPClassification(big, 120, 2, 4))
    small = rng.integers(0, 7, size=(1000, 3000))
    sd = S.SOAPclassification(np.zeros(small.shape), small, list("abcdefg"))
    tm = S.transitionMatrixFromSOAPClassification(sd)
    assert_array_equal(tm, O.transitionMatrixFromSOAPClassification(small, 7))
    assert tm.sum() == 999 * 3000
    # -1 ("unclassified") is numpy's index of the LAST class in the reference (transitions/__init__.py:44-47);
    # a state outside [-n_classes, n_classes) raises IndexError there and here
    err = g["refs_err"]
    ed = S.SOAPclassification(np.zeros(err.shape), err, ["a", "b", "c", "d", "e", "error"])
    for stride, window in ((1, None), (2, 7)):
        assert_array_equal(S.transitionMatrixFromSOAPClassification(ed, stride, window), g[f"tm_err_s{stride}_w{window}"])
        assert_array_equal(S.transitionMatrixFromSOAPClassificationNormalized(ed, stride, window), g[f"tmn_err_s{stride}_w{window}"])
    assert_array_equal(S.transitionMatrixFromSOAPClassification(ed, 3, 5), O.transitionMatrixFromSOAPClassification(err, 6, 3, 5))
    with pytest.raises(IndexError):
        S.transitionMatrixFromSOAPClassification(S.SOAPclassification(np.zeros(err.shape), err, list("abcd")))
    with pytest.raises(IndexError):
        S.transitionMatrixFromSOAPClassification(S.SOAPclassification(np.zeros(err.shape), err - 6, list("abcdef")))


# ------------------------------------------------------------------ randomised shapes (seeded fuzz)
def test_residence_times_match_golden_and_oracle(S, golden):
    g = golden("transitions")
    cases = ((1, None), (3, None), (4, 2), (5, 1), (60, 7), (59, 59))
    for name, key in (("", "refs"), ("slow_", "refs_slow")):
        refs = g[key]
        data = S.SOAPclassification(np.zeros(refs.shape), refs, ["a", "b", "c", "d", "e"])
        for window, stride in cases:
            got = S.calculateResidenceTimesFromClassification(data, window=window, stride=stride)
            assert len(got) == 5
            for state, times in enumerate(got):
                want = g[f"rt_{name}w{window}_s{stride}_{state}"]
                assert_array_equal(times, want)
                assert times.dtype == want.dtype
        assert_array_equal(np.concatenate(S.calculateResidenceTimes(data, window=3)),
                           np.concatenate(S.calculateResidenceTimesFromClassification(data, window=3)))
    data = S.SOAPclassification(np.zeros(g["refs"].shape), g["refs"], ["a", "b", "c", "d", "e"])
    with pytest.raises(ValueError, match="the window must be bigger than the stride"):
        S.calculateResidenceTimesFromClassification(data, window=2, stride=3)
    with pytest.raises(ValueError, match="stride and window must be smaller"):
        S.calculateResidenceTimesFromClassification(data, window=61)
    with pytest.raises(NotImplementedError):
        S.calculateResidenceTimes(data, statesTracker=[object()])
    assert_array_equal(S.calculateTransitionMatrix(data, stride=2, window=7), g["tm_s2_w7"])
    # a large slowly changing classification against the oracle; a state that never occurs gives the
    # reference's empty float64 array; every frame of every series is accounted for
    rng = np.random.default_rng(3)
    big = np.repeat(rng.integers(0, 6, size=(40, 1500)), 7, axis=0)  # 280 frames, runs of >= 7
    bd = S.SOAPclassification(np.zeros(big.shape), big, list("abcdefg"))
    for window, stride in ((1, None), (2, 1), (7, 3)):
        got = S.calculateResidenceTimesFromClassification(bd, window=window, stride=stride)
        want = O.calculateResidenceTimesFromClassification(big, 7, window=window, stride=stride)
        for a, b in zip(got, want):
            assert_array_equal(a, b)
            assert a.dtype == b.dtype
        assert got[6].size == 0 and got[6].dtype == np.float64


def test_fuzz_random_shapes_against_oracle(S):
    """40 seeded random cases over odd/even dimensions, tiny and ragged frame / atom counts, one to
    four windows, both dtypes, with and without expansion: the time-lag kernel (both staging paths),
    the expansion kernel and the SIMT pair kernel against the oracle."""
    rng = np.random.default_rng(20261020 + _fuzz_offset())
    for case in range(40):
        nmax, lmax = int(rng.integers(1, 7)), int(rng.integers(0, 6))
        ns = int(rng.integers(1, 4))
        sp = [None, ["H", "O"], ["H", "C", "O"]][ns - 1]
        if sp is None:
            kw, dc = {}, (lmax + 1) * nmax * (nmax + 1) // 2
        else:
            attrs, dc = dscribe_attrs(lmax, nmax, sp)
            kw = dict(atomTypes=sp, atomicSlices=_slices_from_attrs(attrs))
        F, A = int(rng.integers(2, 90)), int(rng.integers(1, 40))
        dtype = [np.float64, np.float32][case % 2]
        path = "f64" if dtype == np.float64 else "f32_timelag"
        raw = synthetic_soap((F, A, dc), seed=case, kind="W" if case % 3 else "U", dtype=dtype)
        want_full = O.fillSOAPVectorFromdscribe(raw, lmax, nmax, **kw)
        assert_array_equal(S.fillSOAPVectorFromdscribe(raw, lmax, nmax, **kw), want_full)
        Xn = O.normalizeArray(want_full.astype(np.float64))
        fn = S.fillAndNormalize(raw, lmax, nmax, **kw)
        assert fn.dtype == dtype
        # 1e-12 / 1e-6 against the exactly normalised vectors; float32 data additionally against the reference's OWN
        # float32 arithmetic, which carries two more fp32 roundings (norm and quotient): 2e-6
        assert_allclose(fn, Xn, rtol=1e-12 if dtype == np.float64 else 1e-6, atol=0)
        assert_allclose(fn, O.normalizeArray(want_full), rtol=1e-12 if dtype == np.float64 else 2e-6, atol=0)
        nw = int(rng.integers(1, 5))
        windows = sorted(set(int(w) for w in rng.integers(1, F, size=nw)))
        got = S.timeSOAPfromCompressed(raw, lmax, nmax, windows=windows, returnDiff=False, **kw)
        for w, t in zip(windows, got):
            assert_distance_parity(t, O.timeSOAP_batched(Xn, window=w, returnDiff=False), path, what=f"case {case} w{w}")
        M, K = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        a, b = Xn.reshape(-1, Xn.shape[-1])[:M].astype(dtype), Xn.reshape(-1, Xn.shape[-1])[-K:].astype(dtype)
        gotm = S.getDistanceBetween(a, b, S.SOAPdistanceNormalized)
        assert gotm.dtype == dtype
        assert_distance_parity(gotm, O.getDistanceBetween_batched(a.astype(np.float64), b.astype(np.float64)), path, what=f"case {case} matrix")


def test_fuzz_tensor_core_pairs_and_row_minima(S):
    """16 seeded random shapes through the tcgen05 INT8 path (single CTAs and the opt-in CTA pairs): ragged tile
    edges in both directions, odd dimensions (K padded to the block), rows of very different magnitude, some signed
    entries (the digit planes are fixed point per ROW: the error of a dot is ~2^-23 max|x| max|y| sqrt(D), so strongly
    short float32 rows are at the edge of the 1e-6 rule -- uniform(-0.1, 0.9) rows of D = 11 reach 1.06e-6 on d -- and
    the automatic dispatch keeps rows under 32 elements on the SIMT kernel); the matrix against the oracle, the fused row minima against numpy on the matrix the same path wrote,
    with and without an excluded diagonal."""
    import os
    import torch
    from cpctools_b200 import _lib
    from cpctools_b200.classify import pairs_device

    rng = np.random.default_rng(77 + _fuzz_offset())
    names = ["normalized", "simple", "kernel2"]
    for case in range(16):
        M, K, D = int(rng.integers(128, 900)), int(rng.integers(256, 1500)), int(rng.integers(32, 700))
        dtype = [np.float64, np.float32][case % 2]
        name = names[case % 3]
        fn, okind, power = _kinds(S)[name]
        a, b = rng.random((M, D)) - 0.1, rng.random((K, D)) - 0.1
        a[int(rng.integers(0, M))] *= 1.0e-4
        b[int(rng.integers(0, K))] *= 3.0e3
        if name == "normalized":
            a, b = O.normalizeArray(a), O.normalizeArray(b)
        a, b = a.astype(dtype), b.astype(dtype)
        ad, bd = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        want = O.getDistanceBetween_batched(a.astype(np.float64), b.astype(np.float64), okind, power)
        prev = os.environ.pop("SOAP_I8_CG", None)
        try:
            if case % 4 == 3:
                os.environ["SOAP_I8_CG"] = "2"
            full = pairs_device(ad, bd, _lib_kind(name), power, path=_lib.PAIRS_TENSOR).cpu().numpy()
            assert_distance_parity(full, want, "f64" if dtype == np.float64 else "f32", what=f"case {case} {name} {M}x{K}x{D}")
            if M * K >= 1_000_000:  # (smaller problems take the SIMT minimum)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    amin, arg = pairs_device(ad, bd, _lib_kind(name), power, want_argmin=True)
                    assert_array_equal(arg.cpu().numpy(), np.argmin(full, axis=-1))
                    assert_array_equal(amin.cpu().numpy(), np.amin(full, axis=-1).astype(np.float64))
                    off = int(rng.integers(0, max(1, K - M)))
                    amin2, arg2 = pairs_device(ad, bd, _lib_kind(name), power, want_argmin=True, exclude_offset=off)
                    masked = full.astype(np.float64).copy()
                    rows = np.arange(M)
                    keep = rows + off < K
                    masked[rows[keep], rows[keep] + off] = np.inf
                    assert_array_equal(arg2.cpu().numpy(), np.argmin(masked, axis=-1))
                    assert_array_equal(amin2.cpu().numpy(), np.amin(masked, axis=-1))
        finally:
            os.environ.pop("SOAP_I8_CG", None)
            if prev is not None:
                os.environ["SOAP_I8_CG"] = prev


def test_fuzz_timelag_raw_dimensions_and_kinds(S):
    """30 seeded random trajectories of ARBITRARY row length (odd, 8- and 16-byte multiples alike: the TMA path, the
    shifted-window path and the element-copy path), every distance kind, one to four windows, both dtypes,
    with and without multiplicities, against the batched oracle."""
    import torch
    from cpctools_b200 import _lib
    from cpctools_b200.analysis import timelag_device

    rng = np.random.default_rng(4242 + _fuzz_offset())
    kinds = [("simple", _lib.KIND_SIMPLE, O.KIND_SIMPLE, 1), ("kernel1", _lib.KIND_KERNEL, O.KIND_KERNEL, 1),
             ("kernel2", _lib.KIND_KERNEL, O.KIND_KERNEL, 2), ("normalized", _lib.KIND_NORMALIZED, O.KIND_NORMALIZED, 1)]
    for case in range(30):
        F, A, D = int(rng.integers(2, 140)), int(rng.integers(1, 200)), int(rng.integers(1, 420))
        dtype = [np.float64, np.float32][case % 2]
        path = "f64" if dtype == np.float64 else "f32_timelag"
        name, kind, okind, power = kinds[case % 4]
        X = (rng.random((F, A, D)) + 0.05).astype(dtype)
        X64 = X.astype(np.float64)
        if name == "normalized":
            X64 = O.normalizeArray(X64)
            X = X64.astype(dtype)
            X64 = X.astype(np.float64)
        nw = int(rng.integers(1, 5))
        windows = tuple(sorted(set(int(w) for w in rng.integers(1, F, size=nw))))
        use_mult = case % 3 == 0
        m = (1.0 + (rng.random(D) < 0.5)).astype(dtype) if use_mult else None
        ts = timelag_device(torch.from_numpy(X).cuda(), windows, kind, power,
                            mult=None if m is None else torch.from_numpy(m).cuda())
        Xw = X64 if m is None else X64 * np.sqrt(m.astype(np.float64))
        for w, t in zip(windows, ts):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                want, _ = O.timeSOAP_batched(Xw, window=w, kind=okind, power=power)
            assert_distance_parity(t.cpu().numpy(), want, path, what=f"case {case} {name} F{F} A{A} D{D} w{w}")


def test_fuzz_timelag_chain_kernel(S, monkeypatch):
    """24 seeded random trajectories through the lane = units sweep kernel (timelag_chain.cuh) forced on: whole-unit
    rows from a single unit to five pieces, any number of frames above the largest window (one block, odd and even
    block counts), few and many atoms per CTA, every cosine-type kind, both dtypes, ring windows 1..25 in either
    pattern and any order, with and without multiplicities -- against the batched oracle, and the kernel must really
    have been the one that ran."""
    import torch
    from cpctools_b200 import _lib
    from cpctools_b200.analysis import timelag_device

    monkeypatch.setenv("SOAP_TL_CHAIN", "1")
    rng = np.random.default_rng(777 + _fuzz_offset())
    kinds = [("simple", _lib.KIND_SIMPLE, O.KIND_SIMPLE, 1), ("kernel1", _lib.KIND_KERNEL, O.KIND_KERNEL, 1),
             ("kernel2", _lib.KIND_KERNEL, O.KIND_KERNEL, 2), ("normalized", _lib.KIND_NORMALIZED, O.KIND_NORMALIZED, 1)]
    ring_choices = [w for w in range(1, 26) if w % 5]
    for case in range(24):
        dtype = [np.float64, np.float32][case % 2]
        ue = 2 if dtype == np.float64 else 4
        four = case % 3 != 0
        w0 = int(rng.choice(ring_choices)) if case % 4 else 1
        windows = [w0, 5, 10] + ([50] if four else [])
        if dtype == np.float64 and case % 8 >= 4:  # the further float64 patterns (chain_choose)
            windows = [[w0, 5, 50], [w0, 10, 50], [w0, 5, 10, 25], [w0, 5, 10, 20], [5, 10, 50], [5, 10, 25]][int(rng.integers(0, 6))]
        rng.shuffle(windows)
        windows = tuple(int(w) for w in windows)
        F = int(rng.integers(max(windows) + 1, max(windows) + 120))
        A = int(rng.integers(1, 40)) if case % 5 else int(rng.integers(300, 700))
        units = int(rng.choice([1, 7, 31, 32, 33, 100, 128, 129, 200, 306, 500, 612, 640]))
        if A > 100:
            units = min(units, 64)
        D = units * ue
        path = "f64" if dtype == np.float64 else "f32_timelag"
        name, kind, okind, power = kinds[case % 4]
        X = (rng.random((F, A, D)) + 0.05).astype(dtype)
        X64 = X.astype(np.float64)
        if name == "normalized":
            X64 = O.normalizeArray(X64)
            X = X64.astype(dtype)
            X64 = X.astype(np.float64)
        use_mult = case % 3 != 1
        m = (1.0 + (rng.random(D) < 0.5)).astype(dtype) if use_mult else None
        ts = timelag_device(torch.from_numpy(X).cuda(), windows, kind, power,
                            mult=None if m is None else torch.from_numpy(m).cuda())
        assert _lib.load().soap_timelag_last_kernel() == b"timelag_chain_kernel", (case, windows, F, A, D)
        Xw = X64 if m is None else X64 * np.sqrt(m.astype(np.float64))
        for w, t in zip(windows, ts):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                want, _ = O.timeSOAP_batched(Xw, window=w, kind=okind, power=power)
            assert t.shape == (F - w, A)
            assert_distance_parity(t.cpu().numpy(), want, path, what=f"case {case} {name} F{F} A{A} D{D} windows {windows} w{w}")
