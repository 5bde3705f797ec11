"""CPU-side tests of the product: host logic (tables, layouts, attrs, validation, dispatch) and that
the C-ABI library builds, loads and exports every symbol include/soap_b200.h declares.  No compute
call is made (there is no GPU here, and the product has no CPU fallback -- which is also tested).
"""
import functools
import os
import re

import numpy as np
import pytest
from numpy.testing import assert_array_equal

import cpctools_b200 as S
from cpctools_b200 import _lib
from oracle.fake_dataset import FakeSOAPDataset, dscribe_attrs
from oracle.gen_golden import LMAX, NMAX, SPECIES_SETS, _slices_from_attrs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    import torch

    return torch.cuda.is_available()


@pytest.mark.parametrize("ns", sorted(SPECIES_SETS))
def test_product_tables_bit_exact_with_reference_goldens(golden, ns):
    g = golden("tables")
    species = SPECIES_SETS[ns]
    ordered = S.orderByZ(species)
    for n in NMAX:
        for l in LMAX:
            tag = f"S{ns}_n{n}_l{l}"
            attrs, dc = dscribe_attrs(l, n, ordered)
            kw = {} if ns == 1 else dict(atomTypes=species, atomicSlices=_slices_from_attrs(attrs))
            plan = S.getExpansionPlan(l, n, **kw)
            assert plan.index.dtype == np.int64
            assert_array_equal(plan.index, g["fill_" + tag])
            assert plan.d_compressed == dc and plan.d_full == len(g["fill_" + tag])
            assert plan.multiplicity.sum() == plan.d_full
            assert_array_equal(S.getAddressesQuippyLikeDscribe(l, n, species), g["quip_" + tag])
            if ns > 1:  # quippy-composed table == permutation then fill
                q = S.getExpansionPlan(l, n, ordered, _slices_from_attrs(attrs), quippy=True)
                assert q.d_compressed == dc + 1 and q.multiplicity[-1] == 0
                assert_array_equal(q.index, g["quip_" + tag][g["fill_" + tag]])


def test_label_layouts(golden):
    g = golden("tables")
    for ns, n, l in ((2, 2, 1), (3, 6, 6)):
        sp = SPECIES_SETS[ns]
        assert_array_equal(S.getdscribeSOAPMapping(l, n, sp), g[f"labels_dscribe_S{ns}_n{n}_l{l}"])
        assert_array_equal(S.getdscribeSOAPMapping(l, n, sp, crossover=False), g[f"labels_dscribe_nocross_S{ns}_n{n}_l{l}"])
        assert_array_equal(S.getquippySOAPMapping(l, n, sp), g[f"labels_quippy_S{ns}_n{n}_l{l}"])
    assert S.orderByZ(["Au", "H", "C", "W", "O"]) == ["H", "C", "O", "W", "Au"]


def test_plan_cache_and_species_order():
    attrs, _ = dscribe_attrs(8, 8, ["H", "O"])
    sl = _slices_from_attrs(attrs)
    a = S.getExpansionPlan(8, 8, ["O", "H"], sl)
    b = S.getExpansionPlan(8, 8, np.array(["H", "O"]), sl)  # ndarray of str, other order (utils.py:304-306)
    assert a is b
    assert_array_equal(a.index[576:580], [324, 325, 326, 327])
    assert_array_equal(a.index[1152:1160], np.arange(900, 908))
    assert a.index.max() == 1223
    assert set(np.unique(a.multiplicity)) == {1, 2}


def test_settings_from_attrs():
    attrs, dc = dscribe_attrs(4, 4, ["H", "O"])
    ds = FakeSOAPDataset(np.zeros((3, 2, dc)), 2, attrs)
    st = S.getSOAPSettings(ds)
    assert st["nMax"] == 4 and st["lMax"] == 4 and list(st["atomTypes"]) == ["H", "O"]
    up, full = 5 * 10, 5 * 16
    assert st["atomicSlices"]["HH"] == slice(0, up)
    assert st["atomicSlices"]["HO"] == slice(up, up + full) == st["atomicSlices"]["OH"]
    assert st["atomicSlices"]["OO"] == slice(up + full, 2 * up + full)


def test_validation_happens_before_any_device_work():
    """Error conventions of SURVEY.md section 8(b); all raised on the host, GPU or not."""
    X = np.zeros((10, 2, 4))
    with pytest.raises(ValueError, match="the window must be bigger than the stride"):
        S.timeSOAP(X, window=2, stride=3)
    with pytest.raises(ValueError, match="stride and window must be smaller than simulation lenght"):
        S.timeSOAP(X, window=10)
    with pytest.raises(ValueError, match="stride and window must be smaller than simulation lenght"):
        S.timeSOAPsimple(X, window=3, stride=None, backward=True) if False else S.timeSOAPsimple(X, window=12)
    with pytest.raises(NotImplementedError):
        S.timeSOAP(X, distanceFunction=lambda a, b: 0.0)
    with pytest.raises(ValueError):
        S.fillSOAPVectorFromdscribe(np.zeros((5, 325)), 8, 8)
    with pytest.raises(ValueError):
        S.fillSOAPVectorFromdscribe(np.zeros((3, 4, 5, 324)), 8, 8)
    a = S.SOAPReferences(["a"], np.zeros((1, 2)), 8, 8)
    b = S.SOAPReferences(["b"], np.zeros((1, 2)), 7, 8)
    with pytest.raises(ValueError):
        S.mergeReferences(a, b)
    m = S.mergeReferences(a, S.SOAPReferences(["c", "d"], np.ones((2, 2)), 8, 8))
    assert len(m) == 3 and m.spectra.shape == (3, 2)


def test_distance_dispatch():
    from cpctools_b200.distances import resolve_distance

    assert resolve_distance(S.simpleSOAPdistance) == (_lib.KIND_SIMPLE, 1)
    assert resolve_distance(S.SOAPdistance) == (_lib.KIND_KERNEL, 1)
    assert resolve_distance(functools.partial(S.SOAPdistance, n=3)) == (_lib.KIND_KERNEL, 3)
    assert resolve_distance(S.kernelSOAPdistance) == (_lib.KIND_KERNEL, 1)
    assert resolve_distance(S.SOAPdistanceNormalized) == (_lib.KIND_NORMALIZED, 1)
    with pytest.raises(NotImplementedError):
        resolve_distance(np.dot)
    # wrappers (functools.wraps), partial chains, kernelSoap needs its power like the reference
    @functools.wraps(S.SOAPdistanceNormalized)
    def logged(x, y):
        return S.SOAPdistanceNormalized(x, y)

    assert resolve_distance(logged) == (_lib.KIND_NORMALIZED, 1)
    assert resolve_distance(functools.partial(functools.partial(S.SOAPdistance, n=2), n=5)) == (_lib.KIND_KERNEL, 5)
    assert resolve_distance(functools.partial(S.kernelSoap, n=4)) == (_lib.KIND_DOT, 4)
    with pytest.raises(TypeError):
        resolve_distance(functools.partial(S.kernelSoap))
    with pytest.raises(NotImplementedError):
        resolve_distance(functools.partial(S.SOAPdistance, 1.0))
    with pytest.raises(NotImplementedError):
        resolve_distance(functools.partial(S.simpleSOAPdistance, n=2))
    with pytest.raises(NotImplementedError):
        resolve_distance(lambda x, y: 0.0)
    # the reference package's own functions (what an existing script imports) are matched by module + name
    ref = type("f", (), {"__module__": "SOAPify.distances", "__name__": "simpleSOAPdistance", "__call__": lambda s, x, y: 0})()
    assert resolve_distance(ref) == (_lib.KIND_SIMPLE, 1)
    assert S.tempoSOAP.__doc__ and S.getDistancesFromInterval and S.fillSOAPVectorFromQuip


def test_library_builds_loads_and_exports_the_header():
    path = _lib.build_library()
    assert os.path.exists(path)
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "soap_b200.h")).read()
    declared = set(re.findall(r"\b(soap_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.soap_abi_version() == 1
    # argument validation works without a device and reports through soap_last_error()
    assert lib.soap_expand(None, None, None, -1, 4, 4, 8, None) == -1
    assert b"soap_expand" in lib.soap_last_error()


@pytest.mark.skipif(_has_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    """The product path must fail loudly when there is no CUDA device."""
    x = np.ones((2, 3, 324))
    for call in (
        lambda: S.fillSOAPVectorFromdscribe(x, 8, 8),
        lambda: S.normalizeArray(x),
        lambda: S.timeSOAP(x),
        lambda: S.simpleSOAPdistance(x[0, 0], x[0, 1]),
        lambda: S.getDistanceBetween(x[0], x[1], S.SOAPdistanceNormalized),
    ):
        with pytest.raises(S.SoapB200Error, match="no CPU fallback"):
            call()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cpctools_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn
            assert "/root/reference" not in src, fn


def test_references_round_trip_through_an_hdf5_like_group():
    """saveReferences / getReferencesFromDataset only use the h5py duck type (classify.py:208-247)."""
    import cpctools_b200 as S

    class _Attrs(dict):
        def create(self, key, value):
            self[key] = np.asarray(value)

    class _Dataset:
        def __init__(self, shape, dtype, **kw):
            self.data, self.attrs, self.kw = np.zeros(shape, dtype=dtype), _Attrs(), kw

        def __setitem__(self, key, value):
            self.data[key] = value

        def __getitem__(self, key):
            return np.array(self.data[key])

    class _Group(dict):
        def require_dataset(self, name, shape, dtype, **kw):
            self[name] = _Dataset(shape, dtype, **kw)
            return self[name]

    rng = np.random.default_rng(0)
    refs = S.SOAPReferences(["bulk", "surface", "vertex"], rng.random((3, 576)), 8, 8)
    grp = _Group()
    S.saveReferences(grp, "myRefs", refs)
    assert grp["myRefs"].kw == {"compression": "gzip", "compression_opts": 9}
    back = S.getReferencesFromDataset(grp["myRefs"])
    assert back.names == refs.names and back.lmax == 8 and back.nmax == 8 and len(back) == 3
    np.testing.assert_array_equal(back.spectra, refs.spectra)


def test_bench_parses_the_gpu_cpu_affinity_table():
    """bench.py binds every rank to the cores next to its GPU: the parser of `nvidia-smi topo -m`."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    one = "\t\x1b[4mGPU0\tCPU Affinity\tNUMA Affinity\tGPU NUMA ID\x1b[0m\nGPU0\t X \t0-15\t0\t\tN/A\n\nLegend:\n"
    assert bench.parse_topo_cpu_affinity(one, 0) == list(range(16))
    assert bench.parse_topo_cpu_affinity(one, 1) is None
    two = ("\tGPU0\tGPU1\tNIC0\tCPU Affinity\tNUMA Affinity\tGPU NUMA ID\n"
           "GPU0\t X \tNV18\tPIX\t0-3,64-67\t0\t\tN/A\n"
           "GPU1\tNV18\t X \tSYS\t32-35\t1\t\tN/A\n"
           "NIC0\tPIX\tSYS\t X \n")
    assert bench.parse_topo_cpu_affinity(two, 0) == [0, 1, 2, 3, 64, 65, 66, 67]
    assert bench.parse_topo_cpu_affinity(two, 1) == [32, 33, 34, 35]
    assert bench.parse_topo_cpu_affinity("no table here", 0) is None
