#!/usr/bin/env python
"""Accuracy of every CUDA path against the CPU oracle on one B200 -> profiles/rNN_accuracy.txt

    python scripts/accuracy_report.py > profiles/r01_accuracy.txt

For each path: max |d^2 - d^2_ref| (the quantity the parity rule bounds: 1e-12 fp64, 1e-6 fp32) and
max relative error of d where d >= 0.1 / 0.5, on benchmark-shaped random data.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpctools_b200 as S  # noqa: E402
from cpctools_b200 import _lib  # noqa: E402
from cpctools_b200.classify import pairs_device  # noqa: E402
from oracle import soap_oracle as O  # noqa: E402
from oracle.fake_dataset import dscribe_attrs, synthetic_soap  # noqa: E402


def err(got, want, dmin):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    ok = ~(np.isnan(got) | np.isnan(want))
    e2 = np.abs(got[ok] ** 2 - want[ok] ** 2).max()
    big = ok & (want >= dmin)
    rel = (np.abs(got[big] - want[big]) / want[big]).max() if big.any() else 0.0
    return e2, rel


def line(name, e2, rel, tol):
    print(f"{name:70s} max|d2-d2ref| {e2:9.2e}   max rel(d) {rel:9.2e}   tol {tol:g}   {'OK' if e2 <= tol and rel <= tol else 'FAIL'}")


def main():
    print("# accuracy of the CUDA paths against the CPU oracle (B200, seeded data)")
    attrs, dc = dscribe_attrs(8, 8, ["H", "O"])
    sl = {k[len("species_location_"):].replace("-", ""): slice(*v) for k, v in attrs.items() if k.startswith("species_location_")}
    for kind in ("U", "W"):
        raw = synthetic_soap((120, 64, dc), seed=1, kind=kind)
        X = O.normalizeArray(O.fillSOAPVectorFromdscribe(raw, 8, 8, ["H", "O"], sl))
        for dt, path, tol, dmin in ((np.float64, "f64", 1e-12, 0.1), (np.float32, "f32", 1e-6, 0.5)):
            got = S.timeSOAPfromCompressed(raw.astype(dt), 8, 8, ["H", "O"], sl, windows=(1, 5, 10, 50), returnDiff=False)
            ref = O.normalizeArray(O.fillSOAPVectorFromdscribe(raw.astype(dt).astype(np.float64), 8, 8, ["H", "O"], sl))
            for w, t in zip((1, 5, 10, 50), got):
                e2, rel = err(t, O.timeSOAP_batched(ref, window=w, returnDiff=False), dmin)
                line(f"timeSOAP fused-from-compressed  data {kind} {path} window {w}", e2, rel, tol)
            # the same through the lane = units kernel (timelag_chain.cuh; the default for float64 sweeps over many atoms)
            os.environ["SOAP_TL_CHAIN"] = "1"
            got = S.timeSOAPfromCompressed(raw.astype(dt), 8, 8, ["H", "O"], sl, windows=(1, 5, 10, 50), returnDiff=False)
            del os.environ["SOAP_TL_CHAIN"]
            assert _lib.load().soap_timelag_last_kernel() == b"timelag_chain_kernel"
            for w, t in zip((1, 5, 10, 50), got):
                e2, rel = err(t, O.timeSOAP_batched(ref, window=w, returnDiff=False), dmin)
                line(f"timeSOAP fused-from-compressed  data {kind} {path} window {w}  (chain kernel)", e2, rel, tol)
        t = S.timeSOAP(X, window=1, returnDiff=False)
        e2, rel = err(t, O.timeSOAP_batched(X, window=1, returnDiff=False), 0.1)
        line(f"timeSOAP materialised 1728       data {kind} f64 window 1", e2, rel, 1e-12)
    rng = np.random.default_rng(3)
    A = O.normalizeArray(rng.random((3000, 576)))
    B = O.normalizeArray(rng.random((2500, 576)))
    want = O.getDistanceBetween_batched(A, B)
    ta, tb = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    for pname, p in (("SIMT", _lib.PAIRS_SIMT), ("DMMA (native)", _lib.PAIRS_TENSOR_NATIVE), ("tcgen05 int8 x6", _lib.PAIRS_TENSOR)):
        e2, rel = err(pairs_device(ta, tb, _lib.KIND_NORMALIZED, 1, path=p).cpu().numpy(), want, 0.1)
        line(f"all-pairs 3000x2500x576 fp64 {pname}", e2, rel, 1e-12)
    A32, B32 = A.astype(np.float32), B.astype(np.float32)
    want32 = O.getDistanceBetween_batched(A32.astype(np.float64), B32.astype(np.float64))
    ta, tb = torch.from_numpy(A32).cuda(), torch.from_numpy(B32).cuda()
    for pname, p, tol in (("SIMT", _lib.PAIRS_SIMT, 1e-6), ("tcgen05 int8 x3", _lib.PAIRS_TENSOR, 1e-6), ("tcgen05 3xTF32 (native, opt-in)", _lib.PAIRS_TENSOR_NATIVE, 5e-5)):
        e2, rel = err(pairs_device(ta, tb, _lib.KIND_NORMALIZED, 1, path=p).cpu().numpy(), want32, 0.5)
        line(f"all-pairs 3000x2500x576 fp32 {pname}", e2, rel, tol)
    cpu32 = np.sqrt(np.abs(2 - 2 * (A32 @ B32.T)))
    e2, rel = err(cpu32, want32, 0.5)
    line("  (for scale: the reference's own float32 arithmetic, numpy sgemm)", e2, rel, 1e-6)
    return 0


if __name__ == "__main__":
    sys.exit(main())
