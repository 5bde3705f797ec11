#!/usr/bin/env python
"""Per-kernel roofline table on one B200 (CUDA events around each main kernel, inputs >> L2).

    python scripts/kernel_bench.py [--quick]

Prints one JSON line per kernel/config: achieved GB/s (algorithmic bytes / event time) against the
measured HBM copy peak, or TFLOP/s for the pair kernels.  Used to fill DESIGN.md / profiles/.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpctools_b200 as S  # noqa: E402
from cpctools_b200 import _lib  # noqa: E402
from cpctools_b200.analysis import timelag_device  # noqa: E402
from cpctools_b200.classify import pairs_device  # noqa: E402


def peak_hbm():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


def fill(shape, dtype, seed=1):
    x = torch.empty(shape, dtype=dtype, device="cuda")
    _lib.check(_lib.load().soap_fill_uniform(x.data_ptr(), x.numel(), seed, _lib.F64 if dtype == torch.float64 else _lib.F32, None))
    return x


def timed(fn, prof_id, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    ms = _lib.profile_read(prof_id)
    _lib.profile_enable(False)
    return float(np.median(ms)), len(ms) // reps


def slices2(l=8, n=8):
    up, full = (l + 1) * n * (n + 1) // 2, (l + 1) * n * n
    return {"HH": slice(0, up), "HO": slice(up, up + full), "OO": slice(up + full, 2 * up + full)}, 2 * up + full


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    peak = peak_hbm()
    rows = []

    def report(name, cfg, ms, nbytes=None, flops=None, **extra):
        row = {"kernel": name, "config": cfg, "ms": round(ms, 4)}
        if nbytes is not None:
            row.update(GBps=round(nbytes / ms / 1e6, 1), frac_hbm=round(nbytes / ms / 1e6 / peak, 3))
        if flops is not None:
            row.update(TFLOPs=round(flops / ms / 1e9, 2))
        row.update(extra)
        rows.append(row)
        print(json.dumps(row), flush=True)

    sl2, dc2 = slices2()
    want = lambda k: not args.only or k in args.only.split(",")

    # ---- expansion / normalisation (configs[1]: 2 species, 1224 -> 1728), 100-frame tiles of 10k atoms
    if want("expand"):
        F, A = (20, 10000) if args.quick else (100, 10000)
        for dt, es in ((torch.float64, 8), (torch.float32, 4)):
            raw = fill((F, A, dc2), dt)
            nvec = F * A
            ms, _ = timed(lambda: S.fillSOAPVectorFromdscribe(raw, 8, 8, ["H", "O"], sl2), _lib.PROF_EXPAND)
            report("expand", f"{F}x{A} Dc1224->Df1728 {dt}", ms, nvec * (1224 + 1728) * es, vec_per_s=nvec / ms * 1e3)
            ms, _ = timed(lambda: S.fillAndNormalize(raw, 8, 8, ["H", "O"], sl2), _lib.PROF_EXPAND)
            report("expand+normalize", f"{F}x{A} Dc1224->Df1728 {dt}", ms, nvec * (1224 + 1728) * es, vec_per_s=nvec / ms * 1e3)
            full = S.fillSOAPVectorFromdscribe(raw, 8, 8, ["H", "O"], sl2)
            del raw
            ms, _ = timed(lambda: S.normalizeArray(full), _lib.PROF_EXPAND)
            report("normalize", f"{F}x{A} D1728 {dt}", ms, nvec * 2 * 1728 * es, vec_per_s=nvec / ms * 1e3)
            del full
        raw = fill((200 if not args.quick else 50, 20000, 324), torch.float64)
        nvec = raw.shape[0] * raw.shape[1]
        ms, _ = timed(lambda: S.fillAndNormalize(raw, 8, 8), _lib.PROF_EXPAND)
        report("expand+normalize", f"{raw.shape[0]}x20000 Dc324->Df576 f64", ms, nvec * (324 + 576) * 8, vec_per_s=nvec / ms * 1e3)
        del raw

    # ---- time-lag kernel
    if want("timelag"):
        F, A = (1000, 1500) if args.quick else (1000, 4000)
        plan = S.getExpansionPlan(8, 8, ["H", "O"], sl2)
        for dt, es, code in ((torch.float64, 8, _lib.F64), (torch.float32, 4, _lib.F32)):
            raw = fill((F, A, dc2), dt)
            mult = plan.multiplicity_device(dt)
            for wins in ((1,), (50,), (1, 5), (1, 5, 10, 50)):
                ms, _ = timed(lambda: timelag_device(raw, wins, _lib.KIND_SIMPLE, 1, mult=mult), _lib.PROF_TIMELAG)
                nb = F * A * dc2 * es + sum((F - w) * A * 8 for w in wins)
                report("timelag(compressed)", f"{F}x{A}x1224 {dt} windows={wins}", ms, nb, pairs_per_s=sum((F - w) * A for w in wins) / ms * 1e3)
            del raw
        X = fill((F, A // 2, 1728), torch.float64)
        for wins in ((1,), (1, 5, 10, 50)):
            ms, _ = timed(lambda: timelag_device(X, wins, _lib.KIND_SIMPLE, 1), _lib.PROF_TIMELAG)
            nb = X.numel() * 8 + sum((F - w) * (A // 2) * 8 for w in wins)
            report("timelag(materialised)", f"{F}x{A // 2}x1728 f64 windows={wins}", ms, nb)
        ms, _ = timed(lambda: timelag_device(X, (1,), _lib.KIND_EUCLID, 1), _lib.PROF_TIMELAG)
        report("timelag(euclid)", f"{F}x{A // 2}x1728 f64 windows=(1,)", ms, X.numel() * 8 + (F - 1) * (A // 2) * 8)
        del X
        # configs[4]: raw quippy rows, 3 species nMax = lMax = 6 (1198 columns), composed table
        from oracle.fake_dataset import dscribe_attrs
        attrs, dcq = dscribe_attrs(6, 6, ["H", "C", "O"])
        slq = {k[len("species_location_"):].replace("-", ""): slice(*v) for k, v in attrs.items() if k.startswith("species_location_")}
        planq = S.getExpansionPlan(6, 6, ["H", "C", "O"], slq, quippy=True)
        for dt, es in ((torch.float64, 8), (torch.float32, 4)):
            Xq = fill((1000, 4096, dcq + 1), dt)
            ms, _ = timed(lambda: timelag_device(Xq, (1,), _lib.KIND_SIMPLE, 1, mult=planq.multiplicity_device(dt)), _lib.PROF_TIMELAG)
            report("timelag(quippy raw)", f"1000x4096x1198 {dt} windows=(1,)", ms, Xq.numel() * es + 999 * 4096 * 8)
            del Xq
        X = fill((2000, 309 * 8, 324), torch.float64)
        ms, _ = timed(lambda: timelag_device(X, (1,), _lib.KIND_SIMPLE, 1, mult=S.getExpansionPlan(8, 8).multiplicity_device(torch.float64)), _lib.PROF_TIMELAG)
        report("timelag(compressed)", "2000x2472x324 f64 windows=(1,)", ms, X.numel() * 8)
        del X

    # ---- classification against a small dictionary (f1): fused kernel straight from compressed rows
    if want("classify"):
        F, A, K = (50, 20000, 8) if args.quick else (200, 50000, 8)
        for Dc, Df, kw in ((324, 576, {}),):
            raw = fill((F * A, Dc), torch.float64)
            plan = S.getExpansionPlan(8, 8)
            refs = np.random.default_rng(0).random((K, Df))
            refs /= np.linalg.norm(refs, axis=-1, keepdims=True)
            folded = torch.from_numpy(np.stack([np.bincount(plan.index, weights=r, minlength=Dc) for r in refs])).cuda()
            norm2 = torch.ones(K, dtype=torch.float64, device="cuda")
            mult = plan.multiplicity_device(torch.float64)
            amin = torch.empty(F * A, dtype=torch.float64, device="cuda")
            arg = torch.empty(F * A, dtype=torch.int32, device="cuda")
            lib = _lib.load()
            fn = lambda: _lib.check(lib.soap_classify(raw.data_ptr(), mult.data_ptr(), folded.data_ptr(), norm2.data_ptr(), F * A, Dc, K,
                                                      _lib.KIND_NORMALIZED, 1, 1, _lib.F64, None, arg.data_ptr(), amin.data_ptr(), None))
            ms, _ = timed(fn, _lib.PROF_PAIRS)
            report("classify(fused)", f"{F}x{A} Dc{Dc} K={K} f64 argmin", ms, F * A * (Dc * 8 + 12), vec_per_s=F * A / ms * 1e3)
            del raw

    # ---- the step after the classification (f4): transition counts and residence times of int[F, A] states
    if want("transitions"):
        F, A, K = (500, 20000, 8) if args.quick else (2000, 100000, 8)
        states = torch.randint(0, K, (F // 20, A), dtype=torch.int32, device="cuda").repeat_interleave(20, dim=0).contiguous()
        lib = _lib.load()
        counts = torch.empty((K, K), dtype=torch.int64, device="cuda")

        def events(fn, reps=5):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        ms = events(lambda: _lib.check(lib.soap_transition_counts(states.data_ptr(), F, A, K, 1, 1, counts.data_ptr(), None)))
        report("transition_counts", f"{F}x{A} int32 states, {K} classes, window 1", ms, F * A * 4, pairs_per_s=(F - 1) * A / ms * 1e3,
               note="bytes = every state once (window == stride: the `from` state of a step is the previous `to`); round 1 counted both reads")
        cap = int(lib.soap_residence_capacity(F, A, 1, 1))
        ev_s = torch.empty(cap, dtype=torch.int32, device="cuda")
        ev_t = torch.empty(cap, dtype=torch.int64, device="cuda")
        n_ev = torch.zeros(1, dtype=torch.int64, device="cuda")
        ms = events(lambda: _lib.check(lib.soap_residence_events(states.data_ptr(), F, A, K, 1, 1, ev_s.data_ptr(), ev_t.data_ptr(), cap,
                                                                 n_ev.data_ptr(), None)))
        report("residence_events", f"{F}x{A} int32 states, window 1 ({int(n_ev.item())} events)", ms, F * A * 4 + int(n_ev.item()) * 12,
               states_per_s=F * A / ms * 1e3)
        del states, ev_s, ev_t

    # ---- chunked dataset -> pinned ring -> HBM -> kernels (f2): getTimeSOAPSimple on an HDF5-like dataset
    if want("stream"):
        import time
        from oracle.fake_dataset import FakeSOAPDataset, dscribe_attrs
        F, A = (200, 2000) if args.quick else (400, 5000)
        attrs, dc = dscribe_attrs(8, 8, ["H", "O"])
        host = np.random.default_rng(1).random((F, A, dc))
        ds = FakeSOAPDataset(host, 100, attrs)
        S.getTimeSOAPSimple(ds)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        S.getTimeSOAPSimple(ds)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        report("getTimeSOAPSimple(stream)", f"{F}x{A}x{dc} f64 pageable numpy chunks of 100 frames -> host results", ms, host.nbytes,
               note="wall clock incl. chunk read, pinned staging, H2D, expand+normalise, timelag, D2H")
        del host, ds

    # ---- pair kernels (all-pairs matrix, configs[3]: D = 576)
    if want("pairs"):
        N = 8192 if args.quick else 25000
        for dt, code in ((torch.float64, _lib.F64), (torch.float32, _lib.F32)):
            X = torch.nn.functional.normalize(fill((N, 576), dt), dim=-1)
            for path, pname in ((_lib.PAIRS_SIMT, "simt"), (_lib.PAIRS_TENSOR_NATIVE, "tensor-native"), (_lib.PAIRS_TENSOR, "tensor-int8")):
                out = torch.empty((N, N), dtype=dt, device="cuda")
                try:
                    fn = lambda: pairs_device(X, X, _lib.KIND_NORMALIZED, 1, path=path, out=out)
                    fn()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(3):
                        fn()
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / 3
                    report(f"pairs({pname})", f"{N}x{N}x576 {dt}", ms, flops=2.0 * N * N * 576, pairs_per_s=N * N / ms * 1e3)
                except Exception as exc:  # tensor path may be unavailable for a dtype
                    print(json.dumps({"kernel": f"pairs({pname})", "config": str(dt), "error": str(exc)[:200]}), flush=True)
                del out
    return 0


if __name__ == "__main__":
    sys.exit(main())
