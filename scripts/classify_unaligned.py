"""soap_classify on rows that are / are not 16-byte multiples (tiled TMA kernel vs the warp-per-vector kernel):
python scripts/classify_unaligned.py"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpctools_b200 import _lib

lib = _lib.load()
peak = 6538.9
N, K = 4_000_000, 8
for dt, code, es in ((torch.float64, _lib.F64, 8), (torch.float32, _lib.F32, 4)):
    for D in (324, 323, 322, 147, 148, 1197, 1200):
        n = N if D < 1000 else N // 4
        raw = torch.empty((n, D), dtype=dt, device="cuda")
        _lib.check(lib.soap_fill_uniform(raw.data_ptr(), raw.numel(), 7, code, None))
        refs = torch.from_numpy(np.random.default_rng(0).random((K, D))).cuda()
        norm2 = (refs * refs).sum(-1).contiguous()
        amin = torch.empty(n, dtype=torch.float64, device="cuda")
        arg = torch.empty(n, dtype=torch.int32, device="cuda")
        fn = lambda: _lib.check(lib.soap_classify(raw.data_ptr(), None, refs.data_ptr(), norm2.data_ptr(), n, D, K, _lib.KIND_SIMPLE, 1, 1,
                                                  code, None, arg.data_ptr(), amin.data_ptr(), None))
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        nb = n * (D * es + 12)
        print(f"{str(dt)[6:]} D={D:5d} rows of {D * es:5d} B ({'aligned' if D * es % 16 == 0 else 'unaligned'}): {ms:7.3f} ms  {nb / ms / 1e6 / peak:.3f} of peak  sum(arg)={int(arg.long().sum())}", flush=True)
        del raw
