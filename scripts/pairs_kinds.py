"""int8 all-pairs path: time per distance kind (25 000^2 x 576, distinct operands)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpctools_b200 import _lib
from cpctools_b200.classify import pairs_device
N = int(os.environ.get("PD_N", 25000))
for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
    X = torch.rand(N, 576, dtype=dt, device="cuda")
    Y = X.clone()
    out = torch.empty((N, N), dtype=dt, device="cuda")
    for kname, kind, power in (("normalized", _lib.KIND_NORMALIZED, 1), ("simple", _lib.KIND_SIMPLE, 1), ("kernel n=2", _lib.KIND_KERNEL, 2)):
        A, B = (torch.nn.functional.normalize(X, dim=-1), torch.nn.functional.normalize(Y, dim=-1)) if kind == _lib.KIND_NORMALIZED else (X, Y)
        for _ in range(2):
            pairs_device(A, B, kind, power, path=_lib.PAIRS_TENSOR, out=out)
        _lib.profile_enable(True)
        for _ in range(4):
            pairs_device(A, B, kind, power, path=_lib.PAIRS_TENSOR, out=out)
        torch.cuda.synchronize()
        ms = _lib.profile_read(_lib.PROF_PAIRS)
        _lib.profile_enable(False)
        print(f"{name} {kname:12s}: {sum(ms) / len(ms):.3f} ms", flush=True)
    del X, Y, out
