"""Where does the int8 all-pairs kernel spend its time?  (SOAP_I8_DEBUG: 1 = no epilogue, 2 = one MMA per K block)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpctools_b200 import _lib
from cpctools_b200.classify import pairs_device

N = int(os.environ.get("PD_N", 25000))
for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
    X = torch.nn.functional.normalize(torch.rand(N, 576, dtype=dt, device="cuda"), dim=-1)
    Y = X.clone()  # distinct buffer: the non-symmetric schedule
    out = torch.empty((N, N), dtype=dt, device="cuda")
    for dbg in os.environ.get("PD_DBG", "0,1,4,6").split(","):
        os.environ["SOAP_I8_DEBUG"] = dbg
        for _ in range(2):
            pairs_device(X, Y, _lib.KIND_NORMALIZED, 1, path=_lib.PAIRS_TENSOR, out=out)
        _lib.profile_enable(True)
        for _ in range(5):
            pairs_device(X, Y, _lib.KIND_NORMALIZED, 1, path=_lib.PAIRS_TENSOR, out=out)
        torch.cuda.synchronize()
        ms = _lib.profile_read(_lib.PROF_PAIRS)
        _lib.profile_enable(False)
        print(f"{name} debug={dbg}: {sum(ms) / len(ms):.3f} ms", flush=True)
    del X, Y, out
