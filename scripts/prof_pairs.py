"""tiny driver for ncu: one all-pairs call of the requested dtype (python scripts/prof_pairs.py f32|f64 [N])"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpctools_b200 import _lib
from cpctools_b200.classify import pairs_device
dt = torch.float32 if sys.argv[1] == "f32" else torch.float64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 12288
X = torch.nn.functional.normalize(torch.rand(N, 576, dtype=dt, device="cuda"), dim=-1)
out = torch.empty((N, N), dtype=dt, device="cuda")
for _ in range(2):
    pairs_device(X, X, _lib.KIND_NORMALIZED, 1, path=_lib.PAIRS_TENSOR, out=out)
torch.cuda.synchronize()
print("ok", float(out[1, 2]))
