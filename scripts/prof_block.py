"""tiny driver for ncu: the BENCH-SHAPED row block of the all-pairs matrix (25 000 rows x 200 000 vectors, D 576):
python scripts/prof_block.py f32|f64 [rows] [vectors]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpctools_b200 import _lib
from cpctools_b200.classify import pairs_device
dt = torch.float32 if sys.argv[1] == "f32" else torch.float64
M = int(sys.argv[2]) if len(sys.argv) > 2 else 25000
N = int(sys.argv[3]) if len(sys.argv) > 3 else 200000
torch.manual_seed(11)
X = torch.nn.functional.normalize(torch.rand(N, 576, dtype=dt, device="cuda"), dim=-1)
out = torch.empty((M, N), dtype=dt, device="cuda")
for _ in range(2):
    pairs_device(X[:M], X, _lib.KIND_NORMALIZED, 1, path=_lib.PAIRS_TENSOR, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    pairs_device(X[:M], X, _lib.KIND_NORMALIZED, 1, path=_lib.PAIRS_TENSOR, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"ok {float(out[1, 2]):.6f}  {ms:.3f} ms  {M * N / ms / 1e6:.1f} Gpairs/s  tile={os.environ.get('SOAP_I8_TILE', 'default')}")
if os.environ.get("PB_ROWMIN"):  # the fused row minima alone (nothing of the matrix written), self distance excluded
    del out
    for _ in range(2):
        amin, arg = pairs_device(X[:M], X, _lib.KIND_NORMALIZED, 1, path=_lib.PAIRS_TENSOR, want_argmin=True, exclude_offset=0)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        amin, arg = pairs_device(X[:M], X, _lib.KIND_NORMALIZED, 1, path=_lib.PAIRS_TENSOR, want_argmin=True, exclude_offset=0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"rowmin sum(argmin)={int(arg.long().sum())} sum(min)={float(amin.sum()):.9f}  {ms:.3f} ms  {M * N / ms / 1e6:.1f} Gpairs/s")
