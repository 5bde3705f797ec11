"""CTA pairs (cta_group::2) against single CTAs on the int8 all-pairs path: the outputs must be bit-identical
(exact integer sums, same epilogue); ragged shapes included.  python scripts/pairs_cg_check.py"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CHILD = r"""
import sys, torch, hashlib
sys.path.insert(0, %r)
from cpctools_b200 import _lib
from cpctools_b200.classify import pairs_device
torch.manual_seed(3)
for dt in (torch.float32, torch.float64):
    for (M, N, D) in ((300, 256, 576), (129, 333, 64), (1000, 1777, 324), (4096, 8192, 576), (257, 4000, 100)):
        X = torch.nn.functional.normalize(torch.rand(M, D, dtype=dt, device="cuda"), dim=-1)
        Y = torch.nn.functional.normalize(torch.rand(N, D, dtype=dt, device="cuda"), dim=-1)
        out = pairs_device(X, Y, _lib.KIND_NORMALIZED, 1, path=_lib.PAIRS_TENSOR)
        torch.cuda.synchronize()
        ref = torch.sqrt(torch.abs(2 - 2 * (X.double() @ Y.double().T)))
        err = float((out.double() ** 2 - ref ** 2).abs().max())
        print(str(dt)[6:], M, N, D, hashlib.sha1(out.cpu().numpy().tobytes()).hexdigest()[:16], f"{err:.2e}", flush=True)
"""
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
res = {}
for cg in ("1", "2"):
    env = dict(os.environ, SOAP_I8_CG=cg)
    p = subprocess.run([sys.executable, "-c", CHILD % root], env=env, capture_output=True, text=True, timeout=600)
    res[cg] = p.stdout.strip().splitlines()
    if p.returncode:
        print("CG", cg, "FAILED", p.stderr[-2000:])
ok = res["1"] == res["2"] and len(res["1"]) == 10
for a, b in zip(res["1"], res["2"]):
    print(a, "|", b.split()[4], "same" if a == b else "DIFFERENT")
print("pairs == single CTAs:", "ok" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
