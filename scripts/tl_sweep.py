"""Tuning sweep of the timelag kernel's launch shape (consumer warps x batch x split); benchmarks only."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cpctools_b200 as S
from cpctools_b200 import _lib
from cpctools_b200.analysis import timelag_device

F, A = int(os.environ.get('TL_F', 1000)), int(os.environ.get('TL_A', 4096))
lib = _lib.load()
raw = torch.empty((F, A, 1224), dtype=torch.float64, device="cuda")
_lib.check(lib.soap_fill_uniform(raw.data_ptr(), raw.numel(), 7, _lib.F64, None))
plan = S.getExpansionPlan(8, 8, ["H", "O"], {"HH": slice(0, 324), "HO": slice(324, 900), "OO": slice(900, 1224)})
mult = plan.multiplicity_device(torch.float64)
wins = (1, 5, 10, 50)
nb = raw.numel() * 8 + sum((F - w) * A * 8 for w in wins)
peak = 6538.9

def run():
    for _ in range(2):
        timelag_device(raw, wins, _lib.KIND_SIMPLE, 1, mult=mult)
    _lib.profile_enable(True)
    for _ in range(4):
        timelag_device(raw, wins, _lib.KIND_SIMPLE, 1, mult=mult)
    torch.cuda.synchronize()
    ms = _lib.profile_read(_lib.PROF_TIMELAG)
    _lib.profile_enable(False)
    return sum(ms) / len(ms)

cfgs = [(4, 4), (6, 4), (8, 4), (10, 4), (12, 4)]
_cfgs = [(12, 4), (12, 2), (4, 4), (4, 6), (4, 8), (6, 4), (6, 6), (6, 8), (7, 4), (8, 2), (8, 4), (8, 6), (9, 4), (10, 2), (10, 4), (10, 6)]
for (w, b), sp, diag in itertools.product(cfgs, (0,), (0, 2)):
    os.environ["SOAP_TL_DIAG"] = str(diag)
    os.environ["SOAP_TL_WARPS"] = str(w); os.environ["SOAP_TL_BATCH"] = str(b)
    os.environ["SOAP_TL_SPLIT"] = str(sp)
    try:
        ms = run()
        print(f"clk/unit={ms * 1.965e6 / (((F + 31) // 32) * 102 * (A * 6 / 148)):.2f} diag={diag} warps={w} batch={b} split={sp}: {ms:.3f} ms  frac={nb / ms / 1e6 / peak:.3f}", flush=True)
    except Exception as e:
        print(f"warps={w} batch={b} split={sp}: FAIL {str(e)[:80]}", flush=True)
