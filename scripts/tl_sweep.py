"""Tuning runs of the timelag kernel (benchmarks only): every argument is one configuration given as
comma-separated environment knobs, e.g.

    python scripts/tl_sweep.py SOAP_TL_WARPS=8 SOAP_TL_WARPS=6,SOAP_TL_SPLIT=8 SOAP_TL_DIAG=2

Shape and windows come from TL_F, TL_A, TL_W (default 1000 x 4096 x Dc 1224 fp64, windows 1,5,10,50).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import cpctools_b200 as S  # noqa: E402
from cpctools_b200 import _lib  # noqa: E402
from cpctools_b200.analysis import timelag_device  # noqa: E402

F, A = int(os.environ.get("TL_F", 1000)), int(os.environ.get("TL_A", 4096))
wins = tuple(int(w) for w in os.environ.get("TL_W", "1,5,10,50").split(","))
dtype = torch.float32 if os.environ.get("TL_DTYPE", "f64") == "f32" else torch.float64
es = 4 if dtype == torch.float32 else 8
lib = _lib.load()
raw = torch.empty((F, A, 1224), dtype=dtype, device="cuda")
_lib.check(lib.soap_fill_uniform(raw.data_ptr(), raw.numel(), 7, _lib.F64 if es == 8 else _lib.F32, None))
plan = S.getExpansionPlan(8, 8, ["H", "O"], {"HH": slice(0, 324), "HO": slice(324, 900), "OO": slice(900, 1224)})
mult = plan.multiplicity_device(dtype)
nb = raw.numel() * es + sum((F - w) * A * 8 for w in wins)
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


KNOBS = ("SOAP_TL_CHAIN", "SOAP_TL_CHAIN_ALIGN", "SOAP_TL_CHAIN_HALO", "SOAP_TL_CHAIN_BAL", "SOAP_TL_CHAIN_FAST", "SOAP_TL_CHAIN_RED", "SOAP_TL_CHAIN_OV", "SOAP_TL_CHAIN_U", "SOAP_TL_CHAIN_SR", "SOAP_TL_PAIR", "SOAP_TL_PAIR_UB", "SOAP_TL_WARPS", "SOAP_TL_SPLIT", "SOAP_TL_AHEAD", "SOAP_TL_DIAG", "SOAP_TL_GRID", "SOAP_TL_RB", "SOAP_RB_LINES", "SOAP_RB_STAGES", "SOAP_RB_DIAG")
for cfg in sys.argv[1:] or [""]:
    for k in KNOBS:
        os.environ.pop(k, None)
    for kv in filter(None, cfg.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
    try:
        ms = run()
        print(f"{cfg or 'default':40s} {ms:8.3f} ms  frac={nb / ms / 1e6 / peak:.3f}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"{cfg:40s} FAIL {str(e)[:90]}", flush=True)
