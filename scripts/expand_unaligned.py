"""Expansion of rows that are not 16-byte multiples (raw quippy rows: 1198 columns; dscribe 3 species n = l = 6: 1197):
python scripts/expand_unaligned.py   (SOAP_B200_LIB selects an alternative build for A/B timing)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpctools_b200 as S
from cpctools_b200 import _lib
from oracle.fake_dataset import dscribe_attrs

lib = _lib.load()
attrs, dc = dscribe_attrs(6, 6, ["H", "C", "O"])
sl = {k[len("species_location_"):].replace("-", ""): slice(*v) for k, v in attrs.items() if k.startswith("species_location_")}
peak = 6538.9
for dt, es in ((torch.float32, 4), (torch.float64, 8)):
    for name, cols, fn in (
        ("quippy raw 1198 -> 1512, expand", dc + 1, lambda x: S.fillSOAPVectorFromQuip(x, 6, 6, ["H", "C", "O"])),
        ("quippy raw 1198 -> 1512, expand+normalise", dc + 1, lambda x: S.fillAndNormalize(x, 6, 6, ["H", "C", "O"], quippy=True)),
        ("dscribe 1197 -> 1512, expand", dc, lambda x: S.fillSOAPVectorFromdscribe(x, 6, 6, ["H", "C", "O"], sl)),
        ("dscribe 1197 -> 1512, expand+normalise", dc, lambda x: S.fillAndNormalize(x, 6, 6, ["H", "C", "O"], sl)),
    ):
        X = torch.empty((100, 10000, cols), dtype=dt, device="cuda")
        _lib.check(lib.soap_fill_uniform(X.data_ptr(), X.numel(), 7, _lib.F64 if es == 8 else _lib.F32, None))
        for _ in range(2):
            out = fn(X)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = fn(X)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        nb = (X.numel() + out.numel()) * es
        print(f"{str(dt)[6:]} {name:45s} {ms:7.3f} ms  {nb / ms / 1e6:6.0f} GB/s = {nb / ms / 1e6 / peak:.3f} of peak  checksum {float(out.double().sum()):.6e}", flush=True)
        del X, out
