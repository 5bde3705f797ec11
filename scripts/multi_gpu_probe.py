"""One process, all visible GPUs: (1) the box's CONCURRENT host->device ceiling with 1, 2, 4, ... GPUs copying at once
(pinned host buffers, one stream per device) -- the roofline of every end-to-end number at N > 1; (2) the public API
on a pinned host trajectory fanned out over the same device sets (``devices=``), in SOAP vector-pairs per second.

    python scripts/multi_gpu_probe.py [atoms_per_gpu]        -> one JSON line
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import cpctools_b200 as S  # noqa: E402

n_dev = torch.cuda.device_count()
sets = [n for n in (1, 2, 4, 8) if n <= n_dev]
GIB = 1 << 30
out = {"gpus_visible": n_dev, "h2d_concurrent": {}, "api_one_process": {}}

# ---- (1) concurrent H2D
host = [torch.empty(2 * GIB, dtype=torch.uint8).pin_memory() for _ in range(n_dev)]
dev = [torch.empty(2 * GIB, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n_dev)]
streams = [torch.cuda.Stream(device=d) for d in range(n_dev)]


def copy_all(n, reps=3):
    for d in range(n):
        with torch.cuda.stream(streams[d]):
            for _ in range(reps):
                dev[d].copy_(host[d], non_blocking=True)
    for d in range(n):
        streams[d].synchronize()


for n in sets:
    copy_all(n, 1)
    t0 = time.perf_counter()
    copy_all(n, 3)
    dt = time.perf_counter() - t0
    out["h2d_concurrent"][str(n)] = {"GBps_total": n * 3 * 2 * GIB / dt / 1e9, "GBps_per_gpu": 3 * 2 * GIB / dt / 1e9}
del host, dev

# ---- (2) the public API, one process
atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
F, DC, W = 1000, 1224, (1, 5, 10, 50)
slices = {"HH": slice(0, 324), "HO": slice(324, 900), "OO": slice(900, 1224)}
pairs_per_atom = sum(F - w for w in W)
for n in sets:
    x = torch.empty((F, atoms * n, DC), dtype=torch.float64, pin_memory=True)
    x.view(-1)[: F * DC].uniform_()  # values do not matter for the timing; uninitialised pages are touched once below
    x[:, 1:, :] = x[:, :1, :]
    devs = list(range(n))
    run = lambda: S.timeSOAPfromCompressed(x, 8, 8, ["H", "O"], slices, windows=W, devices=devs)
    run()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        res = run()
    dt = (time.perf_counter() - t0) / reps
    out["api_one_process"][str(n)] = {"atoms": atoms * n, "s_per_call": dt, "pairs_per_s": pairs_per_atom * atoms * n / dt,
                                      "h2d_GBps": x.numel() * 8 / dt / 1e9,
                                      "frac_of_concurrent_h2d": x.numel() * 8 / dt / 1e9 / out["h2d_concurrent"][str(n)]["GBps_total"]}
    del x, res
print(json.dumps(out))
