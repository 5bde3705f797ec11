"""Break-down of the end-to-end step of bench.py (pinned host tensor -> host results)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpctools_b200 as S
from cpctools_b200 import _lib, _device
from cpctools_b200.analysis import timelag_device, diff_transposed_device
F, A, DC = 1000, 1250, 1224
host = torch.rand((F, A, DC), dtype=torch.float64).pin_memory()
slices = {"HH": slice(0, 324), "HO": slice(324, 900), "OO": slice(900, 1224)}
plan = S.getExpansionPlan(8, 8, ["H", "O"], slices)
mult = plan.multiplicity_device(torch.float64)
W = (1, 5, 10, 50)
def sync(): torch.cuda.synchronize()
def tm(fn, reps=3):
    fn(); sync(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    sync(); return (time.perf_counter() - t0) / reps * 1e3
print("full API call + .cpu(): %.1f ms" % tm(lambda: [(t.cpu(), dt.cpu()) for t, dt in S.timeSOAPfromCompressed(host, 8, 8, ["H", "O"], slices, windows=W)]))
print("H2D only (to new device tensor): %.1f ms" % tm(lambda: host.to("cuda", non_blocking=True)))
dev = torch.empty_like(host, device="cuda")
print("H2D only (into existing tensor): %.1f ms" % tm(lambda: dev.copy_(host, non_blocking=True)))
print("as_float_tensor: %.1f ms" % tm(lambda: _device.as_float_tensor(host)))
ts = timelag_device(dev, W, _lib.KIND_SIMPLE, 1, mult=mult)
print("kernels (timelag + 4 diff): %.1f ms" % tm(lambda: [diff_transposed_device(t) for t in timelag_device(dev, W, _lib.KIND_SIMPLE, 1, mult=mult)]))
dts = [diff_transposed_device(t) for t in ts]
print("D2H results .cpu(): %.1f ms" % tm(lambda: [(t.cpu(), d.cpu()) for t, d in zip(ts, dts)]))
arr = host.numpy().copy()  # pageable numpy: what a user gets from h5py
t0 = time.perf_counter(); r = S.timeSOAPfromCompressed(arr, 8, 8, ["H", "O"], slices, windows=W); sync(); t1 = time.perf_counter()
t2 = time.perf_counter(); r = S.timeSOAPfromCompressed(arr, 8, 8, ["H", "O"], slices, windows=W); sync(); t3 = time.perf_counter()
print("full API call, pageable numpy in -> numpy out: first %.1f ms, second %.1f ms (%.1f GB/s)" % ((t1 - t0) * 1e3, (t3 - t2) * 1e3, arr.nbytes / (t3 - t2) / 1e9))
# expansion with numpy in / numpy out (the reference's container): 2.45 GB in, 3.46 GB out
x = arr[:200]
for rep in range(2):
    t0 = time.perf_counter(); y = S.fillSOAPVectorFromdscribe(x, 8, 8, ["H", "O"], slices); t1 = time.perf_counter()
    print("fillSOAPVectorFromdscribe numpy->numpy %s -> %s: %.1f ms (%.1f GB/s of in+out)" % (x.shape, y.shape, (t1 - t0) * 1e3, (x.nbytes + y.nbytes) / (t1 - t0) / 1e9))
_device._HostStager.MIN_BYTES = 1 << 50  # the plain paths: whole-array pin for the upload, .cpu() for the download
for rep in range(2):
    t0 = time.perf_counter(); y = S.fillSOAPVectorFromdscribe(x, 8, 8, ["H", "O"], slices); t1 = time.perf_counter()
    print("  same without staging: %.1f ms (%.1f GB/s)" % ((t1 - t0) * 1e3, (x.nbytes + y.nbytes) / (t1 - t0) / 1e9))
