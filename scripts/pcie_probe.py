"""What does the host->device link of this box deliver?  (context for bench.py's e2e number)"""
import torch, time
n = 1 << 30  # 4 GiB of fp32
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
dt = t(lambda: d.copy_(h, non_blocking=True))
print(f"H2D one stream, 4 GiB pinned: {n * 4 / dt / 1e9:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def two():
    with torch.cuda.stream(s1): d[: n // 2].copy_(h[: n // 2], non_blocking=True)
    with torch.cuda.stream(s2): d[n // 2 :].copy_(h[n // 2 :], non_blocking=True)
dt = t(two)
print(f"H2D two streams: {n * 4 / dt / 1e9:.1f} GB/s")
dt = t(lambda: h.copy_(d, non_blocking=True))
print(f"D2H one stream: {n * 4 / dt / 1e9:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): d[: n // 2].copy_(h[: n // 2], non_blocking=True)
    with torch.cuda.stream(s2): h[n // 2 :].copy_(d[n // 2 :], non_blocking=True)
dt = t(both)
print(f"H2D + D2H concurrently (2 GiB each): {n * 4 / dt / 1e9:.1f} GB/s total")
