#!/usr/bin/env python
"""bench.py -- the SOAP post-processing hot path on B200: timeSOAP delay sweep straight from the
compressed power spectrum (expand + normalise + distance fused), in SOAP vector-pairs per second.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # the reference algorithm on host cores

One "step" = one pass of the hot path over the rank's resident trajectory tile
``X[F=1000][A][Dc=1224]`` fp64 (2 species, nMax = lMax = 8; BASELINE.json configs[1]/[2]):
``timeSOAP`` for the delay sweep (1, 5, 10, 50) with the default distance, plus the
``numpy.diff`` derivative of every window -- i.e. what four reference calls
``timeSOAP(normalizeArray(fillSOAPVectorFromdscribe(X, ...)), window=w)`` return.
Atoms are sharded over ranks with NO data-path collective (weak scaling: 12 500 atoms per GPU;
8 GPUs hold configs[2]'s 100 000 atoms).  The JSON line follows the driver's contract.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WINDOWS = (1, 5, 10, 50)
LMAX = NMAX = 8
SPECIES = ["H", "O"]
FRAMES = 1000
ATOMS_PER_GPU = 12500
METRIC = "SOAP vector-pairs/sec (timeSOAP delay sweep 1,5,10,50 from compressed dscribe SOAP, expand+normalise fused)"


def species_slices():
    upper = (LMAX + 1) * NMAX * (NMAX + 1) // 2
    full = (LMAX + 1) * NMAX * NMAX
    return {"HH": slice(0, upper), "HO": slice(upper, upper + full), "OO": slice(upper + full, 2 * upper + full)}, 2 * upper + full


def pairs_per_atom(frames):
    return sum(frames - w for w in WINDOWS)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port) on a bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(raw, slices):
    """fill -> normalise -> timeSOAP(window=w) for the sweep, exactly as the reference is used
    (Python loop over frames x atoms calling scipy's cosine): returns pairs processed."""
    from oracle import soap_oracle as O

    X = O.normalizeArray(O.fillSOAPVectorFromdscribe(raw, LMAX, NMAX, SPECIES, slices))
    pairs = 0
    for w in WINDOWS:
        t, dt = O.timeSOAP(X, window=w)
        pairs += t.size
    return pairs


def cpu_sample(frames, atoms, seed=12345):
    slices, dc = species_slices()
    rng = np.random.default_rng(seed)
    return rng.random((frames, atoms, dc)), slices


def cpu_info():
    info = {"cores": 1, "host_cpus": os.cpu_count()}
    try:
        from threadpoolctl import threadpool_info

        info["blas_threads"] = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        pass
    return info


def _ref_worker(job):
    """one host process = one shard of atoms (the reference loops are single-threaded Python; every
    atom's time series is independent, so this is the same sharding the GPU arm uses across GPUs)"""
    frames, atoms, seed, steps = job
    raw, slices = cpu_sample(frames, atoms, seed=seed)
    pairs = 0
    for _ in range(steps):
        pairs += cpu_reference_step(raw, slices)
    return pairs


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp

    frames, atoms = args.ref_frames, args.ref_atoms
    procs = args.ref_procs if args.ref_procs > 0 else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = "1"  # one BLAS thread per worker: the workers already fill the cores
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    ctx = mp.get_context("fork")  # no CUDA / torch state exists in this process
    with ctx.Pool(procs) as pool:
        # warm-up on a slice: worker start-up, imports, caches
        for _ in range(max(args.warmup, 1)):
            pool.map(_ref_worker, [(80, 2, 1000 + i, 1) for i in range(procs)])
        t0 = time.perf_counter()
        done = pool.map(_ref_worker, [(frames, atoms, 12345 + i, args.steps) for i in range(procs)])
        dt = time.perf_counter() - t0
    pairs = int(sum(done))
    value = pairs / dt
    info = cpu_info()
    info["cores"] = procs
    sample = (f"{procs} host processes x {args.steps} steps x ({frames} frames x {atoms} atoms x Dc 1224 fp64, "
              f"{pairs // (args.steps * procs)} pairs): full pipeline fill+normalise+4x timeSOAP per step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "timeSOAP delay sweep (1,5,10,50), 2 species nMax=lMax=8 (Dc 1224 -> Df 1728), reference "
                               "algorithm (oracle port of SOAPify: numpy gather, numpy normalise, Python loop + scipy cosine) "
                               "on all host cores, atoms sharded over processes; bounded sample of the GPU arm's workload",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "kind": "port", "sample": sample, **info},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def parse_topo_cpu_affinity(text: str, local: int):
    """CPU ids next to ``GPU<local>`` in the output of ``nvidia-smi topo -m`` (the "CPU Affinity" column), or
    None when the table has no such row / column."""
    header = None
    for line in text.splitlines():
        cells = [c.strip() for c in line.replace("\x1b[4m", "").replace("\x1b[0m", "").split("\t") if c.strip()]
        if not cells:
            continue
        if "CPU Affinity" in cells and header is None:
            header = cells
            continue
        if header is not None and cells[0] == f"GPU{local}":
            col = header.index("CPU Affinity") + 1  # the header has no cell above the row labels
            if col >= len(cells):
                return None
            cpus = set()
            for part in cells[col].split(","):
                lo, _, hi = part.partition("-")
                if not lo.strip().isdigit():
                    return None
                cpus.update(range(int(lo), int(hi or lo) + 1))
            return sorted(cpus)
    return None


def bind_to_gpu_numa_node(local: int):
    """Pin this rank's host threads (and so its first-touch / pinned allocations) to the cores next to its
    GPU, as listed by ``nvidia-smi topo -m``: with 8 ranks uploading at once, host buffers on the wrong
    socket halve the end-to-end rate.  Best effort: any failure leaves the affinity untouched."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        cpus = parse_topo_cpu_affinity(out, local)
        if cpus:
            cpus = set(cpus) & os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                return sorted(cpus)
    except Exception:
        pass
    return None


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    import cpctools_b200 as S
    from cpctools_b200 import _lib
    from cpctools_b200.analysis import diff_transposed_device, timelag_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (cpctools_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    host_cpus = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    slices, dc = species_slices()
    plan = S.getExpansionPlan(LMAX, NMAX, SPECIES, slices)
    mult = plan.multiplicity_device(torch.float64)

    # resident tile: F x A_local x Dc fp64 (12 500 atoms = 122.4 GB; shrink if the device is smaller)
    frames, atoms = args.frames, args.atoms
    free, _total = torch.cuda.mem_get_info()
    per_atom = frames * dc * 8 + 6 * 5 * frames * 8 + 8 * frames * 8 * 2
    atoms = int(min(atoms, (free - (6 << 30)) // per_atom))
    X = torch.empty((frames, atoms, dc), dtype=torch.float64, device="cuda")
    _lib.check(lib.soap_fill_uniform(X.data_ptr(), X.numel(), 12345 + rank, _lib.F64, None), "fill")
    torch.cuda.synchronize()

    def step():
        ts = timelag_device(X, WINDOWS, _lib.KIND_SIMPLE, 1, mult=mult)
        dts = [diff_transposed_device(t) for t in ts]
        return ts, dts

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.profile_enable(True)
    launches0 = _lib.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for k in range(args.steps):
        out = step()
        ev[k + 1].record()
    barrier()
    launches = _lib.launch_count() - launches0
    kernel_ms = _lib.profile_read(_lib.PROF_TIMELAG)
    _lib.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = ev[0].elapsed_time(ev[-1])
    if world > 1:
        tmax = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tmax.item())
        tot = torch.tensor([atoms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        total_atoms = int(tot.item())
    else:
        total_atoms = atoms
    pairs_step = pairs_per_atom(frames) * total_atoms
    value = pairs_step * args.steps / (elapsed_ms * 1e-3)

    # roofline of the dominant kernel (timelag_kernel): algorithmic bytes = the compressed tile read
    # once + the partial sums it writes; duration from CUDA events around that kernel alone
    peak, peak_src = peaks()
    scratch_bytes = lib.soap_timelag_scratch_bytes(frames, atoms, dc, _lib.F64, (ctypes_int_array(WINDOWS)), len(WINDOWS))
    algo_bytes = frames * atoms * dc * 8 + sum((frames - w) * atoms * 8 for w in WINDOWS)
    k_ms = float(np.mean(kernel_ms)) if kernel_ms else float("nan")
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_timelag_traffic.json")
    if os.path.exists(tpath):  # dram__bytes_read+write of one `ncu --set full` capture, scaled by the atom count
        tj = json.load(open(tpath))
        traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * atoms / tj["atoms"] * frames / 1000
        traffic_src = f"{tj['capture']} ({tj['config']}), scaled linearly to this tile"
    roofline = {
        "bound": "hbm", "kernel": "timelag_kernel<double,4>", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel_ms": k_ms,
        "kernel_share_of_step": k_ms * len(kernel_ms) / (ev[0].elapsed_time(ev[-1])) if kernel_ms else None,
        "algorithmic_bytes_per_launch": algo_bytes, "scratch_bytes_per_launch": int(scratch_bytes),
    }

    # end-to-end through the public API with HOST buffers (pinned), H2D and D2H inside the timed region
    e2e_atoms = min(args.e2e_atoms, atoms)
    host = torch.empty((frames, e2e_atoms, dc), dtype=torch.float64, pin_memory=True)
    host.copy_(X[:, :e2e_atoms, :])
    torch.cuda.synchronize()

    def e2e_step():
        res = S.timeSOAPfromCompressed(host, LMAX, NMAX, SPECIES, slices, windows=WINDOWS)
        return [(t.cpu(), dt.cpu()) for t, dt in res]

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        res = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tmax = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())
    e2e_pairs = pairs_per_atom(frames) * e2e_atoms * world
    d2h = sum(t.numel() * 8 + dt.numel() * 8 for t, dt in res)
    e2e = {"value": e2e_pairs * e2e_steps / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": int(host.numel() * 8),
           "d2h_bytes_per_step": int(d2h), "atoms_per_gpu": e2e_atoms, "steps": e2e_steps,
           "api": "cpctools_b200.timeSOAPfromCompressed(pinned host tensor) -> host results",
           "host_cpus_bound": None if host_cpus is None else len(host_cpus)}
    del host

    # result collective (reported beside the step, not inside it: results stay sharded by design)
    gather = None
    if world > 1:
        from cpctools_b200.sharding import gather_columns

        ts, _ = out
        full = [gather_columns(t, total_atoms) for t in ts]  # warm-up (NCCL communicator, buffers)
        del full
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        full = [gather_columns(t, total_atoms) for t in ts]
        g1.record()
        torch.cuda.synchronize()
        gms = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        nbytes = sum(int(f.numel()) * 8 for f in full)
        gather = {"collective": "NCCL all_gather of the t[F-w, A/G] slabs (every rank ends with t[F-w, A])",
                  "bytes_per_rank_out": nbytes, "ms": float(gms.item()), "GBps_per_rank": nbytes / float(gms.item()) / 1e6}
        del full

    # single-window calls on the same tile (what one reference call timeSOAP(window=w) costs)
    single = {}
    for w in (1, 50):
        for _ in range(2):
            timelag_device(X, (w,), _lib.KIND_SIMPLE, 1, mult=mult)
        torch.cuda.synchronize()
        _lib.profile_enable(True)
        for _ in range(3):
            timelag_device(X, (w,), _lib.KIND_SIMPLE, 1, mult=mult)
        torch.cuda.synchronize()
        ms = float(np.median(_lib.profile_read(_lib.PROF_TIMELAG)))
        _lib.profile_enable(False)
        nb = frames * atoms * dc * 8 + (frames - w) * atoms * 8
        single[f"window_{w}"] = {"kernel_ms": ms, "GBps": nb / ms / 1e6, "frac": nb / ms / 1e6 / peak,
                                 "pairs_per_s": (frames - w) * atoms / ms * 1e3}
    roofline["single_window_calls"] = single

    # the same sweep launch timed ALONE after the GPU has idled: inside the timed region above the SMs run
    # under sw_power_cap and this kernel (shared-memory bound) follows the SM clock
    time.sleep(3.0)
    _lib.profile_enable(True)
    timelag_device(X, WINDOWS, _lib.KIND_SIMPLE, 1, mult=mult)
    torch.cuda.synchronize()
    alone_ms = float(_lib.profile_read(_lib.PROF_TIMELAG)[0])
    _lib.profile_enable(False)
    roofline["timed_alone"] = {"kernel_ms": alone_ms, "GBps": algo_bytes / alone_ms / 1e6, "frac": algo_bytes / alone_ms / 1e6 / peak,
                               "note": "one launch after 3 s of idle (burst clocks); `frac` above is the sustained figure"}
    del X, out
    torch.cuda.empty_cache()

    allpairs = None
    if not args.no_allpairs:
        allpairs = run_allpairs(args, torch, dist, _lib, rank, world)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        raw, sl = cpu_sample(args.ref_frames, args.cpu_atoms)  # ~10 s of single-core work
        t0 = time.perf_counter()
        npairs = cpu_reference_step(raw, sl)
        dt_cpu = time.perf_counter() - t0
        cpu_baseline = {"value": npairs / dt_cpu, "unit": "pairs/s", "kind": "port",
                        "sample": f"{args.ref_frames} frames x {args.cpu_atoms} atoms x Dc 1224 fp64 ({npairs} pairs, {dt_cpu:.1f} s): "
                                  "oracle port of the reference pipeline (numpy gather + normalise, Python loop + scipy cosine)",
                        **cpu_info()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"timeSOAP delay sweep {WINDOWS} + diff on a resident compressed tile: {frames} frames x "
                                   f"{atoms} atoms/GPU x Dc {dc} fp64 (2 species H/O, nMax=lMax=8, Df 1728); BASELINE configs[1] "
                                   "dataset shape run through configs[2]'s sweep; atoms sharded, no collective",
                       "frames": frames, "atoms_per_gpu": atoms, "atoms_total": total_atoms, "windows": list(WINDOWS),
                       "pairs_per_step": pairs_step, "l2": "inputs (>= 100 GB per GPU) far larger than the 126 MB L2",
                       "timing": "CUDA events on the launching stream, max over ranks"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu_baseline, "allpairs": allpairs, "result_gather": gather,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_allpairs(args, torch, dist, _lib, rank, world):
    """BASELINE configs[3]: all-pairs SOAP distance matrix of 200 000 normalised vectors (D = 576),
    row blocks of 25 000 rows per GPU (8 GPUs = the whole 4e10-pair matrix; fewer GPUs = that share of
    it: weak scaling).  The vector set is generated on rank 0 and broadcast over NCCL; the distance
    blocks stay on their GPU (320 GB in total), only a checksum is reduced."""
    from cpctools_b200.classify import pairs_device
    from cpctools_b200.sharding import shard_range

    n_total, dim, parts = args.allpairs_vectors, 576, 8
    out = {"vectors": n_total, "dim": dim, "rows_per_gpu": None, "metric": "SOAP vector-pairs/sec (all-pairs SOAPdistanceNormalized matrix)"}
    bf16 = bf16_sus = None
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        peaks_json = json.load(open(path))
        bf16 = float(peaks_json["bf16_tflops"])
        bf16_sus = float(peaks_json.get("bf16_tflops_sustained", 0.0)) or None
    for dt, name, nd, kpad in ((torch.float64, "f64", 6, 576), (torch.float32, "f32", 3, 576)):  # K blocks of 64: 576 needs no padding
        X = torch.empty((n_total, dim), dtype=dt, device="cuda")
        if rank == 0:
            _lib.check(_lib.load().soap_fill_uniform(X.data_ptr(), X.numel(), 777, _lib.F64 if dt == torch.float64 else _lib.F32, None), "fill")
            X = torch.nn.functional.normalize(X, dim=-1)
        if world > 1:
            dist.broadcast(X, src=0)
        rows = shard_range(n_total, rank % parts, parts)
        block = torch.empty((rows.stop - rows.start, n_total), dtype=dt, device="cuda")
        fn = lambda: pairs_device(X[rows], X, _lib.KIND_NORMALIZED, 1, path=_lib.PAIRS_TENSOR, out=block)
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        reps = 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _lib.profile_enable(True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        kms = float(np.mean(_lib.profile_read(_lib.PROF_PAIRS)))
        _lib.profile_enable(False)
        ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
        chk = block[:64, :4096].double().sum().reshape(1)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(chk)
        pairs = (rows.stop - rows.start) * n_total * world
        passes = nd * (nd + 1) // 2
        int8_tops = passes * 2.0 * (rows.stop - rows.start) * n_total * kpad / (kms * 1e-3) / 1e12
        out[name] = {
            "value": pairs / (float(ms.item()) * 1e-3), "unit": "pairs/s", "ms_per_step": float(ms.item()),
            "equiv_tflops": 2.0 * pairs * dim / (float(ms.item()) * 1e-3) / 1e12,
            "kernel": f"pairs_i8_kernel<ND={nd}> (tcgen05.mma.kind::i8, {passes} exact digit passes)",
            "roofline": {"bound": "tensor", "achieved": int8_tops, "unit": "TOP/s (int8, executed)",
                         "peak": None if bf16 is None else 2.0 * bf16,
                         "frac": None if bf16 is None else int8_tops / (2.0 * bf16),
                         "peak_source": "2 x measured cuBLAS bf16 burst (MEASURED_PEAKS.json); int8 dense = 2 x bf16 on sm_100",
                         # the launches are timed back to back inside a long run: the sustained bf16 figure is the
                         # like-for-like denominator, the burst one above the conservative one
                         "peak_sustained": None if bf16_sus is None else 2.0 * bf16_sus,
                         "frac_sustained": None if bf16_sus is None else int8_tops / (2.0 * bf16_sus),
                         "kernel_ms": kms},
            "checksum_sample": float(chk.item()),
        }
        out["rows_per_gpu"] = rows.stop - rows.start
        del block
        torch.cuda.empty_cache()
        # the public all-pairs call getDistanceBetween(S, S, f) on one GPU: data is spectra, so only the
        # upper-triangular tiles are computed and mirrored (n_sym^2 ordered pairs delivered)
        n_sym = min(args.allpairs_symmetric, n_total)
        if n_sym >= 1024:
            Xs = X[:n_sym].contiguous()
            full = torch.empty((n_sym, n_sym), dtype=dt, device="cuda")
            sym = lambda: pairs_device(Xs, Xs, _lib.KIND_NORMALIZED, 1, path=_lib.PAIRS_TENSOR, out=full)
            sym()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                sym()
            e1.record()
            torch.cuda.synchronize()
            sms = e0.elapsed_time(e1) / reps
            out[name]["symmetric"] = {"vectors": n_sym, "ms": sms, "pairs_per_s": n_sym * n_sym / (sms * 1e-3),
                                      "equiv_tflops": 2.0 * n_sym * n_sym * dim / (sms * 1e-3) / 1e12,
                                      "symmetric_max_asymmetry": float((full[:2048, :2048] - full[:2048, :2048].T).abs().max().item())}
            del full, Xs
        del X
        torch.cuda.empty_cache()
    return out


def ctypes_int_array(vals):
    import ctypes

    return (ctypes.c_int * len(vals))(*vals)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--atoms", type=int, default=ATOMS_PER_GPU, help="atoms per GPU of the resident tile")
    ap.add_argument("--e2e-atoms", type=int, default=1250, help="atoms per GPU of the host-resident end-to-end tile")
    ap.add_argument("--ref-frames", type=int, default=1000)
    ap.add_argument("--ref-atoms", type=int, default=24, help="atoms of the bounded CPU sample per step and host process")
    ap.add_argument("--ref-procs", type=int, default=0, help="host processes of the reference arm (0 = all cores)")
    ap.add_argument("--cpu-atoms", type=int, default=150, help="atoms of the cpu_baseline sample inside the GPU arm (~10 s)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-allpairs", action="store_true")
    ap.add_argument("--allpairs-vectors", type=int, default=200000)
    ap.add_argument("--allpairs-symmetric", type=int, default=50000, help="vectors of the single-GPU symmetric all-pairs call")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
