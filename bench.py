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
    sweep_kernel = lib.soap_timelag_last_kernel().decode() or "timelag_kernel"
    tpath = os.path.join(ROOT, "profiles", "r03_timelag_chain_traffic.json" if sweep_kernel == "timelag_chain_kernel" else "r02_timelag_traffic.json")
    if os.path.exists(tpath):  # dram__bytes_read+write of one `ncu --set full` capture, scaled by the atom count
        tj = json.load(open(tpath))
        traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * atoms / tj["atoms"] * frames / 1000
        traffic_src = f"{tj['capture']} ({tj['config']}), scaled linearly to this tile"
    roofline = {
        "bound": "hbm", "kernel": f"{sweep_kernel}<double, 4 windows>", "achieved": achieved, "peak": peak, "unit": "GB/s",
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

    # result collective: every rank ends with the full t[F-w, A] of every window.  Timed (a) alone and (b) INSIDE
    # the step loop, where the gather of step k runs on a side stream under the sweep of step k+1
    # (`value_with_gather` is the honest whole-job rate when the results are wanted on every rank)
    gather = None
    parity = None
    if world > 1:
        from cpctools_b200.sharding import gather_columns

        ts, _ = out
        full = [gather_columns(t, total_atoms) for t in ts]  # warm-up (NCCL communicator, buffers)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        full = [gather_columns(t, total_atoms) for t in ts]
        g1.record()
        torch.cuda.synchronize()
        gms = _allreduce(dist, world, torch, g0.elapsed_time(g1), dist.ReduceOp.MAX)
        nbytes = sum(int(f.numel()) * 8 for f in full)
        gather = {"collective": "NCCL all_gather_into_tensor of the t[F-w, A/G] slabs + one interleaving pass (every rank ends with t[F-w, A])",
                  "bytes_per_rank_out": nbytes, "ms": gms, "GBps_per_rank": nbytes / gms / 1e6}
        # (b) overlapped: K steps, each step's slabs gathered on a side stream while the next step computes.
        # The sweep is a persistent kernel that owns every SM's shared memory, so a NCCL kernel cannot run beside
        # it; the gather is done by the COPY ENGINES instead (sharding.PeerGather: symmetric memory, one strided
        # cudaMemcpy2DAsync per slab and peer straight into the [F-w, A] layout).  Fallback if symmetric memory is
        # unavailable: NCCL on the side stream with a few SMs left free by the sweep (soap_reserve_sms).
        side = torch.cuda.Stream()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        peer = None
        try:
            from cpctools_b200.sharding import PeerGather

            if total_atoms % world:
                raise RuntimeError("unequal shards")
            peer = PeerGather([frames - w for w in WINDOWS], atoms)
            method = "copy engines over NVLink into symmetric memory (sharding.PeerGather), no SM used"
        except Exception as exc:  # noqa: BLE001
            method = f"NCCL on a side stream, {args.reserve_sms} SMs reserved (symmetric memory unavailable: {str(exc)[:80]})"
            _lib.check(lib.soap_reserve_sms(args.reserve_sms), "soap_reserve_sms")
        keep = None
        last = 0
        for k in range(2 + args.steps):  # two untimed rounds first (allocator pools, communicators, mappings)
            if k == 2:
                torch.cuda.current_stream().wait_stream(side)
                barrier()
                o0.record()
            ts_k, dts_k = step()
            done = torch.cuda.Event()
            done.record()
            with torch.cuda.stream(side):
                side.wait_event(done)
                for t in ts_k:
                    t.record_stream(side)
                if peer is not None:
                    last = peer.push(ts_k)
                else:
                    keep = [gather_columns(t, total_atoms) for t in ts_k]
        torch.cuda.current_stream().wait_stream(side)
        o1.record()
        barrier()
        if peer is None:
            _lib.check(lib.soap_reserve_sms(0), "soap_reserve_sms")
        oms = _allreduce(dist, world, torch, o0.elapsed_time(o1), dist.ReduceOp.MAX)
        gather["overlapped_method"] = method
        gather["overlapped_ms_per_step"] = oms / args.steps
        gather["value_with_gather"] = pairs_step * args.steps / (oms * 1e-3)
        gather["efficiency_vs_no_gather"] = (elapsed_ms / args.steps) / (oms / args.steps)
        if peer is not None:
            # alone, for the bandwidth figure
            barrier()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            last = peer.push(ts_k)
            p1.record()
            peer.finish()
            pms = _allreduce(dist, world, torch, p0.elapsed_time(p1), dist.ReduceOp.MAX)
            gather["peer_gather_alone_ms"] = pms
            gather["peer_gather_GBps_per_rank"] = nbytes / pms / 1e6
            full = [v.clone() for v in peer.views(last)]
        else:
            full = keep
        del keep

        # ---- correctness of what ran under NCCL: (1) every rank recomputes the first atoms of its RIGHT
        # neighbour's shard from the raw rows (sent over NCCL) and compares with the columns it received in the
        # gather -- the kernels are deterministic, so this is bitwise; (2) the same atoms through an explicit
        # torch fp64 expand -> normalise -> distance chain (1e-12 on d^2)
        ns = min(16, atoms)
        mine = X[:, :ns, :].contiguous()
        samples = torch.empty((world, frames, ns, dc), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(samples, mine)
        nb = (rank + 1) % world
        per = total_atoms // world
        # (16 atoms would take the kernel for few atoms: insist on the one the step ran, whose sums are bitwise
        # independent of the atom count and of the grid)
        forced = sweep_kernel == "timelag_chain_kernel" and "SOAP_TL_CHAIN" not in os.environ
        if forced:
            os.environ["SOAP_TL_CHAIN"] = "1"
        redo = timelag_device(samples[nb], WINDOWS, _lib.KIND_SIMPLE, 1, mult=mult)
        if forced:
            del os.environ["SOAP_TL_CHAIN"]
        idx = plan.index_device().long()
        xa = samples[nb][:, :4, :][..., idx]
        xa = xa / xa.norm(dim=-1, keepdim=True)
        bitwise, worst, worst_ref = True, 0.0, 0.0
        for w, t_full, t_redo in zip(WINDOWS, full, redo):
            got = t_full[:, nb * per : nb * per + ns]
            bitwise = bitwise and bool(torch.equal(got, t_redo))
            worst = max(worst, float((got ** 2 - t_redo ** 2).abs().max().item()))
            ref2 = (2.0 - 2.0 * (xa[w:] * xa[:-w]).sum(-1)).clamp_min(0)
            worst_ref = max(worst_ref, float((got[:, :4] ** 2 - ref2).abs().max().item()))
        ok = worst <= 1e-12 and worst_ref <= 1e-12
        flag = _allreduce(dist, world, torch, 1.0 if ok else 0.0, dist.ReduceOp.MIN)
        parity = {"timesoap": "ok" if flag == 1.0 else "FAIL", "neighbour_recompute_bitwise": bool(_allreduce(dist, world, torch, 1.0 if bitwise else 0.0, dist.ReduceOp.MIN) == 1.0),
                  "max_abs_d2_vs_neighbour_recompute": _allreduce(dist, world, torch, worst, dist.ReduceOp.MAX),
                  "max_abs_d2_vs_torch_f64": _allreduce(dist, world, torch, worst_ref, dist.ReduceOp.MAX),
                  "sample": f"{ns} atoms x {frames} frames x 4 windows per rank, gathered slab vs recomputation on the neighbour rank"}
        del full, samples, redo

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

    # the same sweep on float32 rows (half the bytes per element, the same shared-memory wavefronts per byte)
    X32 = torch.empty((frames, atoms, dc), dtype=torch.float32, device="cuda")
    _lib.check(lib.soap_fill_uniform(X32.data_ptr(), X32.numel(), 12345 + rank, _lib.F32, None), "fill")
    mult32 = plan.multiplicity_device(torch.float32)
    for _ in range(2):
        timelag_device(X32, WINDOWS, _lib.KIND_SIMPLE, 1, mult=mult32)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(5):
        timelag_device(X32, WINDOWS, _lib.KIND_SIMPLE, 1, mult=mult32)
    torch.cuda.synchronize()
    ms32 = float(np.mean(_lib.profile_read(_lib.PROF_TIMELAG)))
    _lib.profile_enable(False)
    nb32 = frames * atoms * dc * 4 + sum((frames - w) * atoms * 8 for w in WINDOWS)
    roofline_f32 = {"bound": "hbm", "kernel": f"{lib.soap_timelag_last_kernel().decode()}<float, 4 windows>", "achieved": nb32 / ms32 / 1e6, "peak": peak, "unit": "GB/s",
                    "frac": nb32 / ms32 / 1e6 / peak, "kernel_ms": ms32, "algorithmic_bytes_per_launch": nb32,
                    "pairs_per_s": pairs_per_atom(frames) * atoms / (ms32 * 1e-3), "traffic": None}
    del X32
    torch.cuda.empty_cache()

    expand = quippy = None
    if not args.no_extra:
        expand = run_expand(args, torch, dist, _lib, S, rank, world, peak)
        quippy = run_quippy(args, torch, dist, _lib, S, rank, world, peak)
    tensor_peaks = probe_tensor_peaks(torch) if (rank == 0 and not args.no_allpairs) else None
    if world > 1:
        box = [tensor_peaks]
        dist.broadcast_object_list(box, src=0)
        tensor_peaks = box[0]

    allpairs = None
    if not args.no_allpairs:
        allpairs = run_allpairs(args, torch, dist, _lib, rank, world, tensor_peaks)
        if parity is not None and allpairs is not None:
            parity["allpairs"] = allpairs.get("parity")

    cpu_baseline = config1 = None
    if rank == 0 and world == 1 and not args.no_cpu:
        config1 = run_config1(S)
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
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "roofline_f32": roofline_f32,
            "cpu_baseline": cpu_baseline, "config1": config1, "expand": expand, "quippy": quippy, "allpairs": allpairs, "tensor_peaks": tensor_peaks,
            "result_gather": gather, "multi_gpu_parity": parity,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_config1(S):
    """BASELINE configs[0], the reference's own CPU-runnable case, in full on both sides: 200 frames x 309 atoms x
    324 (1 species, nMax = lMax = 8) fp64 -> fillSOAPVectorFromdscribe -> normalizeArray ->
    timeSOAP(window=1, SOAPdistanceNormalized), numpy in -> numpy out through the public API on the GPU, and
    the reference algorithm (oracle port) on one host core.  Tiny (fits L2): a parity + latency line."""
    from oracle import soap_oracle as O

    rng = np.random.default_rng(12345)
    raw = rng.random((200, 309, 324))

    def gpu():
        X = S.normalizeArray(S.fillSOAPVectorFromdscribe(raw, 8, 8))
        return S.timeSOAP(X, window=1, distanceFunction=S.SOAPdistanceNormalized)

    gpu()
    t0 = time.perf_counter()
    t_gpu, dt_gpu = gpu()
    gpu_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    (t_fused, _), = S.timeSOAPfromCompressed(raw, 8, 8, windows=(1,), distanceFunction=S.SOAPdistanceNormalized)
    fused_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    Xo = O.normalizeArray(O.fillSOAPVectorFromdscribe(raw, 8, 8))
    t_cpu, dt_cpu = O.timeSOAP(Xo, window=1, distanceFunction=O.SOAPdistanceNormalized)
    cpu_s = time.perf_counter() - t0
    pairs = t_cpu.size
    err = float(np.abs(t_gpu ** 2 - t_cpu ** 2).max())
    err_f = float(np.abs(t_fused ** 2 - t_cpu ** 2).max())
    return {"workload": "timeSOAP(window=1, SOAPdistanceNormalized) after fill + normalise, 200 frames x 309 atoms x 324 -> 576 fp64 (BASELINE configs[0])",
            "pairs": int(pairs), "gpu_three_calls_s": gpu_s, "gpu_fused_call_s": fused_s, "cpu_port_s": cpu_s,
            "gpu_pairs_per_s": pairs / gpu_s, "gpu_fused_pairs_per_s": pairs / fused_s, "cpu_pairs_per_s": pairs / cpu_s,
            "max_abs_d2_err": err, "max_abs_d2_err_fused": err_f, "shapes": [list(t_gpu.shape), list(dt_gpu.shape)],
            "parity": "ok" if max(err, err_f) <= 1e-12 and t_gpu.shape == t_cpu.shape and dt_gpu.shape == dt_cpu.shape else "FAIL",
            "note": "host numpy arrays in and out: the GPU time is launch + PCIe latency, not bandwidth"}


def _allreduce(dist, world, torch, value, op):
    t = torch.tensor([value], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=op)
    return float(t.item())


def run_expand(args, torch, dist, _lib, S, rank, world, peak):
    """BASELINE configs[1]: fillSOAPVectorFromdscribe + normalizeArray fused (one kernel: reads Dc = 1224, writes
    Df = 1728 per vector) on a device-resident tile of 100 frames x 10 000 atoms -- 1/10 of the 1000 x 10 000
    dataset, whose input + output (236 GB in fp64) do not fit one GPU together.  Kernel time from CUDA events
    around expand_kernel; every rank runs its own tile (weak)."""
    slices, dc = species_slices()
    plan = S.getExpansionPlan(LMAX, NMAX, SPECIES, slices)
    frames, atoms = args.expand_frames, args.expand_atoms
    out = {"workload": f"fused expand+normalise, {frames} frames x {atoms} atoms x Dc {dc} -> Df {plan.d_full}, 2 species",
           "vectors_per_gpu": frames * atoms, "target_vec_per_s_per_gpu": 1.94e8}
    for dt, name, code in ((torch.float64, "f64", _lib.F64), (torch.float32, "f32", _lib.F32)):
        es = 8 if dt == torch.float64 else 4
        raw = torch.empty((frames, atoms, dc), dtype=dt, device="cuda")
        _lib.check(_lib.load().soap_fill_uniform(raw.data_ptr(), raw.numel(), 4242 + rank, code, None), "fill")
        for _ in range(3):
            full = S.fillAndNormalize(raw, LMAX, NMAX, SPECIES, slices)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        _lib.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            full = S.fillAndNormalize(raw, LMAX, NMAX, SPECIES, slices)
        e1.record()
        torch.cuda.synchronize()
        kms = float(np.mean(_lib.profile_read(_lib.PROF_EXPAND)))
        _lib.profile_enable(False)
        ms = _allreduce(dist, world, torch, e0.elapsed_time(e1) / reps, dist.ReduceOp.MAX if world > 1 else None)
        nvec = frames * atoms
        nbytes = nvec * (dc + plan.d_full) * es
        # size-independent properties on the full tile: gather == table lookup (bit-exact up to the norm), unit norms
        idx = plan.index_device()
        rows = torch.randint(0, nvec, (4096,), device="cuda")
        src = raw.reshape(nvec, dc)[rows].double()[:, idx.long()]
        want = src / src.norm(dim=-1, keepdim=True)
        err = float((full.reshape(nvec, -1)[rows].double() - want).abs().max().item())
        out[name] = {"kernel_ms": kms, "ms_per_call": ms, "GBps": nbytes / kms / 1e6, "frac": nbytes / kms / 1e6 / peak,
                     "vec_per_s_per_gpu": nvec / (kms * 1e-3), "vec_per_s_total": nvec * world / (ms * 1e-3),
                     "algorithmic_bytes_per_launch": nbytes, "max_abs_err_vs_torch_f64_sample": err,
                     "parity": "ok" if err <= (1e-12 if es == 8 else 1e-6) else "FAIL"}
        del raw, full
        torch.cuda.empty_cache()
    return out


def run_quippy(args, torch, dist, _lib, S, rank, world, peak):
    """BASELINE configs[4]: raw quippy rows (3 species H/C/O, nMax = lMax = 6: 1197 + 1 columns) -> drop the last
    column, permute to the dscribe order, expand to 1512, normalise, timeSOAP(window = 1) -- all folded into ONE pass
    of the time-lag kernel over the raw rows -- in fp64 and in fp32 (the fp64 values rounded), atoms sharded over
    the ranks; the two precisions are compared (BASELINE: <= 1e-6)."""
    species = ["H", "C", "O"]
    frames, atoms = args.quippy_frames, args.quippy_atoms
    plan = S.getExpansionPlan(6, 6, species, S.utils._quippyContainerSlices(6, 6, species), quippy=True)
    dq = plan.d_compressed
    raw = torch.empty((frames, atoms, dq), dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().soap_fill_uniform(raw.data_ptr(), raw.numel(), 99 + rank, _lib.F64, None), "fill")
    out = {"workload": f"raw quippy rows [{frames} x {atoms}/GPU x {dq}] -> permute + expand ({plan.d_full}) + normalise + timeSOAP(window=1), fused",
           "pairs_per_gpu": (frames - 1) * atoms}
    res = {}
    for name, x in (("f64", raw), ("f32", raw.float())):
        es = x.element_size()
        fn = lambda: S.timeSOAPfromCompressed(x, 6, 6, species, None, windows=(1,), quippy=True)
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        _lib.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            (t, dt), = fn()
        e1.record()
        torch.cuda.synchronize()
        kms = float(np.mean(_lib.profile_read(_lib.PROF_TIMELAG)))
        _lib.profile_enable(False)
        ms = _allreduce(dist, world, torch, e0.elapsed_time(e1) / reps, dist.ReduceOp.MAX if world > 1 else None)
        nbytes = frames * atoms * dq * es + (frames - 1) * atoms * 8
        res[name] = t
        out[name] = {"kernel_ms": kms, "ms_per_call": ms, "GBps": nbytes / kms / 1e6, "frac": nbytes / kms / 1e6 / peak,
                     "pairs_per_s_total": (frames - 1) * atoms * world / (ms * 1e-3),
                     "note": None if (dq * es) % 16 == 0 else "rows of 8 mod 16 bytes: fetched by TMA through 16-byte aligned windows, the foreign elements of the first / last unit masked"}
    d2 = float((res["f32"] ** 2 - res["f64"] ** 2).abs().max().item())
    big = res["f64"] >= 0.25
    rel = float(((res["f32"] - res["f64"]).abs() / res["f64"])[big].max().item()) if bool(big.any()) else 0.0
    d2 = _allreduce(dist, world, torch, d2, dist.ReduceOp.MAX if world > 1 else None)
    out["fp32_vs_fp64"] = {"max_abs_d2": d2, "max_rel_d_where_d_ge_0.25": rel, "tolerance": 1e-6, "parity": "ok" if d2 <= 1e-6 and rel <= 1e-6 else "FAIL"}
    # fp64 against torch: a few atoms through the explicit permute -> expand -> normalise -> distance chain
    idx = plan.index_device().long()
    xa = raw[:, :4, :][..., idx]
    xa = xa / xa.norm(dim=-1, keepdim=True)
    ref2 = 2.0 - 2.0 * (xa[1:] * xa[:-1]).sum(-1)
    out["f64"]["max_abs_d2_vs_torch"] = float((res["f64"][:, :4] ** 2 - ref2.clamp_min(0)).abs().max().item())
    del raw, res
    torch.cuda.empty_cache()
    return out


def probe_tensor_peaks(torch):
    """The denominators BASELINE.md section 3 left "to measure": dense INT8 (torch._int_mm -> cuBLASLt IMMA), FP64
    (DGEMM) and TF32 GEMM rates on this GPU, 8192^3, best single call (burst) and a ~1 s back-to-back loop
    (sustained), CUDA events.  Library GEMMs used ONLY as yardsticks."""
    out = {}
    n = 8192

    def rate(fn, flops):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        best = 1e30
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(5):
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(3, int(1000.0 / best))
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return {"burst": flops / best / 1e9, "sustained": flops * reps / e0.elapsed_time(e1) / 1e9}

    try:
        a = torch.randint(-100, 100, (n, n), dtype=torch.int8, device="cuda")
        b = torch.randint(-100, 100, (n, n), dtype=torch.int8, device="cuda").t().contiguous().t()
        out["int8_tops"] = rate(lambda: torch._int_mm(a, b), 2.0 * n ** 3)
        del a, b
    except Exception as e:  # noqa: BLE001
        out["int8_tops"] = {"error": str(e)[:120]}
    try:
        a = torch.randn((n, n), dtype=torch.float64, device="cuda"); b = torch.randn((n, n), dtype=torch.float64, device="cuda")
        out["fp64_tflops"] = rate(lambda: torch.matmul(a, b), 2.0 * n ** 3)
        del a, b
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn((n, n), dtype=torch.float32, device="cuda"); b = torch.randn((n, n), dtype=torch.float32, device="cuda")
        out["tf32_tflops"] = rate(lambda: torch.matmul(a, b), 2.0 * n ** 3)
        torch.backends.cuda.matmul.allow_tf32 = prev
        del a, b
    except Exception as e:  # noqa: BLE001
        out["gemm_error"] = str(e)[:120]
    torch.cuda.empty_cache()
    return out


def run_allpairs(args, torch, dist, _lib, rank, world, tensor_peaks=None):
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
    int8_peak = int8_sus = None
    if tensor_peaks and isinstance(tensor_peaks.get("int8_tops"), dict) and "burst" in tensor_peaks["int8_tops"]:
        int8_peak, int8_sus = tensor_peaks["int8_tops"]["burst"], tensor_peaks["int8_tops"]["sustained"]
    parity_ok = True
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
        # a sampled 128 x 4096 tile of THIS rank's block against torch fp64 sqrt|2 - 2 x.y| (d^2 rule of tests/parity.py)
        xr, xc = X[rows][:128].double(), X[:4096].double()
        ref2 = (2.0 - 2.0 * (xr @ xc.T)).abs()
        err2 = float((block[:128, :4096].double() ** 2 - ref2).abs().max().item())
        tol2 = 1e-12 if dt == torch.float64 else 1e-6
        err2 = _allreduce(dist, world, torch, err2, dist.ReduceOp.MAX if world > 1 else None)
        parity_ok = parity_ok and err2 <= tol2
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
            "sample_tile_max_abs_d2_err": err2, "sample_tile_tolerance": tol2,
            "roofline": {"bound": "tensor", "achieved": int8_tops, "unit": "TOP/s (int8, executed)",
                         "peak": int8_peak if int8_peak else (None if bf16 is None else 2.0 * bf16),
                         "frac": (int8_tops / int8_peak) if int8_peak else (None if bf16 is None else int8_tops / (2.0 * bf16)),
                         "peak_source": ("measured in this run: torch._int_mm (cuBLASLt IMMA) 8192^3, best single call" if int8_peak else
                                         "2 x measured cuBLAS bf16 burst (MEASURED_PEAKS.json); int8 dense = 2 x bf16 on sm_100"),
                         "peak_int8_sustained": int8_sus, "frac_int8_sustained": (int8_tops / int8_sus) if int8_sus else None,
                         "peak_bf16x2_proxy": None if bf16 is None else 2.0 * bf16,
                         # the launches are timed back to back inside a long run: the sustained bf16 figure is the
                         # like-for-like denominator, the burst one above the conservative one
                         "peak_sustained": None if bf16_sus is None else 2.0 * bf16_sus,
                         "frac_sustained": None if bf16_sus is None else int8_tops / (2.0 * bf16_sus),
                         "kernel_ms": kms},
            "checksum_sample": float(chk.item()),
        }
        out["rows_per_gpu"] = rows.stop - rows.start
        # per-row nearest OTHER vector (min / argmin over the whole row of the 200 000-column matrix) fused into the
        # tcgen05 epilogue: nothing of the block is written ("gather only ... per-row min/argmin")
        rm = lambda: pairs_device(X[rows], X, _lib.KIND_NORMALIZED, 1, want_argmin=True, exclude_offset=rows.start)
        amin, arg = rm()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            amin, arg = rm()
        e1.record()
        torch.cuda.synchronize()
        rm_ms = _allreduce(dist, world, torch, e0.elapsed_time(e1) / reps, dist.ReduceOp.MAX if world > 1 else None)
        sub = block[:256].double().clone()
        sub[torch.arange(256, device="cuda"), rows.start + torch.arange(256, device="cuda")] = float("inf")
        smin, sarg = sub.min(dim=1)
        rm_ok = bool(torch.equal(smin, amin[:256]) and torch.equal(sarg.to(torch.int32), arg[:256]))
        parity_ok = parity_ok and rm_ok
        out[name]["row_minima"] = {"ms_per_step": rm_ms, "pairs_per_s": pairs / (rm_ms * 1e-3), "matrix_written": False,
                                   "matches_matrix_rows_0_255": rm_ok,
                                   "api": "soap_pairs_rowmin(out=NULL, exclude_offset=first row of the block)"}
        del block, sub
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
    out["parity"] = "ok" if parity_ok else "FAIL"
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
    ap.add_argument("--reserve-sms", type=int, default=8, help="SMs the sweep leaves to the overlapped NCCL gather (N > 1)")
    ap.add_argument("--no-extra", action="store_true", help="skip the expand (configs[1]) and quippy (configs[4]) objects")
    ap.add_argument("--expand-frames", type=int, default=100)
    ap.add_argument("--expand-atoms", type=int, default=10000)
    ap.add_argument("--quippy-frames", type=int, default=1000)
    ap.add_argument("--quippy-atoms", type=int, default=4096, help="atoms per GPU of the quippy tile (8 x 4096 on 8 GPUs)")
    ap.add_argument("--allpairs-vectors", type=int, default=200000)
    ap.add_argument("--allpairs-symmetric", type=int, default=50000, help="vectors of the single-GPU symmetric all-pairs call")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
