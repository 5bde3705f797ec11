"""The N>1 path on CPU: two gloo ranks shard atoms / row blocks, compute their shard with the CPU
oracle (the partition + gather logic is what is under test; the CUDA compute is covered by
test_gpu_parity.py) and must reproduce the single-process result exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cpctools_b200 import sharding
from oracle import soap_oracle as O
from parity import assert_distance_parity


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_atoms, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(42)
        X = O.normalizeArray(rng.random((23, n_atoms, 40)))
        mine = sharding.atom_shard(n_atoms)
        assert mine == sharding.shard_range(n_atoms, rank, world)

        def compute(traj, windows):
            return [torch.from_numpy(O.timeSOAP_batched(traj.numpy(), window=w, kind=O.KIND_SIMPLE, returnDiff=False)) for w in windows]

        full = sharding.time_soap_sharded(torch.from_numpy(X[:, mine].copy()), (1, 5), compute)
        local = sharding.time_soap_sharded(torch.from_numpy(X[:, mine].copy()), (1, 5), compute, gather=False)
        assert local[0].shape == (22, mine.stop - mine.start)

        S = torch.from_numpy(O.normalizeArray(rng.random((37, 40))))

        def pairs(a, b):
            return torch.from_numpy(O.getDistanceBetween_batched(a.numpy(), b.numpy(), O.KIND_NORMALIZED, 1))

        block, rows, row_min, row_arg, checksum = sharding.all_pairs_row_block(S, pairs)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), t1=full[0].numpy(), t5=full[1].numpy(), block=block.numpy(),
                 r0=rows.start, r1=rows.stop, row_min=row_min.numpy(), row_arg=row_arg.numpy(), checksum=checksum)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_atoms", [8, 7])
def test_two_rank_sharding_matches_single_process(tmp_path, n_atoms):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_atoms, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(42)
    X = O.normalizeArray(rng.random((23, n_atoms, 40)))
    S = O.normalizeArray(rng.random((37, 40)))
    want1 = O.timeSOAP_batched(X, window=1, returnDiff=False)
    want5 = O.timeSOAP_batched(X, window=5, returnDiff=False)
    D = O.getDistanceBetween_batched(S, S, O.KIND_NORMALIZED, 1)
    masked = D.copy()
    np.fill_diagonal(masked, np.inf)
    got = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for g in got:  # every rank holds the full gathered result
        np.testing.assert_array_equal(g["t1"], want1)
        np.testing.assert_array_equal(g["t5"], want5)
        np.testing.assert_allclose(g["row_min"], masked.min(axis=1), rtol=1e-12)  # BLAS blocks rows differently
        np.testing.assert_array_equal(g["row_arg"], masked.argmin(axis=1))
        np.testing.assert_allclose(g["checksum"], D.sum(), rtol=1e-8)  # self distances are ~1e-8 rounding noise
    blocks = np.concatenate([g["block"] for g in got], axis=0)  # row blocks tile the matrix, in rank order
    assert_distance_parity(blocks, D, "f64")  # self distances are cancellation noise: compare by the rule
    assert [int(g["r0"]) for g in got] == [0, 19] and [int(g["r1"]) for g in got] == [19, 37]


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 100000, 100003):
        for world in (1, 2, 3, 8):
            parts = [sharding.shard_range(n, r, world) for r in range(world)]
            assert parts[0].start == 0 and parts[-1].stop == n
            assert all(a.stop == b.start for a, b in zip(parts, parts[1:]))
            sizes = [p.stop - p.start for p in parts]
            assert max(sizes) - min(sizes) <= 1
