"""Multi-GPU sharding of the hot path: one process per GPU (``torch.distributed``, NCCL over NVLink).

Every part of the path shards without a data-path exchange:

* expand / normalise / timeSOAP: every atom's time series is independent -> contiguous ATOM
  ranges per rank (no halo); results stay sharded or are gathered once at the end
  (``all_gather`` of ``[F-w, A/G]`` slabs, a few hundred MB at most);
* all-pairs matrix: the (small) vector set is replicated, each rank computes a ROW BLOCK of the
  output, which stays sharded (it does not fit one GPU at the benchmark size); only reductions
  (row minima / arg-minima, checksums) are gathered.

The collectives only move results.  The functions below take the per-shard computation as a
callable so that the partition / gather logic is testable on CPU with the gloo backend.
"""
from __future__ import annotations

from typing import Callable, Sequence

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def _world(group=None):
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def shard_range(n: int, rank: int, world: int) -> slice:
    """Contiguous, balanced partition of ``range(n)``: the first ``n % world`` ranks get one extra."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return slice(lo, lo + base + (1 if rank < extra else 0))


def atom_shard(n_atoms: int, group=None) -> slice:
    rank, world = _world(group)
    return shard_range(n_atoms, rank, world)


def gather_columns(local, n_total: int, group=None):
    """All-gather ``[rows, n_local]`` slabs that partition the LAST axis by :func:`shard_range` into
    the full ``[rows, n_total]`` tensor on every rank (slabs are padded to the largest shard)."""
    rank, world = _world(group)
    if world == 1:
        return local
    if (n_total % world == 0 and local.is_cuda and local.dtype == torch.float64 and local.dim() == 2
            and local.shape[-1] * world == n_total):
        # equal shards on the GPU: ONE all_gather_into_tensor into [G, rows, A/G] (no padding, no list of
        # tensors, no cat) and one interleaving pass of libsoapb200 into the [rows, A] layout
        from . import _device, _lib

        rows, width = int(local.shape[0]), int(local.shape[1])
        buf = torch.empty((world, rows, width), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(buf, local.contiguous(), group=group)
        out = torch.empty((rows, n_total), dtype=local.dtype, device=local.device)
        _lib.check(_lib.load().soap_interleave_shards(buf.data_ptr(), out.data_ptr(), world, rows, width, _device.stream_ptr()),
                   "soap_interleave_shards")
        return out
    widest = shard_range(n_total, 0, world)
    width = widest.stop - widest.start
    padded = local.new_zeros(local.shape[:-1] + (width,))
    padded[..., : local.shape[-1]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    out = []
    for r, p in enumerate(parts):
        s = shard_range(n_total, r, world)
        out.append(p[..., : s.stop - s.start])
    return torch.cat(out, dim=-1)


class PeerGather:
    """All-gather of the ranks' ``t[F-w, A/G]`` slabs by the COPY ENGINES, straight into the ``[F-w, A]`` layout.

    Every rank owns a symmetric-memory buffer (``torch.distributed._symmetric_memory``: CUDA VMM memory mapped
    into all peers over NVLink) holding the full ``[rows, A]`` arrays of all windows; ``push`` issues, per slab and
    peer, ONE strided ``cudaMemcpy2DAsync`` (``soap_copy_rows``) from the local slab into the column block of this
    rank inside the peer's buffer.  No kernel runs: the gather overlaps the next sweep -- a persistent kernel that
    owns every SM, beside which NCCL's copy kernel cannot start -- without taking SMs from it, and there is no
    staging buffer or interleaving pass.  Equal shards only (``A % world == 0``); ``buffers`` of them rotate so
    that a consumer may still read step k while step k+1 lands."""

    def __init__(self, rows: Sequence[int], width: int, group=None, buffers: int = 2):
        import torch.distributed._symmetric_memory as symm

        from . import _lib

        self.rank, self.world = _world(group)
        self.rows = [int(r) for r in rows]
        self.width = int(width)
        self.total_cols = self.width * self.world
        self.group = group
        self.lib = _lib
        name = (group if group is not None else dist.group.WORLD).group_name
        shape = (sum(self.rows), self.total_cols)
        self.bufs = [symm.empty(shape, dtype=torch.float64, device="cuda") for _ in range(buffers)]
        self.handles = [symm.rendezvous(b, name) for b in self.bufs]
        self.peers = [[h.get_buffer(r, shape, torch.float64) for r in range(self.world)] for h in self.handles]
        self.step = 0

    def push(self, slabs):
        """Queue the copies of this rank's slabs (one ``[rows_i, width]`` float64 CUDA tensor per window) on the
        current stream; returns the index of the buffer they land in (on every rank)."""
        b = self.step % len(self.bufs)
        self.step += 1
        stream = int(torch.cuda.current_stream().cuda_stream)
        lib = self.lib.load()
        pitch = self.total_cols * 8
        off = 0
        for slab, rows in zip(slabs, self.rows):
            assert tuple(slab.shape) == (rows, self.width) and slab.dtype == torch.float64 and slab.is_contiguous()
            for k in range(self.world):  # start with the neighbour: the links are loaded evenly
                r = (self.rank + 1 + k) % self.world
                dst = self.peers[b][r].data_ptr() + (off * self.total_cols + self.rank * self.width) * 8
                self.lib.check(lib.soap_copy_rows(dst, pitch, slab.data_ptr(), self.width * 8, self.width * 8, rows, 2, stream),
                               "soap_copy_rows")
            off += rows
        return b

    def finish(self):
        """Wait until the slabs of EVERY rank have landed here (local stream + a barrier over the group)."""
        torch.cuda.current_stream().synchronize()
        dist.barrier(group=self.group)

    def views(self, b: int):
        """The gathered ``[rows_i, A]`` arrays of buffer ``b`` (valid after :meth:`finish`)."""
        out, off = [], 0
        for rows in self.rows:
            out.append(self.bufs[b][off : off + rows])
            off += rows
        return out


def gather_rows(local, n_total: int, group=None):
    """Same as :func:`gather_columns` for a partition of the FIRST axis."""
    rank, world = _world(group)
    if world == 1:
        return local
    moved = local.movedim(0, -1)
    return gather_columns(moved, n_total, group).movedim(-1, 0).contiguous()


def time_soap_sharded(trajectory, windows: Sequence[int], compute: Callable, n_atoms_total: int = None,
                      gather: bool = True, group=None):
    """timeSOAP over an atom-sharded trajectory.

    ``trajectory`` is this rank's ``[F, A_local, D]`` slab (its :func:`atom_shard` of the atoms);
    ``compute(trajectory, windows)`` returns one ``[F-w, A_local]`` tensor per window (e.g.
    ``cpctools_b200.analysis.timelag_device`` bound to a distance).  With ``gather`` the slabs are
    all-gathered into ``[F-w, A]`` on every rank; the derivative is taken AFTER the gather by the
    caller (or before, per shard: ``diff`` acts along frames, so both orders agree)."""
    local = compute(trajectory, windows)
    if not gather:
        return local
    rank, world = _world(group)
    if n_atoms_total is None:
        counts = torch.tensor([trajectory.shape[1]], dtype=torch.int64, device=local[0].device)
        if world > 1:
            dist.all_reduce(counts, group=group)
        n_atoms_total = int(counts.item())
    return [gather_columns(t, n_atoms_total, group) for t in local]


def all_pairs_row_block(vectors, compute: Callable, group=None, rowmin: Callable = None):
    """Row-block sharded all-pairs matrix: rank r computes ``compute(vectors[rows_r], vectors)``
    (``[n_r, N]``, kept on the rank) and the row-wise minimum over OTHER vectors, its arg-minimum
    and a checksum, which ARE gathered.  Returns ``(block, rows, row_min[N], row_argmin[N], checksum)``.

    ``rowmin(vectors[rows_r], vectors, first_row) -> (min[n_r], argmin[n_r])`` is the fused form (the tcgen05
    epilogue keeps the minima, e.g. ``pairs_device(..., want_argmin=True, exclude_offset=first_row)``): with it
    the block is neither copied nor -- when ``compute`` is ``None`` -- ever materialised (320 GB at 200 000
    vectors); the checksum is then the sum of the row minima.  Without it the minima come from a masked copy of
    the block (the CPU / gloo route of the tests)."""
    rank, world = _world(group)
    n = int(vectors.shape[0])
    rows = shard_range(n, rank, world)
    block = compute(vectors[rows], vectors) if compute is not None else None
    if rowmin is not None:
        local_min, local_arg = rowmin(vectors[rows], vectors, rows.start)
        local_arg = local_arg.to(torch.int64)
    else:
        masked = block.clone()
        idx = torch.arange(rows.start, rows.stop, device=block.device)
        masked[torch.arange(idx.numel(), device=block.device), idx] = float("inf")  # drop the self distance
        masked = torch.nan_to_num(masked, nan=float("inf"))
        local_min, local_arg = masked.min(dim=1)
    row_min = gather_rows(local_min, n, group)
    row_arg = gather_rows(local_arg, n, group)
    source = block if block is not None else local_min
    checksum = torch.nan_to_num(source.double(), nan=0.0).sum().reshape(1)
    if world > 1:
        dist.all_reduce(checksum, group=group)
    return block, rows, row_min, row_arg, float(checksum.item())
