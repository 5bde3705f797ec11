"""Does torch's symmetric memory work on this box, and what does a copy-engine all-gather of the t slabs straight
into the [rows, A] layout deliver?  torchrun --nproc-per-node N scripts/symm_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm
from cpctools_b200 import _lib
lib = _lib.load()
rows, width = 3934, 12500  # the four t slabs of one rank stacked: sum (1000 - w) rows x A/G columns
A = width * world
out = symm.empty((rows, A), dtype=torch.float64, device="cuda")
hdl = symm.rendezvous(out, dist.group.WORLD.group_name)
local_slab = torch.rand((rows, width), dtype=torch.float64, device="cuda")
peers = [hdl.get_buffer(r, (rows, A), torch.float64) for r in range(world)]
def push():
    s = torch.cuda.current_stream().cuda_stream
    for k in range(world):
        r = (rank + k) % world
        dst = peers[r].data_ptr() + rank * width * 8
        rc = lib.soap_copy_rows(dst, A * 8, local_slab.data_ptr(), width * 8, width * 8, rows, 2, s)
        assert rc == 0, _lib.last_error()
def barrier():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
push(); barrier()
ok = True
for r in range(world):
    pass
t0 = time.perf_counter()
for _ in range(5):
    push()
barrier()
dt = (time.perf_counter() - t0) / 5
chk = out[:, rank * width:(rank + 1) * width].equal(local_slab)
if rank == 0:
    print(f"world {world}: copy-engine all-gather into the final layout: {dt*1e3:.2f} ms per gather, "
          f"{rows * A * 8 / dt / 1e9:.0f} GB/s received per rank, own block ok={chk}")
dist.destroy_process_group()
