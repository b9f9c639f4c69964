"""torch.profiler view of the slab trajectory (run under torchrun for world > 1): per-kernel GPU time of a few
trajectories on rank 0 and the wall time per trajectory, to see what the multi-rank step spends outside the
leapfrog kernel (NCCL send/recv, all-reduce, copies)."""
import os, sys, time, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmc_ft_b200.sharding import SlabHMC
from torch.profiler import profile, ProfilerActivity

local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank() if world > 1 else 0
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
slab = SlabHMC(L, 6.0, 50, 0.06)
own = torch.zeros(2, slab.Lx, L, dtype=torch.float64, device="cuda")
for _ in range(4):
    slab.trajectory(own)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
n = 10
t0 = time.perf_counter()
for _ in range(n):
    slab.trajectory(own)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
if rank == 0:
    print(f"world {world} L={L} Lx={slab.Lx} k={slab.k}: host issue {1e3*(t1-t0)/n:.3f} ms/traj, total {1e3*(t2-t0)/n:.3f} ms/traj")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        slab.trajectory(own)
    torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    if ev:
        busy = sum(e.time_range.elapsed_us() for e in ev)
        span = ev[-1].time_range.end - ev[0].time_range.start
        print(f"GPU busy {busy/3e3:.3f} ms/traj over span {span/3e3:.3f} ms/traj")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
