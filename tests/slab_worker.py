"""torchrun worker: slab-decomposed plain HMC over WORLD_SIZE GPUs against the undecomposed chain
computed on rank 0 with the streaming path of HMC_U1.  Prints 'SLAB_OK' on success."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hmc_ft_b200 as hb  # noqa: E402
from hmc_ft_b200.sharding import SlabHMC  # noqa: E402


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    halo = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    n_traj, n_steps, dt, beta = 3, 8, 0.1, 4.0
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if world > 1 else 0
    slab = SlabHMC(L, beta, n_steps, dt, halo=halo, seed=1331, chain=0)
    g = torch.Generator().manual_seed(5)
    full0 = (torch.rand(2, L, L, generator=g, dtype=torch.float64) * 2 - 1) * 3.0
    own = full0[:, rank * slab.Lx:(rank + 1) * slab.Lx].contiguous().cuda()
    outs = [slab.trajectory(own) for _ in range(n_traj)]
    torch.cuda.synchronize()
    # reference: the same chain undecomposed (rank 0 only needs it, every rank computes it cheaply)
    ref = hb.HMC_U1(L, beta, 0, n_steps, dt, device="cuda", if_tune_step_size=False, n_chains=1, seed=1331,
                    kernel_path=1)
    th = full0.cuda()[None].clone()
    r = ref.trajectories(th, n_traj)
    for t in range(n_traj):
        assert abs(outs[t]["dH"].item() - r["dH"][t, 0].item()) < 1e-8, (t, outs[t]["dH"].item(), r["dH"][t, 0].item())
        assert int(outs[t]["accept"].item()) == int(r["accept"][t, 0].item())
        assert abs(outs[t]["plaq"].item() - r["plaq"][t, 0].item()) < 1e-12
        assert outs[t]["topo"].item() == r["topo"][t, 0].item()
    want = th[0][:, rank * slab.Lx:(rank + 1) * slab.Lx]
    err = (own - want).abs().max().item()
    assert err < 1e-10, err
    assert slab.halo.n_exchanges >= n_traj * (2 + (n_steps - 1) // slab.k)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print("SLAB_OK world", world, "L", L, "halo", slab.k, "max|dtheta|", err)


if __name__ == "__main__":
    main()
