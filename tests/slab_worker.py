"""Worker for the slab-decomposed plain HMC tests (BASELINE configs[4] path) against the undecomposed chain
(streaming path of HMC_U1 on one GPU).  Prints 'SLAB_OK' on success.

    python slab_worker.py L halo [transport] [virtual_ranks]
      under torchrun: WORLD_SIZE real ranks, one GPU each (transport peer = CUDA IPC + NVLink, or nccl)
      alone, virtual_ranks = R > 1: R ranks inside this process on ONE GPU (PeerRegions.local), each on its own
      stream - the same kernels, flags and graphs as the multi-GPU path, checkable on a 1-GPU box.
The run is 6 trajectories, so with graphs on (default) it covers 2 eager passes, the capture and 4 replays."""
import os
import sys

# Virtual ranks spin on each other's flags from separate streams of ONE GPU: every stream (main + side stream per
# rank) must own a hardware work queue, or a spinning kernel at the head of a shared queue holds back the very
# kernel it waits for.  The default is 8 queues; 8 virtual ranks use 16 streams.  (Real ranks have a GPU each.)
if len(sys.argv) > 4 and int(sys.argv[4]) > 1:
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hmc_ft_b200 as hb  # noqa: E402
from hmc_ft_b200.sharding import PeerRegions, SlabHMC  # noqa: E402


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    halo = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    transport = sys.argv[3] if len(sys.argv) > 3 else "peer"
    virt = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    n_traj, n_steps, dt, beta = 6, 9, 0.1, 4.0
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if world > 1 else 0
    g = torch.Generator().manual_seed(5)
    # two starts: near-cold (every step is ACCEPTED, dH << 0: theta evolves through all kernels for 6 trajectories) and
    # hot (every step is REJECTED, dH >> 0: the select path must leave theta untouched) - checked with the oracle
    starts = [0.05 * torch.randn(2, L, L, generator=g, dtype=torch.float64),
              (torch.rand(2, L, L, generator=g, dtype=torch.float64) * 2 - 1) * 3.0]
    n_phase = [n_traj, 3]
    if virt > 1:
        assert world == 1
        k = max(1, min(halo, L // virt, n_steps))
        regs = PeerRegions.local(virt, k, L)
        slabs = [SlabHMC(L, beta, n_steps, dt, halo=halo, seed=1331, chain=0, transport="peer", peers=regs[r]) for r in range(virt)]
        streams = [torch.cuda.Stream() for _ in range(virt)]
    else:
        slabs = [SlabHMC(L, beta, n_steps, dt, halo=halo, seed=1331, chain=0, transport=transport)]
        streams = [torch.cuda.current_stream()]
    ranks = [s.rank for s in slabs]
    Lx = slabs[0].Lx
    owns = [starts[0][:, rk * Lx:(rk + 1) * Lx].contiguous().cuda() for rk in ranks]
    ref = hb.HMC_U1(L, beta, 0, n_steps, dt, device="cuda", if_tune_step_size=False, n_chains=1, seed=1331,
                    kernel_path=1)
    th = starts[0].cuda()[None].clone()
    torch.cuda.synchronize()
    worst, n_acc, t_all = 0.0, 0, 0
    for phase, (start, nt) in enumerate(zip(starts, n_phase)):
        if phase > 0:
            for own, rk in zip(owns, ranks):
                own.copy_(start[:, rk * Lx:(rk + 1) * Lx])
            th.copy_(start.cuda()[None])
            torch.cuda.synchronize()
        outs = [[] for _ in slabs]
        for t in range(nt):
            if t_all == 2 and virt > 1:                  # all virtual ranks capture before any of them replays
                torch.cuda.synchronize()
                for r in range(virt):
                    with torch.cuda.stream(streams[r]):
                        assert slabs[r].prepare_graph(owns[r]), slabs[r].graph_error
            for r, slab in enumerate(slabs):
                with torch.cuda.stream(streams[r]):
                    outs[r].append(slab.trajectory(owns[r]))
            t_all += 1
        torch.cuda.synchronize()
        for s in slabs:
            s.check_peers()
            assert s.graph_error is None, s.graph_error
        r = ref.trajectories(th, nt)                     # the same chain undecomposed (streaming path, one GPU)
        for slab, own, out in zip(slabs, owns, outs):
            rk = slab.rank
            for t in range(nt):
                assert abs(out[t]["dH"].item() - r["dH"][t, 0].item()) < 1e-8 * max(1.0, abs(r["dH"][t, 0].item())), \
                    (phase, t, out[t]["dH"].item(), r["dH"][t, 0].item())
                assert int(out[t]["accept"].item()) == int(r["accept"][t, 0].item())
                assert abs(out[t]["plaq"].item() - r["plaq"][t, 0].item()) < 1e-12
                assert out[t]["topo"].item() == r["topo"][t, 0].item()
                n_acc += int(out[t]["accept"].item())
            want = th[0][:, rk * Lx:(rk + 1) * Lx]
            err = (own - want).abs().max().item()
            assert err < 1e-10, (phase, err)
            worst = max(worst, err)
        if phase == 0:
            assert n_acc == nt * len(slabs), "cold start: every step should be accepted"
    for slab in slabs:
        assert slab.halo.n_exchanges >= sum(n_phase) * (2 + (n_steps - 1) // slab.k)
        if slab._graph_mode:
            assert slab._graph is not None, "graph was not captured"
    assert n_acc == n_phase[0] * len(slabs), "hot start: every step should be rejected"
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print("SLAB_OK world", world, "virtual", virt, "L", L, "halo", slabs[0].k, "transport", slabs[0].transport,
              "graph", slabs[0]._graph is not None, "max|dtheta|", worst, "accepted", n_acc, "of", sum(n_phase) * len(slabs))


if __name__ == "__main__":
    main()
