"""Host-side multi-rank logic on CPU (gloo, world_size 2 and 3): chain sharding, observable gather,
ring halo exchange.  No CUDA work is involved."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hmc_ft_b200.sharding import HaloExchanger, gather_chain_values, shard_chains


def test_shard_chains_partition():
    for total in (0, 1, 7, 1024, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_chains(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_chains(4, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, Lx, k, L):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # global field value encodes (f, mu, global row, y) so halos are checkable exactly
        Lg = Lx * world
        rows = torch.arange(Lg, dtype=torch.float64)
        glob = (rows[None, None, :, None] * 1000 + torch.arange(L, dtype=torch.float64)[None, None, None, :]
                + torch.tensor([0.0, 0.25])[None, :, None, None] + torch.tensor([0.0, 0.5])[:, None, None, None])
        ext = torch.full((2, 2, Lx + 2 * k, L), -1.0, dtype=torch.float64)
        ext[:, :, k:k + Lx] = glob[:, :, rank * Lx:(rank + 1) * Lx]
        hx = HaloExchanger(Lx, k)
        hx.exchange(ext)
        idx = [(rank * Lx - k + i) % Lg for i in range(Lx + 2 * k)]
        assert torch.equal(ext, glob[:, :, idx]), f"rank {rank}: halo mismatch"
        # theta-only exchange (3-dim tensor) uses the same code path
        e2 = ext[0].clone(); e2[:, :k] = -7; e2[:, k + Lx:] = -7
        hx.exchange(e2)
        assert torch.equal(e2, glob[0][:, idx])
        # gather of uneven chain shards
        total = 7
        off, cnt = shard_chains(total, world, rank)
        local = torch.arange(off, off + cnt, dtype=torch.float64).repeat(2, 1) + torch.tensor([[0.0], [100.0]])
        full = gather_chain_values(local, total)
        want = torch.arange(total, dtype=torch.float64).repeat(2, 1) + torch.tensor([[0.0], [100.0]])
        assert torch.equal(full, want)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_and_gather_gloo(world):
    mp.spawn(_worker, args=(world, _free_port(), 6, 2, 4), nprocs=world, join=True)


def test_halo_exchange_world1():
    Lx, k, L = 5, 2, 3
    ext = torch.zeros(2, Lx + 2 * k, L, dtype=torch.float64)
    ext[:, k:k + Lx] = torch.arange(Lx, dtype=torch.float64)[None, :, None] + 1
    HaloExchanger(Lx, k).exchange(ext)
    assert ext[0, :, 0].tolist() == [4, 5, 1, 2, 3, 4, 5, 1, 2]


def test_slab_halo_cost_model():
    """SlabHMC._choose_halo (host logic, no GPU): halo widths are multiples of 4 (steps per launch of the temporally blocked
    kernel), never exceed the slab or the trajectory, and at the 8-GPU shape of BASELINE configs[4] (512 rows per rank, 50
    steps) the model picks a width whose extended slab is exactly 10 tile rows of 56 (the minimum of the 8-GPU sweep in
    profiles/r02_slab_8gpu_halo_sweep.json: 20-24)."""
    from hmc_ft_b200.sharding import SlabHMC

    def choose(Lx, n):
        s = SlabHMC.__new__(SlabHMC)
        s.Lx, s.n_steps = Lx, n
        return s._choose_halo()

    for Lx, n in [(512, 50), (1024, 50), (2048, 50), (4096, 50), (64, 9), (512, 10), (512, 4)]:
        k = choose(Lx, n)
        assert k % 4 == 0 and 4 <= k <= 32 and k <= Lx and k <= max(4, n)
    k = choose(512, 50)
    assert k in (20, 24) and -(-(512 + 2 * k) // 56) == 10
