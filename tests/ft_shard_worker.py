"""torchrun worker (BASELINE configs[3] across ranks): FT-HMC chains sharded over WORLD_SIZE GPUs, per-chain
topological-charge histories gathered with sharding.gather_chain_values, autocorrelation Gamma(delta) computed on
the device (utils.auto_from_chi / auto_by_def) - all compared with the SAME chains run unsharded on one GPU
(Philox streams are keyed by the global chain id, so the results must be bit-identical).  Prints 'FTSHARD_OK'."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hmc_ft_b200 as hb  # noqa: E402
from hmc_ft_b200 import utils as U  # noqa: E402
from hmc_ft_b200.sharding import gather_chain_values, shard_chains  # noqa: E402


def run(total, offset, count, L, beta, n_traj, n_steps, dt, theta_all):
    ft = hb.FieldTransformation(L, device="cuda", model_tag="rnet3")
    ft.load_weights_npz(os.path.join(ROOT, "tests", "golden", "weights_rnet3_beta3.0.npz"))
    h = hb.HMC_U1_FT(L, beta, 0, n_steps, dt, ft.field_transformation, ft.compute_jac_logdet, device="cuda",
                     if_tune_step_size=False, n_chains=count, chain_offset=offset, seed=2024)
    th = theta_all[offset:offset + count].cuda().clone()
    r = h.trajectories(th, n_traj)
    return th, r


def main():
    L, beta, n_traj, n_steps, dt, total = 16, 3.0, 12, 6, 0.1, 22          # 22: uneven shards at world 4 and 8
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if world > 1 else 0
    g = torch.Generator().manual_seed(17)
    theta_all = 0.5 * torch.randn(total, 2, L, L, generator=g, dtype=torch.float64)
    off, cnt = shard_chains(total, world, rank)
    th, r = run(total, off, cnt, L, beta, n_traj, n_steps, dt, theta_all)
    topo = gather_chain_values(r["topo"], total)            # [n_traj, total] on every rank
    dH = gather_chain_values(r["dH"], total)
    acc = gather_chain_values(r["accept"].double(), total)
    assert topo.shape == (n_traj, total)
    gam_chi = U.auto_from_chi(topo, 5, beta, L * L).cpu().numpy()
    gam_def = U.auto_by_def(topo, 5).cpu().numpy()
    if rank == 0:
        th1, r1 = run(total, 0, total, L, beta, n_traj, n_steps, dt, theta_all)       # unsharded, one GPU
        assert torch.equal(r1["topo"], topo), "topological-charge histories differ from the unsharded run"
        assert torch.equal(r1["dH"], dH) and torch.equal(r1["accept"].double(), acc)
        assert torch.equal(th1[off:off + cnt], th)
        g1 = U.auto_from_chi(r1["topo"], 5, beta, L * L).cpu().numpy()
        d1 = U.auto_by_def(r1["topo"], 5).cpu().numpy()
        assert np.array_equal(g1, gam_chi) and np.array_equal(d1, gam_def)
        # and against the oracle's restatement of utils.py:224-298 on the gathered histories
        from oracle import u1_oracle as O
        tn = topo.cpu().numpy()
        for b in range(total):
            ref = O.auto_from_chi(tn[:, b], 5, beta, L * L)
            assert np.max(np.abs(gam_chi[:, b] - ref)) < 1e-12
            refd = O.auto_by_def(tn[:, b], 5)
            assert np.max(np.abs(gam_def[:, b] - refd)) < 1e-12
        print("FTSHARD_OK world", world, "chains", total, "dH checksum", float(dH.sum().item()),
              "accept", float(acc.mean().item()))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
