"""FT-HMC leg of bench.py: BASELINE.json configs[3] — L=64 beta=6, rnet3 weights trained at beta=6
(shipped checkpoint, converted to tests/golden/weights_rnet3_beta6.0.npz), 1024 chains in total
sharded over the ranks (strong scaling), 50 leapfrog steps, dt=0.05, fp64."""
from __future__ import annotations

import os

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
MACS = {"simple": 1944, "rnet1": 4536, "rnet3": 9720}      # per site per subset


def fp64_peak_tflops(lib, dev):
    """Measured DFMA peak (TFLOP/s) with the library's probe kernel; best of 5."""
    from hmc_ft_b200 import _lib
    blocks, iters = 148 * 16, 4096
    sink = torch.empty(blocks * 256, dtype=torch.float64, device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    best = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.u1_fp64_probe(sink.data_ptr(), blocks, iters, sp))
        e1.record()
        torch.cuda.synchronize()
        best = max(best, blocks * 256 * iters * 32 * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def bench_ft(args, rank, world, dev, time_steps, arch="rnet3", L=64, beta=6.0, n_steps=50, dt=0.05, total=1024,
             steps=2, warmup=3):
    import hmc_ft_b200 as hb
    from hmc_ft_b200 import _lib
    lib = _lib.load()
    B = total // world
    ft = hb.FieldTransformation(L, device="cuda", model_tag=arch)
    ft.load_weights_npz(os.path.join(ROOT, "tests", "golden", f"weights_{arch}_beta{beta:.1f}.npz"))
    h = hb.HMC_U1_FT(L, beta, 0, n_steps, dt, ft.field_transformation, ft.compute_jac_logdet, device="cuda",
                     if_tune_step_size=False, n_chains=B, chain_offset=rank * B, seed=1331)
    g = torch.Generator().manual_seed(99 + rank)
    theta = (0.3 * torch.randn(B, 2, L, L, generator=g, dtype=torch.float64)).to(dev)
    acc = []

    def step():
        r = h.trajectories(theta, 1)
        acc.append(r["accept"])
    ms = time_steps(step, steps, warmup, world)
    launches = h.gpu_launches // (steps + warmup)
    traj_per_s = world * B / (ms * 1e-3)
    flop_per_traj_chain = 8 * MACS[arch] * 2 * (2 * n_steps + 3) * L * L
    peak = fp64_peak_tflops(lib, dev)
    ach = B * flop_per_traj_chain / (ms * 1e-3) / 1e12
    return {
        "metric": f"FT-HMC trajectories/s, L={L} beta={beta} {arch}, {total} chains over {world} GPU(s), "
                  f"{n_steps} steps dt={dt} fp64",
        "value": traj_per_s, "unit": "trajectories/s", "ms_per_step": ms, "scaling": "strong",
        "chains_per_gpu": B, "gpu_launches_per_step": launches,
        "accept_rate": float(torch.cat(acc).double().mean().item()),
        "roofline": {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                     "peak_source": "measured in this run (u1_fp64_probe DFMA loop)",
                     "note": "algorithmic flops = convolution MACs only, 8 subsets x MAC x 2 x (2 n_steps + 3) "
                             "per site per trajectory (SURVEY §8d); per GPU"},
    }


def bench_slab(args, rank, world, dev, time_steps, L=4096, beta=6.0, n_steps=50, dt=0.06, halo=None, steps=3, warmup=3):
    """BASELINE configs[4]: one L=4096 lattice, plain HMC, slab-decomposed over the ranks with a
    k-row NVLink halo exchange every k steps (hmc_ft_b200/sharding.py SlabHMC)."""
    from hmc_ft_b200.sharding import SlabHMC
    slab = SlabHMC(L, beta, n_steps, dt, halo=halo, seed=1331, chain=0)
    g = torch.Generator().manual_seed(7 + rank)
    own = ((torch.rand(2, slab.Lx, L, generator=g, dtype=torch.float64) * 2 - 1) * 0.5).to(dev)
    outs = []
    ms = time_steps(lambda: outs.append(slab.trajectory(own)), steps, warmup, world)
    su = L * L * n_steps / (ms * 1e-3)
    # the temporally blocked step kernel alone (4 leapfrog steps per launch) on this rank's extended slab
    import ctypes as C
    from hmc_ft_b200 import _lib
    lib = _lib.load()
    cur, nxt = slab._ext
    flag = C.c_int(0)
    sp = torch.cuda.current_stream().cuda_stream
    rows = cur.shape[2]

    def four_steps():
        _lib.check(lib.u1_leapfrog_steps_rect(cur[0].data_ptr(), cur[1].data_ptr(), nxt[0].data_ptr(), nxt[1].data_ptr(),
                                              1, rows, L, beta, dt, dt, 4, C.byref(flag), sp))
    ms4 = time_steps(four_steps, 10, 3, 1)
    su4 = rows * L * 4 / (ms4 * 1e-3)
    return {"metric": f"site-updates/s, single lattice L={L} plain HMC, slab-decomposed over {world} GPU(s), "
                      f"halo {slab.k} rows exchanged every {slab.k} steps",
            "value": su, "unit": "site-updates/s", "ms_per_step": ms, "scaling": "strong",
            "rows_per_gpu": slab.Lx, "halo_exchanges_per_trajectory": slab.halo.n_exchanges // (steps + warmup),
            "halo_bytes_sent_per_trajectory_per_gpu": slab.halo.bytes_sent // (steps + warmup),
            "gpu_launches_per_step": slab.gpu_launches // (steps + warmup),
            "hbm_frac_of_measured_peak_per_gpu": None,
            "cuda_graph_replay": slab._graph is not None, "cuda_graph_error": slab.graph_error,
            "multistep_kernel": {"kernel": "k_leapfrog_multi (4 steps per launch)", "site_updates_per_s_per_gpu": su4,
                                 "algorithmic_GBps": su4 * 64 / 1e9, "ms_per_launch": ms4,
                                 "real_dram_bytes_per_site_update": 32.0 * (64 * 64 + 56 * 56) / (56 * 56 * 4)}}
