"""Secondary legs of bench.py.
bench_ft:   BASELINE.json configs[3] — FT-HMC L=64 beta=6, rnet3 weights trained at beta=6 (shipped checkpoint,
            converted to tests/golden/weights_rnet3_beta6.0.npz), 1024 chains in total sharded over the ranks
            (strong scaling), 50 leapfrog steps, dt=0.05, fp64; with its CPU baseline (oracle FT port), its e2e
            (host buffers) and a cross-GPU-count parity hash.
bench_c1:   configs[0] — single chain L=16 beta=2, 10 steps (latency-bound).
bench_slab: configs[4] — one L=4096 lattice decomposed over the ranks, with parity against the undecomposed chain."""
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


def ft_cpu_baseline(arch, L, beta, n_steps, dt, full=False):
    """The reference algorithm's FT-HMC on the box's host cores (oracle FT port: torch CPU fp64, autograd force like
    hmc_u1_ft.py:128-152), SINGLE chain as the reference class runs it.  Bounded sample: new_force / new_action calls
    timed for ~10-20 s and composed into a trajectory (n_steps forces + 2 actions, hmc_u1_ft.py:181-209); with
    full=True three whole trajectories are timed instead (minutes)."""
    import time
    from oracle import u1_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    _, nets = O.load_weights_npz(os.path.join(ROOT, "tests", "golden", f"weights_{arch}_beta{beta:.1f}.npz"))
    fo = O.FieldTransformationOracle(L, nets)
    g = torch.Generator().manual_seed(3)
    th = 0.3 * torch.randn(1, 2, L, L, generator=g, dtype=torch.float64)
    cores = torch.get_num_threads()
    if full:
        h = O.FTHMCOracle(fo, beta, n_steps, dt)
        n_traj, t0 = 3, None
        for t in range(n_traj + 1):
            if t == 1:
                t0 = time.perf_counter()
            pi = torch.randn(1, 2, L, L, generator=g, dtype=torch.float64)
            th, _, _, _ = h.metropolis_step(th, pi, torch.rand(1, generator=g, dtype=torch.float64))
        s_traj = (time.perf_counter() - t0) / n_traj
        sample = f"3 full single-chain trajectories ({n_steps} steps) after 1 warm-up, oracle FT port"
    else:
        fo.new_force(th, beta)                         # warm-up
        t0, nf = time.perf_counter(), 0
        while nf < 3 or (time.perf_counter() - t0 < 10.0 and nf < 12):
            fo.new_force(th, beta)
            nf += 1
        t_force = (time.perf_counter() - t0) / nf
        t0 = time.perf_counter()
        for _ in range(2):
            fo.new_action(th, beta)
        t_action = (time.perf_counter() - t0) / 2
        s_traj = n_steps * t_force + 2 * t_action
        sample = (f"{nf} new_force calls ({t_force:.3f} s each) + 2 new_action calls ({t_action:.3f} s each), single chain, "
                  f"composed into one trajectory = {n_steps} forces + 2 actions (hmc_u1_ft.py:181-209); oracle FT port")
    return {"value": 1.0 / s_traj, "unit": "trajectories/s", "s_per_trajectory": s_traj, "cores": cores,
            "os_cpu_count": os.cpu_count(), "kind": "port", "sample": sample}


def bench_ft(args, rank, world, dev, time_steps, arch="rnet3", L=64, beta=6.0, n_steps=50, dt=0.05, total=1024,
             steps=2, warmup=3, with_cpu=False, with_e2e=False, ft_cpu_full=False):
    import hmc_ft_b200 as hb
    from hmc_ft_b200 import _lib
    from hmc_ft_b200.sharding import gather_chain_values, shard_chains
    lib = _lib.load()
    off, B = shard_chains(total, world, rank)
    ft = hb.FieldTransformation(L, device="cuda", model_tag=arch)
    ft.load_weights_npz(os.path.join(ROOT, "tests", "golden", f"weights_{arch}_beta{beta:.1f}.npz"))
    h = hb.HMC_U1_FT(L, beta, 0, n_steps, dt, ft.field_transformation, ft.compute_jac_logdet, device="cuda",
                     if_tune_step_size=False, n_chains=B, chain_offset=off, seed=1331)
    # the start depends on the GLOBAL chain id only, so every chain's history is the same for any GPU count
    g = torch.Generator().manual_seed(99)
    theta_all = 0.3 * torch.randn(total, 2, L, L, generator=g, dtype=torch.float64)
    theta = theta_all[off:off + B].to(dev)
    del theta_all
    acc, dH = [], []

    def step():
        r = h.trajectories(theta, 1)
        acc.append(r["accept"]); dH.append(r["dH"])
    ms = time_steps(step, steps, warmup, world)
    launches = h.gpu_launches // (steps + warmup)
    graph_launches = getattr(h, "graph_launches", 0) // (steps + warmup)
    traj_per_s = total / (ms * 1e-3)
    flop_per_traj_chain = 8 * MACS[arch] * 2 * (2 * n_steps + 3) * L * L
    peak = fp64_peak_tflops(lib, dev)
    ach = B * flop_per_traj_chain / (ms * 1e-3) / 1e12
    dH_all = gather_chain_values(torch.cat(dH), total)                     # [T, total], ordered by global chain id
    acc_all = gather_chain_values(torch.cat(acc).double(), total)
    import hashlib
    out = {
        "metric": f"FT-HMC trajectories/s, L={L} beta={beta} {arch}, {total} chains over {world} GPU(s), "
                  f"{n_steps} steps dt={dt} fp64",
        "value": traj_per_s, "unit": "trajectories/s", "ms_per_step": ms, "scaling": "strong",
        "chains_per_gpu": B, "gpu_launches_per_step": launches, "host_launches_per_step": graph_launches or launches,
        "accept_rate": float(acc_all.mean().item()),
        "config": {"start": "theta = 0.3 * N(0,1) per global chain id (NOT thermalised: acceptance is above the reference's "
                            "logged 0.89; timing does not depend on it)", "timed_trajectories": steps, "warmup": warmup},
        "parity": {"dH_sha16_all_chains": hashlib.sha1(dH_all.cpu().contiguous().numpy().tobytes()).hexdigest()[:16],
                   "dH_sum": float(dH_all.sum().item()), "trajectories_hashed": steps + warmup,
                   "note": "per-global-chain dH of every warm-up + timed trajectory; identical at N = 1, 2, 4, 8"},
        "roofline": {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                     "peak_source": "measured in this run (u1_fp64_probe DFMA loop)",
                     "note": "algorithmic flops = convolution MACs only, 8 subsets x MAC x 2 x (2 n_steps + 3) "
                             "per site per trajectory (SURVEY §8d); per GPU"},
    }
    if with_e2e:
        th_host = theta.cpu().pin_memory()
        out_host = torch.empty_like(th_host).pin_memory()

        def e2e_step():
            h.step_host(th_host, out_host)
            torch.cuda.current_stream().synchronize()
        ms_e = time_steps(e2e_step, 1, 1, world)
        out["e2e"] = {"value": total / (ms_e * 1e-3), "unit": "trajectories/s", "ms_per_step": ms_e,
                      "h2d_bytes_per_step": th_host.numel() * 8, "d2h_bytes_per_step": out_host.numel() * 8 + 5 * B * 8,
                      "how": "HMC_U1_FT.step_host: pinned host theta in, trajectory, theta + (dH, accept, H, plaq, topo) out"}
    if with_cpu and rank == 0:
        out["cpu_baseline"] = ft_cpu_baseline(arch, L, beta, n_steps, dt, full=ft_cpu_full)
    return out


def bench_c1(args, rank, world, dev, time_steps):
    """BASELINE configs[0]: plain HMC L=16 beta=2.0, ONE chain, 10 leapfrog steps (dt 0.1), fp64.  One chain is one
    CTA on one of the 148 SMs: a latency figure, not a throughput one.  Three call patterns: n_traj trajectories per
    launch (resident loop), one metropolis_step() call per trajectory on a device tensor (the reference's call
    pattern, host sync per step), and the same on a host tensor (H2D + D2H inside every call)."""
    import time
    import hmc_ft_b200 as hb
    Lc, beta, n, dt = 16, 2.0, 10, 0.1
    h = hb.HMC_U1(Lc, beta, 0, n, dt, device="cuda", if_tune_step_size=False, seed=1331, chain_offset=rank)
    theta = torch.zeros(1, 2, Lc, Lc, dtype=torch.float64, device=dev)
    h.trajectories(theta, 50)
    n_loop = 2000
    ms = time_steps(lambda: h.trajectories(theta, n_loop), 3, 1, world)
    res = {"metric": "plain HMC trajectories/s, C1: L=16 beta=2.0 single chain, 10 steps dt=0.1 fp64 (per GPU; latency-bound)",
           "value": n_loop / (ms * 1e-3), "unit": "trajectories/s", "site_updates_per_s": n_loop * Lc * Lc * n / (ms * 1e-3),
           "how": f"{n_loop} trajectories per launch of the SM-resident kernel (one CTA)", "us_per_trajectory": ms * 1e3 / n_loop}
    t1 = theta[0].clone()
    for _ in range(20):
        t1, _, _ = h.metropolis_step(t1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(300):
        t1, _, _ = h.metropolis_step(t1)
    torch.cuda.synchronize()
    res["per_call_device"] = {"value": 300 / (time.perf_counter() - t0), "unit": "trajectories/s",
                              "how": "HMC_U1.metropolis_step(theta) per trajectory, device tensor, host sync per call "
                                     "(accept/H returned as Python scalars like the reference)"}
    hc = hb.HMC_U1(Lc, beta, 0, n, dt, device="cpu", if_tune_step_size=False, seed=1331)
    tc = torch.zeros(2, Lc, Lc, dtype=torch.float64)
    for _ in range(20):
        tc, _, _ = hc.metropolis_step(tc)
    t0 = time.perf_counter()
    for _ in range(300):
        tc, _, _ = hc.metropolis_step(tc)
    res["e2e"] = {"value": 300 / (time.perf_counter() - t0), "unit": "trajectories/s", "h2d_bytes_per_step": 2 * Lc * Lc * 8,
                  "d2h_bytes_per_step": 2 * Lc * Lc * 8 + 12,
                  "how": "HMC_U1(device='cpu').metropolis_step(theta_host): H2D, trajectory, D2H inside every call"}
    return res


def slab_parity(world, rank, dev, L=512, beta=6.0, n_steps=50, dt=0.06, n_traj=6):
    """The decomposed chain against the SAME chain undecomposed (HMC_U1 streaming path on this GPU), on an L=512
    lattice with the bench's step count: max |d theta| over all ranks, max |d dH|, accept / topo equality.  Runs the
    exact code path that is timed (transport, graph capture after two eager trajectories, replays)."""
    import hmc_ft_b200 as hb
    from hmc_ft_b200.sharding import SlabHMC
    slab = SlabHMC(L, beta, n_steps, dt, seed=4242, chain=3)
    g = torch.Generator().manual_seed(11)
    full0 = 0.05 * torch.randn(2, L, L, generator=g, dtype=torch.float64)       # near-cold: every step is accepted, theta evolves
    own = full0[:, rank * slab.Lx:(rank + 1) * slab.Lx].contiguous().to(dev)
    outs = [slab.trajectory(own) for _ in range(n_traj)]
    ref = hb.HMC_U1(L, beta, 0, n_steps, dt, device="cuda", if_tune_step_size=False, n_chains=1, seed=4242,
                    chain_offset=3, kernel_path=1)
    th = full0.to(dev)[None].clone()
    r = ref.trajectories(th, n_traj)
    torch.cuda.synchronize()
    slab.check_peers()
    d_theta = (own - th[0][:, rank * slab.Lx:(rank + 1) * slab.Lx]).abs().max()
    d_dH = max(abs(outs[t]["dH"].item() - r["dH"][t, 0].item()) / max(1.0, abs(r["dH"][t, 0].item())) for t in range(n_traj))
    acc_eq = all(int(outs[t]["accept"].item()) == int(r["accept"][t, 0].item()) for t in range(n_traj))
    topo_eq = all(outs[t]["topo"].item() == r["topo"][t, 0].item() for t in range(n_traj))
    v = torch.tensor([d_theta.item(), d_dH, 0.0 if acc_eq else 1.0, 0.0 if topo_eq else 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
    v = v.tolist()
    res = {"lattice": L, "trajectories": n_traj, "n_steps": n_steps, "halo": slab.k, "transport": slab.transport,
           "graph_replayed": slab._graph is not None, "max_abs_dtheta": v[0], "max_rel_ddH": v[1],
           "accept_equal": v[2] == 0.0, "topo_equal": v[3] == 0.0,
           "accepts": [int(outs[t]["accept"].item()) for t in range(n_traj)],
           "ok": bool(v[0] < 1e-9 and v[1] < 1e-7 and v[2] == 0.0 and v[3] == 0.0)}
    slab.close()
    return res


def bench_slab(args, rank, world, dev, time_steps, L=4096, beta=6.0, n_steps=50, dt=0.06, halo=None, steps=8, warmup=4):
    """BASELINE configs[4]: one L=4096 lattice, plain HMC, slab-decomposed over the ranks; halo rows pushed into the
    neighbours' memory over NVLink by our kernels every k steps, whole trajectory replayed as one CUDA graph per rank
    (hmc_ft_b200/sharding.py SlabHMC, csrc/u1_slab.cu)."""
    from hmc_ft_b200.sharding import SlabHMC
    parity = slab_parity(world, rank, dev)
    slab = SlabHMC(L, beta, n_steps, dt, halo=halo, seed=1331, chain=0)
    g = torch.Generator().manual_seed(7 + rank)
    own = ((torch.rand(2, slab.Lx, L, generator=g, dtype=torch.float64) * 2 - 1) * 0.5).to(dev)
    outs = []
    ms = time_steps(lambda: outs.append(slab.trajectory(own)), steps, warmup, world)
    slab.check_peers()
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
    ew = 64 if os.environ.get("U1_MULTISTEP_EW") == "64" else 32
    res = {"metric": f"site-updates/s, single lattice L={L} plain HMC, slab-decomposed over {world} GPU(s), "
                     f"halo {slab.k} rows exchanged every {slab.k} steps",
           "value": su, "unit": "site-updates/s", "ms_per_step": ms, "scaling": "strong",
           "rows_per_gpu": slab.Lx, "halo_rows": slab.k, "redundant_row_fraction": 2 * slab.k / slab.Lx,
           "transport": slab.transport,
           "halo_exchanges_per_trajectory": slab.halo.n_exchanges // (steps + warmup),
           "halo_bytes_sent_per_trajectory_per_gpu": slab.halo.bytes_sent // (steps + warmup),
           "gpu_launches_per_step": slab.gpu_launches // (steps + warmup),
           "hbm_frac_of_measured_peak_per_gpu": None,
           "cuda_graph_replay": slab._graph is not None, "cuda_graph_error": slab.graph_error,
           "parity": parity,
           "multistep_kernel": {"kernel": f"k_leapfrog_multi<{ew}> (4 steps per launch)", "site_updates_per_s_per_gpu": su4,
                                "algorithmic_GBps": su4 * 64 / 1e9, "ms_per_launch": ms4,
                                "real_dram_bytes_per_site_update": 32.0 * (64 * ew + 56 * (ew - 8)) / (56 * (ew - 8) * 4)}}
    slab.close()
    return res
