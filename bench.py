#!/usr/bin/env python
"""Benchmark of the 2D U(1) HMC / FT-HMC hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one HMC trajectory (n_steps=50 leapfrog steps + momentum refresh + Metropolis +
observables) of every chain of BASELINE.json configs[1]: plain HMC, L=64, beta=6, B=4096 chains per
GPU, fp64.  metric = site-updates/s (1 site-update = one site, 2 links, through one leapfrog step).
Chains are sharded over ranks with no data-path collective (weak scaling: 4096 chains per GPU).
The same JSON line carries, as extra objects, the other BASELINE configurations: "c1" (configs[0], single
chain L=16), "ft" (configs[3], FT-HMC L=64, strong scaling over the ranks, with its own cpu_baseline and e2e),
"ft_other" (configs[2] and the smaller shipped CNNs) and "slab" (configs[4], one L=4096 lattice decomposed
over the ranks, with a parity block against the undecomposed chain).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L, BETA, N_STEPS, DT, B_PER_GPU = 64, 6.0, 50, 0.06, 4096
C1_L, C1_BETA, C1_N_STEPS, C1_DT = 16, 2.0, 10, 0.1
BYTES_PER_SITE_UPDATE = 64.0      # read+write theta0,theta1,pi0,pi1 in fp64
UNIT = "site-updates/s"
METRIC = "site-updates/sec (force+leapfrog), plain HMC L=64 beta=6 B=4096 chains/GPU fp64"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_constants():
    """Per-launch counters of the dominant kernels taken from the committed ncu captures of THIS round
    (profiles/r02_kernel_constants.json, written by tools/ncu_constants.py from the ncu CSV): FP64-pipe
    instructions per site-update and real DRAM bytes per launch.  Static properties of the SASS / data layout."""
    p = os.path.join(ROOT, "profiles", "r02_kernel_constants.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def time_steps(fn, steps, warmup, world):
    for _ in range(warmup):
        fn()
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier(world)
    return max_over_ranks(e0.elapsed_time(e1), world) / steps      # ms per step, max over ranks


def sha16(t: torch.Tensor) -> str:
    """Bit-exact fingerprint of a tensor (parity across GPU counts is 'same hash')."""
    return hashlib.sha1(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:16]


# ----------------------------------------------------------------------------- CPU baselines
def _plain_oracle_rate(Lc, beta, n_steps, dt, n_chains, n_traj, warm=1):
    from oracle import u1_oracle as O
    o = O.PlainHMCOracle(Lc, beta, n_steps, dt)
    g = torch.Generator().manual_seed(1331)
    th = torch.zeros(n_chains, 2, Lc, Lc, dtype=torch.float64)

    def one():
        nonlocal th
        pi = torch.randn(n_chains, 2, Lc, Lc, generator=g, dtype=torch.float64)
        u = torch.rand(n_chains, generator=g, dtype=torch.float64)
        th, _, _, _ = o.metropolis_step(th, pi, u)

    for _ in range(warm):
        one()
    t0 = time.perf_counter()
    for _ in range(n_traj):
        one()
    return (time.perf_counter() - t0) / n_traj


def cpu_baseline_plain(n_chains=128, n_traj=4):
    """Oracle port (oracle/u1_oracle.py, torch CPU fp64, autograd force like hmc_u1.py:65-69) on the box's host
    cores, bounded samples of the C2 workload: (a) batched over 128 chains so every host thread has work (the
    headline cpu_baseline), (b) the stock single-chain form the reference class has (hmc_u1.py:82-97, B=1),
    which BASELINE.md §3 asks to be shown 'as is'."""
    torch.set_num_threads(os.cpu_count() or 1)
    dt_b = _plain_oracle_rate(L, BETA, N_STEPS, DT, n_chains, n_traj)
    dt_1 = _plain_oracle_rate(L, BETA, N_STEPS, DT, 1, 20, warm=3)
    cores = torch.get_num_threads()
    return {"value": n_chains * L * L * N_STEPS / dt_b, "unit": UNIT, "cores": cores, "os_cpu_count": os.cpu_count(),
            "kind": "port",
            "sample": f"{n_traj} trajectories x {n_chains} chains of the C2 workload (L=64 beta=6 50 steps), "
                      f"oracle torch-CPU fp64 port batched over chains, {dt_b:.3f} s/trajectory",
            "single_chain_stock": {"value": L * L * N_STEPS / dt_1, "unit": UNIT, "traj_per_s": 1.0 / dt_1,
                                   "sample": "20 trajectories, ONE chain (the reference class cannot batch; "
                                             "hmc_u1.py:82-97 semantics), same port, after 3 warm-up trajectories"}}


def cpu_baseline_c1():
    torch.set_num_threads(os.cpu_count() or 1)
    dt_1 = _plain_oracle_rate(C1_L, C1_BETA, C1_N_STEPS, C1_DT, 1, 200, warm=20)
    return {"value": 1.0 / dt_1, "unit": "trajectories/s", "site_updates_per_s": C1_L * C1_L * C1_N_STEPS / dt_1,
            "cores": torch.get_num_threads(), "kind": "port",
            "sample": "200 trajectories of C1 (L=16 beta=2 B=1, 10 steps), oracle torch-CPU fp64 port, single chain"}


def run_reference(args):
    """--impl reference: the reference algorithm's CPU path (oracle port; the reference is pure
    Python/torch and cannot be compiled) on this box's host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_chains = 128
    from oracle import u1_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    o = O.PlainHMCOracle(L, BETA, N_STEPS, DT)
    g = torch.Generator().manual_seed(1331)
    th = torch.zeros(n_chains, 2, L, L, dtype=torch.float64)

    def one():
        nonlocal th
        pi = torch.randn(n_chains, 2, L, L, generator=g, dtype=torch.float64)
        u = torch.rand(n_chains, generator=g, dtype=torch.float64)
        th, _, _, _ = o.metropolis_step(th, pi, u)

    for _ in range(args.warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    ms = (time.perf_counter() - t0) / args.steps * 1e3
    v = n_chains * L * L * N_STEPS / (ms * 1e-3)
    sample = (f"each step = 1 trajectory x {n_chains} chains (bounded sample of the 4096-chain workload), "
              f"oracle torch-CPU fp64 port, {torch.get_num_threads()} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(n):
    return {"workload": f"plain HMC C2: L={L} beta={BETA} n_steps={N_STEPS} dt={DT} fp64, "
                        f"{B_PER_GPU} chains per GPU x {n} GPU(s), hot start U(-pi,pi) + 20 thermalisation trajectories",
            "chains_per_gpu": B_PER_GPU, "lattice": L, "leapfrog_steps": N_STEPS,
            "parallelism": f"chain-sharded x{n}: no data-path collective on the device-resident path; the e2e path "
                           f"moves 537 MB per step and rank through ONE host (all GPUs on the same NUMA node), so "
                           f"host PCIe/DRAM bandwidth, not the GPUs, bounds e2e beyond 1-2 ranks",
            "l2_policy": "inputs larger than L2 (theta 268 MB per GPU, 126 MB L2)"}


# ----------------------------------------------------------------------------- ours
def pinned_copy_ceiling(theta_host, out_host, dev, world, reps=3):
    """H2D + D2H of the e2e step's buffers on two streams with NO kernel in between: the PCIe/host ceiling of
    the e2e path on this box with `world` ranks copying at the same time (max over ranks)."""
    d = torch.empty_like(theta_host, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    cur = torch.cuda.current_stream(dev)

    def both():
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            d.copy_(theta_host, non_blocking=True)
        with torch.cuda.stream(s2):
            out_host.copy_(d, non_blocking=True)
        cur.wait_stream(s1); cur.wait_stream(s2)
    return time_steps(both, reps, 1, world)


def run_ours(args):
    import hmc_ft_b200 as hb
    from hmc_ft_b200 import _lib
    from bench_ft import bench_c1, bench_ft, bench_slab, fp64_peak_tflops
    rank, world, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    peak, peak_src = measured_peaks()
    kc = kernel_constants()

    if args.slab_only:          # quick sweep mode (U1_SLAB_HALO / U1_MULTISTEP_EW studies): only the configs[4] leg
        res = bench_slab(args, rank, world, dev, time_steps)
        if rank == 0:
            print(json.dumps({"slab_only": True, "n_gpus": world, "env": {k: v for k, v in os.environ.items() if k.startswith("U1_")},
                              "slab": res}))
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    B = B_PER_GPU
    h = hb.HMC_U1(L, BETA, 0, N_STEPS, DT, device="cuda", if_tune_step_size=False, n_chains=B,
                  chain_offset=rank * B, seed=1331)
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    theta_host = ((torch.rand(B, 2, L, L, generator=g, dtype=torch.float64) * 2 - 1) * math.pi).pin_memory()
    theta = theta_host.to(dev)
    h.trajectories(theta, 20, want_obs=False)           # thermalise (untimed)
    torch.cuda.synchronize()

    # ---- value: resident inputs, one trajectory of all chains per step ----
    h.gpu_launches = 0
    dH_log = []
    with ClockSampler(local) as cs:
        ms = time_steps(lambda: dH_log.append(h.trajectories(theta, 1)["dH"]), args.steps, args.warmup, world)
    launches_per_step = h.gpu_launches // (args.steps + args.warmup)
    su_per_step = B * L * L * N_STEPS
    value = world * su_per_step / (ms * 1e-3)
    # parity across GPU counts: rank 0 always runs global chains 0..4095 from the same start, so the dH of every
    # warm-up + timed trajectory must hash identically at N = 1, 2, 4, 8 (Philox keyed by global chain id)
    parity = {"dH_sha16_rank0_chains": sha16(torch.cat(dH_log)), "trajectories_hashed": len(dH_log),
              "note": "identical at every --gpus N when steps/warmup are equal"}

    # ---- sustained: the same step back to back for >= 2 s (the 30 ms timed region above is a boost-clock burst) ----
    n_sus = max(args.steps, int(2.2e3 / max(ms, 1e-3)))
    with ClockSampler(local) as cs_sus:
        ms_sus = time_steps(lambda: h.trajectories(theta, 1), n_sus, 1, world)
    sustained = {"value": world * su_per_step / (ms_sus * 1e-3), "unit": UNIT, "ms_per_step": ms_sus, "steps": n_sus,
                 "seconds": n_sus * ms_sus * 1e-3, "clocks": cs_sus.summary()}

    # ---- roofline of the dominant kernel (k_traj_resident): one launch == one step here ----
    fp64_peak = fp64_peak_tflops(lib, dev)
    k_res = kc.get("k_traj_resident", {})
    inst_per_su = k_res.get("fp64_thread_inst_per_site_update")
    ach_alg = su_per_step * BYTES_PER_SITE_UPDATE / (ms * 1e-3) / 1e9
    if inst_per_su:
        ach = inst_per_su * su_per_step * 2 / (ms * 1e-3) / 1e12        # FP64-pipe instructions as DFMA-equivalent flops
        ach_sus = inst_per_su * su_per_step * 2 / (ms_sus * 1e-3) / 1e12
    else:
        ach = ach_sus = None
    roofline = {
        "bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
        "frac": (ach / fp64_peak) if ach else None,
        "frac_sustained": (ach_sus / fp64_peak) if ach_sus else None,
        "traffic": k_res.get("dram_bytes_per_launch"),
        "kernel": "k_traj_resident<64,2,4,PF>",
        "how": "achieved = FP64-pipe thread instructions per site-update (ncu sm__inst_executed_pipe_fp64.sum x 32 / "
               "site-updates per launch, a static property of the SASS) x site-updates per launch x 2 / launch time "
               "measured here with CUDA events; peak = DFMA rate of u1_fp64_probe measured in this run; so frac is the "
               "FP64-pipe utilisation (ncu sm__pipe_fp64_cycles_active of the same kernel: see ncu_pipe_fp64_pct)",
        "fp64_thread_inst_per_site_update": inst_per_su, "ncu_pipe_fp64_pct": k_res.get("ncu_pipe_fp64_pct"),
        "traffic_source": k_res.get("source"),
        "peak_source": "measured in this run (u1_fp64_probe DFMA loop)",
        "algorithmic_hbm": {"achieved": ach_alg, "peak": peak, "unit": "GB/s", "frac": ach_alg / peak,
                            "algorithmic_bytes_per_launch": su_per_step * BYTES_PER_SITE_UPDATE, "peak_source": peak_src,
                            "note": "SURVEY §8(d) figure: 64 B x site-updates.  NOT a roofline fraction for this kernel: theta "
                                    "and pi stay in registers for the whole trajectory, real DRAM traffic is `traffic` "
                                    "(~0.33 B per site-update), so the HBM roof does not bind; it binds the streaming "
                                    "step kernel below"}}

    # ---- the kernel the HBM roof does apply to: the streaming fused step kernel (one HBM pass per leapfrog step) ----
    pi = torch.randn(B, 2, L, L, dtype=torch.float64, device=dev)
    to, po = torch.empty_like(theta), torch.empty_like(pi)
    sp = torch.cuda.current_stream().cuda_stream

    def step_kernel():
        _lib.check(lib.u1_leapfrog_step(theta.data_ptr(), pi.data_ptr(), to.data_ptr(), po.data_ptr(),
                                        B, L, BETA, DT, DT, sp))
    ms_k = time_steps(step_kernel, 20, 5, 1)
    ach_k = B * L * L * BYTES_PER_SITE_UPDATE / (ms_k * 1e-3) / 1e9
    k_step = kc.get("k_leapfrog_step", {})
    roofline_stream = {"bound": "hbm", "achieved": ach_k, "peak": peak, "unit": "GB/s", "frac": ach_k / peak,
                       "traffic": k_step.get("dram_bytes_per_launch"), "traffic_source": k_step.get("source"),
                       "algorithmic_bytes_per_launch": B * L * L * BYTES_PER_SITE_UPDATE, "peak_source": peak_src,
                       "kernel": "k_leapfrog_step<false>", "site_updates_per_s": B * L * L / (ms_k * 1e-3)}
    del pi, to, po

    # ---- e2e: host buffers through the public class API, copies inside the timed region ----
    he = hb.HMC_U1(L, BETA, 0, N_STEPS, DT, device="cpu", if_tune_step_size=False, n_chains=B,
                   chain_offset=rank * B, seed=77)
    theta_host.copy_(theta.cpu())
    out_host = torch.empty_like(theta_host).pin_memory()
    obs_bytes = 4 * B * 8

    def e2e_step():
        # public host-buffer call: sliced H2D -> trajectory kernel -> D2H pipeline, one CUDA stream per engine
        he.step_host(theta_host, out_host)
        torch.cuda.current_stream().synchronize()                      # configurations + observables on the host
    ms_e = time_steps(e2e_step, max(3, args.steps // 2), 3, world)
    ms_copy = pinned_copy_ceiling(theta_host, out_host, dev, world)
    e2e = {"value": world * su_per_step / (ms_e * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": theta_host.numel() * 8, "d2h_bytes_per_step": out_host.numel() * 8 + obs_bytes,
           "ms_per_step": ms_e,
           "bound": "host PCIe / host DRAM (pinned copies), not the GPU",
           "pinned_copy_ceiling": {"value": world * su_per_step / (ms_copy * 1e-3), "unit": UNIT, "ms_per_step": ms_copy,
                                   "how": f"the same H2D + D2H copies on two streams with no kernel, {world} rank(s) at once, "
                                          "max over ranks: e2e cannot exceed this on this box"},
           "frac_of_copy_ceiling": ms_copy / ms_e}
    del out_host

    extra = {}
    if not args.no_c1:
        try:
            extra["c1"] = bench_c1(args, rank, world, dev, time_steps)
            if rank == 0 and world == 1 and not args.no_cpu:
                extra["c1"]["cpu_baseline"] = cpu_baseline_c1()
        except Exception as e:
            extra["c1"] = {"error": f"{type(e).__name__}: {e}"}
    if not args.no_ft:
        try:
            extra["ft"] = bench_ft(args, rank, world, dev, time_steps, with_cpu=(world == 1 and not args.no_cpu),
                                   with_e2e=True, ft_cpu_full=args.ft_cpu_full)
        except Exception as e:      # FT numbers are secondary to the contract line
            extra["ft"] = {"error": f"{type(e).__name__}: {e}"}
        # the other FT configurations SURVEY §8(d) names (one timed trajectory each): the smaller shipped CNNs
        # at L=64 beta=6 and BASELINE configs[2] (L=32 beta=3, rnet3 weights trained at beta=3)
        extra["ft_other"] = []
        for arch, L_o, beta_o in (("simple", 64, 6.0), ("rnet1", 64, 6.0), ("rnet3", 32, 3.0)):
            try:
                r = bench_ft(args, rank, world, dev, time_steps, arch=arch, L=L_o, beta=beta_o, steps=1, warmup=1)
                extra["ft_other"].append({"metric": r["metric"], "value": r["value"], "unit": r["unit"],
                                          "ms_per_step": r["ms_per_step"], "accept_rate": r["accept_rate"],
                                          "fp64_frac": r["roofline"]["frac"], "tflops": r["roofline"]["achieved"],
                                          "gpu_launches_per_step": r["gpu_launches_per_step"],
                                          "parity": r["parity"]})
            except Exception as e:
                extra["ft_other"].append({"config": f"{arch} L={L_o} beta={beta_o}", "error": f"{type(e).__name__}: {e}"})

    if not args.no_slab:
        try:
            extra["slab"] = bench_slab(args, rank, world, dev, time_steps)
            extra["slab"]["hbm_frac_of_measured_peak_per_gpu"] = (
                extra["slab"]["value"] / world * BYTES_PER_SITE_UPDATE / 1e9 / peak)
        except Exception as e:
            extra["slab"] = {"error": f"{type(e).__name__}: {e}"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_plain()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world), "clocks": cs.summary(), "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps, "roofline": roofline, "sustained": sustained,
            "parity": parity,
            "roofline_streaming_step_kernel": roofline_stream, "cpu_baseline": cpu,
            "traj_per_s_per_chain_batch": world * B / (ms * 1e-3), **extra,
        }
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-ft", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-slab", action="store_true")
    ap.add_argument("--no-c1", action="store_true")
    ap.add_argument("--slab-only", action="store_true")
    ap.add_argument("--ft-cpu-full", action="store_true",
                    help="time 3 full single-chain FT trajectories on the CPU (minutes) instead of the bounded sample")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
