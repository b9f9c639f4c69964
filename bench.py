#!/usr/bin/env python
"""Benchmark of the 2D U(1) HMC / FT-HMC hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one HMC trajectory (n_steps=50 leapfrog steps + momentum refresh + Metropolis +
observables) of every chain of BASELINE.json configs[1]: plain HMC, L=64, beta=6, B=4096 chains per
GPU, fp64.  metric = site-updates/s (1 site-update = one site, 2 links, through one leapfrog step).
Chains are sharded over ranks with no data-path collective (weak scaling: 4096 chains per GPU).
The same JSON line carries the FT-HMC trajectories/s figure (configs[3]) under "ft".
"""
from __future__ import annotations

import argparse
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
FT_L, FT_BETA, FT_N_STEPS, FT_DT, FT_B_TOTAL, FT_ARCH = 64, 6.0, 50, 0.05, 1024, "rnet3"
BYTES_PER_SITE_UPDATE = 64.0      # read+write theta0,theta1,pi0,pi1 in fp64
UNIT = "site-updates/s"
METRIC = "site-updates/sec (force+leapfrog), plain HMC L=64 beta=6 B=4096 chains/GPU fp64"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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


# ----------------------------------------------------------------------------- CPU baseline
def cpu_baseline_plain(n_chains=128, n_traj=4):
    """Oracle port (oracle/u1_oracle.py, torch CPU fp64, autograd force like hmc_u1.py:65-69),
    batched over chains so every host thread has work; bounded sample of the C2 workload."""
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

    one()
    t0 = time.perf_counter()
    for _ in range(n_traj):
        one()
    dt = (time.perf_counter() - t0) / n_traj
    return {"value": n_chains * L * L * N_STEPS / dt, "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": "port",
            "sample": f"{n_traj} trajectories x {n_chains} chains of the C2 workload (L=64 beta=6 50 steps), "
                      f"oracle torch-CPU fp64 port batched over chains, {dt:.3f} s/trajectory"}, dt


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
            "parallelism": f"chain-sharded x{n} (no data-path collective)",
            "l2_policy": "inputs larger than L2 (theta 268 MB per GPU, 126 MB L2)"}


# ----------------------------------------------------------------------------- ours
def run_ours(args):
    import hmc_ft_b200 as hb
    from hmc_ft_b200 import _lib
    rank, world, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    peak, peak_src = measured_peaks()

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
    with ClockSampler(local) as cs:
        ms = time_steps(lambda: h.trajectories(theta, 1), args.steps, args.warmup, world)
    launches_per_step = h.gpu_launches // (args.steps + args.warmup)
    su_per_step = B * L * L * N_STEPS
    value = world * su_per_step / (ms * 1e-3)

    # ---- roofline of the dominant kernel (k_traj_resident): one launch == one step here ----
    ach = su_per_step * BYTES_PER_SITE_UPDATE / (ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": 273.3e6, "traffic_source": "ncu --set full, dram__bytes_read.sum+dram__bytes_write.sum per launch "
                                                       "(profiles/r01_ncu_full_k_traj_resident_v5.txt)",
                "algorithmic_bytes_per_launch": su_per_step * BYTES_PER_SITE_UPDATE,
                "kernel": "k_traj_resident<64,2,4>", "peak_source": peak_src,
                "note": "algorithmic bytes = 64 B x site-updates (SURVEY §8d); the kernel keeps theta,pi in "
                        "registers for the whole trajectory so real DRAM traffic is ~n_steps x smaller: a frac "
                        "above 1 is expected; the binding unit is the FP64 pipe (see profiles/)"}

    # ---- secondary: the streaming fused step kernel (one HBM pass per leapfrog step) ----
    pi = torch.randn(B, 2, L, L, dtype=torch.float64, device=dev)
    to, po = torch.empty_like(theta), torch.empty_like(pi)
    sp = torch.cuda.current_stream().cuda_stream

    def step_kernel():
        _lib.check(lib.u1_leapfrog_step(theta.data_ptr(), pi.data_ptr(), to.data_ptr(), po.data_ptr(),
                                        B, L, BETA, DT, DT, sp))
    ms_k = time_steps(step_kernel, 20, 5, 1)
    ach_k = B * L * L * BYTES_PER_SITE_UPDATE / (ms_k * 1e-3) / 1e9
    roofline_stream = {"bound": "hbm", "achieved": ach_k, "peak": peak, "unit": "GB/s", "frac": ach_k / peak,
                       "traffic": 1028.7e6, "traffic_source": "ncu --set full (profiles/r01_ncu_full_k_leapfrog_step_v1.txt)",
                       "algorithmic_bytes_per_launch": B * L * L * BYTES_PER_SITE_UPDATE,
                       "kernel": "k_leapfrog_step<false>", "site_updates_per_s": B * L * L / (ms_k * 1e-3)}

    # ---- e2e: host buffers through the public class API, copies inside the timed region ----
    he = hb.HMC_U1(L, BETA, 0, N_STEPS, DT, device="cpu", if_tune_step_size=False, n_chains=B,
                   chain_offset=rank * B, seed=77)
    theta_host.copy_(theta.cpu())
    out_host = torch.empty_like(theta_host).pin_memory()
    obs_bytes = 4 * B * 8

    def e2e_step():
        # public host-buffer call: sliced H2D -> trajectory kernel -> D2H pipeline over CUDA streams
        he.step_host(theta_host, out_host, n_slices=8)
        torch.cuda.current_stream().synchronize()                      # configurations + observables on the host
    ms_e = time_steps(e2e_step, max(3, args.steps // 2), 3, world)
    e2e = {"value": world * su_per_step / (ms_e * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": theta_host.numel() * 8, "d2h_bytes_per_step": out_host.numel() * 8 + obs_bytes,
           "ms_per_step": ms_e}

    extra = {}
    if not args.no_ft:
        try:
            from bench_ft import bench_ft
            extra["ft"] = bench_ft(args, rank, world, dev, time_steps)
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
                                          "fp64_frac": r["roofline"]["frac"], "tflops": r["roofline"]["achieved"]})
            except Exception as e:
                extra["ft_other"].append({"config": f"{arch} L={L_o} beta={beta_o}", "error": f"{type(e).__name__}: {e}"})

    if not args.no_slab:
        try:
            from bench_ft import bench_slab
            extra["slab"] = bench_slab(args, rank, world, dev, time_steps)
            extra["slab"]["hbm_frac_of_measured_peak_per_gpu"] = (
                extra["slab"]["value"] / world * BYTES_PER_SITE_UPDATE / 1e9 / peak)
        except Exception as e:
            extra["slab"] = {"error": f"{type(e).__name__}: {e}"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu, _ = cpu_baseline_plain()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world), "clocks": cs.summary(), "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps, "roofline": roofline,
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
