"""HMC_U1 with the reference's surface (2d_u1_cluster/hmc_u1.py:7-237), computed by hand-written
sm_100a kernels through the C ABI in include/u1hmc.h.

Differences a reference user should know about (all deliberate):
  * arithmetic is fp64 on the GPU regardless of the caller's dtype; `device` names where the
    caller's tensors live ("cpu": host buffers in/out, copies inside the call; "cuda": resident);
  * every method also accepts a leading chain axis: theta [B,2,L,L] -> per-chain results [B];
  * momenta / accept draws come from an on-device Philox4x32-10 stream keyed by
    (seed, global chain id, trajectory index); `metropolis_step(theta, pi=..., u=...)` injects
    the draws instead (parity tests, golden-log replay);
  * the constructor does not touch torch's global default dtype/device/seed (hmc_u1.py:49-51).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._runtime import WS, as_batch, check_field, ptr, require_cuda, stream_ptr, to_dev, tune_search

try:                                    # progress bars are cosmetic
    from tqdm import tqdm
except Exception:                       # pragma: no cover
    def tqdm(x, **_):
        return x


class HMC_U1:
    def __init__(self, lattice_size, beta, n_thermalization_steps, n_steps, step_size, device="cuda",
                 if_tune_step_size=True, n_chains: Optional[int] = None, seed: int = 1331,
                 chain_offset: int = 0, kernel_path: int = 0, verbose: bool = False):
        """Same leading arguments as hmc_u1.py:8-17.  Extra keyword arguments:
        n_chains     batch of independent chains (None: single chain, reference return types)
        seed         Philox seed (default echoes torch.manual_seed(1331), hmc_u1.py:51)
        chain_offset global id of this process's first chain (chain sharding across GPUs)
        kernel_path  0 auto, 1 streaming step kernels, 2 SM-resident trajectory kernel"""
        if lattice_size < 2 or lattice_size % 2:
            raise ValueError("lattice_size must be even and >= 2")
        self.lattice_size = int(lattice_size)
        self.beta = float(beta)
        self.n_thermalization_steps = int(n_thermalization_steps)
        self.n_steps = int(n_steps)
        self.dt = float(step_size)
        self.device = torch.device(device)
        self.if_tune_step_size = if_tune_step_size
        self.n_chains = n_chains
        self.seed = int(seed)
        self.chain_offset = int(chain_offset)
        self.kernel_path = int(kernel_path)
        self.verbose = verbose
        self.traj_counter = 0              # Philox trajectory index, advances with every step
        self.gpu_launches = 0
        self._lib = _lib.load()            # fail loudly now if the CUDA library is missing
        self._dev = require_cuda(self.device if self.device.type == "cuda" else None)

    # ------------------------------------------------------------------ helpers
    def _print(self, *a):
        if self.verbose:
            print(*a)

    def _stage(self, t: torch.Tensor):
        tb, single = as_batch(t)
        d = to_dev(tb, self._dev)
        check_field(d)
        return d, single

    def _back(self, t: torch.Tensor, like: torch.Tensor, single: bool = False):
        if single:
            t = t[0]
        return t if like.device == t.device else t.to(like.device)

    def _ws(self, B: int):
        return WS.get(self._dev, self._lib.u1_workspace_bytes(B, self.lattice_size))

    # ------------------------------------------------------------------ reference surface
    def initialize(self):
        """hmc_u1.py:53-54 (cold start)."""
        L = self.lattice_size
        shape = [2, L, L] if self.n_chains is None else [self.n_chains, 2, L, L]
        return torch.zeros(shape, dtype=torch.float64, device=self.device)

    def action(self, theta):
        """hmc_u1.py:56-63."""
        d, single = self._stage(theta)
        B, L = check_field(d)
        out = torch.empty(B, dtype=torch.float64, device=self._dev)
        with torch.cuda.device(self._dev):
            ws, nb = self._ws(B)
            _lib.check(self._lib.u1_action(d.data_ptr(), out.data_ptr(), B, L, self.beta, ws, nb,
                                           stream_ptr(self._dev)), "u1_action")
        return self._back(out, theta, single)

    def force(self, theta):
        """hmc_u1.py:65-69."""
        d, single = self._stage(theta)
        B, L = check_field(d)
        out = torch.empty_like(d)
        with torch.cuda.device(self._dev):
            _lib.check(self._lib.u1_force(d.data_ptr(), out.data_ptr(), B, L, self.beta,
                                          stream_ptr(self._dev)), "u1_force")
        return self._back(out, theta, single)

    def leapfrog(self, theta, pi):
        """hmc_u1.py:71-80."""
        d, single = self._stage(theta)
        p, _ = self._stage(pi)
        B, L = check_field(d)
        to, po = torch.empty_like(d), torch.empty_like(d)
        with torch.cuda.device(self._dev):
            ws, nb = self._ws(B)
            _lib.check(self._lib.u1_leapfrog(d.data_ptr(), p.data_ptr(), to.data_ptr(), po.data_ptr(), B, L,
                                             self.beta, self.dt, self.n_steps, ws, nb,
                                             stream_ptr(self._dev)), "u1_leapfrog")
        return self._back(to, theta, single), self._back(po, theta, single)

    def trajectories(self, theta_dev: torch.Tensor, n_traj: int = 1, pi=None, u=None, want_obs: bool = True,
                     _seed=None, _chain0=None, _traj0=None, _advance: bool = True):
        """n_traj Metropolis steps on the device-resident batch `theta_dev` ([B,2,L,L] fp64, updated
        IN PLACE).  Returns dict of device tensors [n_traj,B]: dH, accept, H, plaq, topo.
        (_seed/_chain0/_traj0/_advance: Philox stream override used by tune_step_size.)"""
        B, L = check_field(theta_dev)
        dev = self._dev
        seed = self.seed if _seed is None else int(_seed)
        chain0 = self.chain_offset if _chain0 is None else int(_chain0)
        traj0 = self.traj_counter if _traj0 is None else int(_traj0)
        out = {k: torch.empty((n_traj, B), dtype=torch.float64, device=dev) for k in ("dH", "H", "plaq", "topo")}
        out["accept"] = torch.empty((n_traj, B), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            ws, nb = self._ws(B)
            _lib.check(self._lib.u1_trajectory(
                theta_dev.data_ptr(), ptr(pi), ptr(u), seed, chain0, traj0, n_traj,
                B, L, self.beta, self.dt, self.n_steps,
                out["dH"].data_ptr(), out["accept"].data_ptr(), out["H"].data_ptr(),
                out["plaq"].data_ptr() if want_obs else None, out["topo"].data_ptr() if want_obs else None,
                self.kernel_path, ws, nb, stream_ptr(dev)), "u1_trajectory")
        self.gpu_launches += _lib.last_launch_count()
        if _advance:
            self.traj_counter += n_traj
        return out

    def step_host(self, theta_host: torch.Tensor, out_host: Optional[torch.Tensor] = None, n_slices=12):
        """One Metropolis step of a HOST-resident batch [B,2,L,L] (fp64, ideally pinned): the batch is cut
        into `n_slices` slices (or, with n_slices="tapered", ten slices of 1/32 .. 1/4 .. 1/32 of the batch).
        Three CUDA streams, one per engine: ALL host-to-device copies queue back to back on the first, the
        trajectory kernels of the slices on the second (each waits for its slice's copy event), ALL
        device-to-host copies on the third (each waits for its slice's kernel event), so neither PCIe direction
        ever waits for the other and only the first copy-in and the last copy-out are exposed.
        Returns (theta_out_host, obs_host) with obs_host = pinned fp64 [4,B] = (accept, H, plaq, topo).
        Chains keep their global Philox ids (chain_offset + index), so the result equals the resident path bit
        for bit.  The D2H copies are ASYNCHRONOUS on the current stream: synchronize it (or the device) before
        the host reads `theta_out_host` / `obs_host`."""
        if theta_host.device.type != "cpu" or theta_host.dim() != 4 or theta_host.dtype != torch.float64:
            raise ValueError("step_host expects a CPU fp64 tensor [B,2,L,L]")
        B, L = check_field(theta_host)
        dev = self._dev
        if not theta_host.is_pinned():
            theta_host = theta_host.pin_memory()
        if out_host is None:
            out_host = torch.empty_like(theta_host).pin_memory()
        if n_slices in (0, "tapered"):
            # small slices at both ends, big ones in the middle: the pipeline fill (first H2D) and drain (last
            # kernel + D2H) shrink from 1/8 to 1/32 of the batch while the launch count stays ~10
            frac = [1, 1, 2, 4, 8, 8, 4, 2, 1, 1]
            cuts = [0]
            for f in frac:
                cuts.append(cuts[-1] + f)
            bounds = sorted({(c * B) // cuts[-1] for c in cuts})
        else:
            n_slices = max(1, min(int(n_slices), B))
            bounds = [(i * B) // n_slices for i in range(n_slices + 1)]
        n_slices = len(bounds) - 1
        key = (B, L, tuple(bounds))
        st = getattr(self, "_host_state", None)
        if st is None or st["key"] != key:
            st = {"key": key, "dev": torch.empty((B, 2, L, L), dtype=torch.float64, device=dev),
                  "obs_d": torch.empty((4, B), dtype=torch.float64, device=dev),
                  "obs_h": torch.empty((4, B), dtype=torch.float64).pin_memory(),
                  "acc": torch.empty(B, dtype=torch.int32, device=dev),
                  "streams": [torch.cuda.Stream(device=dev) for _ in range(3)],      # H2D, kernels, D2H
                  "ev_in": [torch.cuda.Event() for _ in range(n_slices)],
                  "ev_k": [torch.cuda.Event() for _ in range(n_slices)]}
            self._host_state = st
        cur = torch.cuda.current_stream(dev)
        s_in, s_k, s_out = st["streams"]
        widest = max(b - a for a, b in zip(bounds[:-1], bounds[1:]))
        with torch.cuda.device(dev):
            ws, nb = WS.get(dev, self._lib.u1_workspace_bytes(widest + 1, L), "host0")
            for s in st["streams"]:
                s.wait_stream(cur)
            o = st["obs_d"]
            for i in range(n_slices):
                a, b = bounds[i], bounds[i + 1]
                d = st["dev"][a:b]
                with torch.cuda.stream(s_in):
                    d.copy_(theta_host[a:b], non_blocking=True)
                    st["ev_in"][i].record(s_in)
                with torch.cuda.stream(s_k):
                    s_k.wait_event(st["ev_in"][i])
                    _lib.check(self._lib.u1_trajectory(
                        d.data_ptr(), None, None, self.seed, self.chain_offset + a, self.traj_counter, 1,
                        b - a, L, self.beta, self.dt, self.n_steps, None, st["acc"][a:b].data_ptr(),
                        o[1, a:b].data_ptr(), o[2, a:b].data_ptr(), o[3, a:b].data_ptr(), self.kernel_path,
                        ws, nb, s_k.cuda_stream), "u1_trajectory")
                    self.gpu_launches += _lib.last_launch_count()
                    o[0, a:b].copy_(st["acc"][a:b])
                    st["ev_k"][i].record(s_k)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(st["ev_k"][i])
                    out_host[a:b].copy_(d, non_blocking=True)
            for s in st["streams"]:
                cur.wait_stream(s)
            st["obs_h"].copy_(st["obs_d"], non_blocking=True)      # one contiguous 4xB read-back
        self.traj_counter += 1
        return out_host, st["obs_h"]

    def metropolis_step(self, theta, pi=None, u=None):
        """hmc_u1.py:82-97.  Single chain: (theta, bool, float) like the reference;
        batch: (theta [B,2,L,L], accepted bool[B], H[B])."""
        d, single = self._stage(theta)
        if d.data_ptr() == theta.data_ptr():
            d = d.clone()                       # the reference never mutates its argument
        B, L = check_field(d)
        pi_d = None if pi is None else to_dev(as_batch(pi)[0], self._dev).reshape(1, B, 2, L, L)
        u_d = None if u is None else to_dev(torch.as_tensor(u, dtype=torch.float64).reshape(-1), self._dev)
        if single:
            # one chain (the reference's own mode, BASELINE configs[0]): latency matters, not throughput - the kernel
            # writes H (fp64) and the accept flag (int32) into ONE 16-byte device buffer, read back with a single
            # synchronising copy (the reference also syncs here: it returns Python scalars)
            dev = self._dev
            buf = torch.empty(2, dtype=torch.float64, device=dev)
            with torch.cuda.device(dev):
                ws, nb = self._ws(1)
                _lib.check(self._lib.u1_trajectory(
                    d.data_ptr(), ptr(pi_d), ptr(u_d), self.seed, self.chain_offset, self.traj_counter, 1,
                    1, L, self.beta, self.dt, self.n_steps, None, buf.data_ptr() + 8, buf.data_ptr(), None, None,
                    self.kernel_path, ws, nb, stream_ptr(dev)), "u1_trajectory")
            self.gpu_launches += _lib.last_launch_count()
            self.traj_counter += 1
            h = buf.cpu()
            return self._back(d, theta, True), bool(h.view(torch.int32)[2].item()), h[0].item()
        r = self.trajectories(d, 1, pi_d, u_d, want_obs=False)
        return self._back(d, theta), self._back(r["accept"][0].bool(), theta), self._back(r["H"][0], theta)

    def tune_step_size(self, n_tune_steps=2000, target_rate=0.65, target_tolerance=0.15,
                       initial_step_size=0.2, max_attempts=10, theta=None, draws=None):
        """hmc_u1.py:99-161: binary search on dt.  As in the reference, every trial trajectory
        starts from the same `theta` (the returned state is discarded, hmc_u1.py:132), so the trials of
        one attempt are independent and run as extra chains of ONE launch (_runtime.tune_search).  With a
        batch of chains the acceptance rate is the mean over chains.  draws=(pi, u) injects the draws in
        the order the reference consumes them (parity test).  Returns the [(dt, rate)] history."""
        theta = self.initialize() if theta is None else theta
        base, _ = self._stage(theta)
        return tune_search(self, base, n_tune_steps, target_rate, target_tolerance, initial_step_size,
                           max_attempts, draws)

    def thermalize(self):
        """hmc_u1.py:163-196: rough thermalisation, optional tuning, restart from cold, then
        n_thermalization_steps trajectories recording the plaquette BEFORE each step."""
        theta = self.initialize()
        d, single = self._stage(theta)
        d = d.clone()
        n = self.n_thermalization_steps
        if n > 0:
            self.trajectories(d, n, want_obs=False)
        if self.if_tune_step_size:
            self.tune_step_size(initial_step_size=self.dt, theta=d)
        else:
            self._print(f">>> Using step size: {self.dt:.2f}")
        d.zero_()
        B = d.shape[0]
        if n == 0:
            return self._back(d, theta, single), [], 0.0
        from .utils import plaq_mean_from_field
        plaq0 = plaq_mean_from_field(d).reshape(1, B)
        r = self.trajectories(d, n)
        plaq = torch.cat([plaq0, r["plaq"][:-1]], dim=0)       # value before each step
        acc = r["accept"].double().mean(dim=0)
        if single:
            return self._back(d, theta, True), plaq[:, 0].tolist(), acc[0].item()
        return self._back(d, theta), plaq.cpu(), acc.cpu()

    def run(self, n_iterations, theta, store_interval=1, save_config=True):
        """hmc_u1.py:198-237.  Observables are accumulated on the device and read back once.
        Single chain: Python lists exactly like the reference; batch: tensors [n_store,B]."""
        d, single = self._stage(theta)
        d = d.clone()
        B, L = check_field(d)
        theta_ls, plaq, topo, ham, acc = [], [], [], [], []
        done = 0
        si = int(store_interval)
        while done < n_iterations:
            if save_config:      # snapshot right after each stored iteration
                chunk = 1 if done % si == 0 else min(si - done % si, n_iterations - done)
            else:                # no snapshots: many trajectories per launch
                chunk = min(512, n_iterations - done)
            r = self.trajectories(d, chunk)
            for k in range(chunk):
                if (done + k) % si == 0:
                    plaq.append(r["plaq"][k]); topo.append(r["topo"][k]); ham.append(r["H"][k])
                    if save_config:
                        theta_ls.append(self._back(d.clone(), theta, single))
            acc.append(r["accept"].double().sum(dim=0))
            done += chunk
        rate = torch.stack(acc).sum(dim=0) / n_iterations
        plaq, topo, ham = torch.stack(plaq), torch.stack(topo), torch.stack(ham)
        if single:
            return theta_ls, plaq[:, 0].tolist(), rate[0].item(), topo[:, 0].tolist(), ham[:, 0].tolist()
        return theta_ls, plaq.cpu(), rate.cpu(), topo.cpu(), ham.cpu()
