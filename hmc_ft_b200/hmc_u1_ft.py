"""HMC_U1_FT with the reference's surface (2d_u1_cluster/hmc_u1_ft.py:7-370) on the sm_100a library.

The reference takes two callables (field_transformation, compute_jac_logdet) and differentiates
through them with autograd.  Here the callables must be bound methods of
hmc_ft_b200.FieldTransformation (the fused CUDA path uses its weights), or an identity pair
(`lambda x: x`, zero log-det; debug/check_hmc.py:75-78) which selects the plain kernels.  Arbitrary
Python callables cannot be fused and are rejected loudly."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._runtime import WS, as_batch, check_field, ptr, require_cuda, stream_ptr, to_dev, tune_search
from .field_trans import FieldTransformation
from .hmc_u1 import HMC_U1
from .utils import plaq_mean_from_field, regularize


def _resolve_ft(field_transformation, compute_jac_logdet, L) -> Optional[FieldTransformation]:
    owner = getattr(field_transformation, "__self__", None)
    if isinstance(field_transformation, FieldTransformation):
        owner = field_transformation
    if isinstance(owner, FieldTransformation):
        o2 = getattr(compute_jac_logdet, "__self__", owner)
        if o2 is not owner:
            raise TypeError("field_transformation and compute_jac_logdet belong to different FieldTransformation objects")
        return owner
    # identity pair?  probe on the host, no arithmetic of ours involved
    probe = torch.arange(2 * L * L, dtype=torch.float64).reshape(2, L, L) * 0.01
    try:
        same = torch.equal(torch.as_tensor(field_transformation(probe)).reshape(2, L, L).double(), probe)
        ld = torch.as_tensor(compute_jac_logdet(probe.unsqueeze(0))).double().reshape(-1)
        if same and bool(torch.all(ld == 0)):
            return None
    except Exception:
        pass
    raise TypeError("HMC_U1_FT needs FieldTransformation-bound callables (fused CUDA path) or an identity pair; "
                    "arbitrary Python callables cannot run on the device")


class HMC_U1_FT:
    def __init__(self, lattice_size, beta, n_thermalization_steps, n_steps, step_size, field_transformation,
                 compute_jac_logdet, device="cuda", if_tune_step_size=True, n_chains: Optional[int] = None,
                 seed: int = 1331, chain_offset: int = 0, verbose: bool = False):
        """Leading arguments as hmc_u1_ft.py:8-19; extras as in HMC_U1."""
        self.lattice_size = int(lattice_size)
        self.beta = float(beta)
        self.n_thermalization_steps = int(n_thermalization_steps)
        self.n_steps = int(n_steps)
        self.dt = float(step_size)
        self.field_transformation = field_transformation
        self.compute_jac_logdet = compute_jac_logdet
        self.device = torch.device(device)
        self.if_tune_step_size = if_tune_step_size
        self.n_chains = n_chains
        self.seed, self.chain_offset, self.verbose = int(seed), int(chain_offset), verbose
        self.traj_counter = 0
        self.gpu_launches = 0
        self.graph_launches = 0
        self._lib = _lib.load()
        self._dev = require_cuda(self.device if self.device.type == "cuda" else None)
        self.ft = _resolve_ft(field_transformation, compute_jac_logdet, self.lattice_size)
        if self.ft is not None and self.ft.L != self.lattice_size:
            raise ValueError("FieldTransformation lattice size differs from HMC_U1_FT lattice size")
        if self.ft is not None and self.ft._dev != self._dev:
            raise ValueError(f"FieldTransformation lives on {self.ft._dev}, HMC_U1_FT on {self._dev}: "
                             "the weight handle, the workspace and theta must share one device")
        self._plain = None
        if self.ft is None:     # identity transformation: the plain kernels are the exact same chain
            self._plain = HMC_U1(lattice_size, beta, n_thermalization_steps, n_steps, step_size, device=device,
                                 if_tune_step_size=if_tune_step_size, n_chains=n_chains, seed=seed,
                                 chain_offset=chain_offset, verbose=verbose)

    # ------------------------------------------------------------------ helpers
    def _print(self, *a):
        if self.verbose:
            print(*a)

    def _stage(self, t):
        tb, single = as_batch(t)
        d = to_dev(tb, self._dev)
        check_field(d)
        return d, single

    def _back(self, t, like, single=False):
        if single:
            t = t[0]
        return t if like.device == t.device else t.to(like.device)

    def _transform(self, d):
        return d if self.ft is None else self.ft.forward(d)

    # ------------------------------------------------------------------ reference surface
    def initialize(self):
        L = self.lattice_size
        shape = [2, L, L] if self.n_chains is None else [self.n_chains, 2, L, L]
        return torch.zeros(shape, dtype=torch.float64, device=self.device)

    def original_action(self, theta):
        """hmc_u1_ft.py:77-98."""
        h = self._plain or HMC_U1(self.lattice_size, self.beta, 0, self.n_steps, self.dt, device=self.device,
                                  if_tune_step_size=False)
        return h.action(theta)

    def new_action(self, theta_new):
        """hmc_u1_ft.py:101-125."""
        if self.ft is None:
            return self._plain.action(theta_new)
        return self.ft.new_action(theta_new, self.beta)

    def new_force(self, theta_new):
        """hmc_u1_ft.py:128-152."""
        if self.ft is None:
            return self._plain.force(theta_new)
        return self.ft.new_force(theta_new, self.beta)

    def leapfrog(self, theta, pi):
        """hmc_u1_ft.py:154-179."""
        if self.ft is None:
            self._plain.dt = self.dt
            return self._plain.leapfrog(theta, pi)
        d, single = self._stage(theta)
        p, _ = self._stage(pi)
        B, L = check_field(d)
        to, po = torch.empty_like(d), torch.empty_like(d)
        with torch.cuda.device(self._dev):
            ws, nb = self.ft.workspace(B, True)
            _lib.check(self._lib.u1ft_leapfrog(self.ft._handle, d.data_ptr(), p.data_ptr(), to.data_ptr(), po.data_ptr(),
                                               B, L, self.beta, self.dt, self.n_steps, ws, nb,
                                               stream_ptr(self._dev)), "u1ft_leapfrog")
        self.gpu_launches += _lib.last_launch_count()
        return self._back(to, theta, single), self._back(po, theta, single)

    def trajectories(self, theta_dev, n_traj=1, pi=None, u=None, want_ori=False,
                     _seed=None, _chain0=None, _traj0=None, _advance: bool = True):
        """n_traj FT-HMC Metropolis steps on the device batch theta_dev (in place).  Returns dict of
        device tensors [n_traj,B]: dH, accept, H, plaq, topo (observables of F(theta), as
        hmc_u1_ft.py:348-358) and optionally theta_ori [B,2,L,L] of the final state."""
        if self.ft is None:
            self._plain.dt, self._plain.traj_counter = self.dt, self.traj_counter
            r = self._plain.trajectories(theta_dev, n_traj, pi, u, _seed=_seed, _chain0=_chain0, _traj0=_traj0,
                                         _advance=_advance)
            if _advance:
                self.traj_counter += n_traj
            self.gpu_launches = self._plain.gpu_launches
            if want_ori:
                r["theta_ori"] = theta_dev.clone()
            return r
        B, L = check_field(theta_dev)
        dev = self._dev
        seed = self.seed if _seed is None else int(_seed)
        chain0 = self.chain_offset if _chain0 is None else int(_chain0)
        traj0 = self.traj_counter if _traj0 is None else int(_traj0)
        out = {k: torch.empty((n_traj, B), dtype=torch.float64, device=dev) for k in ("dH", "H", "plaq", "topo")}
        out["accept"] = torch.empty((n_traj, B), dtype=torch.int32, device=dev)
        ori = torch.empty_like(theta_dev) if want_ori else None
        with torch.cuda.device(dev):
            ws, nb = self.ft.workspace(B, True)
            _lib.check(self._lib.u1ft_trajectory(
                self.ft._handle, theta_dev.data_ptr(), ptr(pi), ptr(u), seed, chain0,
                traj0, n_traj, B, L, self.beta, self.dt, self.n_steps,
                out["dH"].data_ptr(), out["accept"].data_ptr(), out["H"].data_ptr(), out["plaq"].data_ptr(),
                out["topo"].data_ptr(), ptr(ori), ws, nb, stream_ptr(dev)), "u1ft_trajectory")
        self.gpu_launches += _lib.last_launch_count()
        self.graph_launches += int(self._lib.u1ft_last_host_launch_count())     # host launch calls (graph replay = 1)
        if _advance:
            self.traj_counter += n_traj
        if want_ori:
            out["theta_ori"] = ori
        return out

    def step_host(self, theta_host: torch.Tensor, out_host: Optional[torch.Tensor] = None):
        """One FT-HMC Metropolis step of a HOST-resident batch [B,2,L,L] (fp64, ideally pinned): H2D copy,
        u1ft_trajectory, D2H of the new configurations and of the per-chain (dH, accept, H, plaq, topo) as one
        pinned [5,B] block.  The step is compute-bound (seconds of FP64 work against milliseconds of PCIe), so
        unlike HMC_U1.step_host there is nothing to pipeline.  The D2H copies are ASYNCHRONOUS on the current
        stream: synchronize before the host reads the results."""
        if theta_host.device.type != "cpu" or theta_host.dim() != 4 or theta_host.dtype != torch.float64:
            raise ValueError("step_host expects a CPU fp64 tensor [B,2,L,L]")
        B, L = check_field(theta_host)
        if not theta_host.is_pinned():
            theta_host = theta_host.pin_memory()
        if out_host is None:
            out_host = torch.empty_like(theta_host).pin_memory()
        st = getattr(self, "_host_state", None)
        if st is None or st["key"] != (B, L):
            st = {"key": (B, L), "dev": torch.empty((B, 2, L, L), dtype=torch.float64, device=self._dev),
                  "obs_d": torch.empty((5, B), dtype=torch.float64, device=self._dev),
                  "obs_h": torch.empty((5, B), dtype=torch.float64).pin_memory()}
            self._host_state = st
        st["dev"].copy_(theta_host, non_blocking=True)
        r = self.trajectories(st["dev"], 1)
        o = st["obs_d"]
        o[0].copy_(r["dH"][0]); o[1].copy_(r["accept"][0]); o[2].copy_(r["H"][0]); o[3].copy_(r["plaq"][0]); o[4].copy_(r["topo"][0])
        out_host.copy_(st["dev"], non_blocking=True)
        st["obs_h"].copy_(o, non_blocking=True)
        return out_host, st["obs_h"]

    def metropolis_step(self, theta, pi=None, u=None):
        """hmc_u1_ft.py:181-209."""
        d, single = self._stage(theta)
        if d.data_ptr() == theta.data_ptr():
            d = d.clone()
        B, L = check_field(d)
        pi_d = None if pi is None else to_dev(as_batch(pi)[0], self._dev).reshape(1, B, 2, L, L)
        u_d = None if u is None else to_dev(torch.as_tensor(u, dtype=torch.float64).reshape(-1), self._dev)
        r = self.trajectories(d, 1, pi_d, u_d)
        if single:
            return self._back(d, theta, True), bool(r["accept"][0, 0].item()), r["H"][0, 0].item()
        return self._back(d, theta), self._back(r["accept"][0].bool(), theta), self._back(r["H"][0], theta)

    def tune_step_size(self, n_tune_steps=1000, target_rate=0.65, target_tolerance=0.15,
                       initial_step_size=0.2, max_attempts=10, theta=None, draws=None):
        """hmc_u1_ft.py:211-273: binary search on dt; every trial starts from the same theta (as in the
        reference), so one attempt's trials run as extra chains of one launch (_runtime.tune_search)."""
        theta = self.initialize() if theta is None else theta
        base, _ = self._stage(theta)
        return tune_search(self, base, n_tune_steps, target_rate, target_tolerance, initial_step_size,
                           max_attempts, draws)

    def thermalize(self):
        """hmc_u1_ft.py:275-316: plaquette of F(theta) recorded BEFORE each step."""
        theta = self.initialize()
        d, single = self._stage(theta)
        d = d.clone()
        n = self.n_thermalization_steps
        if n > 0:
            self.trajectories(d, n)
        if self.if_tune_step_size:
            self.tune_step_size(theta=d)
        d.zero_()
        B = d.shape[0]
        if n == 0:
            return self._back(d, theta, single), [], 0.0
        plaq0 = plaq_mean_from_field(self._transform(d)).reshape(1, B)
        r = self.trajectories(d, n)
        plaq = torch.cat([plaq0, r["plaq"][:-1]], dim=0)
        acc = r["accept"].double().mean(dim=0)
        if single:
            return self._back(d, theta, True), plaq[:, 0].tolist(), acc[0].item()
        return self._back(d, theta), plaq.cpu(), acc.cpu()

    def run(self, n_iterations, theta, store_interval=1, save_config=True):
        """hmc_u1_ft.py:318-370.  Stored configurations are regularize(F(theta)) moved to the host."""
        d, single = self._stage(theta)
        d = d.clone()
        theta_ls, plaq, topo, ham, acc = [], [], [], [], []
        done, si = 0, int(store_interval)
        while done < n_iterations:
            if save_config:
                chunk = 1 if done % si == 0 else min(si - done % si, n_iterations - done)
            else:
                chunk = min(64, n_iterations - done)
            stored_here = save_config and done % si == 0
            r = self.trajectories(d, chunk, want_ori=stored_here)
            for k in range(chunk):
                if (done + k) % si == 0:
                    plaq.append(r["plaq"][k]); topo.append(r["topo"][k]); ham.append(r["H"][k])
                    if save_config:
                        ori = regularize(r["theta_ori"]).cpu()
                        theta_ls.append(ori[0] if single else ori)
            acc.append(r["accept"].double().sum(dim=0))
            done += chunk
        rate = torch.stack(acc).sum(dim=0) / n_iterations
        plaq, topo, ham = torch.stack(plaq), torch.stack(topo), torch.stack(ham)
        if single:
            return theta_ls, plaq[:, 0].tolist(), rate[0].item(), topo[:, 0].tolist(), ham[:, 0].tolist()
        return theta_ls, plaq.cpu(), rate.cpu(), topo.cpu(), ham.cpu()
