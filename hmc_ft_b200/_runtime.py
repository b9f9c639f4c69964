"""Host-side plumbing shared by the reference-facing classes: device selection, fp64 staging of
caller tensors (host or device), cached workspaces and the current CUDA stream.  torch is used
only for memory, streams and torch.distributed — all arithmetic is in libu1hmc.so."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.U1Error("hmc_ft_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.type != "cuda":
        dev = torch.device("cuda", torch.cuda.current_device())
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def stream_ptr(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def to_dev(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    """Contiguous fp64 copy/view of `t` on the compute device (H2D copy if `t` is on the host)."""
    if t.device != dev:
        if t.device.type == "cpu" and not t.is_pinned() and t.numel() >= (1 << 16):
            t = t.pin_memory()
        t = t.to(dev, non_blocking=True)
    if t.dtype != torch.float64:
        t = t.to(torch.float64)
    return t.contiguous()


def as_batch(theta: torch.Tensor) -> Tuple[torch.Tensor, bool]:
    """[2,L,L] -> ([1,2,L,L], True); [B,2,L,L] -> (same, False)."""
    if theta.dim() == 3:
        return theta.unsqueeze(0), True
    if theta.dim() == 4:
        return theta, False
    raise ValueError(f"expected theta of shape [2,L,L] or [B,2,L,L], got {tuple(theta.shape)}")


def check_field(theta: torch.Tensor) -> Tuple[int, int]:
    B, two, L, L2 = theta.shape
    if two != 2 or L != L2:
        raise ValueError(f"expected [B,2,L,L], got {tuple(theta.shape)}")
    return B, L


class Workspace:
    """Grow-only device scratch buffers keyed by (device, tag)."""

    def __init__(self):
        self._bufs: Dict[Tuple[torch.device, str], torch.Tensor] = {}

    def get(self, dev: torch.device, nbytes: int, tag: str = "main") -> Tuple[int, int]:
        key = (dev, tag)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            self._bufs[key] = buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
        return buf.data_ptr(), buf.numel()


WS = Workspace()


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


TUNE_SEED_SALT = 0x74756E65          # "tune": tuning trials never share a Philox stream with production chains


def tune_search(hmc, base: torch.Tensor, n_tune_steps: int, target_rate: float, target_tolerance: float,
                initial_step_size: float, max_attempts: int, draws=None, max_bytes: int = 1 << 30):
    """Binary search on the step size, hmc_u1.py:99-161 / hmc_u1_ft.py:211-273, for a sampler object with
    `.dt`, `.trajectories(theta, 1, pi, u, _seed=, _chain0=)`, `.seed`, `.chain_offset`, `.traj_counter`.

    As in the reference every trial trajectory starts from the same state `base` [B,2,L,L] (the returned
    state is discarded, hmc_u1.py:132), so the n_tune_steps trials of one attempt are INDEPENDENT: they run
    as extra chains of one launch (replicas of `base`, at most max_bytes of them at a time) instead of
    n_tune_steps sequential launches.  The acceptance rate is the mean over trials and chains.
    draws = (pi [N,B,2,L,L] or [N,2,L,L], u [N,B] or [N]) injects the draws of trial i = attempt *
    n_tune_steps + step (the order the reference consumes torch's generator); without it trial t of chain b
    draws from Philox(seed ^ TUNE_SEED_SALT, chain_offset * n_tune_steps + t * B + b, traj_counter + attempt).
    Returns the history [(dt tried, rate)]; the tuned value is left in hmc.dt."""
    B, _, L, _ = base.shape
    per = B * 2 * L * L * 8
    c = max(1, min(int(n_tune_steps), max_bytes // per))
    hmc.dt = initial_step_size
    step_min, step_max = 1e-6, 1.0
    best_dt, best_diff, rate = hmc.dt, float("inf"), 0.0
    history = []
    trial = 0
    for attempt in range(max_attempts):
        acc = torch.zeros((), dtype=torch.float64, device=base.device)
        t = 0
        while t < n_tune_steps:
            m = min(c, n_tune_steps - t)
            rep = base.unsqueeze(0).expand(m, B, 2, L, L).reshape(m * B, 2, L, L).contiguous()
            pi = u = None
            if draws is not None:
                pi = to_dev(draws[0][trial:trial + m], base.device).reshape(1, m * B, 2, L, L)
                u = to_dev(torch.as_tensor(draws[1][trial:trial + m]), base.device).reshape(m * B)
            # trial t of chain b is Philox chain chain_offset * n_tune_steps + t * B + b: contiguous inside one
            # launch and disjoint between ranks (chain_offset is the number of chains on lower ranks)
            r = hmc.trajectories(rep, 1, pi, u, _seed=hmc.seed ^ TUNE_SEED_SALT,
                                 _chain0=hmc.chain_offset * int(n_tune_steps) + t * B,
                                 _traj0=hmc.traj_counter + attempt, _advance=False)
            acc += r["accept"].double().sum()
            t += m
            trial += m
        rate = (acc / (n_tune_steps * B)).item()
        diff = abs(rate - target_rate)
        history.append((hmc.dt, rate))
        hmc._print(f"Step size: {hmc.dt:.6f}, Acceptance rate: {rate:.2%}")
        if diff < best_diff:
            best_dt, best_diff = hmc.dt, diff
        if diff <= target_tolerance:
            hmc._print(f"Found good step size: {hmc.dt:.6f}")
            break
        if rate > target_rate:
            step_min = hmc.dt
            hmc.dt = min((hmc.dt + step_max) / 2, step_max)
        else:
            step_max = hmc.dt
            hmc.dt = max((hmc.dt + step_min) / 2, step_min)
    if abs(rate - target_rate) > target_tolerance:
        hmc._print(f"Using best found step size: {best_dt:.6f}")
        hmc.dt = best_dt
    hmc.traj_counter += len(history)
    return history
