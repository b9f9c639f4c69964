"""Lattice primitives with the reference's names (2d_u1_cluster/utils.py:21-222), computed by the
sm_100a library.  Inputs may be host or device tensors of any float dtype; arithmetic is fp64 on
the GPU and results come back on the input's device.  Host-side helpers (masks, analytic
expectations, autocorrelation) are plain numpy/torch index logic with no device work."""
from __future__ import annotations

import math
import random

import numpy as np
import torch

from . import _lib
from ._runtime import WS, as_batch, check_field, ptr, require_cuda, stream_ptr, to_dev


def set_seed(seed: int) -> None:
    """utils.py:11-19."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def _run_field_op(name, theta, out_shape_fn):
    lib = _lib.load()
    dev = require_cuda(theta.device if theta.is_cuda else None)
    tb, single = as_batch(theta)
    d = to_dev(tb, dev)
    B, L = check_field(d)
    out = torch.empty(out_shape_fn(B, L), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(getattr(lib, name)(d.data_ptr(), out.data_ptr(), B, L, stream_ptr(dev)), name)
    out = out[0] if single else out
    return out if theta.is_cuda else out.to(theta.device)


def plaq_from_field(theta: torch.Tensor) -> torch.Tensor:
    """utils.py:143-150.  [2,L,L] -> [L,L] (also accepts [B,2,L,L])."""
    return _run_field_op("u1_plaquette", theta, lambda B, L: (B, L, L))


def plaq_from_field_batch(theta: torch.Tensor) -> torch.Tensor:
    """utils.py:118-127.  [B,2,L,L] -> [B,L,L]."""
    return _run_field_op("u1_plaquette", theta, lambda B, L: (B, L, L))


def rect_from_field_batch(theta: torch.Tensor) -> torch.Tensor:
    """utils.py:129-141.  [B,2,L,L] -> [B,2,L,L]."""
    return _run_field_op("u1_rectangle", theta, lambda B, L: (B, 2, L, L))


def regularize(theta: torch.Tensor) -> torch.Tensor:
    """utils.py:182-187: wrap to [-pi, pi)."""
    lib = _lib.load()
    dev = require_cuda(theta.device if theta.is_cuda else None)
    d = to_dev(theta, dev)
    out = torch.empty_like(d)
    with torch.cuda.device(dev):
        _lib.check(lib.u1_regularize(d.data_ptr(), out.data_ptr(), d.numel(), stream_ptr(dev)), "u1_regularize")
    return out if theta.is_cuda else out.to(theta.device)


def _observables(theta: torch.Tensor):
    lib = _lib.load()
    dev = require_cuda(theta.device if theta.is_cuda else None)
    tb, single = as_batch(theta)
    d = to_dev(tb, dev)
    B, L = check_field(d)
    plaq = torch.empty(B, dtype=torch.float64, device=dev)
    topo = torch.empty(B, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        ws, nb = WS.get(dev, lib.u1_workspace_bytes(B, L))
        _lib.check(lib.u1_observables(d.data_ptr(), plaq.data_ptr(), topo.data_ptr(), B, L, ws, nb,
                                      stream_ptr(dev)), "u1_observables")
    if single:
        plaq, topo = plaq[0], topo[0]
    if not theta.is_cuda:
        plaq, topo = plaq.to(theta.device), topo.to(theta.device)
    return plaq, topo


def plaq_mean_from_field(theta: torch.Tensor) -> torch.Tensor:
    """utils.py:173-180."""
    return _observables(theta)[0]


def topo_from_field(theta: torch.Tensor) -> torch.Tensor:
    """utils.py:189-196."""
    return _observables(theta)[1]


# ---- host-side index logic (no arithmetic on fields) ----
def get_field_mask(index: int, batch_size: int, L: int) -> torch.Tensor:
    """utils.py:21-50."""
    m = torch.zeros((batch_size, 2, L, L), dtype=torch.bool)
    mu, k = divmod(index, 4)
    m[:, mu, (k >> 1) & 1::2, k & 1::2] = True
    return m


def get_plaq_mask(index: int, batch_size: int, L: int) -> torch.Tensor:
    """utils.py:52-80."""
    m = torch.zeros((batch_size, L, L), dtype=torch.bool)
    if index < 4:
        m[:, (1 if index < 2 else 0)::2, :] = True
    else:
        m[:, :, (1 if index in (4, 6) else 0)::2] = True
    return m


def get_rect_mask(index: int, batch_size: int, L: int) -> torch.Tensor:
    """utils.py:82-115."""
    m = torch.zeros((batch_size, 2, L, L), dtype=torch.bool)
    k = index % 4
    a, b = (k >> 1) & 1, k & 1          # parity of the active links of this subset
    if index < 4:                        # only R1 kept
        m[:, 1, (1 - a)::2, :] = True
        m[:, 1, a::2, (1 - b)::2] = True
    else:                                # only R0 kept
        m[:, 0, :, (1 - b)::2] = True
        m[:, 0, (1 - a)::2, b::2] = True
    return m


def plaq_mean_theory(beta: float) -> float:
    """utils.py:152-171: I1(beta)/I0(beta)."""
    from scipy.special import i0, i1
    return float(i1(beta) / i0(beta))


def chi_infinity(beta: float) -> float:
    """utils.py:207-222."""
    from scipy.integrate import quad
    num, _ = quad(lambda p: (p / (2 * math.pi)) ** 2 * math.exp(beta * math.cos(p)), -math.pi, math.pi)
    den, _ = quad(lambda p: math.exp(beta * math.cos(p)), -math.pi, math.pi)
    return num / den


def _autocorr(topo, max_lag: int, two_v_chi: float, which: str):
    """Device evaluation for one history [T] or a batch [T,B]; returns the input's kind (numpy in ->
    numpy out, tensor in -> tensor on the same device)."""
    lib = _lib.load()
    as_np = not torch.is_tensor(topo)
    t = torch.as_tensor(np.asarray(topo, dtype=np.float64)) if as_np else topo
    single = t.dim() == 1
    t2 = t.reshape(t.shape[0], -1)
    dev = require_cuda(t2.device if t2.is_cuda else None)
    d = to_dev(t2, dev)
    T, B = d.shape
    if max_lag >= T:
        raise ValueError("max_lag must be smaller than the history length")
    out = torch.empty((max_lag + 1, B), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        ws, nb = WS.get(dev, 16 * B + 256, "autocorr")
        args = (out.data_ptr(), None) if which == "chi" else (None, out.data_ptr())
        _lib.check(lib.u1_autocorr(d.data_ptr(), T, B, int(max_lag), float(two_v_chi), args[0], args[1], ws, nb,
                                   stream_ptr(dev)), "u1_autocorr")
    out = out[:, 0] if single else out
    if as_np:
        return out.cpu().numpy()
    return out if t.is_cuda else out.to(t.device)


def auto_from_chi(topo, max_lag: int, beta: float, volume: int):
    """utils.py:224-262: Gamma(d) = 1 - <(Q_t - Q_{t+d})^2> / (2 V chi_inf); [T] or [T,B] histories."""
    return _autocorr(topo, max_lag, 2 * volume * chi_infinity(beta), "chi")


def auto_by_def(topo, max_lag: int):
    """utils.py:264-298: normalised autocovariance; [T] or [T,B] histories."""
    return _autocorr(topo, max_lag, 1.0, "def")


def auto_from_chi_bootstrap(topo, max_lag: int, beta: float, volume: int, n_bootstrap: int = 1000, random_seed=None):
    """utils.py:300-365: bootstrap mean and standard deviation of auto_from_chi.  The resampling indices come
    from numpy's legacy generator exactly as in the reference (np.random.seed(random_seed), one
    np.random.choice(n, n) per bootstrap sample), so a seeded call reproduces the reference's numbers; the
    n_bootstrap resampled histories are gathered on the device and all their autocorrelation functions are
    evaluated by ONE u1_autocorr launch (one CTA per (lag, sample)) instead of n_bootstrap x max_lag numpy
    passes.  Returns (mean[max_lag+1], std[max_lag+1]) as numpy arrays (ddof=0, as np.std)."""
    if random_seed is not None:
        np.random.seed(random_seed)
    as_t = torch.is_tensor(topo)
    series = torch.round(topo.detach().to(torch.float64)) if as_t else torch.as_tensor(np.round(np.asarray(topo, dtype=np.float64)))
    if series.dim() != 1:
        raise ValueError("auto_from_chi_bootstrap expects one history [T]")
    n = series.shape[0]
    if max_lag >= n:
        raise ValueError("max_lag must be smaller than the history length")
    idx = np.stack([np.random.choice(n, size=n, replace=True) for _ in range(n_bootstrap)], axis=1)      # [T, n_bootstrap]
    dev = require_cuda(series.device if series.is_cuda else None)
    d = to_dev(series, dev)
    resampled = d[torch.as_tensor(idx, device=dev)].contiguous()                                          # [T, n_bootstrap]
    g = _autocorr(resampled, max_lag, 2 * volume * chi_infinity(beta), "chi")                             # [max_lag+1, n_bootstrap]
    mean = g.mean(dim=1)
    std = g.var(dim=1, unbiased=False).sqrt()
    return mean.cpu().numpy(), std.cpu().numpy()
