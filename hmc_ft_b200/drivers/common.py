"""Shared pieces of the drivers: summary print-out of utils.py:367-437 without the figure."""
from __future__ import annotations

import numpy as np
import torch

from ..utils import auto_from_chi, plaq_mean_theory


def _np(x) -> np.ndarray:
    if torch.is_tensor(x):
        return x.detach().cpu().double().numpy()
    if isinstance(x, (list, tuple)) and len(x) and torch.is_tensor(x[0]):
        return torch.stack([t.detach().cpu().double() for t in x]).numpy()
    return np.asarray(x, dtype=np.float64)


def hmc_summary(beta, max_lag, volume, therm_plaq_ls, plaq_ls, topological_charges, hamiltonians,
                therm_acceptance_rate, acceptance_rate, save_path=None):
    """utils.py:367-437: autocorrelation of Q from chi_inf, acceptance and plaquette print-out.
    The 4-panel figure is drawn only if matplotlib is installed; returns (autocorr, fig-or-None)."""
    topo = _np(topological_charges)
    if topo.ndim == 1:
        topo = topo[:, None]
    max_lag = min(max_lag, max(1, topo.shape[0] - 1))
    autocor = np.asarray(auto_from_chi(topo, max_lag, beta, volume)).mean(axis=1)      # all chains in one launch
    print(f"Thermalization acceptance rate: {float(_np(therm_acceptance_rate).mean()):.4f}")
    print(f"Acceptance rate: {float(_np(acceptance_rate).mean()):.4f}")
    plaq = _np(plaq_ls)
    print(">>> Theoretical plaquette: ", plaq_mean_theory(beta))
    print(">>> Mean plaq: ", plaq.mean())
    print(">>> Std of mean plaq: ", plaq.std() / np.sqrt(plaq.size))
    fig = None
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        fig, ax = plt.subplots(2, 2, figsize=(18, 12))
        tp = _np(therm_plaq_ls); tp = tp.reshape(tp.shape[0], -1).mean(axis=1)
        pp = plaq.reshape(plaq.shape[0], -1).mean(axis=1)
        ax[0, 0].plot(np.arange(len(tp)), tp, label="Thermalization Plaquette")
        ax[0, 0].plot(np.arange(len(pp)) + len(tp), pp, label="Plaquette")
        ax[0, 0].axhline(plaq_mean_theory(beta), color="r", linestyle="--", label="Theoretical Plaquette")
        ax[0, 0].legend()
        hm = _np(hamiltonians); ax[0, 1].plot(hm.reshape(hm.shape[0], -1)[:, 0])
        ax[1, 0].plot(topo[:, 0], marker="o", markersize=3)
        ax[1, 1].plot(range(len(autocor)), autocor, marker="o")
        if save_path:
            fig.savefig(save_path, transparent=True)
    except Exception:
        pass
    return autocor, fig
