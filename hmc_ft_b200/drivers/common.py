"""Shared pieces of the drivers: summary print-out of utils.py:367-437 without the figure."""
from __future__ import annotations

import numpy as np

from ..utils import auto_from_chi, plaq_mean_theory


def hmc_summary(beta, max_lag, volume, therm_plaq_ls, plaq_ls, topological_charges, hamiltonians,
                therm_acceptance_rate, acceptance_rate, save_path=None):
    """utils.py:367-437: autocorrelation of Q from chi_inf, acceptance and plaquette print-out.
    The 4-panel figure is drawn only if matplotlib is installed; returns (autocorr, fig-or-None)."""
    topo = np.asarray(topological_charges, dtype=np.float64)
    if topo.ndim == 1:
        topo = topo[:, None]
    max_lag = min(max_lag, max(1, topo.shape[0] - 1))
    autocor = np.mean([auto_from_chi(topo[:, b], max_lag, beta, volume) for b in range(topo.shape[1])], axis=0)
    print(f"Thermalization acceptance rate: {float(np.mean(therm_acceptance_rate)):.4f}")
    print(f"Acceptance rate: {float(np.mean(acceptance_rate)):.4f}")
    plaq = np.asarray(plaq_ls, dtype=np.float64)
    print(">>> Theoretical plaquette: ", plaq_mean_theory(beta))
    print(">>> Mean plaq: ", plaq.mean())
    print(">>> Std of mean plaq: ", plaq.std() / np.sqrt(plaq.size))
    fig = None
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        fig, ax = plt.subplots(2, 2, figsize=(18, 12))
        tp = np.asarray(therm_plaq_ls, dtype=np.float64).reshape(len(therm_plaq_ls), -1).mean(axis=1)
        pp = plaq.reshape(plaq.shape[0], -1).mean(axis=1)
        ax[0, 0].plot(np.arange(len(tp)), tp, label="Thermalization Plaquette")
        ax[0, 0].plot(np.arange(len(pp)) + len(tp), pp, label="Plaquette")
        ax[0, 0].axhline(plaq_mean_theory(beta), color="r", linestyle="--", label="Theoretical Plaquette")
        ax[0, 0].legend()
        ax[0, 1].plot(np.asarray(hamiltonians, dtype=np.float64).reshape(len(hamiltonians), -1)[:, 0])
        ax[1, 0].plot(topo[:, 0], marker="o", markersize=3)
        ax[1, 1].plot(range(len(autocor)), autocor, marker="o")
        if save_path:
            fig.savefig(save_path, transparent=True)
    except Exception:
        pass
    return autocor, fig
