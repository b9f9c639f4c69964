"""python -m hmc_ft_b200.drivers.compare_hmc — flags of 2d_u1_cluster/compare_hmc.py:13-21."""
import argparse
import time

from .. import HMC_U1
from ..utils import set_seed
from .common import hmc_summary


def main(argv=None):
    p = argparse.ArgumentParser(description="Parameters for Comparison")
    p.add_argument("--lattice_size", type=int, default=16)
    p.add_argument("--n_configs", type=int, default=2048)
    p.add_argument("--beta", type=float, default=6)
    p.add_argument("--step_size", type=float, default=0.1)
    p.add_argument("--max_lag", type=int, default=200)
    p.add_argument("--rand_seed", type=int, default=1331)
    p.add_argument("--device", type=str, default="cuda")
    p.add_argument("--n_chains", type=int, default=None, help="extension: batch of independent chains")
    p.add_argument("--no_tune", action="store_true", help="extension: skip tune_step_size")
    a = p.parse_args(argv)
    print("=" * 60 + "\n>>> Arguments:")
    for k, v in vars(a).items():
        print(f"{k}: {v}")
    print("=" * 60)
    set_seed(a.rand_seed)
    n_therm, n_steps, store_interval = 200, 50, 1          # compare_hmc.py:42-45
    print(">>> No Field Transformation HMC Simulation: ")
    hmc = HMC_U1(a.lattice_size, a.beta, n_therm, n_steps, a.step_size, device=a.device,
                 if_tune_step_size=not a.no_tune, n_chains=a.n_chains, seed=a.rand_seed)
    t0 = time.time()
    theta, therm_plaq, therm_acc = hmc.thermalize()
    t1 = time.time()
    print(f">>> Thermalization completed in {t1 - t0:.2f} seconds")
    _, plaq, acc, topo, ham = hmc.run(store_interval * a.n_configs, theta, store_interval, save_config=False)
    t2 = time.time()
    print(f">>> Simulation completed in {t2 - t1:.2f} seconds")
    print(f">>> Total time (Standard HMC): {t2 - t0:.2f} seconds")
    hmc_summary(a.beta, a.max_lag, a.lattice_size ** 2, therm_plaq, plaq, topo, ham, therm_acc, acc)
    return plaq, acc, topo


if __name__ == "__main__":
    main()
