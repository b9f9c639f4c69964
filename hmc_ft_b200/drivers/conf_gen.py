"""python -m hmc_ft_b200.drivers.conf_gen — 2d_u1_cluster/conf_gen.py: thermalise, run, dump the stored
configurations as one fp32 array [n_cfg,2,L,L] (the format train.py:107 consumes)."""
import argparse
import os

import numpy as np
import torch

from .. import HMC_U1
from ..utils import set_seed


def main(argv=None):
    p = argparse.ArgumentParser(description="Parameters for Configuration Generation")
    # the reference's flags and defaults (conf_gen.py:9-18)
    p.add_argument("--lattice_size", type=int, default=8)
    p.add_argument("--beta", type=float, default=3.0)
    p.add_argument("--n_thermalization", type=int, default=200)
    p.add_argument("--store_interval", type=int, default=1)
    p.add_argument("--n_configs", type=int, default=2048)
    # documented extensions
    p.add_argument("--step_size", type=float, default=0.1, help="extension (conf_gen.py:30 hard-codes 0.1)")
    p.add_argument("--rand_seed", type=int, default=1331, help="extension: Philox seed")
    p.add_argument("--n_chains", type=int, default=None, help="extension: configurations come from B parallel chains")
    p.add_argument("--out", type=str, default=None)
    a = p.parse_args(argv)
    set_seed(a.rand_seed)
    n_therm, n_steps = a.n_thermalization, 50
    hmc = HMC_U1(a.lattice_size, a.beta, n_therm, n_steps, a.step_size, device="cuda", if_tune_step_size=True,
                 n_chains=a.n_chains, seed=a.rand_seed)
    theta, therm_plaq, therm_acc = hmc.thermalize()
    per_chain = a.n_configs if a.n_chains is None else -(-a.n_configs // a.n_chains)
    cfgs, plaq, acc, topo, _h = hmc.run(a.store_interval * per_chain, theta, a.store_interval, save_config=True)
    arr = torch.stack(cfgs).reshape(-1, 2, a.lattice_size, a.lattice_size)[: a.n_configs]
    from .common import hmc_summary
    hmc_summary(a.beta, 20, a.lattice_size ** 2, therm_plaq, plaq, topo, _h, therm_acc, acc)      # conf_gen.py:55-56
    out = a.out or f"gauges/theta_ori_L{a.lattice_size}_beta{a.beta:.1f}.npy"
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    np.save(out, arr.detach().cpu().numpy().astype(np.float32))            # conf_gen.py:61
    print(">>> Simulation completed;", arr.shape[0], "configurations ->", out)
    return out


if __name__ == "__main__":
    main()
