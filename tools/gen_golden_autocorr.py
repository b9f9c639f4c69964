"""Golden vectors for the observable post-processing (utils.py:207-298) from the real reference."""
import os, sys, types
import numpy as np
stub = types.ModuleType("matplotlib"); stub.pyplot = types.ModuleType("matplotlib.pyplot")
sys.modules["matplotlib"] = stub; sys.modules["matplotlib.pyplot"] = stub.pyplot
sys.path.insert(0, "/root/reference/2d_u1_cluster")
import utils as R
rng = np.random.default_rng(5)
# random-walk topological charge series like an HMC history, one frozen series, one short series
walk = np.cumsum(rng.integers(-1, 2, size=(600, 4)), axis=0).astype(np.float64)
walk[:, 3] = 2.0
out = {"topo": walk, "beta": 3.0, "volume": 256, "max_lag": 40,
       "chi_inf": R.chi_infinity(3.0), "chi_inf_b6": R.chi_infinity(6.0)}
out["auto_from_chi"] = np.stack([R.auto_from_chi(walk[:, b], 40, 3.0, 256) for b in range(4)], axis=1)
out["auto_by_def"] = np.stack([R.auto_by_def(walk[:, b], 40) for b in range(4)], axis=1)
bm, bs = R.auto_from_chi_bootstrap(walk[:, 0], 40, 3.0, 256, n_bootstrap=64, random_seed=11)
out["boot_mean"], out["boot_std"], out["boot_n"], out["boot_seed"] = bm, bs, 64, 11
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "autocorr.npz"), **out)
print(out["chi_inf"], out["auto_from_chi"][:3], out["auto_by_def"][:3])
