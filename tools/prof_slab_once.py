"""Minimal slab workload for `ncu --set full` of the fixed-cost kernels (momenta + sum pi^2, drift + sums, select): one
rank, L=4096, a few eager trajectories (U1_SLAB_GRAPH=0 so that every kernel is an ordinary launch)."""
import os
import sys

import torch

os.environ.setdefault("U1_SLAB_GRAPH", "0")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hmc_ft_b200.sharding import SlabHMC  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
slab = SlabHMC(L, 6.0, 8, 0.06, transport="peer", world=1, rank=0)
own = torch.zeros(2, slab.Lx, L, dtype=torch.float64, device="cuda")
for _ in range(3):
    out = slab.trajectory(own)
torch.cuda.synchronize()
slab.check_peers()
print("done", int(out["accept"].item()))
