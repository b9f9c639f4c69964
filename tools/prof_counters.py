"""Workload for the ncu counter pass that feeds tools/ncu_constants.py: a few launches of the two plain-HMC kernels
bench.py quotes rooflines for, at the C2 shape (4096 chains, L=64, 50 steps)."""
import math, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hmc_ft_b200 as hb
from hmc_ft_b200 import _lib
L, B = 64, 4096
g = torch.Generator().manual_seed(1234)
th = ((torch.rand(B, 2, L, L, generator=g, dtype=torch.float64) * 2 - 1) * math.pi).cuda()
h = hb.HMC_U1(L, 6.0, 0, 50, 0.06, n_chains=B, if_tune_step_size=False)
h.trajectories(th, 20, want_obs=False)          # thermalise: the profiled launches see the bench's state
for _ in range(3):
    h.trajectories(th, 1)
torch.cuda.synchronize()
lib = _lib.load()
pi = torch.randn(B, 2, L, L, dtype=torch.float64, device="cuda")
to, po = torch.empty_like(th), torch.empty_like(pi)
for _ in range(3):
    _lib.check(lib.u1_leapfrog_step(th.data_ptr(), pi.data_ptr(), to.data_ptr(), po.data_ptr(), B, L, 6.0, 0.06, 0.06,
                                    torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("done")
