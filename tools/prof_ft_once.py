"""Minimal FT workload for `ncu --set full`: rnet3, L=64, B chains, `n` new_force calls and nothing else (no warm-up
loop - ncu serialises and replays the profiled launches anyway).  python tools/prof_ft_once.py [B] [n]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hmc_ft_b200 as hb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 342
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
L, beta = 64, 6.0
ft = hb.FieldTransformation(L, device="cuda", model_tag="rnet3")
ft.load_weights_npz(os.path.join(ROOT, "tests", "golden", "weights_rnet3_beta6.0.npz"))
g = torch.Generator().manual_seed(1)
th = (0.3 * torch.randn(B, 2, L, L, generator=g, dtype=torch.float64)).cuda()
for _ in range(n):
    F = ft.new_force(th, beta)
torch.cuda.synchronize()
print("done", float(F.abs().max().item()))
