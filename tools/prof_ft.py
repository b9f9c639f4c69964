"""Small FT workload for ncu: one chunk (B chains, L=64, rnet3) through new_force / trajectory."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hmc_ft_b200 as hb  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "rnet3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 74
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
L, beta = 64, 6.0
ft = hb.FieldTransformation(L, device="cuda", model_tag=arch)
ft.load_weights_npz(os.path.join(ROOT, "tests", "golden", f"weights_{arch}_beta6.0.npz"))
g = torch.Generator().manual_seed(1)
th = (0.3 * torch.randn(B, 2, L, L, generator=g, dtype=torch.float64)).cuda()
F = ft.new_force(th, beta)
torch.cuda.synchronize()
t_end = time.time() + 1.5            # clocks ramp up from idle: warm up for 1.5 s before timing
while time.time() < t_end:
    F = ft.new_force(th, beta)
    torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    F = ft.new_force(th, beta)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
macs = {"simple": 1944, "rnet1": 4536, "rnet3": 9720}[arch]
fl = 8 * macs * 2 * 2 * L * L * B
print(f"{arch} B={B}: new_force {ms:.3f} ms  -> {fl / ms / 1e9:.2f} TFLOP/s algorithmic (conv MACs fwd+bwd)")
