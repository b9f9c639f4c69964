"""Per-step vs per-trajectory cost of the resident trajectory kernel (C2 shape): t(n) = a + b*n."""
import math, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hmc_ft_b200 as hb
L, B = 64, 4096
th = ((torch.rand(B, 2, L, L, dtype=torch.float64) * 2 - 1) * math.pi).cuda()
res = {}
for n in (10, 50, 100):
    h = hb.HMC_U1(L, 6.0, 0, n, 0.06 * 50 / n, n_chains=B, if_tune_step_size=False)
    d = th.clone()
    t_end = time.time() + 1.0
    while time.time() < t_end:
        h.trajectories(d, 4); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); h.trajectories(d, 10); e1.record(); torch.cuda.synchronize()
    res[n] = e0.elapsed_time(e1) / 10
    print(f"n_steps={n}: {res[n]:.3f} ms/trajectory -> {B*L*L*n/res[n]/1e6:.1f} G site-updates/s")
b = (res[100] - res[10]) / 90; a = res[10] - 10 * b
print(f"per-step {b*1e3:.2f} us, per-trajectory fixed {a*1e3:.1f} us ({a/res[50]*100:.1f}% of the 50-step trajectory)")
