import math, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hmc_ft_b200 as hb
L, B = 64, 4096
h = hb.HMC_U1(L, 6.0, 0, 50, 0.06, device="cpu", if_tune_step_size=False, n_chains=B)
th = ((torch.rand(B, 2, L, L, dtype=torch.float64) * 2 - 1) * math.pi).pin_memory()
out = torch.empty_like(th).pin_memory()
print("pinned:", th.is_pinned(), out.is_pinned(), th[0:512].is_pinned())
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
for ns in (1, 2, 4, 8, 16, 32):
    print("n_slices", ns, f"{timed(lambda: h.step_host(th, out, n_slices=ns)):.2f} ms")
t0 = time.perf_counter(); h.step_host(th, out, n_slices=8); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host-side issue time {1e3*(t1-t0):.2f} ms, total {1e3*(t2-t0):.2f} ms")
