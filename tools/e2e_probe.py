import math, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hmc_ft_b200 as hb
L, B = 64, 4096
h = hb.HMC_U1(L, 6.0, 0, 50, 0.06, device="cpu", if_tune_step_size=False, n_chains=B)
th = ((torch.rand(B, 2, L, L, dtype=torch.float64) * 2 - 1) * math.pi).pin_memory()
out = torch.empty_like(th).pin_memory()
print("pinned:", th.is_pinned(), out.is_pinned(), th[0:512].is_pinned())
def timed(fn, reps=10):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn(); torch.cuda.current_stream().synchronize()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
for rep in range(2):
    for ns in (8, 16, "tapered", 32, 64):
        print("n_slices", ns, f"{timed(lambda: h.step_host(th, out, n_slices=ns)):.2f} ms")
