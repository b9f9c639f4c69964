"""Host-side cost of one FT force call at tiny batch (launch-bound regime)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hmc_ft_b200 as hb
L = 64
ft = hb.FieldTransformation(L, device="cuda", model_tag="rnet3")
ft.load_weights_npz(os.path.join(ROOT, "tests", "golden", "weights_rnet3_beta6.0.npz"))
for B in (1, 4, 16):
    th = (0.3 * torch.randn(B, 2, L, L, dtype=torch.float64)).cuda()
    ft.new_force(th, 6.0); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): F = ft.new_force(th, 6.0)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"B={B}: host issue {1e2*(t1-t0):.3f} ms/force, total {1e2*(t2-t0):.3f} ms/force, launches {ft.gpu_launches}")
    ft.gpu_launches = 0
