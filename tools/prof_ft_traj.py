"""FT-HMC trajectory timing for a given number of chains (per GPU), L=64 rnet3, n_steps leapfrog steps."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hmc_ft_b200 as hb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
L, beta = 64, 6.0
ft = hb.FieldTransformation(L, device="cuda", model_tag="rnet3")
ft.load_weights_npz(os.path.join(ROOT, "tests", "golden", "weights_rnet3_beta6.0.npz"))
h = hb.HMC_U1_FT(L, beta, 0, n_steps, 0.05, ft.field_transformation, ft.compute_jac_logdet, n_chains=B, if_tune_step_size=False)
th = (0.3 * torch.randn(B, 2, L, L, dtype=torch.float64)).cuda()
h.trajectories(th, 1); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); r = h.trajectories(th, 2); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
fl = 8 * 9720 * 2 * (2 * n_steps + 3) * L * L * B
print(f"B={B} n_steps={n_steps}: {ms:.1f} ms/trajectory -> {fl / ms / 1e9:.2f} TFLOP/s algorithmic, acc={r['accept'].double().mean().item():.2f}")
