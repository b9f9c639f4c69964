"""FT-HMC at the per-GPU batch of the 8-GPU strong-scaling run (1024/8 = 128 chains), on ONE GPU: trajectories/s per GPU
for the configurations bench.py reports, so that small-batch tuning (U1FT_TILE_ROWS, U1FT_GRAPH) needs no 8-GPU box."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hmc_ft_b200 as hb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for arch, L, beta in (("rnet3", 32, 3.0), ("rnet3", 64, 6.0), ("rnet1", 64, 6.0), ("simple", 64, 6.0)):
    ft = hb.FieldTransformation(L, device="cuda", model_tag=arch)
    ft.load_weights_npz(os.path.join(ROOT, "tests", "golden", f"weights_{arch}_beta{beta:.1f}.npz"))
    h = hb.HMC_U1_FT(L, beta, 0, 50, 0.05, ft.field_transformation, ft.compute_jac_logdet, device="cuda",
                     if_tune_step_size=False, n_chains=B, seed=1331)
    g = torch.Generator().manual_seed(99)
    th = (0.3 * torch.randn(B, 2, L, L, generator=g, dtype=torch.float64)).cuda()
    for _ in range(2):
        r = h.trajectories(th, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 3
    for _ in range(n):
        r = h.trajectories(th, 1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    macs = {"simple": 1944, "rnet1": 4536, "rnet3": 9720}[arch]
    tf = B * 8 * macs * 2 * 103 * L * L / (ms * 1e-3) / 1e12
    print(f"ROWS={os.environ.get('U1FT_TILE_ROWS','auto')} GRAPH={os.environ.get('U1FT_GRAPH','1')} {arch} L={L} B={B}: "
          f"{ms:.1f} ms/traj -> {B / (ms * 1e-3):.1f} traj/s/GPU, {tf:.2f} TFLOP/s, dH sum {r['dH'].sum().item():.12f}")
