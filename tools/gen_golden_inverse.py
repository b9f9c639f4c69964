"""Golden vectors for FieldTransformation.inverse / compute_K0_K1 from the real reference (fp64)."""
import os, sys, types
os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
import numpy as np, torch
stub = types.ModuleType("matplotlib"); stub.pyplot = types.ModuleType("matplotlib.pyplot")
sys.modules["matplotlib"] = stub; sys.modules["matplotlib.pyplot"] = stub.pyplot
sys.path.insert(0, "/root/reference/2d_u1_cluster")
from field_trans_opt import FieldTransformation
import utils as R
CL = "/root/reference/2d_u1_cluster"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
for arch, ck, L, B, seed, amp in (("rnet3", "best_model_opt_L64_train_beta6.0_rnet3.pt", 8, 3, 21, 0.8),
                                   ("simple", "best_model_L64_train_beta6.0_seed2008.pt", 16, 2, 22, 1.5)):
    ft = FieldTransformation(L, device="cpu", model_tag=arch, save_tag=arch)
    c = torch.load(os.path.join(CL, "models", ck), map_location="cpu", weights_only=False)
    for i, m in enumerate(ft.models):
        m.load_state_dict({k.replace("module.", ""): v for k, v in c[f"model_state_dict_{i}"].items()}); m.double(); m.eval()
    torch.set_default_dtype(torch.float64)
    g = torch.Generator().manual_seed(seed)
    th_new = amp * torch.randn(B, 2, L, L, generator=g, dtype=torch.float64)
    with torch.no_grad():
        th_ori = ft.forward(th_new)
        back = ft.inverse(th_ori)
        plaq = R.plaq_from_field_batch(th_new); rect = R.rect_from_field_batch(th_new)
        K0, K1 = ft.compute_K0_K1(th_new, 6, plaq, rect)
    np.savez_compressed(os.path.join(OUT, f"inverse_{arch}_L{L}.npz"), arch=np.array(arch), L=L, theta_new=th_new.numpy(),
                        theta_ori=th_ori.numpy(), inverse=back.numpy(), K0_6=K0.numpy(), K1_6=K1.numpy())
    print(arch, L, "roundtrip err", (back - th_new).abs().max().item())
