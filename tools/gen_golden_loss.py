"""Golden values of FieldTransformation.loss_fn / evaluate_step (field_trans_opt.py:400-460, 490-505) from the
real reference in fp64: theta_new = inverse(theta_ori); loss = sum_{p=2,4,6,8} ||F_FT(theta_new; train_beta)
- F_plain(theta_new; beta=1)||_p / vol^(1/p)."""
import os, sys, types
os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
import numpy as np, torch
stub = types.ModuleType("matplotlib"); stub.pyplot = types.ModuleType("matplotlib.pyplot")
sys.modules["matplotlib"] = stub; sys.modules["matplotlib.pyplot"] = stub.pyplot
sys.path.insert(0, "/root/reference/2d_u1_cluster")
from field_trans_opt import FieldTransformation
CL = "/root/reference/2d_u1_cluster"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
for arch, ck, L, B, seed, amp, beta in (("rnet3", "best_model_opt_L64_train_beta6.0_rnet3.pt", 8, 3, 31, 0.8, 6.0),
                                         ("simple", "best_model_L64_train_beta6.0_seed2008.pt", 16, 2, 32, 1.5, 3.0)):
    ft = FieldTransformation(L, device="cpu", model_tag=arch, save_tag=arch)
    c = torch.load(os.path.join(CL, "models", ck), map_location="cpu", weights_only=False)
    for i, m in enumerate(ft.models):
        m.load_state_dict({k.replace("module.", ""): v for k, v in c[f"model_state_dict_{i}"].items()}); m.double(); m.eval()
    torch.set_default_dtype(torch.float64)
    ft.train_beta = beta
    g = torch.Generator().manual_seed(seed)
    th_ori = amp * torch.randn(B, 2, L, L, generator=g, dtype=torch.float64)
    loss = ft.evaluate_step(th_ori.clone())
    th_new = ft.inverse(th_ori.clone().requires_grad_(True)).detach()
    f_ori = ft.compute_force(th_new.clone(), beta=1).detach()
    f_new = ft.compute_force(th_new.clone(), beta, transformed=True).detach()
    np.savez_compressed(os.path.join(OUT, f"loss_{arch}_L{L}.npz"), arch=np.array(arch), L=L, train_beta=beta,
                        theta_ori=th_ori.numpy(), theta_new=th_new.numpy(), force_ori=f_ori.numpy(),
                        force_new=f_new.numpy(), loss=np.float64(loss))
    print(arch, L, "loss", loss)
