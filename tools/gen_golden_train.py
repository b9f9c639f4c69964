"""Golden gradients of the training loss (field_trans_opt.py:443-480 train_step: loss_fn -> backward -> AdamW step)
from the REAL reference in fp64 on the CPU: d loss / d (CNN weights of all 8 subsets), and the parameters after ONE
optimizer step (AdamW lr=1e-3 weight_decay=1e-4, field_trans_opt.py:53-56).  The gradient flows through the unrolled
fixed-point inverse (field_trans_opt.py:245-282) and through compute_force(create_graph=True) (:400-441).

    TORCHDYNAMO_DISABLE=1 python tools/gen_golden_train.py
"""
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
        m.load_state_dict({k.replace("module.", ""): v for k, v in c[f"model_state_dict_{i}"].items()}); m.double(); m.train()
    # the optimizers were built on the fp32 parameters; .double() keeps the same Parameter objects
    torch.set_default_dtype(torch.float64)
    ft.train_beta = beta
    g = torch.Generator().manual_seed(seed)
    th_ori = amp * torch.randn(B, 2, L, L, generator=g, dtype=torch.float64)
    out = {"arch": np.array(arch), "L": L, "train_beta": beta, "theta_ori": th_ori.numpy()}
    for s, m in enumerate(ft.models):
        for k, v in m.state_dict().items():
            out[f"w{s}.{k}"] = v.detach().numpy().copy()
    loss = ft.train_step(th_ori.clone())            # loss_fn -> zero_grad -> backward -> optimizer.step
    out["loss"] = np.float64(loss)
    for s, m in enumerate(ft.models):
        for k, p in m.named_parameters():
            out[f"g{s}.{k}"] = p.grad.detach().numpy().copy()
            out[f"p{s}.{k}"] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, f"train_{arch}_L{L}.npz"), **out)
    gn = np.sqrt(sum(float((out[k] ** 2).sum()) for k in out if k.startswith("g")))
    print(arch, L, "loss", loss, "|grad|", gn)
    torch.set_default_dtype(torch.float32)
