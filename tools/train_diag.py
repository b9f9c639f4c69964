"""Component-wise check of the device training gradient against the CPU oracle's autograd (runs on the GPU box):
force self-check, Hessian-vector product, d(v.F_FT)/dW, the inverse adjoint and the total gradient vs the golden file."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hmc_ft_b200 as hb
from hmc_ft_b200 import _lib
from hmc_ft_b200._runtime import WS, stream_ptr
from oracle import u1_oracle as O

tag = sys.argv[1] if len(sys.argv) > 1 else "rnet3_L8"
z = np.load(os.path.join(ROOT, "tests", "golden", f"train_{tag}.npz"))
arch, L, beta = str(z["arch"]), int(z["L"]), float(z["train_beta"])
keys = O.STATE_KEYS[arch]
ft = hb.FieldTransformation(L, device="cuda", model_tag=arch)
for s, m in enumerate(ft.models):
    m.double()
    m.load_state_dict({f"{k}.{wb}": torch.tensor(z[f"w{s}.{k}.{wb}"]) for k in keys for wb in ("weight", "bias")})
ft.sync_weights()
sds = [{f"{k}.{wb}": z[f"w{s}.{k}.{wb}"] for k in keys for wb in ("weight", "bias")} for s in range(8)]
fo = O.FieldTransformationOracle(L, O.weights_from_state_dicts(arch, sds))
th_o = torch.tensor(z["theta_ori"])
B = th_o.shape[0]

def flat(grads):
    return torch.cat([torch.cat([g[0].reshape(-1), g[1].reshape(-1)]) for sub in grads for g in sub])

# ---- oracle pieces ----
params = []
for net in fo.nets:
    for i, (w, b) in enumerate(net.layers):
        w = w.detach().clone().requires_grad_(True); b = b.detach().clone().requires_grad_(True)
        net.layers[i] = (w, b); params += [w, b]
thn_o, iters_o = fo.inverse(th_o)
t = thn_o.detach().clone().requires_grad_(True)
F = torch.autograd.grad(fo.new_action(t, beta).sum(), t, create_graph=True)[0]
Fp = torch.autograd.grad(O.action(t, 1.0).sum(), t, create_graph=True)[0]
d = (F - Fp)
vol = float(L * L)
loss_o = sum(torch.norm(d, p=p) / vol ** (1.0 / p) for p in (2, 4, 6, 8))
v_o = torch.autograd.grad(loss_o, d, retain_graph=True)[0].detach()
phi = (F * v_o).sum()
g = torch.autograd.grad(phi, [t] + params, retain_graph=True, allow_unused=True)
Hv_o, G1_o = g[0], torch.cat([x.reshape(-1) for x in g[1:]])
Hp_o = torch.autograd.grad((Fp * v_o).sum(), t, retain_graph=True)[0]

# ---- device ----
loss, gradW = ft.loss_and_grad(th_o, beta, check=True)
print("loss", loss.item(), "oracle", loss_o.item(), "golden", float(z["loss"]))
print("inverse iterations", ft.inverse_iterations, "oracle", iters_o)
print("force self-check (d phi/dv vs new_force) rel", ft.last_force_check)
dev = ft._dev
lib = ft._lib
thn = ft.inverse(th_o.cuda())
print("theta_new vs oracle", (thn.cpu() - thn_o.detach()).abs().max().item())
f_new = ft.new_force(thn, beta); f_ori = ft.compute_force(thn, 1.0)
sums = torch.empty(4, dtype=torch.float64, device=dev); v = torch.empty_like(f_new)
sp = stream_ptr(dev)
ws, nb = WS.get(dev, lib.u1_diff_power_sums_workspace_bytes(), "loss")
_lib.check(lib.u1_diff_power_sums(f_new.data_ptr(), f_ori.data_ptr(), f_new.numel(), sums.data_ptr(), ws, nb, sp))
_lib.check(lib.u1_loss_grad(f_new.data_ptr(), f_ori.data_ptr(), sums.data_ptr(), vol, f_new.numel(), v.data_ptr(), sp))
print("v vs oracle rel", ((v.cpu() - v_o).abs().max() / v_o.abs().max()).item())
n_w = int(lib.u1ft_weight_count(ft.models[0].n_res_blocks))
g1 = torch.zeros(n_w, dtype=torch.float64, device=dev); Hv = torch.empty_like(f_new); fc = torch.empty_like(f_new)
tws, tnb = WS.get(dev, lib.u1ft_train_workspace_bytes(ft._handle, B, L), "fttrain")
vd = v_o.cuda().contiguous()
_lib.check(lib.u1ft_train_force_grad(ft._handle, thn.data_ptr(), vd.data_ptr(), B, L, beta, None, Hv.data_ptr(), fc.data_ptr(),
                                     g1.data_ptr(), n_w, tws, tnb, sp))
print("Hv_FT vs oracle rel", ((Hv.cpu() - Hv_o).abs().max() / Hv_o.abs().max()).item())
print("G1 = d(v.F)/dW vs oracle rel", ((g1.cpu() - G1_o).abs().max() / G1_o.abs().max()).item(), "|G1|max", G1_o.abs().max().item())
Hp = torch.empty_like(f_new)
_lib.check(lib.u1_force_hvp(thn.data_ptr(), vd.data_ptr(), Hp.data_ptr(), B, L, 1.0, sp))
print("Hv_plain vs oracle rel", ((Hp.cpu() - Hp_o).abs().max() / Hp_o.abs().max()).item())
# per-subset breakdown of G1
per = n_w // 8
for s in range(8):
    a, b = g1.cpu()[s * per:(s + 1) * per], G1_o[s * per:(s + 1) * per]
    print(f"  subset {s}: G1 rel err {((a - b).abs().max() / max(b.abs().max().item(), 1e-300)).item():.3e}")
# total gradient vs golden
gold = torch.cat([torch.cat([torch.tensor(z[f"g{s}.{k}.weight"]).reshape(-1), torch.tensor(z[f"g{s}.{k}.bias"]).reshape(-1)])
                  for s in range(8) for k in keys])
err = (gradW.cpu() - gold).abs().max().item() / gold.abs().max().item()
print("TOTAL grad vs reference golden: rel-to-max err", err, "|g|max", gold.abs().max().item())
G2_o = gold - G1_o
G2 = gradW.cpu() - g1.cpu()
print("G2 (inverse part) rel", ((G2 - G2_o).abs().max() / max(G2_o.abs().max().item(), 1e-300)).item(), "|G2|max", G2_o.abs().max().item())
for s in range(8):
    a, b = G2[s * per:(s + 1) * per], G2_o[s * per:(s + 1) * per]
    print(f"  subset {s}: G2 rel err {((a - b).abs().max() / max(b.abs().max().item(), 1e-300)).item():.3e}  |G2_s|max {b.abs().max().item():.3e}")
