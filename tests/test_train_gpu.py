"""Training gradient: d loss / d (CNN weights) computed by the hand-derived device path
(csrc/u1_ft_train.cuh) against the REAL reference's loss.backward() (tests/golden/train_*.npz, tools/gen_golden_train.py),
and one train_step (AdamW) against the reference's updated parameters."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import u1_oracle as O


def _load(golden_dir, tag):
    import hmc_ft_b200 as hb
    z = np.load(os.path.join(golden_dir, f"train_{tag}.npz"))
    arch, L = str(z["arch"]), int(z["L"])
    keys = O.STATE_KEYS[arch]
    ft = hb.FieldTransformation(L, device="cuda", model_tag=arch)
    for s, m in enumerate(ft.models):
        m.double()
        m.load_state_dict({f"{k}.{wb}": torch.tensor(z[f"w{s}.{k}.{wb}"]) for k in keys for wb in ("weight", "bias")})
    ft.sync_weights()
    ft.train_beta = float(z["train_beta"])
    return ft, z, keys


@pytest.mark.parametrize("tag", ["rnet3_L8", "simple_L16"])
def test_loss_gradient_matches_reference_backward(golden_dir, tag):
    ft, z, keys = _load(golden_dir, tag)
    loss, gradW = ft.loss_and_grad(torch.tensor(z["theta_ori"]), check=True)
    assert abs(loss.item() - float(z["loss"])) < 1e-9 * abs(float(z["loss"]))
    assert ft.last_force_check < 1e-11                       # d(v.F_FT)/dv reproduces the FT force
    gold = torch.cat([torch.cat([torch.tensor(z[f"g{s}.{k}.weight"]).reshape(-1), torch.tensor(z[f"g{s}.{k}.bias"]).reshape(-1)])
                      for s in range(8) for k in keys])
    err = (gradW.cpu() - gold).abs().max().item() / gold.abs().max().item()
    assert err < 1e-9, err                                   # relative to the largest gradient entry


@pytest.mark.parametrize("tag", ["rnet3_L8", "simple_L16"])
def test_train_step_matches_reference_adamw(golden_dir, tag):
    """field_trans_opt.py:463-480: loss -> backward -> AdamW(lr=1e-3, wd=1e-4) step; parameters after the step."""
    ft, z, keys = _load(golden_dir, tag)
    loss = ft.train_step(torch.tensor(z["theta_ori"]))
    assert abs(loss - float(z["loss"])) < 1e-9 * abs(float(z["loss"]))
    for s, m in enumerate(ft.models):
        sd = m.state_dict()
        for k in keys:
            for wb in ("weight", "bias"):
                assert (sd[f"{k}.{wb}"].double() - torch.tensor(z[f"p{s}.{k}.{wb}"])).abs().max().item() < 1e-9
    # the new weights are live on the device: the loss changes
    loss2 = ft.evaluate_step(torch.tensor(z["theta_ori"]))
    assert loss2 != loss and np.isfinite(loss2)


def test_gradient_finite_difference(golden_dir):
    """Independent of any reference: central finite difference of the device loss along a random weight direction."""
    ft, z, keys = _load(golden_dir, "simple_L16")
    th = torch.tensor(z["theta_ori"])
    loss, gradW = ft.loss_and_grad(th)
    g = torch.Generator().manual_seed(0)
    params = [p for m in ft.models for c in m.conv_layers() for p in (c.weight, c.bias)]
    dirs = [torch.randn(p.shape, generator=g, dtype=torch.float64) for p in params]
    flat_dir = torch.cat([d.reshape(-1) for d in dirs])
    eps = 1e-5
    vals = []
    for sgn in (+1, -1):
        with torch.no_grad():
            for p, d in zip(params, dirs):
                p.add_(sgn * eps * d)
        ft.sync_weights()
        vals.append(ft.loss_fn(th).item())
        with torch.no_grad():
            for p, d in zip(params, dirs):
                p.sub_(sgn * eps * d)
    ft.sync_weights()
    fd = (vals[0] - vals[1]) / (2 * eps)
    an = float((gradW.cpu() * flat_dir).sum().item())
    assert abs(fd - an) < 2e-5 * max(1.0, abs(an)), (fd, an)


def test_train_loop_epochs_checkpoint_and_reload(golden_dir, tmp_path, monkeypatch):
    """field_trans_opt.py:507-584: train() runs epochs of train_step / evaluate_step, keeps the best checkpoint in the
    reference's file format (models/best_model_opt_L{L}_train_beta{beta}.pt, per-subset state dicts) and re-loads it at the
    end; the re-loaded weights reproduce the best test loss."""
    ft, z, keys = _load(golden_dir, "simple_L16")
    monkeypatch.chdir(tmp_path)                      # the checkpoint path is CWD-relative like the reference's
    g = torch.Generator().manual_seed(3)
    base = torch.tensor(z["theta_ori"])
    base = base[None] if base.dim() == 3 else base
    data = torch.cat([base[:1] + 0.05 * torch.randn(6, *base.shape[1:], generator=g, dtype=torch.float64)])
    train_data, test_data = data[:4], data[4:]
    beta = float(z["train_beta"])
    tr, te = ft.train(train_data, test_data, beta, n_epochs=3, batch_size=2)
    assert len(tr) == 3 and len(te) == 3 and all(np.isfinite(tr)) and all(np.isfinite(te))
    path = tmp_path / "models" / f"best_model_opt_L16_train_beta{beta:.1f}.pt"
    assert path.exists()
    ck = torch.load(path, weights_only=False)
    assert {f"model_state_dict_{i}" for i in range(8)} <= set(ck) and ck["loss"] == min(te)
    ev = float(np.mean([ft.evaluate_step(test_data[i:i + 2]) for i in range(0, 2, 2)]))
    assert abs(ev - min(te)) < 1e-9 * abs(min(te))
