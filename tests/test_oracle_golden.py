"""Pin the oracle (oracle/u1_oracle.py) against the real reference's outputs
(tests/golden/*.npz, made by tools/gen_golden.py), the reference's own golden log
tests/test.out, and the dense-Jacobian self check (field_trans_opt.py:385-390)."""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import u1_oracle as O

T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("tag", ["L6_b1p5", "L16_b2", "L64_b6"])
def test_plain_matches_reference(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, f"plain_{tag}.npz"))
    th, pi = T(z["theta"]), T(z["pi"])
    beta, n, dt = float(z["beta"]), int(z["n_steps"]), float(z["dt"])
    assert rel(O.plaq_from_field(th), z["plaq_field"]) < 1e-15
    assert rel(O.regularize(th * 3.7), z["regularized"]) == 0.0
    assert abs(O.action(th, beta).item() - float(z["action"])) <= 1e-13 * abs(float(z["action"]))
    assert rel(O.force_autograd(th, beta), z["force"]) < 1e-15
    assert rel(O.force_analytic(th, beta), z["force"]) < 1e-14
    h = O.PlainHMCOracle(int(z["L"]), beta, n, dt)
    th2, pi2 = h.leapfrog(th, pi)
    assert rel(th2, z["theta_out"]) < 1e-13 and rel(pi2, z["pi_out"]) < 1e-13
    assert abs(O.plaq_mean(th).item() - float(z["plaq"])) < 1e-15
    assert O.topo_charge(th).item() == float(z["topo"])
    _, _, _, dH = h.metropolis_step(th[None], pi[None], torch.tensor([0.5], dtype=torch.float64))
    assert abs(dH.item() - float(z["dH"])) < 1e-10 * max(1.0, abs(float(z["dH"])))


@pytest.mark.parametrize("tag,wfile", [
    ("rnet3_L8", "weights_rnet3_beta6.0.npz"), ("simple_L8", "weights_simple_beta6.0.npz"),
    ("rnet1_L8", "weights_rnet1_beta6.0.npz"), ("rnet3_L16_b3", "weights_rnet3_beta3.0.npz"),
    ("rnet3_L32_b3", "weights_rnet3_beta3.0.npz"), ("rnet3_L64_b6", "weights_rnet3_beta6.0.npz"),
    ("simple_L64_b6", "weights_simple_beta6.0.npz"), ("rnet1_L64_b6", "weights_rnet1_beta6.0.npz"),
])
def test_ft_matches_reference(golden_dir, tag, wfile):
    z = np.load(os.path.join(golden_dir, f"ft_{tag}.npz"))
    arch, nets = O.load_weights_npz(os.path.join(golden_dir, wfile))
    assert arch == str(z["arch"])
    L, beta = int(z["L"]), float(z["beta"])
    ft = O.FieldTransformationOracle(L, nets)
    th = T(z["theta"])[None]
    assert rel(ft.ft_phase(th, 0)[0], z["phase0"]) < 1e-13
    assert rel(ft.ft_phase(th, 5)[0], z["phase5"]) < 1e-13
    X, _, _ = ft.features(th, 3)
    K0, K1 = ft.nets[3](X)
    assert rel(K0[0], z["K0_3"]) < 1e-13 and rel(K1[0], z["K1_3"]) < 1e-13
    ori, ld = ft.forward_and_logdet(th)
    assert rel(ori[0], z["theta_ori"]) < 1e-13
    assert abs(ld.item() - float(z["logdet"])) < 1e-13 * abs(float(z["logdet"]))
    assert rel(ft.forward(th)[0], z["theta_ori"]) < 1e-13
    S = ft.new_action(th, beta)
    assert abs(S.item() - float(z["new_action"])) < 1e-13 * abs(float(z["new_action"]))
    assert rel(ft.new_force(th, beta)[0], z["new_force"]) < 1e-12
    if "logdet_dense" in z.files:       # if_check_jac path, field_trans_opt.py:385-390
        assert abs(ld.item() - float(z["logdet_dense"])) < 1e-12 * abs(ld.item())
        assert abs(ft.logdet_dense_autograd(th[0]).item() - ld.item()) < 1e-12 * abs(ld.item())
    if L <= 16:                         # short FT leapfrog (5 steps) incl. dH
        h = O.FTHMCOracle(ft, beta, int(z["n_steps"]), float(z["dt"]))
        th2, pi2 = h.leapfrog(th, T(z["pi"])[None])
        assert rel(th2[0], z["theta_out"]) < 1e-11 and rel(pi2[0], z["pi_out"]) < 1e-11


def test_ft_phase_loop_spec(golden_dir):
    """Explicit index loops against the roll-based restatement."""
    _, nets = O.load_weights_npz(os.path.join(golden_dir, "weights_rnet1_beta6.0.npz"))
    L = 6
    ft = O.FieldTransformationOracle(L, nets)
    g = torch.Generator().manual_seed(7)
    th = torch.randn(1, 2, L, L, generator=g, dtype=torch.float64)
    for k in range(8):
        X, _, _ = ft.features(th, k)
        K0, K1 = ft.nets[k](X)
        d = O.ft_phase_loops(th[0].numpy(), K0[0].numpy(), K1[0].numpy(), k)
        assert np.max(np.abs(d - ft.ft_phase(th, k)[0].numpy())) < 5e-16
        # triangular Jacobian premise: K of subset k does not depend on subset-k links
        th2 = th.clone(); th2[0][ft.fm[k]] += 0.37
        X2, _, _ = ft.features(th2, k)
        assert torch.equal(X, X2) or float((X - X2).abs().max()) < 1e-15


def test_masks_partition():
    L = 8
    tot = sum(O.field_mask(k, L).to(torch.int32) for k in range(8))
    assert torch.all(tot == 1)


def test_replay_reference_golden_log(golden_dir):
    """tests/test.out (tests/hmc_2DU1.py): seed 1984, randn_like then rand([]) per trajectory."""
    with open(os.path.join(golden_dir, "test_out_head.json")) as f:
        G = json.load(f)
    L, beta, n = G["L"], G["beta"], G["nstep"]
    dt = G["tau"] / n
    torch.manual_seed(G["seed"])
    x = torch.zeros(2, L, L, dtype=torch.float64)
    h = O.PlainHMCOracle(L, beta, n, dt)
    for row in G["rows"][:10]:
        p = torch.randn_like(x)
        u = torch.rand([], dtype=torch.float64)
        x1, acc, _, dH = h.metropolis_step(x[None], p[None], u[None])
        x = x1[0]
        assert bool(acc.item()) == row["accept"]
        assert abs(dH.item() - row["dH"]) <= 2e-7 * max(1.0, abs(row["dH"])), (row, dH.item())
        assert abs(O.plaq_mean(x).item() - row["plaq"]) < 5e-8
        assert O.topo_charge(x).item() == row["topo"]


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for c, k, want in kat:
        got = O.philox4x32_10(np.array([c], dtype=np.uint32), np.array([k], dtype=np.uint32))[0]
        assert [int(v) for v in got] == want
    p = O.philox_momenta(1331, 5, 17, 16)
    assert p.shape == (2, 16, 16) and abs(p.mean()) < 0.2 and 0.8 < p.std() < 1.2
    assert 0.0 < O.philox_accept_uniform(1331, 5, 17) < 1.0


def test_autocorrelation_matches_reference(golden_dir):
    """utils.py:207-298 (chi_infinity, auto_from_chi, auto_by_def) against reference outputs."""
    z = np.load(os.path.join(golden_dir, "autocorr.npz"))
    assert abs(O.chi_infinity(3.0) - float(z["chi_inf"])) < 1e-15 and abs(O.chi_infinity(6.0) - float(z["chi_inf_b6"])) < 1e-15
    for b in range(z["topo"].shape[1]):
        a = O.auto_from_chi(z["topo"][:, b], int(z["max_lag"]), float(z["beta"]), int(z["volume"]))
        d = O.auto_by_def(z["topo"][:, b], int(z["max_lag"]))
        assert np.array_equal(a, z["auto_from_chi"][:, b]) and np.array_equal(d, z["auto_by_def"][:, b])
    # utils.py:300-365 bootstrap, seeded legacy numpy generator
    m, sd = O.auto_from_chi_bootstrap(z["topo"][:, 0], int(z["max_lag"]), float(z["beta"]), int(z["volume"]),
                                      n_bootstrap=int(z["boot_n"]), random_seed=int(z["boot_seed"]))
    assert np.array_equal(m, z["boot_mean"]) and np.array_equal(sd, z["boot_std"])


@pytest.mark.parametrize("tag,wfile", [("rnet3_L8", "weights_rnet3_beta6.0.npz"), ("simple_L16", "weights_simple_beta6.0.npz")])
def test_inverse_matches_reference(golden_dir, tag, wfile):
    """field_trans_opt.py:245-282 inverse (same stopping rule) against the reference's output."""
    z = np.load(os.path.join(golden_dir, f"inverse_{tag}.npz"))
    _, nets = O.load_weights_npz(os.path.join(golden_dir, wfile))
    ft = O.FieldTransformationOracle(int(z["L"]), nets)
    back, iters = ft.inverse(T(z["theta_ori"]))
    assert rel(back, z["inverse"]) < 1e-13 and max(iters) < 200
    assert rel(back, z["theta_new"]) < 1e-4                      # debug/check_ft.py:113-118 threshold


@pytest.mark.parametrize("tag,wfile", [("rnet3_L8", "weights_rnet3_beta6.0.npz"), ("simple_L16", "weights_simple_beta6.0.npz")])
def test_training_loss_value_matches_reference(golden_dir, tag, wfile):
    """field_trans_opt.py:400-505 loss_fn / evaluate_step (forward value) against the reference's numbers."""
    z = np.load(os.path.join(golden_dir, f"loss_{tag}.npz"))
    _, nets = O.load_weights_npz(os.path.join(golden_dir, wfile))
    ft = O.FieldTransformationOracle(int(z["L"]), nets)
    th_new = T(z["theta_new"])
    assert rel(ft.compute_force(th_new, 1.0), z["force_ori"]) < 1e-13
    assert rel(ft.compute_force(th_new, float(z["train_beta"]), transformed=True), z["force_new"]) < 1e-12
    loss = ft.loss_fn(T(z["theta_ori"]), float(z["train_beta"]))
    assert abs(loss.item() - float(z["loss"])) < 1e-11 * abs(float(z["loss"]))


@pytest.mark.parametrize("L", [6, 8, 16])
def test_rectangles_match_reference(golden_dir, L):
    """utils.py:118-141 rect_from_field_batch / plaq_from_field_batch, reference outputs (tools/gen_golden_rect_tune.py)."""
    z = np.load(os.path.join(golden_dir, "rect.npz"))
    th = T(z[f"theta_L{L}"])
    assert rel(O.rect_from_field(th), z[f"rect_L{L}"]) == 0.0
    assert rel(O.plaq_from_field(th), z[f"plaq_L{L}"]) == 0.0
    # R0(x,y) = P(x,y) + P(x+1,y), R1(x,y) = P(x,y) + P(x,y+1): the identity the FT reverse pass relies on
    P = O.plaq_from_field(th)
    r = O.rect_from_field(th)
    assert rel(P + torch.roll(P, -1, -2), r[:, 0]) < 1e-14 and rel(P + torch.roll(P, -1, -1), r[:, 1]) < 1e-14


def test_tune_step_size_matches_reference(golden_dir):
    """hmc_u1.py:99-161: the oracle's binary search replays the REAL reference's search (history of (dt, rate)
    and the tuned dt) when fed the draws the reference consumed from torch's generator."""
    z = np.load(os.path.join(golden_dir, "tune.npz"))
    L, beta, n_steps, n_tune, target, tol, dt0, max_att = z["params"]
    h = O.PlainHMCOracle(int(L), float(beta), int(n_steps), float(dt0))
    hist = O.tune_step_size(h, T(z["theta"])[None], T(z["pi"])[:, None], T(z["u"])[:, None], int(n_tune), float(target),
                            float(tol), float(dt0), int(max_att))
    assert len(hist) == len(z["hist_dt"])
    for (dt, rate), dt_ref, rate_ref in zip(hist, z["hist_dt"], z["hist_rate"]):
        assert abs(dt - dt_ref) < 1e-6                       # the reference prints 6 decimals
        assert round(rate * int(n_tune)) == round(float(rate_ref) * int(n_tune))
    assert h.dt == float(z["tuned_dt"])


@pytest.mark.parametrize("tag", ["rnet3_L8", "simple_L16"])
def test_training_gradient_matches_reference(golden_dir, tag):
    """field_trans_opt.py:443-480: loss and d loss / d weights of the oracle's autograd restatement against the REAL
    reference's train_step (tools/gen_golden_train.py): loss to 1e-12, every gradient tensor to 1e-9 of the largest
    gradient entry (double backward through ~10 unrolled inverse iterations per subset)."""
    z = np.load(os.path.join(golden_dir, f"train_{tag}.npz"))
    arch, L = str(z["arch"]), int(z["L"])
    keys = O.STATE_KEYS[arch]
    sds = [{f"{k}.{wb}": z[f"w{s}.{k}.{wb}"] for k in keys for wb in ("weight", "bias")} for s in range(8)]
    fo = O.FieldTransformationOracle(L, O.weights_from_state_dicts(arch, sds))
    loss, grads = fo.train_loss_and_grads(T(z["theta_ori"]), float(z["train_beta"]))
    assert abs(loss.item() - float(z["loss"])) < 1e-12 * abs(float(z["loss"]))
    gmax = max(np.abs(z[f"g{s}.{k}.weight"]).max() for s in range(8) for k in keys)
    for s in range(8):
        for li, k in enumerate(keys):
            assert np.abs(grads[s][li][0].numpy() - z[f"g{s}.{k}.weight"]).max() < 1e-9 * gmax
            assert np.abs(grads[s][li][1].numpy() - z[f"g{s}.{k}.bias"]).max() < 1e-9 * gmax
