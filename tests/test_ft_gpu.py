"""GPU parity of the FT-HMC path against the reference-generated golden vectors and the oracle:
field transformation, Jacobian log-det, transformed action and the hand-derived force.
North-star tolerance: 1e-12 relative on action / force / log-det."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import u1_oracle as O

CASES = [
    ("rnet3_L8", "weights_rnet3_beta6.0.npz"), ("simple_L8", "weights_simple_beta6.0.npz"),
    ("rnet1_L8", "weights_rnet1_beta6.0.npz"), ("rnet3_L16_b3", "weights_rnet3_beta3.0.npz"),
    ("rnet3_L32_b3", "weights_rnet3_beta3.0.npz"), ("rnet3_L64_b6", "weights_rnet3_beta6.0.npz"),
    ("simple_L64_b6", "weights_simple_beta6.0.npz"), ("rnet1_L64_b6", "weights_rnet1_beta6.0.npz"),
]


def rel(a, b):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def make_ft(golden_dir, L, wfile):
    import hmc_ft_b200 as hb
    arch = str(np.load(os.path.join(golden_dir, wfile))["arch"])
    ft = hb.FieldTransformation(L, device="cuda", model_tag=arch)
    ft.load_weights_npz(os.path.join(golden_dir, wfile))
    return ft


@pytest.mark.parametrize("tag,wfile", CASES)
def test_golden_ft(golden_dir, tag, wfile):
    import hmc_ft_b200 as hb
    z = np.load(os.path.join(golden_dir, f"ft_{tag}.npz"))
    L, beta = int(z["L"]), float(z["beta"])
    ft = make_ft(golden_dir, L, wfile)
    th = torch.tensor(z["theta"], dtype=torch.float64, device="cuda")
    assert rel(ft.ft_phase(th[None], 0)[0], z["phase0"]) < 1e-12
    assert rel(ft.ft_phase(th[None], 5)[0], z["phase5"]) < 1e-12
    assert rel(ft.field_transformation(th), z["theta_ori"]) < 1e-12
    ld = ft.compute_jac_logdet(th[None])
    assert ld.shape == (1,) and abs(ld.item() - float(z["logdet"])) < 1e-12 * abs(float(z["logdet"]))
    hmc = hb.HMC_U1_FT(L, beta, 0, int(z["n_steps"]), float(z["dt"]), ft.field_transformation_compiled,
                       ft.compute_jac_logdet_compiled, device="cuda", if_tune_step_size=False)
    S = hmc.new_action(th)
    assert S.dim() == 0 and abs(S.item() - float(z["new_action"])) < 1e-12 * abs(float(z["new_action"]))
    F = hmc.new_force(th)
    assert rel(F, z["new_force"]) < 1e-12, rel(F, z["new_force"])
    if "logdet_dense" in z.files:
        assert abs(ld.item() - float(z["logdet_dense"])) < 1e-12 * abs(ld.item())
    th2, pi2 = hmc.leapfrog(th, torch.tensor(z["pi"], dtype=torch.float64, device="cuda"))
    assert rel(th2, z["theta_out"]) < 1e-10 and rel(pi2, z["pi_out"]) < 1e-10
    S2 = hmc.new_action(th2)
    dH = (S2 + 0.5 * (pi2 ** 2).sum()) - (S + 0.5 * (torch.tensor(z["pi"], device="cuda") ** 2).sum())
    assert abs(dH.item() - float(z["dH"])) < 1e-9


@pytest.mark.parametrize("arch,wfile,L,B", [("rnet1", "weights_rnet1_beta6.0.npz", 16, 5),
                                             ("simple", "weights_simple_beta6.0.npz", 10, 3),
                                             ("rnet3", "weights_rnet3_beta3.0.npz", 12, 4),
                                             ("simple", "weights_simple_beta6.0.npz", 96, 2),
                                             ("rnet1", "weights_rnet1_beta6.0.npz", 128, 2)])
def test_batched_vs_oracle_and_chunking(golden_dir, arch, wfile, L, B):
    """Batch of hot configurations (large angles -> all branches of wrap / masks), any even L,
    and the same batch processed in chunks of 2 chains.  L = 96: generic convolution kernels with the fused
    (direct-load) final layer; L = 128: the tiled kernels with one row group per CTA and the 64-wide reverse fusion."""
    from hmc_ft_b200 import _lib
    ft = make_ft(golden_dir, L, wfile)
    _, nets = O.load_weights_npz(os.path.join(golden_dir, wfile))
    fo = O.FieldTransformationOracle(L, nets)
    g = torch.Generator().manual_seed(11 + L)
    th = (torch.rand(B, 2, L, L, generator=g, dtype=torch.float64) * 2 - 1) * math.pi
    beta = 2.7
    ori_o, ld_o = fo.forward_and_logdet(th)
    S_o, F_o = fo.new_action(th, beta), fo.new_force(th, beta)
    d = th.cuda()
    ori, ld = ft.forward_and_logdet(d)
    assert rel(ori, ori_o) < 1e-12 and rel(ld, ld_o) < 1e-12
    assert rel(ft.new_action(d, beta), S_o) < 1e-12
    F = ft.new_force(d, beta)
    assert rel(F, F_o) < 1e-12, rel(F, F_o)
    prev = _lib.load().u1ft_set_chunk_sites(2 * L * L)
    try:
        assert torch.equal(ft.new_force(d, beta), F)
        assert torch.equal(ft.forward_and_logdet(d)[1], ld)
    finally:
        _lib.load().u1ft_set_chunk_sites(prev)
    # host-buffer call path
    import hmc_ft_b200 as hb
    fc = hb.FieldTransformation(L, device="cpu", model_tag=arch)
    fc.load_weights_npz(os.path.join(golden_dir, wfile))
    Fh = fc.new_force(th, beta)
    assert Fh.device.type == "cpu" and rel(Fh, F_o) < 1e-12


def test_force_is_gradient_of_action(golden_dir):
    """Finite-difference check of the hand-derived reverse pass, independent of any autograd."""
    L, beta = 8, 4.0
    ft = make_ft(golden_dir, L, "weights_rnet3_beta6.0.npz")
    g = torch.Generator().manual_seed(3)
    th = 0.8 * torch.randn(1, 2, L, L, generator=g, dtype=torch.float64).cuda()
    F = ft.new_force(th, beta)[0]
    eps = 1e-5
    for (mu, x, y) in [(0, 0, 0), (1, 3, 5), (0, 7, 2), (1, 4, 4)]:
        tp, tm = th.clone(), th.clone()
        tp[0, mu, x, y] += eps
        tm[0, mu, x, y] -= eps
        fd = (ft.new_action(tp, beta) - ft.new_action(tm, beta)).item() / (2 * eps)
        assert abs(fd - F[mu, x, y].item()) < 1e-6 * max(1.0, abs(fd))


def test_ft_trajectory_vs_oracle(golden_dir):
    import hmc_ft_b200 as hb
    L, B, beta, n, dt = 8, 3, 3.0, 4, 0.07
    wfile = "weights_rnet3_beta3.0.npz"
    ft = make_ft(golden_dir, L, wfile)
    _, nets = O.load_weights_npz(os.path.join(golden_dir, wfile))
    fo = O.FieldTransformationOracle(L, nets)
    g = torch.Generator().manual_seed(21)
    th = 0.7 * torch.randn(B, 2, L, L, generator=g, dtype=torch.float64)
    pi = torch.randn(B, 2, L, L, generator=g, dtype=torch.float64)
    u = torch.tensor([0.999, 1e-9, 0.5], dtype=torch.float64)
    o = O.FTHMCOracle(fo, beta, n, dt)
    th_o, acc_o, H_o, dH_o = o.metropolis_step(th, pi, u)
    h = hb.HMC_U1_FT(L, beta, 0, n, dt, ft.field_transformation, ft.compute_jac_logdet, device="cuda",
                     if_tune_step_size=False, n_chains=B)
    d = th.cuda().clone()
    r = h.trajectories(d, 1, pi.cuda().reshape(1, B, 2, L, L).contiguous(), u.cuda(), want_ori=True)
    assert rel(r["dH"][0], dH_o) < 1e-9
    assert torch.equal(r["accept"][0].cpu().bool(), acc_o)
    assert rel(r["H"][0], H_o) < 1e-11
    assert rel(d, th_o) < 1e-10
    ori_o = fo.forward(th_o)
    assert rel(r["theta_ori"], ori_o) < 1e-10
    assert rel(r["plaq"][0], O.plaq_mean(ori_o)) < 1e-10
    assert torch.equal(r["topo"][0].cpu(), O.topo_charge(ori_o))
    # reference-style single-chain call
    th1, ok, H = h.metropolis_step(th[0].cuda(), pi=pi[0], u=u[0])
    assert isinstance(ok, bool) and ok == bool(acc_o[0]) and abs(H - H_o[0].item()) < 1e-9 * abs(H)


def test_identity_ft_equals_plain_chain():
    """debug/check_hmc.py:75-78: identity transformation + zero log-det reproduces plain HMC."""
    import hmc_ft_b200 as hb
    L, beta, n, dt = 16, 6.0, 10, 0.1
    hf = hb.HMC_U1_FT(L, beta, 0, n, dt, lambda x: x, lambda x: torch.zeros(1), device="cuda", if_tune_step_size=False)
    hp = hb.HMC_U1(L, beta, 0, n, dt, device="cuda", if_tune_step_size=False)
    a = torch.zeros(1, 2, L, L, dtype=torch.float64, device="cuda")
    b = a.clone()
    ra, rb = hf.trajectories(a, 4), hp.trajectories(b, 4)
    assert torch.equal(a, b) and torch.equal(ra["dH"], rb["dH"])
    with pytest.raises(TypeError):
        hb.HMC_U1_FT(L, beta, 0, n, dt, lambda x: 2 * x, lambda x: torch.zeros(1), device="cuda")


def test_reference_style_ft_run(golden_dir):
    import hmc_ft_b200 as hb
    L = 8
    ft = make_ft(golden_dir, L, "weights_simple_beta6.0.npz")
    h = hb.HMC_U1_FT(L, 3.0, 3, 5, 0.1, ft.field_transformation_compiled, ft.compute_jac_logdet_compiled,
                     device="cpu", if_tune_step_size=False)
    theta, plaq_ls, acc = h.thermalize()
    assert theta.shape == (2, L, L) and len(plaq_ls) == 3 and abs(plaq_ls[0] - 1.0) < 1e-12
    cfg, plaq, rate, topo, ham = h.run(4, theta, store_interval=2, save_config=True)
    assert len(cfg) == 2 and cfg[0].shape == (2, L, L) and cfg[0].device.type == "cpu"
    assert len(plaq) == 2 and len(topo) == 2 and len(ham) == 2 and 0 <= rate <= 1
    assert float(cfg[0].abs().max()) <= math.pi + 1e-12


def test_ft_ensemble_plaquette(golden_dir):
    """Ensemble <plaq> of FT-HMC agrees with I1/I0 within statistical error (L=16, beta=3, rnet3
    trained at beta=3, 256 chains) — the exactness of the log-det + force pair shows up here."""
    import hmc_ft_b200 as hb
    from hmc_ft_b200.utils import plaq_mean_theory
    L, beta, n, dt, B = 16, 3.0, 10, 0.1, 256
    ft = make_ft(golden_dir, L, "weights_rnet3_beta3.0.npz")
    h = hb.HMC_U1_FT(L, beta, 0, n, dt, ft.field_transformation, ft.compute_jac_logdet, device="cuda",
                     if_tune_step_size=False, n_chains=B)
    d = torch.zeros(B, 2, L, L, dtype=torch.float64, device="cuda")
    h.trajectories(d, 60)
    r = h.trajectories(d, 40)
    plaq = r["plaq"].mean(dim=0)
    mean, err = plaq.mean().item(), plaq.std().item() / math.sqrt(B)
    assert abs(mean - plaq_mean_theory(beta)) < 5 * err + 2e-4, (mean, err, plaq_mean_theory(beta))
    assert 0.5 < r["accept"].double().mean().item() <= 1.0
    assert abs(torch.exp(-r["dH"]).mean().item() - 1) < 0.05


def test_ft_handle_errors():
    from hmc_ft_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    h = C.c_void_p()
    w = torch.zeros(10, dtype=torch.float64)
    assert lib.u1ft_create(C.byref(h), 3, w.data_ptr(), 10) == -4
    assert lib.u1ft_weight_count(0) == 8 * 1968 and lib.u1ft_weight_count(1) == 8 * 4584 and lib.u1ft_weight_count(3) == 8 * 9816
    assert lib.u1ft_forward(None, None, None, None, 1, 8, None, 0, None) == -4


def test_drivers_smoke(golden_dir, tmp_path):
    """Thin CLI callers (compare_hmc.py / compare_fthmc.py / conf_gen.py flag surface)."""
    from hmc_ft_b200.drivers import compare_fthmc, compare_hmc, conf_gen
    plaq, acc, topo = compare_hmc.main(["--lattice_size", "16", "--n_configs", "8", "--beta", "2.0", "--step_size", "0.1",
                                        "--max_lag", "4", "--no_tune", "--n_chains", "4"])
    assert plaq.shape == (8, 4) and topo.shape == (8, 4)
    plaq, acc, topo = compare_fthmc.main(["--lattice_size", "8", "--n_configs", "3", "--beta", "3.0", "--ft_step_size", "0.1",
                                          "--max_lag", "2", "--model_tag", "simple", "--n_therm", "2",
                                          "--weights", os.path.join(golden_dir, "weights_simple_beta6.0.npz")])
    assert len(plaq) == 3
    out = conf_gen.main(["--lattice_size", "8", "--n_configs", "6", "--beta", "2.0", "--n_chains", "3",
                         "--out", str(tmp_path / "cfg.npy")])
    a = np.load(out)
    assert a.shape == (6, 2, 8, 8) and a.dtype == np.float32 and np.abs(a).max() <= math.pi + 1e-6


def test_ft_small_lattice_and_nan(golden_dir):
    """L=4 (3x3 stencils wrap heavily) against the oracle; NaN input -> NaN dH -> reject."""
    import hmc_ft_b200 as hb
    L, B, beta = 4, 3, 2.0
    wfile = "weights_rnet1_beta6.0.npz"
    ft = make_ft(golden_dir, L, wfile)
    _, nets = O.load_weights_npz(os.path.join(golden_dir, wfile))
    fo = O.FieldTransformationOracle(L, nets)
    g = torch.Generator().manual_seed(44)
    th = (torch.rand(B, 2, L, L, generator=g, dtype=torch.float64) * 2 - 1) * math.pi
    assert rel(ft.new_force(th.cuda(), beta), fo.new_force(th, beta)) < 1e-12
    assert rel(ft.compute_jac_logdet(th.cuda()), fo.compute_jac_logdet(th)) < 1e-12
    h = hb.HMC_U1_FT(L, beta, 0, 3, 0.1, ft.field_transformation, ft.compute_jac_logdet, n_chains=B, if_tune_step_size=False)
    d = th.cuda().clone()
    d[2, 1, 0, 0] = float("nan")
    keep = d.clone()
    r = h.trajectories(d, 1)
    assert r["accept"][0, 2].item() == 0 and torch.isnan(r["dH"][0, 2])
    assert torch.equal(torch.nan_to_num(d[2], nan=5.0), torch.nan_to_num(keep[2], nan=5.0))
    assert not torch.isnan(r["dH"][0, 0])


def test_fthmc_matches_plain_hmc_ensemble(golden_dir):
    """Matched runs: FT-HMC and plain HMC sample the same distribution — <plaq> and <Q^2>/V agree
    within statistical error (L=16, beta=3, 512 chains each)."""
    import hmc_ft_b200 as hb
    L, beta, B = 16, 3.0, 512
    ft = make_ft(golden_dir, L, "weights_rnet3_beta3.0.npz")
    hf = hb.HMC_U1_FT(L, beta, 0, 10, 0.1, ft.field_transformation, ft.compute_jac_logdet, n_chains=B,
                      if_tune_step_size=False, seed=11)
    hp = hb.HMC_U1(L, beta, 0, 10, 0.1, n_chains=B, if_tune_step_size=False, seed=12)
    df = torch.zeros(B, 2, L, L, dtype=torch.float64, device="cuda")
    dp = df.clone()
    hf.trajectories(df, 100); hp.trajectories(dp, 300, want_obs=False)
    rf, rp = hf.trajectories(df, 40), hp.trajectories(dp, 40)

    def stats(x):
        c = x.mean(dim=0)
        return c.mean().item(), c.std().item() / math.sqrt(B)
    (pf, ef), (pp, ep) = stats(rf["plaq"]), stats(rp["plaq"])
    assert abs(pf - pp) < 5 * math.hypot(ef, ep) + 1e-4, (pf, ef, pp, ep)
    (qf, sf), (qp, sp) = stats(rf["topo"] ** 2), stats(rp["topo"] ** 2)
    assert abs(qf - qp) < 5 * math.hypot(sf, sp) + 0.05, (qf, sf, qp, sp)


@pytest.mark.parametrize("tag,wfile", [("rnet3_L8", "weights_rnet3_beta6.0.npz"), ("simple_L16", "weights_simple_beta6.0.npz")])
def test_inverse_and_coefficients(golden_dir, tag, wfile):
    """FieldTransformation.inverse (field_trans_opt.py:245-282) and compute_K0_K1 (:110-135) against the
    reference's outputs; iteration counts against the oracle (same global stopping rule)."""
    z = np.load(os.path.join(golden_dir, f"inverse_{tag}.npz"))
    L = int(z["L"])
    ft = make_ft(golden_dir, L, wfile)
    back = ft.inverse(torch.tensor(z["theta_ori"], device="cuda"))
    assert rel(back, z["inverse"]) < 1e-12
    _, nets = O.load_weights_npz(os.path.join(golden_dir, wfile))
    _, iters = O.FieldTransformationOracle(L, nets).inverse(torch.tensor(z["theta_ori"]))
    assert ft.inverse_iterations == iters
    assert rel(ft.inverse_field_transformation(torch.tensor(z["theta_ori"][0])), z["inverse"][0]) < 1e-4   # single config: its own (global-norm) stopping point
    K0, K1 = ft.compute_K0_K1(torch.tensor(z["theta_new"], device="cuda"), 6)
    m = np.zeros_like(z["K1_6"], dtype=bool); m[:, 4:8, 1::2, 0::2] = True      # subset 6: mu=1, parity (1,0)
    assert rel(K1.cpu().numpy()[m], z["K1_6"][m]) < 1e-12
    m0 = np.zeros_like(z["K0_6"], dtype=bool); m0[:, 2:4, 1::2, 0::2] = True
    assert rel(K0.cpu().numpy()[m0], z["K0_6"][m0]) < 1e-12 and float(K0[:, :2].abs().max()) == 0.0


@pytest.mark.parametrize("tag,wfile", [("rnet3_L8", "weights_rnet3_beta6.0.npz"), ("simple_L16", "weights_simple_beta6.0.npz")])
def test_training_loss_value(golden_dir, tag, wfile):
    """loss_fn / evaluate_step / compute_force (field_trans_opt.py:400-505), forward values, against the real
    reference's numbers (tools/gen_golden_loss.py) and the oracle."""
    z = np.load(os.path.join(golden_dir, f"loss_{tag}.npz"))
    L, beta = int(z["L"]), float(z["train_beta"])
    ft = make_ft(golden_dir, L, wfile)
    th_new = torch.tensor(z["theta_new"], device="cuda")
    assert rel(ft.compute_force(th_new, 1.0), z["force_ori"]) < 1e-12
    assert rel(ft.compute_force(th_new, beta, transformed=True), z["force_new"]) < 1e-12
    ft.train_beta = beta
    loss = ft.evaluate_step(torch.tensor(z["theta_ori"], device="cuda"))
    assert abs(loss - float(z["loss"])) < 1e-9 * abs(float(z["loss"]))
    _, nets = O.load_weights_npz(os.path.join(golden_dir, wfile))
    want = O.FieldTransformationOracle(L, nets).loss_fn(torch.tensor(z["theta_ori"]), beta).item()
    assert abs(ft.loss_fn(torch.tensor(z["theta_ori"]), train_beta=beta).item() - want) < 1e-9 * abs(want)
    # train_step itself: tests/test_train_gpu.py
