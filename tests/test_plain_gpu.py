"""GPU parity of the plain HMC path (C ABI through hmc_ft_b200) against the CPU oracle and the
reference-generated golden vectors.  Tolerances: 1e-12 relative on action / force (north star);
leapfrog/dH after n steps are compared at 1e-9 (round-off of a 50-step symplectic map)."""
import json
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import u1_oracle as O


def rel(a, b):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def hb():
    import hmc_ft_b200 as hb
    assert torch.cuda.is_available()
    return hb


@pytest.mark.parametrize("tag", ["L6_b1p5", "L16_b2", "L64_b6"])
def test_golden_action_force_leapfrog(hb, golden_dir, tag):
    from hmc_ft_b200 import utils as U
    z = np.load(os.path.join(golden_dir, f"plain_{tag}.npz"))
    L, beta, n, dt = int(z["L"]), float(z["beta"]), int(z["n_steps"]), float(z["dt"])
    th = torch.tensor(z["theta"], dtype=torch.float64, device="cuda")
    pi = torch.tensor(z["pi"], dtype=torch.float64, device="cuda")
    h = hb.HMC_U1(L, beta, 0, n, dt, device="cuda", if_tune_step_size=False)
    S = h.action(th)
    assert S.dim() == 0 and abs(S.item() - float(z["action"])) <= 1e-12 * abs(float(z["action"]))
    assert rel(h.force(th), z["force"]) < 1e-12
    assert rel(U.plaq_from_field(th), z["plaq_field"]) < 1e-15
    assert rel(U.regularize(th * 3.7), z["regularized"]) < 1e-15
    assert abs(U.plaq_mean_from_field(th).item() - float(z["plaq"])) < 1e-14
    assert U.topo_from_field(th).item() == float(z["topo"])
    th2, pi2 = h.leapfrog(th, pi)
    assert rel(th2, z["theta_out"]) < 1e-9 and rel(pi2, z["pi_out"]) < 1e-9
    # host-buffer call path (device="cpu": H2D / D2H inside the call)
    hc = hb.HMC_U1(L, beta, 0, n, dt, device="cpu", if_tune_step_size=False)
    Fh = hc.force(torch.tensor(z["theta"]))
    assert Fh.device.type == "cpu" and rel(Fh, z["force"]) < 1e-12


@pytest.mark.parametrize("L,B,path", [(8, 5, 2), (16, 3, 2), (32, 4, 2), (64, 3, 2), (64, 3, 1),
                                      (16, 2, 1), (6, 3, 1), (12, 2, 1), (130, 1, 1), (128, 2, 1), (100, 1, 1),
                                      (232, 1, 1)])
def test_trajectory_vs_oracle(hb, L, B, path):
    """Fused trajectory (both kernel paths) with injected pi/u against the oracle:
    dH, accept, H, plaq, topo, final state."""
    g = torch.Generator().manual_seed(100 + L)
    beta, n, dt = 2.5, 7, 0.11
    th = (torch.rand(B, 2, L, L, generator=g, dtype=torch.float64) * 2 - 1) * math.pi
    pi = torch.randn(B, 2, L, L, generator=g, dtype=torch.float64)
    u = torch.rand(B, generator=g, dtype=torch.float64)
    o = O.PlainHMCOracle(L, beta, n, dt)
    th_o, acc_o, H_o, dH_o = o.metropolis_step(th, pi, u)
    h = hb.HMC_U1(L, beta, 0, n, dt, device="cuda", if_tune_step_size=False, n_chains=B, kernel_path=path)
    d = th.cuda().clone()
    r = h.trajectories(d, 1, pi.cuda().reshape(1, B, 2, L, L).contiguous(), u.cuda())
    assert rel(r["dH"][0], dH_o) < 1e-10
    assert torch.equal(r["accept"][0].cpu().bool(), acc_o)
    assert rel(r["H"][0], H_o) < 1e-11
    assert rel(d, th_o) < 1e-10
    assert rel(r["plaq"][0], O.plaq_mean(th_o)) < 1e-11
    assert torch.equal(r["topo"][0].cpu(), O.topo_charge(th_o))
    assert h.gpu_launches > 0


def test_replay_reference_golden_log(hb, golden_dir):
    """tests/test.out: every trajectory restarted from the oracle's state (no error build-up),
    momenta/uniforms from torch's CPU generator seeded like tests/hmc_2DU1.py."""
    with open(os.path.join(golden_dir, "test_out_head.json")) as f:
        G = json.load(f)
    L, beta, n = G["L"], G["beta"], G["nstep"]
    dt = G["tau"] / n
    torch.manual_seed(G["seed"])
    x = torch.zeros(2, L, L, dtype=torch.float64)
    o = O.PlainHMCOracle(L, beta, n, dt)
    for path in (2, 1):
        torch.manual_seed(G["seed"])
        x = torch.zeros(2, L, L, dtype=torch.float64)
        h = hb.HMC_U1(L, beta, 0, n, dt, device="cuda", if_tune_step_size=False, kernel_path=path)
        for row in G["rows"][:8]:
            p = torch.randn_like(x)
            u = torch.rand([], dtype=torch.float64)
            xo, acc, _, dH = o.metropolis_step(x[None], p[None], u[None])
            d = x.cuda()[None].clone()
            r = h.trajectories(d, 1, p.cuda().reshape(1, 1, 2, L, L), u.cuda().reshape(1))
            assert abs(r["dH"][0, 0].item() - row["dH"]) <= 2e-7 * max(1.0, abs(row["dH"]))
            assert bool(r["accept"][0, 0].item()) == row["accept"]
            assert abs(r["plaq"][0, 0].item() - row["plaq"]) < 5e-8
            assert r["topo"][0, 0].item() == row["topo"]
            assert abs(r["dH"][0, 0].item() - dH.item()) < 1e-9
            x = xo[0]


def test_philox_stream_bit_exact(hb):
    """Device Philox momenta / uniforms against the numpy restatement (oracle) and N(0,1) moments."""
    from hmc_ft_b200 import _lib
    lib = _lib.load()
    B, L, seed, chain0, traj = 3, 16, 1331, 7, 5
    pi = torch.empty(B, 2, L, L, dtype=torch.float64, device="cuda")
    _lib.check(lib.u1_momenta(pi.data_ptr(), seed, chain0, traj, B, L, torch.cuda.current_stream().cuda_stream))
    for b in range(B):
        want = O.philox_momenta(seed, chain0 + b, traj, L)
        assert np.max(np.abs(pi[b].cpu().numpy() - want)) < 1e-13
    big = torch.empty(64, 2, 64, 64, dtype=torch.float64, device="cuda")
    _lib.check(lib.u1_momenta(big.data_ptr(), 1, 0, 0, 64, 64, torch.cuda.current_stream().cuda_stream))
    assert abs(big.mean().item()) < 5e-3 and abs(big.std().item() - 1) < 5e-3
    assert abs((big ** 4).mean().item() - 3) < 0.05


def test_chain_results_independent_of_batching(hb):
    """Philox keyed by global chain id: chains 4..7 run alone == chains 4..7 inside a batch of 8."""
    L, beta, n, dt = 32, 3.0, 10, 0.1
    a = hb.HMC_U1(L, beta, 0, n, dt, n_chains=8, if_tune_step_size=False)
    b = hb.HMC_U1(L, beta, 0, n, dt, n_chains=4, if_tune_step_size=False, chain_offset=4)
    ta = torch.zeros(8, 2, L, L, dtype=torch.float64, device="cuda")
    tb = torch.zeros(4, 2, L, L, dtype=torch.float64, device="cuda")
    ra = a.trajectories(ta, 5)
    rb = b.trajectories(tb, 5)
    assert torch.equal(ta[4:], tb) and torch.equal(ra["dH"][:, 4:], rb["dH"])
    # streaming path produces the same chain up to round-off
    c = hb.HMC_U1(L, beta, 0, n, dt, n_chains=4, if_tune_step_size=False, chain_offset=4, kernel_path=1)
    tc = torch.zeros(4, 2, L, L, dtype=torch.float64, device="cuda")
    rc = c.trajectories(tc, 1)
    assert rel(rc["dH"], rb["dH"][:1]) < 1e-9


def test_ensemble_plaquette_matches_theory(hb):
    """<plaq> = I1(beta)/I0(beta) within statistical error (utils.py:152-171), B=512 chains at full
    C2 lattice size; also energy conservation <exp(-dH)> = 1."""
    from hmc_ft_b200.utils import plaq_mean_theory
    L, beta, n, dt, B = 64, 6.0, 50, 0.06, 512
    h = hb.HMC_U1(L, beta, 0, n, dt, n_chains=B, if_tune_step_size=False)
    d = torch.zeros(B, 2, L, L, dtype=torch.float64, device="cuda")
    h.trajectories(d, 600, want_obs=False)            # cold start relaxes slowly (cf. tests/test.out)
    r = h.trajectories(d, 100)
    plaq = r["plaq"].mean(dim=0)                      # per-chain average over 100 trajectories
    mean, err = plaq.mean().item(), plaq.std().item() / math.sqrt(B)
    assert abs(mean - plaq_mean_theory(beta)) < 5 * err + 1e-5, (mean, err, plaq_mean_theory(beta))
    acc = r["accept"].double().mean().item()
    assert 0.4 < acc < 0.9
    e = torch.exp(-r["dH"]).mean().item()
    assert abs(e - 1) < 0.05
    assert torch.all(r["topo"] == torch.round(r["topo"]))


def test_reference_style_run_and_thermalize(hb):
    h = hb.HMC_U1(16, 2.0, 5, 10, 0.1, device="cpu", if_tune_step_size=False)
    theta, plaq_ls, acc = h.thermalize()
    assert theta.shape == (2, 16, 16) and theta.device.type == "cpu" and len(plaq_ls) == 5 and plaq_ls[0] == 1.0
    cfg, plaq, rate, topo, ham = h.run(6, theta, store_interval=2, save_config=True)
    assert len(cfg) == 3 and len(plaq) == 3 and len(topo) == 3 and len(ham) == 3 and 0 <= rate <= 1
    th, ok, H = h.metropolis_step(theta)
    assert isinstance(ok, bool) and isinstance(H, float) and th.shape == theta.shape


def test_argument_errors(hb):
    from hmc_ft_b200 import _lib
    lib = _lib.load()
    assert lib.u1_force(None, None, 1, 8, 1.0, None) == -1
    t = torch.zeros(1, 2, 7, 7, dtype=torch.float64, device="cuda")
    assert lib.u1_trajectory(t.data_ptr(), None, None, 0, 0, 0, 1, 1, 7, 1.0, 0.1, 2,
                             None, None, None, None, None, 0, None, 0, None) == -2
    t = torch.zeros(1, 2, 12, 12, dtype=torch.float64, device="cuda")
    assert lib.u1_trajectory(t.data_ptr(), None, None, 0, 0, 0, 1, 1, 12, 1.0, 0.1, 2,
                             None, None, None, None, None, 2, None, 0, None) == -5
    assert lib.u1_trajectory(t.data_ptr(), None, None, 0, 0, 0, 1, 1, 12, 1.0, 0.1, 2,
                             None, None, None, None, None, 1, None, 0, None) == -3


def test_step_host_pipeline_matches_resident(hb):
    """Host-buffer step (sliced H2D / kernel / D2H over streams) == resident path, bit for bit."""
    L, B, beta, n, dt = 32, 37, 3.0, 10, 0.1
    g = torch.Generator().manual_seed(9)
    th = (torch.rand(B, 2, L, L, generator=g, dtype=torch.float64) * 2 - 1) * math.pi
    h1 = hb.HMC_U1(L, beta, 0, n, dt, device="cuda", if_tune_step_size=False, n_chains=B, seed=5, chain_offset=3)
    d = th.cuda().clone()
    r = h1.trajectories(d, 1)
    h2 = hb.HMC_U1(L, beta, 0, n, dt, device="cpu", if_tune_step_size=False, n_chains=B, seed=5, chain_offset=3)
    out, obs = h2.step_host(th.clone(), n_slices=5)
    torch.cuda.synchronize()
    assert out.device.type == "cpu" and torch.equal(out, d.cpu())
    assert torch.equal(obs[0], r["accept"][0].double().cpu()) and torch.equal(obs[1], r["H"][0].cpu())
    assert torch.equal(obs[2], r["plaq"][0].cpu()) and torch.equal(obs[3], r["topo"][0].cpu())
    out2, _ = h2.step_host(out, n_slices=1)       # second step continues the same Philox stream
    r2 = h1.trajectories(d, 1)
    torch.cuda.synchronize()
    assert torch.equal(out2, d.cpu())
    out3, _ = h2.step_host(out2, n_slices="tapered")       # small slices at both ends, same chain
    r3 = h1.trajectories(d, 1)
    torch.cuda.synchronize()
    assert torch.equal(out3, d.cpu())


def test_topological_susceptibility_matches_theory(hb):
    """chi_t = <Q^2>/V against chi_infinity(beta) (utils.py:207-222) at a coupling where topology
    moves freely (L=16, beta=2): 1024 chains, matched to the reference's observable definitions."""
    from hmc_ft_b200.utils import chi_infinity, plaq_mean_theory
    L, beta, n, dt, B = 16, 2.0, 10, 0.1, 1024
    h = hb.HMC_U1(L, beta, 0, n, dt, n_chains=B, if_tune_step_size=False, seed=2024)
    d = torch.zeros(B, 2, L, L, dtype=torch.float64, device="cuda")
    h.trajectories(d, 300, want_obs=False)
    r = h.trajectories(d, 50)
    q2 = (r["topo"][::10] ** 2).mean(dim=0) / (L * L)          # 5 decorrelated samples per chain
    mean, err = q2.mean().item(), q2.std().item() / math.sqrt(B)
    assert abs(mean - chi_infinity(beta)) < 5 * err + 0.02 * chi_infinity(beta), (mean, err, chi_infinity(beta))
    assert abs(r["topo"].mean().item()) < 0.2
    plaq = r["plaq"].mean(dim=0)
    assert abs(plaq.mean().item() - plaq_mean_theory(beta)) < 5 * plaq.std().item() / math.sqrt(B) + 1e-4


def test_nan_configuration_is_rejected(hb):
    """NaN dH compares False against the uniform draw -> reject, state unchanged (hmc_u1.py:94)."""
    L, B = 16, 3
    for path in (1, 2):
        h = hb.HMC_U1(L, 2.0, 0, 5, 0.1, n_chains=B, if_tune_step_size=False, kernel_path=path)
        d = torch.zeros(B, 2, L, L, dtype=torch.float64, device="cuda")
        d[1, 0, 3, 3] = float("nan")
        before = d.clone()
        r = h.trajectories(d, 1)
        assert r["accept"][0, 1].item() == 0 and torch.isnan(r["dH"][0, 1])
        assert torch.equal(torch.nan_to_num(d[1], nan=7.0), torch.nan_to_num(before[1], nan=7.0))
        assert r["accept"][0, 0].item() in (0, 1) and not torch.isnan(r["dH"][0, 0])


def test_smallest_lattices(hb):
    """L=2 and L=4 (every stencil wraps onto itself) through the streaming path against the oracle."""
    for L in (2, 4):
        g = torch.Generator().manual_seed(L)
        B, beta, n, dt = 4, 1.3, 3, 0.2
        th = (torch.rand(B, 2, L, L, generator=g, dtype=torch.float64) * 2 - 1) * math.pi
        pi = torch.randn(B, 2, L, L, generator=g, dtype=torch.float64)
        u = torch.rand(B, generator=g, dtype=torch.float64)
        o = O.PlainHMCOracle(L, beta, n, dt)
        th_o, acc_o, H_o, dH_o = o.metropolis_step(th, pi, u)
        h = hb.HMC_U1(L, beta, 0, n, dt, n_chains=B, if_tune_step_size=False, kernel_path=1)
        d = th.cuda().clone()
        r = h.trajectories(d, 1, pi.cuda().reshape(1, B, 2, L, L).contiguous(), u.cuda())
        assert rel(r["dH"][0], dH_o) < 1e-10 and rel(d, th_o) < 1e-10
        assert rel(h.force(th.cuda()), O.force_autograd(th, beta)) < 1e-12


def test_autocorrelation_device(hb, golden_dir):
    """u1_autocorr against the reference's auto_from_chi / auto_by_def outputs (batched over chains)."""
    from hmc_ft_b200 import utils as U
    z = np.load(os.path.join(golden_dir, "autocorr.npz"))
    topo, ml, beta, vol = z["topo"], int(z["max_lag"]), float(z["beta"]), int(z["volume"])
    assert abs(U.chi_infinity(beta) - float(z["chi_inf"])) < 1e-15
    g = U.auto_from_chi(torch.tensor(topo, device="cuda"), ml, beta, vol)
    assert g.shape == (ml + 1, topo.shape[1]) and rel(g, z["auto_from_chi"]) < 1e-14
    d = U.auto_by_def(torch.tensor(topo, device="cuda"), ml)
    assert rel(d, z["auto_by_def"]) < 1e-12 and torch.all(d[:, 3] == 1.0)        # frozen chain -> all ones
    one = U.auto_from_chi(topo[:, 0], ml, beta, vol)                              # numpy in -> numpy out
    assert isinstance(one, np.ndarray) and one.shape == (ml + 1,) and rel(one, z["auto_from_chi"][:, 0]) < 1e-14
    # bootstrap (utils.py:300-365): same seeded resampling, all samples in one launch
    m, sd = U.auto_from_chi_bootstrap(topo[:, 0], ml, beta, vol, n_bootstrap=int(z["boot_n"]), random_seed=int(z["boot_seed"]))
    assert m.shape == (ml + 1,) and rel(m, z["boot_mean"]) < 1e-12 and rel(sd, z["boot_std"]) < 1e-10


@pytest.mark.parametrize("Lx,L,B", [(72, 100, 2), (264, 1000, 3)])
def test_multistep_kernel_bit_identical_to_single_steps(hb, Lx, L, B):
    """u1_leapfrog_steps_rect (temporally blocked, groups of 4 steps) == repeated u1_leapfrog_step,
    bit for bit, on rectangular periodic lattices whose sides are not multiples of the inner block; the
    second case has more windows (630) than resident CTAs, so the persistent prefetching kernel walks
    several windows per CTA."""
    import ctypes as C
    from hmc_ft_b200 import _lib
    lib = _lib.load()
    n, beta, dt = 7, 3.0, 0.07
    g = torch.Generator().manual_seed(77)
    th = ((torch.rand(B, 2, Lx, L, generator=g, dtype=torch.float64) * 2 - 1) * 3).cuda()
    pi = torch.randn(B, 2, Lx, L, generator=g, dtype=torch.float64).cuda()
    sp = torch.cuda.current_stream().cuda_stream
    a_t, a_p, b_t, b_p = th.clone(), pi.clone(), torch.empty_like(th), torch.empty_like(pi)
    flag = C.c_int(0)
    _lib.check(lib.u1_leapfrog_steps_rect(a_t.data_ptr(), a_p.data_ptr(), b_t.data_ptr(), b_p.data_ptr(), B, Lx, L,
                                          beta, 0.5 * dt, dt, n, C.byref(flag), sp))
    res_t, res_p = (b_t, b_p) if flag.value else (a_t, a_p)
    c_t, c_p, d_t, d_p = th.clone(), pi.clone(), torch.empty_like(th), torch.empty_like(pi)
    for s in range(n):
        _lib.check(lib.u1_leapfrog_step_rect(c_t.data_ptr(), c_p.data_ptr(), d_t.data_ptr(), d_p.data_ptr(), B, Lx, L,
                                             beta, 0.5 * dt if s == 0 else dt, dt, sp))
        c_t, d_t, c_p, d_p = d_t, c_t, d_p, c_p
    torch.cuda.synchronize()
    assert torch.equal(res_t, c_t) and torch.equal(res_p, c_p)


def test_slab_building_blocks_match_their_unfused_forms(hb):
    """The fused entry points of the slab trajectory against the separate calls they replace, on a halo-extended
    buffer (Lx own rows + k halo rows per side): u1_leapfrog_steps_rect3 (read-only input) == u1_leapfrog_steps_rect
    bit for bit; u1_slab_drift_sums_ext == u1_slab_drift_wrap then u1_slab_sums_ext (theta' bit-identical, sums to
    reduction round-off); u1_momenta_rows_pi2_dev == u1_momenta_rows_dev plus a sum over the own rows;
    u1_slab_select_state copies state, caller copy and the carried sums on accept and nothing on reject."""
    import ctypes as C
    from hmc_ft_b200 import _lib
    lib = _lib.load()
    Lx, k, L, beta, dt = 120, 8, 512, 4.0, 0.07
    Lxe = Lx + 2 * k
    plane = Lxe * L
    g = torch.Generator().manual_seed(5)
    th = ((torch.rand(2, Lxe, L, generator=g, dtype=torch.float64) * 2 - 1) * 3).cuda()
    pi = torch.randn(2, Lxe, L, generator=g, dtype=torch.float64).cuda()
    sp = torch.cuda.current_stream().cuda_stream
    # -- steps from a read-only input
    a_t, a_p, b_t, b_p = th.clone(), pi.clone(), torch.empty_like(th), torch.empty_like(pi)
    f1 = C.c_int(0)
    _lib.check(lib.u1_leapfrog_steps_rect(a_t.data_ptr(), a_p.data_ptr(), b_t.data_ptr(), b_p.data_ptr(), 1, Lxe, L, beta,
                                          0.5 * dt, dt, 6, C.byref(f1), sp))
    want_t, want_p = (b_t, b_p) if f1.value else (a_t, a_p)
    c_t, c_p, d_t, d_p = (torch.empty_like(th) for _ in range(4))
    th0, pi0 = th.clone(), pi.clone()
    f2 = C.c_int(0)
    _lib.check(lib.u1_leapfrog_steps_rect3(th0.data_ptr(), pi0.data_ptr(), c_t.data_ptr(), c_p.data_ptr(), d_t.data_ptr(),
                                           d_p.data_ptr(), 1, Lxe, L, beta, 0.5 * dt, dt, 6, C.byref(f2), sp))
    got_t, got_p = (d_t, d_p) if f2.value else (c_t, c_p)
    torch.cuda.synchronize()
    assert torch.equal(got_t, want_t) and torch.equal(got_p, want_p)
    assert torch.equal(th0, th) and torch.equal(pi0, pi)                       # input untouched
    # -- final drift fused into the sums
    nred = int(lib.u1_slab_reduce_workspace_bytes(Lxe, L))
    ws_d = torch.zeros(nred, dtype=torch.uint8, device="cuda")
    ws = torch.empty(4096 + 24 * (Lx + 2), dtype=torch.uint8, device="cuda")
    out_t = torch.zeros_like(th)
    sums_f = torch.zeros(3, dtype=torch.float64, device="cuda")
    for _ in range(2):                                                         # twice: the ticket word must reset itself
        _lib.check(lib.u1_slab_drift_sums_ext(th[0, k].data_ptr(), pi[0, k].data_ptr(), plane, out_t[0, k].data_ptr(), 0.5 * dt,
                                              sums_f.data_ptr(), Lx, L, ws_d.data_ptr(), ws_d.numel(), sp))
    ref_t = th.clone()
    # reference: drift the own rows AND the row below them (the fused kernel drifts that row in registers)
    _lib.check(lib.u1_slab_drift_wrap(ref_t[0, k].data_ptr(), pi[0, k].data_ptr(), plane, Lx + 1, L, 0.5 * dt, sp))
    sums_r = torch.zeros(3, dtype=torch.float64, device="cuda")
    _lib.check(lib.u1_slab_sums_ext(ref_t[0, k].data_ptr(), pi[0, k].data_ptr(), plane, sums_r.data_ptr(), Lx, L,
                                    ws.data_ptr(), ws.numel(), sp))
    torch.cuda.synchronize()
    assert torch.equal(out_t[:, k:k + Lx], ref_t[:, k:k + Lx])
    assert torch.all(out_t[:, :k] == 0) and torch.all(out_t[:, k + Lx:] == 0)  # nothing written outside the own rows
    assert rel(sums_f, sums_r) < 1e-13, (sums_f, sums_r)
    # -- momenta with sum pi^2 of the own rows
    ws_m = torch.zeros(nred, dtype=torch.uint8, device="cuda")
    p1, p2 = torch.empty_like(pi), torch.empty_like(pi)
    pi2 = torch.zeros(1, dtype=torch.float64, device="cuda")
    _lib.check(lib.u1_momenta_rows_dev(p1.data_ptr(), 11, 3, 5, None, 1000 - k, Lxe, L, 4096, sp))
    for _ in range(2):
        _lib.check(lib.u1_momenta_rows_pi2_dev(p2.data_ptr(), 11, 3, 5, None, 1000 - k, Lxe, L, 4096, k, Lx, pi2.data_ptr(),
                                               ws_m.data_ptr(), ws_m.numel(), sp))
    torch.cuda.synchronize()
    assert torch.equal(p1, p2)
    assert abs(pi2.item() - (p1[:, k:k + Lx] ** 2).sum().item()) < 1e-10 * pi2.item()
    # -- select with carried sums
    state = th.clone()
    own = torch.zeros(2, Lx, L, dtype=torch.float64, device="cuda")
    s_state = torch.tensor([1.0, 2.0], dtype=torch.float64, device="cuda")
    s_new = torch.tensor([7.0, 8.0, 9.0], dtype=torch.float64, device="cuda")
    for acc in (0, 1):
        flag = torch.tensor([acc], dtype=torch.int32, device="cuda")
        _lib.check(lib.u1_slab_select_state(state[0, k].data_ptr(), out_t[0, k].data_ptr(), plane, own.data_ptr(), Lx, L,
                                            flag.data_ptr(), s_state.data_ptr(), s_new.data_ptr(), sp))
        torch.cuda.synchronize()
        if acc == 0:
            assert torch.equal(state, th) and torch.all(own == 0) and s_state.tolist() == [1.0, 2.0]
        else:
            assert torch.equal(state[:, k:k + Lx], out_t[:, k:k + Lx]) and torch.equal(own, out_t[:, k:k + Lx])
            assert torch.equal(state[:, :k], th[:, :k]) and s_state.tolist() == [7.0, 8.0]


def test_full_size_size_independent_properties(hb):
    """BASELINE configs[1] shape (L=64, beta=6, 4096 chains): properties that need no oracle at this size.
    (1) the leapfrog map is reversible: integrate, flip the momenta, integrate again -> the start (mod 2 pi);
    (2) energy error scales like dt^2 (halving dt at fixed trajectory length divides <|dH|> by ~4);
    (3) a chain's result does not depend on its neighbours in the batch (chains 100..103 alone == inside the batch);
    (4) the resident and the streaming kernels agree on dH to round-off with injected momenta."""
    L, B, beta = 64, 4096, 6.0
    g = torch.Generator().manual_seed(2024)
    th = ((torch.rand(B, 2, L, L, generator=g, dtype=torch.float64) * 2 - 1) * 0.6).cuda()
    pi = torch.randn(B, 2, L, L, generator=g, dtype=torch.float64).cuda()
    h = hb.HMC_U1(L, beta, 0, 50, 0.06, n_chains=B, if_tune_step_size=False)
    t1, p1 = h.leapfrog(th, pi)
    t2, p2 = h.leapfrog(t1, -p1)
    d = torch.remainder(t2 - th + math.pi, 2 * math.pi) - math.pi
    assert d.abs().max().item() < 1e-10 and (p2 + pi).abs().max().item() < 1e-9
    # (2) dt^2 scaling of the energy error, same momenta
    r1 = h.trajectories(th.clone(), 1, pi=pi[None].contiguous(), u=torch.zeros(B, dtype=torch.float64, device="cuda"))
    h2 = hb.HMC_U1(L, beta, 0, 100, 0.03, n_chains=B, if_tune_step_size=False)
    r2 = h2.trajectories(th.clone(), 1, pi=pi[None].contiguous(), u=torch.zeros(B, dtype=torch.float64, device="cuda"))
    ratio = r1["dH"].abs().mean().item() / r2["dH"].abs().mean().item()
    assert 3.0 < ratio < 5.5, ratio
    # (3) batch independence at full size
    sub = slice(100, 104)
    hs = hb.HMC_U1(L, beta, 0, 50, 0.06, n_chains=4, if_tune_step_size=False, chain_offset=100)
    hf = hb.HMC_U1(L, beta, 0, 50, 0.06, n_chains=B, if_tune_step_size=False)      # fresh trajectory counter
    ta, tb = th.clone(), th[sub].clone()
    ra = hf.trajectories(ta, 2)
    rb = hs.trajectories(tb, 2)
    assert torch.equal(ta[sub], tb) and torch.equal(ra["dH"][:, sub], rb["dH"])
    # (4) resident vs streaming path, injected draws
    hk = hb.HMC_U1(L, beta, 0, 50, 0.06, n_chains=64, if_tune_step_size=False, kernel_path=1)
    h0 = hb.HMC_U1(L, beta, 0, 50, 0.06, n_chains=64, if_tune_step_size=False)
    u0 = torch.full((64,), 0.5, dtype=torch.float64, device="cuda")
    pa = pi[None, :64].contiguous()
    x0, x1 = th[:64].clone(), th[:64].clone()
    q0 = h0.trajectories(x0, 1, pi=pa, u=u0)
    q1 = hk.trajectories(x1, 1, pi=pa, u=u0)
    assert (q0["dH"] - q1["dH"]).abs().max().item() < 1e-8 and torch.equal(q0["accept"], q1["accept"])
    assert torch.equal(q0["topo"], q1["topo"]) and (q0["plaq"] - q1["plaq"]).abs().max().item() < 1e-12


@pytest.mark.parametrize("L", [6, 8, 16])
def test_rectangle_golden(hb, golden_dir, L):
    """u1_rectangle / utils.rect_from_field_batch against the reference's rect_from_field_batch (utils.py:129-141)."""
    from hmc_ft_b200 import utils as U
    z = np.load(os.path.join(golden_dir, "rect.npz"))
    th = torch.tensor(z[f"theta_L{L}"], dtype=torch.float64, device="cuda")
    r = U.rect_from_field_batch(th)
    assert r.shape == (3, 2, L, L) and rel(r, z[f"rect_L{L}"]) < 1e-15
    assert rel(U.plaq_from_field_batch(th), z[f"plaq_L{L}"]) < 1e-15
    # single configuration and host-tensor call paths
    assert rel(U.rect_from_field_batch(th[1:2])[0], z[f"rect_L{L}"][1]) < 1e-15
    rh = U.rect_from_field_batch(torch.tensor(z[f"theta_L{L}"]))
    assert rh.device.type == "cpu" and rel(rh, z[f"rect_L{L}"]) < 1e-15


def test_tune_step_size_matches_reference_search(hb, golden_dir):
    """HMC_U1.tune_step_size with the draws the REAL reference consumed (tests/golden/tune.npz): same (dt, rate)
    history, same tuned dt (hmc_u1.py:99-161), and the same answer as the oracle's restatement."""
    z = np.load(os.path.join(golden_dir, "tune.npz"))
    L, beta, n_steps, n_tune, target, tol, dt0, max_att = z["params"]
    L, n_steps, n_tune, max_att = int(L), int(n_steps), int(n_tune), int(max_att)
    th = torch.tensor(z["theta"], dtype=torch.float64)
    pi, u = torch.tensor(z["pi"]), torch.tensor(z["u"])
    for path in (0, 1):
        h = hb.HMC_U1(L, float(beta), 0, n_steps, float(dt0), device="cuda", if_tune_step_size=True, kernel_path=path)
        c0 = h.traj_counter
        hist = h.tune_step_size(n_tune, float(target), float(tol), float(dt0), max_att, theta=th.cuda(), draws=(pi, u))
        assert len(hist) == len(z["hist_dt"])
        for (dt, rate), dt_ref, rate_ref in zip(hist, z["hist_dt"], z["hist_rate"]):
            assert abs(dt - dt_ref) < 1e-6
            assert round(rate * n_tune) == round(float(rate_ref) * n_tune)
        assert h.dt == float(z["tuned_dt"])
        assert h.traj_counter == c0 + len(hist)
    ho = O.PlainHMCOracle(L, float(beta), n_steps, float(dt0))
    O.tune_step_size(ho, th[None], pi[:, None], u[:, None], n_tune, float(target), float(tol), float(dt0), max_att)
    assert ho.dt == h.dt


def test_tune_step_size_batched_philox(hb):
    """Without injected draws: the batched search (all trials of an attempt in one launch, chunked when the
    replicas exceed max_bytes) lands inside the tolerance band, is reproducible, and does not depend on the
    chunking."""
    from hmc_ft_b200._runtime import tune_search
    L, beta = 16, 2.0
    res = []
    for max_bytes in (1 << 30, 5 * 4 * 2 * L * L * 8):            # second: 5 trials per launch
        h = hb.HMC_U1(L, beta, 10, 10, 0.2, device="cuda", if_tune_step_size=True, n_chains=4, seed=11)
        base = torch.zeros(4, 2, L, L, dtype=torch.float64, device="cuda")
        h.trajectories(base, 10)
        hist = tune_search(h, base, 64, 0.65, 0.1, 0.9, 10, None, max_bytes)
        assert abs(hist[-1][1] - 0.65) <= 0.1 and h.dt == hist[-1][0]
        res.append(hist)
    assert res[0] == res[1]


def test_misaligned_pointer_is_rejected(hb):
    """include/u1hmc.h: field pointers must be 16-byte aligned (double2 / TMA bulk copies) -> U1_ERR_ARG, no fault."""
    from hmc_ft_b200 import _lib
    lib = _lib.load()
    L, B = 8, 2
    buf = torch.zeros(B * 2 * L * L + 2, dtype=torch.float64, device="cuda")
    odd = buf[1:]                                   # 8-byte aligned only
    assert odd.data_ptr() % 16 == 8
    ws = torch.empty(lib.u1_workspace_bytes(B, L), dtype=torch.uint8, device="cuda")
    for path in (0, 1, 2):
        st = lib.u1_trajectory(odd.data_ptr(), None, None, 1, 0, 0, 1, B, L, 1.0, 0.1, 2, None, None, None, None, None,
                               path, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
        assert st == -1
    torch.cuda.synchronize()


def test_more_than_65535_chains(hb):
    """Chains ride in gridDim.y of the streaming kernels: batches above 65535 are sliced internally."""
    L, B = 6, 65535 + 300                       # L=6 has no resident kernel: streaming path
    g = torch.Generator().manual_seed(3)
    th = ((torch.rand(B, 2, L, L, generator=g, dtype=torch.float64) * 2 - 1) * 3.0).cuda()
    h = hb.HMC_U1(L, 2.0, 0, 4, 0.1, device="cuda", if_tune_step_size=False, n_chains=B, seed=5)
    S = h.action(th)
    So = O.action(th[-400:].cpu(), 2.0)
    assert rel(S[-400:], So) < 1e-12
    a = th.clone()
    r = h.trajectories(a, 1)
    h2 = hb.HMC_U1(L, 2.0, 0, 4, 0.1, device="cuda", if_tune_step_size=False, n_chains=400, seed=5, chain_offset=B - 400)
    b = th[-400:].clone()
    r2 = h2.trajectories(b, 1)
    # the streaming path's reduction blocking depends on the batch size, so dH agrees to round-off, not bit for bit
    assert (r["dH"][0, -400:] - r2["dH"][0]).abs().max().item() < 1e-10
    assert torch.equal(r["accept"][0, -400:], r2["accept"][0]) and torch.equal(a[-400:], b)
