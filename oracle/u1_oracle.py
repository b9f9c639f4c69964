"""CPU oracle for the 2D U(1) HMC / FT-HMC hot path.  TEST INFRASTRUCTURE ONLY.

This file is a torch-CPU fp64 *restatement* of the reference algorithm
(Greyyy-HJC/hmc_ft, canonical generation ``2d_u1_cluster/``).  It is the checker
used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.  Nothing under ``hmc_ft_b200/`` may
import it: the product path is the CUDA library and fails loudly without it.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function
here against (a) tensors produced by importing the real reference classes in
fp64 (``tools/gen_golden.py`` -> ``tests/golden/*.npz``), (b) the reference's
own fp64 golden log ``tests/test.out`` (first trajectories, replayed with the
torch CPU generator seeded as the reference seeds it) and (c) the dense
autograd Jacobian log-det that the reference's ``if_check_jac`` path uses.

The reference is Python/torch, so the restatement is Python/torch as well:
same operator sequence (``torch.roll`` stencils, ``autograd.grad`` forces),
same conventions, batched over a leading chain axis where the reference is
single-chain.  Every function cites the reference lines it follows
(paths relative to ``/root/reference/2d_u1_cluster``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

DT = torch.float64
TWO_PI = 2.0 * math.pi


# --------------------------------------------------------------------------
# L1a lattice primitives (utils.py:118-196)
# --------------------------------------------------------------------------
def plaq_from_field(theta: torch.Tensor) -> torch.Tensor:
    """P(x,y)=t0(x,y)-t1(x,y)-t0(x,y+1)+t1(x+1,y); utils.py:143-150 (single),
    utils.py:118-127 (batch).  Accepts [...,2,L,L]."""
    t0, t1 = theta[..., 0, :, :], theta[..., 1, :, :]
    return t0 - t1 - torch.roll(t0, shifts=-1, dims=-1) + torch.roll(t1, shifts=-1, dims=-2)


def rect_from_field(theta: torch.Tensor) -> torch.Tensor:
    """2x1 / 1x2 rectangle angles, utils.py:129-141.  [...,2,L,L] -> [...,2,L,L]."""
    t0, t1 = theta[..., 0, :, :], theta[..., 1, :, :]
    r0 = (t0 + torch.roll(t0, -1, -2) + torch.roll(t1, -2, -2)
          - torch.roll(t0, (-1, -1), (-2, -1)) - torch.roll(t0, -1, -1) - t1)
    r1 = (t0 + torch.roll(t1, -1, -2) + torch.roll(t1, (-1, -1), (-2, -1))
          - torch.roll(t0, -2, -1) - torch.roll(t1, -1, -1) - t1)
    return torch.stack([r0, r1], dim=-3)


def regularize(theta: torch.Tensor) -> torch.Tensor:
    """Wrap to [-pi,pi); utils.py:182-187 (exact formula, division included)."""
    w = (theta - math.pi) / (2 * math.pi)
    return 2 * math.pi * (w - torch.floor(w) - 0.5)


def action(theta: torch.Tensor, beta: float) -> torch.Tensor:
    """S=-beta*sum cos(regularize(P)); hmc_u1.py:56-63, hmc_u1_ft.py:77-98.
    [...,2,L,L] -> [...]."""
    return (-beta) * torch.sum(torch.cos(regularize(plaq_from_field(theta))), dim=(-2, -1))


def force_autograd(theta: torch.Tensor, beta: float) -> torch.Tensor:
    """dS/dtheta exactly as the reference obtains it; hmc_u1.py:65-69."""
    t = theta.detach().clone().requires_grad_(True)
    s = action(t, beta).sum()
    return torch.autograd.grad(s, t)[0].detach()


def force_analytic(theta: torch.Tensor, beta: float) -> torch.Tensor:
    """Closed form of hmc_u1.py:65-69:
    F0=beta[sinP(x,y)-sinP(x,y-1)], F1=beta[-sinP(x,y)+sinP(x-1,y)]."""
    s = torch.sin(regularize(plaq_from_field(theta)))
    f0 = beta * (s - torch.roll(s, 1, -1))
    f1 = beta * (-s + torch.roll(s, 1, -2))
    return torch.stack([f0, f1], dim=-3)


def leapfrog(theta, pi, dt: float, n_steps: int, force_fn):
    """hmc_u1.py:71-80 / hmc_u1_ft.py:154-179.  n_steps force evaluations,
    half drifts at both ends, theta wrapped at the end, momentum not negated."""
    th = theta + 0.5 * dt * pi
    p = pi - dt * force_fn(th)
    for _ in range(n_steps - 1):
        th = th + dt * p
        p = p - dt * force_fn(th)
    th = th + 0.5 * dt * p
    return regularize(th), p


def plaq_mean(theta: torch.Tensor) -> torch.Tensor:
    """utils.py:173-180."""
    return torch.mean(torch.cos(regularize(plaq_from_field(theta))), dim=(-2, -1))


def topo_charge(theta: torch.Tensor) -> torch.Tensor:
    """utils.py:189-196: floor(0.1 + sum(wrap P)/2pi)."""
    return torch.floor(0.1 + torch.sum(regularize(plaq_from_field(theta)), dim=(-2, -1)) / (2 * math.pi))


def metropolis_step(theta, pi, u, dt, n_steps, action_fn, force_fn):
    """hmc_u1.py:82-97 / hmc_u1_ft.py:181-209 with the random draws injected.

    theta,pi: [B,2,L,L]; u: [B] uniforms.  Returns (theta_out, accepted[B] bool,
    H_out[B], dH[B])."""
    H_old = action_fn(theta) + 0.5 * torch.sum(pi ** 2, dim=(-3, -2, -1))
    new_theta, new_pi = leapfrog(theta.clone(), pi.clone(), dt, n_steps, force_fn)
    H_new = action_fn(new_theta) + 0.5 * torch.sum(new_pi ** 2, dim=(-3, -2, -1))
    dH = H_new - H_old
    acc = u < torch.exp(-dH)          # NaN compares False -> reject, as in the reference
    theta_out = torch.where(acc[..., None, None, None], new_theta, theta)
    return theta_out, acc, torch.where(acc, H_new, H_old), dH


# --------------------------------------------------------------------------
# Masks (utils.py:21-115)
# --------------------------------------------------------------------------
def field_mask(index: int, L: int) -> torch.Tensor:
    """utils.py:21-50 -> bool [2,L,L]."""
    m = torch.zeros((2, L, L), dtype=torch.bool)
    mu, k = divmod(index, 4)
    m[mu, (k >> 1) & 1::2, k & 1::2] = True
    return m


def plaq_mask(index: int, L: int) -> torch.Tensor:
    """utils.py:52-80 -> bool [L,L]."""
    m = torch.zeros((L, L), dtype=torch.bool)
    if index in (0, 1):
        m[1::2, :] = True
    elif index in (2, 3):
        m[0::2, :] = True
    elif index in (4, 6):
        m[:, 1::2] = True
    else:
        m[:, 0::2] = True
    return m


def rect_mask(index: int, L: int) -> torch.Tensor:
    """utils.py:82-115 -> bool [2,L,L]."""
    m = torch.zeros((2, L, L), dtype=torch.bool)
    if index == 0:
        m[1, 1::2, :] = True; m[1, 0::2, 1::2] = True
    elif index == 1:
        m[1, 1::2, :] = True; m[1, 0::2, 0::2] = True
    elif index == 2:
        m[1, 0::2, :] = True; m[1, 1::2, 1::2] = True
    elif index == 3:
        m[1, 0::2, :] = True; m[1, 1::2, 0::2] = True
    elif index == 4:
        m[0, :, 1::2] = True; m[0, 1::2, 0::2] = True
    elif index == 5:
        m[0, :, 0::2] = True; m[0, 1::2, 1::2] = True
    elif index == 6:
        m[0, :, 1::2] = True; m[0, 0::2, 0::2] = True
    elif index == 7:
        m[0, :, 0::2] = True; m[0, 0::2, 1::2] = True
    return m


# --------------------------------------------------------------------------
# CNN (cnn_model_opt.py:6-247) with explicit weights
# --------------------------------------------------------------------------
def gelu(x: torch.Tensor) -> torch.Tensor:
    """nn.GELU() default = exact erf form."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def conv3x3_circular(x, w, b):
    """nn.Conv2d(..., 3, padding='same', padding_mode='circular'): cross-correlation
    out[o,x,y]=b[o]+sum_{c,i,j} w[o,c,i,j]*in[c,(x+i-1)%L,(y+j-1)%L]."""
    xp = F.pad(x, (1, 1, 1, 1), mode="circular")
    return F.conv2d(xp, w, b)


class CNNWeights:
    """One subset's network.  ``arch`` in {'simple','rnet1','rnet3'}; ``layers`` is the
    ordered list of (weight[o,c,3,3], bias[o]) fp64 tensors:
    simple: conv1, conv2 (cnn_model_opt.py:6-52)
    rnetN : initial_conv, N x (conv1, conv2), final_conv (cnn_model_opt.py:54-87,131-247)."""

    def __init__(self, arch: str, layers: Sequence[Tuple[torch.Tensor, torch.Tensor]]):
        self.arch = arch
        self.layers = [(w.to(DT).contiguous(), b.to(DT).contiguous()) for w, b in layers]
        n = {"simple": 2, "rnet1": 4, "rnet3": 8}[arch]
        assert len(self.layers) == n, (arch, len(self.layers))

    def __call__(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        L = self.layers
        h = gelu(conv3x3_circular(x, *L[0]))
        for i in range(1, len(L) - 1, 2):          # ResBlock: cnn_model_opt.py:74-87
            u = gelu(conv3x3_circular(h, *L[i]))
            h = gelu(conv3x3_circular(u, *L[i + 1]) + h)
        y = gelu(conv3x3_circular(h, *L[-1]))
        k = torch.arctan(y) / math.pi / 2
        return k[:, :4], k[:, 4:]


STATE_KEYS = {
    "simple": ["conv1", "conv2"],
    "rnet1": ["initial_conv", "res_blocks.0.conv1", "res_blocks.0.conv2", "final_conv"],
    "rnet3": ["initial_conv"] + [f"res_blocks.{i}.conv{j}" for i in range(3) for j in (1, 2)] + ["final_conv"],
}


def weights_from_state_dicts(arch: str, state_dicts: Sequence[Dict[str, torch.Tensor]]) -> List[CNNWeights]:
    """field_trans_opt.py:654-692: per-subset state dicts, optional 'module.' prefix."""
    out = []
    for sd in state_dicts:
        sd = {k.replace("module.", ""): torch.as_tensor(np.asarray(v)) for k, v in sd.items()}
        out.append(CNNWeights(arch, [(sd[k + ".weight"], sd[k + ".bias"]) for k in STATE_KEYS[arch]]))
    return out


def load_weights_npz(path: str) -> Tuple[str, List[CNNWeights]]:
    """Load the repo's compact weight fixture (tools/gen_golden.py writes it):
    keys 'arch' and 's{subset}.{layer}.weight|bias'."""
    z = np.load(path, allow_pickle=False)
    arch = str(z["arch"])
    sds = []
    for s in range(8):
        sds.append({k[len(f"s{s}."):]: z[k] for k in z.files if k.startswith(f"s{s}.")})
    return arch, weights_from_state_dicts(arch, sds)


# --------------------------------------------------------------------------
# Field transformation, inference half (field_trans_opt.py:110-398)
# --------------------------------------------------------------------------
class FieldTransformationOracle:
    def __init__(self, L: int, nets: Sequence[CNNWeights]):
        assert L % 2 == 0 and len(nets) == 8
        self.L, self.nets = L, list(nets)
        self.fm = [field_mask(i, L) for i in range(8)]
        self.pm = [plaq_mask(i, L) for i in range(8)]
        self.rm = [rect_mask(i, L) for i in range(8)]

    def features(self, theta, index):
        """field_trans_opt.py:110-131: [sin(mP P),cos(mP P),sin(m0R0),sin(m1R1),cos(m0R0),cos(m1R1)]."""
        plaq = plaq_from_field(theta)
        rect = rect_from_field(theta)
        pmk = plaq * self.pm[index]
        rmk = rect * self.rm[index]
        return torch.cat([torch.sin(pmk)[:, None], torch.cos(pmk)[:, None],
                          torch.sin(rmk), torch.cos(rmk)], dim=1), plaq, rect

    def _terms(self, theta, index, trig_p, trig_r, signs: bool):
        """Shared body of ft_phase (field_trans_opt.py:137-217, trig=sin, signed)
        and the Jacobian shift (field_trans_opt.py:284-375, trig=-cos)."""
        X, plaq, rect = self.features(theta, index)
        K0, K1 = self.nets[index](X)
        r0, r1 = rect[:, 0], rect[:, 1]
        pa = torch.stack([plaq, torch.roll(plaq, 1, 2), plaq, torch.roll(plaq, 1, 1)], dim=1)
        ra = torch.stack([torch.roll(r0, 1, 1), torch.roll(r0, (1, 1), (1, 2)), r0, torch.roll(r0, 1, 2),
                          torch.roll(r1, 1, 2), torch.roll(r1, (1, 1), (1, 2)), r1, torch.roll(r1, 1, 1)], dim=1)
        if signs:
            sp = torch.tensor([-1., 1., 1., -1.], dtype=DT).view(1, 4, 1, 1)
            sr = torch.tensor([-1., 1., -1., 1., 1., -1., 1., -1.], dtype=DT).view(1, 8, 1, 1)
        else:
            sp = -torch.ones(1, 4, 1, 1, dtype=DT)
            sr = -torch.ones(1, 8, 1, 1, dtype=DT)
        tp = K0 * (trig_p(pa) * sp)
        tr = K1 * (trig_r(ra) * sr)
        out = torch.stack([tp[:, 0] + tp[:, 1] + tr[:, 0] + tr[:, 1] + tr[:, 2] + tr[:, 3],
                           tp[:, 2] + tp[:, 3] + tr[:, 4] + tr[:, 5] + tr[:, 6] + tr[:, 7]], dim=1)
        return out * self.fm[index]

    def ft_phase(self, theta, index):
        """field_trans_opt.py:137-217."""
        return self._terms(theta, index, torch.sin, torch.sin, True)

    def jac_shift(self, theta, index):
        """field_trans_opt.py:319-372: J with every +-sin replaced by -cos."""
        return self._terms(theta, index, torch.cos, torch.cos, False)

    def forward(self, theta):
        """field_trans_opt.py:219-235.  [B,2,L,L] -> [B,2,L,L]."""
        cur = theta.clone()
        for k in range(8):
            cur = cur + self.ft_phase(cur, k)
        return cur

    def inverse(self, theta, max_iter: int = 200, tol: float = 1e-6):
        """field_trans_opt.py:245-282: per-subset fixed-point iteration in reverse subset order;
        returns (theta_new, iterations per subset in the order visited)."""
        cur = theta.clone()
        iters = []
        for index in reversed(range(8)):
            it = cur.clone()
            diff, n_it = float("inf"), 0
            for i in range(max_iter):
                nxt = cur + (-self.ft_phase(it, index))
                diff = torch.norm(nxt - it) / torch.norm(it)
                n_it = i + 1
                if diff < tol:
                    cur = nxt
                    break
                it = nxt
            iters.append(n_it)
        return cur, iters

    def forward_and_logdet(self, theta):
        """field_trans_opt.py:284-383 (one CNN evaluation per subset instead of two;
        identical values)."""
        cur = theta.clone()
        logdet = torch.zeros(theta.shape[0], dtype=DT)
        for k in range(8):
            logdet = logdet + torch.log(1 + self.jac_shift(cur, k)).sum(dim=(1, 2, 3))
            cur = cur + self.ft_phase(cur, k)
        return cur, logdet

    def compute_jac_logdet(self, theta):
        return self.forward_and_logdet(theta)[1]

    def logdet_dense_autograd(self, theta1):
        """field_trans_opt.py:385-390 (if_check_jac): logdet of the dense Jacobian; L<=8."""
        jac = torch.autograd.functional.jacobian(self.forward, theta1[None])
        n = theta1.numel()
        return torch.logdet(jac.reshape(n, n))

    # ---- HMC_U1_FT pieces (hmc_u1_ft.py:101-152) ----
    def new_action(self, theta, beta):
        """S_FT = S(F(theta)) - logdet; hmc_u1_ft.py:101-125.  [B,2,L,L]->[B]."""
        ori, ld = self.forward_and_logdet(theta)
        return action(ori, beta) - ld

    def new_force(self, theta, beta):
        """autograd of new_action; hmc_u1_ft.py:128-152."""
        t = theta.detach().clone().requires_grad_(True)
        s = self.new_action(t, beta).sum()
        return torch.autograd.grad(s, t)[0].detach()

    # ---- training loss, forward value only (field_trans_opt.py:400-460, 490-505) ----
    def compute_force(self, theta, beta, transformed: bool = False):
        """field_trans_opt.py:400-441: force of -beta sum cos P (no regularize) or of S_FT."""
        if transformed:
            return self.new_force(theta, beta)
        return force_analytic(theta, beta)

    def train_loss_and_grads(self, theta_ori, train_beta: float):
        """field_trans_opt.py:443-480 (train_step up to the optimizer): loss_fn with autograd through the unrolled
        fixed-point inverse (:245-282) and through compute_force(create_graph=True) (:400-441), then
        d loss / d (every CNN weight and bias).  Returns (loss, grads) with grads[subset][layer] = (dW, db) in the
        layer order of CNNWeights.  Also usable piecewise (see pieces=True) for debugging a device implementation."""
        params = []
        for net in self.nets:
            for i, (w, b) in enumerate(net.layers):
                w = w.detach().clone().requires_grad_(True)
                b = b.detach().clone().requires_grad_(True)
                net.layers[i] = (w, b)
                params += [w, b]
        th_new, _ = self.inverse(theta_ori)
        t = th_new if th_new.requires_grad else th_new.clone().requires_grad_(True)
        f_ori = torch.autograd.grad(action(t, 1.0).sum(), t, create_graph=True)[0]
        f_new = torch.autograd.grad(self.new_action(t, train_beta).sum(), t, create_graph=True)[0]
        vol = float(self.L * self.L)
        d = f_new - f_ori
        loss = sum(torch.norm(d, p=p) / vol ** (1.0 / p) for p in (2, 4, 6, 8))
        g = torch.autograd.grad(loss, params, allow_unused=True)
        grads, it = [], iter(g)
        for net in self.nets:
            grads.append([(next(it), next(it)) for _ in net.layers])
        for net in self.nets:                      # leave the oracle with plain tensors again
            net.layers = [(w.detach(), b.detach()) for w, b in net.layers]
        return loss.detach(), grads

    def loss_fn(self, theta_ori, train_beta: float):
        """field_trans_opt.py:443-462: theta_new = inverse(theta_ori); the four p-norms (p = 2,4,6,8, taken over
        the WHOLE batch tensor) of F_FT(theta_new; train_beta) - F_plain(theta_new; beta=1), each divided by
        vol^(1/p), vol = L^2."""
        th_new, _ = self.inverse(theta_ori)
        d = self.compute_force(th_new, train_beta, transformed=True) - self.compute_force(th_new, 1.0)
        vol = float(self.L * self.L)
        return sum(torch.norm(d, p=p) / vol ** (1.0 / p) for p in (2, 4, 6, 8))


# --------------------------------------------------------------------------
# Index-loop spec of one subset update (small L only; independent of torch.roll)
# --------------------------------------------------------------------------
def ft_phase_loops(theta: np.ndarray, K0: np.ndarray, K1: np.ndarray, index: int) -> np.ndarray:
    """SURVEY §8a F5 restated with explicit loops for one chain: theta[2,L,L],
    K0[4,L,L], K1[8,L,L] -> Delta[2,L,L].  Used to pin the roll conventions."""
    L = theta.shape[-1]
    t0, t1 = theta

    def P(x, y):
        x %= L; y %= L
        return t0[x, y] - t1[x, y] - t0[x, (y + 1) % L] + t1[(x + 1) % L, y]

    def R0(x, y):
        return P(x, y) + P(x + 1, y)

    def R1(x, y):
        return P(x, y) + P(x, y + 1)

    out = np.zeros_like(theta)
    mu, k = divmod(index, 4)
    px, py = (k >> 1) & 1, k & 1
    s = math.sin
    for x in range(px, L, 2):
        for y in range(py, L, 2):
            if mu == 0:
                out[0, x, y] = (-K0[0, x, y] * s(P(x, y)) + K0[1, x, y] * s(P(x, y - 1))
                                - K1[0, x, y] * s(R0(x - 1, y)) + K1[1, x, y] * s(R0(x - 1, y - 1))
                                - K1[2, x, y] * s(R0(x, y)) + K1[3, x, y] * s(R0(x, y - 1)))
            else:
                out[1, x, y] = (K0[2, x, y] * s(P(x, y)) - K0[3, x, y] * s(P(x - 1, y))
                                + K1[4, x, y] * s(R1(x, y - 1)) - K1[5, x, y] * s(R1(x - 1, y - 1))
                                + K1[6, x, y] * s(R1(x, y)) - K1[7, x, y] * s(R1(x - 1, y)))
    return out


# --------------------------------------------------------------------------
# Samplers (single chain semantic of the reference, batched over B)
# --------------------------------------------------------------------------
class PlainHMCOracle:
    """HMC_U1 (hmc_u1.py) with injected randomness, batched over chains."""

    def __init__(self, L, beta, n_steps, dt, analytic_force=False):
        self.L, self.beta, self.n_steps, self.dt = L, beta, n_steps, dt
        self._force = force_analytic if analytic_force else force_autograd

    def action(self, theta):
        return action(theta, self.beta)

    def force(self, theta):
        return self._force(theta, self.beta)

    def leapfrog(self, theta, pi):
        return leapfrog(theta, pi, self.dt, self.n_steps, self.force)

    def metropolis_step(self, theta, pi, u):
        return metropolis_step(theta, pi, u, self.dt, self.n_steps, self.action, self.force)


def tune_step_size(hmc, theta, pi_draws, u_draws, n_tune_steps=2000, target_rate=0.65, target_tolerance=0.15,
                   initial_step_size=0.2, max_attempts=10):
    """Binary search on dt, hmc_u1.py:99-161 (same for hmc_u1_ft.py:211-271), with the draws injected:
    trial i of the whole search (attempt * n_tune_steps + step) consumes pi_draws[i], u_draws[i] - the
    order in which the reference's metropolis_step pulls randn_like / rand from torch's generator.
    Every trial starts from the same theta (the returned state is discarded, hmc_u1.py:132).
    `hmc` is a PlainHMCOracle / FTHMCOracle; its dt is updated like the reference's self.dt.
    Returns the history [(dt tried, acceptance rate)]; the tuned value is hmc.dt."""
    hmc.dt = initial_step_size
    step_min, step_max = 1e-6, 1.0
    best_dt, best_diff = hmc.dt, float("inf")
    history, i, rate = [], 0, 0.0
    for _attempt in range(max_attempts):
        count = 0
        for _ in range(n_tune_steps):
            _, acc, _, _ = hmc.metropolis_step(theta, pi_draws[i], u_draws[i])
            count += int(bool(torch.as_tensor(acc).reshape(-1)[0]))
            i += 1
        rate = count / n_tune_steps
        diff = abs(rate - target_rate)
        history.append((hmc.dt, rate))
        if diff < best_diff:
            best_dt, best_diff = hmc.dt, diff
        if abs(rate - target_rate) <= target_tolerance:
            break
        if rate > target_rate:
            step_min = hmc.dt
            hmc.dt = min((hmc.dt + step_max) / 2, step_max)
        else:
            step_max = hmc.dt
            hmc.dt = max((hmc.dt + step_min) / 2, step_min)
    if abs(rate - target_rate) > target_tolerance:
        hmc.dt = best_dt
    return history


class FTHMCOracle:
    """HMC_U1_FT (hmc_u1_ft.py) with injected randomness, batched over chains."""

    def __init__(self, ft: FieldTransformationOracle, beta, n_steps, dt):
        self.ft, self.beta, self.n_steps, self.dt = ft, beta, n_steps, dt

    def action(self, theta):
        return self.ft.new_action(theta, self.beta)

    def force(self, theta):
        return self.ft.new_force(theta, self.beta)

    def leapfrog(self, theta, pi):
        return leapfrog(theta, pi, self.dt, self.n_steps, self.force)

    def metropolis_step(self, theta, pi, u):
        return metropolis_step(theta, pi, u, self.dt, self.n_steps, self.action, self.force)


# --------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., SC'11; Random123 v1.14 published algorithm) —
# the counter-based generator the CUDA path uses for momenta and accept draws.
# Not part of the reference (it uses torch's global generator); restated here so
# the device stream can be checked bit-for-bit.
# --------------------------------------------------------------------------
PHILOX_M0, PHILOX_M1 = 0xD2511F53, 0xCD9E8D57
PHILOX_W0, PHILOX_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr: uint32 [...,4], key: uint32 [...,2] -> uint32 [...,4]."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0 = key[..., 0].astype(np.uint64)
    k1 = key[..., 1].astype(np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(PHILOX_M0) * c[0]
        p1 = np.uint64(PHILOX_M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [(hi1 ^ c[1] ^ k0) & mask, lo1, (hi0 ^ c[3] ^ k1) & mask, lo0]
        k0 = (k0 + np.uint64(PHILOX_W0)) & mask
        k1 = (k1 + np.uint64(PHILOX_W1)) & mask
    return np.stack(c, axis=-1).astype(np.uint32)


def u01_from_bits(hi: np.ndarray, lo: np.ndarray) -> np.ndarray:
    """53-bit uniform in (0,1): ((hi<<21 | lo>>11) + 0.5) * 2^-53 — matches csrc/u1_common.cuh."""
    m = (hi.astype(np.uint64) << np.uint64(21)) | (lo.astype(np.uint64) >> np.uint64(11))
    return (m.astype(np.float64) + 0.5) * (2.0 ** -53)


def philox_momenta(seed: int, chain: int, traj: int, L: int) -> np.ndarray:
    """Momentum field pi[2,L,L] for (seed, global chain id, trajectory index), as the
    CUDA path generates it: one Philox call per site s=x*L+y with counter
    (s, traj_lo, traj_hi, 0) and key (seed_lo ^ chain-mixed, ...) — see csrc/u1_common.cuh.
    Box-Muller on two 53-bit uniforms gives (pi0, pi1) of that site."""
    s = np.arange(L * L, dtype=np.uint32)
    ctr = np.stack([s, np.full_like(s, traj & 0xFFFFFFFF), np.full_like(s, (traj >> 32) & 0xFFFFFFFF),
                    np.zeros_like(s)], axis=-1)
    key = np.array([(seed & 0xFFFFFFFF) ^ ((chain * 0x9E3779B9) & 0xFFFFFFFF),
                    ((seed >> 32) & 0xFFFFFFFF) ^ (chain & 0xFFFFFFFF)], dtype=np.uint32)
    r = philox4x32_10(ctr, np.broadcast_to(key, (L * L, 2)))
    u1 = u01_from_bits(r[:, 0], r[:, 1])
    u2 = u01_from_bits(r[:, 2], r[:, 3])
    rad = np.sqrt(-2.0 * np.log(u1))
    ang = TWO_PI * u2
    return np.stack([rad * np.cos(ang), rad * np.sin(ang)]).reshape(2, L, L)


def philox_accept_uniform(seed: int, chain: int, traj: int) -> float:
    """Accept/reject uniform of (seed, chain, traj): counter (0xFFFFFFFF, traj_lo, traj_hi, 1)."""
    ctr = np.array([[0xFFFFFFFF, traj & 0xFFFFFFFF, (traj >> 32) & 0xFFFFFFFF, 1]], dtype=np.uint32)
    key = np.array([[(seed & 0xFFFFFFFF) ^ ((chain * 0x9E3779B9) & 0xFFFFFFFF),
                     ((seed >> 32) & 0xFFFFFFFF) ^ (chain & 0xFFFFFFFF)]], dtype=np.uint32)
    r = philox4x32_10(ctr, key)
    return float(u01_from_bits(r[:, 0], r[:, 1])[0])


# --------------------------------------------------------------------------
# Observable post-processing (utils.py:207-298) — numpy, as in the reference
# --------------------------------------------------------------------------
def chi_infinity(beta: float) -> float:
    """utils.py:207-222."""
    from scipy.integrate import quad
    num, _ = quad(lambda p: (p / (2 * math.pi)) ** 2 * math.exp(beta * math.cos(p)), -math.pi, math.pi)
    den, _ = quad(lambda p: math.exp(beta * math.cos(p)), -math.pi, math.pi)
    return num / den


def auto_from_chi(topo, max_lag: int, beta: float, volume: int) -> np.ndarray:
    """utils.py:224-262: Gamma(d) = 1 - <(Q_t - Q_{t+d})^2> / (2 V chi_inf), Gamma(0) = 1."""
    topo = np.round(np.asarray(topo)).astype(int)
    chi = chi_infinity(beta)
    out = np.zeros(max_lag + 1)
    for d in range(max_lag + 1):
        out[d] = 1.0 if d == 0 else 1 - np.mean((topo[:-d] - topo[d:]) ** 2) / (2 * volume * chi)
    return out


def auto_by_def(topo, max_lag: int) -> np.ndarray:
    """utils.py:264-298: normalised autocovariance, all ones if the variance vanishes."""
    topo = np.round(np.asarray(topo)).astype(int)
    mean, var = np.mean(topo), np.var(topo)
    if var == 0:
        return np.ones(max_lag + 1)
    out = np.zeros(max_lag + 1)
    for d in range(max_lag + 1):
        out[d] = 1.0 if d == 0 else np.mean((topo[:-d] - mean) * (topo[d:] - mean)) / var
    return out


def auto_from_chi_bootstrap(topo, max_lag: int, beta: float, volume: int, n_bootstrap: int = 1000, random_seed=None):
    """utils.py:300-365: bootstrap mean / std of auto_from_chi with numpy's legacy generator."""
    if random_seed is not None:
        np.random.seed(random_seed)
    topo = np.round(np.asarray(topo)).astype(int)
    n = len(topo)
    chi = chi_infinity(beta)
    res = np.zeros((n_bootstrap, max_lag + 1))
    for i in range(n_bootstrap):
        bt = topo[np.random.choice(n, size=n, replace=True)]
        res[i, 0] = 1.0
        for d in range(1, max_lag + 1):
            res[i, d] = 1 - np.mean((bt[:-d] - bt[d:]) ** 2) / (2 * volume * chi)
    return np.mean(res, axis=0), np.std(res, axis=0)
