"""FieldTransformation (2d_u1_cluster/field_trans_opt.py:26-692) on the sm_100a library: forward, log-det,
inverse, S_FT and its force, the training loss (loss_fn / evaluate_step) and its hand-derived gradient with respect
to the CNN weights (loss_and_grad -> train_step / train with AdamW on the host), checkpoint save / load."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._runtime import WS, as_batch, check_field, require_cuda, stream_ptr, to_dev
from .cnn_model import choose_cnn_model


class FieldTransformation:
    def __init__(self, lattice_size, device="cuda", n_subsets=8, if_check_jac=False, num_workers=0,
                 identity_init=True, save_tag=None, model_tag="simple", fabric=None):
        """Same signature as field_trans_opt.py:28 (fabric / num_workers are accepted and unused)."""
        if n_subsets != 8:
            raise ValueError("the checkerboard decomposition has exactly 8 subsets (utils.py:21-50)")
        if lattice_size % 2:
            raise ValueError("lattice_size must be even (utils.py:21-115 masks)")
        self.L = int(lattice_size)
        self.device = torch.device(device)
        self.n_subsets = n_subsets
        if if_check_jac and lattice_size > 16:
            raise ValueError("if_check_jac builds the dense (2 L^2)^2 Jacobian: supported for lattice_size <= 16 "
                             "(the reference's autograd version, field_trans_opt.py:385-390, is only feasible there too)")
        self.if_check_jac = if_check_jac
        self.num_workers = num_workers
        self.train_beta = None
        self.model_tag = model_tag
        self.save_tag = save_tag
        self.fabric = fabric
        self.print = print
        self._lib = _lib.load()
        self._dev = require_cuda(self.device if self.device.type == "cuda" else None)
        cnn = choose_cnn_model(model_tag)
        self.models = [cnn() for _ in range(n_subsets)]
        if identity_init:                       # field_trans_opt.py:46-51
            for m in self.models:
                for p in m.parameters():
                    torch.nn.init.normal_(p, mean=0.0, std=0.001)
        self._handle: Optional[C.c_void_p] = None
        self.gpu_launches = 0
        self.sync_weights()

    # ------------------------------------------------------------------ weights
    def sync_weights(self) -> None:
        """(Re)build the C handle from self.models (call after changing parameters)."""
        blob = torch.cat([m.packed_weights() for m in self.models]).contiguous()
        n_res = self.models[0].n_res_blocks
        assert blob.numel() == self._lib.u1ft_weight_count(n_res)
        h = C.c_void_p()
        with torch.cuda.device(self._dev):       # the handle's device table must live on the compute device
            _lib.check(self._lib.u1ft_create(C.byref(h), n_res, blob.data_ptr(), blob.numel()), "u1ft_create")
            self._release()
        self._handle = h

    def _release(self):
        if getattr(self, "_handle", None) is not None and self._handle:
            self._lib.u1ft_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _load_best_model(self, train_beta):
        """field_trans_opt.py:654-692: CWD-relative checkpoint, per-subset state dicts, optional
        'module.' prefix."""
        tag = "" if self.save_tag is None else f"_{self.save_tag}"
        path = f"models/best_model_opt_L{self.L}_train_beta{train_beta:.1f}{tag}.pt"
        self.load_checkpoint(path)

    def load_checkpoint(self, path: str) -> None:
        ck = torch.load(path, map_location="cpu", weights_only=False)
        for i, m in enumerate(self.models):
            key = f"model_state_dict_{i}"
            if key not in ck:
                raise KeyError(f"State dict for model {i} not found in checkpoint")
            sd = {k.replace("module.", ""): v for k, v in ck[key].items()}
            m.load_state_dict(sd)
        self.sync_weights()
        if "epoch" in ck and "loss" in ck:
            self.print(f"Loaded best models from epoch {ck['epoch'] + 1} with loss {ck['loss']:.6f}")

    def load_weights_npz(self, path: str) -> None:
        """Compact fixture format of this repo: keys 'arch', 's{subset}.{layer}.weight|bias'."""
        z = np.load(path, allow_pickle=False)
        if str(z["arch"]) != self.model_tag:
            raise ValueError(f"{path} holds '{z['arch']}' weights, model_tag is '{self.model_tag}'")
        for s, m in enumerate(self.models):
            sd = {k[len(f"s{s}."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"s{s}.")}
            m.load_state_dict(sd)
        self.sync_weights()

    # ------------------------------------------------------------------ helpers
    def _stage(self, theta):
        tb, single = as_batch(theta)
        d = to_dev(tb, self._dev)
        B, L = check_field(d)
        if L != self.L:
            raise ValueError(f"lattice size {L} != {self.L}")
        return d, single, B

    def _back(self, t, like, single=False):
        if single:
            t = t[0]
        return t if like.device == t.device else t.to(like.device)

    def workspace(self, B: int, backward: bool):
        nb = self._lib.u1ft_workspace_bytes(self._handle, B, self.L, 1 if backward else 0)
        return WS.get(self._dev, nb, "ft")

    def _count(self):
        self.gpu_launches += _lib.last_launch_count()

    # ------------------------------------------------------------------ reference surface
    def forward(self, theta):
        """field_trans_opt.py:219-235: [B,2,L,L] -> [B,2,L,L]."""
        d, single, B = self._stage(theta)
        out = torch.empty_like(d)
        with torch.cuda.device(self._dev):
            ws, nb = self.workspace(B, False)
            _lib.check(self._lib.u1ft_forward(self._handle, d.data_ptr(), out.data_ptr(), None, B, self.L, ws, nb,
                                              stream_ptr(self._dev)), "u1ft_forward")
        self._count()
        return self._back(out, theta, single)

    def field_transformation(self, theta):
        """field_trans_opt.py:237-239 (single configuration [2,L,L])."""
        return self.forward(theta)

    def compute_jac_logdet(self, theta):
        """field_trans_opt.py:284-383: [B,2,L,L] -> [B]."""
        d, single, B = self._stage(theta)
        out = torch.empty(B, dtype=torch.float64, device=self._dev)
        with torch.cuda.device(self._dev):
            ws, nb = self.workspace(B, False)
            _lib.check(self._lib.u1ft_forward(self._handle, d.data_ptr(), None, out.data_ptr(), B, self.L, ws, nb,
                                              stream_ptr(self._dev)), "u1ft_forward")
        self._count()
        return self._back(out, theta)

    def forward_and_logdet(self, theta):
        d, single, B = self._stage(theta)
        out = torch.empty_like(d)
        ld = torch.empty(B, dtype=torch.float64, device=self._dev)
        with torch.cuda.device(self._dev):
            ws, nb = self.workspace(B, False)
            _lib.check(self._lib.u1ft_forward(self._handle, d.data_ptr(), out.data_ptr(), ld.data_ptr(), B, self.L,
                                              ws, nb, stream_ptr(self._dev)), "u1ft_forward")
        self._count()
        return self._back(out, theta, single), self._back(ld, theta)

    def ft_phase(self, theta, index):
        """field_trans_opt.py:137-217."""
        d, single, B = self._stage(theta)
        out = torch.empty_like(d)
        with torch.cuda.device(self._dev):
            ws, nb = self.workspace(B, False)
            _lib.check(self._lib.u1ft_phase(self._handle, d.data_ptr(), int(index), out.data_ptr(), B, self.L, ws, nb,
                                            stream_ptr(self._dev)), "u1ft_phase")
        self._count()
        return self._back(out, theta, single)

    def compute_K0_K1(self, theta, index, plaq=None, rect=None):
        """field_trans_opt.py:110-135 -> (K0 [B,4,L,L], K1 [B,8,L,L]).  Only the entries ft_phase reads
        (active sites of the subset, channels of its direction) are defined; everything else is zero.
        plaq / rect are accepted for signature compatibility and recomputed on the device."""
        d, single, B = self._stage(theta)
        K = self._coefficients(d, index)
        mask = torch.zeros_like(K, dtype=torch.bool)
        mu, k = divmod(int(index), 4)
        ch = [2 * mu, 2 * mu + 1] + list(range(4 + 4 * mu, 8 + 4 * mu))
        mask[:, ch, (k >> 1) & 1::2, k & 1::2] = True
        K = torch.where(mask, K, torch.zeros_like(K))
        return self._back(K[:, :4], theta), self._back(K[:, 4:], theta)

    def _coefficients(self, d, index):
        B = d.shape[0]
        K = torch.empty((B, 12, self.L, self.L), dtype=torch.float64, device=self._dev)
        with torch.cuda.device(self._dev):
            ws, nb = self.workspace(B, False)
            _lib.check(self._lib.u1ft_coefficients(self._handle, d.data_ptr(), int(index), K.data_ptr(), B, self.L,
                                                   ws, nb, stream_ptr(self._dev)), "u1ft_coefficients")
        self._count()
        return K

    def inverse(self, theta, max_iter: int = 200, tol: float = 1e-6, _trace=None):
        """field_trans_opt.py:245-282: theta_ori -> theta_new by per-subset fixed-point iteration, same
        stopping rule (global relative change < tol, tested on the host every iteration; a subset
        that does not converge within max_iter is left unchanged, with a warning).  The subset's
        coefficients do not depend on its own links, so the CNN runs once per subset instead of
        once per iteration; the iterates are identical."""
        d, single, B = self._stage(theta)
        cur = d.clone()
        it, nxt = torch.empty_like(cur), torch.empty_like(cur)
        norms = torch.empty(2, dtype=torch.float64, device=self._dev)
        self.inverse_iterations = []
        with torch.cuda.device(self._dev):
            ws, nb = WS.get(self._dev, self._lib.u1ft_inverse_workspace_bytes(B, self.L), "ftinv")
            for index in reversed(range(self.n_subsets)):
                K = self._coefficients(cur, index)
                it.copy_(cur)
                stage_in = cur.clone() if _trace is not None else None
                diff, n_it = float("inf"), 0
                for i in range(max_iter):
                    _lib.check(self._lib.u1ft_inverse_iter(cur.data_ptr(), it.data_ptr(), K.data_ptr(), index,
                                                           nxt.data_ptr(), norms.data_ptr(), B, self.L, ws, nb,
                                                           stream_ptr(self._dev)), "u1ft_inverse_iter")
                    self._count()
                    n = norms.tolist()                        # host sync, as `if diff < tol` in the reference
                    diff = (n[0] ** 0.5) / (n[1] ** 0.5) if n[1] > 0 else float("nan")
                    n_it = i + 1
                    if diff < tol:
                        cur.copy_(nxt)
                        break
                    it, nxt = nxt, it
                if not diff < tol:
                    self.print(f"Warning: Inverse iteration for subset {index} did not converge, final diff = {diff:.2e}")
                self.inverse_iterations.append(n_it)
                if _trace is not None:      # (subset, stage input, iterations; 0 = not converged, stage skipped)
                    _trace.append((index, stage_in, n_it if diff < tol else 0))
        return self._back(cur, theta, single)

    def inverse_field_transformation(self, theta):
        """field_trans_opt.py:280-282."""
        return self.inverse(theta)

    # torch.compile aliases of the reference (field_trans_opt.py:83-108)
    forward_compiled = forward
    field_transformation_compiled = field_transformation
    compute_jac_logdet_compiled = compute_jac_logdet
    ft_phase_compiled = ft_phase

    def compute_action(self, theta, beta):
        """field_trans_opt.py:392-398: -beta * sum cos(P) per chain."""
        d, single, B = self._stage(theta)
        out = torch.empty(B, dtype=torch.float64, device=self._dev)
        with torch.cuda.device(self._dev):
            ws, nb = WS.get(self._dev, self._lib.u1_workspace_bytes(B, self.L))
            _lib.check(self._lib.u1_action(d.data_ptr(), out.data_ptr(), B, self.L, float(beta), ws, nb,
                                           stream_ptr(self._dev)), "u1_action")
        return self._back(out, theta)

    compute_action_compiled = compute_action

    # ------------------------------------------------------------------ HMC_U1_FT pieces
    def new_action(self, theta, beta):
        """hmc_u1_ft.py:101-125: S(F(theta)) - logdet, per chain (0-dim for [2,L,L])."""
        d, single, B = self._stage(theta)
        out = torch.empty(B, dtype=torch.float64, device=self._dev)
        with torch.cuda.device(self._dev):
            ws, nb = self.workspace(B, False)
            _lib.check(self._lib.u1ft_action(self._handle, d.data_ptr(), out.data_ptr(), None, None, B, self.L,
                                             float(beta), ws, nb, stream_ptr(self._dev)), "u1ft_action")
        self._count()
        return self._back(out, theta, single)

    def new_force(self, theta, beta):
        """hmc_u1_ft.py:128-152 (hand-derived reverse pass on the device; no autograd)."""
        d, single, B = self._stage(theta)
        out = torch.empty_like(d)
        with torch.cuda.device(self._dev):
            ws, nb = self.workspace(B, True)
            _lib.check(self._lib.u1ft_force(self._handle, d.data_ptr(), out.data_ptr(), None, B, self.L, float(beta),
                                            ws, nb, stream_ptr(self._dev)), "u1ft_force")
        self._count()
        return self._back(out, theta, single)

    def compute_jac_logdet_autograd(self, theta, eps: float = 1e-5):
        """field_trans_opt.py:385-390 (verification path of if_check_jac): log|det dF/dtheta| of the FIRST
        configuration from the dense Jacobian.  The reference gets the Jacobian from torch.autograd; here it
        comes from central differences of the CUDA forward map (all 2n perturbed configurations in one
        launch, n = 2 L^2; relative error ~eps^2), the determinant from torch.linalg.slogdet on the host.
        Returns a 1-element tensor like the reference."""
        d, single, B = self._stage(theta)
        n = 2 * self.L * self.L
        if self.L > 16:
            raise ValueError("dense Jacobian check supported for lattice_size <= 16")
        base = d[0].reshape(1, n)
        eye = torch.eye(n, dtype=torch.float64, device=self._dev) * eps
        pert = torch.cat([base + eye, base - eye], dim=0).reshape(2 * n, 2, self.L, self.L)
        out = self.forward(pert).reshape(2, n, n)
        jac = ((out[0] - out[1]) / (2 * eps)).T.contiguous()        # jac[i][j] = d out_i / d theta_j
        sign, ld = torch.linalg.slogdet(jac.cpu())
        return (ld if sign > 0 else ld * float("nan")).reshape(1).to(self._dev)

    # ------------------------------------------------------------------ training loss, forward value
    def compute_force(self, theta, beta, transformed: bool = False):
        """field_trans_opt.py:400-441: force of -beta sum cos P (plain, u1_force) or of S_FT (u1ft_force).
        With if_check_jac the analytic log-det of the first configuration is compared with the dense-Jacobian
        one and the reference's message is printed (field_trans_opt.py:420-430)."""
        if transformed:
            if self.if_check_jac:
                ld = self.compute_jac_logdet(theta if theta.dim() == 4 else theta[None])
                ld_dense = self.compute_jac_logdet_autograd(theta)
                diff = (ld_dense[0] - ld[0]) / ld[0]
                if abs(diff.item()) > 1e-4:
                    self.print(f"\nWarning: Jacobian log determinant difference = {diff:.2f}")
                    self.print(">>> Jacobian is not correct!")
                else:
                    self.print(f"\nJacobian log det (manual): {ld[0]:.2e}, (autograd): {ld_dense[0]:.2e}")
                    self.print(">>> Jacobian is all good!")
                self.last_jac_check = float(diff.item())
            return self.new_force(theta, beta)
        d, single, B = self._stage(theta)
        out = torch.empty_like(d)
        with torch.cuda.device(self._dev):
            _lib.check(self._lib.u1_force(d.data_ptr(), out.data_ptr(), B, self.L, float(beta), stream_ptr(self._dev)),
                       "u1_force")
        self._count()
        return self._back(out, theta, single)

    def loss_fn(self, theta_ori, train_beta: Optional[float] = None):
        """field_trans_opt.py:443-462, forward value: theta_new = inverse(theta_ori); d = F_FT(theta_new;
        train_beta) - F_plain(theta_new; 1); loss = sum_{p=2,4,6,8} ||d||_p / vol^(1/p) with the norms over the
        whole batch tensor and vol = L^2.  Returns a 0-dim fp64 tensor on the device (no host sync)."""
        beta = self.train_beta if train_beta is None else train_beta
        if beta is None:
            raise ValueError("loss_fn needs train_beta (set .train_beta or pass train_beta=)")
        th_new = self.inverse(theta_ori)
        d_new, _, B = self._stage(th_new)
        f_ori = self.compute_force(d_new, 1.0)
        f_new = self.compute_force(d_new, float(beta), transformed=True)
        sums = torch.empty(4, dtype=torch.float64, device=self._dev)
        with torch.cuda.device(self._dev):
            ws, nb = WS.get(self._dev, self._lib.u1_diff_power_sums_workspace_bytes(), "loss")
            _lib.check(self._lib.u1_diff_power_sums(f_new.data_ptr(), f_ori.data_ptr(), f_new.numel(), sums.data_ptr(),
                                                    ws, nb, stream_ptr(self._dev)), "u1_diff_power_sums")
        self._count()
        p = torch.tensor([2.0, 4.0, 6.0, 8.0], dtype=torch.float64, device=self._dev)
        vol = float(self.L * self.L)
        return (sums.pow(1.0 / p) / vol ** (1.0 / p)).sum()

    def evaluate_step(self, theta_ori) -> float:
        """field_trans_opt.py:490-505: loss of one evaluation batch as a Python float."""
        return float(self.loss_fn(theta_ori).item())

    # ------------------------------------------------------------------ training (field_trans_opt.py:443-652)
    def loss_and_grad(self, theta_ori, train_beta: Optional[float] = None, check: bool = False):
        """loss_fn (field_trans_opt.py:443-462) and its gradient with respect to every CNN weight, hand-derived on the
        device (csrc/u1_ft_train.cuh) - what the reference obtains from loss.backward() through the unrolled
        fixed-point inverse and compute_force(create_graph=True).  Returns (loss, gradW): a 0-dim fp64 device tensor
        and a flat fp64 device tensor in the layout of the weight blob (per subset, per conv layer: W[o][c][3][3], b[o])."""
        beta = self.train_beta if train_beta is None else train_beta
        if beta is None:
            raise ValueError("loss_and_grad needs train_beta (set .train_beta or pass train_beta=)")
        beta = float(beta)
        lib, dev, L = self._lib, self._dev, self.L
        trace = []
        th_new = self.inverse(theta_ori, _trace=trace)
        d_new, _, B = self._stage(th_new)
        d_new = d_new.contiguous()
        f_ori = self.compute_force(d_new, 1.0)
        f_new = self.new_force(d_new, beta)
        n = f_new.numel()
        sums = torch.empty(4, dtype=torch.float64, device=dev)
        v = torch.empty_like(f_new)
        n_w = int(lib.u1ft_weight_count(self.models[0].n_res_blocks))
        gradW = torch.zeros(n_w, dtype=torch.float64, device=dev)
        Hv, Hp = torch.empty_like(f_new), torch.empty_like(f_new)
        fchk = torch.empty_like(f_new) if check else None
        vol = float(L * L)
        with torch.cuda.device(dev):
            sp = stream_ptr(dev)
            ws, nb = WS.get(dev, lib.u1_diff_power_sums_workspace_bytes(), "loss")
            _lib.check(lib.u1_diff_power_sums(f_new.data_ptr(), f_ori.data_ptr(), n, sums.data_ptr(), ws, nb, sp), "u1_diff_power_sums")
            _lib.check(lib.u1_loss_grad(f_new.data_ptr(), f_ori.data_ptr(), sums.data_ptr(), vol, n, v.data_ptr(), sp), "u1_loss_grad")
            tws, tnb = WS.get(dev, lib.u1ft_train_workspace_bytes(self._handle, B, L), "fttrain")
            # (1) d/dW [v . F_FT(theta_n)] and the Hessian-vector product H_FT v
            _lib.check(lib.u1ft_train_force_grad(self._handle, d_new.data_ptr(), v.data_ptr(), B, L, beta, None, Hv.data_ptr(),
                                                 None if fchk is None else fchk.data_ptr(), gradW.data_ptr(), n_w, tws, tnb, sp),
                       "u1ft_train_force_grad")
            self._count()
            _lib.check(lib.u1_force_hvp(d_new.data_ptr(), v.data_ptr(), Hp.data_ptr(), B, L, 1.0, sp), "u1_force_hvp")
            ubar = Hv - Hp                                   # d loss / d theta_n
            # (2) through the inverse: stages in reverse order of execution (the inverse ran subsets 7..0)
            for index, stage_in, n_it in reversed(trace):
                nxt = torch.empty_like(ubar)
                _lib.check(lib.u1ft_train_inverse_stage_bwd(self._handle, int(index), stage_in.data_ptr(), int(n_it),
                                                            ubar.data_ptr(), nxt.data_ptr(), gradW.data_ptr(), n_w, B, L,
                                                            tws, tnb, sp), "u1ft_train_inverse_stage_bwd")
                self._count()
                ubar = nxt
        p = torch.tensor([2.0, 4.0, 6.0, 8.0], dtype=torch.float64, device=dev)
        loss = (sums.pow(1.0 / p) / vol ** (1.0 / p)).sum()
        if check:       # d(v . F_FT)/dv must reproduce the force: catches a broken reverse sweep without any reference
            self.last_force_check = float((fchk - f_new).abs().max().item() / max(f_new.abs().max().item(), 1e-300))
        self.last_theta_ori_grad = ubar
        return loss, gradW

    def _ensure_optimizers(self):
        """field_trans_opt.py:53-70: one AdamW(lr=1e-3, weight_decay=1e-4) and one ReduceLROnPlateau(0.5, patience 5) per
        subset model (host-side torch.optim on the parameter containers: bookkeeping, not the hot path)."""
        if getattr(self, "optimizers", None) is None:
            self.optimizers = [torch.optim.AdamW(m.parameters(), lr=0.001, weight_decay=1e-4) for m in self.models]
            self.schedulers = [torch.optim.lr_scheduler.ReduceLROnPlateau(o, mode="min", factor=0.5, patience=5)
                               for o in self.optimizers]

    def _assign_grads(self, gradW: torch.Tensor) -> None:
        g = gradW.detach().to("cpu")
        off = 0
        for m in self.models:
            for c in m.conv_layers():
                for prm in (c.weight, c.bias):
                    k = prm.numel()
                    prm.grad = g[off:off + k].reshape(prm.shape).to(prm.dtype).clone()
                    off += k
        assert off == g.numel()

    def train_step(self, theta_ori):
        """field_trans_opt.py:463-480: loss -> gradients -> one AdamW step of all 8 subset models -> new device handle."""
        self._ensure_optimizers()
        loss, gradW = self.loss_and_grad(theta_ori)
        for o in self.optimizers:
            o.zero_grad()
        self._assign_grads(gradW)
        for o in self.optimizers:
            o.step()
        self.sync_weights()
        return float(loss.item())

    def _save_best_model(self, epoch, loss):
        """field_trans_opt.py:619-640: the reference's checkpoint format (loadable by both code bases)."""
        d = {"epoch": epoch, "loss": loss}
        for i, m in enumerate(self.models):
            d[f"model_state_dict_{i}"] = m.state_dict()
        for i, o in enumerate(self.optimizers):
            d[f"optimizer_state_dict_{i}"] = o.state_dict()
        os.makedirs("models", exist_ok=True)
        tag = "" if self.save_tag is None else f"_{self.save_tag}"
        torch.save(d, f"models/best_model_opt_L{self.L}_train_beta{self.train_beta:.1f}{tag}.pt")

    def train(self, train_data, test_data, train_beta, n_epochs=100, batch_size=4):
        """field_trans_opt.py:507-584: epochs of train_step over the training set, evaluate_step over the test set,
        best checkpoint kept, ReduceLROnPlateau on the test loss; the best model is re-loaded at the end.  Returns
        (train_losses, test_losses) (the reference plots them; matplotlib is optional here)."""
        self.train_beta = float(train_beta)
        self._ensure_optimizers()
        tr = torch.utils.data.DataLoader(train_data, batch_size=batch_size, shuffle=True, num_workers=self.num_workers)
        te = torch.utils.data.DataLoader(test_data, batch_size=batch_size, num_workers=self.num_workers)
        train_losses, test_losses, best = [], [], float("inf")
        self.print(f"\n>>> Training the model at beta = {train_beta}\n")
        for epoch in range(n_epochs):
            ep = [self.train_step(b) for b in tr]
            train_losses.append(float(np.mean(ep)))
            ev = [self.evaluate_step(b) for b in te]
            test_losses.append(float(np.mean(ev)))
            self.print(f"Epoch {epoch + 1}/{n_epochs} - Train Loss: {train_losses[-1]:.6f} - Test Loss: {test_losses[-1]:.6f}")
            if test_losses[-1] < best:
                self._save_best_model(epoch, test_losses[-1])
                best = test_losses[-1]
            for sch in self.schedulers:
                sch.step(test_losses[-1])
        self._load_best_model(self.train_beta)
        return train_losses, test_losses

