"""Golden vectors for two rows the round-1 fixtures did not cover, produced by running the REAL
reference (read-only at /root/reference) in fp64 on the CPU of the build container:

  rect.npz  utils.py:118-141  plaq_from_field_batch / rect_from_field_batch on random batches, L = 6, 8, 16
  tune.npz  hmc_u1.py:99-161  HMC_U1.tune_step_size: the (dt, acceptance) history of the binary search and
            the tuned dt, together with every (pi, u) draw the search consumed from torch's generator
            (randn_like(theta) then rand([]) per metropolis_step, hmc_u1.py:83,94) so that the oracle and
            the CUDA path can replay the same search with injected draws.

    TORCHDYNAMO_DISABLE=1 python tools/gen_golden_rect_tune.py
"""
import contextlib
import io
import math
import os
import re
import sys
import types

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
import numpy as np
import torch

CL = "/root/reference/2d_u1_cluster"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
stub = types.ModuleType("matplotlib"); stub.pyplot = types.ModuleType("matplotlib.pyplot")
sys.modules["matplotlib"] = stub; sys.modules["matplotlib.pyplot"] = stub.pyplot
sys.path.insert(0, CL)

from hmc_u1 import HMC_U1          # noqa: E402
import utils as ref_utils          # noqa: E402


def rect_cases():
    out = {}
    for L in (6, 8, 16):
        g = torch.Generator().manual_seed(4242 + L)
        th = (torch.rand(3, 2, L, L, generator=g, dtype=torch.float64) * 2 - 1) * math.pi
        out[f"theta_L{L}"] = th.numpy()
        out[f"plaq_L{L}"] = ref_utils.plaq_from_field_batch(th).numpy()
        out[f"rect_L{L}"] = ref_utils.rect_from_field_batch(th).numpy()
    np.savez_compressed(os.path.join(OUT, "rect.npz"), **out)
    print("rect.npz", {k: v.shape for k, v in out.items()})


def tune_case(L=8, beta=2.0, n_steps=10, n_tune=24, target=0.65, tol=0.05, dt0=0.9, max_attempts=6, seed=777):
    hmc = HMC_U1(L, beta, 0, n_steps, dt0, device="cpu", if_tune_step_size=True)
    torch.set_default_dtype(torch.float64)          # after the ctor (it forces fp32, hmc_u1.py:49)
    g = torch.Generator().manual_seed(99)
    theta = (torch.rand(2, L, L, generator=g, dtype=torch.float64) * 2 - 1) * 0.8
    torch.manual_seed(seed)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(io.StringIO()):
        hmc.tune_step_size(n_tune_steps=n_tune, target_rate=target, target_tolerance=tol,
                           initial_step_size=dt0, max_attempts=max_attempts, theta=theta)
    hist = [(float(a), float(b) / 100.0) for a, b in
            re.findall(r"Step size: ([0-9.]+), Acceptance rate: ([0-9.]+)%", buf.getvalue())]
    # replay the generator: the draws the search consumed, in order
    torch.manual_seed(seed)
    n = len(hist) * n_tune
    pis, us = [], []
    for _ in range(n):
        pis.append(torch.randn_like(theta))
        us.append(torch.rand([]))
    np.savez_compressed(os.path.join(OUT, "tune.npz"), theta=theta.numpy(), pi=torch.stack(pis).numpy(),
                        u=torch.stack(us).numpy(), hist_dt=np.array([h[0] for h in hist]),
                        hist_rate=np.array([h[1] for h in hist]), tuned_dt=np.float64(hmc.dt),
                        params=np.array([L, beta, n_steps, n_tune, target, tol, dt0, max_attempts], dtype=np.float64))
    print("tune.npz history", hist, "tuned dt", hmc.dt, "draws", n)
    torch.set_default_dtype(torch.float32)


if __name__ == "__main__":
    rect_cases()
    tune_case()
