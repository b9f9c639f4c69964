"""Generate tests/golden/* by running the REAL reference (read-only at /root/reference)
in fp64 on CPU.  Runs only in the build container (the GPU box has no /root/reference);
the produced fixtures are committed.  Recipe follows SURVEY.md §8(c).

    TORCHDYNAMO_DISABLE=1 python tools/gen_golden.py
"""
import json
import math
import os
import re
import sys
import types

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
import numpy as np
import torch

REF = "/root/reference"
CL = os.path.join(REF, "2d_u1_cluster")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# matplotlib is not installed; utils.py:1 imports it at module top.
stub = types.ModuleType("matplotlib"); stub.pyplot = types.ModuleType("matplotlib.pyplot")
sys.modules["matplotlib"] = stub; sys.modules["matplotlib.pyplot"] = stub.pyplot
sys.path.insert(0, CL)

from hmc_u1 import HMC_U1                      # noqa: E402
from hmc_u1_ft import HMC_U1_FT                # noqa: E402
from field_trans_opt import FieldTransformation  # noqa: E402
import utils as ref_utils                      # noqa: E402

STATE_KEYS = {
    "simple": ["conv1", "conv2"],
    "rnet1": ["initial_conv", "res_blocks.0.conv1", "res_blocks.0.conv2", "final_conv"],
    "rnet3": ["initial_conv"] + [f"res_blocks.{i}.conv{j}" for i in range(3) for j in (1, 2)] + ["final_conv"],
}
CKPT = {
    ("rnet3", 6.0): "best_model_opt_L64_train_beta6.0_rnet3.pt",
    ("rnet3", 3.0): "best_model_opt_L64_train_beta3.0_rnet3.pt",
    ("rnet1", 6.0): "best_model_opt_L64_train_beta6.0_rnet1.pt",
    ("simple", 6.0): "best_model_L64_train_beta6.0_seed2008.pt",
}


def seeded_inputs(L, seed=12345):
    g = torch.Generator().manual_seed(seed)
    theta = (torch.rand(2, L, L, generator=g, dtype=torch.float64) * 2 - 1) * math.pi
    pi = torch.randn(2, L, L, generator=g, dtype=torch.float64)
    return theta, pi


def plain_case(L, beta, n_steps, dt, tag):
    hmc = HMC_U1(L, beta, 0, n_steps, dt, device="cpu", if_tune_step_size=False)
    torch.set_default_dtype(torch.float64)
    theta, pi = seeded_inputs(L)
    S = hmc.action(theta)
    Fo = hmc.force(theta)
    th2, pi2 = hmc.leapfrog(theta.clone(), pi.clone())
    S2 = hmc.action(th2)
    dH = (S2 + 0.5 * (pi2 ** 2).sum()) - (S + 0.5 * (pi ** 2).sum())
    np.savez_compressed(
        os.path.join(OUT, f"plain_{tag}.npz"),
        L=L, beta=beta, n_steps=n_steps, dt=dt, theta=theta.numpy(), pi=pi.numpy(),
        action=S.item(), force=Fo.numpy(), theta_out=th2.numpy(), pi_out=pi2.numpy(),
        action_out=S2.item(), dH=dH.item(),
        plaq=ref_utils.plaq_mean_from_field(theta).item(), topo=ref_utils.topo_from_field(theta).item(),
        plaq_field=ref_utils.plaq_from_field(theta).numpy(), regularized=ref_utils.regularize(theta * 3.7).numpy(),
    )
    print(f"plain {tag}: S={S.item():.15e} |F|={Fo.norm().item():.15e} dH={dH.item():.13e} topo={ref_utils.topo_from_field(theta).item()}")


def load_ft(L, arch, beta_train):
    ft = FieldTransformation(L, device="cpu", model_tag=arch, save_tag=arch)
    ck = torch.load(os.path.join(CL, "models", CKPT[(arch, beta_train)]), map_location="cpu", weights_only=False)
    sds = []
    for i, m in enumerate(ft.models):
        sd = {k.replace("module.", ""): v for k, v in ck[f"model_state_dict_{i}"].items()}
        m.load_state_dict(sd); m.double(); m.eval()
        sds.append(sd)
    return ft, sds


def dump_weights(arch, beta_train):
    ck = torch.load(os.path.join(CL, "models", CKPT[(arch, beta_train)]), map_location="cpu", weights_only=False)
    d = {"arch": np.array(arch)}
    for s in range(8):
        sd = {k.replace("module.", ""): v for k, v in ck[f"model_state_dict_{s}"].items()}
        for k in STATE_KEYS[arch]:
            d[f"s{s}.{k}.weight"] = sd[k + ".weight"].numpy()
            d[f"s{s}.{k}.bias"] = sd[k + ".bias"].numpy()
    np.savez_compressed(os.path.join(OUT, f"weights_{arch}_beta{beta_train:.1f}.npz"), **d)


def ft_case(L, arch, beta_train, beta, seed, amp, tag, dense_check=False):
    ft, _ = load_ft(L, arch, beta_train)
    hmc = HMC_U1_FT(L, beta, 0, 5, 0.05, ft.field_transformation, ft.compute_jac_logdet,
                    device="cpu", if_tune_step_size=False)
    torch.set_default_dtype(torch.float64)
    g = torch.Generator().manual_seed(seed)
    theta = amp * torch.randn(2, L, L, generator=g, dtype=torch.float64)
    pi = torch.randn(2, L, L, generator=g, dtype=torch.float64)
    ori = ft.field_transformation(theta)
    ld = ft.compute_jac_logdet(theta[None])[0]
    S = hmc.new_action(theta)
    Fo = hmc.new_force(theta)
    phase0 = ft.ft_phase(theta[None], 0)[0]
    phase5 = ft.ft_phase(theta[None], 5)[0]
    plaq = ref_utils.plaq_from_field_batch(theta[None]); rect = ref_utils.rect_from_field_batch(theta[None])
    K0, K1 = ft.compute_K0_K1(theta[None], 3, plaq, rect)
    th2, pi2 = hmc.leapfrog(theta.clone(), pi.clone())
    S2 = hmc.new_action(th2)
    dH = (S2 + 0.5 * (pi2 ** 2).sum()) - (S + 0.5 * (pi ** 2).sum())
    extra = {}
    if dense_check:
        extra["logdet_dense"] = ft.compute_jac_logdet_autograd(theta[None]).item()
    np.savez_compressed(
        os.path.join(OUT, f"ft_{tag}.npz"),
        L=L, beta=beta, arch=np.array(arch), beta_train=beta_train, theta=theta.numpy(), pi=pi.numpy(),
        theta_ori=ori.detach().numpy(), logdet=ld.item(), new_action=S.item(), new_force=Fo.numpy(),
        phase0=phase0.detach().numpy(), phase5=phase5.detach().numpy(),
        K0_3=K0[0].detach().numpy(), K1_3=K1[0].detach().numpy(),
        n_steps=5, dt=0.05, theta_out=th2.numpy(), pi_out=pi2.numpy(), dH=dH.item(), **extra,
    )
    print(f"ft {tag}: S_FT={S.item():.12e} logdet={ld.item():.12e} |F|={Fo.norm().item():.12e} dH={dH.item():.6e}", extra)


def test_out_head(n=48):
    """First n trajectories of the reference's fp64 golden log tests/test.out."""
    rows = []
    pat = re.compile(r"Traj:\s+(\d+)\s+(ACCEPT|REJECT):\s+dH:\s*(\S+)\s+exp\(-dH\):\s*(\S+)\s+plaq:\s*(\S+)\s+topo:\s*(\S+)")
    with open(os.path.join(REF, "tests", "test.out")) as f:
        for line in f:
            m = pat.search(line)
            if m:
                rows.append({"traj": int(m.group(1)), "accept": m.group(2) == "ACCEPT", "dH": float(m.group(3)),
                             "exp_mdH": float(m.group(4)), "plaq": float(m.group(5)), "topo": float(m.group(6))})
            if len(rows) >= n:
                break
    with open(os.path.join(OUT, "test_out_head.json"), "w") as f:
        json.dump({"source": "tests/test.out (tests/hmc_2DU1.py: L=64 beta=6 tau=2 nstep=50 seed=1984, fp64, cold start)",
                   "L": 64, "beta": 6.0, "tau": 2.0, "nstep": 50, "seed": 1984, "rows": rows}, f, indent=0)
    print("test.out rows:", len(rows))


def realistic_configs(n=32):
    a = np.load(os.path.join(REF, "2d_u1_rep", "dump", "theta_ori_L8_beta3.npy"))
    np.save(os.path.join(OUT, "theta_ori_L8_beta3_head.npy"), a[:n])
    print("realistic configs:", a[:n].shape, a.dtype)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    plain_case(16, 2.0, 10, 0.1, "L16_b2")
    plain_case(64, 6.0, 50, 0.06, "L64_b6")
    plain_case(6, 1.5, 3, 0.2, "L6_b1p5")
    for key in CKPT:
        dump_weights(*key)
    ft_case(8, "rnet3", 6.0, 6.0, 3, 0.3, "rnet3_L8", dense_check=True)
    ft_case(8, "simple", 6.0, 6.0, 4, 0.5, "simple_L8", dense_check=True)
    ft_case(8, "rnet1", 6.0, 6.0, 5, 0.5, "rnet1_L8", dense_check=True)
    ft_case(16, "rnet3", 3.0, 3.0, 6, 0.6, "rnet3_L16_b3")
    ft_case(32, "rnet3", 3.0, 3.0, 7, 0.5, "rnet3_L32_b3")
    ft_case(64, "rnet3", 6.0, 6.0, 3, 0.3, "rnet3_L64_b6")
    ft_case(64, "simple", 6.0, 6.0, 8, 0.3, "simple_L64_b6")
    ft_case(64, "rnet1", 6.0, 6.0, 9, 0.3, "rnet1_L64_b6")
    test_out_head()
    realistic_configs()
