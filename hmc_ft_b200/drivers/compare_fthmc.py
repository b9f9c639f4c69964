"""python -m hmc_ft_b200.drivers.compare_fthmc — flags of 2d_u1_cluster/compare_fthmc.py:13-24."""
import argparse
import time

from .. import HMC_U1_FT, FieldTransformation
from ..utils import set_seed
from .common import hmc_summary


def main(argv=None):
    p = argparse.ArgumentParser(description="Parameters for Comparison")
    p.add_argument("--lattice_size", type=int, default=16)
    p.add_argument("--n_configs", type=int, default=2048)
    p.add_argument("--beta", type=float, default=6)
    p.add_argument("--train_beta", type=float, default=6)
    p.add_argument("--step_size", type=float, default=0.1)
    p.add_argument("--ft_step_size", type=float, default=0.1)
    p.add_argument("--max_lag", type=int, default=200)
    p.add_argument("--rand_seed", type=int, default=1331)
    p.add_argument("--model_tag", type=str, default="simple")
    p.add_argument("--save_tag", type=str, default=None)
    p.add_argument("--device", type=str, default="cuda")
    p.add_argument("--n_chains", type=int, default=None, help="extension: batch of independent chains")
    p.add_argument("--weights", type=str, default=None, help="extension: checkpoint (.pt) or fixture (.npz) path")
    p.add_argument("--n_therm", type=int, default=200, help="extension (compare_fthmc.py:49 hard-codes 200)")
    a = p.parse_args(argv)
    print("=" * 60 + "\n>>> Arguments:")
    for k, v in vars(a).items():
        print(f"{k}: {v}")
    print("=" * 60)
    set_seed(a.rand_seed)
    n_steps, store_interval = 50, 1                              # compare_fthmc.py:50-53
    print(">>> Neural Network Field Transformation HMC Simulation: ")
    nn_ft = FieldTransformation(a.lattice_size, device=a.device, n_subsets=8, if_check_jac=False, num_workers=0,
                                identity_init=True, model_tag=a.model_tag, save_tag=a.save_tag)
    t0 = time.time()
    print(">>> Loading trained model")
    if a.weights and a.weights.endswith(".npz"):
        nn_ft.load_weights_npz(a.weights)
    elif a.weights:
        nn_ft.load_checkpoint(a.weights)
    else:
        nn_ft._load_best_model(a.train_beta)
    t1 = time.time()
    print(f">>> Model loaded successfully in {t1 - t0:.2f} seconds")
    hmc = HMC_U1_FT(a.lattice_size, a.beta, a.n_therm, n_steps, a.ft_step_size,
                    field_transformation=nn_ft.field_transformation_compiled,
                    compute_jac_logdet=nn_ft.compute_jac_logdet_compiled, device=a.device,
                    if_tune_step_size=False, n_chains=a.n_chains, seed=a.rand_seed)
    theta, therm_plaq, therm_acc = hmc.thermalize()
    t2 = time.time()
    print(f">>> Thermalization with field transformation completed in {t2 - t1:.2f} seconds")
    _, plaq, acc, topo, ham = hmc.run(store_interval * a.n_configs, theta, store_interval, save_config=False)
    t3 = time.time()
    print(f">>> Simulation with field transformation completed in {t3 - t2:.2f} seconds")
    print(f">>> Total time (Field Transformation HMC): {t3 - t0:.2f} seconds")
    hmc_summary(a.beta, a.max_lag, a.lattice_size ** 2, therm_plaq, plaq, topo, ham, therm_acc, acc)
    return plaq, acc, topo


if __name__ == "__main__":
    main()
