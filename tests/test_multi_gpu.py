"""Domain-decomposed plain HMC (BASELINE configs[4] path) against the undecomposed chain."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "slab_worker.py")


@pytest.mark.parametrize("L,halo,transport", [(64, 1, "peer"), (128, 3, "peer"), (96, 8, "peer"), (128, 3, "nccl")])
def test_slab_world1_matches_plain_chain(L, halo, transport):
    out = subprocess.run([sys.executable, WORKER, str(L), str(halo), transport], capture_output=True, text=True, timeout=600)
    assert "SLAB_OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("L,halo,virt", [(128, 4, 2), (256, 4, 4), (192, 2, 3), (512, 16, 8)])
def test_slab_virtual_ranks_on_one_gpu(L, halo, virt):
    """The NVLink peer path (push kernels, epoch flags, cross-rank Metropolis, CUDA-graph replay) with several
    ranks inside one process on ONE GPU: same code as the multi-GPU run, checkable where only one GPU exists."""
    out = subprocess.run([sys.executable, WORKER, str(L), str(halo), "peer", str(virt)], capture_output=True, text=True,
                         timeout=600)
    assert "SLAB_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


@pytest.mark.parametrize("world,transport", [(2, "peer"), (4, "peer"), (8, "peer"), (2, "nccl"), (8, "nccl")])
def test_slab_multi_gpu_matches_plain_chain(world, transport):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29631", WORKER, "256", "4", transport],
                         capture_output=True, text=True, timeout=900)
    assert "SLAB_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


FT_WORKER = os.path.join(ROOT, "tests", "ft_shard_worker.py")


def test_ft_chain_sharding_world1_observables():
    out = subprocess.run([sys.executable, FT_WORKER], capture_output=True, text=True, timeout=900)
    assert "FTSHARD_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_ft_chain_sharding_multi_gpu_observables(world):
    """C4 'with topological charge and autocorrelation observables' across ranks: gathered Q histories and
    Gamma(delta) equal the unsharded run bit for bit (Philox keyed by global chain id)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29641", FT_WORKER],
                         capture_output=True, text=True, timeout=900)
    assert "FTSHARD_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
