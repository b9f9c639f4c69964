"""CPU-side checks: the C-ABI library builds, loads and exports every symbol include/u1hmc.h
declares, the ctypes table matches the header, and the product path fails loudly without a GPU."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from hmc_ft_b200.build import build
    build()
    from hmc_ft_b200 import _lib
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "u1hmc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(u1(?:ft)?_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from hmc_ft_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/u1hmc.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, set(_lib.SIGNATURES) ^ set(syms)
    assert lib.u1_abi_version() == 1
    assert b"workspace" in lib.u1_status_string(-3)


def test_argument_validation_without_gpu(lib):
    # argument errors are detected before any CUDA call
    assert lib.u1_force(None, None, 1, 8, 1.0, None) == -1
    assert lib.u1_workspace_bytes(4096, 64) > 4 * 4096 * 2 * 64 * 64 * 8
    assert lib.u1_workspace_bytes(0, 64) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    import hmc_ft_b200 as hb
    with pytest.raises(hb.U1Error):
        hb.HMC_U1(16, 2.0, 0, 10, 0.1)
    with pytest.raises(hb.U1Error):
        from hmc_ft_b200 import utils
        utils.regularize(torch.zeros(2, 4, 4))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "hmc_ft_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and False, f"{f} mentions the oracle"


def test_host_masks_match_oracle():
    from hmc_ft_b200 import utils as U
    from oracle import u1_oracle as O
    for L in (4, 8):
        for k in range(8):
            assert torch.equal(U.get_field_mask(k, 2, L)[1], O.field_mask(k, L))
            assert torch.equal(U.get_plaq_mask(k, 2, L)[0], O.plaq_mask(k, L))
            assert torch.equal(U.get_rect_mask(k, 2, L)[1], O.rect_mask(k, L))
