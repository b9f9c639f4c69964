"""ctypes binding of include/u1hmc.h.  There is no CPU fallback: if the CUDA library is
missing or no CUDA device is present, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libu1hmc.so")

_vp, _i, _d, _u64, _sz = C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_size_t

# name -> (restype, argtypes).  Must list every symbol include/u1hmc.h declares.
SIGNATURES = {
    "u1_abi_version": (_i, []),
    "u1_status_string": (C.c_char_p, [_i]),
    "u1_device_info": (_i, [C.POINTER(_i)] * 3),
    "u1_plaquette": (_i, [_vp, _vp, _i, _i, _vp]),
    "u1_rectangle": (_i, [_vp, _vp, _i, _i, _vp]),
    "u1_regularize": (_i, [_vp, _vp, _sz, _vp]),
    "u1_observables": (_i, [_vp, _vp, _vp, _i, _i, _vp, _sz, _vp]),
    "u1_autocorr": (_i, [_vp, _i, _i, _i, _d, _vp, _vp, _vp, _sz, _vp]),
    "u1_workspace_bytes": (_sz, [_i, _i]),
    "u1_action": (_i, [_vp, _vp, _i, _i, _d, _vp, _sz, _vp]),
    "u1_force": (_i, [_vp, _vp, _i, _i, _d, _vp]),
    "u1_leapfrog": (_i, [_vp, _vp, _vp, _vp, _i, _i, _d, _d, _i, _vp, _sz, _vp]),
    "u1_leapfrog_step": (_i, [_vp, _vp, _vp, _vp, _i, _i, _d, _d, _d, _vp]),
    "u1_momenta": (_i, [_vp, _u64, _u64, _u64, _i, _i, _vp]),
    "u1_trajectory": (_i, [_vp, _vp, _vp, _u64, _u64, _u64, _i, _i, _i, _d, _d, _i,
                           _vp, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "u1_last_launch_count": (C.c_longlong, []),
    "u1_fp64_probe": (_i, [_vp, _i, _i, _vp]),
    "u1_slab_leapfrog_step": (_i, [_vp] * 8 + [_i, _i, _d, _d, _d, _vp]),
    "u1_leapfrog_step_rect": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _d, _d, _d, _vp]),
    "u1_leapfrog_steps_rect": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _d, _d, _d, _i, C.POINTER(_i), _vp]),
    "u1_leapfrog_steps_rect3": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _d, _d, _d, _i, C.POINTER(_i), _vp]),
    "u1_slab_drift_sums_ext": (_i, [_vp, _vp, _sz, _vp, _d, _vp, _i, _i, _vp, _sz, _vp]),
    "u1_leapfrog_multi_tile_rows": (_i, [_i]),
    "u1_leapfrog_multi_rows": (_i, [_vp, _vp, _vp, _vp, _i, _i, _d, _d, _d, _i, _i, _i, _vp]),
    "u1_leapfrog_multi_rows2": (_i, [_vp, _vp, _vp, _vp, _i, _i, _d, _d, _d, _i, _i, _i, _i, _i, _vp]),
    "u1_momenta_rows": (_i, [_vp, _u64, _u64, _u64, C.c_longlong, _i, _i, _i, _vp]),
    "u1_momenta_rows_dev": (_i, [_vp, _u64, _u64, _u64, _vp, C.c_longlong, _i, _i, _i, _vp]),
    "u1_drift_wrap": (_i, [_vp, _vp, _vp, _d, _sz, _vp]),
    "u1_metropolis_sums": (_i, [_vp, _vp, _vp, _u64, _u64, _u64, _i, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp]),
    "u1_metropolis_sums_dev": (_i, [_vp, _vp, _vp, _u64, _u64, _u64, _vp, _i, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp]),
    "u1_select_accepted": (_i, [_vp, _vp, _vp, _sz, _i, _vp]),
    "u1_slab_sums": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _sz, _vp]),
    "u1_slab_sums_ext": (_i, [_vp, _vp, C.c_size_t, _vp, _i, _i, _vp, C.c_size_t, _vp]),
    "u1_ipc_handle_bytes": (_sz, []),
    "u1_slab_peer_region_bytes": (_sz, [_i, _i]),
    "u1_peer_alloc": (_i, [C.POINTER(_vp), _sz, _vp]),
    "u1_peer_open": (_i, [_vp, C.POINTER(_vp)]),
    "u1_peer_close": (_i, [_vp]),
    "u1_peer_free": (_i, [_vp]),
    "u1_slab_preload": (_i, []),
    "u1_plain_preload_slab_kernels": (_i, []),
    "u1_slab_peer_status": (_i, [_vp, C.POINTER(C.c_ulonglong)]),
    "u1_slab_exchange": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "u1_slab_metropolis_peer": (_i, [_vp, _vp, _vp, _u64, _u64, _u64, _vp, _d, _d, _i, _i, C.POINTER(_vp),
                                     _vp, _vp, _vp, _vp, _vp, _vp]),
    "u1_slab_drift_wrap": (_i, [_vp, _vp, _sz, _i, _i, _d, _vp]),
    "u1_slab_copy_in": (_i, [_vp, _vp, _sz, _i, _i, _vp]),
    "u1_slab_select_state": (_i, [_vp, _vp, _sz, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "u1_slab_reduce_workspace_bytes": (C.c_size_t, [_i, _i]),
    "u1_momenta_rows_pi2_dev": (_i, [_vp, _u64, _u64, _u64, _vp, C.c_longlong, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "u1_slab_select": (_i, [_vp, _vp, _sz, _i, _i, _vp, _vp]),
    "u1_diff_power_sums_workspace_bytes": (C.c_size_t, []),
    "u1_diff_power_sums": (_i, [_vp, _vp, C.c_size_t, _vp, _vp, C.c_size_t, _vp]),
    "u1ft_create": (_i, [C.POINTER(_vp), _i, _vp, _sz]),
    "u1ft_destroy": (_i, [_vp]),
    "u1ft_weight_count": (_sz, [_i]),
    "u1ft_set_chunk_sites": (C.c_longlong, [C.c_longlong]),
    "u1ft_last_host_launch_count": (C.c_longlong, []),
    "u1ft_workspace_bytes": (_sz, [_vp, _i, _i, _i]),
    "u1ft_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _sz, _vp]),
    "u1ft_phase": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _sz, _vp]),
    "u1ft_coefficients": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _sz, _vp]),
    "u1ft_inverse_workspace_bytes": (_sz, [_i, _i]),
    "u1ft_inverse_iter": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _vp, _sz, _vp]),
    "u1ft_action": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _d, _vp, _sz, _vp]),
    "u1ft_force": (_i, [_vp, _vp, _vp, _vp, _i, _i, _d, _vp, _sz, _vp]),
    "u1ft_leapfrog": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _d, _d, _i, _vp, _sz, _vp]),
    "u1ft_train_workspace_bytes": (_sz, [_vp, _i, _i]),
    "u1_force_hvp": (_i, [_vp, _vp, _vp, _i, _i, _d, _vp]),
    "u1_loss_grad": (_i, [_vp, _vp, _vp, _d, _sz, _vp, _vp]),
    "u1ft_train_force_grad": (_i, [_vp, _vp, _vp, _i, _i, _d, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp]),
    "u1ft_train_inverse_stage_bwd": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _sz, _i, _i, _vp, _sz, _vp]),
    "u1ft_trajectory": (_i, [_vp, _vp, _vp, _vp, _u64, _u64, _u64, _i, _i, _i, _d, _d, _i,
                             _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
}

_lock = threading.Lock()
_lib = None


class U1Error(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen hmc_ft_b200/libu1hmc.so and type every entry point.  Raises if absent."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise U1Error(
                    f"{LIB_PATH} is missing: the sm_100a CUDA library is the only implementation "
                    "(no CPU fallback).  Build it with `python -m hmc_ft_b200.build`.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)          # AttributeError if the ABI and the header diverge
                fn.restype, fn.argtypes = res, args
            if lib.u1_abi_version() != 1:
                raise U1Error("libu1hmc.so ABI version mismatch")
            _lib = lib
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().u1_status_string(status).decode()
        raise U1Error(f"{what or 'u1hmc call'} failed with status {status}: {msg}")


def last_launch_count() -> int:
    return int(load().u1_last_launch_count())
