"""Host-side plumbing shared by the reference-facing classes: device selection, fp64 staging of
caller tensors (host or device), cached workspaces and the current CUDA stream.  torch is used
only for memory, streams and torch.distributed — all arithmetic is in libu1hmc.so."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.U1Error("hmc_ft_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.type != "cuda":
        dev = torch.device("cuda", torch.cuda.current_device())
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def stream_ptr(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def to_dev(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    """Contiguous fp64 copy/view of `t` on the compute device (H2D copy if `t` is on the host)."""
    if t.device != dev:
        if t.device.type == "cpu" and not t.is_pinned() and t.numel() >= (1 << 16):
            t = t.pin_memory()
        t = t.to(dev, non_blocking=True)
    if t.dtype != torch.float64:
        t = t.to(torch.float64)
    return t.contiguous()


def as_batch(theta: torch.Tensor) -> Tuple[torch.Tensor, bool]:
    """[2,L,L] -> ([1,2,L,L], True); [B,2,L,L] -> (same, False)."""
    if theta.dim() == 3:
        return theta.unsqueeze(0), True
    if theta.dim() == 4:
        return theta, False
    raise ValueError(f"expected theta of shape [2,L,L] or [B,2,L,L], got {tuple(theta.shape)}")


def check_field(theta: torch.Tensor) -> Tuple[int, int]:
    B, two, L, L2 = theta.shape
    if two != 2 or L != L2:
        raise ValueError(f"expected [B,2,L,L], got {tuple(theta.shape)}")
    return B, L


class Workspace:
    """Grow-only device scratch buffers keyed by (device, tag)."""

    def __init__(self):
        self._bufs: Dict[Tuple[torch.device, str], torch.Tensor] = {}

    def get(self, dev: torch.device, nbytes: int, tag: str = "main") -> Tuple[int, int]:
        key = (dev, tag)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            self._bufs[key] = buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
        return buf.data_ptr(), buf.numel()


WS = Workspace()


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()
