"""Weight containers with the reference's module/parameter names (2d_u1_cluster/cnn_model_opt.py)
so checkpoints written by the reference load unchanged (field_trans_opt.py:654-692).  These
modules only HOLD parameters: the convolutions run in libu1hmc.so (csrc/u1_ft.cu k_conv)."""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn

WIDTH, IN_CH = 12, 6


def _conv(cin: int, cout: int) -> nn.Conv2d:
    return nn.Conv2d(cin, cout, (3, 3), padding="same", padding_mode="circular")


class ResBlock(nn.Module):
    """cnn_model_opt.py:54-87."""

    def __init__(self, channels: int = WIDTH):
        super().__init__()
        self.conv1 = _conv(channels, channels)
        self.conv2 = _conv(channels, channels)


class _JointCNN(nn.Module):
    n_res_blocks = 0

    def conv_layers(self) -> List[nn.Conv2d]:
        raise NotImplementedError

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("jointCNN_* modules are parameter containers; evaluation happens in the CUDA library "
                           "through FieldTransformation")

    def packed_weights(self) -> torch.Tensor:
        """fp64 blob in the layout u1ft_create expects: per layer W[o][c][3][3] then b[o]."""
        parts = []
        for c in self.conv_layers():
            parts += [c.weight.detach().to("cpu", torch.float64).reshape(-1), c.bias.detach().to("cpu", torch.float64).reshape(-1)]
        return torch.cat(parts)


class jointCNN_simple(_JointCNN):
    """cnn_model_opt.py:6-52: conv1 (6->12), conv2 (12->12)."""
    n_res_blocks = 0

    def __init__(self):
        super().__init__()
        self.conv1 = _conv(IN_CH, WIDTH)
        self.conv2 = _conv(WIDTH, WIDTH)

    def conv_layers(self):
        return [self.conv1, self.conv2]


class _jointCNN_rnet(_JointCNN):
    """cnn_model_opt.py:131-247: initial_conv, res_blocks[n], final_conv."""

    def __init__(self):
        super().__init__()
        self.initial_conv = _conv(IN_CH, WIDTH)
        self.res_blocks = nn.ModuleList([ResBlock(WIDTH) for _ in range(self.n_res_blocks)])
        self.final_conv = _conv(WIDTH, WIDTH)

    def conv_layers(self):
        out = [self.initial_conv]
        for rb in self.res_blocks:
            out += [rb.conv1, rb.conv2]
        return out + [self.final_conv]


class jointCNN_rnet1(_jointCNN_rnet):
    n_res_blocks = 1


class jointCNN_rnet3(_jointCNN_rnet):
    n_res_blocks = 3


def choose_cnn_model(model_tag: str):
    """cnn_model_opt.py:311-324 ('rnet_norm' ships no weights and is out of scope)."""
    try:
        return {"simple": jointCNN_simple, "rnet1": jointCNN_rnet1, "rnet3": jointCNN_rnet3}[model_tag]
    except KeyError:
        raise ValueError(f"Invalid model tag: {model_tag}") from None
