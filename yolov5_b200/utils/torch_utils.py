"""Hot-path edges of the reference's utils/torch_utils.py: BatchNorm folding and the DDP wrapper."""
from __future__ import annotations

import os

import torch
from torch import nn


def fuse_conv_and_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d) -> nn.Conv2d:
    """Conv2d + eval BatchNorm2d -> one Conv2d with bias; same algebra as reference utils/torch_utils.py:224-254
    (W' = diag(g / sqrt(var + eps)) W,  b' = beta + (b - mean) g / sqrt(var + eps))."""
    fused = nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, conv.dilation,
                      conv.groups, bias=True).requires_grad_(False).to(conv.weight.device, conv.weight.dtype)
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    fused.weight.copy_(conv.weight * scale.view(-1, 1, 1, 1))
    b0 = conv.bias if conv.bias is not None else torch.zeros_like(scale)
    fused.bias.copy_(bn.bias + (b0 - bn.running_mean) * scale)
    return fused


def smart_DDP(model):
    """The path's one collective: gradient all-reduce through DistributedDataParallel over NCCL
    (reference utils/torch_utils.py:61-70)."""
    from torch.nn.parallel import DistributedDataParallel as DDP

    local_rank = int(os.getenv("LOCAL_RANK", -1))
    return DDP(model, device_ids=[local_rank], output_device=local_rank, static_graph=True)


def de_parallel(model):
    return model.module if isinstance(model, (nn.parallel.DataParallel, nn.parallel.DistributedDataParallel)) else model
