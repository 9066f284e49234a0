"""box_iou of the metric path (ultralytics.utils.metrics.box_iou, used at reference utils/metrics.py:158,252)."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib


def box_iou(box1: torch.Tensor, box2: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """Pairwise IoU of xyxy boxes: (N,4) x (M,4) -> (N,M) fp32, computed by y5_box_iou on the device."""
    if not (box1.is_cuda and box2.is_cuda):
        raise RuntimeError("y5b200: box_iou runs on CUDA tensors only (no CPU / PyTorch fallback)")
    a = box1.float().contiguous()
    b = box2.float().contiguous()
    out = torch.empty(a.shape[0], b.shape[0], dtype=torch.float32, device=a.device)
    if out.numel():
        _lib.check(_lib.lib().y5_box_iou(a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0], float(eps), out.data_ptr(),
                                         C.c_void_p(_lib.stream_ptr(a.device))), "box_iou")
    return out
