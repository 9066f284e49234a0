"""Checkpoint loading with the reference's entry point (reference models/experimental.py:46-101): ``attempt_load``
unpickles whole-module checkpoints (``{"model": DetectionModel, "ema": ...}``), moves them to the device in fp32, fuses
BatchNorm and returns the model in eval mode; several weights give an ``Ensemble`` whose outputs are concatenated for NMS.

Reference checkpoints pickle classes by their module path (``models.yolo.DetectionModel``, ``models.common.Conv`` ...);
``yolov5_b200.compat.install()`` registers those names as aliases of this package so such files load into the engine's
classes (same attribute names, same state_dict keys).
"""
from __future__ import annotations

import torch
from torch import nn


class Ensemble(nn.ModuleList):
    """NMS ensemble: every member's decoded predictions concatenated along the row dimension."""

    def forward(self, x, augment=False, profile=False, visualize=False):
        y = [m(x, augment, profile, visualize)[0] for m in self]
        return torch.cat(y, 1), None


def attempt_load(weights, device=None, inplace=True, fuse=True):
    from .. import compat
    from .yolo import Detect, DetectionModel

    compat.install()  # reference-pickled class paths resolve to this package
    model = Ensemble()
    for w in weights if isinstance(weights, list) else [weights]:
        ckpt = torch.load(str(w), map_location="cpu", weights_only=False)
        m = (ckpt.get("ema") or ckpt["model"]).to(device).float()
        if not hasattr(m, "stride"):
            m.stride = torch.tensor([32.0])
        if hasattr(m, "names") and isinstance(m.names, (list, tuple)):
            m.names = dict(enumerate(m.names))
        model.append(m.fuse().eval() if fuse and hasattr(m, "fuse") else m.eval())
    for m in model.modules():
        if isinstance(m, (nn.SiLU, Detect, DetectionModel)):
            m.inplace = inplace
        elif isinstance(m, nn.Upsample) and not hasattr(m, "recompute_scale_factor"):
            m.recompute_scale_factor = None
    if len(model) == 1:
        return model[-1]
    for k in ("names", "nc", "yaml"):
        setattr(model, k, getattr(model[0], k))
    model.stride = model[int(torch.argmax(torch.tensor([float(m.stride.max()) for m in model])))].stride
    assert all(model[0].nc == m.nc for m in model), f"Models have different class counts: {[m.nc for m in model]}"
    return model
