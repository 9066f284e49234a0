"""Import shim that lets the UNMODIFIED reference (/root/reference) be imported in the build container.

The reference imports the third-party ``ultralytics`` package (requirements.txt:16) plus
``matplotlib``/``seaborn``; none is installed here and there is no network.  This module installs a
meta-path finder that serves those package names with stand-in modules.  Names on the numeric path
get real implementations restated from the public package's documented behaviour (SURVEY.md
Appendix C -- "parity unpinned" for those pieces, the reference ships no test for them); every
other name resolves to an inert stub so module import succeeds.

Only tests/golden/make_golden.py uses this file, and only in the build container:
/root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import contextlib
import importlib.abc
import importlib.machinery
import logging
import math
import sys
import time
import types

import torch
from torch import nn

REFERENCE_ROOT = "/root/reference"


class _Inert:
    """Callable / decorator / context-manager that does nothing; attribute access yields more of the same."""

    def __init__(self, name="stub"):
        self._name = name

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k and not isinstance(a[0], _Inert):
            return a[0]  # used as a bare decorator
        return _Inert(self._name + "()")

    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        return _Inert(f"{self._name}.{item}")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def __iter__(self):
        return iter(())

    def __bool__(self):
        return False


# ----------------------------------------------------------------------------------------------------------------------
# numeric helpers (Appendix C of SURVEY.md)
# ----------------------------------------------------------------------------------------------------------------------
def make_divisible(x, divisor):
    if isinstance(divisor, torch.Tensor):
        divisor = int(divisor.max())
    return math.ceil(x / divisor) * divisor


def initialize_weights(model):
    for m in model.modules():
        t = type(m)
        if t is nn.BatchNorm2d:
            m.eps = 1e-3
            m.momentum = 0.03
        elif t in {nn.Hardswish, nn.LeakyReLU, nn.ReLU, nn.ReLU6, nn.SiLU}:
            m.inplace = True


def xywh2xyxy(x):
    y = torch.empty_like(x) if isinstance(x, torch.Tensor) else x.copy()
    xy = x[..., :2]
    wh = x[..., 2:] / 2
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y


def clip_boxes(boxes, shape):
    if isinstance(boxes, torch.Tensor):
        boxes[..., 0].clamp_(0, shape[1])
        boxes[..., 1].clamp_(0, shape[0])
        boxes[..., 2].clamp_(0, shape[1])
        boxes[..., 3].clamp_(0, shape[0])
    else:
        boxes[..., [0, 2]] = boxes[..., [0, 2]].clip(0, shape[1])
        boxes[..., [1, 3]] = boxes[..., [1, 3]].clip(0, shape[0])
    return boxes


def box_iou(box1, box2, eps=1e-7):
    (a1, a2), (b1, b2) = box1.float().unsqueeze(1).chunk(2, 2), box2.float().unsqueeze(0).chunk(2, 2)
    inter = (torch.min(a2, b2) - torch.max(a1, b1)).clamp_(0).prod(2)
    return inter / ((a2 - a1).prod(2) + (b2 - b1).prod(2) - inter + eps)


def bbox_iou(box1, box2, xywh=True, GIoU=False, DIoU=False, CIoU=False, eps=1e-7):
    if xywh:
        (x1, y1, w1, h1), (x2, y2, w2, h2) = box1.chunk(4, -1), box2.chunk(4, -1)
        w1_, h1_, w2_, h2_ = w1 / 2, h1 / 2, w2 / 2, h2 / 2
        b1_x1, b1_x2, b1_y1, b1_y2 = x1 - w1_, x1 + w1_, y1 - h1_, y1 + h1_
        b2_x1, b2_x2, b2_y1, b2_y2 = x2 - w2_, x2 + w2_, y2 - h2_, y2 + h2_
    else:
        b1_x1, b1_y1, b1_x2, b1_y2 = box1.chunk(4, -1)
        b2_x1, b2_y1, b2_x2, b2_y2 = box2.chunk(4, -1)
        w1, h1 = b1_x2 - b1_x1, b1_y2 - b1_y1 + eps
        w2, h2 = b2_x2 - b2_x1, b2_y2 - b2_y1 + eps
    inter = (b1_x2.minimum(b2_x2) - b1_x1.maximum(b2_x1)).clamp_(0) * (
        b1_y2.minimum(b2_y2) - b1_y1.maximum(b2_y1)
    ).clamp_(0)
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    if CIoU or DIoU or GIoU:
        cw = b1_x2.maximum(b2_x2) - b1_x1.minimum(b2_x1)
        ch = b1_y2.maximum(b2_y2) - b1_y1.minimum(b2_y1)
        if CIoU or DIoU:
            c2 = cw.pow(2) + ch.pow(2) + eps
            rho2 = ((b2_x1 + b2_x2 - b1_x1 - b1_x2).pow(2) + (b2_y1 + b2_y2 - b1_y1 - b1_y2).pow(2)) / 4
            if CIoU:
                v = (4 / math.pi**2) * ((w2 / h2).atan() - (w1 / h1).atan()).pow(2)
                with torch.no_grad():
                    alpha = v / (v - iou + (1 + eps))
                return iou - (rho2 / c2 + v * alpha)
            return iou - rho2 / c2
        c_area = cw * ch + eps
        return iou - (c_area - union) / c_area
    return iou


def smooth_bce(eps=0.1):
    return 1.0 - 0.5 * eps, 0.5 * eps


def scale_img(img, ratio=1.0, same_shape=False, gs=32):
    if ratio == 1.0:
        return img
    h, w = img.shape[2:]
    s = (int(h * ratio), int(w * ratio))
    img = nn.functional.interpolate(img, size=s, mode="bilinear", align_corners=False)
    if not same_shape:
        h, w = (math.ceil(x * ratio / gs) * gs for x in (h, w))
    return nn.functional.pad(img, [0, w - s[1], 0, h - s[0]], value=0.447)


def copy_attr(a, b, include=(), exclude=()):
    for k, v in b.__dict__.items():
        if (len(include) and k not in include) or k.startswith("_") or k in exclude:
            continue
        setattr(a, k, v)


def is_parallel(model):
    return isinstance(model, (nn.parallel.DataParallel, nn.parallel.DistributedDataParallel))


def intersect_dicts(da, db, exclude=()):
    return {k: v for k, v in da.items() if k in db and all(x not in k for x in exclude) and v.shape == db[k].shape}


def one_cycle(y1=0.0, y2=1.0, steps=100):
    return lambda x: max((1 - math.cos(x * math.pi / steps)) / 2, 0) * (y2 - y1) + y1


def time_sync():
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    return time.time()


def autocast(enabled, device="cuda"):
    return torch.amp.autocast(device, enabled=enabled)


def torch_load(*a, **k):
    k.setdefault("weights_only", False)
    return torch.load(*a, **k)


class Profile(contextlib.ContextDecorator):
    def __init__(self, t=0.0, device=None):
        self.t, self.dt, self.device = t, 0.0, device
        self.cuda = bool(device and str(device).startswith("cuda"))

    def __enter__(self):
        self.start = self.time()
        return self

    def __exit__(self, *exc):
        self.dt = self.time() - self.start
        self.t += self.dt

    def time(self):
        if self.cuda:
            torch.cuda.synchronize(self.device)
        return time.perf_counter()


def colorstr(*args):
    return str(args[-1]) if args else ""


class TryExcept(contextlib.ContextDecorator):
    def __init__(self, msg="", verbose=True):
        self.msg = msg

    def __enter__(self):
        return self

    def __exit__(self, et, ev, tb):
        return True


def threaded(func):
    return func


_LOGGER = logging.getLogger("refshim")
_LOGGER.addHandler(logging.NullHandler())
_LOGGER.propagate = False

_REAL = {
    "ultralytics": {"__version__": "8.4.118"},
    "ultralytics.utils": {
        "LOGGER": _LOGGER,
        "colorstr": colorstr,
        "TryExcept": TryExcept,
        "threaded": threaded,
        "emojis": lambda s="": s,
        "get_default_args": lambda f: {},
    },
    "ultralytics.utils.ops": {
        "Profile": Profile,
        "clip_boxes": clip_boxes,
        "make_divisible": make_divisible,
        "xywh2xyxy": xywh2xyxy,
    },
    "ultralytics.utils.patches": {"torch_load": torch_load},
    "ultralytics.utils.checks": {"is_ascii": lambda s="": all(ord(c) < 128 for c in str(s))},
    "ultralytics.utils.torch_utils": {
        "intersect_dicts": intersect_dicts,
        "one_cycle": one_cycle,
        "autocast": autocast,
        "copy_attr": copy_attr,
        "initialize_weights": initialize_weights,
        "is_parallel": is_parallel,
        "model_info": lambda *a, **k: None,
        "scale_img": scale_img,
        "time_sync": time_sync,
    },
    "ultralytics.utils.metrics": {"box_iou": box_iou, "bbox_iou": bbox_iou, "smooth_bce": smooth_bce},
}
_PREFIXES = ("ultralytics", "matplotlib", "seaborn", "thop")


class _StubModule(types.ModuleType):
    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        return _Inert(f"{self.__name__}.{item}")


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _PREFIXES and fullname.split(".")[0] != "thop":
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        for k, v in _REAL.get(spec.name, {}).items():
            setattr(m, k, v)
        return m

    def exec_module(self, module):
        pass


def install():
    """Put the shim finder and /root/reference on the import path (idempotent)."""
    import packaging.version  # noqa: F401  reference utils/general.py:27,256 relies on this side effect

    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
