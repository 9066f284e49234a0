"""ORACLE (test infrastructure, never on the product path): CPU fp32 restatement of the YOLOv5 forward.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
It is a functional re-expression of what the reference computes, driven by a *state_dict with the
reference's key names* plus the model dict, so the very same weights can be fed to the reference
(tests/golden/make_golden.py, build container only) and to the CUDA engine.

Pinned: tests/golden/model_*.npz hold outputs of the real reference (imported through
tests/golden/refshim.py) for seeded weights/inputs; tests/test_oracle_golden.py checks this file against them.
The training-mode forward (``bn_batch_stats=True``: BatchNorm with batch statistics, models/common.py:86-88 under
model.train()) is pinned the same way by tests/golden/train_step.npz (reference forward + ComputeLoss + backward).

Reference lines restated (paths relative to /root/reference):
  models/common.py:62-92     autopad, Conv (conv -> BN(eps 1e-3) -> SiLU)      -> conv_block
  models/common.py:164-181   Bottleneck                                          -> bottleneck
  models/common.py:230-246   C3                                                  -> c3
  models/common.py:318-340   SPPF (three chained 5x5/s1/p2 max-pools)            -> sppf
  models/common.py:443-453   Concat;  nn.Upsample(None, 2, 'nearest')            -> in forward()
  models/common.py:1104-1117 Proto                                               -> proto
  models/yolo.py:91-128      Detect.forward / _make_grid                         -> detect
  models/yolo.py:144-150     Segment.forward                                     -> forward()
  models/yolo.py:160-170     BaseModel._forward_once routing (f / save)          -> forward()
  models/yolo.py:375-458     parse_model (width/depth scaling)                   -> parse_layers
  utils/torch_utils.py:224-254 fuse_conv_and_bn                                  -> fold_bn
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # set by ultralytics initialize_weights (reference models/yolo.py:259)
_BN_BATCH_STATS = False  # forward(..., bn_batch_stats=True): BatchNorm normalises with batch statistics (module.train())


def make_divisible(x: float, d: int) -> int:
    return math.ceil(x / d) * d


def parse_layers(cfg: dict, ch: int = 3):
    """Return (layers, save): each layer = dict(i, f, kind, c1, c2, n, args). reference models/yolo.py:375-458."""
    anchors, nc, gd, gw = cfg["anchors"], cfg["nc"], cfg["depth_multiple"], cfg["width_multiple"]
    div = cfg.get("channel_multiple") or 8
    na = len(anchors[0]) // 2 if isinstance(anchors, list) else anchors
    no = na * (nc + 5)
    chs: list[int] = [ch]
    layers, save = [], []
    c2 = ch
    for i, (f, n, kind, args) in enumerate(cfg["backbone"] + cfg["head"]):
        args = [({"nc": nc, "anchors": anchors, "None": None}.get(a, a) if isinstance(a, str) else a) for a in args]
        n = max(round(n * gd), 1) if n > 1 else n
        rec = {"i": i, "f": f, "kind": kind, "n": n}
        if kind in ("Conv", "C3", "SPPF"):
            c1, c2 = chs[f], args[0]
            if c2 != no:
                c2 = make_divisible(c2 * gw, div)
            rec.update(c1=c1, c2=c2, args=args[1:])
        elif kind == "Concat":
            c2 = sum(chs[x] for x in f)
            rec.update(c2=c2)
        elif kind in ("Detect", "Segment"):
            rec.update(ch=[chs[x] for x in f], nc=nc, anchors=anchors)
            if kind == "Segment":
                rec.update(nm=args[2], npr=make_divisible(args[3] * gw, div))
        elif kind == "nn.Upsample":
            c2 = chs[f]
            rec.update(c2=c2, scale=args[1])
        else:
            raise NotImplementedError(kind)
        save.extend(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)
        layers.append(rec)
        if i == 0:
            chs = []
        chs.append(c2)
    return layers, sorted(save)


def fold_bn(w, gamma, beta, mean, var, eps=BN_EPS):
    """BN folding exactly as fuse_conv_and_bn does it (utils/torch_utils.py:245-252): returns (W', b')."""
    scale = gamma / torch.sqrt(eps + var)
    w2 = torch.mm(torch.diag(scale), w.reshape(w.shape[0], -1)).view(w.shape)
    b2 = beta - gamma * mean / torch.sqrt(var + eps)
    return w2, b2


def conv_block(sd, p, x, k=1, s=1, pad=None, fused=False, act=True):
    """Conv = conv(no bias) -> BN(running stats) -> SiLU; or, if ``fused``, conv(W', b') -> SiLU."""
    pad = k // 2 if pad is None else pad
    w = sd[f"{p}.conv.weight"]
    if f"{p}.bn.weight" in sd:
        g, b, m, v = (sd[f"{p}.bn.{q}"] for q in ("weight", "bias", "running_mean", "running_var"))
        if fused:
            w2, b2 = fold_bn(w, g, b, m, v)
            y = F.conv2d(x, w2, b2, stride=s, padding=pad)
        else:
            y = F.conv2d(x, w, None, stride=s, padding=pad)
            if _BN_BATCH_STATS:  # nn.BatchNorm2d in training mode (models/common.py:86-88 under model.train())
                y = F.batch_norm(y, None, None, g, b, training=True, eps=BN_EPS)
            else:
                y = F.batch_norm(y, m, v, g, b, training=False, eps=BN_EPS)
    else:  # a state_dict taken from an already fused reference model: conv has a bias, no bn keys
        y = F.conv2d(x, w, sd[f"{p}.conv.bias"], stride=s, padding=pad)
    return F.silu(y) if act else y


def bottleneck(sd, p, x, shortcut, fused):
    y = conv_block(sd, f"{p}.cv2", conv_block(sd, f"{p}.cv1", x, 1, 1, fused=fused), 3, 1, fused=fused)
    return x + y if shortcut else y  # c1 == c2 always holds inside C3 (e=1.0)


def c3(sd, p, x, n, shortcut, fused):
    a = conv_block(sd, f"{p}.cv1", x, fused=fused)
    for j in range(n):
        a = bottleneck(sd, f"{p}.m.{j}", a, shortcut, fused)
    b = conv_block(sd, f"{p}.cv2", x, fused=fused)
    return conv_block(sd, f"{p}.cv3", torch.cat((a, b), 1), fused=fused)


def sppf(sd, p, x, k, fused):
    x = conv_block(sd, f"{p}.cv1", x, fused=fused)
    y1 = F.max_pool2d(x, k, 1, k // 2)
    y2 = F.max_pool2d(y1, k, 1, k // 2)
    y3 = F.max_pool2d(y2, k, 1, k // 2)
    return conv_block(sd, f"{p}.cv2", torch.cat((x, y1, y2, y3), 1), fused=fused)


def proto(sd, p, x, fused):
    x = conv_block(sd, f"{p}.cv1", x, 3, 1, fused=fused)
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    x = conv_block(sd, f"{p}.cv2", x, 3, 1, fused=fused)
    return conv_block(sd, f"{p}.cv3", x, fused=fused)


def detect(sd, p, xs, nc, nm, strides, training):
    """Head: 1x1 conv with bias, reshape to (B,na,ny,nx,no), decode.  Returns (z, raw_list) (eval) or raw_list."""
    anchors = sd[f"{p}.anchors"]  # (nl, na, 2) in grid units
    na = anchors.shape[1]
    no = 5 + nc + nm
    raw, z = [], []
    for i, x in enumerate(xs):
        y = F.conv2d(x, sd[f"{p}.m.{i}.weight"], sd[f"{p}.m.{i}.bias"])
        bs, _, ny, nx = y.shape
        y = y.view(bs, na, no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
        raw.append(y)
        if training:
            continue
        gy, gx = torch.meshgrid(torch.arange(ny, dtype=y.dtype, device=y.device), torch.arange(nx, dtype=y.dtype, device=y.device), indexing="ij")
        grid = torch.stack((gx, gy), 2).expand(1, na, ny, nx, 2) - 0.5
        agrid = (anchors[i] * strides[i]).view(1, na, 1, 1, 2).expand(1, na, ny, nx, 2)
        if nm:
            xy, wh, conf, mask = y.split((2, 2, nc + 1, nm), 4)
            xy = (xy.sigmoid() * 2 + grid) * strides[i]
            wh = (wh.sigmoid() * 2) ** 2 * agrid
            d = torch.cat((xy, wh, conf.sigmoid(), mask), 4)
        else:
            xy, wh, conf = y.sigmoid().split((2, 2, nc + 1), 4)
            xy = (xy * 2 + grid) * strides[i]
            wh = (wh * 2) ** 2 * agrid
            d = torch.cat((xy, wh, conf), 4)
        z.append(d.view(bs, na * ny * nx, no))
    return raw if training else (torch.cat(z, 1), raw)


def model_strides(cfg: dict) -> list[float]:
    """Detect strides, i.e. what models/yolo.py:250-256 measures with a 256x256 probe forward."""
    layers, _ = parse_layers(cfg)
    red = []  # spatial reduction factor of each layer's output w.r.t. the image
    for L in layers:
        f = L["f"]
        if L["kind"] in ("Detect", "Segment"):
            return [float(red[j]) for j in f]
        r = red[f if isinstance(f, int) else f[0]] if red else 1  # negative f indexes from the end, like ys[f]
        if L["kind"] == "Conv":
            r = r * (L["args"][1] if len(L["args"]) > 1 else 1)
        elif L["kind"] == "nn.Upsample":
            r = r / L["scale"]
        red.append(r)
    raise ValueError("model has no Detect/Segment layer")


def forward(cfg: dict, sd: dict, x: torch.Tensor, training: bool = False, fused: bool = False, ch: int = 3, bn_batch_stats: bool = False):
    """See _forward; ``bn_batch_stats`` evaluates every BatchNorm with batch statistics (the training-mode forward)."""
    global _BN_BATCH_STATS
    prev, _BN_BATCH_STATS = _BN_BATCH_STATS, bool(bn_batch_stats)
    try:
        return _forward(cfg, sd, x, training, fused, ch)
    finally:
        _BN_BATCH_STATS = prev


def _forward(cfg: dict, sd: dict, x: torch.Tensor, training: bool = False, fused: bool = False, ch: int = 3):
    """Whole-model forward on CPU fp32 tensors.

    eval Detect : (z (B,N,no), [raw_i (B,na,ny,nx,no)])          models/yolo.py:115
    eval Segment: (z, proto (B,nm,H/4... ), [raw_i])              models/yolo.py:150
    training    : [raw_i]  (Detect)  /  ([raw_i], proto) (Segment)
    """
    layers, save = parse_layers(cfg, ch)
    strides = model_strides(cfg)
    ys = []
    for L in layers:
        f, i, kind = L["f"], L["i"], L["kind"]
        if f != -1:
            x = ys[f] if isinstance(f, int) else [x if j == -1 else ys[j] for j in f]
        p = f"model.{i}"
        if kind == "Conv":
            a = L["args"]  # [k, s, (p)]
            k = a[0] if len(a) > 0 else 1
            s = a[1] if len(a) > 1 else 1
            pad = a[2] if len(a) > 2 else None
            x = conv_block(sd, p, x, k, s, pad, fused)
        elif kind == "C3":
            shortcut = L["args"][0] if L["args"] else True
            x = c3(sd, p, x, L["n"], shortcut, fused)
        elif kind == "SPPF":
            x = sppf(sd, p, x, L["args"][0] if L["args"] else 5, fused)
        elif kind == "nn.Upsample":
            x = F.interpolate(x, scale_factor=L["scale"], mode="nearest")
        elif kind == "Concat":
            x = torch.cat(x, 1)
        elif kind == "Detect":
            x = detect(sd, p, list(x), L["nc"], 0, strides, training)
        elif kind == "Segment":
            pr = proto(sd, f"{p}.proto", x[0], fused)
            d = detect(sd, p, list(x), L["nc"], L["nm"], strides, training)
            x = (d, pr) if training else (d[0], pr, d[1])
        ys.append(x if i in save else None)
    return x


# ----------------------------------------------------------------------------------------------------------------------
# deterministic synthetic weights shared by the golden generator, the tests, smoke() and bench.py
# ----------------------------------------------------------------------------------------------------------------------
def param_shapes(cfg: dict, ch: int = 3) -> dict:
    """state_dict key -> shape for the unfused model, in the reference's own key order."""
    layers, _ = parse_layers(cfg, ch)
    out: dict[str, tuple] = {}

    def conv(p, c1, c2, k):
        out[f"{p}.conv.weight"] = (c2, c1, k, k)
        for q in ("weight", "bias", "running_mean", "running_var"):
            out[f"{p}.bn.{q}"] = (c2,)
        out[f"{p}.bn.num_batches_tracked"] = ()

    for L in layers:
        p, kind = f"model.{L['i']}", L["kind"]
        if kind == "Conv":
            conv(p, L["c1"], L["c2"], L["args"][0] if L["args"] else 1)
        elif kind == "C3":
            c1, c2 = L["c1"], L["c2"]
            c_ = int(c2 * 0.5)
            conv(f"{p}.cv1", c1, c_, 1)
            conv(f"{p}.cv2", c1, c_, 1)
            conv(f"{p}.cv3", 2 * c_, c2, 1)
            for j in range(L["n"]):
                conv(f"{p}.m.{j}.cv1", c_, c_, 1)
                conv(f"{p}.m.{j}.cv2", c_, c_, 3)
        elif kind == "SPPF":
            c_ = L["c1"] // 2
            conv(f"{p}.cv1", L["c1"], c_, 1)
            conv(f"{p}.cv2", c_ * 4, L["c2"], 1)
        elif kind in ("Detect", "Segment"):
            na = len(L["anchors"][0]) // 2
            nm = L.get("nm", 0)
            out[f"{p}.anchors"] = (len(L["anchors"]), na, 2)
            for l, c in enumerate(L["ch"]):
                out[f"{p}.m.{l}.weight"] = (na * (5 + L["nc"] + nm), c, 1, 1)
                out[f"{p}.m.{l}.bias"] = (na * (5 + L["nc"] + nm),)
            if kind == "Segment":
                conv(f"{p}.proto.cv1", L["ch"][0], L["npr"], 3)
                conv(f"{p}.proto.cv2", L["npr"], L["npr"], 3)
                conv(f"{p}.proto.cv3", L["npr"], nm, 1)
    return out


def synth_state_dict(cfg: dict, seed: int = 0, ch: int = 3, head_bias: str = "init") -> dict:
    """Seeded, torch-version-independent weights (numpy RandomState), non-trivial BN statistics.

    conv W ~ U(-a, a), a = 1/sqrt(fan_in) (kaiming-uniform(a=sqrt 5) bound, what nn.Conv2d draws);
    BN gamma ~ U(.5,1.5), beta ~ N(0,.1), mean ~ N(0,.1), var ~ U(.5,1.5)  (SURVEY.md section 8d).
    head_bias='init' reproduces models/yolo.py:314-327 on top of the random bias; 'hot' adds +3 to the
    objectness logit instead so that the decoded output holds NMS candidates.
    """
    import numpy as np

    rs = np.random.RandomState(seed)
    strides = model_strides(cfg)
    nc = cfg["nc"]
    sd = {}
    for k, shp in param_shapes(cfg, ch).items():
        leaf = k.rsplit(".", 2)
        if k.endswith("num_batches_tracked"):
            v = np.zeros((), np.int64)
        elif k.endswith("anchors"):
            a = np.asarray(cfg["anchors"], np.float32).reshape(len(cfg["anchors"]), -1, 2)
            v = a / np.asarray(strides, np.float32).reshape(-1, 1, 1)
        elif ".bn." in k:
            q = leaf[-1]
            if q in ("weight", "running_var"):
                v = rs.uniform(0.5, 1.5, shp)
            else:
                v = rs.normal(0.0, 0.1, shp)
        elif k.endswith(".bias"):  # Detect conv bias
            fan_in = param_shapes(cfg, ch)[k[: -len("bias")] + "weight"][1]
            v = rs.uniform(-1, 1, shp) / math.sqrt(fan_in)
            lvl = int(k.split(".")[-2])
            na = len(cfg["anchors"][0]) // 2
            b = v.reshape(na, -1)
            if head_bias == "init":
                b[:, 4] += math.log(8 / (640 / strides[lvl]) ** 2)
                b[:, 5 : 5 + nc] += math.log(0.6 / (nc - 0.99999))
            else:
                b[:, 4] += 3.0
            v = b.reshape(-1)
        else:  # conv weight
            fan_in = shp[1] * shp[2] * shp[3]
            v = rs.uniform(-1, 1, shp) / math.sqrt(fan_in)
        sd[k] = torch.from_numpy(np.asarray(v, dtype=np.int64 if k.endswith("tracked") else np.float32).copy())
    return sd
