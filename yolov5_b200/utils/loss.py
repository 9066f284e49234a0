"""ComputeLoss with the reference's interface (reference utils/loss.py:101-247): ``ComputeLoss(model)(p, targets) ->
(loss (1,), items (3,))`` and ``.build_targets(p, targets)``.  build_targets, the gather/CIoU/scatter and both BCE
terms -- forward and backward -- run in liby5b200 (y5_loss_fwd_bwd); the returned loss carries a custom autograd
node that hands the kernel-computed gradient of every prediction level back to PyTorch.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib
from .._lib import LossParams
from .torch_utils import de_parallel


def smooth_bce(eps=0.1):
    return 1.0 - 0.5 * eps, 0.5 * eps


class _LossFn(torch.autograd.Function):
    """loss, items = ComputeLoss(p, targets).  The forward launch set computes the loss only; the backward one re-runs it with
    the gradient outputs enabled and the UPSTREAM gradient of the loss (GradScaler's factor x WORLD_SIZE x ..., a device
    scalar) multiplied in fp32 inside the kernels before anything is rounded to the prediction dtype -- exactly where
    autograd applies it for `scaler.scale(loss).backward()` (reference train.py:410).  Scaling fp16 gradients afterwards
    would overflow (65536 is not an fp16 number) and would already have flushed small objectness gradients to zero."""

    @staticmethod
    def forward(ctx, crit, targets, *p):
        out, _ = crit._run(p, targets, want_grad=False)
        ctx.crit, ctx.targets = crit, targets
        ctx.save_for_backward(*p)
        ctx.mark_non_differentiable(out[1])
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_loss, _g_items):
        p = ctx.saved_tensors
        if not any(ctx.needs_input_grad[2:]):
            return (None, None) + (None,) * len(p)
        scale = g_loss.detach().reshape(-1)[:1].to(p[0].device, torch.float32).contiguous()
        _, grads = ctx.crit._run(p, ctx.targets, want_grad=True, grad_scale=scale)
        return (None, None) + tuple(grads)


class ComputeLoss:
    sort_obj_iou = False

    def __init__(self, model, autobalance=False):
        if autobalance:
            raise NotImplementedError("y5b200: autobalance is outside the hot path (reference default False)")
        m = de_parallel(model).model[-1]
        h = model.hyp
        if h.get("fl_gamma", 0.0) > 0:
            raise NotImplementedError("y5b200: focal loss (fl_gamma > 0) is outside the hot path (default hyp uses 0)")
        self.hyp = h
        self.device = next(model.parameters()).device
        self.cp, self.cn = smooth_bce(eps=h.get("label_smoothing", 0.0))
        self.balance = {3: [4.0, 1.0, 0.4]}.get(m.nl, [4.0, 1.0, 0.25, 0.06, 0.02])
        self.na, self.nc, self.nl = m.na, m.nc, m.nl
        self.anchors = m.anchors
        self.gr = 1.0
        self._ws = None

    # -------------------------------------------------------------------------------------------------------------
    def _params(self, p, nt):
        q = LossParams()
        q.nl, q.batch, q.na, q.no, q.nc = self.nl, p[0].shape[0], self.na, p[0].shape[-1], self.nc
        for l, t in enumerate(p):
            q.ny[l], q.nx[l] = t.shape[2], t.shape[3]
            q.balance[l] = self.balance[l]
        q.dtype = _lib.dtype_code(p[0].dtype)
        q.nt = nt
        h = self.hyp
        q.anchor_t, q.box_gain, q.obj_gain, q.cls_gain = h["anchor_t"], h["box"], h["obj"], h["cls"]
        q.cls_pw, q.obj_pw, q.cp, q.cn = h["cls_pw"], h["obj_pw"], self.cp, self.cn
        q.grad_scale = 1.0
        return q

    def _run(self, p, targets, want_grad, grad_scale=None):
        if not all(t.is_cuda for t in p):
            raise RuntimeError("y5b200: ComputeLoss runs on CUDA tensors only (no CPU / PyTorch fallback)")
        lib = _lib.lib()
        dev = p[0].device
        p = [t.contiguous() for t in p]
        tg = targets.to(dev, torch.float32).contiguous().view(-1, 6)
        q = self._params(p, tg.shape[0])
        need = int(lib.y5_loss_workspace_bytes(C.byref(q)))
        if need < 0:
            _lib.check(-1, "loss_workspace_bytes")
        key = (dev.index, _lib.stream_ptr(dev))  # scratch per (device, stream): concurrent streams never share it
        if self._ws is None:
            self._ws = {}
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need + 256:
            ws = self._ws[key] = torch.empty(need + 256, dtype=torch.uint8, device=dev)
        ws_ptr = (ws.data_ptr() + 255) & ~255
        anchors = self.anchors.to(dev, torch.float32).contiguous()
        out = torch.empty(4, dtype=torch.float32, device=dev)
        grads = [torch.empty_like(t) for t in p] if want_grad else None
        pl = (C.c_void_p * self.nl)(*[t.data_ptr() for t in p])
        gl = (C.c_void_p * self.nl)(*[g.data_ptr() for g in grads]) if want_grad else None
        with _lib.on(dev):
            _lib.check(lib.y5_loss_fwd_bwd_scaled(C.byref(q), pl, tg.data_ptr(), anchors.data_ptr(), out.data_ptr(), gl,
                                                  grad_scale.data_ptr() if grad_scale is not None else None, ws_ptr, need,
                                                  C.c_void_p(_lib.stream_ptr(dev))), "loss_fwd_bwd")
        self._last = (q, ws_ptr)
        return (out[0:1], out[1:4]), grads

    def __call__(self, p, targets):
        loss, items = _LossFn.apply(self, targets, *p)
        return loss, items.detach()

    def build_targets(self, p, targets):
        """(tcls, tbox, indices, anch) like reference utils/loss.py:185-247 (int64 indices, fp32 boxes)."""
        self._run(p, targets, want_grad=False)
        q, ws_ptr = self._last
        lib = _lib.lib()
        dev = p[0].device
        cap = max(1, 5 * self.na * q.nt)
        tcls, tbox, indices, anch = [], [], [], []
        import numpy as np

        for l in range(self.nl):
            idx = np.empty((5, cap), np.int64)
            tb = np.empty((cap, 4), np.float32)
            cnt = C.c_int32()
            _lib.check(lib.y5_loss_read_targets(C.byref(q), ws_ptr, l, idx.ctypes.data, tb.ctypes.data, C.byref(cnt),
                                                C.c_void_p(_lib.stream_ptr(dev))), "loss_read_targets")
            n = cnt.value
            ii = torch.from_numpy(idx.reshape(-1)[: 5 * n].reshape(5, n).copy()).to(dev)
            indices.append((ii[0], ii[1], ii[2], ii[3]))
            tcls.append(ii[4])
            tbox.append(torch.from_numpy(tb[:n].copy()).to(dev))
            anch.append(self.anchors.to(dev)[l][ii[1]])
        return tcls, tbox, indices, anch
