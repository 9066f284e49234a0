"""Program planner: turns a DetectionModel (or a single layer) into a fixed sequence of liby5b200 kernel launches.

What the reference does layer by layer through torch.nn (models/yolo.py:160-170 `_forward_once`) becomes, for one
(batch, height, width, dtype):

  * every activation is a channel-slice VIEW of a pre-allocated NHWC buffer; a Concat's inputs are allocated inside
    the Concat's buffer, so torch.cat (models/common.py:246,340,453) disappears;
  * C3's cv1 and cv2 (same input, models/common.py:246) run as ONE GEMM with stacked output channels that lands
    directly in the C3's concat buffer; each Bottleneck's residual add (models/common.py:181) is the epilogue of its
    3x3 conv, in place; SPPF's three pools are one kernel; Upsample writes into its Concat slice;
  * Conv = conv + folded BN + SiLU in one tcgen05 implicit-GEMM kernel (weights are folded/packed once, here);
  * the Detect/Segment head levels run the GEMM with the decode epilogue and write fresh output tensors each call;
  * the whole fixed part is captured in a CUDA graph and replayed.

PyTorch is used for device memory and streams only; all arithmetic is in liby5b200.so.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch

from . import _lib
from ._lib import ConvDesc, DetectDesc

BN_EPS_DEFAULT = 1e-3


class View:
    """Channel slice [coff, coff+c) of an NHWC buffer (B, H, W, pitch)."""

    __slots__ = ("buf", "coff", "c", "h", "w")

    def __init__(self, buf: torch.Tensor, coff: int, c: int):
        self.buf, self.coff, self.c = buf, coff, c
        self.h, self.w = buf.shape[1], buf.shape[2]

    @property
    def pitch(self) -> int:
        return self.buf.shape[3]

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr() + self.coff * self.buf.element_size()

    def slice(self, coff: int, c: int) -> "View":
        assert 0 <= coff and coff + c <= self.c
        return View(self.buf, self.coff + coff, c)

    def dense_nhwc(self) -> torch.Tensor:
        return self.buf[..., self.coff : self.coff + self.c]


def _pad8(c: int) -> int:
    return (c + 7) // 8 * 8


def fold_conv_bn(conv: torch.nn.Conv2d, bn) -> tuple[torch.Tensor, torch.Tensor]:
    """fp32 (W', b') of conv followed by eval-mode BN; same algebra as utils/torch_utils.py:245-252."""
    w = conv.weight.detach().float()
    if bn is None:
        b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
        return w, b
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    w2 = w * scale.view(-1, 1, 1, 1)
    b0 = conv.bias.detach().float() if conv.bias is not None else torch.zeros_like(scale)
    b2 = bn.bias.detach().float() + (b0 - bn.running_mean.detach().float()) * scale
    return w2, b2


def pack_weight(w: torch.Tensor, block_k: int, dtype: torch.dtype) -> torch.Tensor:
    """OIHW fp32 -> [O][kh][kw][cin_pad] (K-major) in the activation dtype, channel dim zero padded to block_k."""
    o, i, kh, kw = w.shape
    cin_pad = (i + block_k - 1) // block_k * block_k
    out = torch.zeros(o, kh, kw, cin_pad, dtype=dtype, device=w.device)
    out[..., :i] = w.permute(0, 2, 3, 1).to(dtype)
    return out.contiguous()


def stem_weight_s2d(w: torch.Tensor) -> torch.Tensor:
    """(O,3,6,6) stride-2 pad-2 filter -> equivalent (O,16,3,3) stride-1 pad-1 filter over the 2x2 space-to-depth
    input (channel = (dy*2+dx)*3 + c, see y5_stem_s2d); 4 of the 16 channels stay zero."""
    o = w.shape[0]
    out = torch.zeros(o, 16, 3, 3, dtype=w.dtype, device=w.device)
    for dy in range(2):
        for dx in range(2):
            for c in range(3):
                out[:, (dy * 2 + dx) * 3 + c] = w[:, c, dy::2, dx::2]
    return out


class _Op:
    """One launch: (function, ctypes args...) bound at plan time; run(stream) issues it."""

    __slots__ = ("fn", "args", "name", "keep")

    def __init__(self, name, fn, args, keep=()):
        self.name, self.fn, self.args, self.keep = name, fn, args, keep

    def run(self, stream: int):
        _lib.check(self.fn(*self.args, C.c_void_p(stream)), self.name)


class Program:
    """Kernel sequence for one module (a whole DetectionModel or a single layer) at a fixed input shape."""

    def __init__(self, module, batch: int, height: int, width: int, dtype: torch.dtype, device, in_channels=None):
        if dtype not in (torch.float16, torch.bfloat16):
            raise TypeError(f"y5b200 engine computes in fp16 or bf16, got {dtype} (call model.half() / .bfloat16())")
        self.lib = _lib.lib()
        self.module = module
        self.B, self.H, self.W = batch, height, width
        self.dtype, self.device = dtype, torch.device(device)
        self.dt_code = _lib.dtype_code(dtype)
        self.ops: list[_Op] = []          # graph-capturable fixed part
        self.head_ops: list = []          # detect levels (fresh outputs per call)
        self._plans: list = []            # (destroy_fn, handle)
        self._keep: list = []             # packed weights / biases / buffers
        self.graph = None
        self.flops = 0                    # 2*MAC of every conv in the program (per batch)
        self.act_bytes = 0                # algorithmic activation bytes: each conv reads its input once, writes its output once
        self.weight_bytes = 0
        self.in_channels = in_channels
        self._build()

    # ------------------------------------------------------------------ allocation helpers
    def new_buf(self, h: int, w: int, c: int) -> torch.Tensor:
        t = torch.zeros(self.B, h, w, _pad8(c), dtype=self.dtype, device=self.device)
        self._keep.append(t)
        return t

    def new_view(self, h: int, w: int, c: int) -> View:
        return View(self.new_buf(h, w, c), 0, c)

    # ------------------------------------------------------------------ op emitters
    def block_k(self, cin: int, cout: int, m_rows: int) -> int:
        bk = C.c_int32()
        _lib.check(self.lib.y5_conv_pick(cin, cout, m_rows, C.byref(bk), None), "conv_pick")
        return bk.value

    def fold_pack(self, parts, m_rows: int):
        """Folded + packed weights of one GEMM from module parameters, one y5_fold_pack launch per part (no ATen arithmetic).
        parts: list of (weight (O,I,kh,kw) tensor, conv bias | None, bn | None); several parts stack along the output channels
        (C3's cv1 | cv2).  Returns (packed [sum O][kh][kw][I_pad] in the activation dtype, fp32 bias [sum O], block_k)."""
        cin, kh, kw = parts[0][0].shape[1:]
        cout = sum(w.shape[0] for w, _, _ in parts)
        bk = self.block_k(cin, cout, m_rows)
        ipad = (cin + bk - 1) // bk * bk
        wp = torch.empty(cout, kh, kw, ipad, dtype=self.dtype, device=self.device)
        bias = torch.empty(cout, dtype=torch.float32, device=self.device)
        self.fold_pack_into(parts, wp, bias, 0, ipad)
        return wp, bias, bk

    def fold_pack_into(self, parts, wp, bias, row0: int, ipad: int, pad_rows_to: int | None = None):
        st = C.c_void_p(_lib.stream_ptr(self.device))
        es = wp.element_size()
        for w, cb, bn in parts:
            w = w.detach()
            if not w.is_contiguous():
                w = w.contiguous()
            if w.device != self.device:
                raise RuntimeError("y5b200: module parameters and the input must live on the same CUDA device")
            o, i, kh, kw = w.shape
            rows = pad_rows_to or o
            keep = [w]
            if cb is not None:
                cb = cb.detach().to(w.dtype).contiguous()
                keep.append(cb)
            g = b_ = mu = var = None
            bn_code, eps = _lib.Y5_F32, 0.0
            if bn is not None:
                g, b_, mu, var = (t.detach().contiguous() for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var))
                if not (g.dtype == b_.dtype == mu.dtype == var.dtype):
                    g, b_, mu, var = g.float(), b_.float(), mu.float(), var.float()
                bn_code, eps = _lib.dtype_code(g.dtype), float(bn.eps)
                keep += [g, b_, mu, var]
            _lib.check(self.lib.y5_fold_pack(w.data_ptr(), _lib.dtype_code(w.dtype), o, i, kh, kw, cb.data_ptr() if cb is not None else None,
                                             g.data_ptr() if g is not None else None, b_.data_ptr() if b_ is not None else None,
                                             mu.data_ptr() if mu is not None else None, var.data_ptr() if var is not None else None, bn_code, eps,
                                             wp.data_ptr() + row0 * kh * kw * ipad * es, ipad, rows, bias.data_ptr() + row0 * 4, self.dt_code, st),
                       "fold_pack")
            self._keep += keep  # the launch is asynchronous: temporaries must outlive it
            row0 += rows

    def conv(self, x: View, out: View, w_fp32, b_fp32, k: int, s: int, p: int, act: bool,
             residual: View | None = None, name: str = "conv", virt=None, packed=None):
        """Emit one fused conv.  Weights come either as fp32 tensors (w_fp32 OIHW already folded, b_fp32) that are packed here,
        or pre-packed by fold_pack: packed = (wp, bias, block_k, cin).  `virt` (stem only) = dict(ptr, in_c, in_w, in_h, x_stride,
        y_stride, n_stride, kw, pad_w): a strided "wide pixel" view of the input and a non-square filter, see y5_conv_desc in
        include/y5b200.h."""
        cout = out.c
        if virt is None:
            cin, in_h, in_w, kw = x.c, x.h, x.w, k
            ho, wo = (x.h + 2 * p - k) // s + 1, (x.w + 2 * p - k) // s + 1
        else:
            cin, in_h, in_w, kw = virt["in_c"], virt["in_h"], virt["in_w"], virt["kw"]
            ho, wo = (in_h + 2 * p - k) // s + 1, (in_w + 2 * virt["pad_w"] - kw) // s + 1
        assert (ho, wo) == (out.h, out.w), (name, ho, wo, out.h, out.w)
        m_rows = self.B * ho * wo
        bk, bn = C.c_int32(), C.c_int32()
        bn.value = int(os.environ.get("Y5_FORCE_BLOCK_N", "0"))  # 0: the library's tile cost model decides (block_n, MT)
        if packed is None:
            assert w_fp32.shape == (cout, cin, k, kw), (w_fp32.shape, cout, cin, k, kw)
            bk.value = self.block_k(cin, cout, m_rows)
            wp = pack_weight(w_fp32, bk.value, self.dtype)
            bias = b_fp32.to(torch.float32).contiguous()
        else:
            wp, bias, bk.value, pc = packed
            assert pc == cin and wp.shape[0] == cout and wp.shape[1:3] == (k, kw), (name, wp.shape, cout, cin, k, kw)
        self._keep += [wp, bias]
        d = ConvDesc()
        if virt is None:
            d.inp, d.in_pitch = x.ptr, x.pitch
        else:
            d.inp, d.in_pitch = virt["ptr"], virt["x_stride"]
            d.in_x_stride, d.in_y_stride, d.in_n_stride = virt["x_stride"], virt["y_stride"], virt["n_stride"]
            d.kw, d.pad_w = kw, virt["pad_w"]
        d.batch, d.in_h, d.in_w, d.in_c = self.B, in_h, in_w, cin
        d.weight, d.bias = wp.data_ptr(), bias.data_ptr()
        d.out, d.out_pitch, d.out_c = out.ptr, out.pitch, cout
        d.residual = residual.ptr if residual is not None else None
        d.res_pitch = residual.pitch if residual is not None else 0
        d.ksize, d.stride, d.pad = k, s, p
        d.act = _lib.ACT_SILU if act else _lib.ACT_NONE
        d.dtype, d.block_k, d.block_n = self.dt_code, bk.value, bn.value
        d.a_mode = int(os.environ.get("Y5_FORCE_A_MODE", "0"))
        if d.a_mode == 2 and s != 1:
            d.a_mode = 0
        d.reserved = 64  # packed once at build time: constant weights, the kernel may fetch them before its dependency wait
        plan = C.c_void_p()
        _lib.check(self.lib.y5_conv_plan_create(C.byref(d), C.byref(plan)), f"conv_plan_create[{name}]")
        self._plans.append((self.lib.y5_conv_plan_destroy, plan))
        self.ops.append(_Op(name, self.lib.y5_conv_plan_run, (plan,)))
        if virt is None:
            self.flops += 2 * m_rows * cout * cin * k * k
            self.act_bytes += 2 * (self.B * x.h * x.w * x.c + m_rows * cout)
            self.weight_bytes += 2 * cout * cin * k * k

    def conv_module(self, m, x: View, out: View, residual: View | None = None, name="conv"):
        """m: models.common.Conv (conv + bn + act) in its fused or unfused state."""
        k, s, p = m.conv.kernel_size[0], m.conv.stride[0], m.conv.padding[0]
        act = isinstance(m.act, torch.nn.SiLU)
        if not act and not isinstance(m.act, torch.nn.Identity):
            raise NotImplementedError(f"y5b200: activation {type(m.act).__name__} (only SiLU / Identity are built)")
        if m.conv.groups != 1 or m.conv.dilation[0] != 1:
            raise NotImplementedError("y5b200: grouped / dilated convolutions are outside the YOLOv5 n..x hot path")
        ho, wo = (x.h + 2 * p - k) // s + 1, (x.w + 2 * p - k) // s + 1
        wp, bias, bk = self.fold_pack([(m.conv.weight, m.conv.bias, getattr(m, "bn", None))], self.B * ho * wo)
        self.conv(x, out, None, None, k, s, p, act, residual, name, packed=(wp, bias, bk, x.c))

    def out_hw(self, m, x: View):
        k, s, p = m.conv.kernel_size[0], m.conv.stride[0], m.conv.padding[0]
        return (x.h + 2 * p - k) // s + 1, (x.w + 2 * p - k) // s + 1

    # ------------------------------------------------------------------ module lowering
    def lower_conv(self, m, x, out: View | None, name):
        from .models.common import Conv  # noqa: F401

        k, s, p, cin = m.conv.kernel_size[0], m.conv.stride[0], m.conv.padding[0], m.conv.in_channels
        if isinstance(x, torch.Tensor):  # network input (B,3,H,W) NCHW: stem path
            if not (k == 6 and s == 2 and p == 2 and cin == 3):
                raise NotImplementedError("y5b200: the first layer must be the YOLOv5 v6 stem Conv(3, c, 6, 2, 2)")
            if self.H % 2 or self.W % 2:
                raise ValueError("y5b200: image height and width must be even")
            h2, w2 = self.H // 2, self.W // 2
            # space-to-depth buffer with one zero cell left and right of every row: [B][h2][w2+2][16]
            s2d_buf = torch.zeros(self.B, h2, w2 + 2, 16, dtype=self.dtype, device=self.device)
            self._keep.append(s2d_buf)
            self.stem_in = s2d_buf
            w, b = fold_conv_bn(m.conv, getattr(m, "bn", None))
            out = out or self.new_view(h2, w2, m.conv.out_channels)
            w3 = stem_weight_s2d(w)  # (O,16,3,3): 3x3/s1/p1 over the 16-channel cells
            act = isinstance(m.act, torch.nn.SiLU)
            # 3 horizontally adjacent cells are contiguous in memory (48 channels): run the stem as a 3x1 conv over
            # overlapping 48-channel "wide pixels" (x stride 16 elements) -> 3 taps of K=48 instead of 9 taps of K=16
            wv = w3.permute(0, 3, 1, 2).reshape(w3.shape[0], 48, 3, 1)  # [o][s*16+c][r][0] = w3[o][c][r][s]
            virt = dict(ptr=s2d_buf.data_ptr(), in_c=48, in_w=w2, in_h=h2, x_stride=16, y_stride=(w2 + 2) * 16,
                        n_stride=h2 * (w2 + 2) * 16, kw=1, pad_w=0)
            n_before = len(self.ops)
            try:
                self.conv(None, out, wv, b, 3, 1, 1, act, None, name + "(s2d 3x1x48)", virt=virt)
            except RuntimeError:
                # driver refused the overlapping-stride tensor map: plain 3x3 over 16-channel cells of the padded buffer
                del self.ops[n_before:]
                virt = dict(ptr=s2d_buf.data_ptr() + 16 * s2d_buf.element_size(), in_c=16, in_w=w2, in_h=h2, x_stride=16,
                            y_stride=(w2 + 2) * 16, n_stride=h2 * (w2 + 2) * 16, kw=3, pad_w=1)
                self.conv(None, out, w3, b, 3, 1, 1, act, None, name + "(s2d 3x3x16)", virt=virt)
            cout = m.conv.out_channels
            self.flops += 2 * self.B * h2 * w2 * cout * 3 * 36
            self.act_bytes += 2 * (self.B * self.H * self.W * 3 + self.B * h2 * w2 * cout)
            self.weight_bytes += 2 * cout * 3 * 36
            return out
        ho, wo = self.out_hw(m, x)
        out = out or self.new_view(ho, wo, m.conv.out_channels)
        self.conv_module(m, x, out, None, name)
        return out

    def lower_c3(self, m, x: View, out: View | None, name):
        c_ = m.cv1.conv.out_channels
        cat = self.new_view(x.h, x.w, 2 * c_)
        # cv1 | cv2 stacked along the output channels: one GEMM, result is already the concat layout
        wp, bias, bk = self.fold_pack([(m.cv1.conv.weight, m.cv1.conv.bias, getattr(m.cv1, "bn", None)),
                                       (m.cv2.conv.weight, m.cv2.conv.bias, getattr(m.cv2, "bn", None))], self.B * x.h * x.w)
        self.conv(x, cat, None, None, 1, 1, 0, True, None, f"{name}.cv1|cv2", packed=(wp, bias, bk, x.c))
        a = cat.slice(0, c_)
        if len(m.m):
            tmp = self.new_view(x.h, x.w, c_)
        for j, bt in enumerate(m.m):
            self.conv_module(bt.cv1, a, tmp, None, f"{name}.m{j}.cv1")
            self.conv_module(bt.cv2, tmp, a, a if bt.add else None, f"{name}.m{j}.cv2")  # in-place residual add
        out = out or self.new_view(x.h, x.w, m.cv3.conv.out_channels)
        self.conv_module(m.cv3, cat, out, None, f"{name}.cv3")
        return out

    def lower_sppf(self, m, x: View, out: View | None, name):
        c_ = m.cv1.conv.out_channels
        k = m.m.kernel_size if isinstance(m.m.kernel_size, int) else m.m.kernel_size[0]
        cat = self.new_view(x.h, x.w, 4 * c_)
        s0, s1, s2, s3 = (cat.slice(i * c_, c_) for i in range(4))
        self.conv_module(m.cv1, x, s0, None, f"{name}.cv1")
        self.ops.append(_Op(f"{name}.pool", self.lib.y5_sppf_pool,
                            (s0.ptr, s0.pitch, s1.ptr, s2.ptr, s3.ptr, cat.pitch, self.B, x.h, x.w, c_, k, self.dt_code)))
        out = out or self.new_view(x.h, x.w, m.cv2.conv.out_channels)
        self.conv_module(m.cv2, cat, out, None, f"{name}.cv2")
        return out

    def lower_upsample(self, m, x: View, out: View | None, name):
        sf = m.scale_factor
        if float(sf) != 2.0 or m.mode != "nearest":
            raise NotImplementedError("y5b200: only nn.Upsample(scale_factor=2, mode='nearest')")
        out = out or self.new_view(2 * x.h, 2 * x.w, x.c)
        self.ops.append(_Op(name, self.lib.y5_upsample2x, (x.ptr, x.pitch, out.ptr, out.pitch, self.B, x.h, x.w, x.c, self.dt_code)))
        return out

    def copy_into(self, x: View, out: View, name):
        self.ops.append(_Op(name, self.lib.y5_copy_view, (x.ptr, x.pitch, out.ptr, out.pitch, self.B * x.h * x.w, x.c, self.dt_code)))

    def lower_proto(self, m, x: View, name):
        a = self.lower_conv(m.cv1, x, None, f"{name}.cv1")
        u = self.lower_upsample(m.upsample, a, None, f"{name}.up")
        b = self.lower_conv(m.cv2, u, None, f"{name}.cv2")
        return self.lower_conv(m.cv3, b, None, f"{name}.cv3")

    def lower_detect(self, m, xs: list[View], name):
        na, no, nc = m.na, m.no, m.nc
        self.z_rows = sum(na * v.h * v.w for v in xs)
        self.det_shapes = [(self.B, na, v.h, v.w, no) for v in xs]
        row0 = 0
        for i, v in enumerate(xs):
            conv = m.m[i]
            HEAD_N = 128  # kHeadN in conv_gemm.cu: one anchor per 128-wide N tile
            if no > HEAD_N:
                raise NotImplementedError(f"y5b200: Detect with no={no} > {HEAD_N} outputs per anchor")
            # every anchor's `no` rows padded to HEAD_N: one fold_pack launch per anchor (weights + bias, no BatchNorm)
            bk = C.c_int32(self.block_k(v.c, na * no, self.B * v.h * v.w))
            ipad = (v.c + bk.value - 1) // bk.value * bk.value
            wp = torch.empty(na * HEAD_N, 1, 1, ipad, dtype=self.dtype, device=self.device)
            bias = torch.empty(na * HEAD_N, dtype=torch.float32, device=self.device)
            w4 = conv.weight.detach().contiguous().view(na, no, v.c, 1, 1)
            b2 = conv.bias.detach().contiguous().view(na, no)
            for a_i in range(na):
                self.fold_pack_into([(w4[a_i], b2[a_i], None)], wp, bias, a_i * HEAD_N, ipad, pad_rows_to=HEAD_N)
            self._keep += [wp, bias, w4, b2]
            d = DetectDesc()
            d.inp, d.in_pitch = v.ptr, v.pitch
            d.batch, d.ny, d.nx, d.in_c = self.B, v.h, v.w, v.c
            d.weight, d.bias = wp.data_ptr(), bias.data_ptr()
            dummy = torch.empty(8, dtype=self.dtype, device=self.device)  # real outputs are bound per call
            self._keep.append(dummy)
            d.raw, d.z = dummy.data_ptr(), dummy.data_ptr()
            d.z_rows, d.z_row0 = self.z_rows, row0
            d.na, d.no, d.nc = na, no, nc
            stride = float(m.stride[i])
            d.stride = stride
            anc = (m.anchors[i].detach().float().cpu() * stride).reshape(-1).tolist()
            for q in range(8):
                d.anchor_wh[q] = anc[q] if q < len(anc) else 0.0
            d.dtype, d.block_k = self.dt_code, bk.value
            plan = C.c_void_p()
            _lib.check(self.lib.y5_detect_plan_create(C.byref(d), C.byref(plan)), f"detect_plan_create[{name}.{i}]")
            self._plans.append((self.lib.y5_detect_plan_destroy, plan))
            self.head_ops.append(plan)
            m_rows = self.B * v.h * v.w
            self.flops += 2 * m_rows * na * no * v.c
            self.act_bytes += 2 * (m_rows * v.c + m_rows * na * no)
            self.weight_bytes += 2 * na * no * v.c
            row0 += na * v.h * v.w

    # ------------------------------------------------------------------ builders
    def _build(self):
        from .models import common as mc
        from .models import yolo as my

        mod = self.module
        self.stem_in = None
        self.proto_view = None
        self.single_out = None
        if isinstance(mod, my.BaseModel):
            self._build_model(mod)
        else:
            cin = self.in_channels
            x = self.new_view(self.H, self.W, cin)
            self.layer_in = x
            self.single_out = self._lower(mod, x, None, type(mod).__name__)

    def _lower(self, m, x, out, name):
        from .models import common as mc

        if isinstance(m, mc.Conv):
            return self.lower_conv(m, x, out, name)
        if isinstance(m, mc.C3):
            return self.lower_c3(m, x, out, name)
        if isinstance(m, mc.SPPF):
            return self.lower_sppf(m, x, out, name)
        if isinstance(m, mc.Bottleneck):
            tmp = self.new_view(x.h, x.w, m.cv1.conv.out_channels)
            self.conv_module(m.cv1, x, tmp, None, f"{name}.cv1")
            out = out or self.new_view(x.h, x.w, m.cv2.conv.out_channels)
            self.conv_module(m.cv2, tmp, out, x if m.add else None, f"{name}.cv2")
            return out
        if isinstance(m, torch.nn.Upsample):
            return self.lower_upsample(m, x, out, name)
        if isinstance(m, mc.Proto):
            return self.lower_proto(m, x, name)
        if isinstance(m, torch.nn.Sequential):
            for j, sub in enumerate(m):
                x = self._lower(sub, x, out if j == len(m) - 1 else None, f"{name}.{j}")
            return x
        raise NotImplementedError(f"y5b200: module {type(m).__name__} is outside the engine's hot path")

    def _build_model(self, model):
        from .models import common as mc
        from .models import yolo as my

        layers = list(model.model)
        n = len(layers)
        # pass 1: symbolic (c, h, w) of every layer output
        shp = []
        for i, m in enumerate(layers):
            f = m.f
            src = ((3, self.H, self.W) if i == 0 else shp[i - 1]) if f == -1 else (
                shp[f] if isinstance(f, int) else [shp[i - 1] if j == -1 else shp[j] for j in f])
            shp.append(self._shape_of(m, src))
        # pass 2: pre-assign concat members to slices of the concat's buffer
        assigned: dict[int, View] = {}
        concat_buf: dict[int, View] = {}
        for i, m in enumerate(layers):
            if isinstance(m, mc.Concat):
                if m.d != 1:
                    raise NotImplementedError("y5b200: Concat along a dimension other than channels")
                c, h, w = shp[i]
                cat = self.new_view(h, w, c)
                concat_buf[i] = cat
                off = 0
                for j in m.f:
                    src = i - 1 if j == -1 else j
                    cj = shp[src][0]
                    if src not in assigned and cj % 8 == 0 and off % 8 == 0:
                        assigned[src] = cat.slice(off, cj)
                    off += cj
        # pass 3: emit
        outs: list = [None] * n
        x = None
        for i, m in enumerate(layers):
            f = m.f
            name = f"model.{i}"
            if isinstance(m, (my.Detect,)):
                xs = [outs[j] for j in f]
                if isinstance(m, my.Segment):
                    self.proto_view = self.lower_proto(m.proto, xs[0], f"{name}.proto")
                self.lower_detect(m, xs, name)
                continue
            if isinstance(m, mc.Concat):
                cat = concat_buf[i]
                off = 0
                for j in f:
                    src = i - 1 if j == -1 else j
                    v = outs[src]
                    want = cat.slice(off, v.c)
                    if not (v.buf is cat.buf and v.coff == want.coff):
                        self.copy_into(v, want, f"{name}.copy{src}")
                    off += v.c
                outs[i] = cat
                continue
            xin = (torch.empty(0) if i == 0 else outs[i - 1]) if f == -1 else outs[f]
            outs[i] = self._lower(m, xin, assigned.get(i), name)
        self.outs = outs

    def _shape_of(self, m, src):
        from .models import common as mc
        from .models import yolo as my

        if isinstance(m, mc.Conv):
            c, h, w = src
            k, s, p = m.conv.kernel_size[0], m.conv.stride[0], m.conv.padding[0]
            return (m.conv.out_channels, (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1)
        if isinstance(m, mc.C3):
            return (m.cv3.conv.out_channels, src[1], src[2])
        if isinstance(m, mc.SPPF):
            return (m.cv2.conv.out_channels, src[1], src[2])
        if isinstance(m, torch.nn.Upsample):
            return (src[0], src[1] * 2, src[2] * 2)
        if isinstance(m, mc.Concat):
            return (sum(s[0] for s in src), src[0][1], src[0][2])
        if isinstance(m, my.Detect):
            return None
        raise NotImplementedError(f"y5b200: module {type(m).__name__} is outside the engine's hot path")

    # ------------------------------------------------------------------ execution
    def _run_fixed(self, stream: int):
        for op in self.ops:
            op.run(stream)

    def capture(self):
        """Capture the fixed part into a CUDA graph (input pointer independent: starts after the stem s2d)."""
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self._run_fixed(s.cuda_stream)  # warm-up outside capture (lazy module loading, attribute setup)
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            self._run_fixed(torch.cuda.current_stream(self.device).cuda_stream)
        self.graph = g

    def run_model(self, img: torch.Tensor, use_graph: bool = True):
        """img: (B,3,H,W) NCHW uint8 (scaled by 1/255 on the fly) or fp16/bf16/fp32 in [0,1]."""
        assert img.is_cuda and img.shape == (self.B, 3, self.H, self.W), (img.shape, (self.B, 3, self.H, self.W))
        if not img.is_contiguous():
            img = img.contiguous()
        st = _lib.stream_ptr(self.device)
        _lib.check(self.lib.y5_stem_s2d(img.data_ptr(), _lib.dtype_code(img.dtype), self.stem_in.data_ptr(), self.dt_code, self.B,
                                        self.H, self.W, self.W // 2 + 2, 1, C.c_void_p(st)), "stem_s2d")
        if use_graph and os.environ.get("Y5_NO_GRAPH") != "1":
            if self.graph is None:
                self.capture()
            self.graph.replay()
        else:
            self._run_fixed(st)
        # head: fresh output tensors every call, like the reference
        no = self.det_shapes[0][-1]
        z = torch.empty(self.B, self.z_rows, no, dtype=self.dtype, device=self.device)
        raws = [torch.empty(s, dtype=self.dtype, device=self.device) for s in self.det_shapes]
        for plan, raw in zip(self.head_ops, raws):
            _lib.check(self.lib.y5_detect_plan_run_to(plan, raw.data_ptr(), z.data_ptr(), C.c_void_p(st)), "detect")
        proto = None
        if self.proto_view is not None:
            pv = self.proto_view
            proto = torch.empty(self.B, pv.c, pv.h, pv.w, dtype=self.dtype, device=self.device)
            _lib.check(self.lib.y5_nhwc_to_nchw(pv.ptr, pv.pitch, proto.data_ptr(), self.B, pv.h, pv.w, pv.c, self.dt_code,
                                                C.c_void_p(st)), "nhwc_to_nchw")
        return z, raws, proto

    def run_layer(self, x: torch.Tensor) -> torch.Tensor:
        """Single-layer program: x (B,C,H,W) NCHW -> (B,C2,H2,W2) NCHW.  The NCHW->NHWC staging of the input is a
        torch copy (plumbing); the layer itself and the NHWC->NCHW export run in liby5b200."""
        v = self.layer_in
        v.buf[..., : v.c].copy_(x.permute(0, 2, 3, 1))
        st = _lib.stream_ptr(self.device)
        self._run_fixed(st)
        o = self.single_out
        y = torch.empty(self.B, o.c, o.h, o.w, dtype=self.dtype, device=self.device)
        _lib.check(self.lib.y5_nhwc_to_nchw(o.ptr, o.pitch, y.data_ptr(), self.B, o.h, o.w, o.c, self.dt_code, C.c_void_p(st)),
                   "nhwc_to_nchw")
        return y

    def launches_per_forward(self) -> int:
        n = len(self.ops) + len(self.head_ops) + (1 if self.stem_in is not None else 0) + (1 if self.proto_view is not None else 0)
        return n

    def __del__(self):
        try:
            for fn, h in self._plans:
                fn(h)
        except Exception:
            pass
