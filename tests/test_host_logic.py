"""CPU: host-side logic of the engine -- model assembly parity with the reference's structure, weight folding /
packing, the stem space-to-depth identity, and the loud failure on non-CUDA inputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import model_ref
from yolov5_b200.cfg import model_cfg
from yolov5_b200.engine import fold_conv_bn, pack_weight, stem_weight_s2d
from yolov5_b200.models.common import C3, Conv
from yolov5_b200.models.yolo import DetectionModel, SegmentationModel
from yolov5_b200.utils.general import _iou_threshold_f32
from yolov5_b200.utils.torch_utils import fuse_conv_and_bn

PARAMS = {"yolov5n": 1872157, "yolov5s": 7235389, "yolov5m": 21190557, "yolov5l": 46563709, "yolov5x": 86749405,
          "yolov5x-seg": 88819517}  # SURVEY.md Appendix A (counted on the reference)


@pytest.mark.parametrize("name", list(PARAMS))
def test_parameter_counts_and_state_dict_keys(name):
    m = (SegmentationModel if name.endswith("-seg") else DetectionModel)(name)
    assert sum(p.numel() for p in m.parameters()) == PARAMS[name]
    assert list(m.state_dict().keys()) == list(model_ref.param_shapes(model_cfg(name)).keys())
    assert [float(s) for s in m.stride] == [8.0, 16.0, 32.0]
    assert m.save == [4, 6, 10, 14, 17, 20, 23]
    assert m.model[0].bn.eps == 1e-3 and m.model[0].bn.momentum == 0.03


def test_stem_space_to_depth_identity():
    """6x6/s2/p2 conv on 3 channels == 3x3/s1/p1 conv on the 2x2 space-to-depth tensor (12 of 16 channels used)."""
    torch.manual_seed(0)
    x = torch.rand(2, 3, 32, 48)
    w = torch.randn(8, 3, 6, 6)
    ref = F.conv2d(x, w, stride=2, padding=2)
    b, _, h, wd = x.shape
    s2d = torch.zeros(b, 16, h // 2, wd // 2)
    for dy in range(2):
        for dx in range(2):
            for c in range(3):
                s2d[:, (dy * 2 + dx) * 3 + c] = x[:, c, dy::2, dx::2]
    got = F.conv2d(s2d, stem_weight_s2d(w), stride=1, padding=1)
    assert torch.allclose(ref, got, atol=1e-4)


def test_fold_matches_oracle_and_fuse():
    torch.manual_seed(1)
    m = Conv(16, 24, 3, 1)
    m.bn.weight.data.uniform_(0.5, 1.5); m.bn.bias.data.normal_(0, 0.1)
    m.bn.running_mean.normal_(0, 0.1); m.bn.running_var.uniform_(0.5, 1.5)
    m.bn.eps = 1e-3
    w, b = fold_conv_bn(m.conv, m.bn)
    w2, b2 = model_ref.fold_bn(m.conv.weight.detach(), m.bn.weight.detach(), m.bn.bias.detach(), m.bn.running_mean, m.bn.running_var)
    assert torch.allclose(w, w2, atol=1e-6) and torch.allclose(b, b2, atol=1e-6)
    f = fuse_conv_and_bn(m.conv, m.bn)
    assert torch.allclose(f.weight, w2, atol=1e-6) and torch.allclose(f.bias, b2, atol=1e-6)
    x = torch.randn(1, 16, 8, 8)
    y_ref = F.silu(F.batch_norm(F.conv2d(x, m.conv.weight, None, 1, 1), m.bn.running_mean, m.bn.running_var, m.bn.weight, m.bn.bias, False, 0.0, 1e-3))
    assert torch.allclose(F.silu(f(x)), y_ref, atol=1e-5)


def test_pack_weight_layout():
    w = torch.arange(2 * 24 * 3 * 3, dtype=torch.float32).view(2, 24, 3, 3)
    p = pack_weight(w, 32, torch.float16)
    assert p.shape == (2, 3, 3, 32)
    assert torch.equal(p[1, 2, 0, :24].float(), w[1, :, 2, 0].half().float())
    assert torch.count_nonzero(p[..., 24:]) == 0


def test_iou_threshold_rule():
    for t in (0.45, 0.6, 0.5, 0.3, 0.65):
        f = np.float32(_iou_threshold_f32(t))
        assert float(f) <= t and float(np.nextafter(f, np.float32(np.inf))) > t


def test_cpu_tensors_fail_loudly():
    m = DetectionModel("yolov5n").eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="CUDA"):
        C3(16, 16).eval()(torch.zeros(1, 16, 8, 8))
    from yolov5_b200.utils.general import non_max_suppression
    from yolov5_b200.utils.metrics import box_iou
    with pytest.raises(RuntimeError, match="CUDA"):
        non_max_suppression(torch.zeros(1, 10, 85))
    with pytest.raises(RuntimeError, match="CUDA"):
        box_iou(torch.zeros(1, 4), torch.zeros(1, 4))


def test_module_pickles_without_engine_state():
    import pickle

    m = DetectionModel("yolov5n")
    m.__dict__["_y5_programs"] = {"x": object()}
    m.__dict__["_y5_pack_plans"] = {("cpu", torch.float16): object()}
    m2 = pickle.loads(pickle.dumps(m))
    assert "_y5_programs" not in m2.__dict__ and "_y5_pack_plans" not in m2.__dict__
    import copy

    assert "_y5_pack_plans" not in copy.deepcopy(m).__dict__  # ModelEMA's deepcopy starts without the training-path buffers
    assert list(m2.state_dict().keys()) == list(m.state_dict().keys())


# ---------------------------------------------------------------------------------------------------------------------
# host logic of the training path (no GPU): index maps, NHWC view detection, zero arena
# ---------------------------------------------------------------------------------------------------------------------
def test_stem_index_maps_match_stem_weight_s2d():
    """train_ops._stem_index must be the gather form of engine.stem_weight_s2d and its exact inverse on the 108 taps."""
    import torch

    from yolov5_b200 import train_ops
    from yolov5_b200.engine import stem_weight_s2d

    w = torch.arange(32 * 3 * 6 * 6, dtype=torch.float32).view(32, 3, 6, 6) + 1
    fwd, inv = train_ops._stem_index(torch.device("cpu"))
    wf = w.flatten(1)
    gathered = torch.cat((wf, wf.new_zeros(32, 1)), 1)[:, fwd].view(32, 16, 3, 3)
    assert torch.equal(gathered, stem_weight_s2d(w))
    assert torch.equal(gathered.reshape(32, -1)[:, inv].view(32, 3, 6, 6), w)  # the weight-gradient path back to (O,3,6,6)
    assert int((fwd == 108).sum()) == 36  # 4 unused channels x 9 taps read the appended zero


def test_nhwc_view_detection():
    import torch

    from yolov5_b200 import train_ops

    buf = torch.zeros(2, 5, 7, 48, dtype=torch.float16)  # NHWC memory
    x = buf.permute(0, 3, 1, 2)                            # logical NCHW, channels_last
    t, pitch = train_ops._nhwc(x)
    assert t.data_ptr() == x.data_ptr() and pitch == 48
    sl = x[:, 16:32]                                       # channel slice (what torch.cat's backward produces)
    t, pitch = train_ops._nhwc(sl)
    assert t.data_ptr() == sl.data_ptr() and pitch == 48 and t.shape[1] == 16
    odd = x[:, 4:20]                                       # 8-byte aligned only: must be copied to a dense tensor
    t, pitch = train_ops._nhwc(odd)
    assert t.data_ptr() != odd.data_ptr() and pitch == 16 and torch.equal(t, odd)
    nchw = torch.zeros(2, 16, 5, 7, dtype=torch.float16)   # NCHW-contiguous input gets a channels_last copy
    t, pitch = train_ops._nhwc(nchw)
    assert pitch == 16 and t.stride() == (5 * 7 * 16, 1, 7 * 16, 16)


def test_zero_arena_slices_are_disjoint_and_zero():
    import torch

    from yolov5_b200.train_ops import _ZeroArena

    a = _ZeroArena()
    dev = torch.device("cpu")
    first = a.take(10, dev)  # before any reset: a fresh zero tensor
    assert first.numel() == 10 and float(first.abs().sum()) == 0
    a.reset(dev)
    s1, s2 = a.take(6, dev), a.take(7, dev)
    assert s1.data_ptr() + 6 * 8 <= s2.data_ptr() and s2.numel() == 8  # rounded to 16 bytes
    s1.fill_(3.0)
    a.reset(dev)
    assert float(a.take(6, dev).abs().sum()) == 0  # cleared by the reset
    big = a.take(1 << 20, dev)  # beyond capacity: fallback allocation, and the arena grows at the next reset
    assert big.numel() == 1 << 20 and float(big.abs().sum()) == 0
    a.reset(dev)
    assert a.buf.numel() >= 1 << 20


def test_training_forward_rejects_cpu_and_fp32():
    import pytest
    import torch

    from yolov5_b200.models.yolo import DetectionModel

    m = DetectionModel("yolov5n").train()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 64, 64))


def test_pack_plan_host_logic(monkeypatch):
    """train_ops.PackPlan without a GPU: registration marks the table dirty, begin() builds one y5_pack_item per registered filter
    (pointers, shapes, pads), covers every element of [fwd | dgrad] with chunk entries exactly once, bumps the epoch, and lookup()
    only hands out buffers packed in the CURRENT forward with the same geometry."""
    import contextlib
    import ctypes

    from yolov5_b200 import _lib, train_ops

    calls = []

    class Fake:
        def y5_weight_pack_chunk_elems(self):
            return 1000

        def y5_weight_pack_multi(self, items, ci, cx, n, dtype, stream):
            calls.append((items, ci, cx, n, dtype))
            return 0

    monkeypatch.setattr(_lib, "lib", lambda: Fake())
    monkeypatch.setattr(_lib, "on", lambda dev: contextlib.nullcontext())
    monkeypatch.setattr(train_ops, "_st", lambda dev: None)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    dev = torch.device("cpu")
    plan = train_ops.PackPlan(dev, torch.float16)
    w1 = torch.nn.Parameter(torch.randn(24, 16, 3, 3))
    w2 = torch.nn.Parameter(torch.randn(8, 24, 1, 1))
    f1, d1 = torch.empty(24, 3, 3, 64, dtype=torch.float16), torch.empty(16, 3, 3, 64, dtype=torch.float16)
    f2 = torch.empty(8, 1, 1, 64, dtype=torch.float16)
    plan.begin()  # nothing registered: no launch
    assert not calls and plan.lookup(w1, 64, 64, True) is None
    plan.register(w1, f1, d1, 64, 64, 64, 64)
    plan.register(w2, f2, None, 64, 0, 64, 0)
    assert plan.dirty and plan.lookup(w1, 64, 64, True) is None  # registered during this forward: packed by the layer itself
    plan.begin()
    assert len(calls) == 1 and not plan.dirty
    items, ci, cx, n, ents = plan.table
    arr = (_lib.PackItem * 2).from_buffer_copy(items.numpy().tobytes())
    assert (arr[0].w, arr[0].fwd, arr[0].dgrad) == (w1.data_ptr(), f1.data_ptr(), d1.data_ptr())
    assert (arr[0].out_c, arr[0].in_c, arr[0].ksize, arr[0].in_c_pad, arr[0].out_c_pad) == (24, 16, 3, 64, 64)
    assert (arr[1].w, arr[1].fwd, arr[1].dgrad) == (w2.data_ptr(), f2.data_ptr(), None) and arr[1].w_dtype == _lib.Y5_F32
    total = [24 * 9 * 64 + 16 * 9 * 64, 8 * 64]
    per_item = [[int(x) for t, x in zip(ci.tolist(), cx.tolist()) if t == i] for i in range(2)]
    for i in range(2):
        assert per_item[i] == list(range((total[i] + 999) // 1000))  # chunks 0..ceil(total/1000)-1, each once
    assert n == sum(len(v) for v in per_item) == calls[0][3]
    assert plan.lookup(w1, 64, 64, True) == (f1, d1) and plan.lookup(w2, 64, 0, False) == (f2, None)
    assert plan.lookup(w2, 64, 64, True) is None      # a data gradient is wanted but was never packed
    assert plan.lookup(w1, 16, 64, True) is None      # another K-block geometry
    w1.data = torch.randn(24, 16, 3, 3)               # storage replaced (e.g. .to()): stale pointer must not be used
    assert plan.lookup(w1, 64, 64, True) is None
    plan.begin()
    assert len(plan.table[4]) == 1                    # the stale entry left the table, w2 stays
    assert ctypes.sizeof(_lib.PackItem) == 48


def test_bn_backward_algebra_of_the_kernels_matches_autograd():
    """The two-pass form the CUDA kernels use for the backward of z = SiLU(BN_batchstats(y)) (train_kernels.cu: the reduce pass
    forms du = dz * silu'(t) and its column sums, the apply pass dy = du*a + y*c1 + c0 from four per-channel constants),
    restated in float64 and compared with torch autograd of the reference expression (models/common.py:86-88)."""
    torch.manual_seed(7)
    rows, c, eps = 257, 24, 1e-3
    y = torch.randn(rows, c, dtype=torch.float64, requires_grad=True)
    gamma = torch.randn(c, dtype=torch.float64, requires_grad=True)
    beta = torch.randn(c, dtype=torch.float64, requires_grad=True)
    dz = torch.randn(rows, c, dtype=torch.float64)
    z = F.silu(F.batch_norm(y, None, None, gamma, beta, training=True, eps=eps))
    z.backward(dz)
    with torch.no_grad():
        mean, var = y.mean(0), y.var(0, unbiased=False)
        invstd = (var + eps).rsqrt()
        a = invstd * gamma                      # constants of the kernels' tables
        b = beta - mean * a
        t = y * a + b                           # BN output
        sg = torch.sigmoid(t)
        du = dz * sg * (1 + t * (1 - sg))       # reduce pass: du, sum du, sum du * xhat
        xhat = y * invstd - mean * invstd
        dbeta, dgamma = du.sum(0), (du * xhat).sum(0)
        c1 = -invstd * (dgamma / rows) * a      # apply pass
        c0 = -(dbeta / rows + (-mean * invstd) * (dgamma / rows)) * a
        dy = du * a + y * c1 + c0
    assert torch.allclose(dy, y.grad, rtol=1e-10, atol=1e-12)
    assert torch.allclose(dgamma, gamma.grad, rtol=1e-10, atol=1e-12) and torch.allclose(dbeta, beta.grad, rtol=1e-10, atol=1e-12)
