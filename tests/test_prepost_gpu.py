"""GPU: the callers either side of the hot path through the C ABI vs their oracles and the
reference-generated fixtures: letterbox pre-processing (bytes bit-exact), process_mask (3 modes), scale_boxes, the batched
metric matching of val.py:282-318 (match matrices bit-exact), and the apriori-label / compat / TTA API edges."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import nms_ref, post_ref, pre_ref
from tests.golden.make_golden_cases import PRE_CASES
from yolov5_b200.utils.augmentations import letterbox, letterbox_batch
from yolov5_b200.utils.general import nms_device, scale_boxes, scale_boxes_batch, scale_meta
from yolov5_b200.utils.metrics import labels_to_native, match_batch, process_batch, val_batch_metrics
from yolov5_b200.utils.segment.general import crop_mask, process_mask, process_mask_batch, process_mask_native

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


# ----------------------------------------------------------------------------------------------------------------- pre
def test_letterbox_bytes_equal_reference_fixture(cuda):
    g = np.load(os.path.join(G, "pre.npz"))
    for i, (h, w, seed, kw) in enumerate(PRE_CASES):
        im = pre_ref.synth_image(h, w, seed)
        out, ratio, pad = letterbox(im, **kw)                       # reference signature, numpy in -> numpy out
        ref, r_ratio, r_pad = pre_ref.letterbox(im, **kw)
        assert np.array_equal(out, ref) and tuple(ratio) == tuple(r_ratio) and tuple(pad) == tuple(r_pad), i
        batch, _, _ = letterbox_batch([torch.from_numpy(im).to(cuda)], auto=kw.get("auto", True), swap_rb=True,
                                      **{k: v for k, v in kw.items() if k != "auto"})
        assert np.array_equal(batch[0].cpu().numpy(), g[f"lb{i}"]), i  # CHW RGB bytes the reference dataloader yields


def test_letterbox_batch_mixed_sizes_all_outputs(cuda):
    """One launch, images of different sizes (640-class canvas, BASELINE shapes): uint8 bytes bit-exact; float outputs ==
    bytes / 255 rounded once; the fused space-to-depth output == y5_stem_s2d of the uint8 batch."""
    import ctypes as C

    from yolov5_b200 import _lib

    rs = np.random.RandomState(3)
    sizes = [(480, 640), (375, 500), (720, 1280), (640, 640), (333, 217), (101, 640)] * 5  # 30 images: two launches of <= 24
    ims = [rs.randint(0, 256, (h, w, 3)).astype(np.uint8) for h, w in sizes]
    dev_ims = [torch.from_numpy(im).to(cuda) for im in ims]
    u8, ratios, pads = letterbox_batch(dev_ims, (640, 640), auto=False)
    for i, im in enumerate(ims):
        ref, r, p = pre_ref.letterbox(im, (640, 640), auto=False)
        assert np.array_equal(u8[i].cpu().numpy(), pre_ref.to_chw_rgb(ref)), i
        assert tuple(ratios[i]) == tuple(r) and tuple(pads[i]) == tuple(p)
    # bytes / 255 as a true fp32 division (what the CPU reference and numpy compute; torch-CUDA multiplies by fp32(1/255), which
    # differs in the last fp32 bit for 126 of the 256 byte values and in none of them after rounding to fp16 / bf16)
    want32 = torch.from_numpy(u8.cpu().numpy().astype(np.float32) / np.float32(255))
    for dt in (torch.float16, torch.bfloat16, torch.float32):
        f, _, _ = letterbox_batch(dev_ims, (640, 640), auto=False, dtype=dt)
        assert torch.equal(f.cpu(), want32.to(dt)), dt
    lib = _lib.lib()
    for dt in (torch.float16, torch.bfloat16):
        want = torch.zeros(len(ims), 320, 322, 16, dtype=dt, device=cuda)
        _lib.check(lib.y5_stem_s2d(u8.data_ptr(), _lib.Y5_U8, want.data_ptr(), _lib.dtype_code(dt), len(ims), 640, 640, 322, 1,
                                   C.c_void_p(_lib.stream_ptr(cuda))))
        got = torch.zeros_like(want)
        letterbox_batch(dev_ims, (640, 640), auto=False, s2d_out=(got, 322, 1))
        assert torch.equal(got, want)


# ---------------------------------------------------------------------------------------------------------------- post
def _unpack(g, key):
    shape = tuple(int(v) for v in g[f"{key}.shape"])
    return np.unpackbits(g[key])[: int(np.prod(shape))].reshape(shape).astype(np.float32)


def test_process_mask_vs_fixture_and_oracle(cuda):
    g = np.load(os.path.join(G, "post.npz"))
    protos, coef, boxes = (torch.from_numpy(g[k]).to(cuda) for k in ("mask.protos", "mask.coef", "mask.boxes"))
    hw = tuple(int(v) for v in g["mask.input_hw"])
    for up in (False, True):
        got = process_mask(protos, coef, boxes, hw, upsample=up).cpu().numpy()
        ref = _unpack(g, f"mask.up{int(up)}")
        _, val = post_ref.process_mask(g["mask.protos"], g["mask.coef"], g["mask.boxes"], hw, upsample=up)
        off = got != ref
        assert got.shape == ref.shape and (not off.any() or np.abs(val[off] - 0.5).max() < 1e-5), (up, int(off.sum()))
    for tag, shp in (("native", (160, 224)), ("native_pad", (128, 224))):
        got = process_mask_native(protos, coef, boxes, shp).cpu().numpy()
        ref = _unpack(g, f"mask.{tag}")
        _, val = post_ref.process_mask_native(g["mask.protos"], g["mask.coef"], g["mask.boxes"], shp)
        off = got != ref
        assert got.shape == ref.shape and (not off.any() or np.abs(val[off] - 0.5).max() < 1e-5), (tag, int(off.sum()))
    m = torch.rand(9, 40, 56, device=cuda)
    assert np.array_equal(crop_mask(m, boxes / 4).cpu().numpy(), post_ref.crop_mask(m.cpu().numpy(), g["mask.boxes"] / 4))


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_process_mask_batched_config5_shape(cuda, dtype):
    """BASELINE config 5 geometry: 1280x1280 input -> 32 prototypes of 320x320 per image, 2 images, strided views into NMS
    rows (coefficients = columns 6.., boxes = columns 0..3), uint8 and float outputs, with and without up-sampling."""
    rs = np.random.RandomState(11)
    protos = torch.from_numpy(rs.randn(2, 32, 320, 320).astype(np.float32)).to(cuda, dtype)
    n_per = (37, 21)
    rows = np.zeros((2, 64, 38), np.float32)
    for b, n in enumerate(n_per):
        xy = rs.uniform(0, 900, (n, 2)); wh = rs.uniform(40, 380, (n, 2))
        rows[b, :n, 0:2], rows[b, :n, 2:4] = xy, xy + wh
        rows[b, :n, 6:] = rs.randn(n, 32) * 0.5
    rows_t = torch.from_numpy(rows).to(cuda)
    flat = torch.cat([rows_t[b, :n] for b, n in enumerate(n_per)])
    idx = torch.cat([torch.full((n,), b, dtype=torch.int32) for b, n in enumerate(n_per)]).to(cuda)
    for up in (False, True):
        got = process_mask_batch(protos, flat[:, 6:], flat[:, :4], idx, (1280, 1280), upsample=up, out_dtype=torch.uint8).cpu().numpy()
        off0 = 0
        for b, n in enumerate(n_per):
            ref, val = post_ref.process_mask(protos[b].float().cpu().numpy(), rows[b, :n, 6:], rows[b, :n, :4], (1280, 1280), upsample=up)
            d = got[off0 : off0 + n].astype(np.float32) != ref
            assert not d.any() or np.abs(val[d] - 0.5).max() < 2e-5, (up, b, int(d.sum()))
            off0 += n


def test_scale_boxes_and_labels_native_bit_exact(cuda):
    g = np.load(os.path.join(G, "post.npz"))
    b = torch.from_numpy(g["scale.in"].copy()).to(cuda)
    assert np.array_equal(scale_boxes((640, 640), b.clone(), (480, 640)).cpu().numpy(), post_ref.scale_boxes((640, 640), g["scale.in"], (480, 640)))
    rp = ((0.75, 0.75), (10.0, 80.0))
    det = torch.zeros(50, 6, device=cuda); det[:, :4] = b
    out = scale_boxes((640, 640), det[:, :4], (480, 640), rp)      # a strided view, edited in place like the reference does
    assert np.array_equal(det[:, :4].cpu().numpy(), post_ref.scale_boxes((640, 640), g["scale.in"], (480, 640), rp)) and out.data_ptr() == det.data_ptr()
    assert np.allclose(det[:, :4].cpu().numpy(), g["scale.given"], rtol=0, atol=1e-4)
    # labels: xywh2xyxy + scale_boxes for a batch of 3 images with different shapes
    rs = np.random.RandomState(2)
    shapes0 = [(480, 640), (375, 500), (640, 427)]
    meta = scale_meta((640, 640), shapes0, [None, ((1.28, 1.28), (0.0, 80.0)), None])
    tg = np.concatenate((rs.randint(0, 3, (40, 1)), rs.randint(0, 80, (40, 1)), rs.uniform(20, 600, (40, 2)), rs.uniform(5, 300, (40, 2))), 1).astype(np.float32)
    got = labels_to_native(torch.from_numpy(tg).to(cuda), meta).cpu().numpy()
    for i in range(40):
        im = int(tg[i, 0])
        xyxy = np.array([tg[i, 2] - tg[i, 4] / np.float32(2), tg[i, 3] - tg[i, 5] / np.float32(2), tg[i, 2] + tg[i, 4] / np.float32(2),
                         tg[i, 3] + tg[i, 5] / np.float32(2)], np.float32)
        rp_i = [None, ((1.28, 1.28), (0.0, 80.0)), None][im]
        assert np.array_equal(got[i, 2:], post_ref.scale_boxes((640, 640), xyxy[None], shapes0[im], rp_i)[0]) and got[i, 0] == tg[i, 0] and got[i, 1] == tg[i, 1]


def test_process_batch_bit_exact_vs_fixture(cuda):
    g = np.load(os.path.join(G, "post.npz"))
    iouv = torch.from_numpy(g["match.iouv"]).to(cuda)
    for case in range(4):
        det, lab = g[f"match{case}.det"], g[f"match{case}.labels"]
        got = process_batch(torch.from_numpy(det).to(cuda), torch.from_numpy(lab).to(cuda), iouv)
        assert got.dtype == torch.bool and got.device == iouv.device
        assert np.array_equal(got.cpu().numpy(), g[f"match{case}.correct"]), case


def test_val_loop_batched_vs_per_image_oracle(cuda):
    """val.py:282-318 for a whole batch: NMS rows -> scale_boxes -> xywh2xyxy/scale labels -> process_batch, one launch set and
    no host round trip, vs the oracle run image by image.  Detections duplicated / tied on purpose (index tie rules)."""
    rs = np.random.RandomState(9)
    B, max_det = 6, 300
    pred = nms_ref.synth_predictions(B, 6300, 80, 0, 31, "fp16")
    rows, _, count = nms_device(torch.from_numpy(pred).to(cuda).half(), 0.001, 0.6, multi_label=True, max_det=max_det)
    shapes0 = [(480, 640), (375, 500), (640, 427), (640, 640), (500, 333), (427, 640)]
    shapes = []
    for h0, w0 in shapes0:  # what the reference's val dataloader yields: ((h0, w0), ((h/h0, w/w0), pad))
        r = min(640 / h0, 640 / w0)
        nh, nw = round(h0 * r), round(w0 * r)
        shapes.append(((h0, w0), ((nh / h0, nw / w0), ((640 - nw) / 2, (640 - nh) / 2))))
    cnt = count.cpu().numpy()
    tg = []
    for b in range(B):  # labels near actual detections so that matches exist; a few exact duplicates
        n = int(cnt[b])
        pick = rs.choice(n, size=min(n, 12), replace=False) if n else []
        d = rows[b].cpu().numpy()[pick]
        xywh = np.stack(((d[:, 0] + d[:, 2]) / 2, (d[:, 1] + d[:, 3]) / 2, d[:, 2] - d[:, 0], d[:, 3] - d[:, 1]), 1) + rs.normal(0, 2, (len(pick), 4))
        t = np.concatenate((np.full((len(pick), 1), b), d[:, 5:6], xywh), 1)
        tg.append(np.concatenate((t, t[:2])))
    targets = np.concatenate(tg).astype(np.float32)
    iouv = np.linspace(0.5, 0.95, 10).astype(np.float32)
    predn, correct = val_batch_metrics(rows, count, torch.from_numpy(targets).to(cuda), (640, 640), shapes, torch.from_numpy(iouv).to(cuda))
    assert correct.shape == (B, max_det, 10) and correct.dtype == torch.bool
    rows_np, predn_np, correct_np = rows.cpu().numpy(), predn.cpu().numpy(), correct.cpu().numpy()
    for b in range(B):
        n = int(cnt[b])
        ref_boxes = post_ref.scale_boxes((640, 640), rows_np[b, :n, :4], shapes[b][0], shapes[b][1])
        assert np.array_equal(predn_np[b, :n, :4], ref_boxes) and np.array_equal(predn_np[b, :n, 4:], rows_np[b, :n, 4:])
        lb = targets[targets[:, 0] == b, 1:]
        half = lb[:, 3:5] / np.float32(2)
        tbox = post_ref.scale_boxes((640, 640), np.concatenate((lb[:, 1:3] - half, lb[:, 1:3] + half), 1), shapes[b][0], shapes[b][1])
        labelsn = np.concatenate((lb[:, 0:1], tbox), 1)
        det = np.concatenate((ref_boxes, rows_np[b, :n, 4:6]), 1)
        assert np.array_equal(correct_np[b, :n], post_ref.process_batch(det, labelsn, iouv)), b
        assert not correct_np[b, n:].any()
    assert correct_np.any()
