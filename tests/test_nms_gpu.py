"""GPU: y5_nms_batched (through yolov5_b200.utils.general.non_max_suppression) vs the oracle -- bit-exact rows and
indices -- on the golden regimes, plus size-independent properties at the full BASELINE size (bs=32, 25200x85)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import nms_ref
from yolov5_b200.utils.general import non_max_suppression
from yolov5_b200.utils.metrics import box_iou

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
TDT = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}


def _run(pred_np, dtype, dev, **kw):
    t = torch.from_numpy(pred_np).to(dev, TDT[dtype])
    out, idx = non_max_suppression(t, return_indices=True, **kw)
    return [o.cpu().numpy() for o in out], [i.cpu().numpy() for i in idx]


def test_golden_regimes_bit_exact(cuda):
    g = np.load(os.path.join(G, "nms.npz"))
    for c in json.loads(str(g["meta"])):
        if c.get("labels"):
            continue  # test_apriori_labels_bit_exact
        pred = nms_ref.synth_predictions(c["bs"], c["n"], c["nc"], c["nm"], c["seed"], c["dtype"])
        got, gidx = _run(pred, c["dtype"], cuda, **c["kw"])
        _, oidx = nms_ref.non_max_suppression(pred, dtype=c["dtype"], return_index=True, **c["kw"])
        for b in range(c["bs"]):
            ref = g[f"{c['tag']}.{b}"]
            assert got[b].shape == ref.shape, (c["tag"], b, got[b].shape, ref.shape)
            assert np.array_equal(gidx[b], oidx[b]), (c["tag"], b, "indices")
            assert np.array_equal(got[b], ref), (c["tag"], b, np.abs(got[b] - ref).max())


def test_apriori_labels_bit_exact(cuda):
    """`labels=` (val.py --save-hybrid, reference utils/general.py:706-712): rows and candidate ids vs the oracle and the
    reference-generated fixture, fp16 and fp32 inputs (an image with labels is processed in fp32 like the reference's cat)."""
    g = np.load(os.path.join(G, "nms.npz"))
    for c in [m for m in json.loads(str(g["meta"])) if m.get("labels")]:
        pred = nms_ref.synth_predictions(c["bs"], c["n"], c["nc"], c["nm"], c["seed"], c["dtype"])
        labels = [g[f"{c['tag']}.labels{b}"] for b in range(c["bs"])]
        t = torch.from_numpy(pred).to(cuda, TDT[c["dtype"]])
        out, idx = non_max_suppression(t, labels=[torch.from_numpy(l).to(cuda) for l in labels], return_indices=True, **c["kw"])
        _, oidx = nms_ref.non_max_suppression(pred, dtype=c["dtype"], labels=labels, return_index=True, **c["kw"])
        for b in range(c["bs"]):
            assert np.array_equal(idx[b].cpu().numpy(), oidx[b]), (c["tag"], b)
            assert np.array_equal(out[b].cpu().numpy(), g[f"{c['tag']}.{b}"]), (c["tag"], b)


@pytest.mark.parametrize("dtype", ["fp16", "fp32"])
def test_edge_cases_vs_oracle(cuda, dtype):
    rs = np.random.RandomState(3)
    cases = []
    # all rows identical boxes / identical scores (pure tie-break + full suppression)
    p = np.zeros((2, 700, 85), np.float32)
    p[..., 0:2] = 100; p[..., 2:4] = 50; p[..., 4] = 0.9; p[..., 5] = 0.8
    cases.append((p, dict(conf_thres=0.25, iou_thres=0.45)))
    # far apart boxes, more survivors than max_det
    p = np.zeros((1, 900, 85), np.float32)
    p[0, :, 0] = np.arange(900) * 60.0; p[0, :, 1] = 30; p[0, :, 2:4] = 20
    p[0, :, 4] = rs.uniform(0.3, 1, 900); p[0, :, 5 + 7] = rs.uniform(0.5, 1, 900)
    cases.append((p, dict(conf_thres=0.25, iou_thres=0.45, max_det=100)))
    # multi-label with very many candidates (exercises the max_nms=30000 selection incl. ties on the cut)
    p = nms_ref.synth_predictions(1, 8000, 80, 0, 11, dtype)
    p[..., 4] = np.maximum(p[..., 4], 0.5)
    p[..., 5:] = nms_ref.round_to(np.round(p[..., 5:] * 16) / 16, dtype)  # few distinct scores -> ties across the cut
    cases.append((nms_ref.round_to(p, dtype), dict(conf_thres=0.01, iou_thres=0.6, multi_label=True, max_det=300)))
    # zero-area boxes (NaN IoU never suppresses) and a single row
    p = np.zeros((1, 64, 85), np.float32); p[..., 0:2] = 10; p[..., 4] = 0.9; p[..., 6] = 0.9
    cases.append((p, dict(conf_thres=0.25, iou_thres=0.45)))
    cases.append((nms_ref.synth_predictions(1, 1, 80, 0, 12, dtype), dict(conf_thres=0.0, iou_thres=0.5)))
    for p, kw in cases:
        p = nms_ref.round_to(p, dtype)
        got, gidx = _run(p, dtype, cuda, **kw)
        ref, ridx = nms_ref.non_max_suppression(p, dtype=dtype, return_index=True, **kw)
        for b in range(p.shape[0]):
            assert np.array_equal(gidx[b], ridx[b]), kw
            assert np.array_equal(got[b], ref[b]), kw


def test_full_size_properties_bs32(cuda):
    """BASELINE config 2 size: (32, 25200, 85) fp16.  Oracle on 3 images; for all 32: sortedness, max_det bound,
    class-wise IoU <= threshold among survivors, idempotence of NMS on its own output set."""
    pred = nms_ref.synth_predictions(32, 25200, 80, 0, 21, "fp16")
    kw = dict(conf_thres=0.25, iou_thres=0.45, max_det=300)
    got, gidx = _run(pred, "fp16", cuda, **kw)
    ref, ridx = nms_ref.non_max_suppression(pred[:3], dtype="fp16", return_index=True, **kw)
    for b in range(3):
        assert np.array_equal(gidx[b], ridx[b]) and np.array_equal(got[b], ref[b])
    for b in range(32):
        d = got[b]
        assert d.shape[0] <= 300 and d.shape[1] == 6
        assert np.all(np.diff(d[:, 4]) <= 0)                      # score-descending
        assert len(np.unique(gidx[b])) == len(gidx[b])            # no duplicate candidates
        boxes = torch.from_numpy(d[:, :4] + d[:, 5:6] * 7680).to(cuda)
        iou = box_iou(boxes, boxes).cpu().numpy()
        np.fill_diagonal(iou, 0)
        assert iou.max() <= 0.45 + 1e-6


def test_box_iou_vs_oracle(cuda):
    rs = np.random.RandomState(5)
    a = rs.uniform(0, 300, (37, 2)).astype(np.float32); b = rs.uniform(0, 300, (91, 2)).astype(np.float32)
    A = np.concatenate((a, a + rs.uniform(0, 80, (37, 2)).astype(np.float32)), 1)
    Bx = np.concatenate((b, b + rs.uniform(0, 80, (91, 2)).astype(np.float32)), 1)
    got = box_iou(torch.from_numpy(A).to(cuda), torch.from_numpy(Bx).to(cuda)).cpu().numpy()
    assert np.array_equal(got, nms_ref.box_iou(A, Bx))
    assert box_iou(torch.zeros(0, 4, device=cuda), torch.from_numpy(Bx).to(cuda)).shape == (0, 91)
