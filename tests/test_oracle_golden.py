"""CPU: the oracle (oracle/*.py, oracle_c.c) against the golden fixtures produced by the real reference
(tests/golden/make_golden.py).  This is what pins the oracle; the GPU parity tests then compare the CUDA path with
the oracle."""
import hashlib
import json
import os

import numpy as np
import torch

from oracle import loss_ref, model_ref, nms_ref
from yolov5_b200.cfg import HYP_SCRATCH_LOW, model_cfg, model_names

G = os.path.join(os.path.dirname(__file__), "golden")


def test_cfg_tables_match_reference_yaml_digest():
    ref = json.load(open(os.path.join(G, "cfg_digest.json")))
    assert set(ref) == set(model_names())
    for name, digest in ref.items():
        got = hashlib.sha256(json.dumps(model_cfg(name), sort_keys=True).encode()).hexdigest()
        assert got == digest, name


def _image(shape, seed):
    return torch.from_numpy(np.random.RandomState(seed).uniform(0, 1, shape).astype(np.float32))


def test_model_forward_oracle_vs_reference_outputs():
    g = np.load(os.path.join(G, "model_forward.npz"))
    for name in ("yolov5n", "yolov5s", "yolov5n-seg"):
        cfg = model_cfg(name)
        seed_w, seed_x = (int(v) for v in g[f"{name}.seed"])
        sd = model_ref.synth_state_dict(cfg, seed=seed_w)
        x = _image(tuple(int(v) for v in g[f"{name}.shape"]), seed_x)
        for tag, fused in (("bn", False), ("fused", True)):
            with torch.no_grad():
                out = model_ref.forward(cfg, sd, x, fused=fused)
            z = out[0].numpy()
            raws = out[2] if name.endswith("-seg") else out[1]
            # fp32 CPU conv results may differ in the last bits between machines (threading / mkldnn blocking)
            np.testing.assert_allclose(z, g[f"{name}.{tag}.z"], rtol=2e-4, atol=2e-4)
            for l, r in enumerate(raws):
                np.testing.assert_allclose(r.numpy(), g[f"{name}.{tag}.raw{l}"], rtol=2e-4, atol=2e-4)
            if name.endswith("-seg"):
                np.testing.assert_allclose(out[1].numpy(), g[f"{name}.{tag}.proto"], rtol=2e-4, atol=2e-4)


def test_model_forward_640_config1_sample():
    """BASELINE.json configs[0]: yolov5n, 1x3x640x640, CPU fp32."""
    g = np.load(os.path.join(G, "model_forward.npz"))
    cfg = model_cfg("yolov5n")
    sw, sx = (int(v) for v in g["yolov5n.640.seed"])
    sd = model_ref.synth_state_dict(cfg, seed=sw, head_bias="hot")
    with torch.no_grad():
        z = model_ref.forward(cfg, sd, _image((1, 3, 640, 640), sx), fused=True)[0]
    assert z.shape == (1, 25200, 85)
    np.testing.assert_allclose(z[0, ::97].numpy(), g["yolov5n.640.z_sample"], rtol=2e-4, atol=2e-4)
    s = g["yolov5n.640.z_sum"]
    assert abs(z.double().sum().item() - s[0]) <= 1e-5 * s[1]


def test_nms_oracle_bit_exact_vs_golden():
    g = np.load(os.path.join(G, "nms.npz"))
    meta = json.loads(str(g["meta"]))
    assert len(meta) >= 9
    for c in meta:
        if c.get("labels"):
            continue  # apriori-label regimes have their own test (they need the label arrays)
        pred = nms_ref.synth_predictions(c["bs"], c["n"], c["nc"], c["nm"], c["seed"], c["dtype"])
        out = nms_ref.non_max_suppression(pred, dtype=c["dtype"], **c["kw"])
        for b, o in enumerate(out):
            ref = g[f"{c['tag']}.{b}"]
            assert o.shape == ref.shape, (c["tag"], b)
            assert np.array_equal(o, ref), (c["tag"], b)


def test_nms_greedy_properties():
    rs = np.random.RandomState(0)
    xy = rs.uniform(0, 100, (300, 2)).astype(np.float32)
    wh = rs.uniform(1, 40, (300, 2)).astype(np.float32)
    boxes = np.concatenate((xy, xy + wh), 1)
    keep = nms_ref.nms_greedy(boxes, 0.5)
    assert keep[0] == 0 and np.all(np.diff(keep) > 0)  # first box always kept, order preserved
    iou = nms_ref.box_iou(boxes[keep], boxes[keep])
    np.fill_diagonal(iou, 0)
    assert iou.max() <= 0.5 + 1e-6  # survivors do not suppress each other
    assert np.array_equal(nms_ref.nms_greedy(boxes[keep], 0.5), np.arange(len(keep)))  # idempotent
    assert len(nms_ref.nms_greedy(boxes, 0.5, max_keep=7)) == 7
    assert np.array_equal(nms_ref.nms_greedy(np.zeros((0, 4), np.float32), 0.5), np.zeros(0, np.int64))


def test_round_to_matches_torch():
    x = np.random.RandomState(1).normal(0, 10, 5000).astype(np.float32)
    t = torch.from_numpy(x)
    assert np.array_equal(nms_ref.round_to(x, "fp16"), t.half().float().numpy())
    assert np.array_equal(nms_ref.round_to(x, "bf16"), t.bfloat16().float().numpy())


def test_loss_oracle_vs_golden():
    g = np.load(os.path.join(G, "loss.npz"))
    cfg = model_cfg("yolov5n")
    anchors = model_ref.synth_state_dict(cfg, seed=30)["model.24.anchors"].numpy()
    for tag in ("a", "b", "none"):
        bs, h, w, seed = (int(v) for v in g[f"{tag}.meta"])
        rs = np.random.RandomState(seed)
        p = [torch.from_numpy(rs.normal(0, 1.5, (bs, 3, h // s, w // s, 85)).astype(np.float32)).requires_grad_(True) for s in (8, 16, 32)]
        tg = loss_ref.synth_targets(bs, seed) if tag != "none" else np.zeros((0, 6), np.float32)
        bt = loss_ref.build_targets(tg, anchors, [tuple(t.shape[2:4]) for t in p], 4.0)
        for i in range(3):
            ref = g[f"{tag}.idx{i}"]
            got = np.stack([bt[i][k] for k in ("b", "a", "gj", "gi", "tcls")])
            assert np.array_equal(got, ref), (tag, i)  # integer result: bit exact
            assert np.array_equal(bt[i]["tbox"], g[f"{tag}.tbox{i}"])
        loss, items = loss_ref.compute_loss(p, tg, anchors, HYP_SCRATCH_LOW)
        loss.backward()
        ref = g[f"{tag}.loss"]
        np.testing.assert_allclose(np.concatenate((loss.detach().numpy(), items.numpy())), ref, rtol=2e-5, atol=1e-6)
        gs = np.array([[t.grad.double().sum().item(), t.grad.double().abs().sum().item()] for t in p])
        np.testing.assert_allclose(gs[:, 1], g[f"{tag}.gradsum"][:, 1], rtol=1e-4)
        np.testing.assert_allclose(p[0].grad.numpy().reshape(-1)[::1009], g[f"{tag}.grad_sample0"], rtol=1e-4, atol=1e-8)


def test_oracle_training_step_matches_reference_fixture():
    """tests/golden/train_step.npz holds the REAL reference's model.train() forward (batch-statistics BN), ComputeLoss,
    backward and BN running-statistic update (generated by tests/golden/make_golden.py train).  The oracle's
    bn_batch_stats forward + loss_ref + torch autograd must reproduce them."""
    import numpy as np
    import torch

    from oracle import loss_ref, model_ref
    from yolov5_b200.cfg import HYP_SCRATCH_LOW, model_cfg

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "train_step.npz"))
    shape, (seed, seed_x, seed_t) = tuple(int(v) for v in g["shape"]), (int(v) for v in g["seed"])
    cfg = model_cfg("yolov5n")
    sd = model_ref.synth_state_dict(cfg, seed=seed)
    x = torch.from_numpy(np.random.RandomState(seed_x).uniform(0, 1, shape).astype(np.float32))
    targets = torch.from_numpy(loss_ref.synth_targets(shape[0], seed=seed_t))
    params = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k and "anchors" not in k) for k, v in sd.items()}
    p = model_ref.forward(cfg, params, x, training=True, bn_batch_stats=True)
    for l, q in enumerate(p):
        assert np.allclose(q.detach().numpy(), g[f"raw{l}"], rtol=1e-4, atol=1e-4)
    loss, items = loss_ref.compute_loss(p, targets, sd["model.24.anchors"], HYP_SCRATCH_LOW)
    assert np.allclose(loss.detach().numpy(), g["loss"], rtol=1e-5) and np.allclose(items.numpy(), g["items"], rtol=1e-5)
    loss.backward()
    for key in g.files:
        if key.startswith("grad."):
            ref = g[key]
            got = params[key[5:]].grad.numpy()
            assert np.abs(got - ref).max() <= 1e-3 * np.abs(ref).max(), key
        elif key.startswith("gnorm."):
            assert abs(float(params[key[6:]].grad.norm()) - float(g[key][0])) <= 1e-3 * float(g[key][0]) + 1e-9, key


def test_post_nms_oracles_match_reference_fixture():
    """oracle/post_ref.py (process_mask, scale_boxes, process_batch: the SURVEY 8(f) rows that come next) against
    tests/golden/post.npz, which holds the real reference's outputs."""
    import numpy as np

    from oracle import post_ref

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "post.npz"))
    hw = tuple(int(v) for v in g["mask.input_hw"])
    for up in (0, 1):
        shape = tuple(int(v) for v in g[f"mask.up{up}.shape"])
        ref = np.unpackbits(g[f"mask.up{up}"])[: int(np.prod(shape))].reshape(shape).astype(np.float32)
        got, val = post_ref.process_mask(g["mask.protos"], g["mask.coef"], g["mask.boxes"], hw, upsample=bool(up))
        off = ref != got
        assert got.shape == shape and (not off.any() or np.abs(val[off] - 0.5).max() < 1e-5), int(off.sum())
    assert np.allclose(post_ref.scale_boxes((640, 640), g["scale.in"], (480, 640)), g["scale.auto"], rtol=0, atol=1e-4)
    assert np.allclose(post_ref.scale_boxes((640, 640), g["scale.in"], (480, 640), ((0.75, 0.75), (10.0, 80.0))), g["scale.given"],
                       rtol=0, atol=1e-4)
    for case in range(4):
        got = post_ref.process_batch(g[f"match{case}.det"], g[f"match{case}.labels"], g["match.iouv"])
        assert np.array_equal(got, g[f"match{case}.correct"]), case


def test_native_mask_oracle_matches_reference_fixture():
    from oracle import post_ref

    g = np.load(os.path.join(G, "post.npz"))
    for tag, hw in (("native", (160, 224)), ("native_pad", (128, 224))):
        shape = tuple(int(v) for v in g[f"mask.{tag}.shape"])
        ref = np.unpackbits(g[f"mask.{tag}"])[: int(np.prod(shape))].reshape(shape).astype(np.float32)
        got, val = post_ref.process_mask_native(g["mask.protos"], g["mask.coef"], g["mask.boxes"], hw)
        off = ref != got
        assert got.shape == shape and (not off.any() or np.abs(val[off] - 0.5).max() < 1e-5), (tag, int(off.sum()))


def test_letterbox_oracle_matches_reference_fixture_and_cv2():
    """oracle/pre_ref.py against tests/golden/pre.npz (the reference's letterbox through cv2 + the dataloader's CHW/RGB
    step), byte for byte; and, when OpenCV is importable, the fixed-point bilinear restatement against cv2 itself."""
    from oracle import pre_ref
    from tests.golden.make_golden_cases import PRE_CASES

    g = np.load(os.path.join(G, "pre.npz"))
    for i, (h, w, seed, kw) in enumerate(PRE_CASES):
        out, ratio, pad = pre_ref.letterbox(pre_ref.synth_image(h, w, seed), **kw)
        assert np.array_equal(pre_ref.to_chw_rgb(out), g[f"lb{i}"]), i
        assert np.allclose([*ratio, *pad], g[f"lb{i}.ratio_pad"], rtol=0, atol=0), i
    try:
        import cv2
    except ImportError:
        return
    rs = np.random.RandomState(1)
    for t in range(40):
        h, w, dh, dw = rs.randint(5, 300), rs.randint(5, 400), rs.randint(4, 300), rs.randint(4, 400)
        img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        assert np.array_equal(cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR), pre_ref.resize_linear_u8(img, (dw, dh))), t


def test_optimizer_oracle_matches_reference_fixture():
    """oracle/optim_ref.py against tests/golden/optim.npz (torch.optim.SGD + clip_grad_norm_ + the reference's ModelEMA)."""
    from oracle import optim_ref

    g = np.load(os.path.join(G, "optim.npz"))
    hyper = json.loads(str(g["hyper"]))
    for case in range(4):
        inv_scale, max_norm, poison, norm = (float(v) for v in g[f"c{case}.cfg"])
        params, grads, moms, emas, groups = optim_ref.synth_problem(40 + case)
        if poison:
            grads[3].flat[5] = np.inf
        p, m, e, eb, gn, skipped = optim_ref.sgd_ema_step(params, grads, moms, emas, groups, hyper, inv_scale, max_norm, 0.9999, 2000.0, 37,
                                                         buffers=[np.linspace(0, 1, 33, dtype=np.float32)],
                                                         ema_buffers=[np.linspace(1, 2, 33, dtype=np.float32)])
        assert skipped == bool(poison)
        if not poison:
            assert abs(gn - norm) <= 1e-5 * norm
        s = optim_ref.FIXTURE_STRIDE
        for i in range(len(params)):
            for tag, arr in (("p", p), ("m", m), ("e", e)):
                assert np.allclose(arr[i].reshape(-1)[::s], g[f"c{case}.{tag}{i}"], rtol=2e-6, atol=1e-7), (case, tag, i)
        assert np.allclose(eb[0], g[f"c{case}.ebuf"], rtol=2e-6, atol=1e-7)


def test_nms_oracle_apriori_labels_fixture():
    g = np.load(os.path.join(G, "nms.npz"))
    meta = [m for m in json.loads(str(g["meta"])) if m.get("labels")]
    assert len(meta) == 2
    for c in meta:
        pred = nms_ref.synth_predictions(c["bs"], c["n"], c["nc"], c["nm"], c["seed"], c["dtype"])
        labels = [g[f"{c['tag']}.labels{b}"] for b in range(c["bs"])]
        out = nms_ref.non_max_suppression(pred, dtype=c["dtype"], labels=labels, **c["kw"])
        for b, o in enumerate(out):
            assert np.array_equal(o, g[f"{c['tag']}.{b}"]), (c["tag"], b)


def test_torch_device_loss_restatement_equals_the_numpy_oracle():
    """oracle.loss_ref.compute_loss_torch (used by bench.py's torch-cuda reference arm) == compute_loss (pinned to the reference)."""
    rs = np.random.RandomState(3)
    p = [torch.from_numpy(rs.normal(0, 1, s).astype(np.float32)).requires_grad_(True) for s in ((4, 3, 16, 16, 85), (4, 3, 8, 8, 85), (4, 3, 4, 4, 85))]
    q = [t.detach().clone().requires_grad_(True) for t in p]
    tg = torch.from_numpy(loss_ref.synth_targets(4, seed=8))
    anc = torch.tensor([[[1.25, 1.625], [2.0, 3.75], [4.125, 2.875]], [[1.875, 3.8125], [3.875, 2.8125], [3.6875, 7.4375]],
                        [[3.625, 2.8125], [4.875, 6.1875], [11.65625, 10.1875]]])
    la, ia = loss_ref.compute_loss(p, tg, anc, HYP_SCRATCH_LOW)
    lb, ib = loss_ref.compute_loss_torch(q, tg, anc, HYP_SCRATCH_LOW)
    assert torch.allclose(la, lb, rtol=1e-5) and torch.allclose(ia, ib, rtol=1e-5)
    la.backward(); lb.backward()
    for a, b in zip(p, q):
        assert torch.allclose(a.grad, b.grad, rtol=1e-4, atol=1e-7)
