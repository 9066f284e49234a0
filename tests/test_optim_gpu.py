"""GPU: the step right after backward -- FusedSGD / ModelEMA (y5_opt_step) vs the oracle and the
reference-generated fixture, vs torch.optim.SGD + clip_grad_norm_ + GradScaler over several real training steps, and the
loss kernel's gradient under GradScaler-sized scales (65536 is not an fp16 number)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import loss_ref, model_ref, optim_ref
from yolov5_b200.cfg import HYP_SCRATCH_LOW, model_cfg
from yolov5_b200.models.yolo import DetectionModel
from yolov5_b200.utils.loss import ComputeLoss
from yolov5_b200.utils.torch_utils import FusedSGD, GraphedTrainStep, ModelEMA, smart_optimizer

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


class _Holder(torch.nn.Module):
    def __init__(self, params, buf):
        super().__init__()
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.from_numpy(p.copy())) for p in params])
        self.register_buffer("stat", torch.from_numpy(buf.copy()))


def test_fused_step_vs_oracle_and_reference_fixture(cuda):
    g = np.load(os.path.join(G, "optim.npz"))
    hyper = json.loads(str(g["hyper"]))
    for case in range(4):
        inv_scale, max_norm, poison, _ = (float(v) for v in g[f"c{case}.cfg"])
        params, grads, moms, emas, groups = optim_ref.synth_problem(40 + case)
        if poison:
            grads[3].flat[5] = np.inf
        model = _Holder(params, np.linspace(0, 1, 33, dtype=np.float32)).to(cuda)
        tp = list(model.ps)
        opt = FusedSGD([dict(params=[tp[i] for i in range(len(tp)) if groups[i] == gi], **hyper[gi]) for gi in range(3)], lr=0.01)
        for i, p in enumerate(tp):
            opt.state[p]["momentum_buffer"] = torch.from_numpy(moms[i].copy()).to(cuda)
            p.grad = torch.from_numpy(grads[i].copy()).to(cuda)
        ema = ModelEMA(model, decay=0.9999, tau=2000, updates=37)
        with torch.no_grad():
            for e, v in zip(ema.ema.ps, emas):
                e.copy_(torch.from_numpy(v))
            ema.ema.stat.copy_(torch.from_numpy(np.linspace(1, 2, 33, dtype=np.float32)))
        scaler = None
        if inv_scale != 1.0:  # a real GradScaler whose scale is 1 / inv_scale: fused_step reads it on the device
            scaler = torch.amp.GradScaler("cuda", init_scale=1.0 / inv_scale, growth_interval=1)
            scaler.scale(torch.zeros(1, device=cuda))  # lazy-init the device scale
        opt.fused_step(scaler=scaler, max_norm=max_norm, ema=ema, model=model)
        p_o, m_o, e_o, eb_o, gn, skipped = optim_ref.sgd_ema_step(params, grads, moms, emas, groups, hyper, inv_scale, max_norm, 0.9999, 2000.0, 37,
                                                                 buffers=[np.linspace(0, 1, 33, dtype=np.float32)],
                                                                 ema_buffers=[np.linspace(1, 2, 33, dtype=np.float32)])
        assert opt.last_step_skipped == skipped == bool(poison) and ema.updates == 38
        if not poison:
            assert abs(opt.last_grad_norm - gn) <= 1e-5 * gn
        s = optim_ref.FIXTURE_STRIDE
        for i, p in enumerate(tp):
            for tag, got, ref in (("p", p.detach(), p_o[i]), ("m", opt.state[p]["momentum_buffer"], m_o[i]), ("e", ema.ema.ps[i].detach(), e_o[i])):
                got = got.cpu().numpy()
                assert np.allclose(got, ref, rtol=3e-6, atol=2e-7), (case, tag, i, np.abs(got - ref).max())
                assert np.allclose(got.reshape(-1)[::s], g[f"c{case}.{tag}{i}"], rtol=3e-6, atol=2e-7), (case, tag, i)  # the real reference objects
        assert np.allclose(ema.ema.stat.cpu().numpy(), eb_o[0], rtol=3e-6, atol=2e-7)
        if scaler is not None:  # GradScaler.update semantics: halve on overflow, double after growth_interval clean steps
            want = (1.0 / inv_scale) * (0.5 if poison else 2.0)
            assert abs(float(scaler.get_scale()) - want) <= 1e-6 * want


def test_model_ema_standalone_update_matches_reference_formula(cuda):
    m = DetectionModel("yolov5n").to(cuda)
    ema = ModelEMA(m)
    before = {k: v.clone() for k, v in ema.ema.state_dict().items()}
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn_like(p) * 0.1)
        m.model[0].bn.running_mean.add_(0.5)
    for step in range(1, 4):
        ema.update(m)
        d = 0.9999 * (1 - np.exp(-step / 2000))
        msd = m.state_dict()
        for k, v in ema.ema.state_dict().items():
            if v.dtype.is_floating_point:
                before[k] = before[k] * d + (1 - d) * msd[k]
                assert torch.allclose(v, before[k], rtol=1e-5, atol=1e-7), (step, k)
            else:
                assert torch.equal(v, before[k])
    assert ema.updates == 3


def _train_setup(dev, seed=0):
    cfg = model_cfg("yolov5n")
    m = DetectionModel("yolov5n")
    m.load_state_dict(model_ref.synth_state_dict(cfg, seed=seed))
    m = m.to(dev).train()
    m.hyp = dict(HYP_SCRATCH_LOW)
    imgs = [torch.from_numpy(np.random.RandomState(10 + i).randint(0, 256, (2, 3, 128, 128)).astype(np.uint8)).to(dev) for i in range(3)]
    tgts = [torch.from_numpy(loss_ref.synth_targets(2, seed=20 + i)).float().to(dev) for i in range(3)]
    return m, imgs, tgts


def test_fused_step_tracks_torch_sgd_clip_gradscaler_over_real_steps(cuda):
    """Three real training steps of yolov5n (fp16 autocast, GradScaler init 65536) drive BOTH optimizers with the same scaled
    gradients: the reference sequence scaler.unscale_ / clip_grad_norm_ / scaler.step / scaler.update / zero_grad / ema.update
    (train.py:413-421) with torch.optim.SGD in the reference's 3-group layout on a twin model, vs smart_optimizer(...).fused_step.
    (Two independent training runs cannot be compared weight by weight: the weight-gradient kernel's reduction order and the
    loss landscape of a random network amplify rounding differences within a few steps.)"""
    ma, imgs, tgts = _train_setup(cuda)
    mb, _, _ = _train_setup(cuda)
    la = ComputeLoss(ma)
    oa = smart_optimizer(ma, "SGD", lr=0.01, momentum=0.937, decay=5e-4)
    groups = [[], [], []]
    for v in mb.modules():
        for n, p in v.named_parameters(recurse=False):
            groups[2 if n == "bias" else 1 if isinstance(v, torch.nn.BatchNorm2d) and n == "weight" else 0].append(p)
    ob = torch.optim.SGD(groups[2], lr=0.01, momentum=0.937, nesterov=True)
    ob.add_param_group({"params": groups[0], "weight_decay": 5e-4})
    ob.add_param_group({"params": groups[1], "weight_decay": 0.0})
    sa, sb = torch.amp.GradScaler("cuda"), torch.amp.GradScaler("cuda")
    sb.scale(torch.zeros(1, device=cuda))  # lazy-init the twin's device scale
    ea, eb = ModelEMA(ma), ModelEMA(mb)
    pa, pb = list(ma.parameters()), list(mb.parameters())
    for i in range(3):
        with torch.autocast("cuda", dtype=torch.float16):
            pred = ma(imgs[i])
        loss_a, _ = la(pred, tgts[i])
        sa.scale(loss_a).backward()
        for qa, qb in zip(pa, pb):
            qb.grad = qa.grad.clone()
        if i == 1:  # an overflow step: both must skip the update and halve the scale
            pa[5].grad.view(-1)[0] = float("inf")
            pb[5].grad.view(-1)[0] = float("inf")
        oa.fused_step(scaler=sa, max_norm=10.0, ema=ea, model=ma)
        oa.zero_grad()
        sb.unscale_(ob)
        torch.nn.utils.clip_grad_norm_(pb, max_norm=10.0)
        sb.step(ob)
        sb.update()
        ob.zero_grad()
        with torch.no_grad():  # the twin does no forward: give its BN buffers the engine's, so the two EMAs see the same state_dict
            for (ka, ba), (kb, bb) in zip(ma.named_buffers(), mb.named_buffers()):
                bb.copy_(ba)
        eb.update(mb)
        assert float(sa.get_scale()) == float(sb.get_scale()), i
        assert oa.last_step_skipped == (i == 1)
    assert float(sa.get_scale()) == 32768.0
    for (k, a), b in zip(ma.named_parameters(), pb):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), (k, float((a - b).abs().max()))
    for (k, a), b in zip(ea.ema.state_dict().items(), eb.ema.state_dict().values()):
        if a.dtype.is_floating_point:
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), k


@pytest.mark.parametrize("scale", [65536.0, 65536.0 * 8])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_loss_gradient_under_gradscaler_scale(cuda, scale, dtype):
    """the upstream gradient of `scaler.scale(loss) * WORLD_SIZE` must be applied in fp32 inside the loss
    kernels: fp16 predictions, scale 65536 (x8 ranks).  Gradients must be finite and equal scale x the fp32-logit gradients
    of the oracle (to fp16/bf16 rounding), including the tiny objectness gradients of confident negatives."""
    rs = np.random.RandomState(5)
    shapes = [(4, 3, 16, 16, 85), (4, 3, 8, 8, 85), (4, 3, 4, 4, 85)]
    p32 = [torch.from_numpy(rs.normal(0, 1.5, s).astype(np.float32)) for s in shapes]
    for t in p32:
        t[..., 4] -= 6.0  # confident negatives: d(BCE)/dx = sigmoid(x) ~ 1e-3 .. 1e-4 before the 1/cells factor
    targets = torch.from_numpy(loss_ref.synth_targets(4, seed=6)).float()
    m = DetectionModel("yolov5n").to(cuda)
    m.hyp = dict(HYP_SCRATCH_LOW)
    crit = ComputeLoss(m)
    p = [t.to(cuda, dtype).requires_grad_(True) for t in p32]
    loss, _ = crit(p, targets.to(cuda))
    (loss * scale).backward()
    pr = [t.to(dtype).float().requires_grad_(True) for t in p32]
    loss_r, _ = loss_ref.compute_loss(pr, targets, m.model[-1].anchors.cpu(), HYP_SCRATCH_LOW)
    loss_r.backward()
    eps = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    for a, b in zip(p, pr):
        ga, gb = a.grad.float().cpu(), b.grad * scale
        assert bool(torch.isfinite(ga).all())
        # whole tensor: box / class gradients of a cell are sums of several contributions accumulated in the prediction dtype (like
        # autograd's index_put backward), so the bound is relative to the largest gradient, as in tests/test_loss_gpu.py
        assert float((ga - gb).abs().max()) <= 4 * eps * float(gb.abs().max())
        # objectness column: one writer per cell -> element-wise one rounding, INCLUDING the tiny gradients of confident negatives
        # (2e-4 x sigmoid(-6): flushed to zero / subnormal if the scale were applied after rounding to fp16)
        oa, ob = ga[..., 4].flatten(), gb[..., 4].flatten()
        rel = (oa - ob).abs() / ob.abs().clamp_min(1e-30)
        # (matched cells can have sigmoid(x) ~ tobj, i.e. a gradient that is itself a cancellation: they are the < 0.1 % tail)
        assert float(torch.quantile(rel, 0.999)) < 2 * eps, float(torch.quantile(rel, 0.999))
        assert float((oa != 0).float().mean()) > 0.999


def test_graphed_train_step_with_loss_scaling_ema_and_schedule(cuda):
    """GraphedTrainStep = the eager fused loop (same kernels), incl. dynamic loss scale on the device, EMA, and a learning
    rate changed between replays (read from param_groups at every call)."""
    ma, imgs, tgts = _train_setup(cuda, seed=1)
    mb, _, _ = _train_setup(cuda, seed=1)
    oa = smart_optimizer(ma, "SGD", lr=0.01, momentum=0.937, decay=5e-4)
    ob = smart_optimizer(mb, "SGD", lr=0.01, momentum=0.937, decay=5e-4)
    ea, eb = ModelEMA(ma), ModelEMA(mb)
    step = GraphedTrainStep(ma, ComputeLoss(ma), oa, batch=2, size=128, ema=ea)
    lb, sb = ComputeLoss(mb), torch.amp.GradScaler("cuda")
    items_a = []
    w0 = torch.cat([v.detach().flatten() for v in ma.parameters()]).clone()
    for i in range(3):
        lr = 0.01 * (1 + i)
        for g_ in oa.param_groups + ob.param_groups:
            g_["lr"] = lr
        items_a.append(step(imgs[i], tgts[i]).clone())
        torch.cuda.synchronize()
        with torch.autocast("cuda", dtype=torch.float16):
            pb = mb(imgs[i])
        loss_b, items_b = lb(pb, tgts[i])
        sb.scale(loss_b).backward()
        ob.fused_step(scaler=sb, max_norm=10.0, ema=eb, model=mb)
        ob.zero_grad()
        assert torch.allclose(items_a[-1], items_b, rtol=3e-2, atol=1e-4), (i, items_a[-1], items_b)
        if i == 0:
            # after ONE step from identical weights the two runs differ only by reduction order (weight-gradient red.add, fp16
            # atomics of the loss gradient): a few percent of the step at most.  Later steps of this random little problem are
            # chaotic (tools/train_diag.py), so only their loss items / scale / counters are compared.
            wa = torch.cat([v.detach().flatten() for v in ma.parameters()])
            wb = torch.cat([v.detach().flatten() for v in mb.parameters()])
            moved = float((wb - w0).norm())
            assert moved > 0 and float((wa - wb).norm()) <= 0.05 * moved, (float((wa - wb).norm()), moved)
            ema_a = torch.cat([v.flatten() for v in ea.ema.parameters()])
            ema_b = torch.cat([v.flatten() for v in eb.ema.parameters()])
            assert float((ema_a - w0).norm()) > 0 and float((ema_a - ema_b).norm()) <= 0.05 * float((ema_b - w0).norm())
    assert ea.updates == eb.updates == 3 and float(step.scaler.get_scale()) == float(sb.get_scale())
    assert int(ma.model[0].bn.num_batches_tracked) == 3 and all(bool(torch.isfinite(v).all()) for v in ma.parameters())


def test_data_parallel_mode_world1_packs_gradients_and_steps_identically(cuda):
    """FusedSGD.data_parallel (the device half of the gradient exchange, y5_grad_pack): in a 1-rank group the arena must hold
    every gradient bit-exactly at its 16-byte aligned offset, and the step taken from the arena must equal the plain step."""
    import torch.distributed as dist

    own_group = not dist.is_initialized()
    if own_group:
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29683", rank=0, world_size=1, device_id=cuda)
    try:
        torch.manual_seed(5)
        shapes = [(33, 7, 3, 3), (33,), (1,), (64, 33, 1, 1), (5,), (70001,)]
        ma = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(s, device=cuda)) for s in shapes])
        mb = torch.nn.ParameterList([torch.nn.Parameter(p.detach().clone()) for p in ma])
        oa = FusedSGD(list(ma), lr=0.05, momentum=0.9, weight_decay=1e-3, nesterov=True)
        ob = FusedSGD(list(mb), lr=0.05, momentum=0.9, weight_decay=1e-3, nesterov=True)
        oa.data_parallel(ma)
        for it in range(3):
            for i, (pa, pb) in enumerate(zip(ma, mb)):
                g = torch.randn_like(pa) * (10.0 if it == 1 else 1.0)
                if i == 4 and it == 2:
                    pa.grad = pb.grad = None  # a parameter without a gradient this step: zeros in the arena, no update
                    continue
                pa.grad, pb.grad = g, g.clone()
            grads = [None if p.grad is None else p.grad.clone() for p in ma]
            oa.fused_step(max_norm=10.0)
            ob.fused_step(max_norm=10.0)
            off = 0
            for p, g in zip(ma, grads):
                n = p.numel()
                want = torch.zeros(n, device=cuda) if g is None else g.flatten()
                assert torch.equal(oa._arena[off : off + n], want)
                off += (n + 3) // 4 * 4
            for pa, pb in zip(ma, mb):
                assert torch.equal(pa, pb)
            assert oa.last_grad_norm == ob.last_grad_norm
            oa.zero_grad()
            ob.zero_grad()
    finally:
        if own_group:
            dist.destroy_process_group()
