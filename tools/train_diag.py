"""Per-tensor gradient report of one training step (engine vs fp32 oracle vs torch-AMP) for a model / shape given on the command
line: python tools/train_diag.py yolov5m 2 192 256 fp16"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from oracle import loss_ref, model_ref
from tests.test_train_gpu import _ref_train_step
from yolov5_b200.cfg import HYP_SCRATCH_LOW, model_cfg
from yolov5_b200.models.yolo import DetectionModel
from yolov5_b200.utils.loss import ComputeLoss

name, b, h, w = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
dtype = {"fp16": torch.float16, "bf16": torch.bfloat16}[sys.argv[5] if len(sys.argv) > 5 else "fp16"]
cuda = torch.device("cuda:0")
cfg = model_cfg(name)
sd = model_ref.synth_state_dict(cfg, seed=21)
g = torch.Generator().manual_seed(22)
img = (torch.rand(b, 3, h, w, generator=g) * 255).to(torch.uint8)
targets = torch.from_numpy(loss_ref.synth_targets(b, seed=23)).float()
p32, loss32, g32 = _ref_train_step(cfg, sd, img, targets, cuda, None)
pamp, lossamp, gamp = _ref_train_step(cfg, sd, img, targets, cuda, dtype)
m = DetectionModel(name)
m.load_state_dict(sd)
m = m.to(cuda).train()
m.hyp = dict(HYP_SCRATCH_LOW)
with torch.autocast("cuda", dtype=dtype):
    p = m(img.to(cuda))
for l, (a, r, lo) in enumerate(zip(p, p32, pamp)):
    sc = float(r.abs().max())
    print(f"raw{l}: mine {float((a.detach().float() - r).abs().max()) / sc:.3e}  amp {float((lo.float() - r).abs().max()) / sc:.3e}")
loss, items = ComputeLoss(m)(p, targets.to(cuda))
print("loss", float(loss), float(loss32), float(lossamp))
# gradient of the loss w.r.t. the head maps: engine loss kernel on ITS maps vs oracle on the fp32 maps
pg = torch.autograd.grad(loss, p, retain_graph=True)
p32r = [q.clone().requires_grad_(True) for q in p32]
l2, _ = loss_ref.compute_loss([q.float().cpu() for q in p32r], targets, sd["model.24.anchors"], HYP_SCRATCH_LOW)
g2 = torch.autograd.grad(l2, p32r)
pe = [q.detach().float().cpu().requires_grad_(True) for q in p]  # oracle loss evaluated AT THE ENGINE'S OWN MAPS: isolates the loss kernel
l3, _ = loss_ref.compute_loss(pe, targets, sd["model.24.anchors"], HYP_SCRATCH_LOW)
g3 = torch.autograd.grad(l3, pe)
pa = [q.detach().float().cpu().requires_grad_(True) for q in pamp]
l4, _ = loss_ref.compute_loss(pa, targets, sd["model.24.anchors"], HYP_SCRATCH_LOW)
g4 = torch.autograd.grad(l4, pa)
for l in range(3):
    a, r = pg[l].float().cpu(), g2[l].float().cpu()
    print(f"dL/draw{l}: kernel@engine-maps vs oracle@fp32-maps rel L2 {float((a - r).norm() / r.norm()):.3e} | kernel vs oracle@engine-maps "
          f"{float((a - g3[l]).norm() / g3[l].norm()):.3e} | oracle@amp-maps vs oracle@fp32-maps {float((g4[l] - r).norm() / r.norm()):.3e} | "
          f"mean signed raw err mine {float((p[l].detach().float().cpu() - p32[l].cpu()).mean()):.2e} amp {float((pamp[l].float().cpu() - p32[l].cpu()).mean()):.2e} "
          f"| obj-col mean err mine {float((p[l].detach().float().cpu() - p32[l].cpu())[..., 4].mean()):.2e} amp {float((pamp[l].float().cpu() - p32[l].cpu())[..., 4].mean()):.2e}")
loss.backward()
named = dict(m.named_parameters())
rows = []
for k, gr in g32.items():
    n = float(gr.norm())
    if n == 0:
        continue
    e, el = float((named[k].grad.float() - gr).norm()) / n, float((gamp[k].float() - gr).norm()) / n
    rows.append((e / (1e-3 + el), k, e, el, tuple(gr.shape)))
rows.sort(reverse=True)
for r in rows[:25]:
    print(f"{r[0]:8.2f} {r[1]:40s} mine {r[2]:.3e} amp {r[3]:.3e} {r[4]}")
print("median ratio", sorted(x[0] for x in rows)[len(rows) // 2])
