"""Exactly ONE training step (forward, loss, scaled backward, fused un-scale/clip/SGD/EMA) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off`:     python tools/one_train_step.py yolov5m 16 640 fp16
The step is the eager one of bench.py's train_ddp record (same public calls); warm-up steps run before the profiled region."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from oracle import loss_ref, model_ref  # synthetic weights / labels only
from yolov5_b200.cfg import HYP_SCRATCH_LOW, model_cfg
from yolov5_b200.models.yolo import DetectionModel
from yolov5_b200.utils.loss import ComputeLoss
from yolov5_b200.utils.torch_utils import ModelEMA, smart_optimizer

name, bs, size = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dt = {"fp16": torch.float16, "bf16": torch.bfloat16}[sys.argv[4]]
dev = torch.device("cuda:0")
cfg = model_cfg(name)
m = DetectionModel(name)
m.load_state_dict(model_ref.synth_state_dict(cfg, seed=0))
m = m.to(dev).train()
m.hyp = dict(HYP_SCRATCH_LOW)
loss_fn = ComputeLoss(m)
opt = smart_optimizer(m, "SGD", lr=1e-3, momentum=0.937, decay=5e-4)
scaler = torch.amp.GradScaler("cuda", enabled=dt == torch.float16)
ema = ModelEMA(m)
img = torch.randint(0, 256, (bs, 3, size, size), dtype=torch.uint8, device=dev)
tgt = torch.from_numpy(loss_ref.synth_targets(bs, seed=1)).float().to(dev)


def step():
    with torch.autocast("cuda", dtype=dt):
        p = m(img)
    loss, items = loss_fn(p, tgt)
    scaler.scale(loss).backward()
    opt.fused_step(scaler=scaler, max_norm=10.0, ema=ema, model=m)
    opt.zero_grad()
    return items


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
items = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(name, bs, size, dt, "loss items", [round(float(v), 4) for v in items])
