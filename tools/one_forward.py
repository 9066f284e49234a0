"""Exactly ONE forward + NMS of a model between cudaProfilerStart/Stop (for `ncu --profile-from-start off`):
    python tools/one_forward.py yolov5l 64 640 bf16
Warm-up (program build, weight packing, graph capture) happens before the profiled region; the profiled forward replays
the same launches the bench's timed step does."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from yolov5_b200.models.yolo import DetectionModel
from yolov5_b200.utils.general import non_max_suppression

name, bs, size = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dt = {"fp16": torch.float16, "bf16": torch.bfloat16}[sys.argv[4]]
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = DetectionModel(name).to(dt).to(dev).eval()
x = torch.rand(bs, 3, size, size, device=dev).to(dt)
for _ in range(3):
    non_max_suppression(m(x)[0], 0.25, 0.45, max_det=300)
torch.cuda.synchronize()
torch.cuda.profiler.start()
det = non_max_suppression(m(x)[0], 0.25, 0.45, max_det=300)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(name, bs, size, dt, "detections", [int(d.shape[0]) for d in det][:4])
