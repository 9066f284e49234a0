"""Run ONE op of a model's program in isolation (for ncu captures and micro-timing):
    python tools/op_probe.py yolov5l 64 640 bf16 detect.0 [reps]
op names are those of tools/layer_profile.py (e.g. model.4.m0.cv1, model.0, detect.0)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from yolov5_b200 import _lib
from yolov5_b200.models.yolo import DetectionModel

name, bs, size = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dt = {"fp16": torch.float16, "bf16": torch.bfloat16}[sys.argv[4]]
want = sys.argv[5]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 5
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = DetectionModel(name).to(dt).to(dev).eval()
x = torch.rand(bs, 3, size, size, device=dev).to(dt)
prog = m._program(x)
m(x)
torch.cuda.synchronize()
st = _lib.stream_ptr(dev)
if want.startswith("detect."):
    i = int(want.split(".")[1])
    no = prog.det_shapes[0][-1]
    zbuf = torch.empty(prog.B, prog.z_rows, no, dtype=prog.dtype, device=dev)
    raw = torch.empty(prog.det_shapes[i], dtype=prog.dtype, device=dev)
    plan = prog.head_ops[i]

    def run():
        _lib.check(prog.lib.y5_detect_plan_run_to(plan, raw.data_ptr(), zbuf.data_ptr(), C.c_void_p(st)), "detect")
else:
    ops = [op for op in prog.ops if op.name.startswith(want)]
    assert ops, [op.name for op in prog.ops]
    op = ops[0]

    def run():
        op.run(st)

e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
run()
torch.cuda.synchronize()
torch.cuda.profiler.start()  # ncu --profile-from-start off: only the probed launches are captured
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"{name} {want}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us per launch (back to back, {reps} reps)")
