# round 2, GPU call 1: correctness of everything new + first bench line of the new default + A/B of the cheap conv_gemm knobs
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 --deselect tests/test_train_ddp_gpu.py 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 > gpurun_out/smoke.log; cat gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err
timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_new.txt 2>&1
Y5_STAGE_CAP=1 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_cap.txt 2>&1
Y5_B_RESIDENT=0 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_nores.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_new.txt 2>&1
Y5_STAGE_CAP=1 Y5_B_RESIDENT=0 timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_old.txt 2>&1
grep -E "fixed ops|full forward" gpurun_out/layers_*.txt
ls -la gpurun_out | head -30
