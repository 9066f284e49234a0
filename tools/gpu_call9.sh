set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_train_ddp_gpu.py > gpurun_out/pytest9_full.log 2>&1
tail -5 gpurun_out/pytest9_full.log
timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v9.txt 2>&1
Y5_B_RESIDENT=1 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v9_bres.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_v9.txt 2>&1
Y5_B_RESIDENT=1 timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_v9_bres.txt 2>&1
grep -E "fixed ops|full forward" gpurun_out/layers_l_v9*.txt gpurun_out/layers_s_v9*.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default9.json 2> gpurun_out/bench_default9.err
tail -3 gpurun_out/bench_default9.err
head -c 400 gpurun_out/bench_default9.json
