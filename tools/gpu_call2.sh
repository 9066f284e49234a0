set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_prepost_gpu.py tests/test_optim_gpu.py -q --maxfail=30 2>&1 | tail -150 > gpurun_out/pytest2.log
tail -25 gpurun_out/pytest2.log
for cfgs in "yolov5m 2 192 256" "yolov5m 2 256 256" "yolov5n 4 128 160" "yolov5s 2 192 256"; do echo "== $cfgs"; timeout 300 python tools/train_diag.py $cfgs 2>&1 | tail -40; done > gpurun_out/train_diag.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err
timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_tma.txt 2>&1
Y5_TMA_STORE=0 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_direct.txt 2>&1
Y5_CG2=1 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_cg2.txt 2>&1
Y5_CG2=2 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_cg2b.txt 2>&1
Y5_CG2=1 Y5_TMA_STORE=0 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_cg2_direct.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_tma.txt 2>&1
Y5_TMA_STORE=0 timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_direct.txt 2>&1
grep -E "fixed ops|full forward" gpurun_out/layers_*.txt
