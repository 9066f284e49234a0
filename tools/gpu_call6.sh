set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -q --maxfail=30 > gpurun_out/pytest6_full.log 2>&1
tail -5 gpurun_out/pytest6_full.log
timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_hb.txt 2>&1
Y5_CG2_MT2=1 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_mt2.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_hb.txt 2>&1
Y5_CG2_MT2=1 timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_mt2.txt 2>&1
grep -E "fixed ops|full forward" gpurun_out/layers_l_hb.txt gpurun_out/layers_l_mt2.txt gpurun_out/layers_s_hb.txt gpurun_out/layers_s_mt2.txt
