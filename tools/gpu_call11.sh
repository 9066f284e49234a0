set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -q --maxfail=10 > gpurun_out/pytest11.log 2>&1
tail -6 gpurun_out/pytest11.log
timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v11.txt 2>&1
Y5_PATCH_WIDE=0 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v11_narrow.txt 2>&1
Y5_PATCH_WIDE=0 Y5_B_RESIDENT=0 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v11_narrow_nores.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_v11.txt 2>&1
Y5_PATCH_WIDE=0 Y5_B_RESIDENT=0 timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_v11_narrow_nores.txt 2>&1
grep -E "fixed ops|full forward" gpurun_out/layers_l_v11*.txt gpurun_out/layers_s_v11*.txt
