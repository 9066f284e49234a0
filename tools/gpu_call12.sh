set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -q --maxfail=10 > gpurun_out/pytest13.log 2>&1
tail -6 gpurun_out/pytest13.log
timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v13.txt 2>&1
Y5_B_GROUP=2 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v13_g2.txt 2>&1
Y5_B_GROUP=0 Y5_CG2_N64=0 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v13_base.txt 2>&1
Y5_B_GROUP=0 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v13_n64.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_v13.txt 2>&1
Y5_B_GROUP=0 Y5_CG2_N64=0 timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_v13_base.txt 2>&1
grep -E "fixed ops|full forward" gpurun_out/layers_l_v13*.txt gpurun_out/layers_s_v13*.txt
