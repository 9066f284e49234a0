set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -q --maxfail=10 > gpurun_out/pytest15.log 2>&1
tail -6 gpurun_out/pytest15.log
Y5B200_LIB=$PWD/yolov5_b200/build/liby5b200_v9.so timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v15_old.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v15_new.txt 2>&1
Y5_B_GROUP=0 Y5_CG2_N64=0 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v15_base.txt 2>&1
Y5_B_GROUP=2 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v15_g2.txt 2>&1
Y5_CG2_N64=0 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v15_non64.txt 2>&1
Y5B200_LIB=$PWD/yolov5_b200/build/liby5b200_v9.so timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_v15_old.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_v15_new.txt 2>&1
grep -E "fixed ops" gpurun_out/layers_l_v15*.txt gpurun_out/layers_s_v15*.txt
