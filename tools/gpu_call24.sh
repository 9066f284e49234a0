set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -q --maxfail=10 > gpurun_out/pytest24.log 2>&1
tail -4 gpurun_out/pytest24.log
for i in 1 2; do
Y5B200_LIB=$PWD/yolov5_b200/build/liby5b200_prev.so timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v24_prev$i.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v24_new$i.txt 2>&1
done
Y5B200_LIB=$PWD/yolov5_b200/build/liby5b200_prev.so timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_v24_prev.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_v24_new.txt 2>&1
grep -E "fixed ops|detect" gpurun_out/layers_l_v24*.txt gpurun_out/layers_s_v24*.txt
