set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv
for i in 1 2; do
Y5B200_LIB=$PWD/yolov5_b200/build/liby5b200_v9.so timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v14_old$i.txt 2>&1
Y5_B_GROUP=0 Y5_CG2_N64=0 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v14_base$i.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v14_new$i.txt 2>&1
done
Y5_B_GROUP=2 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v14_g2.txt 2>&1
Y5B200_LIB=$PWD/yolov5_b200/build/liby5b200_v9.so timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_v14_old.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_v14_new.txt 2>&1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv
grep -E "fixed ops" gpurun_out/layers_l_v14*.txt gpurun_out/layers_s_v14*.txt
