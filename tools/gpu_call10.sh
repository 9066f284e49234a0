set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q -x -k "wide_patch" > gpurun_out/pytest10_wide.log 2>&1
tail -6 gpurun_out/pytest10_wide.log
Y5_PATCH_WIDE_BO=0 timeout 600 python -m pytest tests/test_conv_gpu.py -q -x -k "wide_patch" > gpurun_out/pytest10_wide_nobo.log 2>&1
tail -4 gpurun_out/pytest10_wide_nobo.log
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -q --maxfail=10 > gpurun_out/pytest10_full.log 2>&1
tail -6 gpurun_out/pytest10_full.log
timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v10.txt 2>&1
Y5_PATCH_WIDE=0 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v10_narrow.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_v10.txt 2>&1
grep -E "fixed ops|full forward" gpurun_out/layers_l_v10*.txt gpurun_out/layers_s_v10.txt
