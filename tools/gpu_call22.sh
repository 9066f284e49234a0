set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_train_gpu.py -q --maxfail=10 > gpurun_out/pytest22.log 2>&1
tail -4 gpurun_out/pytest22.log
for r in 0 1 2; do
Y5_BK_RULE=$r timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_v22_r$r.txt 2>&1
Y5_BK_RULE=$r timeout 200 python tools/layer_profile.py yolov5m 32 640 fp16 > gpurun_out/layers_m_v22_r$r.txt 2>&1
Y5_BK_RULE=$r timeout 300 python tools/train_bench.py --model yolov5m --batch 16 --steps 10 --skip-reference > gpurun_out/train_m_v22_r$r.txt 2>&1
done
grep -E "fixed ops" gpurun_out/layers_s_v22*.txt gpurun_out/layers_m_v22*.txt
tail -3 gpurun_out/train_m_v22_r*.txt
