set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_optim_gpu.py tests/test_train_gpu.py tests/test_conv_gpu.py -q --maxfail=30 > gpurun_out/pytest4_full.log 2>&1
tail -12 gpurun_out/pytest4_full.log
for op in detect.0 model.0 model.2.m0.cv1 model.4.m0.cv1 model.6.m0.cv2; do
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm -c 1 -o gpurun_out/ncu2_$op python tools/op_probe.py yolov5l 64 640 bf16 $op 2 > gpurun_out/ncu2_$op.log 2>&1
  tail -1 gpurun_out/ncu2_$op.log
done
timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_rule2.txt 2>&1
grep -E "fixed ops|full forward" gpurun_out/layers_l_rule2.txt
ls -la gpurun_out/*.ncu-rep
