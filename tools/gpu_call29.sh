set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_optim_gpu.py -q -s 2>&1 | grep -E "loss report|passed|failed|Error|assert " | cut -c1-600 > gpurun_out/pytest_train3.log; cat gpurun_out/pytest_train3.log
for u in 4 2; do
  echo "== Y5_BN_RED_U=$u"
  Y5_BN_RED_U=$u PROBE_BN_ONLY=1 PROBE_SHAPES=m timeout 100 python tools/train_kernel_probe.py 2>&1
  Y5_BN_RED_U=$u timeout 300 python tools/train_bench.py --model yolov5m --batch 16 --steps 10 --skip-reference --profile --graph 2>&1 | grep -E "^engine|host-side|bn_act|col_stats|conv_wgrad|Self CUDA time" | cut -c1-60,130-215
done > gpurun_out/bn_v29.txt 2>&1
cat gpurun_out/bn_v29.txt
