set +e
mkdir -p gpurun_out
T='tests/test_train_gpu.py::test_model_training_step_vs_oracle_amp_yardstick'
for env in "Y5_TRAIN_PACK_PLAN=0" "Y5_TRAIN_PACK_PLAN=0 Y5_BN_RED_MIN_ROWS=64" "Y5_TRAIN_PACK_PLAN=0 Y5_BN_RED_MIN_ROWS=2048" "Y5_TRAIN_PACK_PLAN=1"; do
  echo "== $env"; env $env timeout 300 python -m pytest "$T" -q -s 2>&1 | grep -E "gradient report|passed|failed|Error|assert" | cut -c1-400
done > gpurun_out/yardstick_v28.txt 2>&1
cat gpurun_out/yardstick_v28.txt
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_optim_gpu.py -q 2>&1 | tail -8 > gpurun_out/pytest_train2.log; cat gpurun_out/pytest_train2.log
timeout 300 python tools/train_bench.py --model yolov5m --batch 16 --steps 10 --skip-reference --profile --graph > gpurun_out/train_m_v28.txt 2>&1
grep -E "^engine|host-side|bn_act|col_stats|conv_wgrad|weight_pack|Self CUDA time|Memset|aten::" gpurun_out/train_m_v28.txt | cut -c1-60,130-215
