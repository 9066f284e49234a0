set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_optim_gpu.py tests/test_train_gpu.py -q --maxfail=30 > gpurun_out/pytest5_full.log 2>&1
tail -8 gpurun_out/pytest5_full.log
timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_lean.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_lean.txt 2>&1
grep -E "fixed ops|full forward" gpurun_out/layers_l_lean.txt gpurun_out/layers_s_lean.txt
for op in model.0 model.4.m0.cv1; do
  timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --profile-from-start off -k regex:conv_gemm -c 1 --csv --log-file gpurun_out/ncu3_$op.csv python tools/op_probe.py yolov5l 64 640 bf16 $op 2 > gpurun_out/ncu3_$op.log 2>&1
  tail -4 gpurun_out/ncu3_$op.csv | cut -c1-200
done
