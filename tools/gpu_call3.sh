set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_optim_gpu.py tests/test_prepost_gpu.py tests/test_model_gpu.py -q --maxfail=30 2>&1 | tail -40 > gpurun_out/pytest3.log
tail -12 gpurun_out/pytest3.log
for cfgs in "yolov5m 2 192 256" "yolov5s 2 192 256"; do echo "== $cfgs"; timeout 300 python tools/train_diag.py $cfgs 2>&1 | grep -E "dL/draw|median|raw" ; done > gpurun_out/train_diag2.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_rule.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_rule.txt 2>&1
grep -E "fixed ops|full forward" gpurun_out/layers_*rule.txt
# ncu full captures of the three worst layers of yolov5l: Detect P3 head, stem, model.1 (+ a 1x1 layer for reference)
for op in detect.0 model.0 model.1 model.2.cv1 model.4.m0.cv1; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 1 -c 1 -o gpurun_out/ncu_$op python tools/op_probe.py yolov5l 64 640 bf16 $op 2 > gpurun_out/ncu_$op.log 2>&1
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err
ls -la gpurun_out/*.ncu-rep
