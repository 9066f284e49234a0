set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -q --maxfail=10 > gpurun_out/pytest17.log 2>&1
tail -4 gpurun_out/pytest17.log
timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v17.txt 2>&1
Y5_CG2_MT2=0 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v17_nomt2.txt 2>&1
Y5_CG2_MT2=1 timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v17_mt2all.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_v17.txt 2>&1
grep -E "fixed ops" gpurun_out/layers_l_v17*.txt gpurun_out/layers_s_v17*.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default17.json 2> gpurun_out/bench_default17.err
tail -2 gpurun_out/bench_default17.err
head -c 300 gpurun_out/bench_default17.json
