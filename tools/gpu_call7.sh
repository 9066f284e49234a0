set +e
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --deselect tests/test_train_ddp_gpu.py > gpurun_out/pytest7_full.log 2>&1
tail -6 gpurun_out/pytest7_full.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v7.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_v7.txt 2>&1
grep -E "fixed ops|full forward|detect" gpurun_out/layers_l_v7.txt gpurun_out/layers_s_v7.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err
