set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_train_ddp_gpu.py > gpurun_out/pytest23_full.log 2>&1
tail -5 gpurun_out/pytest23_full.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v23.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_v23.txt 2>&1
grep -E "fixed ops" gpurun_out/layers_l_v23.txt gpurun_out/layers_s_v23.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default23.json 2> gpurun_out/bench_default23.err
tail -2 gpurun_out/bench_default23.err
head -c 300 gpurun_out/bench_default23.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref23.json 2> gpurun_out/bench_ref23.err
head -c 300 gpurun_out/bench_ref23.json
