set +e
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_train_gpu.py -x -q 2>&1 | tail -6 > gpurun_out/pytest_train.log; cat gpurun_out/pytest_train.log
PROBE_BN_ONLY=1 timeout 100 python tools/train_kernel_probe.py > gpurun_out/probe_s_v27.txt 2>&1; cat gpurun_out/probe_s_v27.txt
PROBE_BN_ONLY=1 PROBE_SHAPES=m timeout 100 python tools/train_kernel_probe.py > gpurun_out/probe_m_v27.txt 2>&1; cat gpurun_out/probe_m_v27.txt
timeout 300 python tools/train_bench.py --model yolov5m --batch 16 --steps 10 --skip-reference --profile --graph > gpurun_out/train_m_v27.txt 2>&1
grep -E "^engine|host-side|bn_act|col_stats|conv_wgrad|weight_pack|Self CUDA time" gpurun_out/train_m_v27.txt | cut -c1-60,130-215
