set +e
mkdir -p gpurun_out
timeout 300 python tools/train_bench.py --model yolov5m --batch 16 --steps 10 --skip-reference --profile --graph > gpurun_out/train_m_v26.txt 2>&1
tail -60 gpurun_out/train_m_v26.txt | cut -c1-200
