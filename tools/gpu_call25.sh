set +e
mkdir -p gpurun_out
timeout 400 python tools/train_bench.py --model yolov5m --batch 16 --steps 10 --skip-reference --profile --graph > gpurun_out/train_m_v25.txt 2>&1
tail -45 gpurun_out/train_m_v25.txt | cut -c1-160
