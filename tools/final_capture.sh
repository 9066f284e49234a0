set +e
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err
timeout 300 python bench.py --workload yolov5l --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err
timeout 300 python bench.py --workload yolov5s-train --steps 10 --warmup 3 > gpurun_out/bench_s_train.json 2> gpurun_out/bench_s_train.err
timeout 300 python bench.py --workload yolov5m-train --steps 10 --warmup 3 > gpurun_out/bench_m_train.json 2> gpurun_out/bench_m_train.err
timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l.txt 2>&1
timeout 200 python tools/train_kernel_probe.py > gpurun_out/train_probe.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_s.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
timeout 400 ncu --metrics $M --clock-control none -k regex:conv_gemm -s 104 -c 52 --csv --log-file gpurun_out/conv_metrics_s.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_m.log 2>&1
timeout 400 ncu --metrics $M --clock-control none -k regex:conv_gemm -s 192 -c 96 --csv --log-file gpurun_out/conv_metrics_l.csv python bench.py --workload yolov5l --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ml.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3300 -c 1200 --csv --log-file gpurun_out/launches_train.csv python tools/train_bench.py --model yolov5s --batch 16 --steps 1 --skip-reference > gpurun_out/ncu_t.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 70 -c 3 -o gpurun_out/wgrad_full python tools/train_bench.py --model yolov5s --batch 16 --steps 1 --skip-reference > gpurun_out/ncu_w.log 2>&1
ls -la gpurun_out | head -40
