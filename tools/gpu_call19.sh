set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_train_ddp_gpu.py > gpurun_out/pytest19_full.log 2>&1
tail -5 gpurun_out/pytest19_full.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v19.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_v19.txt 2>&1
grep -E "fixed ops" gpurun_out/layers_l_v19.txt gpurun_out/layers_s_v19.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default19.json 2> gpurun_out/bench_default19.err
tail -2 gpurun_out/bench_default19.err
head -c 300 gpurun_out/bench_default19.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02b_launches_fwd_l.csv python tools/one_forward.py yolov5l 64 640 bf16 > gpurun_out/ncu19_f.log 2>&1
timeout 400 ncu --metrics $M --clock-control none --profile-from-start off -k regex:conv_gemm --csv --log-file gpurun_out/r02b_conv_metrics_l.csv python tools/one_forward.py yolov5l 64 640 bf16 > gpurun_out/ncu19_m.log 2>&1
timeout 400 ncu --metrics $M --clock-control none --profile-from-start off -k regex:conv_gemm --csv --log-file gpurun_out/r02b_conv_metrics_s.csv python tools/one_forward.py yolov5s 32 640 fp16 > gpurun_out/ncu19_ms.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02b_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-subrecords > gpurun_out/ncu19_b.log 2>&1
for op in detect.0 model.4.m0.cv2 model.6.m0.cv2 model.4.m0.cv1; do
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm -c 1 -o gpurun_out/ncu19_$op python tools/op_probe.py yolov5l 64 640 bf16 $op 2 > gpurun_out/ncu19_$op.log 2>&1
  tail -1 gpurun_out/ncu19_$op.log
done
ls gpurun_out/r02b_* gpurun_out/ncu19_*.ncu-rep
