set +e
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
# launch list of the bench command (our kernels; build + warm-up + timed steps), and of exactly one forward
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-subrecords > gpurun_out/ncu8_b.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_fwd_l.csv python tools/one_forward.py yolov5l 64 640 bf16 > gpurun_out/ncu8_f.log 2>&1
timeout 400 ncu --metrics $M --clock-control none --profile-from-start off -k regex:conv_gemm --csv --log-file gpurun_out/r02_conv_metrics_l.csv python tools/one_forward.py yolov5l 64 640 bf16 > gpurun_out/ncu8_m.log 2>&1
timeout 400 ncu --metrics $M --clock-control none --profile-from-start off -k regex:conv_gemm --csv --log-file gpurun_out/r02_conv_metrics_s.csv python tools/one_forward.py yolov5s 32 640 fp16 > gpurun_out/ncu8_ms.log 2>&1
for op in model.6.m0.cv2 model.2.m0.cv1 model.2.m0.cv2 model.0; do
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm -c 1 -o gpurun_out/ncu8_$op python tools/op_probe.py yolov5l 64 640 bf16 $op 2 > gpurun_out/ncu8_$op.log 2>&1
  tail -1 gpurun_out/ncu8_$op.log
done
timeout 200 python tools/layer_profile.py yolov5l 64 640 bf16 > gpurun_out/layers_l_v8.txt 2>&1
timeout 200 python tools/layer_profile.py yolov5s 32 640 fp16 > gpurun_out/layers_s_v8.txt 2>&1
grep -E "fixed ops|full forward|detect" gpurun_out/layers_l_v8.txt gpurun_out/layers_s_v8.txt
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_*
