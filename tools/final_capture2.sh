set +e
K='regex:conv_gemm|conv_wgrad|nms_|stem_s2d|sppf|upsample2x|copy_view|nhwc_to_nchw|bn_act|col_stats|weight_pack|zero_stuff|loss_|f32_to_lowp|col_sum'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 210 -c 140 --csv --log-file gpurun_out/launches_s.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 4000 --csv --log-file gpurun_out/launches_train.csv python tools/train_bench.py --model yolov5s --batch 16 --steps 1 --skip-reference --train-only > gpurun_out/ncu_t.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 230 -c 4 -o gpurun_out/conv_full_l python bench.py --workload yolov5l --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fl.log 2>&1
ls -la gpurun_out | grep -v "^total"
