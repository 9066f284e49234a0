set +e
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train_m.csv python tools/one_train_step.py yolov5m 16 640 fp16 > gpurun_out/ncu_train_m.log 2>&1
tail -2 gpurun_out/ncu_train_m.log
python tools/ncu_summary.py launches_bw gpurun_out/launches_train_m.csv gpurun_out/launches_train_m.md "yolov5m training step, 16 x 640x640, fp16 AMP" | head -40
timeout 300 python bench.py --workload yolov5m-train --steps 10 --warmup 3 > gpurun_out/bench_m_train_v30.json 2> gpurun_out/bench_m_train_v30.err
tail -3 gpurun_out/bench_m_train_v30.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_m_train_v30.json'))
print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}, d['e2e'], d['roofline']['frac'], d['cuda_graph_step'], d['torch_cuda_reference_train'].get('value'))
PY
