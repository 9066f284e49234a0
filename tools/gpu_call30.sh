set +e
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train_m.csv python tools/one_train_step.py yolov5m 16 640 fp16 > gpurun_out/ncu_train_m.log 2>&1
tail -2 gpurun_out/ncu_train_m.log
python tools/ncu_summary.py launches_bw gpurun_out/launches_train_m.csv gpurun_out/launches_train_m.md "yolov5m training step, 16 x 640x640, fp16 AMP" | head -40
Y5_ASYNC_WGRAD=1 timeout 300 python -m pytest tests/test_optim_gpu.py tests/test_train_gpu.py -q -k "graph" 2>&1 | tail -5
for a in 0 1; do
  Y5_ASYNC_WGRAD=$a timeout 300 python bench.py --workload yolov5m-train --steps 10 --warmup 3 > gpurun_out/bench_m_train_v30_a$a.json 2> gpurun_out/bench_m_train_v30_a$a.err
  tail -3 gpurun_out/bench_m_train_v30_a$a.err; A=$a python - <<'PY'
import json, os
d=json.load(open('gpurun_out/bench_m_train_v30_a%s.json' % os.environ['A']))
print('async', os.environ['A'], {k:d[k] for k in ('value','ms_per_step','launches_per_step')}, d['e2e'], d['roofline']['frac'], d['cuda_graph_step'], d['torch_cuda_reference_train'].get('value'))
PY
done
