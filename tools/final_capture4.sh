set +e
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -3 gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['clocks'])
t=d['train_ddp']; print('train', t['value'], t['ms_per_step'], t['launches_per_step'], t['roofline']['frac'], t['cuda_graph_step'], t['torch_cuda_reference_train'].get('value'))
print('config2', d['config2']['value'], d['config2']['roofline']['frac'])
PY
