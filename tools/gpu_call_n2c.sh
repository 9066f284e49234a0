set +e
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 300 python -m pytest tests/test_train_ddp_gpu.py -q 2>&1 | tail -6
Y5_BENCH_GRAPH_DP=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_final.json 2> gpurun_out/bench_n2_final.err
echo "bench rc=$?"; tail -4 gpurun_out/bench_n2_final.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2_final.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e'])
t=d['train_ddp']; print('train', t['value'], t['ms_per_step'], t['collective'], t['cuda_graph_step'])
print('weak', d.get('weak_scaling'))
PY
