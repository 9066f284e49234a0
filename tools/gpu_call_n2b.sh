set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_ddp_gpu.py tests/test_optim_gpu.py -q 2>&1 | tail -15
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err
tail -5 gpurun_out/bench_n2b.err
head -c 300 gpurun_out/bench_n2b.json
