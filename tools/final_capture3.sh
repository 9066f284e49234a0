set +e
timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err
timeout 300 python bench.py --workload yolov5s-train --steps 10 --warmup 3 > gpurun_out/bench_s_train.json 2> gpurun_out/bench_s_train.err
timeout 300 python bench.py --workload yolov5m-train --steps 10 --warmup 3 > gpurun_out/bench_m_train.json 2> gpurun_out/bench_m_train.err
for r in 2 3 4 6; do for e in 4 6 8; do echo "== RED_BPS=$r ELT_BPS=$e"; Y5_BN_RED_BPS=$r Y5_BN_ELT_BPS=$e timeout 100 python tools/train_kernel_probe.py 2>&1 | sed -n 2,10p; done; done > gpurun_out/bn_sweep.txt 2>&1
for g in 3 5 8; do for k in 40 56 80; do echo "== GROUP_MAX=$g STAGE_KB=$k"; Y5_WG_GROUP_MAX=$g Y5_WG_STAGE_KB=$k timeout 100 python tools/train_kernel_probe.py 2>&1 | sed -n 12,30p; done; done > gpurun_out/wg_sweep.txt 2>&1
