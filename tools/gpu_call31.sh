set +e
mkdir -p gpurun_out
T="tests/test_train_gpu.py::test_graphed_train_step_matches_eager"
echo "== async"; Y5_ASYNC_WGRAD=1 timeout 200 python -m pytest $T -q 2>&1 | grep -E "^E |passed|failed" | cut -c1-400 | head -12
echo "== async, join now"; Y5_ASYNC_WGRAD=1 Y5_ASYNC_WGRAD_JOIN_NOW=1 timeout 200 python -m pytest $T -q 2>&1 | grep -E "^E |passed|failed" | cut -c1-400 | head -12
echo "== async, no PDL"; Y5_ASYNC_WGRAD=1 Y5_PDL=0 timeout 200 python -m pytest $T -q 2>&1 | grep -E "^E |passed|failed" | cut -c1-400 | head -12
echo "== async, no pack plan"; Y5_ASYNC_WGRAD=1 Y5_TRAIN_PACK_PLAN=0 timeout 200 python -m pytest $T -q 2>&1 | grep -E "^E |passed|failed" | cut -c1-400 | head -12
