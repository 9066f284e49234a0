"""CPU, world_size 2 over gloo: the multi-GPU harness logic (batch sharding with no data-path collective, the
max-over-ranks timing reduction, rank-0 aggregation) -- the N>1 path of bench.py without GPUs."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yolov5_b200.parallel import aggregate_throughput, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(10, rank, world)
    ms = 5.0 + rank  # pretend rank 1 is slower
    total, worst = aggregate_throughput(images=hi - lo, ms=ms, device=torch.device("cpu"))
    q.put((rank, lo, hi, total, worst))
    dist.destroy_process_group()


def test_shard_and_aggregate_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(30) for p in ps]
    assert [(r[1], r[2]) for r in res] == [(0, 5), (5, 10)]          # disjoint, covering shards
    assert all(r[3] == 10 and r[4] == 6.0 for r in res)              # units summed, time = max over ranks


def test_shard_range_covers_everything():
    for total in (1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
