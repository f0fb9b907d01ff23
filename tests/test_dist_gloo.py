"""world_size-2 gloo test of the multi-process plumbing (runs on CPU): image sharding covers the batch
exactly, timings reduce as max over ranks, detections gather to rank 0 in image order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yolov6_b200.dist import gather_detections, max_over_ranks, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(33, rank, world)
    slowest = max_over_ranks(10.0 + rank)
    out = torch.full((hi - lo if hi - lo == 16 else 16, 4, 6), float(rank))
    cnt = torch.full((out.shape[0],), rank + 1, dtype=torch.int32)
    o, c = gather_detections(out, cnt)
    q.put((rank, lo, hi, slowest, None if o is None else (tuple(o.shape), c.tolist())))
    dist.destroy_process_group()


def test_two_rank_gloo_plumbing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, s0, g0), (r1, lo1, hi1, s1, g1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 17, 17, 33)         # exact cover, sizes differ by <= 1
    assert s0 == s1 == 11.0                                 # max over ranks on every rank
    assert g1 is None and g0[0] == (32, 4, 6) and g0[1] == [1] * 16 + [2] * 16


def test_shard_range_properties():
    for n in (0, 1, 7, 32, 33):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
