"""world_size-2 gloo test of the multi-process plumbing (runs on CPU): image sharding covers the batch
exactly, timings reduce as max over ranks, detections gather to rank 0 in image order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yolov6_b200.dist import GradSync, bucket_ranges, gather_detections, max_over_ranks, shard_range


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
    # gradient all-reduce of a flat buffer in contiguous buckets: every rank ends with the SUM of the per-rank buffers
    # (= loss * world_size followed by DDP's average, core/engine.py:171-172, 464-466)
    n = 1003 * 4
    flat = torch.arange(n, dtype=torch.float32) * (rank + 1)
    ranges = bucket_ranges(n, 3)
    sync = GradSync(flat, ranges)
    for k in range(len(ranges)):
        sync.bucket_ready(k)
    sync.finish()
    ok_sum = bool(torch.equal(flat, torch.arange(n, dtype=torch.float32) * sum(range(1, world + 1))))
    avg = torch.full((8,), float(rank))
    s2 = GradSync(avg, [(0, 8)], average=True)
    s2.all_ready()
    s2.finish()
    ok_avg = bool(torch.allclose(avg, torch.full((8,), (world - 1) / 2)))
    q.put((rank, lo, hi, slowest, None if o is None else (tuple(o.shape), c.tolist()), ok_sum and ok_avg, sync.bytes_per_step))
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
    (r0, lo0, hi0, s0, g0, ok0, nb0), (r1, lo1, hi1, s1, g1, ok1, nb1) = res
    assert ok0 and ok1 and nb0 == nb1 == 1003 * 4 * 4      # summed gradients on both ranks, whole buffer covered
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


def test_bucket_ranges_cover_exactly():
    for total in (0, 4, 1000, 4012, 20_000_000):
        for nb in (1, 2, 3, 7):
            r = bucket_ranges(total, nb)
            assert len(r) == nb and r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(lo % 4 == 0 for lo, _ in r)
