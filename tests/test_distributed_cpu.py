"""N>1 host logic on CPU: world_size-2 gloo process group, each rank traces its shard (with the
oracle standing in for the device tracer -- this test is about the sharding + integer reduce
plumbing in frequensee/distributed.py), histograms are reduced, result must equal the 1-rank one."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly(fs):
    from frequensee.distributed import shard_range
    for total in (0, 1, 7, 64, 1 << 20, (1 << 26) + 3):
        for world in (1, 2, 3, 4, 8):
            nxt = 0
            for r in range(world):
                lo, cnt = shard_range(total, r, world)
                assert lo == nxt
                nxt += cnt
            assert nxt == total
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "audio-pathtracer_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import pyoracle as po
    from frequensee import scenes
    from frequensee.distributed import shard_range, reduce_histogram
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = scenes.shoebox()
    cfg = po.default_config()
    S = po.Scene(sc.verts, sc.tri_mat, sc.absorption, use_bvh=False)
    src = np.array([[1.5, 1.2, 1.0], [5.5, 3.9, 2.2]], np.float32)
    n = 1500
    lo, cnt = shard_range(len(src) * n, rank, world)
    h, _ = S.trace(cfg, src, sc.listener, n, 8, 9, g_first=lo, g_count=cnt)
    t = torch.from_numpy(h.view(np.int64).copy())
    reduce_histogram(t, dst=0)
    if rank == 0:
        full, _ = S.trace(cfg, src, sc.listener, n, 8, 9)
        q.put(bool(np.array_equal(t.numpy().view(np.uint64), full)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_reduce_is_bit_exact():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
