"""Run under torchrun on N GPUs: every rank traces its contiguous shard of the global path range into a device histogram,
one NCCL reduce(SUM) brings the shards to rank 0, and rank 0 checks the result BIT FOR BIT against its own single-GPU trace
of the whole range and (small case) against the CPU oracle.  Covers the default mode, 64-source sharing and all-prefix mode.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/check_multigpu_exact.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-pathtracer_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import torch.distributed as dist
import frequensee as fs
from frequensee import scenes, capi
from frequensee.distributed import shard_range

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
cases = [("furnished_room", {}, 1, 300000, 16, 0), ("furnished_room", {}, 5, 100000, 16, capi.FLAG_SHARE_LISTENER),
         ("shoebox", {}, 1, 20000, 8, capi.FLAG_CONNECT_ALL), ("concert_hall", {"target_tris": 400000}, 3, 150000, 24, 0)]
for name, kw, S, n, depth, flags in cases:
    sc = scenes.by_name(name, **kw)
    src = np.stack([sc.sources[0] + np.float32([0.35 * i, 0.2 * i, 0.05 * i]) for i in range(S)]).astype(np.float32)
    with fs.Context(device=local, flags=flags) as ctx:
        ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
        B, K = ctx.cfg.n_bands, ctx.cfg.n_bins
        d_hist = torch.zeros((S, B, K), dtype=torch.int64, device="cuda")
        g0, gc = shard_range(S * n, rank, world)
        ctx.trace_range_device(src, sc.listener, n, g0, gc, depth, 4242, d_hist.data_ptr(), True)
        torch.cuda.synchronize()
        dist.reduce(d_hist, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            whole = ctx.trace(src, sc.listener, n, depth, 4242)
            got = d_hist.cpu().numpy().view(np.uint64)
            same = bool(np.array_equal(got, whole))
            msg = "%s S=%d n=%d depth=%d flags=%d world=%d: NCCL-reduced shards == single GPU: %s" % (name, S, n, depth, flags, world, same)
            if name == "shoebox":
                import pyoracle as po
                So = po.Scene(sc.verts, sc.tri_mat, sc.absorption, use_bvh=False)
                ho, _ = So.trace(po.default_config(flags=flags), src, sc.listener, n, depth, 4242, n_threads=16)
                same_o = bool(np.array_equal(got, ho)); msg += ", == CPU oracle: %s" % same_o
                same = same and same_o
            print(msg, flush=True)
            ok = ok and same
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
