"""Sweeps FS_TUNE_* environments over the room (2^20 pairs, depth 16) and the hall share (1 310 720 pairs, depth 32):
median last_trace_ms (CUDA events on the context stream around one fs_trace, production configuration) per setting, and a
histogram checksum so that every variant is seen to produce the same integers.
usage: python tools/tune_sweep.py [room|hall|both] 'K=V,K=V;K=V;...'   (';' separates settings, '' = defaults)"""
import os, sys, json, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-pathtracer_b200"))
import numpy as np
import frequensee as fs
from frequensee import scenes

which = sys.argv[1] if len(sys.argv) > 1 else "both"
settings = [dict(kv.split("=") for kv in s.split(",") if kv) for s in (sys.argv[2] if len(sys.argv) > 2 else "").split(";")]
jobs = []
if which in ("room", "both"):
    jobs.append(("room", scenes.furnished_room(), 1 << 20, 16))
if which in ("hall", "both"):
    jobs.append(("hall", scenes.concert_hall(), 1310720, 32))
if which == "mid":
    jobs.append(("room128k", scenes.furnished_room(), 1 << 17, 16))
    jobs.append(("room256k", scenes.furnished_room(), 1 << 18, 16))
    jobs.append(("hall164k", scenes.concert_hall(), 163840, 32))
if which == "big":
    jobs.append(("room2M", scenes.furnished_room(), 1 << 21, 16))
    jobs.append(("room512k", scenes.furnished_room(), 1 << 19, 16))
    jobs.append(("hall2.6M", scenes.concert_hall(), 2621440, 32))
    jobs.append(("hall655k", scenes.concert_hall(), 655360, 32))
if which == "small":
    jobs.append(("room64k", scenes.furnished_room(), 1 << 16, 16))
    jobs.append(("shoebox1M", scenes.shoebox(), 1 << 20, 8))
for name, sc, n, depth in jobs:
    for env in settings:
        kw = {k[4:]: int(v) for k, v in env.items() if k.startswith("cfg.")}      # cfg.max_batch_paths=262144 -> fs_config field
        env = {k: v for k, v in env.items() if not k.startswith("cfg.")}
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            ctx = fs.Context(**kw)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
        ms = []
        crc = None
        for i in range(7):
            h = ctx.trace(sc.sources[:1], sc.listener, n, depth, 1000 + (i % 2), want_hist=(i == 0))
            ctx.build_ir(0, want_ir=False)
            st = ctx.stats()
            ms.append(st["last_trace_ms"])
            if i == 0:
                crc = zlib.crc32(h.tobytes())
        ctx.close()
        print(json.dumps({"scene": name, "env": dict(env, **{"cfg." + k: v for k, v in kw.items()}), "trace_ms_median": float(np.median(ms[2:])), "min": float(min(ms[2:])), "crc": crc}), flush=True)
