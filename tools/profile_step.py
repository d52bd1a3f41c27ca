"""Minimal driver for ncu: furnished room, W warm-up + S steps of the device-resident IR update."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-pathtracer_b200"))
import frequensee as fs
from frequensee import scenes, capi
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sc = scenes.by_name(sys.argv[3]) if len(sys.argv) > 3 else scenes.furnished_room()
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 16
import time
ctx = fs.Context(flags=flags | (0 if os.environ.get('PS_NOTIME') else capi.FLAG_TIME_KERNELS))   # PS_NOTIME=1: no per-kernel events (batch lanes overlap)
t0 = time.time(); ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption); print('commit %d tris %.3f s' % (sc.n_tris, time.time() - t0))
for i in range(steps):
    ctx.trace(sc.sources[:int(os.environ.get('PS_SOURCES', 1))], sc.listener, int(os.environ.get('PS_PATHS', 1 << 20)), depth, 1000 + i, want_hist=False)
    ctx.build_ir(0, want_ir=False)
    st = ctx.stats()
    print("step", i, {k: st[k] for k in ("last_trace_ms", "extend_ms", "connect_ms", "node_visits", "tri_tests", "ext_rays")})
ctx.close()
