import os, sys, time
sys.path.insert(0, "audio-pathtracer_b200")
import frequensee as fs
from frequensee import scenes, capi
sc = scenes.furnished_room()
for flags, tag in ((0, "endpoints"), (capi.FLAG_CONNECT_ALL, "all prefixes")):
    with fs.Context(flags=flags) as ctx:
        ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
        for i in range(3):
            ctx.trace(sc.sources[:1], sc.listener, 1 << 18, 16, 1000 + i, want_hist=False)
            st = ctx.stats()
        print(tag, "2^18 pairs depth 16: %.2f ms, %d shadow rays, %d connected" % (st["last_trace_ms"], st["shadow_rays"], st["connected"]))
