"""Differential check at full scene size: the queue kernels (default) against the phased kernels (FS_TUNE_TQ=0) on random
rays -- closest hit (t, original triangle id) and any hit must agree bit for bit.  usage: kernel_agreement.py [scene] [n_rays]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-pathtracer_b200"))
import numpy as np
import frequensee as fs
from frequensee import scenes
name = sys.argv[1] if len(sys.argv) > 1 else "concert_hall"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000000
sc = scenes.by_name(name)
lo, hi = sc.verts.reshape(-1, 3).min(0), sc.verts.reshape(-1, 3).max(0)
rng = np.random.default_rng(7)
o = rng.uniform(lo + 0.05 * (hi - lo), hi - 0.05 * (hi - lo), size=(n, 3)).astype(np.float32)
d = rng.normal(size=(n, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
rays = np.concatenate([o, d], axis=1)
res = {}
for tq in ("0", "2"):
    os.environ["FS_TUNE_TQ"] = tq
    with fs.Context() as ctx:
        ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
        t, i = ctx.closest_hits(rays)
        tmax = np.where(np.isfinite(t), t * rng.uniform(0.5, 1.5, n), 10.0).astype(np.float32) if tq == "0" else res["tmax"]
        h = ctx.any_hits(rays, tmax)
        res[tq] = (t, i, h); res["tmax"] = tmax
(t0, i0, h0), (t2, i2, h2) = res["0"], res["2"]
print("%s: %d triangles, %d rays: closest t equal %s, triangle equal %s, any-hit equal %s; hit fraction %.3f, occluded fraction %.3f"
      % (name, sc.n_tris, n, np.array_equal(t0, t2), np.array_equal(i0, i2), np.array_equal(h0, h2),
         float((i0 != 0xffffffff).mean()), float(h0.mean())))
assert np.array_equal(t0, t2) and np.array_equal(i0, i2) and np.array_equal(h0, h2)
