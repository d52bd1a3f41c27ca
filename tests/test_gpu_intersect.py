"""Device BVH (LBVH build + stack traversal) against device brute force and the oracle: the
closest hit is min (t, triangle id) of an exact-sequence Moller-Trumbore, so all three must agree
bit for bit; the padded slab test only has to be conservative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rays(rng, lo, hi, n):
    o = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    return np.concatenate([o, d], axis=1)


@pytest.mark.parametrize("scene_name,kw,lo,hi", [
    ("furnished_room", {}, [0.3, 0.3, 0.2], [11.7, 8.7, 3.4]),
    ("mine_tunnels", {"target_tris": 200000}, None, None),
    ("concert_hall", {"target_tris": 300000}, [1, 1, 0.5], [47, 29, 17]),
])
def test_bvh_equals_brute_force_and_oracle(fs, oracle, scene_name, kw, lo, hi):
    from frequensee import scenes, capi
    rng = np.random.default_rng(42)
    sc = scenes.by_name(scene_name, **kw)
    n = 60000
    if lo is None:                                           # tunnels: origins near the source positions
        base = sc.sources[rng.integers(0, len(sc.sources), n)]
        rays = _rays(rng, [0, 0, 0], [1, 1, 1], n)
        rays[:, :3] = base + rng.uniform(-0.3, 0.3, size=(n, 3)).astype(np.float32)
    else:
        rays = _rays(rng, lo, hi, n)
    rays[:64, 3:] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 64)] * rng.choice([-1, 1], (64, 1))  # axis-parallel
    import os
    os.environ["FS_TUNE_W8"] = "1"               # FS_TUNE_* are read by fs_create: the 8-wide compressed nodes (optional format)
    try:
        w8 = fs.Context()
    finally:
        os.environ.pop("FS_TUNE_W8", None)
    with fs.Context() as a, fs.Context(flags=capi.FLAG_BRUTE_FORCE) as b, fs.Context(flags=capi.FLAG_FUSED_EXTEND) as c, w8:
        for ctx in (a, b, c, w8):
            ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
        t8, i8 = w8.closest_hits(rays)
        ta, ia = a.closest_hits(rays)            # production kernels: 4-wide quantised nodes, smem stack, ray replacement
        tc, ic = c.closest_hits(rays)            # per-thread traversal of the BVH2 float nodes
        nb = 6000
        tb, ib = b.closest_hits(rays[:nb])
        assert np.array_equal(ta, tc) and np.array_equal(ia, ic)
        assert np.array_equal(ta, t8) and np.array_equal(ia, i8)
        assert np.array_equal(ta[:nb], tb) and np.array_equal(ia[:nb], ib)
        assert (ia != 0xffffffff).mean() > 0.5
        tmax = (ta * rng.uniform(0.5, 1.5, n)).astype(np.float32)
        tmax[~np.isfinite(tmax)] = 10.0
        ha = a.any_hits(rays, tmax)
        assert np.array_equal(ha, c.any_hits(rays, tmax))
        assert np.array_equal(ha, w8.any_hits(rays, tmax))
        hb = b.any_hits(rays[:nb], tmax[:nb])
        assert np.array_equal(ha[:nb], hb)
    S = oracle.Scene(sc.verts, sc.tri_mat, sc.absorption, use_bvh=True)
    for i in range(3000):
        hit, t, tri = S.closest_hit(rays[i, :3], rays[i, 3:])
        assert hit == (ia[i] != 0xffffffff)
        if hit:
            assert np.float32(t) == ta[i] and tri == ia[i]
        assert S.any_hit(rays[i, :3], rays[i, 3:], tmax[i]) == ha[i]


def test_bvh_statistics(fs):
    from frequensee import scenes
    sc = scenes.furnished_room()
    with fs.Context() as ctx:
        ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
        st = ctx.stats()
    assert st["bvh_nodes"] == sc.n_tris - 1 and 1 <= st["bvh_max_leaf"] <= 4
