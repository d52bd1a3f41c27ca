"""SURVEY 8f rank 3 -- FS_FLAG_MATERIAL_MODEL: Transmission / Scattering / ThicknessCm of the material asset (MAT.h:26-33) drive
the walk: pass-through (COMP.cpp:271-275), mirror reflection (GetReflectionVector, COMP.cpp:186) or the cosine lobe, with the
authors' energy split (MaterialAcousticProcessor.cpp:50-66).  Known-answer tests of the oracle's restatement (CPU), then
bit-exact parity of the CUDA path (GPU)."""
import math

import numpy as np
import pytest

EV_DIFFUSE, EV_SPECULAR, EV_TRANSMIT = 0, 1, 2


def _model_cfg(oracle, **over):
    flags = over.pop("flags", 0) | oracle.FLAG_MATERIAL_MODEL
    return oracle.default_config(flags=flags, **over)


def two_walls():
    """a 100 x 100 m sheet at x = 2 (material 0) and another at x = 6 (material 1); free space otherwise"""
    def sheet(x):
        a, b, c, d = [x, -50, -50], [x, 50, -50], [x, 50, 50], [x, -50, 50]
        return [[a, b, c], [a, c, d]]
    verts = np.array(sheet(2.0) + sheet(6.0), np.float32)
    return verts, np.array([0, 0, 1, 1], np.uint32)


def test_defaults_reproduce_the_reference_model(oracle):
    from frequensee import scenes
    sc = scenes.furnished_room(target_tris=12000)
    S = oracle.Scene(sc.verts, sc.tri_mat, sc.absorption, use_bvh=True)
    h0, s0 = S.trace(oracle.default_config(), sc.sources, sc.listener, 2048, 12, 3, n_threads=8)
    S.set_material_model()                                             # tau = 0, sigma = 1, 2.5 cm
    h1, s1 = S.trace(_model_cfg(oracle), sc.sources, sc.listener, 2048, 12, 3, n_threads=8)
    assert h0.any() and np.array_equal(h0, h1) and s0 == s1


def test_mirror_walls_follow_the_reflection_law(oracle):
    from frequensee import scenes
    sc = scenes.shoebox()
    S = oracle.Scene(sc.verts, sc.tri_mat, sc.absorption, use_bvh=False)
    S.set_material_model(scattering=np.zeros_like(sc.absorption))      # sigma = 0: every reflection is a mirror
    cfg = _model_cfg(oracle)
    checked = 0
    for g in range(200):
        F, B = S.debug_path(cfg, sc.sources[0], sc.listener, g, 1000, 12, 5)
        for N in (F, B):
            n = len(N["p"])
            assert all(N["ev"][1:n - 1] == EV_SPECULAR) and N["ev"][n - 1] == EV_DIFFUSE
            for i in range(1, n - 1):                                  # in -> node i -> out
                din = N["p"][i] - N["p"][i - 1]; dout = N["p"][i + 1] - N["p"][i]
                if min(np.linalg.norm(din), np.linalg.norm(dout)) < 0.5:
                    continue                                           # the 1 mm node offsets tilt very short segments
                din /= np.linalg.norm(din); dout /= np.linalg.norm(dout)
                same, neg = np.isclose(din, dout, atol=1e-2), np.isclose(din, -dout, atol=1e-2)
                assert (same | neg).all() and (neg & ~same).sum() == 1   # a shoebox wall flips exactly one component
                # discrete event: P = rr * P_spec = 0.9 * 1
                assert abs(N["prob"][i + 1] - 0.9) < 1e-6
                checked += 1
    assert checked > 500
    # nothing but the direct path can be connected through a mirror: every other end node has a zero diffuse factor
    h, st = S.trace(cfg, sc.sources, sc.listener, 4096, 8, 1, n_threads=4)
    d = np.linalg.norm(np.asarray(sc.listener, np.float64) - np.asarray(sc.sources[0], np.float64))
    assert list(np.flatnonzero(h[0, 0])) == [int(d / 343.0 * 1000)]


def test_pass_through_keeps_the_direction_and_applies_the_layer_gain(oracle):
    verts, tri_mat = two_walls()
    ab = np.array([[0.5] * 8, [0.2] * 8], np.float32)
    tr = np.array([[0.5] * 8, [0.0] * 8], np.float32)
    S = oracle.Scene(verts, tri_mat, ab, use_bvh=False)
    S.set_material_model(transmission=tr, scattering=np.ones_like(ab), thickness_cm=np.array([5.0, 2.5], np.float32))
    cfg = _model_cfg(oracle, min_seg=1e-6)
    src, lis = [0.0, 0.0, 0.0], [4.0, 1.0, 0.5]
    found = 0
    for g in range(400):
        F, B = S.debug_path(cfg, src, lis, g, 1000, 6, 9)
        if len(F["p"]) >= 3 and F["ev"][1] == EV_TRANSMIT and F["mat"][1] == 0 and F["mat"][2] == 1:
            x0, x1, x2 = F["p"][0].astype(np.float64), F["p"][1].astype(np.float64), F["p"][2].astype(np.float64)
            d01 = (x1 - x0) / np.linalg.norm(x1 - x0); d12 = (x2 - x1) / np.linalg.norm(x2 - x1)
            assert np.allclose(d01, d12, atol=1e-3)                    # straight on
            assert abs(F["ro"][1][0] - (2.0 + 1e-3)) < 1e-5 and abs(F["p"][1][0] - (2.0 - 1e-3)) < 1e-5   # node in front, ray from behind
            # lobe probabilities: mean Refl 0.5, mean tau 0.5 -> P_T = 0.5, P_D = 0.5 (sigma = 1)
            assert abs(F["prob"][2] - 0.9 * 0.5) < 1e-6
            found += 1
            if len(F["p"]) == 3 and len(B["p"]) == 1:
                # S -> x1 (through) -> x2 (ends) -> L, evaluated by hand: tau_eff = 0.5 ^ (5 / 2.5) = 0.25
                h, st, dbg = S.trace(cfg, [src], lis, 1000, 6, 9, g_first=g, g_count=1, debug=True)
                if dbg[0]["connected"]:
                    d0 = np.linalg.norm(x1 - x0); d1 = np.linalg.norm(x2 - F["ro"][1].astype(np.float64)); d2 = np.linalg.norm(np.array(lis) - x2)
                    air = 1e-4
                    G = lambda d: 1.0 / (4 * math.pi * d * d)          # noqa: E731
                    e = (1.0 * G(d0) * math.exp(-air * d0) / F["prob"][0] ** 0.1) * \
                        (0.25 * G(d1) * math.exp(-air * d1) / F["prob"][1] ** 0.1) * \
                        (0.8 * 1.0 / math.pi * G(d2) * math.exp(-air * d2) / F["prob"][2] ** 0.1)
                    assert abs(dbg[0]["energy"][0] / (min(e, 1.0) * 10.0) - 1.0) < 1e-4
                    found += 100
    assert found % 100 >= 20 and found >= 100                         # both the geometric and the energy case were seen


def test_lobe_split_is_energy_conserving(oracle):
    """tau is limited to 1 - Refl (MaterialAcousticProcessor.cpp:59-60); the three event probabilities add up to one"""
    rng = np.random.default_rng(1)
    ab = rng.uniform(0, 1, (5, 8)).astype(np.float32)
    verts, tri_mat = two_walls()
    S = oracle.Scene(verts, np.array([0, 1, 2, 3], np.uint32), ab, use_bvh=False)
    S.set_material_model(transmission=rng.uniform(0, 1, ab.shape).astype(np.float32), scattering=rng.uniform(0, 1, ab.shape).astype(np.float32),
                         thickness_cm=rng.uniform(0.5, 20, 5).astype(np.float32))
    h, st = S.trace(_model_cfg(oracle), [[0, 0, 0]], [4, 1, 0.5], 2048, 8, 2, n_threads=4)
    assert h.any() and st["connected"] > 0


# ---- the CUDA path -----------------------------------------------------------------------------------------------------
def _gpu_ctx(fs, sc_verts, sc_mat, ab, tr, scat, th, **over):
    from frequensee import capi
    over["flags"] = over.get("flags", 0) | capi.FLAG_MATERIAL_MODEL
    ctx = fs.Context(**over)
    ctx.set_scene_ex(sc_verts, sc_mat, ab, tr, scat, th)
    return ctx


@pytest.mark.gpu
def test_gpu_material_model_bit_exact(fs, oracle):
    from frequensee import scenes, capi
    rng = np.random.default_rng(6)
    room = scenes.furnished_room(target_tris=30000)
    M, B = room.absorption.shape
    tr = rng.uniform(0, 0.6, (M, B)).astype(np.float32); tr[0] = 0.0
    scat = rng.uniform(0, 1, (M, B)).astype(np.float32); scat[1] = 1.0; scat[2] = 0.0
    th = rng.uniform(0.5, 12.0, M).astype(np.float32); th[3] = 2.5
    S = oracle.Scene(room.verts, room.tri_mat, room.absorption, use_bvh=True)
    S.set_material_model(tr, scat, th)
    for oflags, gflags, env, n, depth in ((0, 0, {}, 20000, 16), (0, 0, {"FS_TUNE_MEGA": "1"}, 20000, 16),
                                          (oracle.FLAG_CONNECT_ALL, capi.FLAG_CONNECT_ALL, {}, 1500, 12),
                                          (0, capi.FLAG_COUNT_VISITS, {}, 6000, 33)):
        ho, so = S.trace(_model_cfg(oracle, flags=oflags), room.sources, room.listener, n, depth, 21, n_threads=16)
        import os
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            ctx = _gpu_ctx(fs, room.verts, room.tri_mat, room.absorption, tr, scat, th, flags=gflags)
        finally:
            for k, v in old.items():
                os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
        with ctx:
            h = ctx.trace(room.sources, room.listener, n, depth, 21)
            st = ctx.stats()
            assert ho.any() and np.array_equal(h, ho), (oflags, env)
            assert [st["ext_rays"], st["shadow_rays"], st["connected"]] == [so["ext_rays"], so["shadow_rays"], so["connected"]]
            if not oflags and not gflags and not env:
                dg = ctx.trace_debug(room.sources, room.listener, n, 0, 512, depth, 21)
                _, _, do = S.trace(_model_cfg(oracle), room.sources, room.listener, n, depth, 21, g_first=0, g_count=512, debug=True)
                assert dg.tobytes() == do.tobytes()
    # shared listener subpaths with the model, several sources
    srcs = np.array([room.sources[0], room.sources[0] + np.float32([0.7, 0.4, 0.1]), room.sources[0] + np.float32([1.5, -0.3, 0.2])], np.float32)
    ho, _ = S.trace(_model_cfg(oracle, flags=oracle.FLAG_SHARE_LISTENER), srcs, room.listener, 3000, 12, 4, n_threads=16)
    with _gpu_ctx(fs, room.verts, room.tri_mat, room.absorption, tr, scat, th, flags=capi.FLAG_SHARE_LISTENER) as ctx:
        assert np.array_equal(ctx.trace(srcs, room.listener, 3000, 12, 4), ho)
    # an open scene: rays that pass through a sheet and then miss everything end at the node in front of it
    verts, tri_mat = two_walls()
    ab = np.array([[0.5] * 8, [0.2] * 8], np.float32)
    tr2 = np.array([[0.5] * 8, [0.3] * 8], np.float32)
    S2 = oracle.Scene(verts, tri_mat, ab, use_bvh=False)
    S2.set_material_model(tr2, np.full_like(ab, 0.5), np.array([5.0, 1.0], np.float32))
    ho, so = S2.trace(_model_cfg(oracle), [[0, 0, 0]], [4, 1, 0.5], 30000, 8, 5, n_threads=8)
    with _gpu_ctx(fs, verts, tri_mat, ab, tr2, np.full_like(ab, 0.5), np.array([5.0, 1.0], np.float32)) as ctx:
        h = ctx.trace([[0, 0, 0]], [4, 1, 0.5], 30000, 8, 5)
        assert ho.any() and np.array_equal(h, ho) and ctx.stats()["connected"] == so["connected"]
    # the flag with the asset's defaults is the default mode
    with fs.Context(flags=capi.FLAG_MATERIAL_MODEL) as a, fs.Context() as b:
        a.set_scene(room.verts, room.tri_mat, room.absorption); b.set_scene(room.verts, room.tri_mat, room.absorption)
        assert np.array_equal(a.trace(room.sources, room.listener, 8192, 16, 2), b.trace(room.sources, room.listener, 8192, 16, 2))
    with pytest.raises(fs.FrequenSeeError):
        with fs.Context(flags=capi.FLAG_MATERIAL_MODEL | capi.FLAG_FUSED_EXTEND) as c:
            c.set_scene(room.verts, room.tri_mat, room.absorption)
            c.trace(room.sources, room.listener, 16, 4, 1)
