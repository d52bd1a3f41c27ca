"""Parity tests proper: the CUDA path, called through the C-ABI (include/frequensee.h), against
the CPU oracle on identical seeded inputs.  Histograms are integer: the bar is BIT-EXACT."""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def _dense(idx, val, shape):
    h = np.zeros(int(np.prod(shape)), np.uint64)
    h[idx] = val
    return h.reshape(shape)


@pytest.fixture(scope="module")
def shoebox():
    from frequensee import scenes
    return scenes.shoebox()


@pytest.fixture(scope="module")
def room():
    from frequensee import scenes
    return scenes.furnished_room()


def _ctx(fs, sc, **over):
    ctx = fs.Context(**over)
    ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
    return ctx


def test_config1_shoebox_bit_exact_vs_oracle_and_golden(fs, oracle, shoebox):
    """BASELINE config 1: 12 tris / 6 materials, 16k paths, depth 8, 1 s histogram"""
    g = _golden()
    S = oracle.Scene(shoebox.verts, shoebox.tri_mat, shoebox.absorption, use_bvh=False)
    for tag, rr in (("rr09", 0.9), ("rr10", 1.0)):
        with _ctx(fs, shoebox, rr_prob=rr) as ctx:
            h = ctx.trace(shoebox.sources, shoebox.listener, 16384, 8, 0x5EED)
            st = ctx.stats()
        ho, so = S.trace(oracle.default_config(rr_prob=rr), shoebox.sources, shoebox.listener, 16384, 8, 0x5EED,
                         n_threads=8)
        assert np.array_equal(h, ho)
        assert np.array_equal(h, _dense(g["shoebox_%s_idx" % tag], g["shoebox_%s_val" % tag], h.shape))
        assert [st["ext_rays"], st["shadow_rays"], st["connected"]] == [so["ext_rays"], so["shadow_rays"], so["connected"]]
        assert st["kernel_launches"] > 0


def test_per_path_records_bit_exact(fs, oracle, shoebox):
    g = _golden()
    for tag, rr in (("rr09", 0.9), ("rr10", 1.0)):
        with _ctx(fs, shoebox, rr_prob=rr) as ctx:
            dbg = ctx.trace_debug(shoebox.sources, shoebox.listener, 16384, 0, 64, 8, 0x5EED)
        assert dbg.tobytes() == g["shoebox_%s_dbg64" % tag].tobytes()


def test_every_traversal_mode_gives_the_same_histogram(fs, oracle, room):
    from frequensee import capi
    S = oracle.Scene(room.verts, room.tri_mat, room.absorption, use_bvh=True)
    ho, so = S.trace(oracle.default_config(), room.sources, room.listener, 8192, 16, 1, n_threads=16)
    g = _golden()
    assert np.array_equal(ho, _dense(g["room_idx"], g["room_val"], ho.shape))
    for flags in (0, capi.FLAG_FUSED_EXTEND, capi.FLAG_FUSED_EXTEND | capi.FLAG_SMEM_TREELET, capi.FLAG_NO_SPLAT_AGG,
                  capi.FLAG_COUNT_VISITS, capi.FLAG_COUNT_VISITS | capi.FLAG_FUSED_EXTEND, capi.FLAG_TIME_KERNELS):
        with _ctx(fs, room, flags=flags) as ctx:
            h = ctx.trace(room.sources, room.listener, 8192, 16, 1)
            st = ctx.stats()
        assert np.array_equal(h, ho), "flags=%d" % flags
        assert st["ext_rays"] == so["ext_rays"] and st["connected"] == so["connected"]
        if flags & capi.FLAG_COUNT_VISITS:
            assert st["node_visits"] > 0 and st["tri_tests"] > 0
    with _ctx(fs, room, flags=capi.FLAG_BRUTE_FORCE) as ctx:          # every triangle tested: small N
        h = ctx.trace(room.sources, room.listener, 512, 16, 1)
    hb, _ = S.trace(oracle.default_config(), room.sources, room.listener, 512, 16, 1, n_threads=16)
    assert np.array_equal(h, hb)


@pytest.mark.parametrize("over", [
    dict(n_bands=1), dict(n_bands=3), dict(n_bins=48000, bin_ms=1000.0 / 48000.0), dict(pdf_exponent=1.0),
    dict(min_seg=1.0), dict(energy_clamp=0.01, energy_gain=1.0), dict(rr_prob=0.5), dict(eps_offset=1e-2),
])
def test_config_variants_bit_exact(fs, oracle, over):
    from frequensee import scenes
    sc = scenes.furnished_room(target_tris=20000, n_bands=over.get("n_bands", 8))
    S = oracle.Scene(sc.verts, sc.tri_mat, sc.absorption, use_bvh=True)
    with _ctx(fs, sc, **over) as ctx:
        h = ctx.trace(sc.sources, sc.listener, 4096, 12, 77)
    ho, _ = S.trace(oracle.default_config(**over), sc.sources, sc.listener, 4096, 12, 77, n_threads=16)
    assert h.any() and np.array_equal(h, ho)


def test_multi_source_and_shard_invariance(fs, oracle, shoebox):
    """g = source * n_paths + i; any partition of g sums to the same integer histogram"""
    src = np.array([[1.5, 1.2, 1.0], [3.0, 2.0, 1.5], [6.0, 4.0, 2.0]], np.float32)
    n = 5000
    S = oracle.Scene(shoebox.verts, shoebox.tri_mat, shoebox.absorption, use_bvh=False)
    ho, _ = S.trace(oracle.default_config(), src, shoebox.listener, n, 8, 5, n_threads=8)
    with _ctx(fs, shoebox) as ctx:
        h = ctx.trace(src, shoebox.listener, n, 8, 5)
        assert np.array_equal(h, ho)
        for world in (2, 4, 8):
            from frequensee.distributed import shard_range
            acc = np.zeros_like(h)
            for r in range(world):
                lo, cnt = shard_range(3 * n, r, world)
                ctx.trace_range(src, shoebox.listener, n, lo, cnt, 8, 5, hist=acc)
            assert np.array_equal(acc, ho), "world=%d" % world
    with _ctx(fs, shoebox, max_batch_paths=1000) as ctx:                 # several wavefront batches
        assert np.array_equal(ctx.trace(src, shoebox.listener, n, 8, 5), ho)


def test_full_size_config2_linearity(fs, room):
    """BASELINE config 2 at full size (1M paths, depth 16, 8 bands): size-independent properties.
    hist([0,N)) == hist([0,N/2)) + hist([N/2,N)) bit for bit, and a second run reproduces it."""
    N = 1 << 20
    with _ctx(fs, room) as ctx:
        full = ctx.trace(room.sources, room.listener, N, 16, 2024)
        st = ctx.stats()
        again = ctx.trace(room.sources, room.listener, N, 16, 2024)
        acc = np.zeros_like(full)
        ctx.trace_range(room.sources, room.listener, N, 0, N // 2, 16, 2024, hist=acc)
        ctx.trace_range(room.sources, room.listener, N, N // 2, N - N // 2, 16, 2024, hist=acc)
    assert np.array_equal(full, again) and np.array_equal(full, acc)
    assert st["connected"] > 0 and st["ext_rays"] > 10 * N // 2
    # every band of a bin is populated together; energy per path is bounded by clamp * gain = 10
    assert int(full.sum(axis=2).max()) <= 10 * 2 ** 32 * st["connected"]


def test_full_size_config2_bit_exact_vs_oracle(fs, oracle, room):
    """BASELINE config 2 at FULL size -- 2^20 path pairs, depth 16, 8 bands, the default two-lane batching, i.e. exactly what
    bench.py times -- bit for bit against the CPU oracle on all host threads (queue wrap, batch split and lane overlap at
    full occupancy are what the small cases cannot exercise)"""
    N = 1 << 20
    S = oracle.Scene(room.verts, room.tri_mat, room.absorption, use_bvh=True)
    ho, so = S.trace(oracle.default_config(), room.sources, room.listener, N, 16, 1000, n_threads=os.cpu_count() or 8)
    with _ctx(fs, room) as ctx:
        h = ctx.trace(room.sources, room.listener, N, 16, 1000)
        st = ctx.stats()
        assert np.array_equal(h, ho)
        assert [st["ext_rays"], st["shadow_rays"], st["connected"]] == [so["ext_rays"], so["shadow_rays"], so["connected"]]
        # the shard form the multi-GPU path uses: 8 device-resident shards summed == the whole job
        from frequensee.distributed import shard_range
        acc = np.zeros_like(h)
        for r in range(8):
            lo, cnt = shard_range(N, r, 8)
            ctx.trace_range(room.sources, room.listener, N, lo, cnt, 16, 1000, hist=acc)
        assert np.array_equal(acc, ho)
    with _env(FS_TUNE_MEGA=1):
        ctx = _ctx(fs, room)
    with ctx:                                                           # the persistent per-batch kernel at full size
        assert np.array_equal(ctx.trace(room.sources, room.listener, N, 16, 1000), ho)


def test_edge_cases(fs, oracle, shoebox):
    empty_v, empty_m = np.zeros((0, 3, 3), np.float32), np.zeros(0, np.uint32)
    ab = np.full((1, 8), 0.5, np.float32)
    S0 = oracle.Scene(empty_v, empty_m, ab, use_bvh=False)
    with fs.Context() as ctx:                                           # empty scene: free field KAT on the GPU
        ctx.set_scene(empty_v, empty_m, ab)
        h = ctx.trace([[1.0, 2.0, 3.0]], [4.0, 6.0, 3.0], 64, 8, 7)
        ho, _ = S0.trace(oracle.default_config(), [[1.0, 2.0, 3.0]], [4.0, 6.0, 3.0], 64, 8, 7)
        assert np.array_equal(h, ho)
        assert list(np.flatnonzero(h[0, 0])) == [14]
        e = 10.0 / (4 * math.pi * 25.0) * math.exp(-1e-4 * 5.0)
        assert abs(int(h[0, 0, 14]) / 64 / 2.0 ** 32 / e - 1) < 1e-6
        h = ctx.trace([[0, 0, 0]], [400.0, 0, 0], 16, 8, 7)             # late energy -> last bin
        assert list(np.flatnonzero(h[0, 0])) == [999]
        h = ctx.trace([[1, 1, 1]], [1, 1, 1], 8, 8, 7)                  # coincident endpoints
        assert int(h[0, 0, 0]) == 8 * 10 * 2 ** 32
    one = np.array([[[0, 0, 2], [4, 0, 2], [0, 4, 2]]], np.float32)      # single triangle
    S1 = oracle.Scene(one, np.zeros(1, np.uint32), ab, use_bvh=False)
    with fs.Context() as ctx:
        ctx.set_scene(one, np.zeros(1, np.uint32), ab)
        h = ctx.trace([[1, 1, 1]], [1.5, 1.2, 0.5], 4096, 4, 3)
        ho, so = S1.trace(oracle.default_config(), [[1, 1, 1]], [1.5, 1.2, 0.5], 4096, 4, 3)
        assert np.array_equal(h, ho) and ctx.stats()["ext_rays"] == so["ext_rays"]
    S = oracle.Scene(shoebox.verts, shoebox.tri_mat, shoebox.absorption, use_bvh=False)
    with _ctx(fs, shoebox) as ctx:
        for depth in (0, 1, 2, 33):                                       # ragged depths incl. direct-only
            h = ctx.trace(shoebox.sources, shoebox.listener, 2048, depth, 9)
            ho, _ = S.trace(oracle.default_config(), shoebox.sources, shoebox.listener, 2048, depth, 9)
            assert np.array_equal(h, ho), "depth=%d" % depth
        h = ctx.trace(shoebox.sources, shoebox.listener, 1, 8, 9)           # a single path
        ho, _ = S.trace(oracle.default_config(), shoebox.sources, shoebox.listener, 1, 8, 9)
        assert np.array_equal(h, ho)


def test_error_behaviour(fs, shoebox):
    from frequensee import capi
    with fs.Context() as ctx:
        with pytest.raises(fs.FrequenSeeError) as ei:
            ctx.trace(shoebox.sources, shoebox.listener, 16, 8, 1)
        assert ei.value.code == capi.FS_ERR_STATE                           # trace before commit
        with pytest.raises(fs.FrequenSeeError):
            ctx.set_scene(shoebox.verts, shoebox.tri_mat + 100, shoebox.absorption)     # material id out of range
        with pytest.raises(fs.FrequenSeeError):
            ctx.set_scene(shoebox.verts, shoebox.tri_mat, shoebox.absorption[:, :4])   # band count mismatch
        with pytest.raises(fs.FrequenSeeError):
            ctx.set_scene(shoebox.verts, shoebox.tri_mat, shoebox.absorption * 0 + 1.5)  # absorption > 1
        ctx.set_scene(shoebox.verts, shoebox.tri_mat, shoebox.absorption)
        with pytest.raises(fs.FrequenSeeError) as ei:
            ctx.trace(shoebox.sources, shoebox.listener, 0, 8, 1)
        assert ei.value.code == capi.FS_ERR_INVALID
        with pytest.raises(fs.FrequenSeeError):
            ctx.build_ir(5)                                                 # no histogram for that source
    with pytest.raises(fs.FrequenSeeError):
        fs.Context(n_bands=9)
    with pytest.raises(fs.FrequenSeeError):
        fs.Context(conv_block=1000)


class _env:
    """FS_TUNE_* knobs are read by fs_create: set them around the creation of a context"""
    def __init__(self, **kv):
        self.kv = {k: str(v) for k, v in kv.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_every_traversal_kernel_and_bvh_layout_gives_the_same_histogram(fs, oracle, room):
    """queue kernel (default) / phased kernel, greedy / grandchild collapse, dense / sparse wide-node layout,
    BVH2 float nodes, LBVH with multi-triangle leaves: the closest hit does not depend on the structure"""
    S = oracle.Scene(room.verts, room.tri_mat, room.absorption, use_bvh=True)
    ho, so = S.trace(oracle.default_config(), room.sources, room.listener, 8192, 16, 1, n_threads=16)
    for env in ({}, {"FS_TUNE_TQ": 0}, {"FS_TUNE_TQ": 1}, {"FS_TUNE_COLLAPSE": 0}, {"FS_TUNE_COLLAPSE": 1}, {"FS_TUNE_COLLAPSE": 16}, {"FS_TUNE_PLOC_R": 3}, {"FS_TUNE_ROTATE": 3}, {"FS_TUNE_MEGA": 1}, {"FS_TUNE_MEGA": 1, "FS_TUNE_TEX": 0}, {"FS_TUNE_COLLAPSE": 3}, {"FS_TUNE_TQ": 0, "FS_TUNE_COLLAPSE": 2},
                {"FS_TUNE_WIDE": 0}, {"FS_TUNE_BUILDER": 0}, {"FS_TUNE_BUILDER": 0, "FS_TUNE_LEAF_MAX": 1},
                {"FS_TUNE_TQ_FLUSH": 1}, {"FS_TUNE_TQ_FLUSH": 32, "FS_TUNE_TQ_NODE_MIN": 0}, {"FS_TUNE_REFILL": 1},
                {"FS_TUNE_L2PIN": 8}, {"FS_TUNE_W8": 1}, {"FS_TUNE_W8": 1, "FS_TUNE_MEGA": 1}, {"FS_TUNE_W8": 1, "FS_TUNE_COLLAPSE": 9},
                {"FS_TUNE_W8": 1, "FS_TUNE_TQ_FLUSH": 1}):
        with _env(**env):
            ctx = _ctx(fs, room)
        with ctx:
            h = ctx.trace(room.sources, room.listener, 8192, 16, 1)
            st = ctx.stats()
        assert np.array_equal(h, ho), "env=%r" % (env,)
        assert st["ext_rays"] == so["ext_rays"] and st["connected"] == so["connected"], "env=%r" % (env,)


def test_batch_lanes_and_batch_splits_are_bit_exact(fs, oracle, shoebox):
    """jobs above 2^18 pairs run as equal batches on two lanes (streams): same integers as one lane, as many
    small batches, and as the oracle"""
    n = 600000
    S = oracle.Scene(shoebox.verts, shoebox.tri_mat, shoebox.absorption, use_bvh=False)
    ho, so = S.trace(oracle.default_config(), shoebox.sources, shoebox.listener, n, 8, 77, n_threads=16)
    for env, over in (({}, {}), ({"FS_TUNE_STREAMS": 1}, {}), ({"FS_TUNE_STREAMS": 3}, {"max_batch_paths": 70001}),
                      ({"FS_TUNE_STREAMS": 4}, {}), ({"FS_TUNE_STREAMS": 2}, {"max_batch_paths": 1 << 16}),
                      ({"FS_TUNE_MEGA": 1}, {}), ({"FS_TUNE_MEGA": 1, "FS_TUNE_STREAMS": 1}, {"max_batch_paths": 100000})):
        with _env(**env):
            ctx = _ctx(fs, shoebox, **over)
        with ctx:
            h = ctx.trace(shoebox.sources, shoebox.listener, n, 8, 77)
            st = ctx.stats()
            h2 = ctx.trace(shoebox.sources, shoebox.listener, n, 8, 77)       # buffers reused, lanes re-joined
        assert np.array_equal(h, ho), "env=%r over=%r" % (env, over)
        assert np.array_equal(h2, ho)
        assert [st["ext_rays"], st["shadow_rays"], st["connected"]] == [so["ext_rays"], so["shadow_rays"], so["connected"]]


def test_connect_all_prefixes_bit_exact(fs, oracle, shoebox, room):
    """FS_FLAG_CONNECT_ALL: every prefix pair connected, weight 1 / (s + t - 1); bit-exact against
    the oracle, through both connection-ray kernels, with batch splits, and for the counters"""
    from frequensee import capi
    ALL = capi.FLAG_CONNECT_ALL
    for sc, n, depth, use_bvh in ((shoebox, 4096, 8, False), (room, 1500, 16, True), (shoebox, 777, 0, False), (shoebox, 300, 33, False)):
        S = oracle.Scene(sc.verts, sc.tri_mat, sc.absorption, use_bvh=use_bvh)
        ho, so = S.trace(oracle.default_config(flags=oracle.FLAG_CONNECT_ALL), sc.sources[:1], sc.listener, n, depth, 13, n_threads=16)
        for env, over in (({}, {}), ({"FS_TUNE_TQ": 0}, {}), ({}, {"max_batch_paths": 500}), ({}, {"flags": ALL | capi.FLAG_COUNT_VISITS})):
            with _env(**env):
                ctx = _ctx(fs, sc, **dict({"flags": ALL}, **over))
            with ctx:
                h = ctx.trace(sc.sources[:1], sc.listener, n, depth, 13)
                st = ctx.stats()
            assert np.array_equal(h, ho), "n=%d depth=%d env=%r over=%r" % (n, depth, env, over)
            assert [st["ext_rays"], st["shadow_rays"], st["connected"]] == [so["ext_rays"], so["shadow_rays"], so["connected"]]
    with _ctx(fs, shoebox, flags=ALL) as ctx:
        with pytest.raises(fs.FrequenSeeError):
            ctx.trace(shoebox.sources, shoebox.listener, 64, 64, 1)              # depth > 63 does not fit the (s, t) ids


def test_shared_listener_subpaths_bit_exact(fs, oracle, shoebox, room):
    """FS_FLAG_SHARE_LISTENER: listener subpaths keyed by the path index only and traced once per call;
    same integers as the oracle (which regenerates them per source with the same keying), fewer rays"""
    from frequensee import capi
    SH = capi.FLAG_SHARE_LISTENER
    srcs = np.array([[1.5, 1.2, 1.0], [3.0, 2.0, 1.5], [5.5, 4.0, 2.0], [2.0, 4.2, 0.8]], np.float32)
    S = oracle.Scene(shoebox.verts, shoebox.tri_mat, shoebox.absorption, use_bvh=False)
    cfgo = oracle.default_config(flags=oracle.FLAG_SHARE_LISTENER)
    n = 3000
    ho, so = S.trace(cfgo, srcs, shoebox.listener, n, 8, 17, n_threads=16)
    h_plain, _ = S.trace(oracle.default_config(), srcs, shoebox.listener, n, 8, 17, n_threads=16)
    assert np.array_equal(ho[0], h_plain[0]) and not np.array_equal(ho[1], h_plain[1])       # source 0 keeps its stream
    for over in ({}, {"max_batch_paths": 700}, {"flags": SH | capi.FLAG_COUNT_VISITS}):
        with _ctx(fs, shoebox, **dict({"flags": SH}, **over)) as ctx:
            h = ctx.trace(srcs, shoebox.listener, n, 8, 17)
            st = ctx.stats()
            assert np.array_equal(h, ho), "over=%r" % (over,)
            assert st["connected"] == so["connected"] and st["shadow_rays"] == so["shadow_rays"]
            assert st["ext_rays"] < 0.7 * so["ext_rays"]                                      # 4 + 1 subpaths per index instead of 8
            # shards: a range inside one source (traced in place) and ranges across sources (cache) sum to the whole
            parts = np.zeros_like(ho)
            for g0, gc in ((0, 1000), (1000, 5500), (6500, 4 * n - 6500)):
                parts += ctx.trace_range(srcs, shoebox.listener, n, g0, gc, 8, 17)
            assert np.array_equal(parts, ho)
            h1 = ctx.trace(srcs[:1], shoebox.listener, n, 8, 17)                             # one source: nothing to share
            assert np.array_equal(h1[0], ho[0])
    Sr = oracle.Scene(room.verts, room.tri_mat, room.absorption, use_bvh=True)
    rs = np.array([room.sources[0], room.sources[0] + np.float32([0.7, 0.4, 0.1])], np.float32)
    hr, sr = Sr.trace(cfgo, rs, room.listener, 2048, 16, 3, n_threads=16)
    with _ctx(fs, room, flags=SH) as ctx:
        assert np.array_equal(ctx.trace(rs, room.listener, 2048, 16, 3), hr)


def test_full_material_schema_is_carried_but_unused(fs, oracle, shoebox):
    """MAT.h:16-34: Transmission / Scattering / ThicknessCm travel through the ABI, are validated, and -- as in the
    reference's tracers -- do not change the result"""
    import ctypes as C
    M, B = shoebox.absorption.shape
    with _ctx(fs, shoebox) as ctx:
        h0 = ctx.trace(shoebox.sources, shoebox.listener, 2048, 8, 4)
        ab = np.ascontiguousarray(shoebox.absorption, np.float32)
        tr = np.full((M, B), 0.3, np.float32); scat = np.full((M, B), 0.7, np.float32); th = np.full(M, 2.5, np.float32)
        assert ctx.L.fs_scene_set_materials_ex(ctx.h, ab.ctypes.data, tr.ctypes.data, scat.ctypes.data, th.ctypes.data, M, B) == 0
        assert ctx.L.fs_scene_commit(ctx.h) == 0
        assert np.array_equal(ctx.trace(shoebox.sources, shoebox.listener, 2048, 8, 4), h0)
        tr[1, 2] = 1.5
        assert ctx.L.fs_scene_set_materials_ex(ctx.h, ab.ctypes.data, tr.ctypes.data, None, None, M, B) != 0
