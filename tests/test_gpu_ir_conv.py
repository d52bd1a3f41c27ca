"""IR reconstruction and partitioned FFT convolution on the device vs the oracle.
Floating point: tolerance is BASELINE.json's 1e-5 relative L2."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _rel(a, b):
    dt = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    return float(np.linalg.norm(np.asarray(a, dt) - np.asarray(b, dt)) / max(np.linalg.norm(np.asarray(b, dt)), 1e-30))


def test_ir_from_trace_matches_oracle(fs, oracle):
    from frequensee import scenes
    sc = scenes.shoebox()
    with fs.Context() as ctx:
        ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
        h = ctx.trace(sc.sources, sc.listener, 16384, 8, 0x5EED)
        ir = ctx.build_ir(0)
        # caller-owned, page-locked result buffers (fs_host_alloc), reused across updates: the same integers and floats
        hb = fs.capi.host_alloc(h.shape, np.uint64); ib = fs.capi.host_alloc(ir.shape)
        for _ in range(2):
            hb[:] = 7; ib[:] = -1.0
            assert ctx.trace(sc.sources, sc.listener, 16384, 8, 0x5EED, out=hb) is hb and ctx.build_ir(0, out=ib) is ib
            assert np.array_equal(np.asarray(hb), h) and np.array_equal(np.asarray(ib), ir)
        with pytest.raises(ValueError):
            ctx.build_ir(0, out=np.zeros((2, 100), np.float32))
        # fs_update = UpdateSource in one call (trace + IR rebuild, one synchronisation): same results, pageable or page-locked
        ib3 = fs.capi.host_alloc((1,) + ir.shape)
        hb[:] = 7; ib3[:] = -1.0
        ctx.update(sc.sources, sc.listener, 16384, 8, 0x5EED, hist_out=hb, ir_out=ib3)
        assert np.array_equal(np.asarray(hb), h) and np.array_equal(np.asarray(ib3)[0], ir)
        h2, ir2 = ctx.update(sc.sources, sc.listener, 16384, 8, 0x5EED, hist_out=np.zeros_like(h), ir_out=np.zeros((1,) + ir.shape, np.float32))
        assert np.array_equal(h2, h) and np.array_equal(ir2[0], ir)
        assert ctx.update(sc.sources, sc.listener, 16384, 8, 0x5EED) == (None, None) and np.array_equal(ctx.get_histogram(), h)
    iro = oracle.build_ir(oracle.default_config(), h[0], 16384)
    assert ir.shape == (2, 48000) and np.array_equal(ir[0], ir[1])
    assert np.abs(iro).max() > 0 and _rel(ir, iro) < TOL


def test_ir_from_energy_seam_and_edge_histograms(fs, oracle):
    """seam 1 of the reference taken literally: EnergyBuffer float[1000] -> ReconstructImpulseResponse"""
    rng = np.random.default_rng(2)
    cfg = oracle.default_config()
    with fs.Context() as ctx:
        for e in (np.zeros(1000, np.float32),
                  (rng.uniform(0, 0.05, 1000) * (rng.uniform(size=1000) < 0.3)).astype(np.float32),
                  np.where(np.arange(1000) == 999, 0.02, 0).astype(np.float32),
                  np.full(1000, 5e-7, np.float32)):
            ir = ctx.build_ir_from_energy(e)
            iro = oracle.build_ir_from_energy(cfg, e)
            if not iro.any():
                assert not ir.any()
            else:
                assert _rel(ir, iro) < TOL
        h = np.zeros((2, 8, 1000), np.uint64)                      # set_histogram path, 2 sources
        h[1, :, 100:400] = rng.integers(0, 2 ** 30, size=(8, 300)).astype(np.uint64)
        ctx.set_histogram(h, 1000)
        assert not ctx.build_ir(0).any()
        assert _rel(ctx.build_ir(1), oracle.build_ir(cfg, h[1], 1000)) < TOL
        assert np.array_equal(ctx.get_histogram(), h)


def test_ir_l2_normalisation_flag(fs, oracle):
    """FS_FLAG_IR_NORMALIZE (NormalizeImpulseResponse, COMP.cpp:382-406): every build path gives the oracle's unit-norm IR,
    and the convolver runs on it"""
    from frequensee import scenes, capi
    sc = scenes.shoebox()
    cfg = oracle.default_config(flags=oracle.FLAG_IR_NORMALIZE)
    srcs = np.array([[1.5, 1.2, 1.0], [3.0, 2.0, 1.5], [5.5, 4.0, 2.0]], np.float32)
    x = np.zeros((2, 1024, 2), np.float32); x[0, 0] = 1.0; x[1, 7] = -0.25
    with fs.Context(flags=capi.FLAG_IR_NORMALIZE) as ctx:
        ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
        h = ctx.trace(srcs, sc.listener, 8192, 8, 77)
        one = ctx.build_ir(1)
        allir = ctx.build_ir_all(3)
        bands = ctx.build_ir_bands(9, hist_source=2, conv_source=2)
        e = np.where((np.arange(1000) > 30) & (np.arange(1000) < 200), 0.01, 0).astype(np.float32)
        from_e = ctx.build_ir_from_energy(e, source=0)
        ctx.build_ir(1)
        ctx.conv_init_source(1)
        y = ctx.conv_process_many(x, 1)
    assert np.allclose(np.linalg.norm(one.astype(np.float64), axis=1), 1.0, rtol=1e-5)
    assert _rel(one, oracle.build_ir(cfg, h[1], 8192)) < TOL and np.array_equal(one, allir[1])
    for s in range(3):
        assert _rel(allir[s], oracle.build_ir(cfg, h[s], 8192)) < TOL
    assert _rel(bands, oracle.build_ir_bands(cfg, h[2], 8192, 9)) < TOL
    assert _rel(from_e, oracle.build_ir_from_energy(cfg, e)) < TOL
    cv = oracle.Conv(cfg); cv.set_ir(oracle.build_ir(cfg, h[1], 8192))
    yo = np.stack([cv.process(b) for b in x])
    assert _rel(y, yo) < TOL


def test_device_fft_vs_reference_kissfft_and_numpy(fs, oracle):
    rng = np.random.default_rng(0)
    with fs.Context() as ctx:
        for n in (64, 256, 2048, 4096):
            x = rng.uniform(-1, 1, n).astype(np.float32)
            X = ctx.rfft(x)
            assert _rel(X, np.fft.rfft(x.astype(np.float64))) < 1e-6
            if oracle.ref_lib() is not None:                        # the reference's own kiss_fftr
                assert _rel(X, oracle.kiss_fftr(x)) < TOL


def test_convolution_matches_direct_form(fs, oracle):
    """config 3 semantics: y_block = history (*) current IR (REV.cpp:172-213); unit impulse at frame 0
    plus white noise; IR refreshed mid-stream takes effect at the next block boundary."""
    rng = np.random.default_rng(3)
    cfg = oracle.default_config()
    def mk_ir(seed):
        r = np.random.default_rng(seed)
        ir = np.zeros((2, 48000), np.float32)
        ir[:, :30000] = (r.normal(size=(2, 30000)) * np.exp(-np.arange(30000) / 6000.0) * 0.01).astype(np.float32)
        ir[0, 47999] = 0.02
        return ir
    ir_a, ir_b = mk_ir(1), mk_ir(2)
    x = rng.uniform(-0.5, 0.5, size=(60, 1024, 2)).astype(np.float32)
    x[0, 0] = 1.0
    cv = oracle.Conv(cfg); cv.set_ir(ir_a)
    with fs.Context() as ctx:
        ctx.conv_init_source(0); ctx.set_ir(ir_a, 0)
        num = den = 0.0
        ys = []
        for b in range(60):
            if b == 52:
                cv.set_ir(ir_b); ctx.set_ir(ir_b, 0)
            y = ctx.conv_process(x[b], 0); yo = cv.process(x[b])
            ys.append(y)
            num += float(((y - yo) ** 2).sum()); den += float((yo ** 2).sum())
        assert den > 0 and (num / den) ** 0.5 < TOL
        # offline form == streaming form
        ctx.conv_init_source(1); ctx.set_ir(ir_a, 1)
        many = ctx.conv_process_many(x[:52], 1)
        assert np.array_equal(many, np.stack(ys[:52]))
        # release + re-init clears the history
        ctx.conv_release_source(1)
        with pytest.raises(fs.FrequenSeeError):
            ctx.conv_process(x[0], 1)
        with pytest.raises(fs.FrequenSeeError):
            ctx.conv_process(x[0, :512], 0)                          # wrong frame count


def test_clamp_and_wet_mix(fs, oracle):
    rng = np.random.default_rng(4)
    ir = np.zeros((2, 48000), np.float32); ir[:, 0] = 3.0           # gain 3 -> clamps
    x = rng.uniform(-0.9, 0.9, size=(2, 1024, 2)).astype(np.float32)
    for over in (dict(), dict(conv_clamp=0), dict(conv_wet=0.25)):
        cv = oracle.Conv(oracle.default_config(**over)); cv.set_ir(ir)
        with fs.Context(**over) as ctx:
            ctx.conv_init_source(0); ctx.set_ir(ir, 0)
            for b in range(2):
                assert np.allclose(ctx.conv_process(x[b], 0), cv.process(x[b]), rtol=1e-5, atol=2e-6)


def test_saved_ir_text_round_trip_drives_the_convolver(fs, oracle, tmp_path):
    """COMP.cpp:454-505: an IR saved as saved_ir.txt and loaded back gives bit-identical convolver output"""
    from frequensee import scenes
    sc = scenes.shoebox()
    rng = np.random.default_rng(2)
    x = rng.uniform(-0.5, 0.5, size=(3, 1024, 2)).astype(np.float32)
    with fs.Context() as ctx:
        ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
        ctx.trace(sc.sources, sc.listener, 4096, 8, 11)
        ir = ctx.build_ir(0)
        ctx.conv_init_source(0)
        y0 = ctx.conv_process_many(x, 0)
        p = str(tmp_path / "saved_ir.txt")
        fs.save_float_array(p, ir[0])
        mono = fs.load_float_array(p)
        assert np.array_equal(mono, ir[0])
        ctx.conv_init_source(1)
        ctx.set_ir(np.stack([mono, mono]), 1)
        y1 = ctx.conv_process_many(x, 1)
    assert np.array_equal(ir[0], ir[1])                     # both channels read the same mono histogram (COMP.cpp:325-329)
    assert np.array_equal(y0, y1)


def test_per_band_ir_matches_oracle(fs, oracle):
    """fs_build_ir_bands against the oracle on the histogram of a real trace; tolerance 1e-5 relative L2"""
    from frequensee import scenes
    sc = scenes.shoebox()
    with fs.Context() as ctx:
        ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
        h = ctx.trace(sc.sources, sc.listener, 16384, 8, 21)
        ir = ctx.build_ir_bands(99)
        ir_again = ctx.build_ir_bands(99)
        ir_other = ctx.build_ir_bands(100)
        ctx.conv_init_source(0)
        x = np.zeros((1, 1024, 2), np.float32); x[0, 0] = 1.0
        y = ctx.conv_process_many(x, 0)                                  # the convolver got the per-band IR (seed 100)
    iro = oracle.build_ir_bands(oracle.default_config(), h[0], 16384, 99)
    assert _rel(ir, iro) < 1e-5
    assert np.array_equal(ir, ir_again) and not np.array_equal(ir, ir_other)
    assert not np.array_equal(ir[0], ir[1])                               # decorrelated channels
    assert _rel(y[0, :, 0], np.clip(ir_other[0, :1024], -1, 1)) < 1e-5


def test_build_ir_all_equals_per_source_builds(fs, oracle):
    """fs_build_ir_all (one launch per kernel for all emitters) gives exactly the IRs and convolver state of fs_build_ir"""
    from frequensee import scenes
    sc = scenes.shoebox()
    srcs = np.array([[1.5, 1.2, 1.0], [3.0, 2.0, 1.5], [5.5, 4.0, 2.0], [2.0, 4.2, 0.8], [6.1, 1.1, 2.4]], np.float32)
    x = np.zeros((2, 1024, 2), np.float32); x[0, 0] = 1.0; x[1, 5] = -0.5
    with fs.Context() as ctx:
        ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
        h = ctx.trace(srcs, sc.listener, 4096, 8, 31)
        one = np.stack([ctx.build_ir(s) for s in range(len(srcs))])
        for s in range(len(srcs)):
            ctx.conv_init_source(s)
        y_one = np.stack([ctx.conv_process_many(x, s) for s in range(len(srcs))])
    with fs.Context() as ctx:
        ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
        ctx.trace(srcs, sc.listener, 4096, 8, 31)
        allir = ctx.build_ir_all(len(srcs))
        # a page-locked destination (fs_host_alloc) is written by the copy engine directly: same floats
        pinned = fs.capi.host_alloc((len(srcs), 2, 48000))
        pinned[:] = -1.0
        assert ctx.build_ir_all(len(srcs), out=pinned) is pinned
        for s in range(len(srcs)):
            ctx.conv_init_source(s)
        y_all = np.stack([ctx.conv_process_many(x, s) for s in range(len(srcs))])
    assert np.array_equal(one, allir) and np.array_equal(y_one, y_all)
    assert np.array_equal(np.asarray(pinned), allir)
    del pinned
    cfg = oracle.default_config()
    for s in range(len(srcs)):
        assert _rel(allir[s], oracle.build_ir(cfg, h[s], 4096)) < 1e-5
    assert not np.array_equal(allir[0], allir[1])
