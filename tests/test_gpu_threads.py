"""The threading contract of include/frequensee.h on the device: the game thread (fs_trace + fs_build_ir) and the audio
thread (fs_conv_process*) use ONE context concurrently (reference: SUB.cpp:55 vs REV.cpp:118, which share ImpulseBuffer
with no synchronisation).  Plus the multi-emitter callback and the API-robustness cases of the round-1 review."""
import threading
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_audio_thread_runs_beside_the_game_thread(fs, oracle):
    """One thread loops fs_trace + fs_build_ir (a 2^20-pair room update, ~5 ms each), another loops fs_conv_process.
    Every update re-traces the same seed, so the IR never changes and the audio must equal the single-threaded answer
    exactly, whenever the swaps happen; the callbacks must not queue behind the traces."""
    from frequensee import scenes
    sc = scenes.furnished_room()
    rng = np.random.default_rng(4)
    n_blocks = 96
    x = rng.uniform(-0.5, 0.5, size=(n_blocks, 1024, 2)).astype(np.float32)
    with fs.Context(conv_clamp=0) as ctx:
        ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
        ctx.trace(sc.sources, sc.listener, 1 << 20, 16, 7, want_hist=False)
        ir = ctx.build_ir(0)                                        # synchronises: in effect for every later callback
        assert np.abs(ir).max() > 0
        ctx.conv_init_source(0)
        expect = ctx.conv_process_many(x, 0)                        # single-threaded answer
        ctx.conv_init_source(0)                                     # history back to zero (OnInitSource)
        stop = threading.Event()
        errors, updates, trace_ms = [], [0], []

        def game():
            try:
                while not stop.is_set():
                    t0 = time.perf_counter()
                    ctx.trace(sc.sources, sc.listener, 1 << 20, 16, 7, want_hist=False)
                    ctx.build_ir(0, want_ir=(updates[0] % 3 == 0))  # both the asynchronous and the synchronising form
                    ctx.synchronize()
                    trace_ms.append(1e3 * (time.perf_counter() - t0))
                    updates[0] += 1
            except Exception as e:                                  # pragma: no cover
                errors.append(e)

        th = threading.Thread(target=game)
        th.start()
        lat, got = [], np.zeros_like(x)
        try:
            while updates[0] < 1:
                time.sleep(0.001)
            for b in range(n_blocks):
                t0 = time.perf_counter()
                got[b] = ctx.conv_process(x[b], 0)
                lat.append(1e3 * (time.perf_counter() - t0))
                time.sleep(0.002)                                   # callbacks arrive while traces are in flight
        finally:
            stop.set()
            th.join()
        assert not errors, errors
        assert updates[0] >= 5
        assert np.array_equal(got, expect)                          # same IR every time: bit-identical audio
        lat = np.sort(np.array(lat))
        p50, p99 = lat[len(lat) // 2], lat[int(len(lat) * 0.99)]
        t_update = float(np.median(trace_ms))
        print("callback latency beside %d updates of %.2f ms: p50 %.3f ms, p99 %.3f ms" % (updates[0], t_update, p50, p99))
        assert t_update > 2.0                                       # the traces really are long compared with a callback
        assert p99 < 0.5 * t_update and p50 < 0.5                   # not queued behind them (21.3 ms budget per callback)


def test_ir_update_is_adopted_at_a_block_boundary_after_sync(fs, oracle):
    """a new IR takes effect for callbacks issued after a synchronising call; the history (FDL) carries over"""
    rng = np.random.default_rng(9)
    ir_a = (rng.normal(size=(2, 48000)) * np.exp(-np.arange(48000) / 5000.0) * 0.02).astype(np.float32)
    ir_b = (rng.normal(size=(2, 48000)) * np.exp(-np.arange(48000) / 9000.0) * 0.02).astype(np.float32)
    x = rng.uniform(-0.5, 0.5, size=(6, 1024, 2)).astype(np.float32)
    cfg = oracle.default_config(conv_clamp=0)
    cv = oracle.Conv(cfg)
    with fs.Context(conv_clamp=0) as ctx:
        ctx.conv_init_source(0)
        ctx.set_ir(ir_a, 0); cv.set_ir(ir_a)
        for b in range(6):
            if b == 3:
                ctx.set_ir(ir_b, 0); cv.set_ir(ir_b)
            assert _rel(ctx.conv_process(x[b], 0), cv.process(x[b])) < 1e-5


def test_multi_emitter_callback_equals_serial_calls(fs):
    """fs_conv_process_multi: 64 sources in one launch (grid = sources x channels) == 64 fs_conv_process calls"""
    rng = np.random.default_rng(12)
    S = 64
    irs = (rng.normal(size=(S, 2, 48000)) * np.exp(-np.arange(48000) / 7000.0) * 0.02).astype(np.float32)
    x = rng.uniform(-0.5, 0.5, size=(3, S, 1024, 2)).astype(np.float32)
    ids = np.arange(S, dtype=np.uint32)[::-1].copy() + 5             # any distinct ids, any order
    with fs.Context() as a, fs.Context() as b:
        for i, s in enumerate(ids):
            for c in (a, b):
                c.conv_init_source(int(s)); c.set_ir(irs[i], int(s))
        for blk in range(3):
            y_multi = a.conv_process_multi(x[blk], ids)
            y_serial = np.stack([b.conv_process(x[blk, i], int(s)) for i, s in enumerate(ids)])
            assert np.array_equal(y_multi, y_serial)
        with pytest.raises(fs.FrequenSeeError):
            a.conv_process_multi(x[0, :2], np.array([5, 5], np.uint32))          # duplicate id
        with pytest.raises(fs.FrequenSeeError):
            a.conv_process_multi(x[0, :2], np.array([5, 4000], np.uint32))       # not initialised


def test_radix4_fft_sizes_and_block_sizes(fs, oracle):
    """every conv_block the ABI allows (2 Bk-point FFT: odd and even log2) against the direct form; fs_build_ir_all with
    conv_block = 2048 (64 KB of dynamic shared memory in the multi-source spectra kernel)"""
    rng = np.random.default_rng(3)
    for bk in (32, 64, 256, 512, 2048):
        ir = (rng.normal(size=(2, 48000)) * np.exp(-np.arange(48000) / 3000.0) * 0.02).astype(np.float32)
        x = rng.uniform(-0.5, 0.5, size=(3, bk, 2)).astype(np.float32)
        cfg = oracle.default_config(conv_block=bk, conv_clamp=0)
        cv = oracle.Conv(cfg); cv.set_ir(ir)
        with fs.Context(conv_block=bk, conv_clamp=0) as ctx:
            ctx.conv_init_source(0); ctx.set_ir(ir, 0)
            for b in range(3):
                assert _rel(ctx.conv_process(x[b], 0), cv.process(x[b])) < 1e-5, bk
            if bk == 2048:
                h = np.zeros((3, 8, 1000), np.uint64)
                h[:, :, 50:300] = rng.integers(0, 2 ** 30, size=(3, 8, 250)).astype(np.uint64)
                ctx.set_histogram(h, 1000)
                irs = ctx.build_ir_all(3)
                for s in range(3):
                    assert _rel(irs[s], oracle.build_ir(cfg, h[s], 1000)) < 1e-5


def test_histogram_readback_uses_the_source_count_of_the_last_trace(fs, oracle):
    """round-1 review: fs_get_histogram copied the allocation's high-water mark (4 sources) into a 1-source buffer"""
    from frequensee import scenes
    sc = scenes.shoebox()
    src4 = np.array([[1.5, 1.2, 1.0], [3.0, 2.0, 1.5], [6.0, 4.0, 2.0], [2.0, 4.0, 0.8]], np.float32)
    with fs.Context() as ctx:
        ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
        h4 = ctx.trace(src4, sc.listener, 512, 8, 1)
        assert ctx.get_histogram().shape[0] == 4 and np.array_equal(ctx.get_histogram(), h4)
        ctx.trace(src4[:1], sc.listener, 512, 8, 1, want_hist=False)
        h1 = ctx.get_histogram()
        assert h1.shape == (1, 8, 1000) and np.array_equal(h1[0], h4[0])
        with pytest.raises(fs.FrequenSeeError):
            ctx.build_ir(2)                                          # sources 1..3 of the older trace are gone


def test_source_id_limits_and_lowpass_validation(fs, oracle):
    from frequensee import capi
    ir = np.zeros((2, 48000), np.float32)
    with fs.Context() as ctx:
        for bad in (4096, 1 << 31):
            with pytest.raises(fs.FrequenSeeError) as ei:
                ctx.set_ir(ir, bad)
            assert ei.value.code == capi.FS_ERR_INVALID
            with pytest.raises(fs.FrequenSeeError):
                ctx.build_ir_from_energy(np.zeros(1000, np.float32), bad)
            with pytest.raises(fs.FrequenSeeError):
                ctx.conv_init_source(bad)
        ctx.set_ir(ir, 4095)
    for bad in (0.0, -0.1, 1.5, 1e-4):
        with pytest.raises(fs.FrequenSeeError):
            fs.Context(ir_lowpass=bad)
    with pytest.raises(fs.FrequenSeeError):
        fs.Context(bin_ms=1e-3)                                      # less than one sample per bin
    rng = np.random.default_rng(8)
    e = (rng.uniform(0, 0.05, 1000) * (rng.uniform(size=1000) < 0.3)).astype(np.float32)
    for a in (0.02, 0.25, 0.9, 1.0):                                 # the warm-up window follows the coefficient
        with fs.Context(ir_lowpass=a) as ctx:
            assert _rel(ctx.build_ir_from_energy(e), oracle.build_ir_from_energy(oracle.default_config(ir_lowpass=a), e)) < 1e-5


def test_last_error_is_per_thread(fs):
    """fs_last_error returns the message of the calling thread's last failure"""
    with fs.Context() as ctx:
        msgs = {}

        def worker():
            try:
                ctx.conv_process(np.zeros((1024, 2), np.float32), 77)
            except fs.FrequenSeeError as e:
                msgs["audio"] = str(e)

        try:
            ctx.build_ir(0)
        except fs.FrequenSeeError as e:
            msgs["game"] = str(e)
        t = threading.Thread(target=worker); t.start(); t.join()
        assert "not initialised" in msgs["audio"] and "no histogram" in msgs["game"]
        assert b"no histogram" in ctx.L.fs_last_error(ctx.h)         # this thread still sees its own message
