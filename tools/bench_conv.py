"""BASELINE.json configs[2]: IR build + uniformly partitioned FFT convolution.

IR of the furnished room (config 2), 10 s x 48 kHz stereo dry signal (seeded noise in [-0.5, 0.5] + a unit impulse
at frame 0), 1024-frame callbacks, a new IR published at every block boundary >= k * 800 frames (60 Hz refresh).
Prints one JSON line:
  ms_per_ir_refresh      fs_build_ir: histogram -> IR -> 47 x 2 partition spectra (CUDA events, device resident)
  ms_per_block_callback  fs_conv_process, host buffers in and out, one 1024-frame block per call (wall clock)
  ms_per_block_stream    fs_conv_process_many over all 469 blocks (wall clock / blocks)
  rel_l2_vs_direct       against the double-precision direct-form convolution with the IR held fixed (tolerance 1e-5)
  multi_emitter          fs_conv_process_multi: one callback of 64 sources in ONE launch (grid = sources x channels), device time
                         (CUDA events on the convolver's stream) and wall clock, with the `roofline` block of SURVEY.md 8(d):
                         algorithmic bytes = S * C * [(P+1)(Bk+1) * 16 + 2 Bk * 8] per callback against the measured HBM peak
  cpu_reference          the reference's own scheme (3 x 65 536-point KissFFT per channel per callback, REV.cpp:172-213)
                         from oracle/_ref on one host thread, when that library was built
The oracle is used here only as the checker / CPU baseline (tools/, not the product)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-pathtracer_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import frequensee as fs
from frequensee import scenes


def main():
    sc = scenes.furnished_room()
    src = sc.sources[:1]
    FR, BLK, NB = 480000, 1024, 469
    rng = np.random.default_rng(3)
    x = rng.uniform(-0.5, 0.5, size=(NB * BLK, 2)).astype(np.float32)
    x[FR:] = 0.0
    x[0] = 1.0
    out = {"config": "furnished room IR (2^20 pairs, depth 16), 10 s 48 kHz stereo, 1024-frame blocks, IR refresh at 60 Hz"}
    with fs.Context(conv_clamp=0) as ctx:
        ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
        ctx.conv_init_source(0)
        # --- IR refresh
        ctx.trace(src, sc.listener, 1 << 20, 16, 1000, want_hist=False)
        ms = []
        for i in range(20):
            ctx.build_ir(0)                                         # last_ir_ms: CUDA events around histogram -> IR -> spectra
            ms.append(ctx.stats()["last_ir_ms"])
        out["ms_per_ir_refresh"] = float(np.median(ms[2:]))
        ir = ctx.build_ir(0)
        # --- per-callback latency with a refresh every 800 frames of audio time
        lis = np.array(sc.listener, np.float32)
        t_cb, next_refresh, refreshes = [], 0, 0
        ys = np.zeros_like(x)
        t_trace = []
        for b in range(NB):
            if b * BLK >= next_refresh and refreshes < 8:          # a few real re-traces (seeded listener perturbation)
                t0 = time.perf_counter()
                l2 = lis + np.float32(0.01) * np.float32(refreshes) * np.array([1, 0, 0], np.float32)
                ctx.trace(src, l2, 1 << 20, 16, 2000 + refreshes, want_hist=False)
                ctx.build_ir(0, want_ir=False)
                ctx.stats()                                         # synchronises the context stream
                t_trace.append(1e3 * (time.perf_counter() - t0))
                refreshes += 1
            while next_refresh <= b * BLK:
                next_refresh += 800
            t0 = time.perf_counter()
            ys[b * BLK:(b + 1) * BLK] = ctx.conv_process(x[b * BLK:(b + 1) * BLK], 0)
            t_cb.append(1e3 * (time.perf_counter() - t0))
        out["ms_per_block_callback"] = float(np.median(t_cb))
        out["ms_per_block_callback_p99"] = float(np.percentile(t_cb, 99))
        out["ms_per_trace_plus_ir_refresh_wall"] = float(np.median(t_trace))
        # --- streaming, fixed IR, checked against the direct form
        ctx.conv_release_source(0); ctx.conv_init_source(0); ctx.set_ir(ir, 0)
        ctx.conv_process_many(x[:BLK * 4].reshape(4, BLK, 2), 0)
        ctx.conv_release_source(0); ctx.conv_init_source(0); ctx.set_ir(ir, 0)
        t0 = time.perf_counter()
        y = ctx.conv_process_many(x.reshape(NB, BLK, 2), 0).reshape(-1, 2)
        out["ms_per_block_stream"] = 1e3 * (time.perf_counter() - t0) / NB
        out["ms_per_block_device"] = ctx.stats()["last_conv_ms"] / NB
    # --- multi-emitter callback (BASELINE configs[3]: 64 sources), every source with its own IR and history
    S = 64
    with fs.Context(conv_clamp=0) as ctx:
        c = ctx.cfg
        P = (c.sample_rate + c.conv_block - 1) // c.conv_block
        irs = (rng.normal(size=(S, 2, 48000)) * np.exp(-np.arange(48000) / 7000.0) * 0.02).astype(np.float32)
        ids = np.arange(S, dtype=np.uint32)
        for i in range(S):
            ctx.conv_init_source(i); ctx.set_ir(irs[i], i)
        xb = rng.uniform(-0.5, 0.5, size=(S, BLK, 2)).astype(np.float32)
        dev, wall = [], []
        for it in range(60):
            t0 = time.perf_counter()
            ctx.conv_process_multi(xb, ids)
            wall.append(1e3 * (time.perf_counter() - t0))
            dev.append(ctx.stats()["last_conv_ms"])
        t_serial = []
        for it in range(5):
            t0 = time.perf_counter()
            for i in range(S):
                ctx.conv_process(xb[i], i)
            t_serial.append(1e3 * (time.perf_counter() - t0))
        bytes_cb = S * c.n_channels * ((P + 1) * (BLK + 1) * 16 + 2 * BLK * 8)
        peak = 6545.9
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        d = float(np.median(dev[10:]))
        out["multi_emitter"] = {"sources": S, "ms_per_callback_device": d, "ms_per_callback_wall": float(np.median(wall[10:])),
                                "ms_per_callback_wall_p99": float(np.percentile(wall[10:], 99)),
                                "ms_64_serial_calls_wall": float(np.median(t_serial)),
                                "roofline": {"bound": "hbm", "kernel": "k_conv_blocks (grid = 64 sources x 2 channels, radix-4 Stockham FFT, 47-partition spectral MAC)",
                                             "achieved": bytes_cb / (d * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                             "frac": bytes_cb / (d * 1e-3) / 1e9 / peak, "bytes_per_launch": bytes_cb,
                                             "note": "64 x (FDL + H) = 99 MB per callback: partly L2-resident between callbacks"}}
    num = den = 0.0
    n_chk = 96000                                                    # first 2 s: direct form in double is O(n * taps)
    for c in range(2):
        yd = np.convolve(x[:n_chk, c].astype(np.float64), ir[c].astype(np.float64))[:n_chk]
        num += float(((y[:n_chk, c] - yd) ** 2).sum()); den += float((yd ** 2).sum())
    out["rel_l2_vs_direct"] = (num / den) ** 0.5
    try:
        import pyoracle as po
        if po.ref_lib() is not None:
            rc = po.RefKissConv(); rc.set_ir(ir)
            t = []
            for b in range(24):
                t0 = time.perf_counter(); rc.process(x[b * BLK:(b + 1) * BLK]); t.append(1e3 * (time.perf_counter() - t0))
            out["cpu_reference"] = {"ms_per_block_callback": float(np.median(t)), "cores": 1, "kind": "reference",
                                    "what": "reference's vendored KissFFT, 65 536-point scheme of REV.cpp:172-213 (oracle/_ref)"}
    except Exception as e:                                            # oracle/_ref is optional on the GPU box
        out["cpu_reference"] = {"unavailable": str(e)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
