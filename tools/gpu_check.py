"""Quick on-GPU diagnosis: device BVH vs device brute force vs CPU oracle, first timings.
Run under gpurun; writes gpurun_out/gpu_check.log"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-pathtracer_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import frequensee as fs
from frequensee import scenes, capi
import pyoracle as po

def cmp_hist(a, b, tag):
    eq = np.array_equal(a, b)
    print(f"[{tag}] hist equal: {eq}  sum gpu={int(a.sum())} cpu={int(b.sum())} nnz gpu={np.count_nonzero(a)} cpu={np.count_nonzero(b)}", flush=True)
    return eq

def locate(ctx, S, cfgo, sc, n, D, seed):
    dg = ctx.trace_debug(sc.sources, sc.listener, n, 0, n * len(sc.sources), D, seed)
    _, _, do = S.trace(cfgo, sc.sources, sc.listener, n, D, seed, debug=True)
    bad = 0
    for i in range(len(dg)):
        a, b = dg[i], do[i]
        same = all(np.array_equal(a[k], b[k]) for k in dg.dtype.names)
        if not same:
            bad += 1
            if bad <= 5:
                print("  path", i, "GPU", a, "\n          CPU", b, flush=True)
    print("  mismatching paths:", bad, "of", len(dg), flush=True)

def main():
    rng = np.random.default_rng(0)
    out = {}
    # 1. shoebox parity
    sc = scenes.shoebox()
    cfgo = po.default_config()
    S = po.Scene(sc.verts, sc.tri_mat, sc.absorption, use_bvh=False)
    for flags, tag in ((0, "shoebox/split"), (capi.FLAG_FUSED_EXTEND, "shoebox/fused"), (capi.FLAG_BRUTE_FORCE, "shoebox/brute"),
                       (capi.FLAG_NO_SPLAT_AGG, "shoebox/noagg"), (capi.FLAG_COUNT_VISITS, "shoebox/count")):
        ctx = fs.Context(flags=flags)
        ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
        hg = ctx.trace(sc.sources, sc.listener, 16384, 8, 0x5EED)
        ho, so = S.trace(cfgo, sc.sources, sc.listener, 16384, 8, 0x5EED, n_threads=8)
        ok = cmp_hist(hg, ho, tag)
        st = ctx.stats(); print("   gpu stats", st, "\n   cpu stats", so, flush=True)
        if not ok: locate(ctx, S, cfgo, sc, 2048, 8, 0x5EED)
        out[tag] = ok
        ctx.close()
    # 2. furnished room: intersector parity + hist parity
    fr = scenes.furnished_room()
    S2 = po.Scene(fr.verts, fr.tri_mat, fr.absorption, use_bvh=True)
    ctx = fs.Context(flags=capi.FLAG_COUNT_VISITS)
    t0 = time.time(); ctx.set_scene(fr.verts, fr.tri_mat, fr.absorption); print("commit %d tris: %.3fs" % (fr.n_tris, time.time() - t0), ctx.stats(), flush=True)
    n = 200000
    o = rng.uniform([0.5, 0.5, 0.3], [11.5, 8.5, 3.3], size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays = np.concatenate([o, d], axis=1)
    t, tri = ctx.closest_hits(rays)
    ctxb = fs.Context(flags=capi.FLAG_BRUTE_FORCE); ctxb.set_scene(fr.verts, fr.tri_mat, fr.absorption)
    nb = 20000
    tb, trib = ctxb.closest_hits(rays[:nb])
    print("gpu bvh vs gpu brute: t equal", np.array_equal(t[:nb], tb), "tri equal", np.array_equal(tri[:nb], trib), flush=True)
    bad = np.nonzero((t[:nb] != tb) | (tri[:nb] != trib))[0]
    for i in bad[:5]: print("   ray", i, rays[i], "bvh", t[i], tri[i], "brute", tb[i], trib[i])
    nc = 20000; mism = 0
    for i in range(nc):
        h, tt, ti = S2.closest_hit(rays[i, :3], rays[i, 3:])
        if h != (tri[i] != 0xffffffff) or (h and (np.float32(tt) != t[i] or ti != tri[i])):
            mism += 1
            if mism <= 5: print("   ray", i, "gpu", t[i], tri[i], "oracle", h, tt, ti)
    print("gpu bvh vs oracle closest: mismatches", mism, "of", nc, flush=True)
    out["fr/intersect"] = (len(bad) == 0 and mism == 0)
    cfgo = po.default_config()
    N = 1 << 14
    hg = ctx.trace(fr.sources, fr.listener, N, 16, 1)
    ho, so = S2.trace(cfgo, fr.sources, fr.listener, N, 16, 1, n_threads=16)
    ok = cmp_hist(hg, ho, "furnished/16k")
    st = ctx.stats(); print("   gpu stats", st, "\n   cpu stats", so, flush=True)
    if not ok: locate(ctx, S2, cfgo, fr, 4096, 16, 1)
    out["fr/hist"] = ok
    irg = ctx.build_ir(0); iro = po.build_ir(cfgo, ho[0], N)
    rel = np.linalg.norm(irg - iro) / max(np.linalg.norm(iro), 1e-30)
    print("IR rel-L2", rel, flush=True); out["fr/ir_rel"] = float(rel)
    ctx.close(); ctxb.close()
    # 3. timing, 1M paths
    for flags, tag in ((0, "split"), (capi.FLAG_FUSED_EXTEND, "fused"), (capi.FLAG_FUSED_EXTEND | capi.FLAG_SMEM_TREELET, "fused+top"), (capi.FLAG_NO_SPLAT_AGG, "split/noagg")):
        ctx = fs.Context(flags=flags | capi.FLAG_TIME_KERNELS)
        ctx.set_scene(fr.verts, fr.tri_mat, fr.absorption)
        N = 1 << 20
        for it in range(3):
            ctx.trace(fr.sources, fr.listener, N, 16, 100 + it, want_hist=False); ctx.build_ir(0, want_ir=False)
            st = ctx.stats()
        rays = st["ext_rays"] + st["shadow_rays"]
        print(f"[time {tag}] ext {st['extend_ms']:.3f} con {st['connect_ms']:.3f} eval {st['eval_ms']:.3f} | 1M paths: trace {st['last_trace_ms']:.3f} ms  paths/s {N / st['last_trace_ms'] * 1e3:.3e}  Mrays/s {rays / st['last_trace_ms'] / 1e3:.1f}  connected {st['connected']}", flush=True)
        out["time/" + tag] = st["last_trace_ms"]
        ctx.close()
    # 4. conv + fft
    ctx = fs.Context()
    x = rng.uniform(-1, 1, 2048).astype(np.float32)
    Xg = ctx.rfft(x); Xn = np.fft.rfft(x.astype(np.float64))
    print("rfft rel err vs numpy", np.linalg.norm(Xg - Xn) / np.linalg.norm(Xn), flush=True)
    cfgo = po.default_config()
    ir = np.zeros((2, 48000), np.float32); ir[:, :4000] = (rng.normal(size=(2, 4000)) * np.exp(-np.arange(4000) / 800.0)).astype(np.float32) * 0.05
    ctx.conv_init_source(0); ctx.set_ir(ir, 0)
    cv = po.Conv(cfgo); cv.set_ir(ir)
    blocks = rng.uniform(-0.5, 0.5, size=(6, 1024, 2)).astype(np.float32)
    yg = ctx.conv_process_many(blocks, 0)
    yo = np.stack([cv.process(b) for b in blocks])
    rel = np.linalg.norm(yg - yo) / np.linalg.norm(yo)
    print("conv rel-L2 vs direct double", rel, flush=True); out["conv_rel"] = float(rel)
    ctx.close()
    print(json.dumps(out))

if __name__ == "__main__":
    main()
