"""SURVEY 8f rank 1, second half -- FS_FLAG_MIS: all prefix connections combined with the balance heuristic (the goal of the
reference's unfinished getPdf / getExpectedWeight / MISEnergy, SUB.cpp:537-597).  CPU: the oracle's three estimators of the
same integral -- balance-weighted all strategies, listener-connection only (t = 1), source-connection only (s = 1) -- agree
within their confidence intervals, which pins every area-measure density.  GPU: bit-exact parity."""
import numpy as np
import pytest


def _cfg(oracle, extra=0, **over):
    kw = dict(energy_clamp=1e30, energy_gain=1.0, min_seg=1e-3)
    kw.update(over)
    return oracle.default_config(flags=oracle.FLAG_MIS | extra, **kw)


def test_balance_heuristic_equals_single_strategy_estimators(oracle):
    from frequensee import scenes
    sc = scenes.shoebox()
    ab = np.full((8, 8), 0.5, np.float32)                  # rho = 0.5: paths beyond the depth cut carry < 1e-6 of the energy
    ab[:, 4:] = 0.7
    S = oracle.Scene(sc.verts, sc.tri_mat, ab, use_bvh=False)
    N, D, seeds = 12000, 20, range(5)
    est = {}
    for name, extra in (("balance", 0), ("t1", oracle.FLAG_MIS_T1), ("s1", oracle.FLAG_MIS_S1)):
        tot, coarse = [], []
        for seed in seeds:
            h, st = S.trace(_cfg(oracle, extra), sc.sources, sc.listener, N, D, seed, n_threads=8)
            e = h[0].astype(np.float64) / 2.0 ** 32 / N
            tot.append(e.sum(axis=1)); coarse.append(e[0].reshape(50, 20).sum(axis=1))
        est[name] = (np.array(tot), np.array(coarse))
    mean = {k: v[0].mean(axis=0) for k, v in est.items()}
    sem = {k: v[0].std(axis=0, ddof=1) / np.sqrt(len(seeds)) for k, v in est.items()}
    for other in ("t1", "s1"):
        z = np.abs(mean["balance"] - mean[other]) / np.sqrt(sem["balance"] ** 2 + sem[other] ** 2)
        assert z.max() < 4.5, (other, z, mean["balance"], mean[other])
        assert np.abs(mean["balance"] / mean[other] - 1.0).max() < 0.02
    # the first 20 ms bins of band 0 (direct sound + first reflections) agree too
    cb, ct = est["balance"][1].mean(axis=0), est["t1"][1].mean(axis=0)
    assert np.allclose(cb[:6], ct[:6], rtol=0.05, atol=1e-6)
    # the direct path: strategy (1, 1) alone builds it, weight 1, value 1 / (4 pi d^2) exp(-air d)
    d = float(np.linalg.norm(np.asarray(sc.listener, np.float64) - np.asarray(sc.sources[0], np.float64)))
    h, _ = S.trace(_cfg(oracle), sc.sources, sc.listener, 64, 0, 1)
    e = h[0, 0].astype(np.float64) / 2.0 ** 32 / 64
    assert list(np.flatnonzero(e)) == [int(d / 343.0 * 1000)]
    assert abs(e.sum() / (np.exp(-1e-4 * d) / (4 * np.pi * d * d)) - 1.0) < 1e-5


@pytest.mark.gpu
def test_gpu_mis_bit_exact(fs, oracle):
    from frequensee import scenes, capi
    room = scenes.furnished_room(target_tris=30000)
    shoebox = scenes.shoebox()
    for sc, use_bvh, n, depth, over in ((shoebox, False, 3000, 8, {}), (room, True, 1200, 16, {}), (shoebox, False, 500, 32, {}),
                                        (shoebox, False, 2000, 6, dict(energy_clamp=1e30, energy_gain=1.0, min_seg=1e-3, rr_prob=1.0)),
                                        (shoebox, False, 700, 0, {})):
        S = oracle.Scene(sc.verts, sc.tri_mat, sc.absorption, use_bvh=use_bvh)
        ho, so = S.trace(oracle.default_config(flags=oracle.FLAG_MIS, **over), sc.sources[:1], sc.listener, n, depth, 13, n_threads=16)
        for ctx_over in ({}, {"max_batch_paths": 400}):
            with fs.Context(flags=capi.FLAG_MIS, **dict(over, **ctx_over)) as ctx:
                ctx.set_scene(sc.verts, sc.tri_mat, sc.absorption)
                h = ctx.trace(sc.sources[:1], sc.listener, n, depth, 13)
                st = ctx.stats()
            assert ho.any() and np.array_equal(h, ho), (n, depth, over, ctx_over)
            assert [st["ext_rays"], st["shadow_rays"], st["connected"]] == [so["ext_rays"], so["shadow_rays"], so["connected"]]
    with fs.Context(flags=capi.FLAG_MIS) as ctx:
        ctx.set_scene(shoebox.verts, shoebox.tri_mat, shoebox.absorption)
        with pytest.raises(fs.FrequenSeeError):
            ctx.trace(shoebox.sources, shoebox.listener, 64, 33, 1)
