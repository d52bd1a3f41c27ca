"""The CUDA path against the REFERENCE'S OWN CODE: one UpdateSource (SUB.cpp:128-195, unmodified, 1000 path pairs) computed by
oracle/_ref/libref_ue_bodies.so -- or, where oracle/_ref did not travel, its committed output tests/golden/ref_ue_v1.npz --
against fs_trace + fs_build_ir through the C-ABI, configured with the reference's constants (pyoracle.ue_pin_config)."""
import os

import numpy as np
import pytest

from test_oracle_ref_ue import pin_scene, closet_scene

pytestmark = pytest.mark.gpu


def _reference(oracle, scene_fn, seed, tag):
    verts, tri_mat, ab, src, lis = scene_fn()
    if oracle.ref_ue_lib() is not None:
        W = oracle.RefUEWorld(verts, tri_mat, 1.0 - ab[:, 2])
        e, ir, n = W.update_source(src, lis, seed)
        return e, ir[0], n
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_ue_v1.npz"))
    return g[tag + "_energy"], (g[tag + "_ir0"] if tag + "_ir0" in g else None), int(g[tag + "_traces"])


def _pin_kwargs(oracle, **over):
    c = oracle.ue_pin_config(**over)
    return dict(n_bands=1, n_bins=1000, bin_ms=c.bin_ms, rr_prob=c.rr_prob, eps_offset=c.eps_offset, eps_connect=c.eps_connect,
                min_seg=c.min_seg, sound_speed=c.sound_speed, pdf_exponent=c.pdf_exponent, energy_clamp=c.energy_clamp,
                energy_gain=c.energy_gain, air_absorption=list(c.air_absorption))


@pytest.mark.parametrize("scene_fn, seed, tag", [(pin_scene, 1, "pin_seed1"), (closet_scene, 5, "closet_seed5")])
def test_gpu_update_equals_reference_update_source(fs, oracle, scene_fn, seed, tag):
    verts, tri_mat, ab, src, lis = scene_fn()
    e_ref, ir_ref, n_traces = _reference(oracle, scene_fn, seed, tag)
    ab1 = ab[:, 2:3].copy()
    with fs.Context(**_pin_kwargs(oracle)) as ctx:
        ctx.set_scene(verts, tri_mat, ab1)
        h = ctx.trace([src], lis, 1000, 512, seed)
        st = ctx.stats()
    # the same rays as the reference traced, and the oracle's integers bit for bit
    assert st["ext_rays"] + st["shadow_rays"] == n_traces
    S = oracle.Scene(verts, tri_mat, ab1, use_bvh=False)
    ho, so = S.trace(oracle.ue_pin_config(), [src], lis, 1000, 512, seed)
    assert np.array_equal(h, ho)
    # the reference's float EnergyBuffer (1/1000 per path) against the GPU's Q32.32 histogram
    e_gpu = h[0, 0].astype(np.float64) / 2.0 ** 32 / 1000
    assert e_ref.sum() > 0 and np.abs(e_ref - e_gpu).sum() / e_gpu.sum() < 2e-5
    # the IR as the reference computes it (49 samples per bin): a context with that bin width rebuilds it from the histogram
    if ir_ref is not None:
        with fs.Context(**_pin_kwargs(oracle, bin_ms=49.0 / 48.0)) as c2:
            c2.set_histogram(h, 1000)
            ir = c2.build_ir(0)
        if ir_ref.any():
            assert np.linalg.norm(ir[0] - ir_ref) / np.linalg.norm(ir_ref) < 1e-5
            assert np.array_equal(ir[0], ir[1])
        else:
            assert not ir.any()
