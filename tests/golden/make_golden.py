"""Generates the committed golden vectors under tests/golden/ from the CPU oracle.

The reference (henreedev/audio-pathtracer) has no tests, fixtures or golden vectors for this path
and its BDPT code cannot be built or imported here (Unreal Engine 5.4 C++),
so the vectors are produced by oracle/fs_oracle.c -- whose primitives are pinned separately by
the analytic known-answer tests in tests/test_oracle_kat.py and, for the FFT stage, by the
reference's own KissFFT (oracle/_ref).  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "audio-pathtracer_b200"))
import pyoracle as po  # noqa: E402
from frequensee import scenes  # noqa: E402


def sparse(h):
    idx = np.flatnonzero(h)
    return idx.astype(np.uint32), h.reshape(-1)[idx]


def main():
    out = {}
    # config 1: shoebox, 16k paths, depth 8, seed 0x5EED, RR 0.9 and RR disabled
    sc = scenes.shoebox()
    S = po.Scene(sc.verts, sc.tri_mat, sc.absorption, use_bvh=False)
    for tag, rr in (("rr09", 0.9), ("rr10", 1.0)):
        cfg = po.default_config(rr_prob=rr)
        h, st, dbg = S.trace(cfg, sc.sources, sc.listener, 16384, 8, 0x5EED, debug=True)
        i, v = sparse(h)
        out["shoebox_%s_idx" % tag] = i
        out["shoebox_%s_val" % tag] = v
        out["shoebox_%s_stats" % tag] = np.array([st["ext_rays"], st["shadow_rays"], st["connected"]], dtype=np.uint64)
        out["shoebox_%s_dbg64" % tag] = dbg[:64]
        if rr == 0.9:
            ir = po.build_ir(cfg, h[0], 16384)
            out["shoebox_ir_first4800"] = ir[0, :4800]
    # furnished room (config 2 geometry) at a size the oracle finishes in seconds
    fr = scenes.furnished_room()
    S2 = po.Scene(fr.verts, fr.tri_mat, fr.absorption, use_bvh=True)
    cfg = po.default_config()
    h, st = S2.trace(cfg, fr.sources, fr.listener, 8192, 16, 1, n_threads=8)
    i, v = sparse(h)
    out["room_idx"], out["room_val"] = i, v
    out["room_stats"] = np.array([st["ext_rays"], st["shadow_rays"], st["connected"], fr.n_tris], dtype=np.uint64)
    # primitives
    out["philox_kat"] = np.array([po.philox([0, 0, 0, 0], [0, 0]),
                                  po.philox([0xffffffff] * 4, [0xffffffff] * 2),
                                  po.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                                            [0xa4093822, 0x299f31d0])], dtype=np.uint32)
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote golden_v1.npz:", {k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == "__main__":
    main()
