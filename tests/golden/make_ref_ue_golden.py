"""Generates tests/golden/ref_ue_v1.npz from the REFERENCE's own code (oracle/_ref/libref_ue_bodies.so: the unmodified
UpdateSource of SUB.cpp:128-195 compiled against oracle/ue_shim).  Run here, where /root/reference exists:

    python tests/golden/make_ref_ue_golden.py

Contents: EnergyBuffer (float[1000]) and line-trace count of one UpdateSource on the pin scene (seed 1), and EnergyBuffer +
ImpulseBuffer[0] on the closet scene (seed 5).  tests/test_oracle_ref_ue.py checks the fixture against the library,
tests/test_gpu_ref_ue.py checks the CUDA path against it."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "audio-pathtracer_b200"), os.path.dirname(HERE)]

import pyoracle as po                                       # noqa: E402
from test_oracle_ref_ue import pin_scene, closet_scene      # noqa: E402

out = {}
verts, tri_mat, ab, src, lis = pin_scene()
W = po.RefUEWorld(verts, tri_mat, 1.0 - ab[:, 2])
e, ir, n = W.update_source(src, lis, 1)
out["pin_seed1_energy"], out["pin_seed1_traces"] = e, np.uint64(n)
verts, tri_mat, ab, src, lis = closet_scene()
W = po.RefUEWorld(verts, tri_mat, 1.0 - ab[:, 2])
e, ir, n = W.update_source(src, lis, 5)
out["closet_seed5_energy"], out["closet_seed5_ir0"], out["closet_seed5_traces"] = e, ir[0], np.uint64(n)
np.savez_compressed(os.path.join(HERE, "ref_ue_v1.npz"), **out)
print({k: (v.shape, float(np.abs(v).sum())) for k, v in out.items()})
