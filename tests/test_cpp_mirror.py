"""The C++ host-side mirror of the reference interface (include/frequensee.hpp) compiles against the C-ABI with a
plain g++, fails loudly without a GPU, and -- on the GPU -- reproduces the oracle through the reference's own call
sequence (RegisterGeometry / ForceUpdateSources / GetImpulseResponse / ProcessSourceAudio)."""
import math
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "audio-pathtracer_b200", "lib")
EXE = os.path.join(ROOT, "tests", "cpp", "mirror_smoke")


@pytest.fixture(scope="module")
def exe():
    src = os.path.join(ROOT, "tests", "cpp", "mirror_smoke.cpp")
    hpp = os.path.join(ROOT, "include", "frequensee.hpp")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(hpp)):
        env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                               "-L", LIBDIR, "-lfrequensee", "-Wl,-rpath," + LIBDIR], env=env)
    return EXE


def test_mirror_compiles_and_fails_loudly_without_gpu(exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([exe, "nogpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_mirror_reproduces_oracle_through_reference_call_sequence(exe, oracle):
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    kv = dict(re.findall(r"(\w+)=([-+.\w]+)", r.stdout))
    # the same scene through the oracle
    X, Y, Z = 7.0, 5.0, 3.0
    quads = [[[0, 0, 0], [0, Y, 0], [0, Y, Z], [0, 0, Z]], [[X, 0, 0], [X, Y, 0], [X, Y, Z], [X, 0, Z]],
             [[0, 0, 0], [X, 0, 0], [X, 0, Z], [0, 0, Z]], [[0, Y, 0], [X, Y, 0], [X, Y, Z], [0, Y, Z]],
             [[0, 0, 0], [X, 0, 0], [X, Y, 0], [0, Y, 0]], [[0, 0, Z], [X, 0, Z], [X, Y, Z], [0, Y, Z]]]
    tris, mats = [], []
    for f, q in enumerate(quads):
        tris += [[q[0], q[1], q[2]], [q[0], q[2], q[3]]]
        mats += [f, f]
    alpha = np.repeat(np.array([[0.02], [0.05], [0.12], [0.07], [0.37], [0.75]], np.float32), 8, axis=1)
    S = oracle.Scene(np.array(tris, np.float32), np.array(mats, np.uint32), alpha, use_bvh=False)
    cfg = oracle.default_config()
    h, st = S.trace(cfg, [[1.5, 1.2, 1.0]], [5.0, 3.5, 1.6], 4096, 8, 0x5EED)
    ir = oracle.build_ir(cfg, h[0], 4096)
    assert int(kv["connected"]) == st["connected"]
    assert abs(float(kv["ir_energy"]) / float((ir[0].astype(np.float64) ** 2).sum()) - 1) < 1e-5
    assert int(kv["ir_peak"]) == int(np.argmax(ir[0]))
    assert float(kv["conv_rel"]) < 1e-5                              # impulse in -> IR out (REV.cpp:172-213)
    assert float(kv["bin10"]) == pytest.approx(0.04) and float(kv["bin999"]) == pytest.approx(0.01)   # COMP.h:87-91
    e = np.zeros(1000, np.float32); e[10] = 0.04; e[999] = 0.01
    ir2 = oracle.build_ir_from_energy(cfg, e)
    assert float(kv["ir2_sample"]) == pytest.approx(float(ir2[1, 10 * 48 + 47]), rel=1e-5)
    assert int(kv["passthrough"]) == 1                               # bApplyReverb == false (REV.cpp:128-133)
