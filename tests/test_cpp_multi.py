"""Multi-GPU inside the C-ABI (fs_multi_*), driven by a plain C++ program: no Python, no torch, no collective library on the
data path.  On a one-GPU box the devices are two / three contexts on GPU 0 (the same code path: shard, peer-store into the
staging buffer of context 0, sum); with >= 2 GPUs it also runs across real devices."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "audio-pathtracer_b200", "lib")
EXE = os.path.join(ROOT, "tests", "cpp", "multi_smoke")


@pytest.fixture(scope="module")
def exe():
    src = os.path.join(ROOT, "tests", "cpp", "multi_smoke.cpp")
    hdr = os.path.join(ROOT, "include", "frequensee.h")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                               "-L", LIBDIR, "-lfrequensee", "-Wl,-rpath," + LIBDIR], env=env)
    return EXE


def test_multi_smoke_compiles(exe):
    assert os.path.exists(exe)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 3])
def test_multi_contexts_on_one_gpu_equal_single(exe, n):
    r = subprocess.run([exe, str(n), "same"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    j = json.loads(r.stdout.strip().splitlines()[-1])
    assert j["multi_equals_single"] and j["async_multi_equals_single"] and j["ir_equal"] and j["hist_sum"] > 0


@pytest.mark.gpu
def test_multi_real_devices_equal_single(exe):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU on this box (the 2- and 8-GPU runs are kept under profiles/)")
    r = subprocess.run([exe, str(min(n, 8))], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    j = json.loads(r.stdout.strip().splitlines()[-1])
    assert j["multi_equals_single"] and j["async_multi_equals_single"] and j["ir_equal"]


@pytest.mark.gpu
def test_multi_python_binding(fs):
    """the ctypes view of the same API"""
    import numpy as np
    from frequensee import scenes
    sc = scenes.shoebox()
    with fs.Context() as one, fs.MultiContext([0, 0]) as m:
        one.set_scene(sc.verts, sc.tri_mat, sc.absorption); m.set_scene(sc.verts, sc.tri_mat, sc.absorption)
        h1 = one.trace(sc.sources, sc.listener, 40001, 8, 3)
        assert np.array_equal(m.trace(sc.sources, sc.listener, 40001, 8, 3), h1)
        m.trace(sc.sources, sc.listener, 40001, 8, 3, want_hist=False)
        c0 = m.context(0)
        assert np.array_equal(c0.build_ir(0), one.build_ir(0))
        total, red = m.last_ms()
        assert total > 0 and red >= 0
