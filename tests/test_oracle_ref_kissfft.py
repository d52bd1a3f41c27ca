"""Pins the oracle's convolution semantics against the REFERENCE's own vendored KissFFT
(oracle/_ref, compiled in place from /root/reference by oracle/Makefile) running the reference's
own scheme: 65 536-point kiss_fftr of the 49 023-sample history and of the IR, multiply,
kiss_fftri (REV.cpp:172-213)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def ref(oracle):
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built (reference tree absent at build time)")
    return oracle


def test_kiss_fftr_matches_numpy(ref):
    rng = np.random.default_rng(0)
    for n in (2048, 65536):
        x = rng.uniform(-1, 1, n).astype(np.float32)
        X = ref.kiss_fftr(x)
        Xn = np.fft.rfft(x.astype(np.float64))
        assert np.linalg.norm(X - Xn) / np.linalg.norm(Xn) < 1e-6


def test_reference_fft_size_is_65536(ref):
    # RoundUpToPowerOfTwo(47 999 + 1024) (REV.cpp:82, 90); the "97022" comment there is wrong
    assert ref.RefKissConv().fft_size == 65536


def test_reference_scheme_equals_oracle_direct_form(ref):
    rng = np.random.default_rng(5)
    cfg = ref.default_config()
    ir = np.zeros((2, 48000), np.float32)
    ir[:, :6000] = (rng.normal(size=(2, 6000)) * np.exp(-np.arange(6000) / 1500.0) * 0.03).astype(np.float32)
    ir[0, 47999] = 0.01                                         # last tap matters too
    rc = ref.RefKissConv(); rc.set_ir(ir)
    cv = ref.Conv(cfg); cv.set_ir(ir)
    num = den = 0.0
    for b in range(50):                                         # > 47 blocks: history fully populated
        x = rng.uniform(-0.5, 0.5, size=(1024, 2)).astype(np.float32)
        if b == 0:
            x[0] = 1.0                                          # unit impulse at frame 0
        yr = rc.process(x); yo = cv.process(x)
        num += float(((yr - yo) ** 2).sum()); den += float((yo ** 2).sum())
    assert (num / den) ** 0.5 < 1e-5
