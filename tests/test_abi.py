"""The C-ABI library loads on a CPU-only box and exports every symbol include/frequensee.h
declares; the struct layouts the Python binding and the oracle assume match the header."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "frequensee.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(fs):
    L = fs.capi.load()
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), "libfrequensee.so does not export %s" % s
    assert sorted(fs.capi.ABI_SYMBOLS) == syms


def test_config_layout_matches_oracle(fs, oracle):
    a, b = fs.capi.Config, oracle.Config
    assert C.sizeof(a) == C.sizeof(b) == 116
    for (na, _), (nb, _) in zip(a._fields_[:19], b._fields_[:19]):
        assert na == nb and getattr(a, na).offset == getattr(b, nb).offset
    ca, cb = fs.default_config(), oracle.default_config()
    assert bytes(ca)[:104] == bytes(cb)[:104]              # same defaults up to the product-only tail


def test_path_dbg_layout(fs, oracle):
    assert fs.capi.PATH_DBG_DTYPE == oracle.PATH_DBG_DTYPE and fs.capi.PATH_DBG_DTYPE.itemsize == 80


def test_no_cpu_fallback(fs):
    """without a GPU the product must fail loudly, never compute on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fs.FrequenSeeError) as ei:
        fs.Context()
    assert ei.value.code == fs.capi.FS_ERR_CUDA and "no CPU fallback" in str(ei.value)


def test_product_does_not_reference_oracle():
    """nothing under the product tree may import / include / link the oracle"""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "audio-pathtracer_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"pyoracle|fs_oracle|liboracle|oracle/", txt):
                    bad.append(os.path.join(dp, f))
    # mentions in comments are allowed only in docs, not in code files
    assert bad == [], bad


def test_float_array_text_files_round_trip(fs, tmp_path):
    """saved_ir.txt format of the reference (COMP.cpp:454-505): one float per line; host only, no GPU needed"""
    import numpy as np
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.normal(size=2000).astype(np.float32) * np.float32(1e-3),
                        np.array([0.0, 1.0, -1.0, 1e-30, 3.4e38, -2.5e-7, 16777216.0], np.float32)])
    p = str(tmp_path / "saved_ir.txt")
    fs.save_float_array(p, x)
    lines = open(p).read().split("\n")
    assert len(lines) == len(x) and lines[-7:-4] == ["0.0", "1.0", "-1.0"]       # SanitizeFloat keeps one fractional digit
    assert np.array_equal(fs.load_float_array(p), x)                             # bit-exact round trip
    # files written by the reference: "%f"-style lines, blank lines, CRLF, a trailing newline, junk -> Atof gives 0
    open(p, "w").write("0.500000\r\n\n-0.125000\n1e-3\nabc\n  2.000000  \n")
    assert np.array_equal(fs.load_float_array(p), np.array([0.5, -0.125, 1e-3, 0.0, 2.0], np.float32))
    import pytest
    with pytest.raises(fs.FrequenSeeError):
        fs.load_float_array(str(tmp_path / "missing.txt"))
