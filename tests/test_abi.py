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
