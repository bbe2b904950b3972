"""The C-ABI library loads and exports every symbol include/mpassit_rg.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

from tests.conftest import ROOT


def _declared():
    txt = open(os.path.join(ROOT, "include", "mpassit_rg.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mprg_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(engine_lib):
    L = ctypes.CDLL(engine_lib)
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_binding_lists_every_declared_symbol(engine_lib):
    from mpassit_b200 import lib

    assert sorted(lib.EXPORTS) == _declared()
    lib.load()
    assert lib.load().mprg_version().startswith(b"mpassit-rg")


def test_init_without_gpu_fails_loudly(engine_lib, have_gpu):
    """No CPU fallback: on a box without CUDA the engine refuses to initialise."""
    import pytest

    from mpassit_b200.regrid import MprgError, Regridder

    if have_gpu:
        pytest.skip("GPU present")
    with pytest.raises(MprgError) as e:
        Regridder(device=0)
    assert "no CUDA device" in str(e.value)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure; the product package must not reference it."""
    pkg = os.path.join(ROOT, "mpassit_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        if "_build" in d:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "mpassit_oracle" in src \
                        or "orc_" in src:
                    bad.append(os.path.join(d, f))
    assert not bad, bad
