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


def test_fortran_module_binds_every_reference_facing_entry():
    """fortran/mpassit_rg_mod.F90 (what interp.F90 / model_grid.F90 / write_data.F90 would `use`) declares an
    ISO_C_BINDING interface for every entry point of the header, under its exact C name, and nothing else."""
    txt = open(os.path.join(ROOT, "fortran", "mpassit_rg_mod.F90")).read()
    bound = sorted(set(re.findall(r'bind\(C,\s*name="(mprg_[a-z0-9_]+)"\)', txt)))
    declared = _declared()
    # instrumentation that only the bench / tuning scripts use
    tooling = {"mprg_profile_enable", "mprg_profile_read", "mprg_profile_reset", "mprg_set_stream", "mprg_version",
               "mprg_route_export_csr", "mprg_route_export_w2", "mprg_route_import_csr", "mprg_host_alloc", "mprg_host_free", "mprg_device_alloc",
               "mprg_device_free", "mprg_kernel_launches", "mprg_last_ms", "mprg_route_src_referenced", "mprg_clear_routes",
               "mprg_route_info", "mprg_scratch", "mprg_has_rotation", "mprg_synchronize", "mprg_get_slab", "mprg_get_option", "mprg_weight_cache_stats", "mprg_get_target_lonlat"}
    assert not [b for b in bound if b not in declared], "Fortran binds a name the header does not declare"
    missing = [d for d in declared if d not in bound and d not in tooling]
    assert not missing, missing
    # the stagger / method / memory constants agree with the header
    hdr = open(os.path.join(ROOT, "include", "mpassit_rg.h")).read()
    for name in ("MPRG_BILINEAR", "MPRG_CONSERVE", "MPRG_NEAREST_STOD", "MPRG_CENTER", "MPRG_EDGE1", "MPRG_EDGE2", "MPRG_CORNER",
                 "MPRG_CENTER_HALO", "MPRG_F32", "MPRG_F64", "MPRG_HOST", "MPRG_DEVICE", "MPRG_EPI_ROT_U", "MPRG_EPI_ROT_V",
                 "MPRG_GRID_NOPERI", "MPRG_GRID_1PERI_MONOPOLE"):
        hv = re.search(name + r"\s*=\s*(\d+)", hdr)
        fv = re.search(name + r"\s*=\s*(\d+)", txt)
        assert hv and fv and hv.group(1) == fv.group(1), name


def test_host_header_symbols_exported(engine_lib):
    """libmpassit_host.so (the C++ mirror of the Fortran host stages) exports every function include/mpassit_host.h
    declares, and the Python binding gives each of them an argument list."""
    from mpassit_b200 import build, host

    build.build_host()
    txt = open(os.path.join(ROOT, "include", "mpassit_host.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"typedef[^;]*\(\*\w+\)[^;]*;", "", txt)          # function-pointer typedefs are not exports
    names = sorted(set(re.findall(r"\b(mpassit_[a-z0-9_]+)\s*\(", txt)))
    assert len(names) >= 20, names
    L = host.load()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    unbound = [n for n in names if getattr(L, n).argtypes is None]
    assert not unbound, unbound
