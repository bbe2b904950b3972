"""`mpassit <namelist>` with files on both sides (mpassit_b200/host/run.cpp) on the GPU:
MPAS NetCDF-classic files -> mapped sources, byte swap in HBM, interp_data, WRF post-ops in HBM, per-slab pwrite ->
output file, compared with the in-memory pass of the same fields (itself held to the oracle by test_gpu_methods)
plus the writer's arithmetic (write_data.F90:1339-1432) restated in numpy."""
import ctypes as C
import os

import numpy as np
import pytest

from mpassit_b200 import check
from tests import mpas_files

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host(engine_lib):
    from mpassit_b200 import build, host

    build.build_host()
    host.load()
    return host


def _reference_pass(wl):
    """Device-buffer interp_data of the synthetic fields: {output name: [nlev][nj][ni] fp32}, plus the sources."""
    from mpassit_b200 import lib as l
    from mpassit_b200 import workload
    from mpassit_b200.regrid import Regridder

    rg = Regridder(device=0)
    workload.load_geometry(rg, wl)
    F = workload.make_fields(wl, device="cuda:0")
    workload.run_interp(rg, wl, F["dev"], l.DEVICE)
    rg.synchronize()
    D = F["dev"]
    src = {g: [(s.name, s.src.cpu().numpy()) for s in D[g]] for g in ("diag", "hist_2d", "hist_3d", "soil")}
    ter = D["ter"].cpu().numpy()
    njM, niM = wl.grids["M"][0].shape
    got = {}
    wrf = bool(wl.cfg.wrf_mod_vars)
    for g in ("diag", "hist_2d", "hist_3d", "soil"):
        for s in D[g]:
            if wrf and s.name.startswith("uReconstruct"):
                continue  # they leave through the staggered U / V
            got[s.target_name] = s.dst.cpu().numpy().reshape(s.nlev, njM, niM)
    if wl.cfg.interp_hist:
        got["HGT"] = D["hgt"].cpu().numpy().reshape(njM, niM)
    if wrf:
        got["U"] = D["u_stag"].cpu().numpy().reshape((wl.nz,) + wl.grids["U"][0].shape)
        got["V"] = D["v_stag"].cpu().numpy().reshape((wl.nz,) + wl.grids["V"][0].shape)
    rg.close()
    return got, src, ter


def _expect_file(got, wl):
    """What write_to_file puts in the file for the regridded fields `got` (wrf_mod_vars = .true.)."""
    f32 = np.float32
    # 2-D variables are [south_north][west_east] in the file (record dimension dropped by the reader)
    want = {k: (v[0] if (v.ndim == 3 and v.shape[0] == 1) else v) for k, v in got.items()}
    wrf = bool(wl.cfg.wrf_mod_vars)
    if wrf:
        want["T"] = got["T"] * f32(1.0) + f32(-300.0)           # write_data.F90:1339-1345 (the `< 10` guard is a no-op)
    phb = got["PHB"]
    zc = np.empty_like(phb)
    zc[:-1] = f32(0.5) * (phb[1:] + phb[:-1])                    # :1406-1412
    zc[-1] = 9.9692099683868690e+36                              # the level the reference never writes: fill value
    want["Z_C"] = zc
    want["PHB"] = phb * f32(9.81)                                # :1414 (keyed on the name alone, like Z_C)
    if not wrf:
        return want
    want["PB"] = got["P_HYD"]                                    # :1376 writes dum3dt, which still holds P_HYD
    for z, like in (("MU", "MUB"), ("P", "P_HYD"), ("PH", "PHB")):
        want[z] = np.zeros_like(got[like])                       # :1356, :1468, :1424
    p = got["P_HYD"].astype(np.float64)
    top = p[-1]
    cand = top[top >= 10.0] * 0.8
    want["P_TOP"] = f32(min(p.max(), cand.min() if cand.size else np.inf))   # :1364-1373
    return want


def _check(out, want, exact=("XLAND", "TSLB", "SMOIS", "SH2O", "MU", "P", "PH", "PB")):
    checked = 0
    for k, w in want.items():
        g = out[k]
        assert g.shape == w.shape, (k, g.shape, w.shape)
        if k in exact:
            assert np.array_equal(g, w), k
        else:
            scale = max(float(np.abs(w[np.abs(w) < 1e30]).max()), 1e-30)
            assert np.abs(g.astype(np.float64) - w.astype(np.float64)).max() <= 2e-6 * scale, (k, np.abs(g - w).max(), scale)
        checked += 1
    return checked


def test_files_in_files_out_single_rank(host, tmp_path):
    from mpassit_b200 import workload

    wl = workload.make("mini", rundir=str(tmp_path))
    got, src, ter = _reference_pass(wl)
    nl, paths = mpas_files.write_case(wl, str(tmp_path), src, ter)
    st = host.run(nl, str(tmp_path), device=0)
    out, g, va, dims, order = mpas_files.read_output(paths["out"])
    want = _expect_file(got, wl)
    assert set(want) <= set(out)
    assert _check(out, want) >= 45
    assert st.n_vars_written == len(got) - 1 and st.output_version == 2  # HGT is written apart
    assert abs(st.p_top - float(want["P_TOP"])) <= 1e-6 * abs(float(want["P_TOP"]))
    nM = wl.n_mass
    assert st.bytes_in == sum(a.size for gl in src.values() for _, a in gl) * 4
    assert st.bytes_out >= sum(v.size for k, v in got.items()) * 4 + got["P_HYD"].size * 4 + (wl.nz * nM) * 4
    # a field the engine never wrote stays zero; the grid description is there
    assert not out["P"].any() and out["XLAT"].shape == wl.grids["M"][0].shape


def test_files_double_precision_sources(host, tmp_path):
    """An MPAS file written in double precision (input_data.F90 reads into R8 either way): NC_DOUBLE sources are
    swapped as 8-byte words and regridded to the same fp32 outputs."""
    from mpassit_b200 import workload

    wl = workload.make("mini", rundir=str(tmp_path))
    got, src, ter = _reference_pass(wl)
    nl, paths = mpas_files.write_case(wl, str(tmp_path), src, ter, real="f8")
    host.run(nl, str(tmp_path), device=0)
    out = mpas_files.read_output(paths["out"])[0]
    want = _expect_file(got, wl)
    # fp64 sources accumulate in fp64: equal to the fp32 pass within its accumulation error
    for k in ("T", "QVAPOR", "PSFC", "U", "V", "REFL_10CM", "SNOW", "HGT"):
        check.assert_field_close(out[k], want[k], k)
    for k in ("XLAND", "TSLB"):
        assert np.array_equal(out[k], want[k]), k


def test_files_two_rank_slabs_tile_the_single_rank_file(host, tmp_path):
    """Two ranks (emulated one after the other on one device) write their own rows of the same file with pwrite;
    the result equals the single-rank file byte for byte, P_TOP included (its max / min partials are combined by
    the host's comm callback, here replayed from a recording pass)."""
    from mpassit_b200 import workload

    wl = workload.make("mini", rundir=str(tmp_path))
    got, src, ter = _reference_pass(wl)
    nl, paths = mpas_files.write_case(wl, str(tmp_path), src, ter)
    host.run(nl, str(tmp_path), device=0)
    single = open(paths["out"], "rb").read()
    os.remove(paths["out"])

    rec = {host.COMM_MAX: [], host.COMM_MIN: []}

    def recording(_a, op, vals, n):
        if op != host.COMM_BARRIER:
            rec[op].append(vals[0])

    def combined(_a, op, vals, n):
        if op == host.COMM_MAX:
            vals[0] = max(rec[op])
        elif op == host.COMM_MIN:
            vals[0] = min(rec[op])

    for cb in (recording, combined):
        fn = host.COMM_FN(cb)
        for rank in (0, 1):
            host.run(nl, str(tmp_path), device=0, rank=rank, nranks=2, comm=fn)
    assert len(rec[host.COMM_MAX]) == 2
    assert open(paths["out"], "rb").read() == single


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_unpinned_big_endian_sources_take_the_bounce_ring(engine_lib, dtype):
    """A large host source that is not page-locked (what a mapped file variable is) goes through the engine's
    parallel bounce ring (capi.cu: upload_unpinned) and, flagged big-endian, is swapped in HBM: same result, bit for
    bit, as the native-order array uploaded from a small (direct-copy) or device buffer."""
    import torch

    from mpassit_b200.regrid import Regridder

    rg = Regridder(device=0)
    rng = np.random.default_rng(11)
    nSrc, nDst, nlev = 450_001, 30_000, 23          # 41 / 83 MB per field: several 8-MiB chunks, ragged tail
    lens = rng.integers(0, 4, nDst)
    rp = np.zeros(nDst + 1, np.int32)
    np.cumsum(lens, out=rp[1:])
    col = rng.integers(0, nSrc, rp[-1]).astype(np.int32)
    col[:2] = (0, nSrc - 1)
    r = rg.import_csr(nSrc, rp, col, rng.random(rp[-1]))
    a = rng.standard_normal((nSrc, nlev)).astype(dtype)
    b = rng.standard_normal((nSrc, nlev)).astype(dtype)
    want = [torch.empty((nlev, nDst), dtype=torch.float32, device="cuda") for _ in range(2)]
    rg.apply(r, [torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()], want, nlev=[nlev, nlev])
    got = [np.empty((nlev, nDst), np.float32) for _ in range(2)]
    rg.apply(r, [a, b], got, nlev=[nlev, nlev])                     # native order, unpinned
    for g, w in zip(got, want):
        assert np.array_equal(g, w.cpu().numpy())
    got = [np.empty((nlev, nDst), np.float32) for _ in range(2)]
    rg.set_source_byte_order(True)
    rg.apply(r, [a.byteswap(), b.byteswap()], got, nlev=[nlev, nlev])  # file order
    rg.set_source_byte_order(False)
    for g, w in zip(got, want):
        assert np.array_equal(g, w.cpu().numpy())
    # device-side swap is an involution
    t = torch.from_numpy(a[:1000]).cuda()
    rg.bswap(t)
    assert np.array_equal(t.cpu().numpy().view(np.uint8), a[:1000].byteswap().view(np.uint8))
    rg.bswap(t)
    assert np.array_equal(t.cpu().numpy(), a[:1000])
    r.release()
    rg.close()


def test_files_global_latlon_target(host, tmp_path):
    """BASELINE configs[0] through files: 40,962-cell global mesh, 55 levels -> 1-degree lat-lon, hist lists only,
    no rotation (the projection is not Lambert), winds staggered on the non-periodic grid."""
    from mpassit_b200 import workload

    wl = workload.make("c1", rundir=str(tmp_path))
    got, src, ter = _reference_pass(wl)
    nl, paths = mpas_files.write_case(wl, str(tmp_path), src, ter)
    st = host.run(nl, str(tmp_path), device=0)
    out, g, va, dims, order = mpas_files.read_output(paths["out"])
    want = _expect_file(got, wl)
    assert _check(out, want) >= 28 and "SINALPHA" not in out and g["MAP_PROJ"] == 0
    assert st.n_cells == 40962 and dims["bottom_top"] == 55


def test_files_without_wrf_mod_vars_and_cdf5_inputs(host, tmp_path):
    """wrf_mod_vars = .false.: the winds are ordinary 3d_nz fields at mass points (no rotation, no staggering), T is
    written as regridded, none of MU / P_TOP / PH / P / PB exists; the inputs are in the 64-bit-data format (CDF-5)."""
    from mpassit_b200 import workload

    wl = workload.make("mini", rundir=str(tmp_path))
    wl.cfg.wrf_mod_vars = 0
    got, src, ter = _reference_pass(wl)
    assert got["U"].shape == (wl.nz,) + wl.grids["M"][0].shape
    nl, paths = mpas_files.write_case(wl, str(tmp_path), src, ter)
    for k in ("init", "diag", "history"):
        host.nc_copy(paths[k], paths[k] + ".cdf5", 5)
        os.replace(paths[k] + ".cdf5", paths[k])
    host.run(nl, str(tmp_path), device=0)
    out, g, va, dims, order = mpas_files.read_output(paths["out"])
    want = _expect_file(got, wl)
    assert _check(out, want) >= 44
    assert np.array_equal(out["T"], got["T"]) and not ({"MU", "P_TOP", "PH", "P", "PB"} & set(order))


def test_engine_weights_round_trip_through_esmf_weight_files(engine_lib, host, tmp_path):
    """The engine's bilinear / nearest / conservative matrices written in ESMF's weight-file format, read back and
    imported as routes (what one would do with ESMF's own file) regrid a field to the same result."""
    import torch

    from mpassit_b200 import lib as l
    from mpassit_b200 import workload
    from mpassit_b200.regrid import Regridder
    from tools import esmf_kit

    paths = esmf_kit.export("mini", str(tmp_path))
    wl = workload.make("mini", rundir=str(tmp_path))
    rg = Regridder(device=0)
    workload.load_geometry(rg, wl)
    src = torch.from_numpy(np.random.default_rng(4).standard_normal((wl.mesh.lonCell.size, 8)).astype(np.float32)).cuda()
    for name, method in (("bilinear", l.BILINEAR), ("nearest", l.NEAREST_STOD), ("conserve", l.CONSERVE)):
        r = rg.store(method, l.SRC_MESH_ELEMENT, l.CENTER)
        want = torch.empty((8, wl.n_mass), dtype=torch.float32, device="cuda")
        rg.apply(r, [src], [want], nlev=[8])
        na, nb, rp, col, w = host.read_esmf_weights(paths[name])
        assert (na, nb) == (wl.mesh.lonCell.size, wl.n_mass)
        r2 = rg.import_csr(na, rp, col, w)
        got = torch.empty_like(want)
        rg.apply(r2, [src], [got], nlev=[8])
        if name == "nearest":
            assert torch.equal(got, want), name
        else:  # same matrix, same kernels; the imported route may pick another launch shape
            assert float((got - want).abs().max()) <= 1e-6 * float(want.abs().max()), name
        assert esmf_kit.compare(paths[name], paths[name])["verdict"].startswith("identical structure, weights")
        r.release()
        r2.release()
    rg.close()


def test_files_target_grid_from_a_wrf_style_file(host, tmp_path):
    """target_grid_type = 'file': the target comes from a WRF-style file (here: the output of a parameter-mode run).
    The run must equal an in-memory pass on the grid such a file holds -- float-rounded centres, staggers and rotation
    angles, corners synthesised by get_cell_corners -- plus the writer's arithmetic."""
    import dataclasses

    from mpassit_b200 import workload

    wl = workload.make("mini", rundir=str(tmp_path))
    got, src, ter = _reference_pass(wl)
    nl, paths = mpas_files.write_case(wl, str(tmp_path), src, ter)
    host.run(nl, str(tmp_path), device=0)
    target = str(tmp_path / "wrf_target.nc")
    os.replace(paths["out"], target)
    t = mpas_files.read_output(target)[0]
    f8 = lambda k: t[k].astype(np.float64)  # noqa: E731
    grids = {"M": (f8("XLAT"), f8("XLONG")), "U": (f8("XLAT_U"), f8("XLONG_U")), "V": (f8("XLAT_V"), f8("XLONG_V"))}
    grids["CORNER"] = host.get_cell_corners(*grids["M"], wl.cfg.dx)
    wl2 = dataclasses.replace(wl, grids=grids, cosa=f8("COSALPHA"), sina=f8("SINALPHA"))
    got2, _, _ = _reference_pass(wl2)
    nl2, paths2 = mpas_files.write_case(wl, str(tmp_path), src, ter, target_file=target)
    host.run(nl2, str(tmp_path), device=0)
    out, g, va, dims, order = mpas_files.read_output(paths2["out"])
    want = _expect_file(got2, wl)
    # the reference's corner bearings turn the last column / row of target cells inside out: conservative
    # fields are compared as the same bits, whatever they are there
    for k in ("SNOW", "SNOWH"):
        assert np.array_equal(out[k], want.pop(k), equal_nan=True), k
    assert _check(out, want) >= 44
    assert np.array_equal(out["XLAT"], t["XLAT"]) and np.array_equal(out["MAPFAC_U"], t["MAPFAC_U"])
    # and the float-rounded grid really is another grid: the parameter-mode result differs in the last bits
    assert not np.array_equal(got2["PSFC"], got["PSFC"])
    assert np.abs(got2["PSFC"] - got["PSFC"]).max() <= 1e-3 * np.abs(got["PSFC"]).max()
