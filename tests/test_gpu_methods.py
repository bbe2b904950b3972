"""GPU parity for the remaining regrid classes and for the whole interp_data pass:
grid->grid stagger bilinear, conservative, node-based bilinear, and the host mirror's
call sequence (interp.F90:92-465) end to end, all through the C ABI."""
import numpy as np
import pytest

from mpassit_b200 import check, defaults
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rg(engine_lib):
    from mpassit_b200.regrid import Regridder

    r = Regridder(device=0)
    yield r
    r.close()


@pytest.fixture(scope="module")
def host(engine_lib):
    from mpassit_b200 import build, host

    build.build_host()
    host.load()
    return host


@pytest.fixture(scope="module")
def lc_case(host, tmp_path_factory):
    """30-km Lambert target (60 x 40 mass points) over an irregular regional mesh that does
    not cover it completely (unmapped rows + extrapolating nearest rows both occur)."""
    d = tmp_path_factory.mktemp("lc")
    cfg = host.read_setup_namelist(defaults.write_namelist(str(d / "namelist.input"), nx=61, ny=41, dx=30000.0))
    grids = {k: host.target_coords(cfg, s) for k, s in (("M", 0), ("U", 1), ("V", 2), ("CORNER", 3))}
    mesh = H.synth.regional_delaunay_mesh(9000, extent_x_m=1700e3, extent_y_m=1150e3, seed=21, lloyd_iters=3)
    cosa, sina = host.get_rotang(*grids["M"])
    return cfg, grids, mesh, cosa, sina


def _load(rg, grids, mesh):
    rg.set_mesh(mesh.lonCell, mesh.latCell, mesh.lonVertex, mesh.latVertex, mesh.verticesOnCell)
    for k, s in (("M", 0), ("U", 1), ("V", 2), ("CORNER", 3)):
        lat, lon = grids[k]
        rg.set_target(s, lon, lat)
    cosa, sina = __import__("mpassit_b200.host", fromlist=["x"]).get_rotang(*grids["M"])
    rg.set_rotation(cosa, sina)


def test_stagger_routes_bit_exact(rg, orc, lc_case):
    from mpassit_b200 import lib as l

    cfg, grids, mesh, cosa, sina = lc_case
    _load(rg, grids, mesh)
    lat, lon = grids["M"]
    sx = orc.sph_deg_to_cart(lon, lat).reshape(lat.shape[0], lat.shape[1], 3)
    for key, stag in (("U", l.EDGE1), ("V", l.EDGE2)):
        slat, slon = grids[key]
        e, c, w = orc.bilinear_quadgrid(sx, orc.sph_deg_to_cart(slon, slat), brute=True)
        rp, cc, ww = orc.ell_to_csr(e >= 0, c, w)
        r = rg.store(l.BILINEAR, l.SRC_GRID_CENTER, stag)
        grp, gc, gw = r.export_csr()
        assert np.array_equal(grp, rp) and np.array_equal(gc, cc)
        assert np.abs(gw - ww).max() <= 1e-12
        un = (e < 0).reshape(slat.shape)
        # edge points outside the hull of the centres are unmapped: first/last column of U, first/last row of V
        if key == "U":
            assert un[:, 0].all() and un[:, -1].all() and not un[1:-1, 1:-1].any()
        else:
            assert un[0, :].all() and un[-1, :].all() and not un[1:-1, 1:-1].any()
        # apply: fp64 level-slowest source -> fp32
        src = np.random.default_rng(1).standard_normal((5, lat.size))
        got = np.empty((5, slat.size), np.float32)
        import torch

        dsrc = torch.from_numpy(src).cuda()
        dgot = torch.empty((5, slat.size), dtype=torch.float32, device="cuda")
        rg.apply(r, [dsrc], [dgot], nlev=[5])
        rg.synchronize()
        got = dgot.cpu().numpy()
        want = orc.apply_planes(rp, cc, ww, src).astype(np.float32)
        np.testing.assert_allclose(got, want, rtol=2e-7, atol=1e-7)
        r.release()


def test_conserve_structure_weights_and_conservation(rg, orc, lc_case):
    from mpassit_b200 import lib as l

    cfg, grids, mesh, cosa, sina = lc_case
    _load(rg, grids, mesh)
    cxyz, vxyz, tri = H.oracle_geometry(orc, mesh)
    clat, clon = grids["CORNER"]
    cor = orc.sph_deg_to_cart(clon, clat).reshape(clat.shape[0], clat.shape[1], 3)
    rp, cc, ww = orc.conserve(cxyz, vxyz, mesh.verticesOnCell, cor)
    r = rg.store(l.CONSERVE, l.SRC_MESH_ELEMENT, l.CENTER)
    grp, gc, gw = r.export_csr()
    assert np.array_equal(grp, rp)            # which destination cells are touched, and by how many sources
    assert np.array_equal(gc, cc)             # which sources (ascending ids)
    assert np.abs(gw - ww).max() <= 1e-12
    frac = np.zeros(rp.size - 1)
    np.add.at(frac, np.repeat(np.arange(rp.size - 1), np.diff(rp)), gw)
    # (the planar-Delaunay synthetic mesh is not exactly Delaunay on the sphere, so neighbouring polygons can
    #  overlap by slivers of ~1e-7 of a cell; a true Voronoi tiling would give <= 1 to rounding)
    assert frac.max() <= 1 + 1e-6 and (frac < 0.5).any() and (np.abs(frac - 1) < 1e-6).any()
    # constant field -> covered fraction; patchy snow field vs oracle
    lo, la = mesh.lonCell, mesh.latCell
    snow = H.synth.patchy_field(lo, la)
    ones = np.ones(mesh.nCells, np.float32)
    o1 = np.empty((1, frac.size), np.float32)
    o2 = np.empty((1, frac.size), np.float32)
    rg.apply(r, [ones, snow], [o1, o2], nlev=[1, 1])
    np.testing.assert_allclose(o1[0], frac.astype(np.float32), rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(o2, orc.apply(rp, cc, ww, snow, np.float32), rtol=1e-5, atol=1e-6)
    r.release()


def test_node_bilinear(rg, orc, lc_case):
    from mpassit_b200 import lib as l

    cfg, grids, mesh, cosa, sina = lc_case
    _load(rg, grids, mesh)
    cxyz, vxyz, tri = H.oracle_geometry(orc, mesh)
    lat, lon = grids["M"]
    e, c, w = orc.bilinear_node(cxyz, vxyz, mesh.verticesOnCell, orc.sph_deg_to_cart(lon, lat), brute=True)
    rp, cc, ww = orc.ell_to_csr(e >= 0, c, w)
    r = rg.store(l.BILINEAR, l.SRC_MESH_NODE, l.CENTER)
    grp, gc, gw = r.export_csr()
    assert np.array_equal(grp, rp) and np.array_equal(gc, cc) and np.abs(gw - ww).max() <= 1e-12
    vort = H.synth.smooth_field(mesh.lonVertex, mesh.latVertex, 12, seed=4)
    got = np.empty((12, lat.size), np.float32)
    rg.apply(r, [vort], [got])
    np.testing.assert_allclose(got, orc.apply(rp, cc, ww, vort, np.float32), rtol=1e-6)
    r.release()


def _fields(mesh, nz=8, nsoil=4):
    lo, la = mesh.lonCell, mesh.latCell
    S = H.synth
    diag = []
    for k, (nm, _) in enumerate(defaults.DIAGLIST):
        diag.append((nm, S.smooth_field(lo, la, nz if nm == "refl10cm" else 1, seed=100 + k).reshape(mesh.nCells, -1)))
    h2 = []
    for k, (nm, _) in enumerate(defaults.HISTLIST_2D):
        if nm in ("snow", "snowh"):
            a = S.patchy_field(lo, la, seed=k)
        elif nm == "xland":
            a = S.integer_field(mesh.nCells, 2)
        else:
            a = S.smooth_field(lo, la, 1, seed=200 + k).reshape(-1)
        h2.append((nm, a.reshape(mesh.nCells, 1)))
    h3 = []
    for k, (nm, _) in enumerate(defaults.HISTLIST_3D):
        nl = nz + 1 if nm in ("zgrid", "w") else nz
        a = S.moisture_field(lo, la, nl, seed=300 + k) if nm.startswith("q") else S.smooth_field(lo, la, nl, seed=300 + k)
        h3.append((nm, a))
    soil = [(nm, np.ascontiguousarray(np.stack([S.integer_field(mesh.nCells, 30, seed=400 + 7 * k + s) for s in range(nsoil)], 1)))
            for k, (nm, _) in enumerate(defaults.HISTLIST_SOIL)]
    ter = S.smooth_field(lo, la, 1, seed=9).reshape(mesh.nCells, 1)
    return dict(diag=diag, hist_2d=h2, hist_3d=h3, soil=soil, ter=ter)


def test_interp_data_end_to_end(rg, orc, host, lc_case):
    """Default parm lists + wrf_mod_vars: diag bundle, 2d patch, hgt, 3d nz/nzp1, wind chain with rotation and
    staggering, conservative snow, nearest xland, and the soil bundle inheriting NEAREST_STOD."""
    cfg, grids, mesh, cosa, sina = lc_case
    _load(rg, grids, mesh)
    nz, nsoil = 8, 4
    F = _fields(mesh, nz, nsoil)
    want = H.oracle_interp(orc, mesh, grids, F, cosa, sina)
    nM, nU, nV = grids["M"][0].size, grids["U"][0].size, grids["V"][0].size

    def specs(items, table):
        tn = dict(table)
        return [host.FieldSpec(nm, tn.get(nm, nm), a.shape[1], a, np.full((a.shape[1], nM), np.nan, np.float32)) for nm, a in items]

    diag, h2, h3, soil = (specs(F["diag"], defaults.DIAGLIST), specs(F["hist_2d"], defaults.HISTLIST_2D),
                          specs(F["hist_3d"], defaults.HISTLIST_3D), specs(F["soil"], defaults.HISTLIST_SOIL))
    hgt = np.full((1, nM), np.nan, np.float32)
    ust = np.full((nz, nU), np.nan, np.float32)
    vst = np.full((nz, nV), np.nan, np.float32)
    n0 = rg.kernel_launches
    io = host.interp_data(rg, cfg, diag=diag, hist_2d=h2, hist_3d=h3, soil=soil, ter=F["ter"], hgt=hgt, u_stag=ust,
                          v_stag=vst, nz=nz)
    assert rg.kernel_launches > n0
    got = {s.name: s.dst for s in diag + h2 + h3 + soil}
    got["HGT"], got["U"], got["V"] = hgt, ust, vst
    # regrid classes the mirror assigned (input_data.F90:840-911)
    k3 = {F["hist_3d"][i][0]: io.hist_3d[i].klass for i in range(len(h3))}
    assert k3["uReconstructZonal"] == host.CLASS_U and k3["zgrid"] == host.CLASS_3D_NZP1 and k3["theta"] == host.CLASS_3D_NZ
    exact = {"xland", "tslb", "smois", "sh2o"}            # nearest-neighbour classes: bit-exact
    for nm, w in want.items():
        if nm in ("uReconstructZonal", "uReconstructMeridional"):
            continue
        g = got[nm]
        assert not np.isnan(g).any(), nm
        if nm in exact:
            assert np.array_equal(g, w), nm
        else:
            check.assert_field_close(g, w, nm)      # per element: |g - w| <= 1e-5 |w| + 1e-7 max|w|
    # the wind inputs are peeled off into U/V (their own dst buffers stay untouched)
    assert np.isnan(got["uReconstructZonal"]).all()
    # soil really used nearest neighbour (values are integers from the source, never blended)
    assert np.array_equal(got["tslb"], np.round(got["tslb"]))
    # unmapped bilinear points are exactly zero, and some exist in this case
    assert (got["theta"] == 0).any() and (got["xland"] != 0).all()


@pytest.mark.parametrize("nranks", [2, 3])
def test_interp_data_rank_slabs_equal_single_rank(engine_lib, host, lc_case, nranks):
    """The whole pass on `nranks` row slabs (ranks emulated one after another on one device; the regrid
    itself needs no collective) reproduces the single-rank output bit-for-bit for every field, including
    the staggered winds whose mass-point halo rows are recomputed locally (MPRG_CENTER_HALO)."""
    from mpassit_b200 import lib as l
    from mpassit_b200.regrid import Regridder

    cfg, grids, mesh, cosa, sina = lc_case
    nz, nsoil = 8, 4
    F = _fields(mesh, nz, nsoil)
    shapes = {k: grids[k][0].shape for k in ("M", "U", "V")}

    def run(rank, n):
        r = Regridder(device=0, rank=rank, nranks=n)
        _load(r, grids, mesh)
        rows = {k: r.slab(s) for k, s in (("M", l.CENTER), ("U", l.EDGE1), ("V", l.EDGE2))}
        nM = (rows["M"][1] - rows["M"][0]) * shapes["M"][1]
        nU = (rows["U"][1] - rows["U"][0]) * shapes["U"][1]
        nV = (rows["V"][1] - rows["V"][0]) * shapes["V"][1]

        def specs(items, table):
            tn = dict(table)
            return [host.FieldSpec(nm, tn.get(nm, nm), a.shape[1], a, np.full((a.shape[1], nM), np.nan, np.float32))
                    for nm, a in items]

        diag, h2, h3, soil = (specs(F["diag"], defaults.DIAGLIST), specs(F["hist_2d"], defaults.HISTLIST_2D),
                              specs(F["hist_3d"], defaults.HISTLIST_3D), specs(F["soil"], defaults.HISTLIST_SOIL))
        hgt = np.full((1, nM), np.nan, np.float32)
        ust = np.full((nz, nU), np.nan, np.float32)
        vst = np.full((nz, nV), np.nan, np.float32)
        host.interp_data(r, cfg, diag=diag, hist_2d=h2, hist_3d=h3, soil=soil, ter=F["ter"], hgt=hgt, u_stag=ust,
                         v_stag=vst, nz=nz)
        out = {s.name: (s.dst, "M") for s in diag + h2 + h3 + soil if not s.name.startswith("uReconstruct")}
        out["HGT"], out["U"], out["V"] = (hgt, "M"), (ust, "U"), (vst, "V")
        r.close()
        return out, rows

    single, _ = run(0, 1)
    parts = [run(k, nranks) for k in range(nranks)]
    for name, (full, stag) in single.items():
        ni = shapes[stag][1]
        pieces = [p[0][name][0].reshape(full.shape[0], -1, ni) for p in parts]
        got = np.concatenate(pieces, axis=1).reshape(full.shape[0], -1)
        assert got.shape == full.shape, name
        assert not np.isnan(got).any(), name
        assert np.array_equal(got, full), name
    # slabs tile each stagger's rows exactly once
    for k in ("M", "U", "V"):
        spans = [p[1][k] for p in parts]
        assert spans[0][0] == 0 and spans[-1][1] == shapes[k][0]
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_interp_data_global_latlon_target(engine_lib, orc, host, tmp_path):
    """BASELINE.json configs[0] in miniature: quasi-uniform global mesh -> global lat-lon target
    (`target_grid_type='lat-lon'`, no wind rotation: interp.F90:138,291 rotate on Lambert only),
    histlist_2d/3d/soil through the host mirror, checked against the oracle's interp_data."""
    from mpassit_b200.regrid import Regridder

    f = tmp_path / "namelist.input"
    f.write_text("&config\n target_grid_type = 'lat-lon'\n nx = 73\n ny = 37\n is_regional = .false.\n"
                 " stand_lon = -180.\n interp_diag = .false.\n interp_hist = .true.\n wrf_mod_vars = .true.\n/\n")
    cfg = host.read_setup_namelist(str(f))
    assert cfg.proj_code == host.PROJ_LATLON
    grids = {k: host.target_coords(cfg, s) for k, s in (("M", 0), ("U", 1), ("V", 2), ("CORNER", 3))}
    assert grids["M"][0].shape == (36, 72)
    mesh = H.small_global(2562, jitter=0.1, seed=3)
    nz, nsoil = 7, 4            # 7 levels: columns not 16-byte aligned
    F = _fields(mesh, nz, nsoil)
    F["diag"] = []
    want = H.oracle_interp(orc, mesh, grids, F, None, None, lc=False, periodic=True)   # is_regional = .false.
    rg = Regridder(device=0)
    rg.set_mesh(mesh.lonCell, mesh.latCell, mesh.lonVertex, mesh.latVertex, mesh.verticesOnCell)
    for k, s in (("M", 0), ("U", 1), ("V", 2), ("CORNER", 3)):
        rg.set_target(s, grids[k][1], grids[k][0])
    nM, nU, nV = grids["M"][0].size, grids["U"][0].size, grids["V"][0].size

    def specs(items, table):
        tn = dict(table)
        return [host.FieldSpec(nm, tn.get(nm, nm), a.shape[1], a, np.full((a.shape[1], nM), np.nan, np.float32)) for nm, a in items]

    h2, h3, soil = (specs(F["hist_2d"], defaults.HISTLIST_2D), specs(F["hist_3d"], defaults.HISTLIST_3D),
                    specs(F["soil"], defaults.HISTLIST_SOIL))
    hgt = np.full((1, nM), np.nan, np.float32)
    ust = np.full((nz, nU), np.nan, np.float32)
    vst = np.full((nz, nV), np.nan, np.float32)
    host.interp_data(rg, cfg, diag=[], hist_2d=h2, hist_3d=h3, soil=soil, ter=F["ter"], hgt=hgt, u_stag=ust, v_stag=vst, nz=nz)
    got = {s.name: s.dst for s in h2 + h3 + soil}
    got["HGT"], got["U"], got["V"] = hgt, ust, vst
    for nm, w in want.items():
        if nm.startswith("uReconstruct"):
            continue
        g = got[nm]
        assert not np.isnan(g).any(), nm
        if nm in {"xland", "tslb", "smois", "sh2o"}:
            assert np.array_equal(g, w), nm
        else:
            check.assert_field_close(g, w, nm)
    # a global mesh covers every target point: nothing is unmapped (snow: every destination cell fully covered)
    assert (got["theta"] != 0).all()
    rg.close()


@pytest.mark.parametrize("name", ["mid", "c1"])
def test_interp_data_workload_against_oracle(engine_lib, orc, host, name):
    """The bench's own code path (workload.make / run_interp, device buffers, stock var-lists, wind chain
    with fused rotation) against the oracle, every output field:
    mid -- the 12-km miniature of the CONUS case (153 k cells, 60 levels -> 450 x 265, Lambert, rotation);
    c1  -- BASELINE.json configs[0] at its full size (40,962-cell global mesh, 55 levels -> 1-degree lat-lon)."""
    import torch

    from mpassit_b200 import lib as l
    from mpassit_b200 import workload
    from mpassit_b200.regrid import Regridder

    wl = workload.make(name)
    rg = Regridder(device=0)
    workload.load_geometry(rg, wl)
    F = workload.make_fields(wl, device="cuda:0")
    workload.run_interp(rg, wl, F["dev"], l.DEVICE)
    rg.synchronize()
    fields = {g: [(s.name, s.src.cpu().numpy()) for s in F["dev"][g]] for g in ("diag", "hist_2d", "hist_3d", "soil")}
    fields["ter"] = F["dev"]["ter"].cpu().numpy()
    want = H.oracle_interp(orc, wl.mesh, wl.grids, fields, wl.cosa, wl.sina, lc=wl.cosa is not None,
                           periodic=not wl.cfg.is_regional)
    got = {s.name: s.dst.cpu().numpy() for g in ("diag", "hist_2d", "hist_3d", "soil") for s in F["dev"][g]}
    got["HGT"], got["U"], got["V"] = (F["dev"][k].cpu().numpy() for k in ("hgt", "u_stag", "v_stag"))
    exact = {"xland", "tslb", "smois", "sh2o"}
    checked = 0
    for nm, w in want.items():
        if nm.startswith("uReconstruct"):
            continue
        g = got[nm].reshape(w.shape)
        if nm in exact:
            assert np.array_equal(g, w), nm
        else:
            check.assert_field_close(g, w, nm)
        checked += 1
    assert checked >= (40 if name == "mid" else 25)
    rg.close()


def test_cuda_graph_replay_of_a_pass(engine_lib, host):
    """mprg_capture_begin / _end record a device-buffer interp_data pass; replaying the graph after the sources
    changed gives exactly what an eager pass gives, and calls that cannot be captured are refused."""
    import torch

    from mpassit_b200 import lib as l
    from mpassit_b200 import workload
    from mpassit_b200.regrid import Regridder

    wl = workload.make("mini")
    rg = Regridder(device=0)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        rg.use_torch_stream()
        workload.load_geometry(rg, wl)
        F = workload.make_fields(wl, device="cuda:0")
        workload.run_interp(rg, wl, F["dev"], l.DEVICE)          # builds (memoises) every route, sizes scratch
        rg.synchronize()
        n0 = rg.kernel_launches
        rg.capture_begin()
        workload.run_interp(rg, wl, F["dev"], l.DEVICE)
        with pytest.raises(l.MprgError):                          # host buffers cannot be captured
            r = rg.store(l.BILINEAR, l.SRC_MESH_ELEMENT, l.CENTER)
            rg.apply(r, [np.zeros((wl.mesh.nCells, 3), np.float32)], [np.zeros((3, wl.n_mass), np.float32)], nlev=[3])
        g = rg.capture_end()
        assert rg.kernel_launches == n0                           # recorded, not run
        # new data in the same buffers
        for grp in ("diag", "hist_2d", "hist_3d", "soil"):
            for s in F["dev"][grp]:
                s.src.mul_(1.25).add_(0.5) if grp != "soil" and s.name != "xland" else None
                s.dst.fill_(float("nan"))
        rg.graph_launch(g)
        rg.synchronize()
        assert rg.kernel_launches > n0
        got = {s.name: s.dst.clone() for grp in ("diag", "hist_2d", "hist_3d", "soil") for s in F["dev"][grp]}
        gu, gv = F["dev"]["u_stag"].clone(), F["dev"]["v_stag"].clone()
        workload.run_interp(rg, wl, F["dev"], l.DEVICE)           # eager pass on the same data
        rg.synchronize()
        for grp in ("diag", "hist_2d", "hist_3d", "soil"):
            for s in F["dev"][grp]:
                if s.name.startswith("uReconstruct"):
                    continue
                assert torch.equal(got[s.name], s.dst), s.name
        assert torch.equal(gu, F["dev"]["u_stag"]) and torch.equal(gv, F["dev"]["v_stag"])
        rg.graph_release(g)
        rg.close()
