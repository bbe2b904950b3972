"""Oracle-side mirror of the reference's hot path, interp_data (interp.F90:92-465), on top of the C
restatement (oracle/mpassit_oracle.c): every RegridStore / Regrid pair of interp_diag_data and
interp_hist_data in the reference's order, the `method` carry-over (interp.F90:203-204, 370, 420,
436-443), rotate_winds_cgrid at the mass points (:291-293, v' from the rotated u') and the
centre -> EDGE1 / EDGE2 grid-to-grid regrid of the winds (:295-328).

TEST INFRASTRUCTURE ONLY (see the header of mpassit_oracle.c; PARITY UNPINNED).  Callers: tests/,
__graft_entry__.smoke(), bench.py's `cpu_baseline` leg and `--impl reference` arm -- as the checker
and as the timed CPU baseline, never on the product path.
"""
from __future__ import annotations

import time

import numpy as np

from . import oracle as orc

CONS_VARS = ("snow", "snowh")                                # input_data.F90:840
NSTD_VARS = ("ivgtyp", "isltyp", "xland", "landmask")        # input_data.F90:841
WIND_U, WIND_V = "uReconstructZonal", "uReconstructMeridional"


def geometry(mesh):
    """Mesh arrays -> (cell xyz, vertex xyz, dual triangles) the way the reference builds its ESMF_Mesh
    (model_grid.F90:446-497: degrees, > 180 -> -360 wrap) and ESMF turns SPH_DEG into Cartesian."""
    lo, la = orc.mesh_rad_to_deg(mesh.lonCell, mesh.latCell)
    cxyz = orc.sph_deg_to_cart(lo, la)
    lov, lav = orc.mesh_rad_to_deg(mesh.lonVertex, mesh.latVertex)
    vxyz = orc.sph_deg_to_cart(lov, lav)
    tri = orc.dual_triangles(mesh.verticesOnCell, mesh.nVertices)
    return cxyz, vxyz, tri


def interp_data(mesh, grids, fields, cosa=None, sina=None, wrf_mod_vars=True, lc=True, periodic=False,
                rows=None, keep=True, tiled=False, timing=None):
    """One interp_data pass.

    grids   : stagger name -> (lat, lon) [nj][ni] degrees for M, U, V, CORNER
    fields  : dict with keys diag, hist_2d, hist_3d, soil -> lists of (name, array [n][nlev] | [n]) and 'ter'
    rows    : (j0, j1) block of CENTER rows to regrid (EDGE2 / CORNER get j1 + 1); None = the whole grid
    periodic: is_regional=.false. -> ESMF_GridCreate1PeriDim (model_grid.F90:684-696): the centre->edge
              source grid wraps in i
    keep    : return the outputs (name -> fp32 [nlev][n]); False = timing only (outputs dropped as produced)
    timing  : optional dict, filled with weights_s / apply_s / units
    """
    t0 = time.perf_counter()
    cxyz, vxyz, tri = geometry(mesh)
    lat, lon = grids["M"]
    nj, ni = lat.shape
    j0, j1 = (0, nj) if rows is None else rows
    njb = j1 - j0
    dxyz = orc.sph_deg_to_cart(lon[j0:j1], lat[j0:j1])
    n = dxyz.shape[0]
    voc = mesh.verticesOnCell
    diag, h2, h3, soil = (fields.get(k, []) for k in ("diag", "hist_2d", "hist_3d", "soil"))
    have_u = wrf_mod_vars and any(nm == WIND_U for nm, _ in h3)
    have_v = wrf_mod_vars and any(nm == WIND_V for nm, _ in h3)

    # ---- weight generation: one matrix per (method, source location, stagger), as the engine memoises them;
    #      the reference regenerates the bilinear one 7 times (interp.F90:123-366)
    e, c, w = orc.bilinear(cxyz, tri, voc, dxyz)
    bil = orc.ell_to_csr(e >= 0, c, w)
    cons = nst = node = ucsr = vcsr = None
    if any(nm in CONS_VARS for nm, _ in h2):
        clat, clon = grids["CORNER"]
        cor = orc.sph_deg_to_cart(clon[j0:j1 + 1], clat[j0:j1 + 1]).reshape(njb + 1, clat.shape[1], 3)
        cons = orc.conserve(cxyz, vxyz, voc, cor)
    if any(nm in NSTD_VARS for nm, _ in h2):
        nst = orc.nearest_to_csr(orc.nearest(cxyz, dxyz))
    if any(nm == "vorticity" for nm, _ in h3):
        en, cn, wn = orc.bilinear_node(cxyz, vxyz, voc, dxyz)
        node = orc.ell_to_csr(en >= 0, cn, wn)
    if have_u or have_v:
        # the grid-to-grid source is the CENTER grid itself; a row block interpolates from its own rows only
        # when it is the whole grid, so blocks carry one halo row either side (clipped to the grid)
        h0, h1 = max(j0 - 1, 0), min(j1 + 1, nj)
        sx = orc.sph_deg_to_cart(lon[h0:h1], lat[h0:h1]).reshape(h1 - h0, ni, 3)
        topo = 0
        if periodic:    # ESMF_GridCreate1PeriDim + MONOPOLE; a block carries a cap only when it holds the grid's end row
            topo = orc.TOPO_PERI | (orc.TOPO_SPOLE if h0 == 0 else 0) | (orc.TOPO_NPOLE if h1 == nj else 0)
        if have_u:
            ulat, ulon = grids["U"]
            eu, cu, wu = orc.bilinear_quadgrid(sx, orc.sph_deg_to_cart(ulon[j0:j1], ulat[j0:j1]), topo=topo)
            ucsr = orc.quadgrid_csr(ni, h1 - h0, eu, cu, wu, topo)
        if have_v:
            vlat, vlon = grids["V"]
            jv1 = j1 + 1 if j1 == nj else j1
            ev, cv, wv = orc.bilinear_quadgrid(sx, orc.sph_deg_to_cart(vlon[j0:jv1], vlat[j0:jv1]), topo=topo)
            vcsr = orc.quadgrid_csr(ni, h1 - h0, ev, cv, wv, topo)
    t1 = time.perf_counter()

    # ---- weight application
    out = {}
    units = 0

    def app(csr, arr, dt=np.float32, name=None):
        nonlocal units
        a = np.ascontiguousarray(arr)
        r = orc.apply(*csr, a, dt, tiled=tiled and dt == np.float32)
        units += r.size
        if keep and name is not None:
            out[name] = r
        return r

    for nm, a in diag:                                      # interp_diag_data, interp.F90:107-141
        app(bil, a, name=nm)
    if keep and "u10" in out and "v10" in out and lc:        # :138-139
        orc.rotate_winds(out["u10"], out["v10"], cosa[j0:j1].reshape(-1), sina[j0:j1].reshape(-1))
    method = None
    if any(nm not in CONS_VARS + NSTD_VARS for nm, _ in h2):
        method = "bil"                                      # interp.F90:203-204
    for nm, a in h2:
        if nm not in CONS_VARS + NSTD_VARS:
            app(bil, a, name=nm)
    if fields.get("ter") is not None:                       # hgt, :226-238
        app(bil, fields["ter"], name="HGT")
    u = v = None
    for nm, a in h3:
        if wrf_mod_vars and nm == WIND_U:
            u = a
        elif wrf_mod_vars and nm == WIND_V:
            v = a
        elif nm == "vorticity":
            app(node, a, name=nm)                           # :350-366
        else:
            app(bil, a, name=nm)                            # 3d_nz :240-254, 3d_nzp1 :331-347
    if u is not None or v is not None:                      # winds, :256-328, R8 until the write
        if rows is None:
            hb, hcsr = (0, nj), bil
        else:                                               # mass-point winds on the halo rows of the block
            hb = (max(j0 - 1, 0), min(j1 + 1, nj))
            hx = orc.sph_deg_to_cart(lon[hb[0]:hb[1]], lat[hb[0]:hb[1]])
            eh, ch, wh = orc.bilinear(cxyz, tri, voc, hx)
            hcsr = orc.ell_to_csr(eh >= 0, ch, wh)
        um = orc.apply(*hcsr, np.ascontiguousarray(u), np.float64) if u is not None else None
        vm = orc.apply(*hcsr, np.ascontiguousarray(v), np.float64) if v is not None else None
        units += (0 if um is None else wl_rows(um, hb, (j0, j1), ni)) + (0 if vm is None else wl_rows(vm, hb, (j0, j1), ni))
        if um is not None and vm is not None and lc:        # :291-293
            orc.rotate_winds(um, vm, cosa[hb[0]:hb[1]].reshape(-1), sina[hb[0]:hb[1]].reshape(-1))
        if keep:
            if um is not None:
                out[WIND_U] = um.reshape(um.shape[0], hb[1] - hb[0], ni)[:, j0 - hb[0]:j1 - hb[0]].reshape(um.shape[0], -1).astype(np.float32)
            if vm is not None:
                out[WIND_V] = vm.reshape(vm.shape[0], hb[1] - hb[0], ni)[:, j0 - hb[0]:j1 - hb[0]].reshape(vm.shape[0], -1).astype(np.float32)
        if um is not None:
            r = orc.apply_planes(*ucsr, um)
            units += r.size
            if keep:
                out["U"] = r.astype(np.float32)
        if vm is not None:
            r = orc.apply_planes(*vcsr, vm)
            units += r.size
            if keep:
                out["V"] = r.astype(np.float32)
    if cons is not None:                                    # 2d_cons, :368-416
        method = "cons"
        for nm, a in h2:
            if nm in CONS_VARS:
                app(cons, a, name=nm)
    if nst is not None:                                     # 2d_nstd, :418-434
        method = "nstd"
        for nm, a in h2:
            if nm in NSTD_VARS:
                app(nst, a, name=nm)
    soil_csr = {"bil": bil, None: bil, "cons": cons, "nstd": nst}[method]   # :436-447 whatever `method` holds
    for nm, a in soil:
        app(soil_csr, a, name=nm)
    t2 = time.perf_counter()
    if timing is not None:
        timing.update(weights_s=t1 - t0, apply_s=t2 - t1, units=units, rows=njb, of_rows=nj)
    return out


def wl_rows(arr, halo, own, ni):
    """Output values of a mass-point wind field that belong to the block's own rows."""
    return arr.shape[0] * (own[1] - own[0]) * ni


def wind_chain(mesh, grids, u, v, cosa, sina):
    """The reference's wind chain alone, in R8 (interp.F90:256-328): cell-centre (u, v) [nCells][nlev] -> mass points
    (bilinear), rotate_winds_cgrid there, -> EDGE1 / EDGE2 (grid-to-grid bilinear).  Returns {"U", "V"} fp64
    [nlev][points]: what the composed wind routes of the engine must reproduce."""
    cxyz, vxyz, tri = geometry(mesh)
    lat, lon = grids["M"]
    nj, ni = lat.shape
    dxyz = orc.sph_deg_to_cart(lon, lat)
    e, c, w = orc.bilinear(cxyz, tri, mesh.verticesOnCell, dxyz)
    bil = orc.ell_to_csr(e >= 0, c, w)
    um = orc.apply(*bil, np.ascontiguousarray(u, np.float64), np.float64)
    vm = orc.apply(*bil, np.ascontiguousarray(v, np.float64), np.float64)
    orc.rotate_winds(um, vm, np.asarray(cosa, np.float64).reshape(-1), np.asarray(sina, np.float64).reshape(-1))
    sx = dxyz.reshape(nj, ni, 3)
    out = {}
    for key, src in (("U", um), ("V", vm)):
        slat, slon = grids[key]
        es, cs, ws = orc.bilinear_quadgrid(sx, orc.sph_deg_to_cart(slon, slat), topo=0)
        out[key] = np.asarray(orc.apply_planes(*orc.quadgrid_csr(ni, nj, es, cs, ws, 0), src), np.float64)
    return out
