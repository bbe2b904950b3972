"""Shared test helpers: small synthetic cases + oracle-side geometry."""
import numpy as np

from mpassit_b200 import synth


def latlon_grid(ni, nj, lon0=-180.0, lon1=180.0, lat0=-90.0, lat1=90.0):
    """Cell-centred regular lat-lon target [nj][ni] in degrees."""
    dlon, dlat = (lon1 - lon0) / ni, (lat1 - lat0) / nj
    lon = lon0 + (np.arange(ni) + 0.5) * dlon
    lat = lat0 + (np.arange(nj) + 0.5) * dlat
    LO, LA = np.meshgrid(lon, lat)
    return np.ascontiguousarray(LO), np.ascontiguousarray(LA)


def oracle_geometry(orc, mesh):
    lo, la = orc.mesh_rad_to_deg(mesh.lonCell, mesh.latCell)
    cxyz = orc.sph_deg_to_cart(lo, la)
    lov, lav = orc.mesh_rad_to_deg(mesh.lonVertex, mesh.latVertex)
    vxyz = orc.sph_deg_to_cart(lov, lav)
    tri = orc.dual_triangles(mesh.verticesOnCell, mesh.nVertices)
    return cxyz, vxyz, tri


def small_global(n=2562, jitter=0.15, seed=7):
    return synth.global_mesh(n, kind="icos", jitter=jitter, seed=seed)


def small_regional(n=3000, seed=11):
    return synth.regional_delaunay_mesh(n, seed=seed)


# --------------------------------------------------------------------------
# oracle-side mirror of interp_data (interp.F90:92-465) for end-to-end parity
# --------------------------------------------------------------------------
def oracle_interp(orc, mesh, grids, fields, cosa, sina, wrf_mod_vars=True, lc=True):
    """grids: dict stagger-name -> (lat, lon) for M, U, V, CORNER.  fields: dict with keys diag, hist_2d,
    hist_3d, soil (lists of (name, array [n][nlev] or [n])) and 'ter'.  Returns dict name -> fp32 [nlev][n]."""
    import numpy as np

    from mpassit_b200 import host as hostmod  # only for the class constants

    cxyz, vxyz, tri = oracle_geometry(orc, mesh)
    lat, lon = grids["M"]
    n = lat.size
    dxyz = orc.sph_deg_to_cart(lon, lat)
    e, c, w = orc.bilinear(cxyz, tri, mesh.verticesOnCell, dxyz)
    bil = orc.ell_to_csr(e >= 0, c, w)
    out = {}

    def app(csr, arr, dt=np.float32):
        return orc.apply(*csr, np.ascontiguousarray(arr), dt)

    names = {}
    for nm, a in fields.get("diag", []):
        out[nm] = app(bil, a)
        names[nm] = 1
    if "u10" in out and "v10" in out and lc:
        orc.rotate_winds(out["u10"], out["v10"], cosa.reshape(-1), sina.reshape(-1))
    method = None
    h2 = fields.get("hist_2d", [])
    cons_names = ("snow", "snowh")
    nstd_names = ("ivgtyp", "isltyp", "xland", "landmask")
    if any(nm not in cons_names + nstd_names for nm, _ in h2):
        method = "bil"
    for nm, a in h2:
        if nm not in cons_names + nstd_names:
            out[nm] = app(bil, a)
    if fields.get("ter") is not None:
        out["HGT"] = app(bil, fields["ter"])
    u = v = None
    for nm, a in fields.get("hist_3d", []):
        if wrf_mod_vars and nm == "uReconstructZonal":
            u = a
        elif wrf_mod_vars and nm == "uReconstructMeridional":
            v = a
        elif nm == "vorticity":
            en, cn, wn = orc.bilinear_node(cxyz, vxyz, mesh.verticesOnCell, dxyz)
            out[nm] = app(orc.ell_to_csr(en >= 0, cn, wn), a)
        else:
            out[nm] = app(bil, a)
    if u is not None or v is not None:
        um = app(bil, u, np.float64) if u is not None else None
        vm = app(bil, v, np.float64) if v is not None else None
        if um is not None and vm is not None and lc:
            orc.rotate_winds(um, vm, cosa.reshape(-1), sina.reshape(-1))
        sx = dxyz.reshape(lat.shape[0], lat.shape[1], 3)
        if um is not None:
            ulat, ulon = grids["U"]
            eu, cu, wu = orc.bilinear_quadgrid(sx, orc.sph_deg_to_cart(ulon, ulat))
            out["U"] = orc.apply_planes(*orc.ell_to_csr(eu >= 0, cu, wu), um).astype(np.float32)
        if vm is not None:
            vlat, vlon = grids["V"]
            ev, cv, wv = orc.bilinear_quadgrid(sx, orc.sph_deg_to_cart(vlon, vlat))
            out["V"] = orc.apply_planes(*orc.ell_to_csr(ev >= 0, cv, wv), vm).astype(np.float32)
    if any(nm in cons_names for nm, _ in h2):
        method = "cons"
        clat, clon = grids["CORNER"]
        cor = orc.sph_deg_to_cart(clon, clat).reshape(clat.shape[0], clat.shape[1], 3)
        cons = orc.conserve(cxyz, vxyz, mesh.verticesOnCell, cor)
        for nm, a in h2:
            if nm in cons_names:
                out[nm] = app(cons, a)
    nst = None
    if any(nm in nstd_names for nm, _ in h2):
        method = "nstd"
        nst = orc.nearest_to_csr(orc.nearest(cxyz, dxyz))
        for nm, a in h2:
            if nm in nstd_names:
                out[nm] = app(nst, a)
    soil_csr = {"bil": bil, None: bil, "cons": locals().get("cons"), "nstd": nst}[method]
    for nm, a in fields.get("soil", []):
        out[nm] = app(soil_csr, a)
    return out
