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
# oracle-side mirror of interp_data (interp.F90:92-465) for end-to-end parity: oracle/interp_oracle.py
# --------------------------------------------------------------------------
def oracle_interp(orc, mesh, grids, fields, cosa, sina, wrf_mod_vars=True, lc=True, periodic=False, rows=None):
    """grids: dict stagger-name -> (lat, lon) for M, U, V, CORNER.  fields: dict with keys diag, hist_2d,
    hist_3d, soil (lists of (name, array [n][nlev] or [n])) and 'ter'.  Returns dict name -> fp32 [nlev][n].
    periodic: the target is a global grid (is_regional = .false.: ESMF_GridCreate1PeriDim + monopoles)."""
    from oracle import interp_oracle

    return interp_oracle.interp_data(mesh, grids, fields, cosa, sina, wrf_mod_vars=wrf_mod_vars, lc=lc,
                                     periodic=periodic, rows=rows)
