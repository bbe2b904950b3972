"""numpy restatement of the WRF-compatibility post-ops the reference applies while writing
(write_data.F90:1339-1432, `wrf_mod_vars`).  TEST INFRASTRUCTURE ONLY (checks the engine's fused
epilogues and mprg_post_*).  Parity unpinned (no reference run available here).

All arrays are [nlev][n] (C order of the reference's (i, j, lev))."""
import numpy as np


def t_minus_300(theta):
    """write_data.F90:1339-1345.  The guard `if (dum3d(i,j,1) < 10) continue` is a no-op in Fortran
    (`continue` does nothing), so EVERY point is shifted: unmapped points (0.0) become -300."""
    return (theta.astype(np.float64) - 300.0).astype(np.float32)


def z_c(zgrid_nzp1):
    """write_data.F90:1406-1412: Z_C(k-1) = 0.5 (PHB(k) + PHB(k-1)), k = 2..nzp1, on the regridded
    zgrid BEFORE the x9.81 scaling."""
    z = zgrid_nzp1.astype(np.float64)
    return (0.5 * (z[1:] + z[:-1])).astype(np.float32)


def phb(zgrid_nzp1):
    """write_data.F90:1417: PHB = zgrid * 9.81."""
    return (zgrid_nzp1.astype(np.float64) * 9.81).astype(np.float32)


def p_top(p_hyd):
    """write_data.F90:1364-1373: starts from maxval over the whole field, then the minimum of
    0.8 P_HYD(top level) over the columns whose top-level value is >= 10."""
    p = p_hyd.astype(np.float64)
    out = p.max()
    top = p[-1]
    sel = top >= 10.0
    if sel.any():
        out = min(out, (top[sel] * 0.80).min())
    return out
