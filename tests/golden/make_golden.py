#!/usr/bin/env python
"""Generate tests/golden/lc_small.npz -- golden vectors for the interpolation path.

The reference ships no fixtures and cannot run here (Fortran + ESMF + MPI + NetCDF are all
missing), so these vectors come from a SECOND, independent restatement of the
same ESMF rules, written in plain numpy with different algorithms from oracle/mpassit_oracle.c:

  nearest     dense distance matrix + argmin             (oracle: kd-tree, running minimum)
  bilinear    np.linalg.solve on [e1 e2 -p] per pair     (oracle: closed-form triple products)
  quad grid   vectorised Newton over all (point, quad)   (oracle: per-point search + scalar Newton)
  conserve    Sutherland-Hodgman in the GNOMONIC plane of the destination cell (great circles are
              straight lines there), areas by L'Huilier's theorem from arc lengths
              (oracle: clipping by great-circle planes in 3-D, areas by the atan2 triple-product form)
  node        fan triangles of the Voronoi polygon, same solve as bilinear
  apply       dense fp64 matrix product, one rounding to fp32
  rotation    closed form with the sequential u' -> v' rule (interp.F90:739-745)

Two implementations that agree pin each other (tests/test_golden.py runs the C oracle against this
file; tests/test_gpu_golden.py runs the CUDA engine against it through the C ABI).  Neither is ESMF:
parity with the real reference stays UNPINNED until ESMF_RegridWeightGen dumps can be produced.

    python tests/golden/make_golden.py        # rewrites lc_small.npz (deterministic)
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from mpassit_b200 import synth  # noqa: E402  (inputs only: mesh generator, field recipes)
from oracle import proj_oracle  # noqa: E402  (inputs only: Lambert coordinates of the target grid)

TOL = 1e-10  # ESMF's parametric tolerance for point-in-element tests


# ------------------------------------------------------------------ coordinates
def mesh_deg(lon_rad, lat_rad):
    """model_grid.F90:450-454,464-468."""
    pi = 4.0 * np.arctan(1.0)
    lo = lon_rad * 180.0 / pi
    lo = np.where(lo > 180.0, lo - 360.0, lo)
    return lo, lat_rad * 180.0 / pi


def cart(lon_deg, lat_deg):
    d2r = 3.141592653589793238 / 180.0
    th, ph = np.asarray(lon_deg).reshape(-1) * d2r, (90.0 - np.asarray(lat_deg).reshape(-1)) * d2r
    return np.stack([np.cos(th) * np.sin(ph), np.sin(th) * np.sin(ph), np.cos(ph)], 1)


# ------------------------------------------------------------------ nearest
def nearest(src, dst):
    d2 = ((dst[:, None, :] - src[None, :, :]) ** 2).sum(-1)
    return d2.argmin(1).astype(np.int32)  # first minimum == smallest source id


# ------------------------------------------------------------------ bilinear on the dual mesh
def dual_triangles(voc, nV):
    tri = np.full((nV, 3), -1, np.int32)
    for v in range(nV):
        cells = np.nonzero((voc == v + 1).any(1))[0]
        if cells.size == 3:
            tri[v] = cells  # ascending cell ids
    return tri


def solve_tri(v0, v1, v2, p):
    """v0 + a (v1-v0) + b (v2-v0) = s p for broadcast arrays [...,3] -> a, b, s (nan where singular)."""
    v0, v1, v2, p = np.broadcast_arrays(v0, v1, v2, p)
    M = np.stack([v1 - v0, v2 - v0, -p], -1)
    ok = np.abs(np.linalg.det(M)) > 0
    M = np.where(ok[..., None, None], M, np.eye(3))
    x = np.linalg.solve(M, (-v0)[..., None])[..., 0]
    x[~ok] = np.nan
    return x[..., 0], x[..., 1], x[..., 2]


def bilinear(cxyz, tri, dst):
    valid = np.nonzero(tri[:, 0] >= 0)[0]
    t = tri[valid]
    a, b, s = solve_tri(cxyz[t[:, 0]][None], cxyz[t[:, 1]][None], cxyz[t[:, 2]][None], dst[:, None, :])
    with np.errstate(invalid="ignore"):
        acc = (a >= -TOL) & (b >= -TOL) & (a + b <= 1 + TOL) & (s > 0)
    n = dst.shape[0]
    elem = np.full(n, -1, np.int32)
    col = np.zeros((n, 3), np.int32)
    w = np.zeros((n, 3))
    for i in range(n):
        k = np.nonzero(acc[i])[0]
        if k.size:
            k = k[0]  # valid[] ascending: smallest dual-element id wins
            elem[i] = valid[k]
            col[i] = t[k]
            w[i] = (1.0 - a[i, k] - b[i, k], a[i, k], b[i, k])
    return elem, col, w


# ------------------------------------------------------------------ grid -> grid (centre quads)
def quadgrid(sxyz_grid, dst):
    nj, ni = sxyz_grid.shape[:2]
    q0 = sxyz_grid[:-1, :-1].reshape(-1, 3)
    q1 = sxyz_grid[:-1, 1:].reshape(-1, 3)
    q2 = sxyz_grid[1:, 1:].reshape(-1, 3)
    q3 = sxyz_grid[1:, :-1].reshape(-1, 3)
    A, B, Cc = (q0 - q1 + q2 - q3)[None], (q1 - q0)[None], (q3 - q0)[None]
    P = dst[:, None, :]
    X = np.zeros((dst.shape[0], q0.shape[0], 3))
    live = np.ones(X.shape[:2], bool)
    with np.errstate(all="ignore"):
        for _ in range(100):
            u, v, t = X[..., 0:1], X[..., 1:2], X[..., 2:3]
            F = u * v * A + u * B + v * Cc - t * P + q0[None]
            live &= ~((F ** 2).sum(-1) < 1e-20)
            if not live.any():
                break
            J = np.stack([A * v + B, A * u + Cc, -np.broadcast_to(P, F.shape)], -1)
            good = live & (np.abs(np.linalg.det(J)) > 0) & np.isfinite(J).all((-1, -2))
            Jg = np.where(good[..., None, None], J, np.eye(3))
            step = np.linalg.solve(Jg, F[..., None])[..., 0]
            X = np.where(good[..., None], X - step, X)
            live &= good
        u, v, t = X[..., 0], X[..., 1], X[..., 2]
        acc = (t > 0) & (u >= -TOL) & (u <= 1 + TOL) & (v >= -TOL) & (v <= 1 + TOL) & np.isfinite(X).all(-1)
        F = (u * v)[..., None] * A + u[..., None] * B + v[..., None] * Cc - t[..., None] * P + q0[None]
        acc &= (F ** 2).sum(-1) < 1e-18  # converged
    n = dst.shape[0]
    elem = np.full(n, -1, np.int64)
    col = np.zeros((n, 4), np.int32)
    w = np.zeros((n, 4))
    for i in range(n):
        k = np.nonzero(acc[i])[0]
        if k.size:
            k = k[0]
            jq, iq = divmod(k, ni - 1)
            elem[i] = k
            col[i] = (jq * ni + iq, jq * ni + iq + 1, (jq + 1) * ni + iq + 1, (jq + 1) * ni + iq)
            uu, vv = u[i, k], v[i, k]
            w[i] = ((1 - uu) * (1 - vv), uu * (1 - vv), uu * vv, (1 - uu) * vv)
    return elem, col, w


# ------------------------------------------------------------------ conservative
def arc(a, b):
    return 2.0 * np.arcsin(min(1.0, 0.5 * np.linalg.norm(a - b)))


def lhuilier(a, b, c):
    A, B, Cc = arc(b, c), arc(c, a), arc(a, b)
    s = 0.5 * (A + B + Cc)
    t = np.tan(0.5 * s) * np.tan(0.5 * (s - A)) * np.tan(0.5 * (s - B)) * np.tan(0.5 * (s - Cc))
    return 4.0 * np.arctan(np.sqrt(max(t, 0.0)))


def poly_area(v):
    return sum(lhuilier(v[0], v[k], v[k + 1]) for k in range(1, len(v) - 1))


def gnomonic_frame(c):
    c = c / np.linalg.norm(c)
    e = np.cross([0.0, 0.0, 1.0], c)
    e /= np.linalg.norm(e)
    return c, e, np.cross(c, e)


def clip2d(poly, a, b):
    """Keep the part of `poly` on the left of the directed line a -> b (2-D Sutherland-Hodgman)."""
    out = []
    side = lambda p: (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])
    for k in range(len(poly)):
        s, e = poly[k - 1], poly[k]
        ds, de = side(s), side(e)
        if (de >= 0) != (ds >= 0):
            tau = ds / (ds - de)
            out.append(s + tau * (e - s))
        if de >= 0:
            out.append(e)
    return out


def ccw2d(p):
    p = np.asarray(p)
    x, y = p[:, 0], p[:, 1]
    return p if (x * np.roll(y, -1) - np.roll(x, -1) * y).sum() >= 0 else p[::-1]


def conserve(vxyz, voc, corner_grid):
    njc, nic = corner_grid.shape[:2]
    nj, ni = njc - 1, nic - 1
    polys = [vxyz[row[row > 0] - 1] for row in voc]
    cen_s = np.array([p.mean(0) / np.linalg.norm(p.mean(0)) for p in polys])
    rad_s = np.array([np.linalg.norm(p - c, axis=1).max() for p, c in zip(polys, cen_s)])
    rowptr, col, w = [0], [], []
    for j in range(nj):
        for i in range(ni):
            dq = np.array([corner_grid[j, i], corner_grid[j, i + 1], corner_grid[j + 1, i + 1], corner_grid[j + 1, i]])
            c, e1, e2 = gnomonic_frame(dq.sum(0))
            proj = lambda v: np.stack([v @ e1, v @ e2], -1) / (v @ c)[..., None]
            lift = lambda q: (lambda x: x / np.linalg.norm(x))(c + q[0] * e1 + q[1] * e2)
            d2 = ccw2d(proj(dq))
            darea = poly_area([lift(q) for q in d2])
            rd = np.linalg.norm(dq - c, axis=1).max()
            near = np.nonzero(np.linalg.norm(cen_s - c, axis=1) <= (rd + rad_s) * 1.01 + 1e-9)[0]
            for s in near:  # ascending source ids
                p2 = list(ccw2d(proj(polys[s])))
                for k in range(4):
                    p2 = clip2d(p2, d2[k], d2[(k + 1) % 4])
                    if len(p2) < 3:
                        break
                if len(p2) < 3:
                    continue
                ar = poly_area([lift(q) for q in p2])
                if ar > 0.0:
                    col.append(s)
                    w.append(ar / darea)
            rowptr.append(len(col))
    return np.array(rowptr, np.int32), np.array(col, np.int32), np.array(w)


# ------------------------------------------------------------------ node-based bilinear
def bilinear_node(vxyz, voc, dst):
    n = dst.shape[0]
    elem = np.full(n, -1, np.int32)
    col = np.zeros((n, 3), np.int32)
    w = np.zeros((n, 3))
    fans = []  # (cell, v0, vk, vk+1) in cell order, then fan order
    for c, row in enumerate(voc):
        vs = row[row > 0] - 1
        for k in range(1, len(vs) - 1):
            fans.append((c, vs[0], vs[k], vs[k + 1]))
    fans = np.array(fans)
    a, b, s = solve_tri(vxyz[fans[:, 1]][None], vxyz[fans[:, 2]][None], vxyz[fans[:, 3]][None], dst[:, None, :])
    with np.errstate(invalid="ignore"):
        acc = (a >= -TOL) & (b >= -TOL) & (a + b <= 1 + TOL) & (s > 0)
    for i in range(n):
        k = np.nonzero(acc[i])[0]
        if k.size:
            k = k[0]
            elem[i] = fans[k, 0]
            col[i] = fans[k, 1:]
            w[i] = (1.0 - a[i, k] - b[i, k], a[i, k], b[i, k])
    return elem, col, w


# ------------------------------------------------------------------ apply / rotation
def ell_csr(mask, col, w):
    k = col.shape[1]
    rowptr = np.zeros(mask.size + 1, np.int32)
    np.cumsum(np.where(mask, k, 0), out=rowptr[1:])
    return rowptr, col[mask].reshape(-1).astype(np.int32), w[mask].reshape(-1)


def apply_dense(rowptr, col, w, src, nsrc):
    """src [nsrc][nlev] -> fp32 [nlev][ndst] via a dense fp64 matrix (zero rows stay 0)."""
    W = np.zeros((rowptr.size - 1, nsrc))
    rows = np.repeat(np.arange(rowptr.size - 1), np.diff(rowptr))
    np.add.at(W, (rows, col), w)
    return (W @ src.astype(np.float64)).T.astype(np.float32)


def rotate(u, v, cosa, sina):
    tana = sina / cosa
    u2 = (u + v * tana) / (cosa + sina * tana)
    v2 = (v - u2 * sina) / cosa
    return u2, v2


def build_case():
    mesh = synth.regional_delaunay_mesh(420, extent_x_m=560e3, extent_y_m=400e3, seed=5, lloyd_iters=2)
    kw = dict(nx=17, ny=13, dx=40000.0, ref_lat=38.5, ref_lon=-97.5, truelat1=38.5, truelat2=38.5, stand_lon=-97.5)
    grids = {s: proj_oracle.lc_grid(stagger=s, **kw) for s in ("M", "U", "V", "CORNER")}
    return mesh, grids


def main():
    mesh, grids = build_case()
    voc = mesh.verticesOnCell
    lo, la = mesh_deg(mesh.lonCell, mesh.latCell)
    cxyz = cart(lo, la)
    lov, lav = mesh_deg(mesh.lonVertex, mesh.latVertex)
    vxyz = cart(lov, lav)
    latM, lonM = grids["M"]
    dM = cart(lonM, latM)
    out = dict(lonCell=mesh.lonCell, latCell=mesh.latCell, lonVertex=mesh.lonVertex, latVertex=mesh.latVertex,
               verticesOnCell=voc.astype(np.int32))
    for s in grids:
        out[f"lat_{s}"], out[f"lon_{s}"] = grids[s]
    cosa, sina = proj_oracle.rotang(latM, lonM)
    out["cosa"], out["sina"] = cosa, sina

    out["nearest_idx"] = nearest(cxyz, dM)
    tri = dual_triangles(voc, mesh.nVertices)
    out["tri"] = tri
    e, c, w = bilinear(cxyz, tri, dM)
    out["bil_elem"], out["bil_col"], out["bil_w"] = e, c, w
    sgrid = dM.reshape(*latM.shape, 3)
    for s in ("U", "V"):
        eq, cq, wq = quadgrid(sgrid, cart(grids[s][1], grids[s][0]))
        out[f"quad{s}_elem"], out[f"quad{s}_col"], out[f"quad{s}_w"] = eq, cq, wq
    clat, clon = grids["CORNER"]
    rp, cc, ww = conserve(vxyz, voc, cart(clon, clat).reshape(*clat.shape, 3))
    out["cons_rowptr"], out["cons_col"], out["cons_w"] = rp, cc, ww
    en, cn, wn = bilinear_node(vxyz, voc, dM)
    out["node_elem"], out["node_col"], out["node_w"] = en, cn, wn

    # fields and their regridded values
    nlev = 5
    theta = synth.smooth_field(mesh.lonCell, mesh.latCell, nlev, seed=3)
    xland = synth.integer_field(mesh.nCells, 2, seed=9)
    snow = synth.patchy_field(mesh.lonCell, mesh.latCell)
    uu = synth.smooth_field(mesh.lonCell, mesh.latCell, nlev, seed=4) - 280.0
    vv = synth.smooth_field(mesh.lonCell, mesh.latCell, nlev, seed=6) - 285.0
    vort = synth.smooth_field(mesh.lonVertex, mesh.latVertex, nlev, seed=8)
    out.update(src_theta=theta, src_xland=xland, src_snow=snow, src_u=uu, src_v=vv, src_vort=vort)
    bil = ell_csr(e >= 0, c, w)
    out["dst_theta"] = apply_dense(*bil, theta, mesh.nCells)
    n = dM.shape[0]
    out["dst_xland"] = apply_dense(np.arange(n + 1, dtype=np.int32), out["nearest_idx"], np.ones(n), xland[:, None], mesh.nCells)
    out["dst_snow"] = apply_dense(rp, cc, ww, snow[:, None], mesh.nCells)
    out["dst_vort"] = apply_dense(*ell_csr(en >= 0, cn, wn), vort, mesh.nVertices)
    # winds: mass-point bilinear in fp64, rotation, then centre -> edge stagger (interp.F90:259-325)
    W = np.zeros((n, mesh.nCells))
    rows = np.repeat(np.arange(n), np.diff(bil[0]))
    np.add.at(W, (rows, bil[1]), bil[2])
    um, vm = (W @ uu.astype(np.float64)).T, (W @ vv.astype(np.float64)).T
    um, vm = rotate(um, vm, cosa.reshape(-1)[None], sina.reshape(-1)[None])
    out["dst_umass"], out["dst_vmass"] = um, vm
    for s, f in (("U", um), ("V", vm)):
        q = ell_csr(out[f"quad{s}_elem"] >= 0, out[f"quad{s}_col"], out[f"quad{s}_w"])
        Wq = np.zeros((q[0].size - 1, n))
        np.add.at(Wq, (np.repeat(np.arange(q[0].size - 1), np.diff(q[0])), q[1]), q[2])
        out[f"dst_{s}"] = (Wq @ f.T).T.astype(np.float32)
    path = os.path.join(HERE, "lc_small.npz")
    np.savez_compressed(path, **out)
    unm = int((e < 0).sum())
    print(f"wrote {path}: {mesh.nCells} cells, {mesh.nVertices} vertices, target {latM.shape[::-1]}, "
          f"bilinear unmapped {unm}/{n}, conserve nnz {cc.size}, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
