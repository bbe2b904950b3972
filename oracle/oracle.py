"""ctypes loader for the CPU oracle (oracle/mpassit_oracle.c).

TEST INFRASTRUCTURE ONLY -- may be imported from tests/, from
__graft_entry__.smoke() and from bench.py's cpu_baseline / --impl reference legs.
The product package (mpassit_b200/) never imports this module.

PARITY UNPINNED: see the header of mpassit_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmpassit_oracle.so")
_lib = None

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mpassit_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_mesh_rad_to_deg.argtypes = [C.c_int64, _f64p, _f64p, _f64p, _f64p]
        L.orc_sph_deg_to_cart.argtypes = [C.c_int64, _f64p, _f64p, _f64p]
        L.orc_dual_triangles.argtypes = [C.c_int32, C.c_int32, C.c_int32, _i32p, _i32p]
        L.orc_dual_triangles.restype = C.c_int
        L.orc_nearest.argtypes = [C.c_int32, _f64p, C.c_int64, _f64p, _i32p, C.c_int]
        L.orc_nearest.restype = C.c_int
        L.orc_bilinear.argtypes = [C.c_int32, _f64p, C.c_int32, _i32p, C.c_int32, _i32p, C.c_int64, _f64p,
                                   _i32p, _i32p, _f64p, C.c_int]
        L.orc_bilinear.restype = C.c_int
        for name, tin, tout in (("orc_apply_f32_f32", _f32p, _f32p), ("orc_apply_f32_f64", _f32p, _f64p),
                                ("orc_apply_f64_f64", _f64p, _f64p), ("orc_apply_f64_f32", _f64p, _f32p),
                                ("orc_apply_f32_f32_tiled", _f32p, _f32p)):
            getattr(L, name).argtypes = [C.c_int64, _i32p, _i32p, _f64p, C.c_int32, tin, tout]
        L.orc_rotate_winds.argtypes = [C.c_int64, C.c_int32, _f64p, _f64p, _f64p, _f64p]
        L.orc_rotate_winds_f32.argtypes = [C.c_int64, C.c_int32, _f32p, _f32p, _f64p, _f64p]
        _i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
        L.orc_bilinear_quadgrid.argtypes = [C.c_int32, C.c_int32, _f64p, C.c_int64, _f64p, _i64p, _i32p, _f64p, C.c_int]
        L.orc_bilinear_quadgrid.restype = C.c_int
        L.orc_bilinear_quadgrid_topo.argtypes = [C.c_int32, C.c_int32, _f64p, C.c_int64, _f64p, _i64p, _i32p, _f64p, C.c_int, C.c_int]
        L.orc_bilinear_quadgrid_topo.restype = C.c_int
        L.orc_apply_planes_f64.argtypes = [C.c_int64, _i32p, _i32p, _f64p, C.c_int32, C.c_int64, _f64p, _f64p]
        L.orc_conserve.argtypes = [C.c_int32, _f64p, _f64p, C.c_int32, _i32p, C.c_int32, C.c_int32, _f64p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_conserve.restype = C.c_int
        L.orc_bilinear_node.argtypes = [C.c_int32, _f64p, _f64p, C.c_int32, _i32p, C.c_int64, _f64p, _i32p, _i32p,
                                        _f64p, C.c_int]
        L.orc_bilinear_node.restype = C.c_int
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


# ---------------------------------------------------------------- geometry
def mesh_rad_to_deg(lon_rad, lat_rad):
    lon_rad = np.ascontiguousarray(lon_rad, np.float64)
    lat_rad = np.ascontiguousarray(lat_rad, np.float64)
    lo = np.empty_like(lon_rad)
    la = np.empty_like(lat_rad)
    lib().orc_mesh_rad_to_deg(lon_rad.size, lon_rad, lat_rad, lo, la)
    return lo, la


def sph_deg_to_cart(lon_deg, lat_deg):
    lon_deg = np.ascontiguousarray(lon_deg, np.float64).reshape(-1)
    lat_deg = np.ascontiguousarray(lat_deg, np.float64).reshape(-1)
    xyz = np.empty((lon_deg.size, 3), np.float64)
    lib().orc_sph_deg_to_cart(lon_deg.size, lon_deg, lat_deg, xyz)
    return xyz


def dual_triangles(verticesOnCell, nVertices):
    voc = np.ascontiguousarray(verticesOnCell, np.int32)
    tri = np.empty((nVertices, 3), np.int32)
    rc = lib().orc_dual_triangles(voc.shape[0], nVertices, voc.shape[1], voc, tri)
    if rc != 0:
        raise ValueError(f"orc_dual_triangles rc={rc}")
    return tri


# ---------------------------------------------------------------- weights
def nearest(src_xyz, dst_xyz, brute=False):
    src_xyz = np.ascontiguousarray(src_xyz, np.float64)
    dst_xyz = np.ascontiguousarray(dst_xyz, np.float64)
    idx = np.empty(dst_xyz.shape[0], np.int32)
    rc = lib().orc_nearest(src_xyz.shape[0], src_xyz, dst_xyz.shape[0], dst_xyz, idx, int(brute))
    if rc != 0:
        raise ValueError(f"orc_nearest rc={rc}")
    return idx


def bilinear(cell_xyz, tri, verticesOnCell, dst_xyz, brute=False):
    cell_xyz = np.ascontiguousarray(cell_xyz, np.float64)
    tri = np.ascontiguousarray(tri, np.int32)
    voc = np.ascontiguousarray(verticesOnCell, np.int32)
    dst_xyz = np.ascontiguousarray(dst_xyz, np.float64)
    n = dst_xyz.shape[0]
    elem = np.empty(n, np.int32)
    col = np.empty((n, 3), np.int32)
    w = np.empty((n, 3), np.float64)
    rc = lib().orc_bilinear(cell_xyz.shape[0], cell_xyz, tri.shape[0], tri, voc.shape[1], voc, n, dst_xyz,
                            elem, col, w, int(brute))
    if rc != 0:
        raise ValueError(f"orc_bilinear rc={rc}")
    return elem, col, w


def ell_to_csr(mask, col, w):
    """Fixed-width rows (col [n][k], w [n][k]) + mapped mask -> CSR (rowptr, col, w)."""
    n, k = col.shape
    cnt = np.where(mask, k, 0).astype(np.int64)
    rowptr = np.zeros(n + 1, np.int32)
    np.cumsum(cnt, out=rowptr[1:])
    return rowptr, np.ascontiguousarray(col[mask].reshape(-1), np.int32), \
        np.ascontiguousarray(w[mask].reshape(-1), np.float64)


def nearest_to_csr(idx):
    n = idx.shape[0]
    return np.arange(n + 1, dtype=np.int32), np.ascontiguousarray(idx, np.int32), np.ones(n, np.float64)


# ---------------------------------------------------------------- apply
def apply(rowptr, col, w, src, out_dtype=np.float32, tiled=False):
    """src [nSrc][nlev] (or [nSrc]) -> dst [nlev][nDst]."""
    src = np.ascontiguousarray(src)
    if src.ndim == 1:
        src = src.reshape(-1, 1)
    nlev = src.shape[1]
    nDst = rowptr.shape[0] - 1
    dst = np.empty((nlev, nDst), out_dtype)
    key = ("f32" if src.dtype == np.float32 else "f64") + "_" + ("f32" if dst.dtype == np.float32 else "f64")
    name = "orc_apply_" + key + ("_tiled" if tiled and key == "f32_f32" else "")
    getattr(lib(), name)(nDst, np.ascontiguousarray(rowptr, np.int32), np.ascontiguousarray(col, np.int32),
                         np.ascontiguousarray(w, np.float64), nlev, src, dst)
    return dst


def rotate_winds(u, v, cosa, sina):
    """In place on [nlev][n] arrays (fp64 or fp32)."""
    n = cosa.size
    nlev = u.size // n
    cosa = np.ascontiguousarray(cosa, np.float64).reshape(-1)
    sina = np.ascontiguousarray(sina, np.float64).reshape(-1)
    if u.dtype == np.float64:
        lib().orc_rotate_winds(n, nlev, u, v, cosa, sina)
    else:
        lib().orc_rotate_winds_f32(n, nlev, u, v, cosa, sina)
    return u, v


TOPO_PERI, TOPO_SPOLE, TOPO_NPOLE = 1, 2, 4


def bilinear_quadgrid(src_xyz_grid, dst_xyz, brute=False, topo=0):
    """src_xyz_grid [nj][ni][3] (CENTER points) -> (elem, col [n][4], w [n][4]).
    topo: 0 = ESMF_GridCreateNoPeriDim; TOPO_PERI (| TOPO_SPOLE | TOPO_NPOLE) = ESMF_GridCreate1PeriDim with monopole
    caps at the bottom / top row of this block (model_grid.F90:684-696).  Cap rows come back 4 wide (see
    mpassit_oracle.c); quadgrid_csr expands them."""
    src = np.ascontiguousarray(src_xyz_grid, np.float64)
    nj, ni = src.shape[0], src.shape[1]
    dst = np.ascontiguousarray(dst_xyz, np.float64).reshape(-1, 3)
    n = dst.shape[0]
    elem = np.empty(n, np.int64)
    col = np.empty((n, 4), np.int32)
    w = np.empty((n, 4), np.float64)
    rc = lib().orc_bilinear_quadgrid_topo(ni, nj, src.reshape(-1, 3), n, dst, elem, col, w, int(brute), int(topo))
    if rc != 0:
        raise ValueError(f"orc_bilinear_quadgrid rc={rc}")
    return elem, col, w


def quadgrid_csr(ni, nj, elem, col, w, topo=0):
    """CSR of a centre -> edge matrix.  Quad rows: 4 entries.  Monopole-cap rows (elem >= number of quads): the
    artificial pole's value is the average of its row, so the row holds ni entries ws/ni, plus a and b on the
    cap triangle's two row points -- emitted in ascending column order."""
    n = elem.shape[0]
    nquads = (ni if topo & TOPO_PERI else ni - 1) * (nj - 1)
    cap = elem >= nquads
    cnt = np.where(elem < 0, 0, np.where(cap, ni, 4)).astype(np.int64)
    rowptr = np.zeros(n + 1, np.int32)
    np.cumsum(cnt, out=rowptr[1:])
    C = np.empty(int(rowptr[-1]), np.int32)
    W = np.empty(int(rowptr[-1]), np.float64)
    q = np.flatnonzero((elem >= 0) & ~cap)
    idx = rowptr[q][:, None] + np.arange(4)[None, :]
    C[idx] = col[q]
    W[idx] = w[q]
    for t in np.flatnonzero(cap):
        base = int(col[t, 2])
        a_col, b_col, a, b, ws = int(col[t, 0]), int(col[t, 1]), w[t, 0], w[t, 1], w[t, 2]
        cc = base + np.arange(ni, dtype=np.int32)
        ww = np.full(ni, ws / ni)
        ww[a_col - base] = ww[a_col - base] + a
        ww[b_col - base] = ww[b_col - base] + b
        C[rowptr[t]:rowptr[t + 1]] = cc
        W[rowptr[t]:rowptr[t + 1]] = ww
    return rowptr, C, W


def apply_planes(rowptr, col, w, src_planes):
    """src [nlev][plane] fp64 -> dst [nlev][nDst] fp64 (level-slowest source)."""
    src = np.ascontiguousarray(src_planes, np.float64)
    if src.ndim == 1:
        src = src.reshape(1, -1)
    nlev, plane = src.shape
    nDst = rowptr.shape[0] - 1
    dst = np.empty((nlev, nDst), np.float64)
    lib().orc_apply_planes_f64(nDst, np.ascontiguousarray(rowptr, np.int32), np.ascontiguousarray(col, np.int32),
                               np.ascontiguousarray(w, np.float64), nlev, plane, src, dst)
    return dst


def conserve(cell_xyz, vert_xyz, verticesOnCell, corner_xyz_grid, brute=False):
    """corner_xyz_grid [(nj+1)][(ni+1)][3] -> CSR (rowptr, col, w) over the ni x nj destination cells."""
    cxyz = np.ascontiguousarray(cell_xyz, np.float64)
    vxyz = np.ascontiguousarray(vert_xyz, np.float64)
    voc = np.ascontiguousarray(verticesOnCell, np.int32)
    cor = np.ascontiguousarray(corner_xyz_grid, np.float64)
    nj, ni = cor.shape[0] - 1, cor.shape[1] - 1
    n = ni * nj
    cnt = np.zeros(n, np.int32)
    L = lib()
    rc = L.orc_conserve(cxyz.shape[0], cxyz, vxyz, voc.shape[1], voc, ni, nj, cor.reshape(-1, 3),
                        cnt.ctypes.data, None, None, None, int(brute))
    if rc != 0:
        raise ValueError(f"orc_conserve rc={rc}")
    rowptr = np.zeros(n + 1, np.int32)
    np.cumsum(cnt, out=rowptr[1:])
    col = np.empty(max(int(rowptr[-1]), 1), np.int32)
    w = np.empty(max(int(rowptr[-1]), 1), np.float64)
    rc = L.orc_conserve(cxyz.shape[0], cxyz, vxyz, voc.shape[1], voc, ni, nj, cor.reshape(-1, 3),
                        None, rowptr.ctypes.data, col.ctypes.data, w.ctypes.data, int(brute))
    if rc != 0:
        raise ValueError(f"orc_conserve rc={rc}")
    return rowptr, col[: rowptr[-1]], w[: rowptr[-1]]


def bilinear_node(cell_xyz, vert_xyz, verticesOnCell, dst_xyz, brute=False):
    """Source on mesh nodes (Voronoi vertices): (elem=cell id or -1, col [n][3] vertex ids, w [n][3])."""
    cxyz = np.ascontiguousarray(cell_xyz, np.float64)
    vxyz = np.ascontiguousarray(vert_xyz, np.float64)
    voc = np.ascontiguousarray(verticesOnCell, np.int32)
    dst = np.ascontiguousarray(dst_xyz, np.float64).reshape(-1, 3)
    n = dst.shape[0]
    elem = np.empty(n, np.int32)
    col = np.empty((n, 3), np.int32)
    w = np.empty((n, 3), np.float64)
    rc = lib().orc_bilinear_node(cxyz.shape[0], cxyz, vxyz, voc.shape[1], voc, n, dst, elem, col, w, int(brute))
    if rc != 0:
        raise ValueError(f"orc_bilinear_node rc={rc}")
    return elem, col, w
