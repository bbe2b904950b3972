"""Synthetic MPAS-shaped meshes and fields (test / bench infrastructure only).

Emits exactly the arrays the reference reads from an MPAS grid file
(/root/reference/model_grid.F90:292-417): ``lonCell, latCell, lonVertex,
latVertex`` in RADIANS and ``verticesOnCell`` as ``[nCells][maxEdges]`` int32,
1-based with 0 marking an unused slot (model_grid.F90:448,479).  Nothing here
is part of the regridding product path; it only fabricates inputs of the
shapes named in BASELINE.json / SURVEY.md §8(d).

Mesh construction: cell centres on the unit sphere + a triangle list (the
Delaunay triangulation of the centres).  The Voronoi vertex of a triangle is
its spherical circumcentre; a cell's polygon is the ring of circumcentres of
its incident triangles in counter-clockwise order (seen from outside the
sphere).  Regional meshes are culled the way MPAS limited-area meshes are:
cells whose ring is incomplete are dropped, every vertex of a kept cell is
kept, so boundary vertices touch fewer than three kept cells and therefore
carry no dual (Delaunay) triangle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

SEED = 20240625
EARTH_RADIUS_M = 6370000.0  # /root/reference/constants_module.F90:27


@dataclass
class MpasMesh:
    """Arrays of one synthetic MPAS mesh, in MPAS grid-file conventions."""

    lonCell: np.ndarray  # [nCells] f64 radians, [0, 2pi)
    latCell: np.ndarray  # [nCells] f64 radians
    lonVertex: np.ndarray  # [nVertices] f64 radians
    latVertex: np.ndarray  # [nVertices] f64 radians
    verticesOnCell: np.ndarray  # [nCells][maxEdges] int32, 1-based, 0 = pad
    meta: dict = field(default_factory=dict)

    @property
    def nCells(self) -> int:
        return int(self.lonCell.shape[0])

    @property
    def nVertices(self) -> int:
        return int(self.lonVertex.shape[0])

    @property
    def maxEdges(self) -> int:
        return int(self.verticesOnCell.shape[1])


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------

def _unit(v: np.ndarray) -> np.ndarray:
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def lonlat_to_xyz(lon_rad: np.ndarray, lat_rad: np.ndarray) -> np.ndarray:
    cl = np.cos(lat_rad)
    return np.stack([cl * np.cos(lon_rad), cl * np.sin(lon_rad), np.sin(lat_rad)], axis=-1)


def xyz_to_lonlat(xyz: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    lon = np.arctan2(xyz[:, 1], xyz[:, 0])
    lon = np.where(lon < 0.0, lon + 2.0 * np.pi, lon)  # MPAS stores [0, 2pi)
    lat = np.arcsin(np.clip(xyz[:, 2], -1.0, 1.0))
    return lon, lat


def _circumcentres(xyz: np.ndarray, tri: np.ndarray) -> np.ndarray:
    a, b, c = xyz[tri[:, 0]], xyz[tri[:, 1]], xyz[tri[:, 2]]
    n = np.cross(b - a, c - a)
    n = _unit(n)
    s = np.sign(np.einsum("ij,ij->i", n, a))
    s[s == 0] = 1.0
    return n * s[:, None]


def mesh_from_triangulation(xyz: np.ndarray, tri: np.ndarray, keep: np.ndarray | None = None,
                            max_edges: int = 10, meta: dict | None = None) -> MpasMesh:
    """Build MPAS arrays from cell centres ``xyz`` [n][3] and triangles ``tri`` [m][3].

    ``keep`` (bool per cell) selects the cells that survive culling; by default
    every cell whose triangle ring is closed (degree == number of distinct
    neighbours) is kept, which on a closed sphere is every cell.
    """
    n = xyz.shape[0]
    tri = np.asarray(tri, dtype=np.int64)
    m = tri.shape[0]
    vert_xyz = _circumcentres(xyz, tri)

    # (cell, triangle) incidence pairs
    cell_of = tri.reshape(-1)
    tri_of = np.repeat(np.arange(m, dtype=np.int64), 3)
    deg = np.bincount(cell_of, minlength=n)

    if keep is None:
        # ring closed <=> #triangles == #distinct neighbours (Euler on the star)
        e = np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]], axis=0)
        e = np.concatenate([e, e[:, ::-1]], axis=0)
        ekey = np.unique(e[:, 0] * n + e[:, 1])
        nnbr = np.bincount((ekey // n).astype(np.int64), minlength=n)
        keep = (deg == nnbr) & (deg >= 3)
    keep = np.asarray(keep, dtype=bool)

    # angular order of incident vertices around each cell, CCW seen from outside
    c = xyz[cell_of]
    v = vert_xyz[tri_of]
    zhat = np.array([0.0, 0.0, 1.0])
    east = np.cross(zhat, c)
    en = np.linalg.norm(east, axis=1)
    polar = en < 1e-12
    east[polar] = np.array([1.0, 0.0, 0.0])
    east = _unit(east)
    north = np.cross(c, east)
    d = v - c
    ang = np.arctan2(np.einsum("ij,ij->i", d, north), np.einsum("ij,ij->i", d, east))
    order = np.lexsort((ang, cell_of))
    cell_s, tri_s = cell_of[order], tri_of[order]
    start = np.concatenate([[0], np.cumsum(deg)])[:-1]
    slot = np.arange(cell_s.shape[0]) - start[cell_s]

    sel = keep[cell_s]
    if int(deg[keep].max()) > max_edges:
        max_edges = int(deg[keep].max())
    new_cell = -np.ones(n, dtype=np.int64)
    new_cell[keep] = np.arange(int(keep.sum()))
    used_tri = np.zeros(m, dtype=bool)
    used_tri[tri_s[sel]] = True
    new_vert = -np.ones(m, dtype=np.int64)
    new_vert[used_tri] = np.arange(int(used_tri.sum()))

    voc = np.zeros((int(keep.sum()), max_edges), dtype=np.int32)
    voc[new_cell[cell_s[sel]], slot[sel]] = (new_vert[tri_s[sel]] + 1).astype(np.int32)

    lonC, latC = xyz_to_lonlat(xyz[keep])
    lonV, latV = xyz_to_lonlat(vert_xyz[used_tri])
    md = dict(meta or {})
    md.setdefault("nCells", int(keep.sum()))
    return MpasMesh(lonC, latC, lonV, latV, voc, md)


# --------------------------------------------------------------------------
# global meshes
# --------------------------------------------------------------------------

def fibonacci_points(n: int) -> np.ndarray:
    k = np.arange(n, dtype=np.float64) + 0.5
    z = 1.0 - 2.0 * k / n
    phi = k * (math.pi * (3.0 - math.sqrt(5.0)))
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    return np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)


def icosahedral_points(level: int) -> np.ndarray:
    """Recursive midpoint subdivision of the icosahedron: 10*4**level + 2 points."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t],
                  [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v = _unit(v)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4],
                  [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8],
                  [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(level):
        nv = v.shape[0]
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        es = np.sort(e, axis=1)
        key = es[:, 0] * nv + es[:, 1]
        ukey, inv = np.unique(key, return_inverse=True)
        mid = _unit(v[ukey // nv] + v[ukey % nv])
        mid_id = nv + inv
        nf = f.shape[0]
        a, b, c = f[:, 0], f[:, 1], f[:, 2]
        ab, bc, ca = mid_id[:nf], mid_id[nf:2 * nf], mid_id[2 * nf:]
        f = np.concatenate([np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1),
                            np.stack([c, ca, bc], 1), np.stack([ab, bc, ca], 1)], axis=0)
        v = np.concatenate([v, mid], axis=0)
    return v


def geodesic_mesh_points(freq: int):
    """Class-I geodesic subdivision of the icosahedron at frequency `freq` (every edge cut into `freq` parts):
    10 freq^2 + 2 points and their 20 freq^2 triangles, written down analytically -- no O(n log n) hull, so the
    6.5 M-cell mesh of BASELINE.json configs[3] (freq = 806) takes seconds instead of minutes.  Point ids: the
    12 corners, then the interior points of the 30 edges, then the interior points of the 20 faces row by row
    (spatially coherent, like an MPAS mesh ordered by its generator)."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = _unit(np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t],
                        [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64))
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4],
                  [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8],
                  [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    n = int(freq)
    edges = {}
    for a, b, c in f.tolist():
        for x, y in ((a, b), (a, c), (b, c)):
            edges.setdefault((min(x, y), max(x, y)), len(edges))
    ne, T = len(edges), (n - 1) * (n - 2) // 2
    e0, f0 = 12, 12 + ne * (n - 1)
    npts = f0 + 20 * T
    pts = np.empty((npts, 3), np.float64)
    pts[:12] = v
    k = np.arange(1, n, dtype=np.float64)[:, None] / n
    for (a, b), e in edges.items():                      # edge points from the smaller corner id: shared by both faces
        pts[e0 + e * (n - 1): e0 + (e + 1) * (n - 1)] = v[a] + k * (v[b] - v[a])
    I, J = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
    inside = (I >= 1) & (J >= 1) & (I + J <= n - 1)
    # interior index of (i, j): rows j = 1 .. n-2, each with n-1-j points i = 1 .. n-1-j
    row0 = np.concatenate([[0], np.cumsum(n - 1 - np.arange(1, n - 1))])[:-1] if n > 2 else np.zeros(0, np.int64)
    tris = []
    for fi, (a, b, c) in enumerate(f.tolist()):
        ids = np.full((n + 1, n + 1), -1, np.int64)
        if T:
            ii, jj = I[inside], J[inside]
            loc = row0[jj - 1] + (ii - 1)
            ids[ii, jj] = f0 + fi * T + loc
            pts[f0 + fi * T + loc] = v[a] + (ii[:, None] / n) * (v[b] - v[a]) + (jj[:, None] / n) * (v[c] - v[a])
        ids[0, 0], ids[n, 0], ids[0, n] = a, b, c

        def edge_ids(x, y, par):       # par = steps from x towards y, 1 .. n-1
            e = edges[(min(x, y), max(x, y))]
            kk = par if x < y else n - par
            return e0 + e * (n - 1) + (kk - 1)

        m = np.arange(1, n)
        ids[m, 0] = edge_ids(a, b, m)
        ids[0, m] = edge_ids(a, c, m)
        ids[n - m, m] = edge_ids(b, c, m)
        up = (I + J <= n - 1) & (I < n) & (J < n)
        iu, ju = I[up], J[up]
        tris.append(np.stack([ids[iu, ju], ids[iu + 1, ju], ids[iu, ju + 1]], 1))
        dn = (I + J <= n - 2) & (I < n) & (J < n)
        idn, jdn = I[dn], J[dn]
        tris.append(np.stack([ids[idn + 1, jdn], ids[idn + 1, jdn + 1], ids[idn, jdn + 1]], 1))
    return _unit(pts), np.concatenate(tris, 0)


def schmidt_focus(xyz: np.ndarray, focus_lat: float, focus_lon: float, spacing_ratio: float) -> np.ndarray:
    """Schmidt transformation (sin lat' = (D + sin lat) / (1 + D sin lat), D = (r - 1) / (r + 1)) followed by a
    rotation of the pole onto the focus.  It is a Moebius map of the sphere: circles go to circles, so the
    (spherical) Delaunay triangulation of the points is unchanged."""
    D = (spacing_ratio - 1.0) / (spacing_ratio + 1.0)
    z = xyz[:, 2]
    z2 = (D + z) / (1.0 + D * z)
    s = np.sqrt(np.maximum(0.0, 1.0 - z2 * z2)) / np.sqrt(np.maximum(1e-300, 1.0 - z * z))
    p = np.stack([xyz[:, 0] * s, xyz[:, 1] * s, z2], axis=1)
    la, lo = math.radians(focus_lat), math.radians(focus_lon)
    ez = np.array([math.cos(la) * math.cos(lo), math.cos(la) * math.sin(lo), math.sin(la)])
    ex = _unit(np.cross([0.0, 0.0, 1.0], ez)[None, :])[0]
    ey = np.cross(ez, ex)
    return _unit(p[:, :1] * ex + p[:, 1:2] * ey + p[:, 2:] * ez)


def variable_geodesic_mesh(freq: int = 806, focus_lat: float = 38.5, focus_lon: float = -97.5,
                           spacing_ratio: float = 5.0, max_edges: int = 10) -> MpasMesh:
    """BASELINE.json configs[3] at its stated size: freq = 806 -> 6,496,362 cells ("~6.5 M"), spacing graded 5:1
    (15 km -> 3 km) towards CONUS.  Topology from geodesic_mesh_points, geometry through schmidt_focus."""
    xyz, tri = geodesic_mesh_points(freq)
    # a tiny rotation first so that no icosahedron corner sits exactly on the Schmidt axis
    c, s_ = math.cos(0.3), math.sin(0.3)
    xyz = np.stack([xyz[:, 0], c * xyz[:, 1] - s_ * xyz[:, 2], s_ * xyz[:, 1] + c * xyz[:, 2]], 1)
    p = schmidt_focus(xyz, focus_lat, focus_lon, spacing_ratio)
    return mesh_from_triangulation(p, tri, keep=np.ones(p.shape[0], bool), max_edges=max_edges,
                                   meta={"kind": "global-geodesic-variable", "freq": freq, "spacing_ratio": spacing_ratio})


def _sphere_delaunay(xyz: np.ndarray) -> np.ndarray:
    from scipy.spatial import ConvexHull

    return ConvexHull(xyz).simplices.astype(np.int64)


def global_mesh(n_cells: int = 40962, kind: str = "icos", lloyd_iters: int = 0, jitter: float = 0.0,
                seed: int = SEED, max_edges: int = 10) -> MpasMesh:
    """Quasi-uniform global mesh (config C1: 40,962 cells = icosahedral level 6)."""
    rng = np.random.default_rng(seed)
    if kind == "icos":
        level = int(round(math.log((n_cells - 2) / 10.0, 4)))
        xyz = icosahedral_points(level)
    elif kind == "fib":
        xyz = fibonacci_points(n_cells)
    elif kind == "random":
        xyz = _unit(rng.standard_normal((n_cells, 3)))
    else:
        raise ValueError(kind)
    if jitter > 0.0:
        h = math.sqrt(4.0 * math.pi / xyz.shape[0])
        xyz = _unit(xyz + jitter * h * rng.standard_normal(xyz.shape))
    # shuffle-free: MPAS ids keep generator order (spatially coherent for icos/fib)
    for _ in range(lloyd_iters):
        tri = _sphere_delaunay(xyz)
        cc = _circumcentres(xyz, tri)
        acc = np.zeros_like(xyz)
        for k in range(3):
            np.add.at(acc, tri[:, k], cc)
        xyz = _unit(acc)
    tri = _sphere_delaunay(xyz)
    return mesh_from_triangulation(xyz, tri, max_edges=max_edges,
                                   meta={"kind": f"global-{kind}", "seed": seed})


def variable_global_mesh(n_cells: int = 655362, focus_lat: float = 38.5, focus_lon: float = -97.5,
                         spacing_ratio: float = 5.0, max_edges: int = 10) -> MpasMesh:
    """Variable-resolution global mesh (config C4's shape: 15 km -> 3 km is spacing_ratio = 5): Fibonacci
    points pulled towards a focus by a Schmidt transformation (sin(lat') = (D + sin lat) / (1 + D sin lat),
    D = (r - 1) / (r + 1) for a coarse/fine spacing ratio r), then the pole rotated onto the focus.  Cell ids
    follow the Fibonacci spiral, i.e. rings around the focus."""
    xyz = fibonacci_points(n_cells)
    D = (spacing_ratio - 1.0) / (spacing_ratio + 1.0)
    z = xyz[:, 2]
    z2 = (D + z) / (1.0 + D * z)
    s = np.sqrt(np.maximum(0.0, 1.0 - z2 * z2)) / np.sqrt(np.maximum(1e-300, 1.0 - z * z))
    p = np.stack([xyz[:, 0] * s, xyz[:, 1] * s, z2], axis=1)
    la, lo = math.radians(focus_lat), math.radians(focus_lon)
    ez = np.array([math.cos(la) * math.cos(lo), math.cos(la) * math.sin(lo), math.sin(la)])
    ex = _unit(np.cross([0.0, 0.0, 1.0], ez)[None, :])[0]
    ey = np.cross(ez, ex)
    p = _unit(p[:, :1] * ex + p[:, 1:2] * ey + p[:, 2:] * ez)
    return mesh_from_triangulation(p, _sphere_delaunay(p), max_edges=max_edges,
                                   meta={"kind": "global-variable", "spacing_ratio": spacing_ratio})


# --------------------------------------------------------------------------
# Lambert-conformal plane <-> sphere (generator-side only; the product's
# projection code lives in mpassit_b200/host/)
# --------------------------------------------------------------------------

def lc_inverse(x_m: np.ndarray, y_m: np.ndarray, ref_lat: float, ref_lon: float,
               truelat: float, radius: float = EARTH_RADIUS_M) -> tuple[np.ndarray, np.ndarray]:
    """Tangent-cone Lambert conformal inverse; (x,y) metres from (ref_lat, ref_lon) -> radians."""
    phi1 = math.radians(truelat)
    nn = math.sin(phi1)
    F = math.cos(phi1) * math.tan(math.pi / 4 + phi1 / 2) ** nn / nn
    rho0 = radius * F / math.tan(math.pi / 4 + math.radians(ref_lat) / 2) ** nn
    rho = np.sign(nn) * np.sqrt(x_m * x_m + (rho0 - y_m) ** 2)
    theta = np.arctan2(x_m, rho0 - y_m)
    lat = 2.0 * np.arctan((radius * F / rho) ** (1.0 / nn)) - math.pi / 2
    lon = math.radians(ref_lon) + theta / nn
    return lon, lat


def regional_hex_mesh(spacing_m: float = 3000.0, extent_x_m: float = 5600e3, extent_y_m: float = 3400e3,
                      ref_lat: float = 38.5, ref_lon: float = -97.5, truelat: float = 38.5,
                      jitter: float = 0.05, seed: int = SEED, max_edges: int = 10) -> MpasMesh:
    """Jittered hexagonal lattice in the LC plane lifted to the sphere (configs C2/C3/C5).

    With jitter << spacing the Delaunay topology equals the perfect lattice's,
    so connectivity is written down analytically (no O(n log n) triangulation):
    2.4 M cells take a few seconds.
    """
    rng = np.random.default_rng(seed)
    a = float(spacing_m)
    dy = a * math.sqrt(3.0) / 2.0
    NC = int(math.ceil(extent_x_m / a)) + 3
    NR = int(math.ceil(extent_y_m / dy)) + 3
    r, c = np.meshgrid(np.arange(NR), np.arange(NC), indexing="ij")
    x = (c + 0.5 * (r & 1)) * a
    y = r * dy
    x = x - x.mean()
    y = y - y.mean()
    x = x + jitter * a * rng.standard_normal(x.shape)
    y = y + jitter * a * rng.standard_normal(y.shape)
    lon, lat = lc_inverse(x.reshape(-1), y.reshape(-1), ref_lat, ref_lon, truelat)
    xyz = lonlat_to_xyz(lon, lat)

    idx = (r * NC + c)
    rr, cc_ = r[:-1, :-1], c[:-1, :-1]
    even = (rr & 1) == 0
    p = idx[:-1, :-1]          # (r, c)
    pr = idx[:-1, 1:]          # (r, c+1)
    u = idx[1:, :-1]           # (r+1, c)
    ur = idx[1:, 1:]           # (r+1, c+1)
    # even row r: A=(p, pr, u)  B=(u, ur, pr);  odd row r: A=(p, pr, ur)  B=(u, ur, p)
    A = np.where(even[..., None], np.stack([p, pr, u], -1), np.stack([p, pr, ur], -1))
    B = np.where(even[..., None], np.stack([u, ur, pr], -1), np.stack([u, ur, p], -1))
    tri = np.concatenate([A.reshape(-1, 3), B.reshape(-1, 3)], axis=0)

    keep = np.ones((NR, NC), dtype=bool)
    keep[0, :] = keep[-1, :] = False
    keep[:, 0] = keep[:, -1] = False
    return mesh_from_triangulation(xyz, tri, keep=keep.reshape(-1), max_edges=max_edges,
                                   meta={"kind": "regional-hex", "spacing_m": a, "seed": seed,
                                         "ref_lat": ref_lat, "ref_lon": ref_lon, "truelat": truelat})


def regional_delaunay_mesh(n_points: int = 4000, extent_x_m: float = 600e3, extent_y_m: float = 400e3,
                           ref_lat: float = 38.5, ref_lon: float = -97.5, truelat: float = 38.5,
                           seed: int = SEED, lloyd_iters: int = 2, max_edges: int = 10) -> MpasMesh:
    """Irregular regional mesh (random points, optional Lloyd smoothing) for parity tests:
    exercises 5/7-gons, obtuse triangles and a ragged boundary."""
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    pts = np.stack([(rng.random(n_points) - 0.5) * extent_x_m, (rng.random(n_points) - 0.5) * extent_y_m], 1)
    for _ in range(lloyd_iters):
        d = Delaunay(pts)
        t = d.simplices
        A_, B_, C_ = pts[t[:, 0]], pts[t[:, 1]], pts[t[:, 2]]
        cen = (A_ + B_ + C_) / 3.0
        acc = np.zeros_like(pts)
        cnt = np.zeros(n_points)
        for k in range(3):
            np.add.at(acc, t[:, k], cen)
            np.add.at(cnt, t[:, k], 1.0)
        interior = np.ones(n_points, dtype=bool)
        interior[np.unique(d.convex_hull)] = False
        pts = np.where(interior[:, None], acc / np.maximum(cnt, 1)[:, None], pts)
    d = Delaunay(pts)
    lon, lat = lc_inverse(pts[:, 0], pts[:, 1], ref_lat, ref_lon, truelat)
    xyz = lonlat_to_xyz(lon, lat)
    return mesh_from_triangulation(xyz, d.simplices.astype(np.int64), max_edges=max_edges,
                                   meta={"kind": "regional-delaunay", "seed": seed})


def renumber_cells(mesh: MpasMesh, order: str, seed: int = SEED) -> MpasMesh:
    """The same mesh with its cells renumbered (vertices keep their ids): `rowmajor` = as generated,
    `morton` = along a Z-order curve of (lon, lat) (what a locality sort such as MPAS-Tools' sort_mesh
    produces), `random` = no locality at all.  The apply kernel fetches runs of consecutively numbered
    cells with one bulk copy, so its speed depends on the numbering; results do not."""
    if order in ("rowmajor", "", None):
        return mesh
    n = mesh.nCells
    if order == "random":
        perm = np.random.default_rng(seed).permutation(n)          # new id k holds old cell perm[k]
    elif order == "morton":
        def spread(v):
            v = v.astype(np.uint64) & np.uint64(0xFFFF)
            v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF)
            v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F)
            v = (v | (v << np.uint64(2))) & np.uint64(0x33333333)
            v = (v | (v << np.uint64(1))) & np.uint64(0x55555555)
            return v
        lon = np.where(mesh.lonCell > np.pi, mesh.lonCell - 2 * np.pi, mesh.lonCell)
        qx = ((lon - lon.min()) / max(np.ptp(lon), 1e-30) * 65535).astype(np.uint64)
        qy = ((mesh.latCell - mesh.latCell.min()) / max(np.ptp(mesh.latCell), 1e-30) * 65535).astype(np.uint64)
        perm = np.argsort(spread(qx) | (spread(qy) << np.uint64(1)), kind="stable")
    else:
        raise ValueError(order)
    meta = dict(mesh.meta, cell_order=order)
    return MpasMesh(np.ascontiguousarray(mesh.lonCell[perm]), np.ascontiguousarray(mesh.latCell[perm]), mesh.lonVertex,
                    mesh.latVertex, np.ascontiguousarray(mesh.verticesOnCell[perm]), meta)


# --------------------------------------------------------------------------
# fields (SURVEY.md §8(d) value recipes)
# --------------------------------------------------------------------------

def smooth_field(lon: np.ndarray, lat: np.ndarray, nlev: int, seed: int = SEED, noise: float = 0.5,
                 dtype=np.float32) -> np.ndarray:
    """[n][nlev] level-fastest (MPAS file order, input_data.F90:630)."""
    rng = np.random.default_rng(seed)
    z = np.sin(lat)
    base = 280.0 + 20.0 * z + 5.0 * np.sin(3.0 * lon) * np.cos(2.0 * lat)
    lev = np.arange(nlev, dtype=np.float64)
    f = base[:, None] - 0.1 * lev[None, :]
    if noise > 0:
        f = f + noise * rng.standard_normal(f.shape)
    return np.ascontiguousarray(f.astype(dtype))


def moisture_field(lon, lat, nlev, seed=SEED, dtype=np.float32):
    f = smooth_field(lon, lat, nlev, seed=seed, noise=0.5, dtype=np.float64) - 285.0
    return np.ascontiguousarray(np.maximum(0.0, f).astype(dtype) * dtype(1e-3))


def integer_field(n: int, nmax: int, seed: int = SEED, dtype=np.float32) -> np.ndarray:
    ids = np.arange(n, dtype=np.uint64)
    h = (ids * np.uint64(2654435761) + np.uint64(seed)) % np.uint64(2 ** 32)
    return ((h >> np.uint64(7)) % np.uint64(nmax) + np.uint64(1)).astype(dtype)


def patchy_field(lon, lat, seed=SEED, dtype=np.float32):
    """Non-negative with ~70 % exact zeros (snow-like)."""
    v = np.sin(5.0 * lon + 0.3) * np.cos(7.0 * lat - 0.2) + 0.1 * np.sin(40.0 * lon) * np.sin(31.0 * lat)
    return np.ascontiguousarray(np.maximum(0.0, v - 0.45).astype(dtype) * dtype(100.0))
