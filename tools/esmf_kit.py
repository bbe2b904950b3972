"""Pinning kit: everything a maintainer with ESMF needs to hold the engine's weights to ESMF's own.

This image has no ESMF, so parity is pinned to the oracle / golden vectors.  On a machine with ESMF:

    python tools/esmf_kit.py export --workload mini --dir kit/          # needs a GPU for the engine's weights
    ESMF_RegridWeightGen -s kit/mesh.nc --src_type ESMF --src_loc center -d kit/grid.nc --dst_type SCRIP \
        -m bilinear -i --dst_regional -w kit/esmf_bilinear.nc            # same arguments as interp.F90:118-128:
    ESMF_RegridWeightGen ... -m neareststod -w kit/esmf_nearest.nc       #   unmapped action IGNORE (-i), defaults otherwise
    ESMF_RegridWeightGen ... -m conserve   -w kit/esmf_conserve.nc
    python tools/esmf_kit.py compare kit/esmf_bilinear.nc kit/ours_bilinear.nc

`export` writes the source mesh as an ESMF unstructured mesh file built the way the reference builds its ESMF_Mesh
(model_grid.F90:446-497: nodes = MPAS vertices in degrees with the > 180 wrap, elements = verticesOnCell entries != 0 in
file order, element coordinates = cell centres), the target as a SCRIP grid file (centres + the 4 CORNER-stagger points
around each centre, model_grid.F90:949-1038), and -- with a GPU -- the engine's matrices in ESMF's weight-file format.
`compare` reports, for two weight files, whether the row structure is identical, how many rows reference the same
source set, and the largest weight difference.  Files are NetCDF classic (ESMF's default weight-file format).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def mesh_degrees(lon_rad, lat_rad):
    """model_grid.F90:450-454 / 464-468: degrees with PI = 4 atan(1), longitudes above 180 wrapped."""
    pi = 4.0 * np.arctan(1.0)
    lon = np.asarray(lon_rad, np.float64) * 180.0 / pi
    lon = np.where(lon > 180.0, lon - 360.0, lon)
    return lon, np.asarray(lat_rad, np.float64) * 180.0 / pi


def write_esmf_mesh(path: str, mesh) -> None:
    from scipy.io import netcdf_file

    lonV, latV = mesh_degrees(mesh.lonVertex, mesh.latVertex)
    lonC, latC = mesh_degrees(mesh.lonCell, mesh.latCell)
    voc = np.asarray(mesh.verticesOnCell, np.int32)
    nC, maxE = voc.shape
    # entries != 0, kept in file order and packed to the front (model_grid.F90:476-486)
    conn = np.full((nC, maxE), -1, np.int32)
    num = (voc > 0).sum(axis=1).astype(np.int32)
    for c in range(nC):
        v = voc[c][voc[c] > 0]
        conn[c, :v.size] = v
    with netcdf_file(path, "w", version=2) as f:
        f.createDimension("nodeCount", lonV.size)
        f.createDimension("elementCount", nC)
        f.createDimension("maxNodePElement", maxE)
        f.createDimension("coordDim", 2)
        f.gridType = b"unstructured"
        f.version = b"0.9"
        v = f.createVariable("nodeCoords", "f8", ("nodeCount", "coordDim"))
        v.units = b"degrees"
        v[:] = np.stack([lonV, latV], axis=1)
        v = f.createVariable("elementConn", "i4", ("elementCount", "maxNodePElement"))
        v.long_name = b"Node Indices that define the element connectivity"
        v._FillValue = np.int32(-1)
        v.start_index = np.int32(1)
        v[:] = conn
        v = f.createVariable("numElementConn", "i4", ("elementCount",))
        v.long_name = b"Number of nodes per element"
        v[:] = num
        v = f.createVariable("centerCoords", "f8", ("elementCount", "coordDim"))
        v.units = b"degrees"
        v[:] = np.stack([lonC, latC], axis=1)


def write_scrip_grid(path: str, lat_c, lon_c, lat_corner, lon_corner) -> None:
    """Target grid: centres [nj][ni] and the CORNER stagger [nj+1][ni+1] (counter-clockwise corners per cell)."""
    from scipy.io import netcdf_file

    nj, ni = lat_c.shape
    cl = np.stack([lat_corner[:-1, :-1], lat_corner[:-1, 1:], lat_corner[1:, 1:], lat_corner[1:, :-1]], axis=-1).reshape(-1, 4)
    co = np.stack([lon_corner[:-1, :-1], lon_corner[:-1, 1:], lon_corner[1:, 1:], lon_corner[1:, :-1]], axis=-1).reshape(-1, 4)
    with netcdf_file(path, "w", version=2) as f:
        f.createDimension("grid_size", ni * nj)
        f.createDimension("grid_corners", 4)
        f.createDimension("grid_rank", 2)
        f.title = b"mpassit target grid"
        f.createVariable("grid_dims", "i4", ("grid_rank",))[:] = (ni, nj)
        for n, a in (("grid_center_lat", lat_c.reshape(-1)), ("grid_center_lon", lon_c.reshape(-1))):
            v = f.createVariable(n, "f8", ("grid_size",))
            v.units = b"degrees"
            v[:] = a
        f.createVariable("grid_imask", "i4", ("grid_size",))[:] = 1
        for n, a in (("grid_corner_lat", cl), ("grid_corner_lon", co)):
            v = f.createVariable(n, "f8", ("grid_size", "grid_corners"))
            v.units = b"degrees"
            v[:] = a


def compare(path_a: str, path_b: str, tol: float = 1e-12) -> dict:
    from mpassit_b200 import host

    na, nb, rpa, ca, wa = host.read_esmf_weights(path_a)
    na2, nb2, rpb, cb, wb = host.read_esmf_weights(path_b)
    out = {"n_a": (na, na2), "n_b": (nb, nb2), "n_s": (int(rpa[-1]), int(rpb[-1]))}
    if (na, nb) != (na2, nb2):
        out["verdict"] = "different grids"
        return out
    same_len = np.diff(rpa) == np.diff(rpb)
    out["rows_same_length"] = int(same_len.sum())
    out["rows_mapped"] = (int((np.diff(rpa) > 0).sum()), int((np.diff(rpb) > 0).sum()))
    same_set = 0
    max_dw = 0.0
    worst = -1
    for i in np.flatnonzero(same_len):
        a0, a1, b0, b1 = rpa[i], rpa[i + 1], rpb[i], rpb[i + 1]
        oa, ob = np.argsort(ca[a0:a1], kind="stable"), np.argsort(cb[b0:b1], kind="stable")
        if np.array_equal(ca[a0:a1][oa], cb[b0:b1][ob]):
            same_set += 1
            if a1 > a0:
                d = float(np.abs(wa[a0:a1][oa] - wb[b0:b1][ob]).max())
                if d > max_dw:
                    max_dw, worst = d, int(i)
    out["rows_same_sources"] = same_set
    out["max_weight_difference"] = max_dw
    out["worst_row"] = worst
    out["verdict"] = ("identical structure, weights within %g" % tol if same_set == nb and max_dw <= tol else
                      "identical structure" if same_set == nb else "structures differ")
    return out


def export(workload_name: str, directory: str) -> dict:
    from mpassit_b200 import host, workload

    os.makedirs(directory, exist_ok=True)
    host.load()
    wl = workload.make(workload_name, rundir=directory)
    paths = {"mesh": os.path.join(directory, "mesh.nc"), "grid": os.path.join(directory, "grid.nc")}
    write_esmf_mesh(paths["mesh"], wl.mesh)
    write_scrip_grid(paths["grid"], wl.grids["M"][0], wl.grids["M"][1], wl.grids["CORNER"][0], wl.grids["CORNER"][1])
    try:
        import torch

        gpu = torch.cuda.is_available()
    except Exception:
        gpu = False
    if gpu:
        from mpassit_b200 import lib as L
        from mpassit_b200.regrid import Regridder

        rg = Regridder(device=0)
        workload.load_geometry(rg, wl)
        for name, method, label in (("bilinear", L.BILINEAR, "Bilinear"), ("nearest", L.NEAREST_STOD, "Nearest source to destination"),
                                    ("conserve", L.CONSERVE, "Conservative")):
            r = rg.store(method, L.SRC_MESH_ELEMENT, L.CENTER)
            rp, col, w = r.export_csr()
            paths[name] = os.path.join(directory, f"ours_{name}.nc")
            host.write_esmf_weights(paths[name], wl.mesh.lonCell.size, rp, col, w, label)
            r.release()
        rg.close()
    return paths


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    e = sub.add_parser("export")
    e.add_argument("--workload", default="mini")
    e.add_argument("--dir", default="esmf_kit_out")
    c = sub.add_parser("compare")
    c.add_argument("a")
    c.add_argument("b")
    a = ap.parse_args()
    if a.cmd == "export":
        for k, v in export(a.workload, a.dir).items():
            print(f"{k}: {v}")
    else:
        for k, v in compare(a.a, a.b).items():
            print(f"{k}: {v}")


if __name__ == "__main__":
    main()
