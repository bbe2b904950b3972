"""Synthetic MPAS grid / diag / history files in the NetCDF classic format (bench / test infrastructure, like synth.py),
written with scipy -- an implementation independent of mpassit_b200/host/ncio.cpp -- and a reader for the output file.  Shapes and names follow what the
reference reads: model_grid.F90:286-419 (init file), input_data.F90:150-260 (diag), :337-812 (history).
"""
from __future__ import annotations

import numpy as np
from scipy.io import netcdf_file

START_TIME = "2024-03-25_00:00:00"
VALID_TIME = "2024-03-25_09:00:00"


def _xtime(f, stamp):
    v = f.createVariable("xtime", "S1", ("Time", "StrLen"))
    v[0, :] = np.frombuffer(stamp.ljust(64).encode(), "S1")


def write_grid_file(path, mesh, nz, nsoil, ter, version=2, real="f4", start=None):
    nC, nV = mesh.lonCell.size, mesh.lonVertex.size
    with netcdf_file(path, "w", version=version) as f:
        f.createDimension("Time", None)
        for n, l in (("nCells", nC), ("nVertices", nV), ("nVertLevels", nz), ("nVertLevelsP1", nz + 1),
                     ("maxEdges", mesh.verticesOnCell.shape[1]), ("nSoilLevels", nsoil), ("StrLen", 64)):
            f.createDimension(n, l)
        for n, a, d in (("lonCell", mesh.lonCell, "nCells"), ("latCell", mesh.latCell, "nCells"),
                        ("lonVertex", mesh.lonVertex, "nVertices"), ("latVertex", mesh.latVertex, "nVertices")):
            v = f.createVariable(n, "f8", (d,))
            v[:] = a
        v = f.createVariable("verticesOnCell", "i4", ("nCells", "maxEdges"))
        v[:] = mesh.verticesOnCell
        v = f.createVariable("ter", real, ("nCells",))
        v[:] = ter
        v = f.createVariable("zs", real, ("Time", "nCells", "nSoilLevels"))
        v[0, :, :] = np.tile(np.array([0.05, 0.25, 0.7, 1.5, 2.5, 3.5, 4.5, 5.5, 6.5][:nsoil]), (nC, 1))
        _xtime(f, start or START_TIME)


def write_field_file(path, fields, nCells, nVertices, nz, nsoil, attrs, version=2, vert_names=(), valid=None):
    """fields: list of (name, array [n][nlev] or [n]); every variable gets a Time record dimension, units and long_name."""
    with netcdf_file(path, "w", version=version) as f:
        f.createDimension("Time", None)
        for n, l in (("nCells", nCells), ("nVertices", nVertices), ("nVertLevels", nz), ("nVertLevelsP1", nz + 1),
                     ("nSoilLevels", nsoil), ("StrLen", 64)):
            f.createDimension(n, l)
        for k, v in attrs.items():
            setattr(f, k, v)
        _xtime(f, valid or VALID_TIME)
        for name, a in fields:
            a = np.asarray(a)
            hdim = "nVertices" if name in vert_names else "nCells"
            if a.ndim == 1 or a.shape[1] == 1:
                dims = ("Time", hdim)
                a = a.reshape(-1)
            else:
                zd = {nz: "nVertLevels", nz + 1: "nVertLevelsP1", nsoil: "nSoilLevels"}[a.shape[1]]
                dims = ("Time", hdim, zd)
            v = f.createVariable(name, a.dtype.str[1:], dims)
            v.units = f"unit_of_{name}".encode()
            v.long_name = f"long name of {name}".encode()
            v[0] = a


def read_output(path):
    """{name: array} + global attributes + per-variable attributes of an output file (record 0 dropped)."""
    with netcdf_file(path, "r", mmap=False) as f:
        out = {}
        vatts = {}
        for n, v in f.variables.items():
            a = v[:].copy()
            out[n] = a[0] if v.isrec else a
            vatts[n] = {k: (x.decode() if isinstance(x, bytes) else x) for k, x in v._attributes.items()}
        gatts = {k: (x.decode() if isinstance(x, bytes) else x) for k, x in f._attributes.items()}
        dims = dict(f.dimensions)
        order = list(f.variables.keys())
    return out, gatts, vatts, dims, order


def write_case(wl, rundir, fields, ter, real="f4", target_file=None, start=None, valid=None, config_dt=18.0):
    """The three input files of a workload + a namelist pointing at them.  fields: {group: [(mpas_name, array)]}
    for group in diag / hist_2d / hist_3d / soil ([n][nlev] level-fastest, as MPAS stores them)."""
    import os

    from mpassit_b200 import defaults

    m = wl.mesh
    paths = {k: os.path.join(rundir, f"mpas.{k}.nc") for k in ("init", "diag", "history")}
    paths["out"] = os.path.join(rundir, "mpassit_out.nc")
    start = start or START_TIME
    write_grid_file(paths["init"], m, wl.nz, wl.nsoil, np.asarray(ter).reshape(-1), real=real, start=start)
    nC, nV = m.lonCell.size, m.lonVertex.size
    cast = (lambda a: np.asarray(a, np.float64)) if real == "f8" else (lambda a: np.asarray(a, np.float32))
    write_field_file(paths["diag"], [(n, cast(a)) for n, a in fields.get("diag", [])], nC, nV, wl.nz, wl.nsoil,
                     dict(config_start_time=start.encode(), config_dt=np.float64(config_dt), output_interval=np.int32(3600)),
                     valid=valid)
    hist = [(n, cast(a)) for g in ("hist_2d", "hist_3d", "soil") for n, a in fields.get(g, [])]
    write_field_file(paths["history"], hist, nC, nV, wl.nz, wl.nsoil,
                     dict(config_start_time=start.encode(), config_dt=np.float64(config_dt), config_lsm_scheme=b"noah",
                          config_microp_scheme=b"mp_nssl2m", config_convection_scheme=b"cu_grell_freitas"),
                     vert_names=("vorticity",), valid=valid)
    c = wl.cfg
    nl = os.path.join(rundir, "namelist.files")
    tf = lambda b: ".true." if b else ".false."  # noqa: E731
    lines = ["&config", f' grid_file_input_grid = "{paths["init"]}"', f' hist_file_input_grid = "{paths["history"]}"',
             f' diag_file_input_grid = "{paths["diag"]}"', f' output_file = "{paths["out"]}"',
             f" interp_diag = {tf(c.interp_diag)}", f" interp_hist = {tf(c.interp_hist)}", f" wrf_mod_vars = {tf(c.wrf_mod_vars)}",
             " esmf_log = .false.", f" nx = {c.nx}", f" ny = {c.ny}"]
    if target_file:  # define_target_grid_file: sizes and projection come from a WRF-style file
        lines = [ln for ln in lines if not ln.startswith((" nx", " ny"))]
        lines += [" target_grid_type = 'file'", f' file_target_grid = "{target_file}"']
    elif c.proj_code == 1:
        lines += [" target_grid_type = 'lambert'", f" dx = {c.dx}", f" dy = {c.dy}", f" ref_lat = {c.ref_lat}",
                  f" ref_lon = {c.ref_lon}", f" truelat1 = {c.truelat1}", f" truelat2 = {c.truelat2}", f" stand_lon = {c.stand_lon}"]
    elif c.proj_code == 0:
        lines += [" target_grid_type = 'lat-lon'", f" is_regional = {tf(c.is_regional)}", f" stand_lon = {c.stand_lon}"]
        if c.is_regional:
            lines += [f" ref_lat = {c.ref_lat}", f" ref_lon = {c.ref_lon}", f" dx = {c.dx}", f" dy = {c.dy}"]
    else:
        raise NotImplementedError("write_case: Lambert and lat-lon workloads only")
    with open(nl, "w") as fh:
        fh.write("\n".join(lines) + "\n/\n")
    return nl, paths
