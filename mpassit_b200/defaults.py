"""Default variable lists of the reference's run directory, as data.

The reference reads four two-column text files with fixed names from the CWD
(`diaglist`, `histlist_2d`, `histlist_3d`, `histlist_soil`; input_data.F90:276,
851-855): column 1 = MPAS variable, column 2 = output (WRF-style) name.  These
tuples are the stock lists of /root/reference/parm/ so that benches and tests can
materialise the same files (`write_varlists`) without shipping copies of them.
"""
from __future__ import annotations

import os

DIAGLIST = [
    ("rainc", "RAINC"), ("rainnc", "RAINNC"), ("snowncv", "SNOWNCV"), ("rainncv", "RAINNCV"),
    ("graupelncv", "GRAUPELNCV"), ("prec_acc_c", "PREC_ACC_C"), ("prec_acc_nc", "PREC_ACC_NC"),
    ("snow_acc_nc", "SNOW_ACC_NC"), ("refl10cm", "REFL_10CM"), ("refl10cm_max", "COMPOSITE_REFL_10CM"),
    ("refl10cm_1km", "REFL_10CM_1KM"), ("refl10cm_1km_max", "REFL_10CM_1KM_MAX"), ("u10", "U10"), ("v10", "V10"),
    ("q2", "Q2"), ("t2m", "T2"), ("th2m", "TH2"), ("updraft_helicity_max", "UP_HELI_MAX"),
    ("w_velocity_max", "W_UP_MAX"),
]
HISTLIST_2D = [
    ("surface_pressure", "PSFC"), ("xland", "XLAND"), ("skintemp", "TSK"), ("snow", "SNOW"), ("snowh", "SNOWH"),
    ("sst", "SST"),
]
HISTLIST_3D = [
    ("zgrid", "PHB"), ("w", "W"), ("theta", "T"), ("uReconstructZonal", "U"), ("uReconstructMeridional", "V"),
    ("qv", "QVAPOR"), ("qc", "QCLOUD"), ("qr", "QRAIN"), ("qi", "QICE"), ("qs", "QSNOW"), ("qg", "QGRAUP"),
    ("ni", "QNICE"), ("nr", "QNRAIN"), ("pressure", "P_HYD"), ("rho", "MUB"),
]
HISTLIST_SOIL = [("tslb", "TSLB"), ("smois", "SMOIS"), ("sh2o", "SH2O")]

FILES = {"diaglist": DIAGLIST, "histlist_2d": HISTLIST_2D, "histlist_3d": HISTLIST_3D, "histlist_soil": HISTLIST_SOIL}


def write_varlists(directory: str, lists: dict | None = None) -> dict:
    """Write the four var-list files (tab separated, as the reference's own) into `directory`."""
    os.makedirs(directory, exist_ok=True)
    out = {}
    for fname, items in (lists or FILES).items():
        path = os.path.join(directory, fname)
        with open(path, "w") as fh:
            for a, b in items:
                fh.write(f"{a}\t\t{b}\n")
        out[fname] = path
    return out


NAMELIST_CONUS_3KM = """&config
  grid_file_input_grid = "{grid}"
  hist_file_input_grid = "{hist}"
  diag_file_input_grid = "{diag}"
  output_file = "{out}"
  interp_diag = .true.
  interp_hist = .true.
  wrf_mod_vars = .true.
  esmf_log = .false.
  nx = {nx}
  ny = {ny}
  dx = {dx}
  dy = {dx}
  ref_lat = 38.5
  ref_lon = -97.5
  truelat1 = 38.5
  truelat2 = 38.5
  stand_lon = -97.5
  target_grid_type = 'lambert'
/
"""


def write_namelist(path: str, nx: int = 1801, ny: int = 1061, dx: float = 3000.0, **kw) -> str:
    """A well-formed &config namelist of the shape of the reference's sample run
    (README.md:60-72: 3-km CONUS Lambert target, nx/ny = staggered counts)."""
    txt = NAMELIST_CONUS_3KM.format(grid=kw.get("grid", "mpas.init.nc"), hist=kw.get("hist", "mpas.history.nc"),
                                    diag=kw.get("diag", "mpas.diag.nc"), out=kw.get("out", "mpassit_out.nc"),
                                    nx=nx, ny=ny, dx=dx)
    with open(path, "w") as fh:
        fh.write(txt)
    return path
