"""The bench workloads for the CPU reference arm, built WITHOUT the product libraries.

`bench.py --impl reference` times the CPU restatement of the reference's path and must not map
libmpassit_rg.so / libmpassit_host.so: the namelist values are taken from the workload table, the
target coordinates come from oracle/proj_oracle.py (numpy restatement of the WPS formulas,
module_map_utils.F90:1083-1233, 1398-1428; held to the C++ host mirror by tests/test_host.py) and the
mesh from mpassit_b200/synth.py (pure numpy).  TEST / BENCH INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from types import SimpleNamespace

import numpy as np

from mpassit_b200 import defaults, synth

from . import proj_oracle as po

# name -> (mesh kwargs, (nx, ny, dx), nz, nsoil)   == mpassit_b200/workload.py:_SPECS
SPECS = {
    "c2": (dict(spacing_m=3000.0, extent_x_m=5600e3, extent_y_m=3400e3), (1801, 1061, 3000.0), 60, 4),
    "c3": (dict(spacing_m=3000.0, extent_x_m=5600e3, extent_y_m=3400e3), (1801, 1061, 3000.0), 60, 4),
    "mid": (dict(spacing_m=12000.0, extent_x_m=5600e3, extent_y_m=3400e3), (451, 266, 12000.0), 60, 4),
    "c5": (dict(spacing_m=1000.0, extent_x_m=1100e3, extent_y_m=1100e3), (1001, 1001, 1000.0), 60, 4),
    "mini": (dict(spacing_m=30000.0, extent_x_m=2000e3, extent_y_m=1400e3), (61, 41, 30000.0), 8, 4),
}


@dataclass
class RefWorkload:
    name: str
    cfg: SimpleNamespace
    mesh: synth.MpasMesh
    grids: dict
    cosa: np.ndarray | None
    sina: np.ndarray | None
    nz: int
    nsoil: int
    lists: dict = field(default_factory=dict)

    @property
    def n_mass(self) -> int:
        return self.grids["M"][0].size

    def levels_of(self, group: str, name: str) -> int:
        if group == "diag":
            return self.nz if name == "refl10cm" else 1
        if group == "hist_2d":
            return 1
        if group == "hist_3d":
            return self.nz + 1 if name in ("zgrid", "w") else self.nz
        return self.nsoil

    def units_per_pass(self) -> int:
        nM, nU, nV = self.n_mass, self.grids["U"][0].size, self.grids["V"][0].size
        u = 0
        for g in ("diag", "hist_2d", "hist_3d", "soil"):
            for nm, _ in self.lists[g]:
                if g == "hist_3d" and self.cfg.wrf_mod_vars and nm in ("uReconstructZonal", "uReconstructMeridional"):
                    u += self.nz * (nM + (nU if nm == "uReconstructZonal" else nV))
                else:
                    u += self.levels_of(g, nm) * nM
        return u + nM


def make(name: str = "c2", seed: int = synth.SEED, cell_order: str = "rowmajor") -> RefWorkload:
    lists = {"diag": list(defaults.DIAGLIST), "hist_2d": list(defaults.HISTLIST_2D), "hist_3d": list(defaults.HISTLIST_3D),
             "soil": list(defaults.HISTLIST_SOIL)}
    if name == "c1":
        mesh = synth.global_mesh(40962)
        nx, ny, nz, nsoil = 361, 181, 55, 4
        cfg = SimpleNamespace(nx=nx, ny=ny, dxkm=0.0, wrf_mod_vars=1, is_regional=0, proj="lat-lon")
        grids = {k: po.latlon_global_grid(nx, ny, -180.0, k) for k in ("M", "U", "V", "CORNER")}
        lists["diag"] = []
        return RefWorkload(name, cfg, mesh, grids, None, None, nz, nsoil, lists)
    mk, (nx, ny, dx), nz, nsoil = SPECS[name]
    mesh = synth.renumber_cells(synth.regional_hex_mesh(seed=seed, **mk), cell_order, seed)
    cfg = SimpleNamespace(nx=nx, ny=ny, dxkm=dx, wrf_mod_vars=1, is_regional=1, proj="lambert")
    grids = {k: po.lc_grid(nx, ny, dx, 38.5, -97.5, 38.5, 38.5, -97.5, k) for k in ("M", "U", "V", "CORNER")}
    cosa, sina = po.rotang(*grids["M"])
    return RefWorkload(name, cfg, mesh, grids, cosa, sina, nz, nsoil, lists)


def workload_name(wl) -> str:
    """The `config.workload` string shared by both arms of bench.py."""
    if wl.name == "c1":
        return "c1: 120-km global MPAS (40962 cells, 55 levels) -> 1 deg lat-lon, histlist_2d/3d/soil, one interp_data pass"
    return (f"{wl.name}: 3-km regional MPAS ({wl.mesh.nCells} cells, {wl.nz} levels) -> Lambert {wl.cfg.nx}x{wl.cfg.ny} "
            f"dx={wl.cfg.dxkm:.0f} m, diaglist+histlist_2d/3d/soil, one interp_data pass")


def synthetic_sources(wl) -> dict:
    """One source array per listed variable (values by the SURVEY.md 8(d) recipes; arrays of equal shape are
    shared -- the CPU arm's timing does not depend on the values)."""
    m = wl.mesh
    cache = {}

    def arr(n):
        if n not in cache:
            cache[n] = synth.smooth_field(m.lonCell, m.latCell, n, seed=n)
        return cache[n]

    out = {g: [(nm, arr(wl.levels_of(g, nm))) for nm, _ in wl.lists[g]] for g in ("diag", "hist_2d", "hist_3d", "soil")}
    out["ter"] = arr(1)
    return out
