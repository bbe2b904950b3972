"""Synthetic MPASSIT run set-ups (bench / smoke infrastructure; no arithmetic of the path).

A workload is what one `mpassit namelist.input` run hands to interp_data: the
&config namelist and the four var-lists (written to a scratch run directory and read
back through the host mirror, i.e. through the user contract), an MPAS-shaped
source mesh, the projected target grid, and one source array per listed variable.
BASELINE.json configs:
  c1  120-km quasi-uniform global mesh (40,962 cells, 55 levels) -> 1 deg lat-lon
  c2  3-km regional mesh (~2.4 M cells, 60 levels) -> Lambert 1801 x 1061, dx = 3 km,
      diaglist + histlist_2d/3d/soil  (the configuration the metric is quoted on)
  c3  the c2 geometry with every 3-D + soil field stacked into ONE batched apply on the bilinear route:
      11 x 60 + 2 x 61 + 2 x 60 + 3 x 4 = 914 level-columns (stacked_fields / run_stacked)
  c4  15-3 km variable-resolution global mesh (6,496,362 cells) -> 0.03 deg global lat-lon (12000 x 6000),
      one bilinear 55-level field + one nearest-neighbour integer field (CENTER stagger only)
  c5  1-km regional mesh -> 1-km Lambert 1001 x 1001, conservative snow fields, weights rebuilt per run
  mini  a 30-km, 8-level miniature of c2 (smoke / tests)
"""
from __future__ import annotations

import os
import tempfile
from dataclasses import dataclass, field

import numpy as np

from . import defaults, host, synth
from . import lib as L


@dataclass
class Workload:
    name: str
    cfg: host.Config
    mesh: synth.MpasMesh
    grids: dict          # "M","U","V","CORNER" -> (lat, lon) [nj][ni] degrees
    cosa: np.ndarray | None
    sina: np.ndarray | None
    nz: int
    nsoil: int
    rundir: str
    lists: dict = field(default_factory=dict)   # diag / hist_2d / hist_3d / soil -> [(mpas, target)]

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
        """Output values of one interp_data pass (target points x levels x fields), zero-filled points included."""
        nM, nU, nV = self.n_mass, self.grids["U"][0].size, self.grids["V"][0].size
        u = 0
        wrf = bool(self.cfg.wrf_mod_vars)
        for g in ("diag", "hist_2d", "hist_3d", "soil"):
            for nm, _ in self.lists[g]:
                if g == "hist_3d" and wrf and nm in ("uReconstructZonal", "uReconstructMeridional"):
                    # mass-point wind (UMASS/VMASS) + its staggered field
                    u += self.nz * (nM + (nU if nm == "uReconstructZonal" else nV))
                else:
                    u += self.levels_of(g, nm) * nM
        return u + nM  # + HGT


_SPECS = {
    #        mesh builder kwargs                                  namelist                      nz  nsoil
    "c2": (dict(spacing_m=3000.0, extent_x_m=5600e3, extent_y_m=3400e3), dict(nx=1801, ny=1061, dx=3000.0), 60, 4),
    "mid": (dict(spacing_m=12000.0, extent_x_m=5600e3, extent_y_m=3400e3), dict(nx=451, ny=266, dx=12000.0), 60, 4),
    # BASELINE.json configs[4]: 1-km regional mesh -> 1-km Lambert 1001 x 1001 (conservative snow fields)
    "c5": (dict(spacing_m=1000.0, extent_x_m=1100e3, extent_y_m=1100e3), dict(nx=1001, ny=1001, dx=1000.0), 60, 4),
    "mini": (dict(spacing_m=30000.0, extent_x_m=2000e3, extent_y_m=1400e3), dict(nx=61, ny=41, dx=30000.0), 8, 4),
}


def make(name: str = "c2", rundir: str | None = None, seed: int = synth.SEED, cell_order: str | None = None) -> Workload:
    """cell_order (or $MPASSIT_CELL_ORDER): rowmajor (as generated) | morton | random -- see synth.renumber_cells."""
    rundir = rundir or tempfile.mkdtemp(prefix=f"mpassit_{name}_")
    cell_order = cell_order or os.environ.get("MPASSIT_CELL_ORDER", "rowmajor")
    if name == "c4":
        mesh = synth.variable_geodesic_mesh(int(os.environ.get("MPASSIT_C4_FREQ", "806")))
        nl = os.path.join(rundir, "namelist.input")
        with open(nl, "w") as fh:
            fh.write("&config\n target_grid_type = 'lat-lon'\n nx = 12001\n ny = 6001\n is_regional = .false.\n"
                     " stand_lon = -180.\n interp_diag = .false.\n interp_hist = .true.\n wrf_mod_vars = .false.\n/\n")
        paths = defaults.write_varlists(rundir)
        cfg = host.read_setup_namelist(nl)
        # only the CENTER stagger is needed (one bilinear 3-D field + one nearest 2-D field): 72 M points
        grids = {"M": host.target_coords(cfg, L.CENTER)}
        lists = {"diag": [], "hist_2d": [("ivgtyp", "IVGTYP")], "hist_3d": [("theta", "T")], "soil": []}
        return Workload(name, cfg, mesh, grids, None, None, 55, 4, rundir, lists)
    if name == "c1":
        mesh = synth.global_mesh(40962)
        nl = os.path.join(rundir, "namelist.input")
        with open(nl, "w") as fh:
            fh.write("&config\n target_grid_type = 'lat-lon'\n nx = 361\n ny = 181\n is_regional = .false.\n"
                     " stand_lon = -180.\n interp_diag = .false.\n interp_hist = .true.\n wrf_mod_vars = .true.\n/\n")
        nz, nsoil = 55, 4
    else:
        mk, nk, nz, nsoil = _SPECS["c2" if name == "c3" else name]
        mesh = synth.renumber_cells(synth.regional_hex_mesh(seed=seed, **mk), cell_order, seed)
        nl = defaults.write_namelist(os.path.join(rundir, "namelist.input"), **nk)
    paths = defaults.write_varlists(rundir)
    cfg = host.read_setup_namelist(nl)
    lists = {"diag": host.read_varlist(paths["diaglist"]), "hist_2d": host.read_varlist(paths["histlist_2d"]),
             "hist_3d": host.read_varlist(paths["histlist_3d"]), "soil": host.read_varlist(paths["histlist_soil"])}
    if not cfg.interp_diag:
        lists["diag"] = []
    grids = {k: host.target_coords(cfg, s) for k, s in (("M", L.CENTER), ("U", L.EDGE1), ("V", L.EDGE2), ("CORNER", L.CORNER))}
    cosa = sina = None
    if cfg.proj_code == host.PROJ_LC:
        cosa, sina = host.get_rotang(*grids["M"])
    return Workload(name, cfg, mesh, grids, cosa, sina, nz, nsoil, rundir, lists)


def load_geometry(rg, wl: Workload) -> None:
    """mprg_set_mesh + mprg_set_target for every stagger (define_input_grid / define_target_grid)."""
    m = wl.mesh
    rg.set_mesh(m.lonCell, m.latCell, m.lonVertex, m.latVertex, m.verticesOnCell)
    rg.set_grid_kind(L.GRID_NOPERI if wl.cfg.is_regional else L.GRID_1PERI_MONOPOLE)   # model_grid.F90:684-703
    for k, s in (("M", L.CENTER), ("U", L.EDGE1), ("V", L.EDGE2), ("CORNER", L.CORNER)):
        if k not in wl.grids:
            continue
        lat, lon = wl.grids[k]
        rg.set_target(s, lon, lat)
    if wl.cosa is not None:  # cosa/sina_target_grid, model_grid.F90:1113-1185
        rg.set_rotation(wl.cosa, wl.sina)


def _field_values_torch(wl: Workload, group: str, name: str, nlev: int, k: int, device):
    """Synthetic values per SURVEY.md §8(d), generated on `device` ([nCells][nlev] fp32)."""
    import torch

    m = wl.mesh
    lon = torch.from_numpy(m.lonCell).to(device=device, dtype=torch.float32)
    lat = torch.from_numpy(m.latCell).to(device=device, dtype=torch.float32)
    g = torch.Generator(device=device)
    g.manual_seed(synth.SEED + 1000 * ("diag", "hist_2d", "hist_3d", "soil").index(group) + k)
    if name in ("xland", "ivgtyp", "isltyp", "landmask") or group == "soil":
        nmax = {"xland": 2, "landmask": 2}.get(name, 20)
        return torch.randint(1, nmax + 1, (m.nCells, nlev), generator=g, device=device).to(torch.float32)
    base = 280.0 + 20.0 * torch.sin(lat) + 5.0 * torch.sin(3.0 * lon) * torch.cos(2.0 * lat)
    lev = torch.arange(nlev, device=device, dtype=torch.float32)
    f = base[:, None] - 0.1 * lev[None, :] + 0.5 * torch.randn((m.nCells, nlev), generator=g, device=device)
    if name in ("snow", "snowh") or name.startswith("q") or name in ("ni", "nr") or name.startswith("rain"):
        f = torch.clamp(f - 285.0, min=0.0) * (100.0 if name.startswith("snow") else 1e-3)
    return f.contiguous()


def make_fields(wl: Workload, device="cuda", pinned_host: bool = False, rg=None) -> dict:
    """Source arrays + destination slabs for one interp_data pass.

    device fields: torch CUDA tensors.  pinned_host=True additionally returns page-locked
    host copies (numpy views) for the end-to-end (host-buffer) path.  Destination slabs are
    sized for rank `rg` (full grid when rg is None / single rank)."""
    import torch

    nM, nU, nV = wl.n_mass, wl.grids["U"][0].size, wl.grids["V"][0].size
    niM, niU, niV = wl.grids["M"][0].shape[1], wl.grids["U"][0].shape[1], wl.grids["V"][0].shape[1]
    if rg is not None and rg.nranks > 1:
        j0, j1 = rg.slab(L.CENTER)
        nM = (j1 - j0) * niM
        j0, j1 = rg.slab(L.EDGE1)
        nU = (j1 - j0) * niU
        j0, j1 = rg.slab(L.EDGE2)
        nV = (j1 - j0) * niV
    out = {"dev": {}, "host": {}, "sizes": (nM, nU, nV)}

    def host_copy(t):
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        return h

    for where in (["dev", "host"] if pinned_host else ["dev"]):
        groups = {}
        for g in ("diag", "hist_2d", "hist_3d", "soil"):
            specs = []
            for k, (nm, tn) in enumerate(wl.lists[g]):
                nlev = wl.levels_of(g, nm)
                if where == "dev":
                    src = _field_values_torch(wl, g, nm, nlev, k, device)
                    dst = torch.empty((nlev, nM), dtype=torch.float32, device=device)
                else:
                    src = host_copy(out["dev"][g][k].src)
                    dst = torch.empty((nlev, nM), dtype=torch.float32, pin_memory=True)
                specs.append(host.FieldSpec(nm, tn, nlev, src, dst))
            groups[g] = specs
        if where == "dev":
            ter = _field_values_torch(wl, "hist_2d", "ter", 1, 99, device)
            extra = dict(ter=ter, hgt=torch.empty((1, nM), dtype=torch.float32, device=device),
                         u_stag=torch.empty((wl.nz, nU), dtype=torch.float32, device=device),
                         v_stag=torch.empty((wl.nz, nV), dtype=torch.float32, device=device))
        else:
            extra = dict(ter=host_copy(out["dev"]["ter"]), hgt=torch.empty((1, nM), dtype=torch.float32, pin_memory=True),
                         u_stag=torch.empty((wl.nz, nU), dtype=torch.float32, pin_memory=True),
                         v_stag=torch.empty((wl.nz, nV), dtype=torch.float32, pin_memory=True))
        groups.update(extra)
        out[where] = groups
    return out


def stacked_fields(wl: Workload) -> list[tuple[str, int]]:
    """BASELINE.json configs[2]: every histlist_3d field (winds included, as plain fields) and every histlist_soil
    field in ONE batched apply on the bilinear route: (name, levels); 914 level-columns on the stock lists."""
    return ([(nm, wl.levels_of("hist_3d", nm)) for nm, _ in wl.lists["hist_3d"]] +
            [(nm, wl.nsoil) for nm, _ in wl.lists["soil"]])


def make_stacked(wl: Workload, device="cuda"):
    """(names, sources [nCells][nlev], destinations [nlev][nMass]) of the c3 stack, on `device`."""
    import torch

    names, srcs, dsts = [], [], []
    for k, (nm, nlev) in enumerate(stacked_fields(wl)):
        grp = "hist_3d" if k < len(wl.lists["hist_3d"]) else "soil"
        src = _field_values_torch(wl, grp, nm, nlev, k, device)
        if grp == "soil":   # bilinear here: use smooth values, not class labels
            src = _field_values_torch(wl, "hist_3d", "soil_" + nm, nlev, 50 + k, device)
        names.append(nm)
        srcs.append(src)
        dsts.append(torch.empty((nlev, wl.n_mass), dtype=torch.float32, device=device))
    return names, srcs, dsts


def _np(x):
    """numpy view of a (pinned) host torch tensor, or the tensor itself if on device."""
    if isinstance(x, int):
        return x
    if type(x).__module__.startswith("torch") and not x.is_cuda:
        return x.numpy()
    return x


def run_interp(rg, wl: Workload, F: dict, mem: int, dst_full: bool = False):
    """One interp_data pass (interp.F90:92) through the host mirror + C ABI.  dst_full: the destinations in
    F are full-grid fields (the writing rank's, own or mapped) and every rank stores its rows into them."""
    def conv(specs):
        return [host.FieldSpec(s.name, s.target_name, s.nlev, _np(s.src), _np(s.dst)) for s in specs]

    return prepare_interp(rg, wl, F, mem, dst_full)()


def prepare_interp(rg, wl: Workload, F: dict, mem: int, dst_full: bool = False):
    """The pass of run_interp with its argument block marshalled once (host.PreparedInterp): a time loop
    over many output times pays one C call per pass, as a compiled host would."""
    def conv(specs):
        return [host.FieldSpec(s.name, s.target_name, s.nlev, _np(s.src), _np(s.dst)) for s in specs]

    keep = [conv(F[g]) for g in ("diag", "hist_2d", "hist_3d", "soil")]
    p = host.prepare_interp(rg, wl.cfg, diag=keep[0], hist_2d=keep[1], hist_3d=keep[2], soil=keep[3], ter=_np(F["ter"]),
                            hgt=_np(F["hgt"]), u_stag=_np(F["u_stag"]), v_stag=_np(F["v_stag"]), nz=wl.nz, mem=mem,
                            dst_full=dst_full)
    p._fields = (keep, F)  # the buffers behind the raw pointers stay alive with the call
    return p


# ---- fused gather: full-grid outputs on the writing rank, mapped into the other ranks with CUDA IPC --------
def full_outputs(wl: Workload, device) -> dict:
    """Full-grid output fields [nlev][nj*ni] of one pass (allocated on the writing rank)."""
    import torch

    nM, nU, nV = wl.n_mass, wl.grids["U"][0].size, wl.grids["V"][0].size
    out = {g: [torch.empty((wl.levels_of(g, nm), nM), dtype=torch.float32, device=device) for nm, _ in wl.lists[g]]
           for g in ("diag", "hist_2d", "hist_3d", "soil")}
    out["hgt"] = torch.empty((1, nM), dtype=torch.float32, device=device)
    out["u_stag"] = torch.empty((wl.nz, nU), dtype=torch.float32, device=device)
    out["v_stag"] = torch.empty((wl.nz, nV), dtype=torch.float32, device=device)
    return out


def export_full(rg, full: dict) -> dict:
    """Picklable (handle, offset) pairs for every field of `full` (mprg_ipc_export)."""
    return {k: ([rg.ipc_export(t) for t in v] if isinstance(v, list) else rg.ipc_export(v)) for k, v in full.items()}


def open_full(rg, exported: dict) -> dict:
    """Device addresses of the writing rank's fields in this process (mprg_ipc_open)."""
    return {k: ([rg.ipc_open(*h) for h in v] if isinstance(v, list) else rg.ipc_open(*v)) for k, v in exported.items()}


def with_destinations(F: dict, dst: dict) -> dict:
    """The field set F with every destination replaced by the matching entry of `dst`."""
    out = {}
    for g in ("diag", "hist_2d", "hist_3d", "soil"):
        out[g] = [host.FieldSpec(s.name, s.target_name, s.nlev, s.src, d) for s, d in zip(F[g], dst[g])]
    out["ter"] = F["ter"]
    for k in ("hgt", "u_stag", "v_stag"):
        out[k] = dst[k]
    return out


def io_bytes(wl: Workload, F: dict) -> tuple[int, int]:
    """(host->device, device->host) bytes of one host-buffer pass, counted from the tensors copied."""
    h2d = d2h = 0
    wrf = bool(wl.cfg.wrf_mod_vars)
    for g in ("diag", "hist_2d", "hist_3d", "soil"):
        for s in F[g]:
            h2d += s.src.numel() * 4
            if not (g == "hist_3d" and wrf and s.name in ("uReconstructZonal", "uReconstructMeridional")):
                d2h += s.dst.numel() * 4
    h2d += F["ter"].numel() * 4
    d2h += F["hgt"].numel() * 4 + F["u_stag"].numel() * 4 + F["v_stag"].numel() * 4
    return h2d, d2h
