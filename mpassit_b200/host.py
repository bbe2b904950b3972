"""ctypes binding of libmpassit_host.so (include/mpassit_host.h): the C++ mirror of the
reference's Fortran host stages -- ``read_setup_namelist`` (program_setup.F90:87),
``read_varlist`` (input_data.F90:1146), the regrid-class tables
(input_data.F90:840-911), ``define_target_grid_params`` coordinates
(model_grid.F90:644-1201), ``get_rotang`` (model_grid.F90:2450) and ``interp_data``
(interp.F90:92-465).  Python adds nothing but marshalling.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import lib as _l

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libmpassit_host.so")
STRLEN, NAMELEN = 512, 64
NAN = 1.0e20
PROJ_LATLON, PROJ_LC, PROJ_PS, PROJ_MERC = 0, 1, 2, 3
(CLASS_2D_PATCH, CLASS_2D_CONS, CLASS_2D_NSTD, CLASS_3D_NZ, CLASS_3D_NZP1, CLASS_3D_VERT, CLASS_U, CLASS_V,
 CLASS_SOIL, CLASS_DIAG_2D, CLASS_DIAG_3D) = range(11)


class HostError(RuntimeError):
    """error_handler(msg, rc) of the reference (utils.F90:16-33), minus the mpi_abort."""

    def __init__(self, rc, msg):
        super().__init__(f"rc={rc}: {msg}")
        self.rc, self.msg = rc, msg


class Config(C.Structure):
    _fields_ = ([(n, C.c_char * STRLEN) for n in
                 ("grid_file_input_grid", "diag_file_input_grid", "hist_file_input_grid", "file_target_grid",
                  "output_file", "block_decomp_file", "target_grid_type")] +
                [(n, C.c_int32) for n in
                 ("interp_diag", "interp_hist", "wrf_mod_vars", "esmf_log", "is_regional", "interp_as_bundle", "nx", "ny")] +
                [(n, C.c_double) for n in
                 ("dx", "dy", "ref_lat", "ref_lon", "ref_x", "ref_y", "truelat1", "truelat2", "stand_lon", "pole_lat",
                  "pole_lon")] +
                [(n, C.c_int32) for n in ("i_target", "j_target", "proj_code")] +
                [(n, C.c_double) for n in
                 ("dxkm", "dykm", "dlondeg", "dlatdeg", "known_lat", "known_lon", "known_x", "known_y")] +
                [("map_proj_char", C.c_char * NAMELEN)])


class Field(C.Structure):
    _fields_ = [("name", C.c_char * NAMELEN), ("target_name", C.c_char * NAMELEN), ("nlev", C.c_int32),
                ("klass", C.c_int32), ("src", C.c_void_p), ("dst", C.c_void_p)]


class InterpIO(C.Structure):
    _fields_ = [("src_dtype", C.c_int32), ("dst_dtype", C.c_int32), ("mem", C.c_int32), ("nz", C.c_int32),
                ("n_diag", C.c_int32), ("diag", C.POINTER(Field)),
                ("n_hist_2d", C.c_int32), ("hist_2d", C.POINTER(Field)),
                ("n_hist_3d", C.c_int32), ("hist_3d", C.POINTER(Field)),
                ("n_soil", C.c_int32), ("soil", C.POINTER(Field)),
                ("ter", C.c_void_p), ("hgt", C.c_void_p), ("u_stag", C.c_void_p), ("v_stag", C.c_void_p),
                ("dst_full", C.c_int32), ("dst_device", C.c_int32)]


class RunStats(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("setup_ms", "read_ms", "interp_ms", "write_ms", "total_ms", "init_ms", "target_ms",
                                          "gridfile_ms", "mesh_ms", "alloc_ms", "download_ms", "writer_wait_ms")] + \
               [(n, C.c_int64) for n in ("n_cells", "bytes_in", "bytes_out")] + \
               [("n_vars_written", C.c_int32), ("output_version", C.c_int32), ("p_top", C.c_double)]


COMM_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int)
COMM_BARRIER, COMM_MAX, COMM_MIN, COMM_ABORT = 0, 1, 2, 3

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(HOST_LIB_PATH):
        raise HostError(-1, f"{HOST_LIB_PATH} not built; run `python -m mpassit_b200.build`")
    _l.load()  # the engine first (resolved by rpath anyway)
    L = C.CDLL(HOST_LIB_PATH)
    cp, i32 = C.c_char_p, C.c_int32
    L.mpassit_read_setup_namelist.argtypes = [cp, C.POINTER(Config), cp, C.c_size_t]
    L.mpassit_read_varlist.argtypes = [cp, i32, C.POINTER(i32), cp, cp, cp, C.c_size_t]
    L.mpassit_classify_hist_2d.argtypes = [cp]
    L.mpassit_classify_hist_3d.argtypes = [cp, C.c_int]
    L.mpassit_classify_diag.argtypes = [cp]
    L.mpassit_para_range.argtypes = [i32, i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    L.mpassit_para_range.restype = None
    L.mpassit_read_block_decomp_file.argtypes = [cp, i32, i32, C.c_void_p, cp, C.c_size_t]
    L.mpassit_target_dims.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(i32), C.POINTER(i32)]
    L.mpassit_projection.argtypes = [C.POINTER(Config), C.POINTER(_l.Projection), cp, C.c_size_t]
    L.mpassit_target_coords.argtypes = [C.POINTER(Config), C.c_int, C.c_void_p, C.c_void_p, cp, C.c_size_t]
    L.mpassit_get_rotang.argtypes = [C.c_void_p, C.c_void_p, i32, i32, C.c_void_p, C.c_void_p]
    L.mpassit_get_rotang.restype = None
    L.mpassit_get_cell_corners.argtypes = [C.c_void_p, C.c_void_p, i32, i32, C.c_double, C.c_void_p, C.c_void_p]
    L.mpassit_get_cell_corners.restype = None
    L.mpassit_classify_fields.argtypes = [C.POINTER(Config), C.POINTER(InterpIO)] + [C.POINTER(i32)] * 4
    L.mpassit_interp_data.argtypes = [C.c_void_p, C.POINTER(Config), C.POINTER(InterpIO), cp, C.c_size_t]
    L.mpassit_xytoll.argtypes = [C.POINTER(Config), C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.mpassit_get_map_factor.argtypes = [C.POINTER(Config), C.c_void_p, C.c_int64, C.c_void_p]
    L.mpassit_run.argtypes = [cp, cp, C.c_int, C.c_int, C.c_int, COMM_FN, C.c_void_p, C.POINTER(RunStats), cp, C.c_size_t]
    L.mpassit_nc_describe.argtypes = [cp, cp, C.c_size_t, cp, C.c_size_t]
    L.mpassit_nc_get.argtypes = [cp, cp, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, cp, C.c_size_t]
    L.mpassit_nc_copy.argtypes = [cp, cp, C.c_int, cp, C.c_size_t]
    i64p = C.POINTER(C.c_int64)
    L.mpassit_weights_sizes.argtypes = [cp, i64p, i64p, i64p, cp, C.c_size_t]
    L.mpassit_weights_read_csr.argtypes = [cp, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, cp, C.c_size_t]
    L.mpassit_weights_write.argtypes = [cp, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, cp, cp, C.c_size_t]
    _lib = L
    return L


def _err():
    return C.create_string_buffer(2048)


def read_setup_namelist(path: str) -> Config:
    cfg, e = Config(), _err()
    rc = load().mpassit_read_setup_namelist(path.encode(), C.byref(cfg), e, len(e))
    if rc:
        raise HostError(rc, e.value.decode())
    return cfg


def projection(cfg: Config):
    """The per-grid scalars of the target projection (map_set / set_lc) for Regridder.set_target_projected."""
    p, e = _l.Projection(), _err()
    rc = load().mpassit_projection(C.byref(cfg), C.byref(p), e, len(e))
    if rc:
        raise HostError(rc, e.value.decode())
    return p


def read_varlist(path: str) -> list[tuple[str, str]]:
    n, e = C.c_int32(), _err()
    cap = 512
    a = C.create_string_buffer(cap * NAMELEN)
    b = C.create_string_buffer(cap * NAMELEN)
    rc = load().mpassit_read_varlist(path.encode(), cap, C.byref(n), a, b, e, len(e))
    if rc:
        raise HostError(rc, e.value.decode())
    out = []
    for k in range(n.value):
        out.append((a.raw[k * NAMELEN:(k + 1) * NAMELEN].split(b"\0")[0].decode(),
                    b.raw[k * NAMELEN:(k + 1) * NAMELEN].split(b"\0")[0].decode()))
    return out


def classify_hist_2d(name: str) -> int:
    return load().mpassit_classify_hist_2d(name.encode())


def classify_hist_3d(name: str, wrf_mod_vars: bool) -> int:
    return load().mpassit_classify_hist_3d(name.encode(), int(wrf_mod_vars))


def classify_diag(name: str) -> int:
    return load().mpassit_classify_diag(name.encode())


def para_range(n1: int, n2: int, nprocs: int, irank: int) -> tuple[int, int]:
    a, b = C.c_int32(), C.c_int32()
    load().mpassit_para_range(n1, n2, nprocs, irank, C.byref(a), C.byref(b))
    return a.value, b.value


def slab_rows(nj: int, nranks: int, rank: int) -> tuple[int, int]:
    """Target rows [j0, j1) (0-based) owned by `rank`: para_range over 1..nj, model_grid.F90:2428."""
    a, b = para_range(1, nj, nranks, rank)
    return a - 1, b


def gather_runs(ni: int, nj: int, nlev: int, nranks: int, rank: int) -> list[tuple[int, int, int]]:
    """Placement of one rank's slab [nlev][j1-j0][ni] inside the gathered field [nlev][nj][ni]
    (what ESMF_FieldGather does for the reference, write_data.F90:1006-1453): one contiguous
    run per level, as (offset in the slab, offset in the full field, element count).  The
    engine's mprg_gather posts exactly these runs as grouped ncclSend/ncclRecv pairs."""
    j0, j1 = slab_rows(nj, nranks, rank)
    n = (j1 - j0) * ni
    if n <= 0:
        return []
    return [(l * n, l * nj * ni + j0 * ni, n) for l in range(nlev)]


def read_block_decomp_file(path: str, ncells: int, npets: int) -> np.ndarray:
    owner = np.empty(ncells, np.int32)
    e = _err()
    rc = load().mpassit_read_block_decomp_file(path.encode(), ncells, npets, owner.ctypes.data, e, len(e))
    if rc:
        raise HostError(rc, e.value.decode())
    return owner


def target_dims(cfg: Config, stagger: int) -> tuple[int, int]:
    ni, nj = C.c_int32(), C.c_int32()
    load().mpassit_target_dims(C.byref(cfg), stagger, C.byref(ni), C.byref(nj))
    return ni.value, nj.value


def target_coords(cfg: Config, stagger: int) -> tuple[np.ndarray, np.ndarray]:
    """(lat, lon) [nj][ni] degrees of one stagger of the projected target grid."""
    ni, nj = target_dims(cfg, stagger)
    lat = np.empty((nj, ni), np.float64)
    lon = np.empty((nj, ni), np.float64)
    e = _err()
    rc = load().mpassit_target_coords(C.byref(cfg), stagger, lat.ctypes.data, lon.ctypes.data, e, len(e))
    if rc:
        raise HostError(rc, e.value.decode())
    return lat, lon


def get_rotang(lat: np.ndarray, lon: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    lat = np.ascontiguousarray(lat, np.float64)
    lon = np.ascontiguousarray(lon, np.float64)
    nj, ni = lat.shape
    cosa, sina = np.empty_like(lat), np.empty_like(lat)
    load().mpassit_get_rotang(lat.ctypes.data, lon.ctypes.data, ni, nj, cosa.ctypes.data, sina.ctypes.data)
    return cosa, sina


def get_cell_corners(lat: np.ndarray, lon: np.ndarray, dx: float) -> tuple[np.ndarray, np.ndarray]:
    """get_cell_corners (model_grid.F90:1902-1972): CORNER-stagger (lat, lon) [nj+1][ni+1] of a file-mode target."""
    lat = np.ascontiguousarray(lat, np.float64)
    lon = np.ascontiguousarray(lon, np.float64)
    nj, ni = lat.shape
    clat, clon = np.empty((nj + 1, ni + 1)), np.empty((nj + 1, ni + 1))
    load().mpassit_get_cell_corners(lat.ctypes.data, lon.ctypes.data, ni, nj, float(dx), clat.ctypes.data, clon.ctypes.data)
    return clat, clon


@dataclass
class FieldSpec:
    name: str
    target_name: str
    nlev: int
    src: object  # numpy array or torch CUDA tensor, [n][nlev] level-fastest (or [n] when nlev == 1)
    dst: object  # [nlev][nj_slab][ni]


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, int):  # raw device address (a mapped peer buffer)
        return x
    if type(x).__module__.startswith("torch"):
        return x.data_ptr()
    return x.ctypes.data


def _farr(specs):
    arr = (Field * max(len(specs), 1))()
    for k, s in enumerate(specs):
        arr[k].name = s.name.encode()
        arr[k].target_name = s.target_name.encode()
        arr[k].nlev = int(s.nlev)
        arr[k].src = _ptr(s.src)
        arr[k].dst = _ptr(s.dst)
    return arr


class PreparedInterp:
    """interp_data with its argument block marshalled once: calling the object is one C call
    (mpassit_interp_data), which is what a compiled host pays per output time."""

    def __init__(self, rg, cfg: Config, io: "InterpIO", keep):
        self.rg, self.cfg, self.io, self._keep = rg, cfg, io, keep
        self._fn = load().mpassit_interp_data
        self._err = _err()

    def __call__(self) -> "InterpIO":
        if self.io.mem == _l.DEVICE and not getattr(self.rg, "_torch_stream", None):
            import torch   # device tensors came from torch's stream; the engine is on its own (regrid.order_after_torch)

            torch.cuda.current_stream().synchronize()
        rc = self._fn(self.rg.ctx, C.byref(self.cfg), C.byref(self.io), self._err, len(self._err))
        if rc:
            raise HostError(rc, self._err.value.decode())
        return self.io


def prepare_interp(rg, cfg: Config, **kw) -> PreparedInterp:
    """Same arguments as interp_data; returns the reusable call."""
    io, keep = _build_io(**kw)
    return PreparedInterp(rg, cfg, io, keep)


def interp_data(rg, cfg: Config, **kw) -> InterpIO:
    """interp_data (interp.F90:92) on Regridder ``rg``; field lists are FieldSpec sequences in
    var-list order.  Returns the InterpIO (with the regrid class of every field filled in)."""
    return prepare_interp(rg, cfg, **kw)()


def _build_io(*, diag=(), hist_2d=(), hist_3d=(), soil=(), ter=None, hgt=None, u_stag=None,
              v_stag=None, nz=0, src_dtype=_l.F32, dst_dtype=_l.F32, mem=_l.HOST, dst_full=False):
    io = InterpIO()
    io.src_dtype, io.dst_dtype, io.mem, io.nz = src_dtype, dst_dtype, mem, nz
    keep = [_farr(diag), _farr(hist_2d), _farr(hist_3d), _farr(soil)]
    io.n_diag, io.diag = len(diag), keep[0]
    io.n_hist_2d, io.hist_2d = len(hist_2d), keep[1]
    io.n_hist_3d, io.hist_3d = len(hist_3d), keep[2]
    io.n_soil, io.soil = len(soil), keep[3]
    io.ter, io.hgt, io.u_stag, io.v_stag = _ptr(ter), _ptr(hgt), _ptr(u_stag), _ptr(v_stag)
    io.dst_full = int(bool(dst_full))
    io._keep = keep
    return io, keep


# ---- program mpassit with files on both sides (host/run.cpp) -------------------------------------------
def xytoll(cfg: Config, x: float, y: float, stagger: int = 1) -> tuple[float, float]:
    lat, lon = C.c_double(), C.c_double()
    load().mpassit_xytoll(C.byref(cfg), x, y, stagger, C.byref(lat), C.byref(lon))
    return lat.value, lon.value


def get_map_factor(cfg: Config, lat: np.ndarray) -> np.ndarray:
    lat = np.ascontiguousarray(lat, np.float64)
    out = np.empty_like(lat)
    load().mpassit_get_map_factor(C.byref(cfg), lat.ctypes.data, lat.size, out.ctypes.data)
    return out


def torch_comm(group=None):
    """The `comm` callback of mpassit_run on top of torch.distributed (what MPI is to the reference's host)."""
    import torch
    import torch.distributed as dist

    def fn(_arg, op, vals, n):
        if op == COMM_ABORT:
            # error_handler -> mpi_abort (utils.F90:16-33): peers may be blocked in a collective, so the whole job
            # goes down, the failing rank's code first
            import os
            import sys

            sys.stderr.write(f"mpassit_run: FATAL ERROR on a rank (rc {int(vals[0])}); aborting the job\n")
            sys.stderr.flush()
            os._exit(int(vals[0]) & 0xff or 1)
        if op == COMM_BARRIER:
            dist.barrier(group=group)
            return
        t = torch.tensor([vals[i] for i in range(n)], dtype=torch.float64)
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == COMM_MAX else dist.ReduceOp.MIN, group=group)
        for i, v in enumerate(t.cpu().tolist()):
            vals[i] = v

    return COMM_FN(fn)


def run(namelist: str, varlist_dir: str | None = None, device: int = 0, rank: int = 0, nranks: int = 1, comm=None) -> RunStats:
    """`mpassit <namelist>` (mpassit.F90:23-146): MPAS NetCDF-classic files in, one WRF-style file out."""
    st, e = RunStats(), _err()
    cb = comm if comm is not None else C.cast(None, COMM_FN)
    rc = load().mpassit_run(namelist.encode(), varlist_dir.encode() if varlist_dir else None, device, rank, nranks, cb, None,
                            C.byref(st), e, len(e))
    if rc:
        raise HostError(rc, e.value.decode())
    return st


def nc_describe(path: str) -> list[list[str]]:
    out, e = C.create_string_buffer(1 << 20), _err()
    rc = load().mpassit_nc_describe(path.encode(), out, len(out), e, len(e))
    if rc:
        raise HostError(rc, e.value.decode())
    return [ln.split(" ") for ln in out.value.decode().splitlines()]


def nc_get(path: str, var: str, n: int, rec: int = 0, first: int = 0) -> np.ndarray:
    out, e = np.empty(n, np.float64), _err()
    rc = load().mpassit_nc_get(path.encode(), var.encode(), rec, first, n, out.ctypes.data, e, len(e))
    if rc:
        raise HostError(rc, e.value.decode())
    return out


def nc_copy(src: str, dst: str, version: int = 0) -> None:
    e = _err()
    rc = load().mpassit_nc_copy(src.encode(), dst.encode(), version, e, len(e))
    if rc:
        raise HostError(rc, e.value.decode())


# ---- ESMF regrid weight files (host/weights.cpp) --------------------------------------------------------
def read_esmf_weights(path: str) -> tuple[int, int, np.ndarray, np.ndarray, np.ndarray]:
    """(n_a, n_b, rowptr[n_b+1], col[n_s] 0-based, S[n_s]) of an ESMF weight file, rows sorted, ESMF's order inside a row."""
    na, nb, ns, e = C.c_int64(), C.c_int64(), C.c_int64(), _err()
    rc = load().mpassit_weights_sizes(path.encode(), C.byref(na), C.byref(nb), C.byref(ns), e, len(e))
    if rc:
        raise HostError(rc, e.value.decode())
    rowptr = np.empty(nb.value + 1, np.int32)
    col = np.empty(ns.value, np.int32)
    w = np.empty(ns.value, np.float64)
    rc = load().mpassit_weights_read_csr(path.encode(), nb.value, ns.value, rowptr.ctypes.data, col.ctypes.data, w.ctypes.data, e, len(e))
    if rc:
        raise HostError(rc, e.value.decode())
    return na.value, nb.value, rowptr, col, w


def write_esmf_weights(path: str, n_a: int, rowptr, col, w, method: str = "Bilinear") -> None:
    rowptr = np.ascontiguousarray(rowptr, np.int32)
    col = np.ascontiguousarray(col, np.int32)
    w = np.ascontiguousarray(w, np.float64)
    e = _err()
    rc = load().mpassit_weights_write(path.encode(), n_a, rowptr.size - 1, rowptr.ctypes.data, col.ctypes.data if col.size else None,
                                      w.ctypes.data if w.size else None, method.encode(), e, len(e))
    if rc:
        raise HostError(rc, e.value.decode())
