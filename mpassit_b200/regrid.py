"""Python host-side view of the engine: one ``Regridder`` per rank / GPU.

Mirrors the ESMF object vocabulary the reference uses on this path
(/root/reference/interp.F90): a *mesh* (ESMF_Mesh, model_grid.F90:488), a
*target grid* with staggers (ESMF_Grid, model_grid.F90:684-728), ``store`` ->
route handle (ESMF_FieldBundleRegridStore, interp.F90:123), ``apply``
(ESMF_FieldBundleRegrid, interp.F90:134), ``release``
(ESMF_FieldBundleRegridRelease, interp.F90:450) and ``gather``
(ESMF_FieldGather, write_data.F90:1006).  All arithmetic happens in
libmpassit_rg.so; this file only marshals pointers.

Fields may be numpy arrays (host buffers, copied through the engine's pipelined
staging) or torch CUDA tensors (used in place).
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import lib as _l
from .lib import (BILINEAR, CENTER, CENTER_HALO, CONSERVE, CORNER, DEVICE, EDGE1, EDGE2, F32, F64, HOST,  # noqa: F401
                  NEAREST_STOD, SRC_GRID_CENTER, SRC_MESH_ELEMENT, SRC_MESH_NODE, MprgError)


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _dtype_code(x) -> int:
    if _is_torch(x):
        import torch

        if x.dtype == torch.float32:
            return F32
        if x.dtype == torch.float64:
            return F64
    else:
        if x.dtype == np.float32:
            return F32
        if x.dtype == np.float64:
            return F64
    raise TypeError(f"unsupported field dtype {x.dtype}")


def _ptr(x) -> int:
    if isinstance(x, int):       # a raw device address (e.g. the writing rank's buffer mapped with ipc_open)
        return x
    if _is_torch(x):
        if not x.is_contiguous():
            raise ValueError("device fields must be contiguous")
        return x.data_ptr()
    if not x.flags["C_CONTIGUOUS"]:
        raise ValueError("host fields must be C-contiguous")
    return x.ctypes.data


class Route:
    """Opaque route handle (weights + schedule) == type(esmf_routehandle), interp.F90:86."""

    def __init__(self, owner: "Regridder", handle: int, method: int, src_loc: int, dst_stagger: int):
        self.owner, self.handle = owner, handle
        self.method, self.src_loc, self.dst_stagger = method, src_loc, dst_stagger

    def info(self) -> dict:
        v = [C.c_int64() for _ in range(4)]
        rc = _l.load().mprg_route_info(self.handle, *[C.byref(x) for x in v])
        _l.check(self.owner.ctx, rc)
        t = [C.c_int64() for _ in range(3)]
        _l.load().mprg_route_schedule_info(self.handle, *[C.byref(x) for x in t])
        return dict(nDst=v[0].value, nnz=v[1].value, nUnmapped=v[2].value, nSrc=v[3].value,
                    nSrcRef=int(_l.load().mprg_route_src_referenced(self.handle)),
                    tiles=t[0].value, tile_columns=t[1].value, tile_runs=t[2].value)

    def export_csr(self):
        i = self.info()
        rowptr = np.empty(i["nDst"] + 1, np.int32)
        col = np.empty(max(i["nnz"], 1), np.int32)
        w = np.empty(max(i["nnz"], 1), np.float64)
        rc = _l.load().mprg_route_export_csr(self.owner.ctx, self.handle, rowptr.ctypes.data, col.ctypes.data,
                                             w.ctypes.data)
        _l.check(self.owner.ctx, rc)
        return rowptr, col[: i["nnz"]], w[: i["nnz"]]

    def export_w2(self):
        """Second weights (meridional source) of a composed wind route."""
        w2 = np.empty(max(self.info()["nnz"], 1), np.float64)
        _l.check(self.owner.ctx, _l.load().mprg_route_export_w2(self.owner.ctx, self.handle, w2.ctypes.data))
        return w2[: self.info()["nnz"]]

    def release(self) -> None:
        if self.handle:
            _l.check(self.owner.ctx, _l.load().mprg_release(self.owner.ctx, self.handle))
            self.handle = 0


class Regridder:
    def __init__(self, device: int = 0, rank: int = 0, nranks: int = 1):
        L = _l.load()
        ctx = C.c_void_p()
        rc = L.mprg_init(device, rank, nranks, C.byref(ctx))
        if rc != 0:
            raise MprgError(rc, (L.mprg_last_error(None) or b"").decode())
        self.ctx = ctx
        self.L = L
        self.device, self.rank, self.nranks = device, rank, nranks
        self.shape = {}  # stagger -> (nj, ni) full grid
        self.nSrc = {}

    # ---- lifetime -------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "ctx", None):
            self.L.mprg_finalize(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int) -> None:
        _l.check(self.ctx, rc)

    def use_torch_stream(self) -> None:
        """Run engine work on torch's current CUDA stream (so torch events time it)."""
        import torch

        self._torch_stream = torch.cuda.current_stream().cuda_stream
        self._ck(self.L.mprg_set_stream(self.ctx, C.c_void_p(self._torch_stream)))

    def order_after_torch(self, *bufs) -> None:
        """Device tensors handed to the engine were produced on torch's current stream; the engine runs on its own
        (non-blocking) stream unless use_torch_stream() was called.  Make what torch has queued visible first --
        marshalling only: a compiled host orders its own streams (mprg_set_stream)."""
        if not any(_is_torch(b) for b in bufs):
            return
        import torch

        cur = torch.cuda.current_stream()
        if getattr(self, "_torch_stream", None) != cur.cuda_stream:
            cur.synchronize()

    def set_async(self, on: bool) -> None:
        """Host-buffer applies return once queued; call synchronize() before reading their outputs."""
        self._ck(self.L.mprg_set_async(self.ctx, int(on)))

    # ---- CUDA graphs ------------------------------------------------------
    def capture_begin(self) -> None:
        self._ck(self.L.mprg_capture_begin(self.ctx))

    def capture_end(self) -> int:
        g = C.c_void_p()
        self._ck(self.L.mprg_capture_end(self.ctx, C.byref(g)))
        return g.value

    def graph_launch(self, graph: int) -> None:
        self._ck(self.L.mprg_graph_launch(self.ctx, C.c_void_p(graph)))

    def graph_release(self, graph: int) -> None:
        self._ck(self.L.mprg_graph_release(self.ctx, C.c_void_p(graph)))

    def synchronize(self) -> None:
        self._ck(self.L.mprg_synchronize(self.ctx))

    @property
    def kernel_launches(self) -> int:
        return int(self.L.mprg_kernel_launches(self.ctx))

    def set_source_byte_order(self, big_endian: bool) -> None:
        """Host sources of the following applies hold big-endian words (a variable mapped from a NetCDF classic file)."""
        self._ck(self.L.mprg_set_source_byte_order(self.ctx, int(big_endian)))

    def bswap(self, buf) -> None:
        """In-place word swap of a torch CUDA tensor (fp32 / int32 / fp64)."""
        self._ck(self.L.mprg_bswap(self.ctx, _ptr(buf), buf.numel(), F32 if buf.element_size() == 4 else F64))

    def post_affine(self, buf, scale: float, offset: float) -> None:
        self._ck(self.L.mprg_post_affine(self.ctx, _ptr(buf), buf.numel(), _dtype_code(buf), scale, offset))

    def io_bytes(self) -> tuple[int, int]:
        """(host->device, device->host) field bytes moved by host-buffer applies since init."""
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self.L.mprg_io_bytes(self.ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def last_ms(self) -> float:
        return float(self.L.mprg_last_ms(self.ctx))

    # ---- profiling ------------------------------------------------------
    def profile(self, on: bool = True) -> None:
        self._ck(self.L.mprg_profile_enable(self.ctx, int(on)))
        self._ck(self.L.mprg_profile_reset(self.ctx))

    def profile_read(self) -> list[dict]:
        n = self.L.mprg_profile_read(self.ctx, 0, None, None, None, None)
        kind = np.zeros(max(n, 1), np.int32)
        ms, by, un = (np.zeros(max(n, 1), np.float64) for _ in range(3))
        self.L.mprg_profile_read(self.ctx, n, kind.ctypes.data, ms.ctypes.data, by.ctypes.data, un.ctypes.data)
        return [dict(kind=int(kind[i]), ms=float(ms[i]), alg_bytes=float(by[i]), units=float(un[i])) for i in range(n)]

    # ---- geometry -------------------------------------------------------
    def set_mesh(self, lonCell, latCell, lonVertex, latVertex, verticesOnCell) -> None:
        lonC = np.ascontiguousarray(lonCell, np.float64)
        latC = np.ascontiguousarray(latCell, np.float64)
        lonV = np.ascontiguousarray(lonVertex, np.float64)
        latV = np.ascontiguousarray(latVertex, np.float64)
        voc = np.ascontiguousarray(verticesOnCell, np.int32)
        if voc.ndim != 2 or voc.shape[0] != lonC.size:
            raise ValueError("verticesOnCell must be [nCells][maxEdges]")
        self._ck(self.L.mprg_set_mesh(self.ctx, lonC.size, lonV.size, voc.shape[1], lonC.ctypes.data, latC.ctypes.data,
                                      lonV.ctypes.data, latV.ctypes.data, voc.ctypes.data))
        self.nSrc[SRC_MESH_ELEMENT] = lonC.size
        self.nSrc[SRC_MESH_NODE] = lonV.size

    def set_target(self, stagger: int, lon_deg, lat_deg) -> None:
        lon = np.ascontiguousarray(lon_deg, np.float64)
        lat = np.ascontiguousarray(lat_deg, np.float64)
        if lon.ndim != 2 or lon.shape != lat.shape:
            raise ValueError("target coordinates must be [nj][ni]")
        nj, ni = lon.shape
        self._ck(self.L.mprg_set_target(self.ctx, stagger, ni, nj, lon.ctypes.data, lat.ctypes.data))
        self.shape[stagger] = (nj, ni)
        if stagger == CENTER:
            self.nSrc[SRC_GRID_CENTER] = nj * ni

    def set_weight_cache(self, directory: str | None) -> None:
        """Cross-run weight cache directory (mprg_set_weight_cache); None switches it off."""
        self._ck(self.L.mprg_set_weight_cache(self.ctx, directory.encode() if directory else None))

    def weight_cache_stats(self) -> tuple[int, int]:
        """(routes loaded from the cache, routes written to it) since init."""
        h, s_ = C.c_int64(), C.c_int64()
        self.L.mprg_weight_cache_stats(self.ctx, C.byref(h), C.byref(s_))
        return h.value, s_.value

    def set_option(self, key: str, value) -> None:
        """Tuning knob (include/mpassit_rg.h: mprg_set_option), e.g. ("accumulate", "f64"), ("staging", "ldg")."""
        self._ck(self.L.mprg_set_option(self.ctx, key.encode(), str(value).encode()))

    def get_option(self, key: str) -> str:
        import ctypes as C
        buf = C.create_string_buffer(64)
        rc = self.L.mprg_get_option(self.ctx, key.encode(), buf, len(buf))
        if rc:
            raise KeyError(key)
        return buf.value.decode()

    def set_target_projected(self, stagger: int, ni: int, nj: int, proj) -> None:
        """Coordinates of one stagger generated on the device from the projection scalars (mprg_set_target_projected)."""
        self._ck(self.L.mprg_set_target_projected(self.ctx, int(stagger), int(ni), int(nj), C.byref(proj)))
        self.shape[stagger] = (nj, ni)
        if stagger == CENTER:
            self.nSrc[SRC_GRID_CENTER] = nj * ni

    def target_lonlat(self, stagger: int):
        nj, ni = self.shape[stagger]
        lon, lat = np.empty((nj, ni), np.float64), np.empty((nj, ni), np.float64)
        self._ck(self.L.mprg_get_target_lonlat(self.ctx, int(stagger), lon.ctypes.data, lat.ctypes.data))
        return lon, lat

    def target_map_factor(self, stagger: int, proj_code: int, truelat1: float, truelat2: float):
        nj, ni = self.shape[stagger]
        out = np.empty((nj, ni), np.float64)
        self._ck(self.L.mprg_target_map_factor(self.ctx, int(stagger), int(proj_code), float(truelat1), float(truelat2), out.ctypes.data))
        return out

    def set_rotation_from_target(self):
        """get_rotang on the device from the CENTER stagger + mprg_set_rotation; returns (cosa, sina)."""
        nj, ni = self.shape[CENTER]
        ca, sa = np.empty((nj, ni), np.float64), np.empty((nj, ni), np.float64)
        self._ck(self.L.mprg_set_rotation_from_target(self.ctx, ca.ctypes.data, sa.ctypes.data))
        return ca, sa

    def set_grid_kind(self, kind: int) -> None:
        """ESMF_GridCreateNoPeriDim (0) or ESMF_GridCreate1PeriDim + MONOPOLE (1), model_grid.F90:684-703."""
        self._ck(self.L.mprg_set_grid_kind(self.ctx, int(kind)))

    def slab(self, stagger: int) -> tuple[int, int]:
        j0, j1 = C.c_int32(), C.c_int32()
        self._ck(self.L.mprg_get_slab(self.ctx, stagger, C.byref(j0), C.byref(j1)))
        return j0.value, j1.value

    # ---- weights --------------------------------------------------------
    def store(self, method: int, src_loc: int = SRC_MESH_ELEMENT, dst_stagger: int = CENTER) -> Route:
        h = C.c_void_p()
        self._ck(self.L.mprg_store(self.ctx, method, src_loc, dst_stagger, C.byref(h)))
        return Route(self, h.value, method, src_loc, dst_stagger)

    def import_csr(self, nSrc: int, rowptr, col, w) -> Route:
        rowptr = np.ascontiguousarray(rowptr, np.int32)
        col = np.ascontiguousarray(col, np.int32)
        w = np.ascontiguousarray(w, np.float64)
        h = C.c_void_p()
        self._ck(self.L.mprg_route_import_csr(self.ctx, nSrc, rowptr.size - 1, rowptr.ctypes.data,
                                              col.ctypes.data if col.size else None,
                                              w.ctypes.data if w.size else None, C.byref(h)))
        return Route(self, h.value, -1, SRC_MESH_ELEMENT, -1)

    def clear_routes(self) -> None:
        self._ck(self.L.mprg_clear_routes(self.ctx))

    # ---- apply ----------------------------------------------------------
    def apply(self, route: Route, srcs: Sequence, dsts: Sequence, nlev: Sequence[int] | None = None,
              epi_op: Sequence[int] | None = None, epi_arg: Sequence[float] | None = None) -> None:
        """dsts[f][lev][slab] = W . srcs[f]; buffers are numpy (host) or torch CUDA (device)."""
        n = len(srcs)
        if n == 0:
            return
        if len(dsts) != n:
            raise ValueError("srcs/dsts length mismatch")
        self.order_after_torch(srcs[0], dsts[0])
        s_mem = DEVICE if _is_torch(srcs[0]) else HOST
        d_mem = DEVICE if _is_torch(dsts[0]) else HOST
        s_dt, d_dt = _dtype_code(srcs[0]), _dtype_code(dsts[0])
        if nlev is None:
            nlev = []
            for s in srcs:
                if route.src_loc == SRC_GRID_CENTER:
                    nlev.append(1 if s.ndim == 2 else int(s.shape[0]))
                else:
                    nlev.append(1 if s.ndim == 1 else int(s.shape[1]))
        sp = (C.c_void_p * n)(*[_ptr(s) for s in srcs])
        dp = (C.c_void_p * n)(*[_ptr(d) for d in dsts])
        nl = (C.c_int32 * n)(*[int(v) for v in nlev])
        for s, d in zip(srcs, dsts):
            if _dtype_code(s) != s_dt or _dtype_code(d) != d_dt:
                raise TypeError("all stacked fields of one apply share a dtype")
        if epi_op is None:
            rc = self.L.mprg_apply(self.ctx, route.handle, n, sp, nl, s_dt, s_mem, dp, d_dt, d_mem)
        else:
            eo = (C.c_int32 * n)(*[int(v) for v in epi_op])
            ea = (C.c_double * n)(*[float(v) for v in (epi_arg or [0.0] * n)])
            rc = self.L.mprg_apply_ex(self.ctx, route.handle, n, sp, nl, s_dt, s_mem, dp, d_dt, d_mem, eo, ea)
        self._ck(rc)

    def apply_into(self, route: Route, srcs: Sequence, dsts_full: Sequence, nlev: Sequence[int], dst_dtype: int = F32,
                   epi_op: Sequence[int] | None = None, epi_arg: Sequence[float] | None = None) -> None:
        """Like apply, but dsts_full[f] is the FULL [nlev][nj][ni] field (a CUDA tensor of this process or a
        raw address from ipc_open) and this rank writes only its own rows: the gather fused into the store."""
        n = len(srcs)
        if n == 0:
            return
        self.order_after_torch(srcs[0], dsts_full[0])
        s_mem = DEVICE if _is_torch(srcs[0]) else HOST
        sp = (C.c_void_p * n)(*[_ptr(s) for s in srcs])
        dp = (C.c_void_p * n)(*[_ptr(d) for d in dsts_full])
        nl = (C.c_int32 * n)(*[int(v) for v in nlev])
        eo = (C.c_int32 * n)(*[int(v) for v in epi_op]) if epi_op is not None else None
        ea = (C.c_double * n)(*[float(v) for v in (epi_arg or [0.0] * n)]) if epi_op is not None else None
        self._ck(self.L.mprg_apply_into(self.ctx, route.handle, n, sp, nl, _dtype_code(srcs[0]), s_mem, dp, dst_dtype, eo, ea))

    def put_slab(self, stagger: int, nlev: int, slab, full, dtype: int = F32) -> None:
        self.order_after_torch(slab, full)
        self._ck(self.L.mprg_put_slab(self.ctx, int(stagger), int(nlev), dtype, _ptr(slab), _ptr(full)))

    def ipc_export(self, buf) -> tuple[bytes, int]:
        """(64-byte handle, offset) of a device buffer of this process, for the other ranks' ipc_open."""
        h = C.create_string_buffer(64)
        off = C.c_size_t()
        self._ck(self.L.mprg_ipc_export(self.ctx, _ptr(buf), h, C.byref(off)))
        return h.raw, off.value

    def ipc_open(self, handle: bytes, offset: int) -> int:
        p = C.c_void_p()
        self._ck(self.L.mprg_ipc_open(self.ctx, C.create_string_buffer(handle, 64), offset, C.byref(p)))
        return int(p.value)

    def ipc_close_all(self) -> None:
        self._ck(self.L.mprg_ipc_close_all(self.ctx))

    # ---- winds ----------------------------------------------------------
    def set_rotation(self, cosa, sina) -> None:
        ca = np.ascontiguousarray(cosa, np.float64)
        sa = np.ascontiguousarray(sina, np.float64)
        self._ck(self.L.mprg_set_rotation(self.ctx, ca.ctypes.data, sa.ctypes.data))

    def store_wind(self, dst_stagger: int) -> Route | None:
        """Composed wind route (stagger x rotate_winds_cgrid x bilinear, interp.F90:256-328) or None when the grid /
        mesh is not composable and the three-step chain must be used."""
        h = C.c_void_p()
        self._ck(self.L.mprg_store_wind(self.ctx, int(dst_stagger), C.byref(h)))
        return Route(self, h.value, BILINEAR, _l.SRC_MESH_WIND, dst_stagger) if h.value else None

    def apply_wind(self, route: Route, u_src, v_src, dst, nlev: int, into_full: bool = False) -> None:
        """dst = A u_src + B v_src: cell-centre winds (torch CUDA tensors, [nCells][nlev]) -> this rank's slab of the
        rotated, staggered wind [nlev][nj_slab][ni]."""
        self.order_after_torch(u_src, v_src, dst)
        self._ck(self.L.mprg_apply_wind(self.ctx, route.handle, _ptr(u_src), _ptr(v_src), int(nlev), _dtype_code(u_src),
                                        _ptr(dst), _dtype_code(dst), 1 if into_full else 0))

    def rotate_winds(self, u, v, nlev: int, stagger: int = CENTER) -> None:
        self.order_after_torch(u, v)
        mem = DEVICE if _is_torch(u) else HOST
        self._ck(self.L.mprg_rotate_winds_on(self.ctx, int(stagger), _ptr(u), _ptr(v), int(nlev), _dtype_code(u), mem))

    # ---- WRF-compatibility post-ops (write_data.F90:1364-1373, 1406-1412) -----------
    def post_midlevels(self, x, mid, nlev: int, stagger: int = CENTER) -> None:
        self.order_after_torch(x, mid)
        mem = DEVICE if _is_torch(x) else HOST
        self._ck(self.L.mprg_post_midlevels(self.ctx, int(stagger), int(nlev), _dtype_code(x), mem, _ptr(x), _ptr(mid)))

    def post_ptop(self, x, nlev: int, stagger: int = CENTER) -> tuple[float, float]:
        self.order_after_torch(x)
        mem = DEVICE if _is_torch(x) else HOST
        a, b = C.c_double(), C.c_double()
        self._ck(self.L.mprg_post_ptop(self.ctx, int(stagger), int(nlev), _dtype_code(x), mem, _ptr(x), C.byref(a), C.byref(b)))
        return a.value, b.value

    # ---- gather ---------------------------------------------------------
    def comm_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._ck(self.L.mprg_comm_id(self.ctx, buf))
        return buf.raw

    def comm_init(self, id128: bytes) -> None:
        buf = C.create_string_buffer(id128, 128)
        self._ck(self.L.mprg_comm_init(self.ctx, buf))

    def gather(self, stagger: int, nlev: int, slab, root: int = 0, full=None) -> None:
        self._ck(self.L.mprg_gather(self.ctx, stagger, int(nlev), _dtype_code(slab), _ptr(slab), int(root),
                                    _ptr(full) if full is not None else None))

    def gather_many(self, items, root: int = 0) -> None:
        """items: [(stagger, nlev, slab, full-or-None)], one NCCL group for all of them (mprg_gather_v)."""
        n = len(items)
        if n == 0:
            return
        st = (C.c_int * n)(*[int(i[0]) for i in items])
        nl = (C.c_int32 * n)(*[int(i[1]) for i in items])
        sl = (C.c_void_p * n)(*[_ptr(i[2]) for i in items])
        fu = (C.c_void_p * n)(*[(_ptr(i[3]) if i[3] is not None else None) for i in items])
        self._ck(self.L.mprg_gather_v(self.ctx, n, st, nl, _dtype_code(items[0][2]), sl, int(root), fu))
