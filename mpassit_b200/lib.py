"""ctypes binding of libmpassit_rg.so (include/mpassit_rg.h).

Thin by design: every function maps 1:1 onto a C-ABI entry point, checks the rc
and raises ``MprgError`` with ``mprg_last_error`` -- the Python analogue of the
reference's ``if (rc /= 0) call error_handler(msg, rc)`` (utils.F90:16-33).
There is no CPU fallback: if the CUDA library is missing or no GPU is visible,
calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmpassit_rg.so")

BILINEAR, CONSERVE, NEAREST_STOD = 0, 1, 2
SRC_MESH_ELEMENT, SRC_MESH_NODE, SRC_GRID_CENTER, SRC_MESH_WIND = 0, 1, 2, 3
CENTER, EDGE1, EDGE2, CORNER, CENTER_HALO = 0, 1, 2, 3, 4
F32, F64 = 0, 1
HOST, DEVICE = 0, 1
GRID_NOPERI, GRID_1PERI_MONOPOLE = 0, 1
EPI_NONE, EPI_ADD, EPI_MUL, EPI_ROT_U, EPI_ROT_V = 0, 1, 2, 3, 4

EXPORTS = [
    "mprg_init", "mprg_finalize", "mprg_last_error", "mprg_version", "mprg_set_stream", "mprg_synchronize", "mprg_set_async", "mprg_get_async", "mprg_set_option", "mprg_get_option", "mprg_download",
    "mprg_host_bind_to_device", "mprg_host_alloc", "mprg_host_free", "mprg_device_alloc", "mprg_device_free", "mprg_scratch", "mprg_has_rotation", "mprg_set_mesh", "mprg_set_target", "mprg_set_target_projected", "mprg_get_target_lonlat", "mprg_target_map_factor", "mprg_set_rotation_from_target", "mprg_set_grid_kind", "mprg_get_slab", "mprg_store",
    "mprg_release", "mprg_clear_routes", "mprg_set_weight_cache", "mprg_weight_cache_stats", "mprg_route_info", "mprg_route_export_csr", "mprg_route_import_csr",
    "mprg_apply", "mprg_apply_ex", "mprg_apply_into", "mprg_put_slab", "mprg_ipc_export", "mprg_ipc_open", "mprg_ipc_close_all", "mprg_set_rotation", "mprg_rotate_winds", "mprg_rotate_winds_on", "mprg_comm_id", "mprg_comm_init",
    "mprg_post_midlevels", "mprg_post_ptop", "mprg_gather", "mprg_gather_v", "mprg_kernel_launches", "mprg_io_bytes", "mprg_capture_begin", "mprg_capture_end", "mprg_graph_launch", "mprg_graph_release", "mprg_last_ms", "mprg_profile_enable", "mprg_profile_read",
    "mprg_profile_reset", "mprg_route_src_referenced", "mprg_route_schedule_info", "mprg_set_source_byte_order", "mprg_bswap", "mprg_post_affine",
    "mprg_store_wind", "mprg_apply_wind", "mprg_route_export_w2",
]


class Projection(C.Structure):
    """mprg_projection (include/mpassit_rg.h): per-grid scalars of the target projection."""
    _fields_ = [("code", C.c_int32), ("nxmin", C.c_int32), ("nxmax", C.c_int32)] + \
               [(n, C.c_double) for n in ("lat1", "lon1", "knowni", "knownj", "latinc", "loninc", "stdlon", "truelat1",
                                          "truelat2", "hemi", "cone", "polei", "polej", "rebydx")]


class MprgError(RuntimeError):
    def __init__(self, rc: int, msg: str):
        super().__init__(f"mprg rc={rc}: {msg}")
        self.rc = rc


_lib = None


def load() -> C.CDLL:
    """Load the engine.  Raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MprgError(-1, f"{LIB_PATH} not built; run `python -m mpassit_b200.build` (or __graft_entry__.build())")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    pp = C.POINTER(C.c_void_p)
    L.mprg_version.restype = C.c_char_p
    L.mprg_last_error.restype = C.c_char_p
    L.mprg_last_error.argtypes = [vp]
    L.mprg_init.argtypes = [C.c_int, C.c_int, C.c_int, pp]
    L.mprg_finalize.argtypes = [vp]
    L.mprg_set_stream.argtypes = [vp, vp]
    L.mprg_synchronize.argtypes = [vp]
    L.mprg_host_bind_to_device.argtypes = [C.c_int, C.POINTER(C.c_int)]
    L.mprg_host_alloc.argtypes = [vp, C.c_size_t, pp]
    L.mprg_host_free.argtypes = [vp, vp]
    L.mprg_device_alloc.argtypes = [vp, C.c_size_t, pp]
    L.mprg_device_free.argtypes = [vp, vp]
    L.mprg_scratch.argtypes = [vp, C.c_int, C.c_size_t, pp]
    L.mprg_has_rotation.argtypes = [vp]
    L.mprg_set_mesh.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp]
    L.mprg_set_target.argtypes = [vp, C.c_int, i32, i32, vp, vp]
    L.mprg_set_grid_kind.argtypes = [vp, C.c_int]
    L.mprg_set_target_projected.argtypes = [vp, C.c_int, i32, i32, C.POINTER(Projection)]
    L.mprg_get_target_lonlat.argtypes = [vp, C.c_int, vp, vp]
    L.mprg_target_map_factor.argtypes = [vp, C.c_int, C.c_int, dbl, dbl, vp]
    L.mprg_set_rotation_from_target.argtypes = [vp, vp, vp]
    L.mprg_set_option.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.mprg_set_weight_cache.argtypes = [vp, C.c_char_p]
    L.mprg_weight_cache_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.mprg_get_option.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_size_t]
    L.mprg_get_slab.argtypes = [vp, C.c_int, C.POINTER(i32), C.POINTER(i32)]
    L.mprg_store.argtypes = [vp, C.c_int, C.c_int, C.c_int, pp]
    L.mprg_release.argtypes = [vp, vp]
    L.mprg_store_wind.argtypes = [vp, C.c_int, pp]
    L.mprg_apply_wind.argtypes = [vp, vp, vp, vp, i32, C.c_int, vp, C.c_int, C.c_int]
    L.mprg_clear_routes.argtypes = [vp]
    L.mprg_route_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    L.mprg_route_export_csr.argtypes = [vp, vp, vp, vp, vp]
    L.mprg_route_export_w2.argtypes = [vp, vp, vp]
    L.mprg_route_import_csr.argtypes = [vp, i64, i64, vp, vp, vp, pp]
    L.mprg_apply.argtypes = [vp, vp, i32, pp, C.POINTER(i32), C.c_int, C.c_int, pp, C.c_int, C.c_int]
    L.mprg_apply_ex.argtypes = [vp, vp, i32, pp, C.POINTER(i32), C.c_int, C.c_int, pp, C.c_int, C.c_int,
                                C.POINTER(i32), C.POINTER(dbl)]
    L.mprg_set_rotation.argtypes = [vp, vp, vp]
    L.mprg_rotate_winds.argtypes = [vp, vp, vp, i32, C.c_int, C.c_int]
    L.mprg_rotate_winds_on.argtypes = [vp, C.c_int, vp, vp, i32, C.c_int, C.c_int]
    L.mprg_comm_id.argtypes = [vp, vp]
    L.mprg_comm_init.argtypes = [vp, vp]
    L.mprg_post_midlevels.argtypes = [vp, C.c_int, i32, C.c_int, C.c_int, vp, vp]
    L.mprg_post_ptop.argtypes = [vp, C.c_int, i32, C.c_int, C.c_int, vp, vp, vp]
    L.mprg_apply_into.argtypes = [vp, vp, i32, pp, C.POINTER(i32), C.c_int, C.c_int, pp, C.c_int, C.POINTER(i32), C.POINTER(dbl)]
    L.mprg_put_slab.argtypes = [vp, C.c_int, i32, C.c_int, vp, vp]
    L.mprg_ipc_export.argtypes = [vp, vp, vp, vp]
    L.mprg_ipc_open.argtypes = [vp, vp, C.c_size_t, vp]
    L.mprg_ipc_close_all.argtypes = [vp]
    L.mprg_route_schedule_info.argtypes = [vp, vp, vp, vp]
    L.mprg_io_bytes.argtypes = [vp, vp, vp]
    L.mprg_capture_begin.argtypes = [vp]
    L.mprg_capture_end.argtypes = [vp, pp]
    L.mprg_graph_launch.argtypes = [vp, vp]
    L.mprg_graph_release.argtypes = [vp, vp]
    L.mprg_set_source_byte_order.argtypes = [vp, C.c_int]
    L.mprg_bswap.argtypes = [vp, vp, C.c_size_t, C.c_int]
    L.mprg_post_affine.argtypes = [vp, vp, C.c_size_t, C.c_int, dbl, dbl]
    L.mprg_set_async.argtypes = [vp, C.c_int]
    L.mprg_get_async.argtypes = [vp]
    L.mprg_download.argtypes = [vp, vp, vp, C.c_size_t]
    L.mprg_gather.argtypes = [vp, C.c_int, i32, C.c_int, vp, C.c_int, vp]
    L.mprg_gather_v.argtypes = [vp, i32, vp, vp, C.c_int, vp, C.c_int, vp]
    L.mprg_kernel_launches.argtypes = [vp]
    L.mprg_kernel_launches.restype = i64
    L.mprg_last_ms.argtypes = [vp]
    L.mprg_last_ms.restype = dbl
    L.mprg_profile_enable.argtypes = [vp, C.c_int]
    L.mprg_profile_reset.argtypes = [vp]
    L.mprg_profile_read.argtypes = [vp, i32, vp, vp, vp, vp]
    L.mprg_route_src_referenced.argtypes = [vp]
    L.mprg_route_src_referenced.restype = i64
    _lib = L
    return L


def check(ctx, rc: int) -> None:
    if rc != 0:
        msg = load().mprg_last_error(ctx)
        raise MprgError(rc, msg.decode() if msg else "")


def host_bind_to_device(device: int) -> int:
    """Pin this process to the NUMA node of `device` (CPUs + preferred memory); returns the node or -1."""
    node = C.c_int(-1)
    load().mprg_host_bind_to_device(int(device), C.byref(node))
    return node.value


def np_ptr(a: np.ndarray) -> int:
    return a.ctypes.data
