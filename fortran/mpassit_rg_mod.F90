!> mpassit_rg_mod -- ISO_C_BINDING interfaces to libmpassit_rg.so (include/mpassit_rg.h).
!!
!! This is the thin layer BASELINE.json's north_star asks for: the Fortran host code of
!! MPASSIT (interp.F90, model_grid.F90, input_data.F90, write_data.F90) keeps its structure
!! and calls CUDA through these bind(C) interfaces instead of ESMF.  Each interface names
!! the ESMF call it replaces.  NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Fortran compiler
!! is installed there); the C ABI itself is exercised from C++/Python (tests/, bench.py) and
!! the call sites a maintainer would change are listed in INTEGRATION.md.
!!
!! Conventions: every function returns integer(c_int) rc, 0 = success (== ESMF_SUCCESS), so
!! the reference's pattern is kept verbatim:
!!     rc = mprg_store(ctx, MPRG_BILINEAR, MPRG_SRC_MESH_ELEMENT, MPRG_CENTER, rh_patch)
!!     if (rc /= 0) call error_handler("IN FieldBundleRegridStore", rc)
!! Fortran arrays map without copies: a (nz, nCells) file-order variable is the C
!! [nCells][nz] the engine wants (the reference's transpose at input_data.F90:653-655
!! disappears); a target (i, j, lev) array is the C [lev][j][i] the engine writes.
module mpassit_rg_mod
  use iso_c_binding
  implicit none
  private

  ! enum values of include/mpassit_rg.h
  integer(c_int), parameter, public :: MPRG_BILINEAR = 0, MPRG_CONSERVE = 1, MPRG_NEAREST_STOD = 2
  integer(c_int), parameter, public :: MPRG_SRC_MESH_ELEMENT = 0, MPRG_SRC_MESH_NODE = 1, MPRG_SRC_GRID_CENTER = 2
  integer(c_int), parameter, public :: MPRG_CENTER = 0, MPRG_EDGE1 = 1, MPRG_EDGE2 = 2, MPRG_CORNER = 3, &
                                        MPRG_CENTER_HALO = 4
  integer(c_int), parameter, public :: MPRG_GRID_NOPERI = 0, MPRG_GRID_1PERI_MONOPOLE = 1
  integer(c_int), parameter, public :: MPRG_F32 = 0, MPRG_F64 = 1
  integer(c_int), parameter, public :: MPRG_HOST = 0, MPRG_DEVICE = 1
  integer(c_int), parameter, public :: MPRG_EPI_NONE = 0, MPRG_EPI_ADD = 1, MPRG_EPI_MUL = 2, &
                                        MPRG_EPI_ROT_U = 3, MPRG_EPI_ROT_V = 4

  !> mirrors struct mprg_projection of include/mpassit_rg.h
  type, bind(C), public :: mprg_projection
     integer(c_int32_t) :: code, nxmin, nxmax
     real(c_double) :: lat1, lon1, knowni, knownj, latinc, loninc, stdlon, truelat1, truelat2, hemi, cone, polei, polej, rebydx
  end type mprg_projection

  public :: mprg_set_target_projected, mprg_target_map_factor, mprg_set_rotation_from_target, mprg_host_bind_to_device
  public :: mprg_init, mprg_finalize, mprg_last_error, mprg_error_message
  public :: mprg_set_mesh, mprg_set_target, mprg_set_grid_kind, mprg_set_option, mprg_set_weight_cache, mprg_get_slab
  public :: mprg_store, mprg_release, mprg_clear_routes, mprg_route_info
  public :: mprg_apply, mprg_apply_ex, mprg_set_rotation, mprg_rotate_winds, mprg_rotate_winds_on
  public :: mprg_store_wind, mprg_apply_wind
  public :: mprg_comm_id, mprg_comm_init, mprg_gather, mprg_gather_v
  public :: mprg_set_async, mprg_get_async, mprg_download, mprg_io_bytes
  public :: mprg_post_midlevels, mprg_post_ptop, mprg_route_schedule_info
  public :: mprg_set_source_byte_order, mprg_bswap, mprg_post_affine
  public :: mprg_apply_into, mprg_put_slab, mprg_ipc_export, mprg_ipc_open, mprg_ipc_close_all
  public :: mprg_device_alloc, mprg_device_free
  public :: mprg_capture_begin, mprg_capture_end, mprg_graph_launch, mprg_graph_release
  public :: mprg_host_alloc, mprg_host_free, mprg_scratch, mprg_synchronize

  interface
     !> replaces ESMF_Initialize + ESMF_VMGet (mpassit.F90:84-94)
     integer(c_int) function mprg_init(device, rank, nranks, ctx) bind(C, name="mprg_init")
       import :: c_int, c_ptr
       integer(c_int), value :: device, rank, nranks
       type(c_ptr), intent(out) :: ctx
     end function
     !> replaces ESMF_finalize (mpassit.F90:140)
     integer(c_int) function mprg_finalize(ctx) bind(C, name="mprg_finalize")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
     end function
     type(c_ptr) function mprg_last_error(ctx) bind(C, name="mprg_last_error")
       import :: c_ptr
       type(c_ptr), value :: ctx
     end function
     integer(c_int) function mprg_synchronize(ctx) bind(C, name="mprg_synchronize")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
     end function
     integer(c_int) function mprg_host_alloc(ctx, bytes, ptr) bind(C, name="mprg_host_alloc")
       import :: c_int, c_ptr, c_size_t
       type(c_ptr), value :: ctx
       integer(c_size_t), value :: bytes
       type(c_ptr), intent(out) :: ptr
     end function
     integer(c_int) function mprg_host_free(ctx, ptr) bind(C, name="mprg_host_free")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, ptr
     end function
     integer(c_int) function mprg_scratch(ctx, slot, bytes, ptr) bind(C, name="mprg_scratch")
       import :: c_int, c_ptr, c_size_t
       type(c_ptr), value :: ctx
       integer(c_int), value :: slot
       integer(c_size_t), value :: bytes
       type(c_ptr), intent(out) :: ptr
     end function

     !> replaces ESMF_MeshCreate (model_grid.F90:488-497).  Pass lonCell, latCell, lonVert,
     !! latVert (radians, as read at model_grid.F90:354-384) and vertOnCell(maxEdges,nCells)
     !! (model_grid.F90:416) unchanged; no unique_sort / FINDLOC connectivity build is needed.
     integer(c_int) function mprg_set_mesh(ctx, nCells, nVertices, maxEdges, lonCell, latCell, lonVertex, &
                                           latVertex, verticesOnCell) bind(C, name="mprg_set_mesh")
       import :: c_int, c_ptr, c_int32_t, c_double
       type(c_ptr), value :: ctx
       integer(c_int32_t), value :: nCells, nVertices, maxEdges
       real(c_double), intent(in) :: lonCell(*), latCell(*), lonVertex(*), latVertex(*)
       integer(c_int32_t), intent(in) :: verticesOnCell(*)
     end function
     !> replaces ESMF_GridAddCoord / GridGetCoord fills (model_grid.F90:707-728, 949-1038):
     !! one call per stagger with the (i,j) lon/lat arrays in degrees
     !! (longitude_one / latitude_one etc. from get_lat_lon_fields, model_grid.F90:747-794)
     integer(c_int) function mprg_set_target(ctx, stagger, ni, nj, lon_deg, lat_deg) bind(C, name="mprg_set_target")
       import :: c_int, c_ptr, c_int32_t, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: stagger
       integer(c_int32_t), value :: ni, nj
       real(c_double), intent(in) :: lon_deg(*), lat_deg(*)
     end function
     !> target coordinates / map factors / rotation angles generated on the device (include/mpassit_rg.h);
     !! proj = the scalars that map_set / set_lc leave in proj_info (module_map_utils.F90:243-568, 1083-1121)
     integer(c_int) function mprg_set_target_projected(ctx, stagger, ni, nj, proj) bind(C, name="mprg_set_target_projected")
       import :: c_int, c_ptr, c_int32_t, mprg_projection
       type(c_ptr), value :: ctx
       integer(c_int), value :: stagger
       integer(c_int32_t), value :: ni, nj
       type(mprg_projection), intent(in) :: proj
     end function
     integer(c_int) function mprg_target_map_factor(ctx, stagger, proj_code, truelat1, truelat2, mapfac) &
         bind(C, name="mprg_target_map_factor")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: stagger, proj_code
       real(c_double), value :: truelat1, truelat2
       real(c_double), intent(out) :: mapfac(*)
     end function
     integer(c_int) function mprg_set_rotation_from_target(ctx, cosa, sina) bind(C, name="mprg_set_rotation_from_target")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       real(c_double), intent(out) :: cosa(*), sina(*)
     end function
     !> replaces the choice between ESMF_GridCreate1PeriDim (polekindflag=MONOPOLE, periodicDim=1) and
     !! ESMF_GridCreateNoPeriDim (model_grid.F90:684-703): kind = MPRG_GRID_1PERI_MONOPOLE when .not. is_regional
     integer(c_int) function mprg_set_grid_kind(ctx, kind) bind(C, name="mprg_set_grid_kind")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int), value :: kind
     end function
     !> bind this rank's process to the NUMA node of its GPU before allocating host buffers (include/mpassit_rg.h)
     integer(c_int) function mprg_host_bind_to_device(device, node) bind(C, name="mprg_host_bind_to_device")
       import :: c_int
       integer(c_int), value :: device
       integer(c_int), intent(out) :: node
     end function
     !> cross-run weight cache directory (NUL-terminated); the reference regenerates every matrix in every run
     !! (program_setup.F90:72-75)
     integer(c_int) function mprg_set_weight_cache(ctx, dir) bind(C, name="mprg_set_weight_cache")
       import :: c_int, c_ptr, c_char
       type(c_ptr), value :: ctx
       character(kind=c_char), intent(in) :: dir(*)
     end function
     !> tuning knob (include/mpassit_rg.h); key and value are NUL-terminated C strings,
     !! e.g. mprg_set_option(ctx, "accumulate"//c_null_char, "f64"//c_null_char) for the reference's R8 arithmetic
     integer(c_int) function mprg_set_option(ctx, key, val) bind(C, name="mprg_set_option")
       import :: c_int, c_ptr, c_char
       type(c_ptr), value :: ctx
       character(kind=c_char), intent(in) :: key(*), val(*)
     end function
     !> rows [j0, j1) (0-based) of a stagger owned by this rank == ESMF_GridGet bounds (clb/cub)
     integer(c_int) function mprg_get_slab(ctx, stagger, j0, j1) bind(C, name="mprg_get_slab")
       import :: c_int, c_ptr, c_int32_t
       type(c_ptr), value :: ctx
       integer(c_int), value :: stagger
       integer(c_int32_t), intent(out) :: j0, j1
     end function

     !> replaces ESMF_FieldRegridStore / ESMF_FieldBundleRegridStore
     !! (interp.F90:123,207,226,241,259,277,298,316,334,353,372,394,421,437)
     integer(c_int) function mprg_store(ctx, method, src_loc, dst_stagger, rh) bind(C, name="mprg_store")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int), value :: method, src_loc, dst_stagger
       type(c_ptr), intent(out) :: rh
     end function
     !> replaces ESMF_FieldBundleRegridRelease (interp.F90:450-463)
     integer(c_int) function mprg_release(ctx, rh) bind(C, name="mprg_release")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, rh
     end function
     integer(c_int) function mprg_clear_routes(ctx) bind(C, name="mprg_clear_routes")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
     end function
     integer(c_int) function mprg_route_info(rh, nDst, nnz, nUnmapped, nSrc) bind(C, name="mprg_route_info")
       import :: c_int, c_ptr, c_int64_t
       type(c_ptr), value :: rh
       integer(c_int64_t), intent(out) :: nDst, nnz, nUnmapped, nSrc
     end function

     !> replaces ESMF_FieldRegrid / ESMF_FieldBundleRegrid
     !! (interp.F90:134,219,236,251,268,286,307,325,344,363,382,404,431,443).
     !! src / dst are arrays of c_loc() addresses, one per field of the bundle.
     integer(c_int) function mprg_apply(ctx, rh, nfields, src, nlev, src_dtype, src_mem, dst, dst_dtype, dst_mem) &
          bind(C, name="mprg_apply")
       import :: c_int, c_ptr, c_int32_t
       type(c_ptr), value :: ctx, rh
       integer(c_int32_t), value :: nfields
       type(c_ptr), intent(in) :: src(*), dst(*)
       integer(c_int32_t), intent(in) :: nlev(*)
       integer(c_int), value :: src_dtype, src_mem, dst_dtype, dst_mem
     end function
     !> same with a fused per-field epilogue (T-300, PHB = zgrid*9.81: write_data.F90:1339-1345, 1417)
     integer(c_int) function mprg_apply_ex(ctx, rh, nfields, src, nlev, src_dtype, src_mem, dst, dst_dtype, dst_mem, &
                                           epi_op, epi_arg) bind(C, name="mprg_apply_ex")
       import :: c_int, c_ptr, c_int32_t, c_double
       type(c_ptr), value :: ctx, rh
       integer(c_int32_t), value :: nfields
       type(c_ptr), intent(in) :: src(*), dst(*)
       integer(c_int32_t), intent(in) :: nlev(*), epi_op(*)
       integer(c_int), value :: src_dtype, src_mem, dst_dtype, dst_mem
       real(c_double), intent(in) :: epi_arg(*)
     end function

     !> cosa_target_grid / sina_target_grid (model_grid.F90:1113-1185) registered once
     integer(c_int) function mprg_set_rotation(ctx, cosa, sina) bind(C, name="mprg_set_rotation")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       real(c_double), intent(in) :: cosa(*), sina(*)
     end function
     !> replaces rotate_winds_cgrid (interp.F90:689-749)
     integer(c_int) function mprg_rotate_winds(ctx, u, v, nlev, dtype, mem) bind(C, name="mprg_rotate_winds")
       import :: c_int, c_ptr, c_int32_t
       type(c_ptr), value :: ctx, u, v
       integer(c_int32_t), value :: nlev
       integer(c_int), value :: dtype, mem
     end function
     !> same on the rows of MPRG_CENTER or MPRG_CENTER_HALO (the mass-point winds of the U/V chain)
     integer(c_int) function mprg_rotate_winds_on(ctx, stagger, u, v, nlev, dtype, mem) &
         bind(C, name="mprg_rotate_winds_on")
       import :: c_int, c_ptr, c_int32_t
       type(c_ptr), value :: ctx, u, v
       integer(c_int), value :: stagger
       integer(c_int32_t), value :: nlev
       integer(c_int), value :: dtype, mem
     end function

     !> the wind chain of interp_hist_data (interp.F90:256-328: regrid to mass points, rotate_winds_cgrid, regrid to
     !! EDGE1 / EDGE2) composed into ONE matrix per staggered grid; rh = c_null_ptr with rc 0: not composable
     !! (global / periodic grid, no rotation), keep the three steps
     integer(c_int) function mprg_store_wind(ctx, dst_stagger, rh) bind(C, name="mprg_store_wind")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int), value :: dst_stagger
       type(c_ptr), intent(out) :: rh
     end function
     !> dst = A u_src + B v_src: cell-centre winds (device, file order) -> this rank's slab of rotated U / V
     integer(c_int) function mprg_apply_wind(ctx, rh, u_src, v_src, nlev, src_dtype, dst, dst_dtype, into_full) &
         bind(C, name="mprg_apply_wind")
       import :: c_int, c_ptr, c_int32_t
       type(c_ptr), value :: ctx, rh, u_src, v_src, dst
       integer(c_int32_t), value :: nlev
       integer(c_int), value :: src_dtype, dst_dtype, into_full
     end function

     !> NCCL bootstrap: rank 0 calls mprg_comm_id, the 128 bytes are MPI_Bcast'ed by the host
     !! (mpi_comm_world is already up, mpassit.F90:71), every rank calls mprg_comm_init
     integer(c_int) function mprg_comm_id(ctx, id128) bind(C, name="mprg_comm_id")
       import :: c_int, c_ptr, c_char
       type(c_ptr), value :: ctx
       character(kind=c_char), intent(out) :: id128(128)
     end function
     integer(c_int) function mprg_comm_init(ctx, id128) bind(C, name="mprg_comm_init")
       import :: c_int, c_ptr, c_char
       type(c_ptr), value :: ctx
       character(kind=c_char), intent(in) :: id128(128)
     end function
     !> replaces ESMF_FieldGather(rootPet=0) (write_data.F90:1006-1453); device buffers
     integer(c_int) function mprg_gather(ctx, stagger, nlev, dtype, slab_dev, root, full_dev) bind(C, name="mprg_gather")
       import :: c_int, c_ptr, c_int32_t
       type(c_ptr), value :: ctx, slab_dev, full_dev
       integer(c_int), value :: stagger, dtype, root
       integer(c_int32_t), value :: nlev
     end function
     !> asynchronous host-buffer applies (see mpassit_rg.h); synchronise with mprg_synchronize
     integer(c_int) function mprg_set_async(ctx, on) bind(C, name="mprg_set_async")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int), value :: on
     end function
     integer(c_int) function mprg_get_async(ctx) bind(C, name="mprg_get_async")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
     end function
     integer(c_int) function mprg_download(ctx, dev, host, bytes) bind(C, name="mprg_download")
       import :: c_int, c_ptr, c_size_t
       type(c_ptr), value :: ctx, dev, host
       integer(c_size_t), value :: bytes
     end function
     !> CUDA graph of a device-buffer pass: record once, replay per output time
     integer(c_int) function mprg_capture_begin(ctx) bind(C, name="mprg_capture_begin")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
     end function
     integer(c_int) function mprg_capture_end(ctx, graph) bind(C, name="mprg_capture_end")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       type(c_ptr), intent(out) :: graph
     end function
     integer(c_int) function mprg_graph_launch(ctx, graph) bind(C, name="mprg_graph_launch")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, graph
     end function
     integer(c_int) function mprg_graph_release(ctx, graph) bind(C, name="mprg_graph_release")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, graph
     end function
     !> device memory for buffers that live on the GPU (the writing rank's full fields of the fused gather)
     integer(c_int) function mprg_device_alloc(ctx, bytes, ptr) bind(C, name="mprg_device_alloc")
       import :: c_int, c_ptr, c_size_t
       type(c_ptr), value :: ctx
       integer(c_size_t), value :: bytes
       type(c_ptr), intent(out) :: ptr
     end function
     integer(c_int) function mprg_device_free(ctx, ptr) bind(C, name="mprg_device_free")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, ptr
     end function
     !> apply fused with the gather: dst_full(f) is the whole field of the destination stagger (the
     !! writing rank's buffer, own or mapped with mprg_ipc_open); each rank stores its rows into it
     integer(c_int) function mprg_apply_into(ctx, rh, nfields, src, nlev, src_dtype, src_mem, dst_full, dst_dtype, &
                                             epi_op, epi_arg) bind(C, name="mprg_apply_into")
       import :: c_int, c_ptr, c_int32_t
       type(c_ptr), value :: ctx, rh
       integer(c_int32_t), value :: nfields
       type(c_ptr), intent(in) :: src(*), dst_full(*)
       integer(c_int32_t), intent(in) :: nlev(*)
       integer(c_int), value :: src_dtype, src_mem, dst_dtype
       type(c_ptr), value :: epi_op, epi_arg   ! c_null_ptr: no epilogues
     end function
     integer(c_int) function mprg_put_slab(ctx, stagger, nlev, dtype, slab_dev, full_dev) bind(C, name="mprg_put_slab")
       import :: c_int, c_ptr, c_int32_t
       type(c_ptr), value :: ctx, slab_dev, full_dev
       integer(c_int), value :: stagger, dtype
       integer(c_int32_t), value :: nlev
     end function
     !> CUDA IPC: the writing rank exports (handle, offset) of each full field (MPI_Bcast them), the others map them
     integer(c_int) function mprg_ipc_export(ctx, dev_ptr, handle64, offset) bind(C, name="mprg_ipc_export")
       import :: c_int, c_ptr, c_char, c_size_t
       type(c_ptr), value :: ctx, dev_ptr
       character(kind=c_char), intent(out) :: handle64(64)
       integer(c_size_t), intent(out) :: offset
     end function
     integer(c_int) function mprg_ipc_open(ctx, handle64, offset, peer_ptr) bind(C, name="mprg_ipc_open")
       import :: c_int, c_ptr, c_char, c_size_t
       type(c_ptr), value :: ctx
       character(kind=c_char), intent(in) :: handle64(64)
       integer(c_size_t), value :: offset
       type(c_ptr), intent(out) :: peer_ptr
     end function
     integer(c_int) function mprg_ipc_close_all(ctx) bind(C, name="mprg_ipc_close_all")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
     end function
     integer(c_int) function mprg_route_schedule_info(rh, tiles, columns, runs) bind(C, name="mprg_route_schedule_info")
       import :: c_int, c_ptr, c_int64_t
       type(c_ptr), value :: rh
       integer(c_int64_t), intent(out) :: tiles, columns, runs
     end function
     !> Z_C = 0.5 (PHB(k) + PHB(k-1)) (write_data.F90:1406-1412) on this rank's slab
     integer(c_int) function mprg_post_midlevels(ctx, stagger, nlev, dtype, mem, x, mid) bind(C, name="mprg_post_midlevels")
       import :: c_int, c_ptr, c_int32_t
       type(c_ptr), value :: ctx, x, mid
       integer(c_int), value :: stagger, dtype, mem
       integer(c_int32_t), value :: nlev
     end function
     !> this rank's share of P_TOP (write_data.F90:1364-1373): combine with MPI_MAX / MPI_MIN
     integer(c_int) function mprg_post_ptop(ctx, stagger, nlev, dtype, mem, x, maxval, mincand) bind(C, name="mprg_post_ptop")
       import :: c_int, c_ptr, c_int32_t, c_double
       type(c_ptr), value :: ctx, x
       integer(c_int), value :: stagger, dtype, mem
       integer(c_int32_t), value :: nlev
       real(c_double), intent(out) :: maxval, mincand
     end function
     !> sources handed over straight from a classic-NetCDF file mapping are big-endian: swapped on the device
     integer(c_int) function mprg_set_source_byte_order(ctx, big_endian) bind(C, name="mprg_set_source_byte_order")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int), value :: big_endian
     end function
     !> in-place word swap of a device buffer (the writer's last step before mprg_download)
     integer(c_int) function mprg_bswap(ctx, dev, count, dtype) bind(C, name="mprg_bswap")
       import :: c_int, c_ptr, c_size_t
       type(c_ptr), value :: ctx, dev
       integer(c_size_t), value :: count
       integer(c_int), value :: dtype
     end function
     !> x = x*scale + offset on the device: T-300, PHB = 9.81 zgrid, zero fields (write_data.F90:1339-1432)
     integer(c_int) function mprg_post_affine(ctx, dev, count, dtype, scale, offset) bind(C, name="mprg_post_affine")
       import :: c_int, c_ptr, c_size_t, c_double
       type(c_ptr), value :: ctx, dev
       integer(c_size_t), value :: count
       integer(c_int), value :: dtype
       real(c_double), value :: scale, offset
     end function
     integer(c_int) function mprg_io_bytes(ctx, h2d, d2h) bind(C, name="mprg_io_bytes")
       import :: c_int, c_ptr, c_int64_t
       type(c_ptr), value :: ctx
       integer(c_int64_t), intent(out) :: h2d, d2h
     end function
     !> nfields fields in one NCCL group (the ~20 back-to-back FieldGather calls of write_to_file)
     integer(c_int) function mprg_gather_v(ctx, nfields, stagger, nlev, dtype, slab_dev, root, full_dev) &
         bind(C, name="mprg_gather_v")
       import :: c_int, c_ptr, c_int32_t
       type(c_ptr), value :: ctx
       integer(c_int32_t), value :: nfields
       integer(c_int), intent(in) :: stagger(*)
       integer(c_int32_t), intent(in) :: nlev(*)
       integer(c_int), value :: dtype, root
       type(c_ptr), intent(in) :: slab_dev(*), full_dev(*)
     end function
  end interface

contains

  !> mprg_last_error as a Fortran string, for error_handler(trim(msg), rc) (utils.F90:16-33)
  function mprg_error_message(ctx) result(msg)
    type(c_ptr), intent(in) :: ctx
    character(len=:), allocatable :: msg
    character(kind=c_char), pointer :: p(:)
    type(c_ptr) :: cp
    integer :: n, k
    cp = mprg_last_error(ctx)
    msg = ""
    if (.not. c_associated(cp)) return
    call c_f_pointer(cp, p, [1024])
    n = 0
    do while (n < 1024)
       if (p(n + 1) == c_null_char) exit
       n = n + 1
    end do
    ! msg is already allocated (zero length) by the assignment above: re-assign, never ALLOCATE it again
    msg = repeat(" ", n)
    do k = 1, n
       msg(k:k) = p(k)
    end do
  end function mprg_error_message

end module mpassit_rg_mod
