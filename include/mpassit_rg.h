/*
 * mpassit_rg.h -- C ABI of libmpassit_rg.so, the B200 (sm_100a) regridding engine
 * that stands in for the ESMF calls on MPASSIT's interpolation path.
 *
 * Plain C: pointers, sizes and opaque handles only.  Every entry point returns
 * an int rc, 0 = success (mirrors ESMF_SUCCESS); nonzero = failure, message via
 * mprg_last_error().  The library never aborts or throws; the Fortran host keeps
 * its `if (rc /= 0) call error_handler(msg, rc)` pattern
 * (/root/reference/utils.F90:16-33, e.g. interp.F90:130-131).
 *
 * All calls on one context are collective in the reference's sense: every rank
 * (one process per GPU) makes the same calls in the same order
 * (interp.F90 has no rank-conditional regrid call).  A context is not
 * thread-safe; work is stream-ordered on the context's stream.
 *
 * Index conventions: MPAS ids are 1-based in verticesOnCell with 0 = unused
 * slot (model_grid.F90:448,479); everything this library exports (CSR) is
 * 0-based.  Target arrays are C order [nj][ni] == Fortran (i,j), i fastest.
 *
 * Each declaration names the reference interface it replaces (file:line under
 * /root/reference).
 */
#ifndef MPASSIT_RG_H
#define MPASSIT_RG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mprg_ctx mprg_ctx;     /* one per rank / GPU */
typedef struct mprg_route mprg_route; /* weights + schedule == type(esmf_routehandle), interp.F90:86,192 */

/* regridmethod, interp.F90:119,204 (BILINEAR) :370 (CONSERVE) :420 (NEAREST_STOD) */
enum { MPRG_BILINEAR = 0, MPRG_CONSERVE = 1, MPRG_NEAREST_STOD = 2 };
/* where the source field lives: ESMF_MESHLOC_ELEMENT (input_data.F90:970-1132),
 * ESMF_MESHLOC_NODE (vorticity bundle, interp.F90:353), or the target grid's
 * CENTER stagger (u/v_target_grid_nostag, interp.F90:298,316) */
enum { MPRG_SRC_MESH_ELEMENT = 0, MPRG_SRC_MESH_NODE = 1, MPRG_SRC_GRID_CENTER = 2,
       /* the (u, v) pair on mesh cells: source of a composed wind route (mprg_store_wind) */
       MPRG_SRC_MESH_WIND = 3 };
/* ESMF_STAGGERLOC_* of the destination, interp.F90:477-520, model_grid.F90:707-728 */
enum { MPRG_CENTER = 0, MPRG_EDGE1 = 1, MPRG_EDGE2 = 2, MPRG_CORNER = 3,
       /* CENTER rows of this rank widened by one halo row on each side (clipped to the grid): the
        * mass-point rows that this rank's EDGE1 / EDGE2 points interpolate from.  The reference gets
        * them through ESMF's halo exchange inside the Grid->Grid regrid (interp.F90:298,316); here the
        * wind chain recomputes them from the replicated source mesh instead, so no rank waits on a
        * neighbour.  Defined as soon as CENTER is set; equals CENTER on one rank. */
       MPRG_CENTER_HALO = 4 };
enum { MPRG_F32 = 0, MPRG_F64 = 1 };
enum { MPRG_HOST = 0, MPRG_DEVICE = 1 };
/* per-field fused epilogues (WRF-compat post-ops the reference runs serially on
 * PET 0, write_data.F90:1339-1345 and :1406-1425) */
enum { MPRG_EPI_NONE = 0, MPRG_EPI_ADD = 1, MPRG_EPI_MUL = 2,
       /* fields f (ROT_U: zonal) and f+1 (ROT_V: meridional), same level count, on CENTER or CENTER_HALO
        * rows: rotate_winds_cgrid (interp.F90:689-749, v' from the already rotated u') is applied to the
        * pair as it is stored, with the angles given to mprg_set_rotation -- the separate rotation pass
        * over both fields disappears.  Bit-identical to mprg_apply + mprg_rotate_winds_on. */
       MPRG_EPI_ROT_U = 3, MPRG_EPI_ROT_V = 4 };

/* ---- lifetime: replaces ESMF_Initialize / ESMF_VMGet / ESMF_finalize
 *      (mpassit.F90:84-94,140).  `device` is the CUDA ordinal; rank/nranks give
 *      the target row-slab this context owns (para_range, model_grid.F90:2428). */
int mprg_init(int device, int rank, int nranks, mprg_ctx **out);
int mprg_finalize(mprg_ctx *ctx);
const char *mprg_last_error(const mprg_ctx *ctx); /* ctx may be NULL: last init error */
const char *mprg_version(void);
/* run all work of this context on an existing CUDA stream (cudaStream_t passed
 * as void*); NULL restores the context's own stream */
int mprg_set_stream(mprg_ctx *ctx, void *cuda_stream);
int mprg_synchronize(mprg_ctx *ctx);
/* Host-buffer applies are blocking by default, like ESMF_FieldRegrid.  With async on they return once
 * their copies and kernels are queued (the H2D / kernel / D2H pipeline then runs across consecutive
 * applies, and beside weight generation, which has its own stream); the caller must not touch the host
 * buffers of queued applies until mprg_synchronize returns.  Environment MPASSIT_GPU_ASYNC=1 sets the
 * initial mode. */
int mprg_set_async(mprg_ctx *ctx, int on);
/* Tuning knobs (never needed for correctness).  Each has an environment variable that is read ONCE, by mprg_init;
 * mprg_set_option changes it afterwards.  key / values [environment variable]:
 *   "accumulate"  "f32" (default) | "f64": arithmetic of fp32-in / fp32-out applies and of the wind rotation; f64 is
 *                 the reference's R8 arithmetic (one rounding on store)                          [MPASSIT_GPU_ACC]
 *   "pipe_split"  "0" (default) | "1": the column kernel runs an apply's plain aligned fields in a lean first phase
 *                 and wind pairs / unaligned level counts in a second phase of the SAME launch; with 1 the two
 *                 groups get a launch each                                                  [MPASSIT_GPU_PIPE_SPLIT]
 *   "apply"       "pipe" (default) | "direct": register-gather kernels only                    [MPASSIT_GPU_APPLY]
 *   "pipe_minb"   0 (default: by shared memory) | 4 | 5 resident CTAs per SM               [MPASSIT_GPU_PIPE_MINB]
 *   "cols_minb"   2 | 3 (default) | 4: register cap of the register-gather kernel               [MPASSIT_GPU_MINB]
 *   "wind"        "chain" (default) | "composed": with chain, mprg_store_wind always declines and hosts run the
 *                 reference's three regrid / rotate steps one after the other; composed lets it build the
 *                 one-matrix route (measured slower than the chain on a mesh as fine as the grid)   [MPASSIT_GPU_WIND]
 *   "upload_threads"  host threads of the unpinned-source bounce ring (0 = 3/4 of the cores) [MPASSIT_UPLOAD_THREADS] */
int mprg_set_option(mprg_ctx *ctx, const char *key, const char *value);
int mprg_get_option(const mprg_ctx *ctx, const char *key, char *value, size_t len);
int mprg_get_async(const mprg_ctx *ctx);
/* device -> host copy ordered after everything queued on the context so far (blocking unless async) */
int mprg_download(mprg_ctx *ctx, const void *dev, void *host, size_t bytes);

/* Host placement.  Every rank moves its share of the fields over ITS GPU's PCIe link; when the page-locked buffers
 * (or the threads that fill them) live on the other socket, that traffic crosses the inter-socket link and the
 * ranks of a box end up sharing one bottleneck (round 1: 8 ranks moved 16.8 GB at 142 GB/s in aggregate).
 * mprg_host_bind_to_device restricts the calling process to the CPUs of the NUMA node the device hangs off and makes
 * that node the preferred one for its future allocations (sched_setaffinity + set_mempolicy; what `numactl
 * --cpunodebind --preferred` does).  Call it once per rank BEFORE allocating host buffers.  *node receives the node
 * (-1: unknown / single-node host, nothing changed).  Needs no context. */
int mprg_host_bind_to_device(int device, int *node);
/* pinned host memory for callers that want full-speed H2D/D2H (optional) */
int mprg_host_alloc(mprg_ctx *ctx, size_t bytes, void **ptr);
int mprg_host_free(mprg_ctx *ctx, void *ptr);
/* device memory for hosts that keep intermediates on the GPU (the mass-point
 * winds between interp.F90:268 and :307 never need to visit the host) */
int mprg_device_alloc(mprg_ctx *ctx, size_t bytes, void **ptr);
int mprg_device_free(mprg_ctx *ctx, void *ptr);
/* engine-owned device scratch that persists across calls (slot 0..7, grown on demand,
 * freed by mprg_finalize): lets a host run interp_data repeatedly without allocating */
int mprg_scratch(mprg_ctx *ctx, int slot, size_t bytes, void **ptr);

/* ---- source mesh: replaces ESMF_MeshCreate, model_grid.F90:488-497.
 *      Takes the arrays exactly as read from the MPAS grid file
 *      (model_grid.F90:354-417): radians, verticesOnCell [nCells][maxEdges]
 *      (== Fortran (maxEdges,nCells)).  The engine applies the reference's
 *      *180/PI and >180 -> -360 wrap (model_grid.F90:450-454,464-468) itself. */
int mprg_set_mesh(mprg_ctx *ctx, int32_t nCells, int32_t nVertices, int32_t maxEdges,
                  const double *lonCell_rad, const double *latCell_rad,
                  const double *lonVertex_rad, const double *latVertex_rad,
                  const int32_t *verticesOnCell);

/* ---- target grid: replaces ESMF_GridCreate* + GridAddCoord/GetCoord fills,
 *      model_grid.F90:684-728, 949-1038.  One call per stagger; lon/lat in
 *      degrees, full grid [nj][ni] on every rank.  CENTER is ni x nj =
 *      i_target x j_target, EDGE1 (ni+1) x nj, EDGE2 ni x (nj+1),
 *      CORNER (ni+1) x (nj+1)  (interp.F90:477-520). */
int mprg_set_target(mprg_ctx *ctx, int stagger, int32_t ni, int32_t nj,
                    const double *lon_deg, const double *lat_deg);
/* ---- target grid generated on the device.  The reference computes the coordinates of all four
 *      staggers, the map factors and the rotation angles on the host, point by point, through the WPS map utilities
 *      (get_lat_lon_fields model_grid.F90:2188-2219 -> ijll_lc / ijll_latlon module_map_utils.F90:1160-1233,1398-1428;
 *      get_map_factor :2229-2365; get_rotang :2450-2507).  mprg_projection carries the per-grid scalars of the projection
 *      (what map_set / set_lc leave in proj_info, module_map_utils.F90:243-568,1083-1121), computed once by the host;
 *      everything O(nx x ny) then runs on the device in fp64 with the reference's formulas:
 *        mprg_set_target_projected      one stagger's lon / lat (xytoll offsets: EDGE1 x-0.5, EDGE2 y-0.5, CORNER both)
 *                                       and its Cartesian coordinates; equivalent to mprg_set_target with those arrays
 *        mprg_get_target_lonlat         the generated (or given) coordinates, [nj][ni] degrees, for the output file
 *        mprg_target_map_factor         MAPFAC_M / _U / _V of a stagger (0 for lat-lon, where the reference leaves it unset)
 *        mprg_set_rotation_from_target  get_rotang on the CENTER stagger + mprg_set_rotation; cosa / sina returned if not NULL
 *      Device trig differs from the host libm in the last 1-2 ulp, so coordinates are not bit-identical to a CPU run
 *      (mprg_set_target with host arrays is); tests/test_gpu_targetgen.py bounds the difference and checks that the
 *      weight matrices' structure is unchanged on the BASELINE configurations. */
typedef struct mprg_projection {
    int32_t code;               /* 0 = cylindrical lat-lon (PROJ_LATLON), 1 = Lambert conformal (PROJ_LC) */
    int32_t nxmin, nxmax;       /* lat-lon: periodic index range (ijll_latlon) */
    double lat1, lon1;          /* lat-lon: coordinates of the known point */
    double knowni, knownj;      /* its (i, j) */
    double latinc, loninc;      /* lat-lon increments, degrees */
    double stdlon, truelat1, truelat2, hemi, cone, polei, polej, rebydx;   /* Lambert: proj_info after set_lc */
} mprg_projection;
int mprg_set_target_projected(mprg_ctx *ctx, int stagger, int32_t ni, int32_t nj, const mprg_projection *proj);
int mprg_get_target_lonlat(mprg_ctx *ctx, int stagger, double *lon_deg, double *lat_deg);
int mprg_target_map_factor(mprg_ctx *ctx, int stagger, int proj_code, double truelat1, double truelat2, double *mapfac);
int mprg_set_rotation_from_target(mprg_ctx *ctx, double *cosa, double *sina);

/* Topology of the target grid: which ESMF_GridCreate* call the reference makes, model_grid.F90:684-703.
 *   MPRG_GRID_NOPERI         ESMF_GridCreateNoPeriDim (is_regional = .true., the default here);
 *   MPRG_GRID_1PERI_MONOPOLE ESMF_GridCreate1PeriDim(periodicDim=1, poleDim=2, polekindflag=MONOPOLE)
 *                            (is_regional = .false.): the grid wraps in i and is closed at both poles.
 * It only matters where the grid itself is a regrid SOURCE -- the centre -> EDGE1 / EDGE2 staggering of the
 * winds (interp.F90:298,316): with a periodic grid the quad between the last and the first centre column
 * exists (U on the seam columns is interpolated across the seam), and the polar caps (ESMF's default
 * polemethod ALLAVG: an artificial pole node whose value is the average of the end row) map V on the pole
 * rows.  Set it before the first mprg_store; changing it drops the memoised grid-source routes. */
enum { MPRG_GRID_NOPERI = 0, MPRG_GRID_1PERI_MONOPOLE = 1 };
int mprg_set_grid_kind(mprg_ctx *ctx, int kind);
/* rows [j0, j1) of `stagger` owned by this rank (0-based; para_range) */
int mprg_get_slab(const mprg_ctx *ctx, int stagger, int32_t *j0, int32_t *j1);

/* ---- weights: replaces ESMF_FieldRegridStore / ESMF_FieldBundleRegridStore
 *      (interp.F90:123,207,226,241,259,277,298,316,334,353,372,394,421,437)
 *      with srcTermProcessing=1, unmappedaction=IGNORE, other arguments default.
 *      Routes are memoised per (method, src_loc, dst_stagger): the reference's 12
 *      Store calls hit 4-6 distinct matrices.  The returned handle is reference
 *      counted; pair every store with one mprg_release. */
int mprg_store(mprg_ctx *ctx, int method, int src_loc, int dst_stagger, mprg_route **rh);
/* replaces ESMF_FieldBundleRegridRelease, interp.F90:450-463 */
int mprg_release(mprg_ctx *ctx, mprg_route *rh);
/* drop every memoised route (forces the next store to rebuild) */
int mprg_clear_routes(mprg_ctx *ctx);
/* Cross-run weight cache: the reference regenerates every matrix in every run (weights are never
 * kept, program_setup.F90:72-75).  With a cache directory set (or MPASSIT_WEIGHT_CACHE in the environment at
 * mprg_init), mprg_store looks the route up under a 128-bit key of everything that determines it -- source mesh,
 * the destination points of this rank's slab, grid topology, method, source location, stagger, slab bounds, engine
 * version; computed on the device from the arrays weight generation reads -- loads its CSR from
 * <dir>/mprg_<key>.w when present and writes that file after generating it otherwise.  NULL or "" switches it off.
 * Results with and without the cache are byte-identical.  hits / stores count routes loaded / written since init. */
int mprg_set_weight_cache(mprg_ctx *ctx, const char *dir);
int mprg_weight_cache_stats(const mprg_ctx *ctx, int64_t *hits, int64_t *stores);

/* sizes of this rank's part of the route: destination points (slab), stored
 * weights, unmapped destination points, source entities (cells/nodes/points) */
int mprg_route_info(const mprg_route *rh, int64_t *nDst, int64_t *nnz, int64_t *nUnmapped, int64_t *nSrc);
/* test / weight-cache hooks: CSR of this rank's slab, 0-based, host buffers
 * rowptr[nDst+1], col[nnz], w[nnz] */
int mprg_route_export_csr(mprg_ctx *ctx, const mprg_route *rh, int32_t *rowptr, int32_t *col, double *w);
/* composed wind routes (mprg_store_wind): the entries' second weights (meridional source), w2[nnz] */
int mprg_route_export_w2(mprg_ctx *ctx, const mprg_route *rh, double *w2);
int mprg_route_import_csr(mprg_ctx *ctx, int64_t nSrc, int64_t nDst, const int32_t *rowptr,
                          const int32_t *col, const double *w, mprg_route **rh);

/* ---- apply: replaces ESMF_FieldRegrid / ESMF_FieldBundleRegrid
 *      (interp.F90:134,219,236,251,268,286,307,325,344,363,382,404,431,443).
 *      nfields stacked fields in ONE batched launch sequence.
 *      src[f]: [nSrc][nlev[f]] level-fastest == MPAS file order
 *              (input_data.F90:630,645); for MPRG_SRC_GRID_CENTER the source is a
 *              previous output on this rank's MPRG_CENTER_HALO rows,
 *              [nlev[f]][rows][ni] (level-slowest; the whole grid on one rank).
 *      dst[f]: this rank's slab [nlev[f]][nj_slab][ni] (the full grid on 1 rank).
 *      Unmapped destination points are written 0 (zeroregion=TOTAL default).
 *      src_mem / dst_mem: MPRG_HOST buffers are copied through the engine's
 *      pipelined staging; MPRG_DEVICE pointers are used in place. */
int mprg_apply(mprg_ctx *ctx, mprg_route *rh, int32_t nfields,
               const void *const *src, const int32_t *nlev, int src_dtype, int src_mem,
               void *const *dst, int dst_dtype, int dst_mem);
/* same, with one fused epilogue per field: dst = op(dst, epi_arg[f])
 * (T-300: write_data.F90:1339-1345; PHB=zgrid*9.81: write_data.F90:1417) */
int mprg_apply_ex(mprg_ctx *ctx, mprg_route *rh, int32_t nfields,
                  const void *const *src, const int32_t *nlev, int src_dtype, int src_mem,
                  void *const *dst, int dst_dtype, int dst_mem,
                  const int32_t *epi_op, const double *epi_arg);

/* ---- apply fused with the gather: like mprg_apply_ex, but dst_full[f] is the FULL field
 *      [nlev[f]][nj][ni] of the route's destination stagger (device memory) and this rank writes only
 *      its own rows into it.  On the writing rank that is its own buffer; on the other ranks it is the
 *      writing rank's buffer mapped with mprg_ipc_open, so the apply kernels store straight into the
 *      writer's memory over NVLink and ESMF_FieldGather (write_data.F90:1006-1453) needs no pass of its
 *      own.  Wind pairs (MPRG_EPI_ROT_*) are intermediates and not accepted here.  Every rank must have
 *      finished (mprg_synchronize + the host's barrier) before the writer reads the fields.
 *      mprg_ipc_export: handle (64 bytes) + offset of a device buffer of this process;
 *      mprg_ipc_open: the same buffer in this process's address space (mappings are cached per
 *      allocation and released by mprg_ipc_close_all / mprg_finalize). */
int mprg_apply_into(mprg_ctx *ctx, mprg_route *rh, int32_t nfields,
                    const void *const *src, const int32_t *nlev, int src_dtype, int src_mem,
                    void *const *dst_full, int dst_dtype, const int32_t *epi_op, const double *epi_arg);
/* this rank's slab [nlev][nj_slab][ni] -> its rows of a full field (own or mapped), on the context's stream */
int mprg_put_slab(mprg_ctx *ctx, int stagger, int32_t nlev, int dtype, const void *slab_dev, void *full_dev);
int mprg_ipc_export(mprg_ctx *ctx, const void *dev_ptr, void *handle64, size_t *offset);
int mprg_ipc_open(mprg_ctx *ctx, const void *handle64, size_t offset, void **peer_ptr);
int mprg_ipc_close_all(mprg_ctx *ctx);

/* ---- wind rotation: replaces rotate_winds_cgrid, interp.F90:689-749.
 *      cosa/sina: CENTER stagger, full grid [nj][ni], fp64 (cosa_target_grid /
 *      sina_target_grid, model_grid.F90:1113-1185).  u, v: this rank's CENTER
 *      slab [nlev][nj_slab][ni], rotated in place. */
int mprg_set_rotation(mprg_ctx *ctx, const double *cosa, const double *sina);
int mprg_has_rotation(const mprg_ctx *ctx); /* 1 once mprg_set_rotation succeeded */
int mprg_rotate_winds(mprg_ctx *ctx, void *u, void *v, int32_t nlev, int dtype, int mem);
/* same on the rows of `stagger` = MPRG_CENTER or MPRG_CENTER_HALO (u, v: [nlev][rows][ni]) */
int mprg_rotate_winds_on(mprg_ctx *ctx, int stagger, void *u, void *v, int32_t nlev, int dtype, int mem);

/* ---- the wind chain as ONE matrix per staggered grid.  interp_hist_data regrids the cell-centre
 *      winds to the mass points (interp.F90:256-289), rotates them there (rotate_winds_cgrid, :291-293) and regrids
 *      the rotated mass-point fields to the EDGE1 / EDGE2 points (:295-328).  All three steps are linear in (u, v), so
 *          U = S_U R_u W (u, v)      V = S_V R_v W (u, v)
 *      is one sparse matrix per staggered grid whose entries carry two weights (one for the zonal, one for the
 *      meridional source column).  mprg_store_wind composes it from the memoised bilinear (mesh -> CENTER_HALO) and
 *      stagger (CENTER -> EDGE) routes and the angles of mprg_set_rotation; mprg_apply_wind applies it to the pair of
 *      cell-centre sources (device buffers, file order [nCells][nlev], 16-byte aligned columns) and writes this rank's
 *      slab of the staggered field [nlev][nj_slab][ni] (into_full != 0: its rows of the full field, as
 *      mprg_apply_into).  The mass-point fields UMASS / VMASS are never materialised: one write and one read of two
 *      3-D fields and the grid-source launches disappear from the pass.  Results differ from the three-step chain by
 *      rounding only (the chain rounds the mass-point values to the output type).
 *      *rh == NULL with rc 0 means "not composable here" -- a periodic / global target grid, no rotation registered,
 *      a destination point that draws on more than 12 mesh cells, or option "wind" = "chain" -- and the caller keeps
 *      the chain.  Released with mprg_release; mprg_set_rotation drops the memoised composed routes. */
int mprg_store_wind(mprg_ctx *ctx, int dst_stagger /* MPRG_EDGE1 | MPRG_EDGE2 */, mprg_route **rh);
int mprg_apply_wind(mprg_ctx *ctx, mprg_route *rh, const void *u_src, const void *v_src, int32_t nlev, int src_dtype,
                    void *dst, int dst_dtype, int into_full);

/* ---- file byte order.  NetCDF classic / CDF-5 data is big-endian; the reference lets the NetCDF library
 *      swap on the host inside nf90_get_var / nf90_put_var (input_data.F90:186,205,437..., write_data.F90:1010...).
 *      Here file bytes cross PCIe untouched and are swapped in HBM.
 *      mprg_set_source_byte_order(1): MPRG_HOST sources of mprg_apply* hold big-endian words (a variable
 *      mapped straight from the input file); each uploaded range is swapped on the device before the apply.
 *      mprg_bswap: swap `count` words (4 bytes for MPRG_F32 / int32, 8 for MPRG_F64) of a device buffer in
 *      place -- the writer's last step before mprg_download + pwrite. */
int mprg_set_source_byte_order(mprg_ctx *ctx, int big_endian);
int mprg_bswap(mprg_ctx *ctx, void *dev, size_t count, int dtype);

/* ---- WRF-compatibility post-ops of the writer (write_data.F90:1339-1432, wrf_mod_vars), on this
 *      rank's slab, device or host buffers.  T-300 and PHB = 9.81 zgrid are fused epilogues of
 *      mprg_apply_ex; the two below need neighbouring levels / a reduction.
 *      mprg_post_midlevels: mid[k] = 0.5 (x[k+1] + x[k]), k = 0..nlev-2  (Z_C from the regridded
 *      zgrid, write_data.F90:1406-1412; x is [nlev][slab], mid is [nlev-1][slab]).
 *      mprg_post_ptop: this rank's share of P_TOP (write_data.F90:1364-1373): *maxval = max of the
 *      whole field, *mincand = min of 0.8 x[nlev-1][.] over the points whose top-level value is >= 10
 *      (+inf if none).  P_TOP = min(MAX over ranks of maxval, MIN over ranks of mincand). */
/* x = x * scale + offset over `count` values of a device buffer: T - 300 (write_data.F90:1339-1345, applied to
 * every column: the reference's `< 10` guard is a no-op `continue`), PHB = 9.81 zgrid (:1417), and the
 * all-zero MU / PH / P fields (:1356,1424,1468) with scale = offset = 0 */
int mprg_post_affine(mprg_ctx *ctx, void *dev, size_t count, int dtype, double scale, double offset);
int mprg_post_midlevels(mprg_ctx *ctx, int stagger, int32_t nlev, int dtype, int mem, const void *x, void *mid);
int mprg_post_ptop(mprg_ctx *ctx, int stagger, int32_t nlev, int dtype, int mem, const void *x, double *maxval,
                   double *mincand);

/* ---- gather: replaces ESMF_FieldGather(rootPet=0), write_data.F90:1006-1453.
 *      Collects every rank's slab of a [nlev][nj][ni] field on `root`.
 *      Device buffers; NCCL over NVLink.  With nranks == 1 it is a device copy.
 *      mprg_comm_id / mprg_comm_init bootstrap the communicator: rank 0 calls
 *      mprg_comm_id, the host program broadcasts the 128 bytes by whatever
 *      transport it already has (MPI_Bcast in the Fortran host), then every
 *      rank calls mprg_comm_init. */
int mprg_comm_id(mprg_ctx *ctx, void *id128);
int mprg_comm_init(mprg_ctx *ctx, const void *id128);
int mprg_gather(mprg_ctx *ctx, int stagger, int32_t nlev, int dtype,
                const void *slab_dev, int root, void *full_dev);
/* the same for nfields fields in ONE NCCL group (one launch, all peers and fields in flight together:
 * what write_to_file's ~20 back-to-back FieldGather calls amount to).  full_dev[] is read on root only. */
int mprg_gather_v(mprg_ctx *ctx, int32_t nfields, const int *stagger, const int32_t *nlev, int dtype,
                  const void *const *slab_dev, int root, void *const *full_dev);

/* ---- CUDA graphs: a device-buffer pass (memoised stores, mprg_apply* with MPRG_DEVICE buffers,
 *      rotations, post-ops on device buffers) issued between capture_begin and capture_end is recorded
 *      instead of run; mprg_graph_launch replays it on the context's stream with one launch call.  For
 *      time loops that regrid many output times through the same buffers: at 8 GPUs a pass is 0.6 ms of
 *      kernels and the nine separate launches show.  Nothing that allocates, synchronises or copies
 *      through host buffers may be called while capturing (rc != 0, capture abandoned). */
typedef struct mprg_graph mprg_graph;
int mprg_capture_begin(mprg_ctx *ctx);
int mprg_capture_end(mprg_ctx *ctx, mprg_graph **graph);
int mprg_graph_launch(mprg_ctx *ctx, mprg_graph *graph);
int mprg_graph_release(mprg_ctx *ctx, mprg_graph *graph);

/* ---- instrumentation */
/* number of engine kernels launched on this context since init */
int64_t mprg_kernel_launches(const mprg_ctx *ctx);
/* field bytes copied host->device / device->host by host-buffer applies and downloads since init.
 * Host sources are halo-sharded: only the id range the rank's weights reference is uploaded. */
int mprg_io_bytes(const mprg_ctx *ctx, uint64_t *h2d, uint64_t *d2h);
/* device milliseconds of the most recent store / apply (CUDA events on the
 * context's stream) */
double mprg_last_ms(const mprg_ctx *ctx);

/* per-launch timing of the apply kernels (CUDA events on the context's stream):
 * enable, run applies, then read the records accumulated since the last reset.
 * kind: 0 = columns kernel, 16-byte loads; 1 = columns kernel, 4-byte loads;
 * 2 = flat (2-D fields); 3 = planes (grid source).  alg_bytes is the algorithmic
 * traffic of that launch, units its target-point-levels.
 * mprg_profile_read synchronises the stream; returns the number of records
 * (arrays may be NULL to query the count). */
int mprg_profile_enable(mprg_ctx *ctx, int on);
int mprg_profile_read(mprg_ctx *ctx, int32_t max, int32_t *kind, double *ms, double *alg_bytes, double *units);
int mprg_profile_reset(mprg_ctx *ctx);
/* distinct source entities referenced by a route's weights (nSrcT of the roofline model) */
int64_t mprg_route_src_referenced(const mprg_route *rh);
/* the route's tile schedule: 32-target tiles, distinct source columns summed over tiles, and runs of
 * consecutively numbered columns summed over tiles.  The column kernel issues one bulk copy per run (all-
 * aligned launches), so columns / runs says how much the mesh numbering helps it. */
int mprg_route_schedule_info(const mprg_route *rh, int64_t *tiles, int64_t *columns, int64_t *runs);

#ifdef __cplusplus
}
#endif
#endif /* MPASSIT_RG_H */
