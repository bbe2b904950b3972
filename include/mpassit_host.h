/*
 * mpassit_host.h -- C ABI of libmpassit_host.so: the host-side mirror (C++) of the
 * reference's Fortran stages either side of the regridding engine.  It exists
 * because this image has no Fortran compiler; a Fortran host binds the engine
 * directly through fortran/mpassit_rg_mod.F90 instead.
 *
 * Same names, argument meaning and error behaviour as the reference routines
 * cited at each declaration (file:line under /root/reference).  All reals are
 * fp64: the reference is built with -r8 / -fdefault-real-8 (CMakeLists.txt:80-82).
 */
#ifndef MPASSIT_HOST_H
#define MPASSIT_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "mpassit_rg.h"

#ifdef __cplusplus
extern "C" {
#endif

#define MPASSIT_NAN 1.0e20 /* "unset" sentinel, misc_definitions_module.F90:12 */
#define MPASSIT_STRLEN 512
#define MPASSIT_NAMELEN 64

/* projection codes, misc_definitions_module.F90:38-47 */
enum { MPASSIT_PROJ_LATLON = 0, MPASSIT_PROJ_LC = 1, MPASSIT_PROJ_PS = 2, MPASSIT_PROJ_MERC = 3 };
/* stagger ids used by xytoll, misc_definitions_module.F90:30 */
enum { MPASSIT_M = 1, MPASSIT_U = 2, MPASSIT_V = 3, MPASSIT_CORNER = 6 };

/* &config namelist + derived module variables, program_setup.F90:22-75,103-106,155-243 */
typedef struct mpassit_config {
    char grid_file_input_grid[MPASSIT_STRLEN];
    char diag_file_input_grid[MPASSIT_STRLEN];
    char hist_file_input_grid[MPASSIT_STRLEN];
    char file_target_grid[MPASSIT_STRLEN];
    char output_file[MPASSIT_STRLEN];
    char block_decomp_file[MPASSIT_STRLEN];
    char target_grid_type[MPASSIT_STRLEN];
    int32_t interp_diag, interp_hist, wrf_mod_vars, esmf_log, is_regional, interp_as_bundle;
    int32_t nx, ny;
    double dx, dy, ref_lat, ref_lon, ref_x, ref_y, truelat1, truelat2, stand_lon, pole_lat, pole_lon;
    /* derived (program_setup.F90:155-243) */
    int32_t i_target, j_target, proj_code;
    double dxkm, dykm, dlondeg, dlatdeg, known_lat, known_lon, known_x, known_y;
    char map_proj_char[MPASSIT_NAMELEN];
} mpassit_config;

/* read_setup_namelist, program_setup.F90:87-249.  rc 0 = ok; on failure `err`
 * holds the reference's error_handler message. */
int mpassit_read_setup_namelist(const char *filename, mpassit_config *cfg, char *err, size_t errlen);

/* read_varlist, input_data.F90:1146-1194: two whitespace-separated columns per
 * non-blank line.  names are [max][MPASSIT_NAMELEN]. */
int mpassit_read_varlist(const char *file, int32_t max, int32_t *nfields, char *field_names,
                         char *field_names_target, char *err, size_t errlen);

/* regrid classes of init_input_hist_fields, input_data.F90:840-911 */
enum {
    MPASSIT_CLASS_2D_PATCH = 0, MPASSIT_CLASS_2D_CONS = 1, MPASSIT_CLASS_2D_NSTD = 2,
    MPASSIT_CLASS_3D_NZ = 3, MPASSIT_CLASS_3D_NZP1 = 4, MPASSIT_CLASS_3D_VERT = 5,
    MPASSIT_CLASS_U = 6, MPASSIT_CLASS_V = 7, MPASSIT_CLASS_SOIL = 8, MPASSIT_CLASS_DIAG_2D = 9,
    MPASSIT_CLASS_DIAG_3D = 10
};
int mpassit_classify_hist_2d(const char *mpas_name);                   /* input_data.F90:858-866 */
int mpassit_classify_hist_3d(const char *mpas_name, int wrf_mod_vars); /* input_data.F90:896-911 */
int mpassit_classify_diag(const char *mpas_name);                      /* input_data.F90:283 */

/* para_range, model_grid.F90:2428-2441 (1-based inclusive, as in the reference) */
void mpassit_para_range(int32_t n1, int32_t n2, int32_t nprocs, int32_t irank, int32_t *ista, int32_t *iend);
/* read_block_decomp_file, model_grid.F90:2367-2426: owner[k] = rank of cell k (0-based) */
int mpassit_read_block_decomp_file(const char *file, int32_t ncells, int32_t npets, int32_t *owner,
                                   char *err, size_t errlen);

/* define_target_grid_params, model_grid.F90:644-1201.
 * Size of each ESMF stagger (MPRG_CENTER/EDGE1/EDGE2/CORNER). */
int mpassit_target_dims(const mpassit_config *cfg, int mprg_stagger, int32_t *ni, int32_t *nj);
/* get_lat_lon_fields + xytoll + ij_to_latlon (model_grid.F90:2188-2219,
 * llxy_module.F90:166-216, module_map_utils.F90:1160-1233,1398-1428);
 * lat, lon: [nj][ni] degrees.  LC and lat-lon projections. */
/* push_source_projection + map_set / set_lc (llxy_module.F90:38-160, module_map_utils.F90:243-568, 1083-1121): the
 * per-grid scalars of the target projection, for mprg_set_target_projected (coordinates generated on the device) */
int mpassit_projection(const mpassit_config *cfg, mprg_projection *out, char *err, size_t errlen);
int mpassit_target_coords(const mpassit_config *cfg, int mprg_stagger, double *lat, double *lon,
                          char *err, size_t errlen);
/* get_cell_corners, model_grid.F90:1902-1972 (target_grid_type = 'file'): CORNER-stagger points [nj+1][ni+1]
 * synthesised from the mass points [nj][ni] and the cell size, bearings and truncated pi as in the reference */
void mpassit_get_cell_corners(const double *lat, const double *lon, int32_t ni, int32_t nj, double dx, double *clat,
                              double *clon);
/* get_rotang, model_grid.F90:2450-2507, on an [nj][ni] lat/lon pair */
void mpassit_get_rotang(const double *lat, const double *lon, int32_t ni, int32_t nj, double *cosa, double *sina);

/* ---- interp_data, interp.F90:92-465 ------------------------------------- */
typedef struct mpassit_field {
    char name[MPASSIT_NAMELEN];        /* MPAS name (var-list column 1) */
    char target_name[MPASSIT_NAMELEN]; /* output name (var-list column 2) */
    int32_t nlev;                      /* 1, nVertLevels, nVertLevelsP1 or nSoilLevels */
    int32_t klass;                     /* MPASSIT_CLASS_*, filled by mpassit_classify_fields */
    const void *src;                   /* [nCells|nVertices][nlev] level-fastest (file order) */
    void *dst;                         /* this rank's slab [nlev][nj_slab][ni] (caller allocated) */
} mpassit_field;

typedef struct mpassit_interp_io {
    int32_t src_dtype, dst_dtype;      /* MPRG_F32 / MPRG_F64 */
    int32_t mem;                       /* MPRG_HOST or MPRG_DEVICE for every src/dst below */
    int32_t nz;                        /* nVertLevels (nz_input) */
    int32_t n_diag;  mpassit_field *diag;      /* diaglist order */
    int32_t n_hist_2d; mpassit_field *hist_2d; /* histlist_2d order */
    int32_t n_hist_3d; mpassit_field *hist_3d; /* histlist_3d order (incl. uReconstruct*) */
    int32_t n_soil;  mpassit_field *soil;      /* histlist_soil order */
    const void *ter; void *hgt;                /* hgt_input_grid -> HGT (interp.F90:226-238); may be NULL */
    /* wrf_mod_vars wind chain (interp.F90:256-328): U on EDGE1 [nz][nj][ni+1], V on EDGE2
     * [nz][nj+1][ni]; the rotated mass-point winds (UMASS/VMASS, never written to the
     * output file) stay on the device */
    void *u_stag, *v_stag;
    /* dst_full != 0 (device buffers only): every dst / hgt / u_stag / v_stag above is the FULL field
     * [nlev][nj][ni] of its stagger -- the writing rank's own buffer, or that buffer mapped with
     * mprg_ipc_open on the other ranks -- and each rank stores its rows straight into it
     * (mprg_apply_into): the FieldGather of write_to_file is fused into the regrid */
    int32_t dst_full;
    /* dst_device != 0 with mem == MPRG_HOST: sources are host buffers (e.g. variables mapped from the input
     * file) but every dst / hgt / u_stag / v_stag is a DEVICE slab -- the file writer keeps the outputs in
     * HBM for the WRF post-ops and downloads them in file byte order */
    int32_t dst_device;
    /* wind rotation (interp.F90:138,291) uses the angles registered once with mprg_set_rotation,
     * the analogue of cosa/sina_target_grid created in define_target_grid (model_grid.F90:1113-1185) */
} mpassit_interp_io;

/* fills klass for every field (input_data.F90:283,858-911) and returns the
 * do_u_interp / do_v_interp / do_u10 / do_v10 flags */
int mpassit_classify_fields(const mpassit_config *cfg, mpassit_interp_io *io, int32_t *do_u, int32_t *do_v,
                            int32_t *u10_ind, int32_t *v10_ind);
/* interp_data: every Store/Regrid pair in the reference's order, through the engine */
int mpassit_interp_data(mprg_ctx *ctx, const mpassit_config *cfg, mpassit_interp_io *io, char *err, size_t errlen);


/* xytoll, llxy_module.F90:166-216: grid index (x, y) of stagger MPASSIT_M/_U/_V/_CORNER -> lat, lon (degrees) */
int mpassit_xytoll(const mpassit_config *cfg, double x, double y, int stagger, double *lat, double *lon);
/* get_map_factor, model_grid.F90:2229-2365, on n latitudes (Lambert; 0 for lat-lon targets, which the
 * reference leaves unassigned) */
int mpassit_get_map_factor(const mpassit_config *cfg, const double *xlat, int64_t n, double *mapfac);

/* ---- program mpassit, mpassit.F90:23-146, files on both sides ------------------------------------------
 * read_setup_namelist -> define_target_grid -> define_input_grid -> read_input_data -> interp_data ->
 * write_to_file for this rank's row slab.  Input files: NetCDF classic (CDF-1/2/5) MPAS grid / diag / history
 * files, mapped and consumed in file order and byte order (no transpose, no host swap); with target_grid_type =
 * 'file' the target comes from a WRF-style file (define_target_grid_file, model_grid.F90:1203-1890); output: one NetCDF
 * classic file holding the reference's dimensions, variables and attributes (write_data.F90:170-994), each
 * rank writing its own rows with pwrite.  The var-list files diaglist / histlist_2d / histlist_3d /
 * histlist_soil are read from `varlist_dir` (NULL = the working directory, as the reference does).
 * `comm` stands in for the MPI the reference's host already has: it is called collectively by every rank
 * with op = MPASSIT_COMM_BARRIER (vals NULL), or MPASSIT_COMM_MAX / _MIN to all-reduce vals[0..n) in place
 * (P_TOP, write_data.F90:1364-1373).  A rank that fails calls it ONCE, not collectively, with op =
 * MPASSIT_COMM_ABORT and vals[0] = its rc before it returns: the counterpart of error_handler's mpi_abort
 * (utils.F90:16-33) -- the communicator must bring the other ranks down (MPI_Abort), which may be blocked in a
 * barrier or a reduction.  May be NULL when nranks == 1. */
enum { MPASSIT_COMM_BARRIER = 0, MPASSIT_COMM_MAX = 1, MPASSIT_COMM_MIN = 2, MPASSIT_COMM_ABORT = 3 };
typedef void (*mpassit_comm_fn)(void *arg, int op, double *vals, int n);
typedef struct mpassit_run_stats {
    double setup_ms, read_ms, interp_ms, write_ms, total_ms; /* wall clock of the stages on this rank */
    double init_ms, target_ms, gridfile_ms, mesh_ms;         /* setup_ms split: mprg_init, target grid, grid file, mprg_set_mesh */
    double alloc_ms;                                         /* output slabs in HBM (between read and interp) */
    double download_ms, writer_wait_ms;                      /* write_ms split: device->host copies; waiting for the pwrite thread */
    int64_t n_cells, bytes_in, bytes_out;                    /* source bytes referenced in the input files; bytes this rank wrote */
    int32_t n_vars_written, output_version;                  /* regridded variables; 2 = CDF-2, 5 = CDF-5 */
    double p_top;
} mpassit_run_stats;
int mpassit_run(const char *namelist_file, const char *varlist_dir, int device, int rank, int nranks,
                mpassit_comm_fn comm, void *comm_arg, mpassit_run_stats *stats, char *err, size_t errlen);

/* ---- NetCDF classic container (host/ncio.cpp), test hooks.  No CUDA involved.
 * mpassit_nc_describe: one line per header item -- "version V numrecs N", "dim NAME LEN", "gatt NAME TYPE NELEMS",
 *   "var NAME TYPE BEGIN DIM...", "vatt VAR NAME TYPE NELEMS" -- into out (truncated to outlen).
 * mpassit_nc_get: elements [first, first+n) of record `rec` of a numeric variable as doubles.
 * mpassit_nc_copy: read src with the reader and write every dimension, attribute, variable and record again
 *   with the writer in format `version` (1, 2, 5; 0 = automatic). */
int mpassit_nc_describe(const char *path, char *out, size_t outlen, char *err, size_t errlen);
int mpassit_nc_get(const char *path, const char *var, int64_t rec, int64_t first, int64_t n, double *out,
                   char *err, size_t errlen);
int mpassit_nc_copy(const char *src, const char *dst, int version, char *err, size_t errlen);


/* ---- ESMF regrid weight files (host/weights.cpp): dimensions n_a / n_b / n_s, variables col(n_s), row(n_s) (1-based)
 * and S(n_s) -- what `ESMF_RegridWeightGen -w` and ESMF_SparseMatrixWrite produce.  The bridge for pinning parity on a
 * machine that has ESMF: read ESMF's matrix as a 0-based CSR (rows sorted, ESMF's order kept inside a row) to compare
 * with mprg_route_export_csr or to run the engine on it (mprg_route_import_csr); write the engine's matrix in the
 * same format for ESMF-side tools.  (Classic NetCDF has no zero-length fixed dimension: an empty matrix is written
 * as one entry (row 1, col 1) with weight 0.) */
int mpassit_weights_sizes(const char *path, int64_t *n_a, int64_t *n_b, int64_t *n_s, char *err, size_t errlen);
int mpassit_weights_read_csr(const char *path, int64_t n_b, int64_t n_s, int32_t *rowptr /*[n_b+1]*/, int32_t *col /*[n_s]*/,
                             double *w /*[n_s]*/, char *err, size_t errlen);
int mpassit_weights_write(const char *path, int64_t n_a, int64_t n_b, const int32_t *rowptr, const int32_t *col,
                          const double *w, const char *method, char *err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif
