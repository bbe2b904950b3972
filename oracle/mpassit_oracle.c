/*
 * mpassit_oracle.c -- CPU restatement (fp64) of the regridding arithmetic that
 * MPASSIT delegates to ESMF.  TEST INFRASTRUCTURE ONLY: nothing under
 * mpassit_b200/ may link, import or call this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY UNPINNED: the reference (/root/reference, Fortran + ESMF 8.x + MPI +
 * NetCDF) cannot be built in this image and ships no tests, golden vectors or
 * fixtures.  The arithmetic lives in the un-vendored
 * third-party library ESMF (CMakeLists.txt:48 `find_package(ESMF 8.3.0)`,
 * author builds pin 8.6.0: modulefiles/build.jet.intel.lua:29-30).  This file
 * restates ESMF's published algorithms for exactly the argument set the
 * reference passes (interp.F90:118-128: regridmethod in {BILINEAR, CONSERVE,
 * NEAREST_STOD}, srcTermProcessing=1, unmappedaction=IGNORE, all else default:
 * lineType CART for bilinear/nearest, GREAT_CIRCLE for conserve,
 * normType DSTAREA, no masks, no extrapolation) and is pinned only by the
 * analytic known-answer tests in tests/test_oracle_kat.py.
 *
 * Each function cites the reference call site (file:line under /root/reference)
 * whose ESMF call it stands in for.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off -shared).
 * -ffp-contract=off matters: index/mask decisions must not depend on FMA
 * contraction so that a GPU build with -fmad=false reproduces them bit-exactly.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_TOL 1e-10 /* ESMF point-in-element tolerance (parametric coords) */

/* ------------------------------------------------------------------------ */
/* small vector helpers (fixed operation order; mirrored on the device)      */
/* ------------------------------------------------------------------------ */
static inline void cross3(const double *a, const double *b, double *c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double dot3(const double *a, const double *b) {
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}
static inline double dist2(const double *a, const double *b) {
    double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return (dx * dx + dy * dy) + dz * dz;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------ */
/* (1) coordinates                                                           */
/* ------------------------------------------------------------------------ */

/* model_grid.F90:450-454 (cells) and :464-468 (vertices): MPAS radians ->
 * degrees with PI = 4*atan(1) (model_grid.F90:280), longitudes > 180 wrapped
 * by -360.  These are the elemCoords / nodeCoords handed to ESMF_MeshCreate
 * (model_grid.F90:488-497). */
void orc_mesh_rad_to_deg(int64_t n, const double *lon_rad, const double *lat_rad,
                         double *lon_deg, double *lat_deg) {
    const double PI = 4.0 * atan(1.0);
    for (int64_t i = 0; i < n; ++i) {
        double lo = lon_rad[i] * 180.0 / PI;
        if (lo > 180.0) lo = lo - 360.0;
        lon_deg[i] = lo;
        lat_deg[i] = lat_rad[i] * 180.0 / PI;
    }
}

/* ESMF_COORDSYS_SPH_DEG -> unit-sphere Cartesian (what ESMF does internally
 * with the coordinates given at model_grid.F90:488-497 and :949-1038):
 *   theta = lon*DEG2RAD, phi = (90-lat)*DEG2RAD,
 *   x = cos(theta) sin(phi), y = sin(theta) sin(phi), z = cos(phi). */
void orc_sph_deg_to_cart(int64_t n, const double *lon_deg, const double *lat_deg, double *xyz) {
    const double DEG2RAD = 3.141592653589793238 / 180.0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double th = lon_deg[i] * DEG2RAD;
        double ph = (90.0 - lat_deg[i]) * DEG2RAD;
        double sp = sin(ph);
        xyz[3 * i + 0] = cos(th) * sp;
        xyz[3 * i + 1] = sin(th) * sp;
        xyz[3 * i + 2] = cos(ph);
    }
}

/* ------------------------------------------------------------------------ */
/* (2) dual mesh                                                             */
/* ------------------------------------------------------------------------ */

/* ESMF bilinear from MESHLOC_ELEMENT fields (every hist/diag field created at
 * input_data.F90:970-1132 on `input_grid`) interpolates on the DUAL mesh: one
 * dual element per original node, whose corners are the centres of the
 * elements around that node.  For an MPAS mesh (vertexDegree 3) these are the
 * Delaunay triangles.  Built here by inverting verticesOnCell, the only
 * connectivity the reference reads (model_grid.F90:409-417).  A vertex touched
 * by fewer than 3 cells (regional boundary) has no dual element: tri = -1.
 * Corner order: ascending cell id.  Returns 0, or -1 if a vertex has > 3 cells.
 * verticesOnCell: [nCells][maxEdges], 1-based, 0 = unused (model_grid.F90:448). */
int orc_dual_triangles(int32_t nCells, int32_t nVertices, int32_t maxEdges, const int32_t *voc,
                       int32_t *tri /* [nVertices][3] */) {
    int32_t *cnt = (int32_t *)calloc((size_t)nVertices, sizeof(int32_t));
    for (int64_t i = 0; i < (int64_t)nVertices * 3; ++i) tri[i] = -1;
    int rc = 0;
    for (int32_t c = 0; c < nCells; ++c) {
        for (int32_t k = 0; k < maxEdges; ++k) {
            int32_t v = voc[(int64_t)c * maxEdges + k];
            if (v <= 0) continue;
            v -= 1;
            if (v >= nVertices) { rc = -2; continue; }
            if (cnt[v] >= 3) { rc = -1; cnt[v]++; continue; }
            tri[3 * (int64_t)v + cnt[v]] = c; /* c ascending by construction */
            cnt[v]++;
        }
    }
    for (int32_t v = 0; v < nVertices; ++v)
        if (cnt[v] != 3) tri[3 * (int64_t)v] = tri[3 * (int64_t)v + 1] = tri[3 * (int64_t)v + 2] = -1;
    free(cnt);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* kd-tree over points (oracle-side acceleration; validated against the      */
/* brute-force paths in tests/test_oracle_kat.py)                            */
/* ------------------------------------------------------------------------ */
typedef struct {
    int32_t n;
    const double *xyz;
    int32_t *perm;   /* permutation; node over [lo,hi) splits at mid=(lo+hi)/2 */
    uint8_t *axis;   /* split axis stored at index mid */
} kdtree;

#define KD_LEAF 8

static void kd_select(const double *xyz, int32_t *p, int32_t lo, int32_t hi, int32_t k, int ax) {
    /* quickselect on coordinate ax, ties broken by id for determinism */
    while (hi - lo > 1) {
        int32_t piv = p[lo + (hi - lo) / 2];
        double pv = xyz[3 * (int64_t)piv + ax];
        int32_t i = lo, j = hi - 1;
        while (i <= j) {
            while (xyz[3 * (int64_t)p[i] + ax] < pv || (xyz[3 * (int64_t)p[i] + ax] == pv && p[i] < piv)) ++i;
            while (xyz[3 * (int64_t)p[j] + ax] > pv || (xyz[3 * (int64_t)p[j] + ax] == pv && p[j] > piv)) --j;
            if (i <= j) { int32_t t = p[i]; p[i] = p[j]; p[j] = t; ++i; --j; }
        }
        if (k <= j) hi = j + 1;
        else if (k >= i) lo = i;
        else return;
    }
}

static void kd_build_rec(kdtree *t, int32_t lo, int32_t hi) {
    if (hi - lo <= KD_LEAF) return;
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int32_t i = lo; i < hi; ++i)
        for (int a = 0; a < 3; ++a) {
            double v = t->xyz[3 * (int64_t)t->perm[i] + a];
            if (v < mn[a]) mn[a] = v;
            if (v > mx[a]) mx[a] = v;
        }
    int ax = 0;
    if (mx[1] - mn[1] > mx[ax] - mn[ax]) ax = 1;
    if (mx[2] - mn[2] > mx[ax] - mn[ax]) ax = 2;
    int32_t mid = lo + (hi - lo) / 2;
    kd_select(t->xyz, t->perm, lo, hi, mid, ax);
    t->axis[mid] = (uint8_t)ax;
    kd_build_rec(t, lo, mid);
    kd_build_rec(t, mid + 1, hi);
}

static kdtree *kd_build(int32_t n, const double *xyz) {
    kdtree *t = (kdtree *)malloc(sizeof(kdtree));
    t->n = n;
    t->xyz = xyz;
    t->perm = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    t->axis = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
    for (int32_t i = 0; i < n; ++i) t->perm[i] = i;
    kd_build_rec(t, 0, n);
    return t;
}
static void kd_free(kdtree *t) {
    free(t->perm);
    free(t->axis);
    free(t);
}

static void kd_nearest_rec(const kdtree *t, int32_t lo, int32_t hi, const double *q, double *best, int32_t *bid) {
    if (hi - lo <= KD_LEAF) {
        for (int32_t i = lo; i < hi; ++i) {
            int32_t id = t->perm[i];
            double d = dist2(q, t->xyz + 3 * (int64_t)id);
            if (d < *best || (d == *best && id < *bid)) { *best = d; *bid = id; }
        }
        return;
    }
    int32_t mid = lo + (hi - lo) / 2;
    int32_t id = t->perm[mid];
    int ax = t->axis[mid];
    double d = dist2(q, t->xyz + 3 * (int64_t)id);
    if (d < *best || (d == *best && id < *bid)) { *best = d; *bid = id; }
    double dp = q[ax] - t->xyz[3 * (int64_t)id + ax];
    if (dp <= 0.0) {
        kd_nearest_rec(t, lo, mid, q, best, bid);
        if (dp * dp <= *best) kd_nearest_rec(t, mid + 1, hi, q, best, bid);
    } else {
        kd_nearest_rec(t, mid + 1, hi, q, best, bid);
        if (dp * dp <= *best) kd_nearest_rec(t, lo, mid, q, best, bid);
    }
}

typedef void (*kd_visit)(int32_t id, void *ctx);
static void kd_radius_rec(const kdtree *t, int32_t lo, int32_t hi, const double *q, double r2, kd_visit f, void *ctx) {
    if (hi - lo <= KD_LEAF) {
        for (int32_t i = lo; i < hi; ++i) {
            int32_t id = t->perm[i];
            if (dist2(q, t->xyz + 3 * (int64_t)id) <= r2) f(id, ctx);
        }
        return;
    }
    int32_t mid = lo + (hi - lo) / 2;
    int32_t id = t->perm[mid];
    int ax = t->axis[mid];
    if (dist2(q, t->xyz + 3 * (int64_t)id) <= r2) f(id, ctx);
    double dp = q[ax] - t->xyz[3 * (int64_t)id + ax];
    if (dp <= 0.0 || dp * dp <= r2) kd_radius_rec(t, lo, mid, q, r2, f, ctx);
    if (dp >= 0.0 || dp * dp <= r2) kd_radius_rec(t, mid + 1, hi, q, r2, f, ctx);
}

/* ------------------------------------------------------------------------ */
/* (3) NEAREST_STOD  (interp.F90:420-431 nstd bundle; :436-443 soil bundle   */
/*     when `method` still holds NEAREST_STOD)                               */
/* ------------------------------------------------------------------------ */
/* Nearest source point by 3-D Cartesian distance; ties -> smallest source
 * index; every destination is mapped.  idx is 0-based.  brute!=0 scans all. */
int orc_nearest(int32_t nSrc, const double *sxyz, int64_t nDst, const double *dxyz, int32_t *idx, int brute) {
    if (nSrc <= 0) return -1;
    kdtree *t = brute ? NULL : kd_build(nSrc, sxyz);
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < nDst; ++i) {
        const double *q = dxyz + 3 * i;
        double best = 1e300;
        int32_t bid = 0x7fffffff;
        if (brute) {
            for (int32_t s = 0; s < nSrc; ++s) {
                double d = dist2(q, sxyz + 3 * (int64_t)s);
                if (d < best) { best = d; bid = s; } /* ascending s => lowest id on ties */
            }
        } else {
            kd_nearest_rec(t, 0, nSrc, q, &best, &bid);
        }
        idx[i] = bid;
    }
    if (t) kd_free(t);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* (4) BILINEAR, Mesh(element) -> Grid  (interp.F90:123 diag; :207 2d_patch;  */
/*     :226 hgt; :241 3d_nz; :259,:277 u,v; :334 3d_nzp1)                    */
/* ------------------------------------------------------------------------ */
/* Point-in-dual-triangle + weights, ESMF lineType CART: intersect the ray
 * origin->p with the plane of the flat triangle (v0,v1,v2):
 *     v0 + a (v1-v0) + b (v2-v0) = s p
 * accept iff a >= -tol, b >= -tol, a+b <= 1+tol, s > 0; weights (1-a-b, a, b).
 * Returns 1 if accepted. */
static inline int tri_locate(const double *v0, const double *v1, const double *v2, const double *p, double *w) {
    double e1[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]};
    double e2[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
    double q[3], r[3], n[3];
    cross3(e2, p, q);            /* q = e2 x p  */
    double det = dot3(e1, q);    /* e1 . (e2 x p) */
    if (det == 0.0) return 0;
    cross3(v0, p, r);            /* r = v0 x p  */
    cross3(e2, v0, n);           /* n = e2 x v0 */
    double a = -dot3(v0, q) / det;
    double b = -dot3(e1, r) / det;
    double s = dot3(e1, n) / det;
    if (!(s > 0.0)) return 0;
    if (a < -ORC_TOL || b < -ORC_TOL || (a + b) > 1.0 + ORC_TOL) return 0;
    w[0] = (1.0 - a) - b;
    w[1] = a;
    w[2] = b;
    return 1;
}

typedef struct {
    const double *cxyz;
    const int32_t *tri;
    const int32_t *voc;
    int32_t maxEdges;
    const double *p;
    int32_t best;
    double w[3];
} bil_ctx;

static void bil_try(int32_t v, bil_ctx *c) {
    const int32_t *t = c->tri + 3 * (int64_t)v;
    if (t[0] < 0) return;
    if (c->best >= 0 && v >= c->best) return; /* smallest dual-element id wins */
    double w[3];
    if (tri_locate(c->cxyz + 3 * (int64_t)t[0], c->cxyz + 3 * (int64_t)t[1], c->cxyz + 3 * (int64_t)t[2], c->p, w)) {
        c->best = v;
        c->w[0] = w[0]; c->w[1] = w[1]; c->w[2] = w[2];
    }
}
static void bil_visit_cell(int32_t cell, void *vctx) {
    bil_ctx *c = (bil_ctx *)vctx;
    for (int32_t k = 0; k < c->maxEdges; ++k) {
        int32_t v = c->voc[(int64_t)cell * c->maxEdges + k];
        if (v > 0) bil_try(v - 1, c);
    }
}

/* elem[i] = winning dual-element id (= MPAS vertex id, 0-based) or -1 if the
 * point is in no dual triangle (unmapped: UNMAPPEDACTION_IGNORE, interp.F90:127,
 * destination stays at the zero fill).  col [nDst][3] (ascending cell ids),
 * w [nDst][3]. */
int orc_bilinear(int32_t nCells, const double *cxyz, int32_t nVertices, const int32_t *tri,
                 int32_t maxEdges, const int32_t *voc, int64_t nDst, const double *dxyz,
                 int32_t *elem, int32_t *col, double *w, int brute) {
    kdtree *t = NULL;
    double r2 = 0.0;
    if (!brute) {
        /* any triangle whose cone holds p has all corners within its longest
         * edge of p; search radius = longest dual edge (+ slack). */
        double m2 = 0.0;
        for (int32_t v = 0; v < nVertices; ++v) {
            const int32_t *tv = tri + 3 * (int64_t)v;
            if (tv[0] < 0) continue;
            for (int k = 0; k < 3; ++k) {
                double d = dist2(cxyz + 3 * (int64_t)tv[k], cxyz + 3 * (int64_t)tv[(k + 1) % 3]);
                if (d > m2) m2 = d;
            }
        }
        double r = sqrt(m2) * 1.001 + 1e-9;
        r2 = r * r;
        t = kd_build(nCells, cxyz);
    }
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < nDst; ++i) {
        bil_ctx c;
        c.cxyz = cxyz; c.tri = tri; c.voc = voc; c.maxEdges = maxEdges;
        c.p = dxyz + 3 * i; c.best = -1; c.w[0] = c.w[1] = c.w[2] = 0.0;
        if (brute) {
            for (int32_t v = 0; v < nVertices; ++v) bil_try(v, &c);
        } else {
            kd_radius_rec(t, 0, nCells, c.p, r2, bil_visit_cell, &c);
        }
        elem[i] = c.best;
        if (c.best >= 0) {
            const int32_t *tv = tri + 3 * (int64_t)c.best;
            for (int k = 0; k < 3; ++k) { col[3 * i + k] = tv[k]; w[3 * i + k] = c.w[k]; }
        } else {
            for (int k = 0; k < 3; ++k) { col[3 * i + k] = -1; w[3 * i + k] = 0.0; }
        }
    }
    if (t) kd_free(t);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* (5) weight application (ESMF_FieldRegrid / FieldBundleRegrid,             */
/*     interp.F90:134,219,236,251,268,286,307,325,344,363,382,404,431,443)   */
/* ------------------------------------------------------------------------ */
/* dst zeroed first (zeroregion=TOTAL default) then dst = W src in fp64, terms
 * added in CSR order.  src is [nSrc][nlev] level-fastest (MPAS file order,
 * input_data.F90:630,645); dst is [nlev][nDst] (C order of the Fortran
 * (i,j,lev) target arrays, interp.F90:155-160).  Output fp32 is the fp64 sum
 * rounded once (what NetCDF does when write_data.F90:587 stores R8 into
 * NF90_FLOAT). */
#define ORC_APPLY(NAME, TIN, TOUT)                                                                   \
    void NAME(int64_t nDst, const int32_t *rowptr, const int32_t *col, const double *w, int32_t nlev, \
              const TIN *src, TOUT *dst) {                                                            \
        _Pragma("omp parallel for schedule(static)") for (int64_t t = 0; t < nDst; ++t) {             \
            int32_t b = rowptr[t], e = rowptr[t + 1];                                                 \
            for (int32_t l = 0; l < nlev; ++l) {                                                      \
                double acc = 0.0;                                                                     \
                for (int32_t k = b; k < e; ++k) acc = acc + w[k] * (double)src[(int64_t)col[k] * nlev + l]; \
                dst[(int64_t)l * nDst + t] = (TOUT)acc;                                               \
            }                                                                                         \
        }                                                                                             \
    }
ORC_APPLY(orc_apply_f32_f32, float, float)
ORC_APPLY(orc_apply_f32_f64, float, double)
ORC_APPLY(orc_apply_f64_f64, double, double)
ORC_APPLY(orc_apply_f64_f32, double, float)

/* Cache-blocked variant used only for CPU-baseline timing (same arithmetic,
 * level loop innermost over a target tile so stores are contiguous). */
void orc_apply_f32_f32_tiled(int64_t nDst, const int32_t *rowptr, const int32_t *col, const double *w,
                             int32_t nlev, const float *src, float *dst) {
    enum { TB = 64 };
#pragma omp parallel
    {
        double *buf = (double *)malloc(sizeof(double) * (size_t)TB * (size_t)(nlev > 0 ? nlev : 1));
#pragma omp for schedule(static)
        for (int64_t t0 = 0; t0 < nDst; t0 += TB) {
            int64_t t1 = t0 + TB < nDst ? t0 + TB : nDst;
            for (int64_t t = t0; t < t1; ++t) {
                double *acc = buf + (t - t0) * nlev;
                for (int32_t l = 0; l < nlev; ++l) acc[l] = 0.0;
                for (int32_t k = rowptr[t]; k < rowptr[t + 1]; ++k) {
                    const float *s = src + (int64_t)col[k] * nlev;
                    double wk = w[k];
                    for (int32_t l = 0; l < nlev; ++l) acc[l] = acc[l] + wk * (double)s[l];
                }
            }
            for (int32_t l = 0; l < nlev; ++l)
                for (int64_t t = t0; t < t1; ++t) dst[(int64_t)l * nDst + t] = (float)buf[(t - t0) * nlev + l];
        }
        free(buf);
    }
}

/* ------------------------------------------------------------------------ */
/* (6) rotate_winds_cgrid  (interp.F90:689-749)                              */
/* ------------------------------------------------------------------------ */
/* Earth-relative -> grid-relative at mass points; v' uses the ALREADY
 * rotated u' (interp.F90:741-742 / :744-745).  u, v: [nlev][n]; cosa, sina: [n]. */
void orc_rotate_winds(int64_t n, int32_t nlev, double *u, double *v, const double *cosa, const double *sina) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double tana = sina[i] / cosa[i];
        for (int32_t l = 0; l < nlev; ++l) {
            double uu = u[(int64_t)l * n + i], vv = v[(int64_t)l * n + i];
            uu = (uu + vv * tana) / (cosa[i] + sina[i] * tana);
            vv = (vv - uu * sina[i]) / cosa[i];
            u[(int64_t)l * n + i] = uu;
            v[(int64_t)l * n + i] = vv;
        }
    }
}
void orc_rotate_winds_f32(int64_t n, int32_t nlev, float *u, float *v, const double *cosa, const double *sina) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double tana = sina[i] / cosa[i];
        for (int32_t l = 0; l < nlev; ++l) {
            double uu = (double)u[(int64_t)l * n + i], vv = (double)v[(int64_t)l * n + i];
            uu = (uu + vv * tana) / (cosa[i] + sina[i] * tana);
            vv = (vv - uu * sina[i]) / cosa[i];
            u[(int64_t)l * n + i] = (float)uu;
            v[(int64_t)l * n + i] = (float)vv;
        }
    }
}
