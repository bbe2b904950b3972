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
 * Build: see oracle/Makefile (gcc -O3 -march=x86-64-v3 -fopenmp -ffp-contract=off -shared).
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

/* ------------------------------------------------------------------------ */
/* (7) BILINEAR, Grid(CENTER) -> Grid(EDGE1 / EDGE2)  (interp.F90:298,316)    */
/* ------------------------------------------------------------------------ */
/* Source elements are the quads of 4 adjacent centre points; element id =
 * j*(ni-1)+i (grid sequence order, i fastest).  ESMF maps a point into a quad
 * by Newton iteration on   q0 + u B + v C + u v A = t p   (B=q1-q0, C=q3-q0,
 * A=q0-q1+q2-q3; start (0,0,0); stop when |F|^2 < 1e-20; at most 100 steps),
 * accepts iff u,v in [-tol, 1+tol] and t > 0; weights
 * ((1-u)(1-v), u(1-v), u v, (1-u) v) on (q0,q1,q2,q3) = (i,j),(i+1,j),(i+1,j+1),(i,j+1). */
static int invert3(const double *J, double *inv) {
    double c00 = J[4] * J[8] - J[5] * J[7];
    double c01 = J[5] * J[6] - J[3] * J[8];
    double c02 = J[3] * J[7] - J[4] * J[6];
    double det = (J[0] * c00 + J[1] * c01) + J[2] * c02;
    if (det == 0.0) return 0;
    double id = 1.0 / det;
    inv[0] = c00 * id;
    inv[1] = (J[2] * J[7] - J[1] * J[8]) * id;
    inv[2] = (J[1] * J[5] - J[2] * J[4]) * id;
    inv[3] = c01 * id;
    inv[4] = (J[0] * J[8] - J[2] * J[6]) * id;
    inv[5] = (J[2] * J[3] - J[0] * J[5]) * id;
    inv[6] = c02 * id;
    inv[7] = (J[1] * J[6] - J[0] * J[7]) * id;
    inv[8] = (J[0] * J[4] - J[1] * J[3]) * id;
    return 1;
}

static int quad_locate(const double *q0, const double *q1, const double *q2, const double *q3, const double *p,
                       double *w) {
    double A[3], B[3], C[3], X[3] = {0.0, 0.0, 0.0}, F[3], J[9], inv[9];
    for (int d = 0; d < 3; ++d) {
        A[d] = ((q0[d] - q1[d]) + q2[d]) - q3[d];
        B[d] = q1[d] - q0[d];
        C[d] = q3[d] - q0[d];
    }
    for (int it = 0; it < 100; ++it) {
        for (int d = 0; d < 3; ++d)
            F[d] = (((X[0] * X[1]) * A[d] + X[0] * B[d]) + X[1] * C[d]) - X[2] * p[d] + q0[d];
        if ((F[0] * F[0] + F[1] * F[1]) + F[2] * F[2] < 1.0e-20) break;
        for (int d = 0; d < 3; ++d) {
            J[3 * d + 0] = A[d] * X[1] + B[d];
            J[3 * d + 1] = A[d] * X[0] + C[d];
            J[3 * d + 2] = -p[d];
        }
        if (!invert3(J, inv)) return 0;
        for (int r = 0; r < 3; ++r)
            X[r] = X[r] - ((inv[3 * r] * F[0] + inv[3 * r + 1] * F[1]) + inv[3 * r + 2] * F[2]);
    }
    double u = X[0], v = X[1], t = X[2];
    if (!(t > 0.0)) return 0;
    if (!(u >= -ORC_TOL && u <= 1.0 + ORC_TOL && v >= -ORC_TOL && v <= 1.0 + ORC_TOL)) return 0;
    w[0] = (1.0 - u) * (1.0 - v);
    w[1] = u * (1.0 - v);
    w[2] = u * v;
    w[3] = (1.0 - u) * v;
    return 1;
}

/* Grid topology of the centre -> edge source grid (the reference's target_grid):
 *   ORC_TOPO_NOPERI  ESMF_GridCreateNoPeriDim (model_grid.F90:698-703, regional targets): (ni-1) x (nj-1) quads;
 *   ORC_TOPO_PERI    bit 0: ESMF_GridCreate1PeriDim(periodicDim=1) (model_grid.F90:684-696, is_regional=.false.):
 *                    ni quads per row, the last one joins column ni-1 to column 0 across the seam;
 *   ORC_TOPO_SPOLE / ORC_TOPO_NPOLE  bits 1, 2: polekindflag = MONOPOLE at the bottom / top row of THIS row block
 *                    (a row block carries the cap only when it holds grid row 0 / nj-1).  ESMF's default
 *                    polemethod (ALLAVG) closes the grid with an artificial pole node "in the centre of the row,
 *                    projected onto the sphere", whose value is the average of the row's ni values; the cap is
 *                    the fan of triangles (pole, P_i, P_i+1).  A destination point inside cap triangle i with
 *                    triangle weights (ws, a, b) therefore has ni entries: ws/ni on every point of the row, plus
 *                    a on P_i and b on P_i+1.
 * Element ids (tie rule: smallest id wins): quads j*nqi + i, then the south cap's triangles, then the north cap's.
 * Output stays 4 wide: quads -> col = (q0,q1,q2,q3), w = bilinear weights; cap triangles -> elem >= nquads,
 * col = (P_i, P_i+1, first point of the pole row, -1), w = (a, b, ws, 0): orc.quadgrid_to_csr expands them. */
#define ORC_TOPO_PERI 1
#define ORC_TOPO_SPOLE 2
#define ORC_TOPO_NPOLE 4

typedef struct {
    const double *sxyz;
    int32_t ni, nj, nqi, topo;
    int64_t nquads;
    const double *spole, *npole;
    const double *p;
    int64_t best;
    int32_t col[4];
    double w[4];
} quad_ctx;

static void quad_try(int32_t i, int32_t j, quad_ctx *c) {
    if (c->topo & ORC_TOPO_PERI) i = (i % c->ni + c->ni) % c->ni;
    if (i < 0 || j < 0 || i >= c->nqi || j >= c->nj - 1) return;
    int64_t id = (int64_t)j * c->nqi + i;
    if (c->best >= 0 && id >= c->best) return;
    int32_t i1 = (i + 1 == c->ni) ? 0 : i + 1;
    int32_t b0 = j * c->ni + i, b1 = j * c->ni + i1;
    const double *q0 = c->sxyz + 3 * (int64_t)b0, *q1 = c->sxyz + 3 * (int64_t)b1;
    const double *q3 = q0 + 3 * (int64_t)c->ni, *q2 = q1 + 3 * (int64_t)c->ni;
    double w[4];
    if (quad_locate(q0, q1, q2, q3, c->p, w)) {
        c->best = id;
        c->col[0] = b0; c->col[1] = b1; c->col[2] = b1 + c->ni; c->col[3] = b0 + c->ni;
        for (int k = 0; k < 4; ++k) c->w[k] = w[k];
    }
}
/* cap triangle i of the south (north = 0) or north (north = 1) pole */
static void cap_try(int32_t i, int north, quad_ctx *c) {
    if (!(c->topo & (north ? ORC_TOPO_NPOLE : ORC_TOPO_SPOLE))) return;
    i = (i % c->ni + c->ni) % c->ni;
    int64_t id = c->nquads + ((north && (c->topo & ORC_TOPO_SPOLE)) ? c->ni : 0) + i;
    if (c->best >= 0 && id >= c->best) return;
    int32_t row = north ? c->nj - 1 : 0;
    int32_t i1 = (i + 1 == c->ni) ? 0 : i + 1;
    int32_t a = row * c->ni + i, b = row * c->ni + i1;
    double w[3];
    if (tri_locate(north ? c->npole : c->spole, c->sxyz + 3 * (int64_t)a, c->sxyz + 3 * (int64_t)b, c->p, w)) {
        c->best = id;
        c->col[0] = a; c->col[1] = b; c->col[2] = row * c->ni; c->col[3] = -1;
        c->w[0] = w[1]; c->w[1] = w[2]; c->w[2] = w[0]; c->w[3] = 0.0;
    }
}
static void quad_visit_point(int32_t pt, void *vctx) {
    quad_ctx *c = (quad_ctx *)vctx;
    int32_t i = pt % c->ni, j = pt / c->ni;
    quad_try(i - 1, j - 1, c);
    quad_try(i, j - 1, c);
    quad_try(i - 1, j, c);
    quad_try(i, j, c);
}

/* artificial pole node: centre of the row's points, projected onto the unit sphere */
static void pole_node(const double *row, int32_t ni, double *out) {
    double s[3] = {0.0, 0.0, 0.0};
    for (int32_t i = 0; i < ni; ++i)
        for (int d = 0; d < 3; ++d) s[d] = s[d] + row[3 * (int64_t)i + d];
    double inv = 1.0 / sqrt(dot3(s, s));
    for (int d = 0; d < 3; ++d) out[d] = s[d] * inv;
}

/* elem = winning element id or -1; col [nDst][4], w [nDst][4] as described above. */
int orc_bilinear_quadgrid_topo(int32_t ni, int32_t nj, const double *sxyz, int64_t nDst, const double *dxyz,
                               int64_t *elem, int32_t *col, double *w, int brute, int topo) {
    if (ni < 2 || nj < 2) return -1;
    const int32_t nqi = (topo & ORC_TOPO_PERI) ? ni : ni - 1;
    double spole[3] = {0, 0, -1}, npole[3] = {0, 0, 1};
    if (topo & ORC_TOPO_SPOLE) pole_node(sxyz, ni, spole);
    if (topo & ORC_TOPO_NPOLE) pole_node(sxyz + 3 * (int64_t)(nj - 1) * ni, ni, npole);
    kdtree *t = NULL;
    double r2 = 0.0;
    if (!brute) {
        double m2 = 0.0;
        for (int32_t j = 0; j < nj - 1; ++j)
            for (int32_t i = 0; i < nqi; ++i) {
                int32_t i1 = (i + 1 == ni) ? 0 : i + 1;
                const double *q0 = sxyz + 3 * ((int64_t)j * ni + i), *q1 = sxyz + 3 * ((int64_t)j * ni + i1);
                double d1 = dist2(q0, q1 + 3 * (int64_t)ni), d2 = dist2(q1, q0 + 3 * (int64_t)ni);
                if (d1 > m2) m2 = d1;
                if (d2 > m2) m2 = d2;
            }
        double r = sqrt(m2) * 1.001 + 1e-9;
        r2 = r * r;
        t = kd_build(ni * nj, sxyz);
    }
    /* polar caps: a destination point within the cap's angular radius (+ margin) tries every cap triangle */
    double scap = 2.0, ncap = 2.0; /* cos of the cap radius; 2 = no cap */
    if (topo & ORC_TOPO_SPOLE) {
        scap = 1.0;
        for (int32_t i = 0; i < ni; ++i) { double d = dot3(spole, sxyz + 3 * (int64_t)i); if (d < scap) scap = d; }
        scap -= 1e-9;
    }
    if (topo & ORC_TOPO_NPOLE) {
        ncap = 1.0;
        for (int32_t i = 0; i < ni; ++i) { double d = dot3(npole, sxyz + 3 * ((int64_t)(nj - 1) * ni + i)); if (d < ncap) ncap = d; }
        ncap -= 1e-9;
    }
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t n = 0; n < nDst; ++n) {
        quad_ctx c;
        c.sxyz = sxyz; c.ni = ni; c.nj = nj; c.nqi = nqi; c.topo = topo; c.nquads = (int64_t)nqi * (nj - 1);
        c.spole = spole; c.npole = npole;
        c.p = dxyz + 3 * n; c.best = -1;
        for (int k = 0; k < 4; ++k) { c.w[k] = 0.0; c.col[k] = -1; }
        if (brute) {
            for (int32_t j = 0; j < nj - 1; ++j)
                for (int32_t i = 0; i < nqi; ++i) quad_try(i, j, &c);
        } else {
            kd_radius_rec(t, 0, ni * nj, c.p, r2, quad_visit_point, &c);
        }
        if (c.best < 0) { /* quads have the smaller ids: caps are tried only when no quad accepted the point */
            if (dot3(c.p, spole) >= scap) for (int32_t i = 0; i < ni; ++i) cap_try(i, 0, &c);
            if (c.best < 0 && dot3(c.p, npole) >= ncap) for (int32_t i = 0; i < ni; ++i) cap_try(i, 1, &c);
        }
        elem[n] = c.best;
        for (int k = 0; k < 4; ++k) { col[4 * n + k] = c.col[k]; w[4 * n + k] = c.best >= 0 ? c.w[k] : 0.0; }
    }
    if (t) kd_free(t);
    return 0;
}

int orc_bilinear_quadgrid(int32_t ni, int32_t nj, const double *sxyz, int64_t nDst, const double *dxyz,
                          int64_t *elem, int32_t *col, double *w, int brute) {
    return orc_bilinear_quadgrid_topo(ni, nj, sxyz, nDst, dxyz, elem, col, w, brute, 0);
}

/* Level-slowest source variant of the apply (source is itself a [lev][plane]
 * grid field: u/v_target_grid_nostag, interp.F90:307,325). */
void orc_apply_planes_f64(int64_t nDst, const int32_t *rowptr, const int32_t *col, const double *w, int32_t nlev,
                          int64_t srcPlane, const double *src, double *dst) {
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < nDst; ++t)
        for (int32_t l = 0; l < nlev; ++l) {
            double acc = 0.0;
            for (int32_t k = rowptr[t]; k < rowptr[t + 1]; ++k) acc = acc + w[k] * src[(int64_t)l * srcPlane + col[k]];
            dst[(int64_t)l * nDst + t] = acc;
        }
}

/* ------------------------------------------------------------------------ */
/* (8) CONSERVE, first order  (interp.F90:370-407; snow, snowh)              */
/* ------------------------------------------------------------------------ */
/* w_ij = area(S_j ^ D_i) / area(D_i)   (normType DSTAREA, no renormalisation
 * of partially covered destination cells).  S_j = Voronoi polygon of source
 * cell j (verticesOnCell order, model_grid.F90:446-486), D_i = quad of the 4
 * CORNER-stagger points around centre i (model_grid.F90:962-980).  Edges are
 * great circles (lineType GREAT_CIRCLE is ESMF's default for conservative).
 * Intersection: Sutherland-Hodgman clip of S_j by the 4 great-circle planes of
 * D_i; areas: fan of spherical triangles,
 *   E = 2 atan2(|a.((b-a)x(c-a))|, 1 + a.b + b.c + c.a). */
#define ORC_MAXPOLY 40

static double sph_tri_area(const double *a, const double *b, const double *c) {
    double ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    double ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    double n[3];
    cross3(ab, ac, n);
    double num = fabs(dot3(a, n));
    double den = ((1.0 + dot3(a, b)) + dot3(b, c)) + dot3(c, a);
    return 2.0 * atan2(num, den);
}

static double sph_poly_area(const double *v, int n) {
    double s = 0.0;
    for (int k = 1; k + 1 < n; ++k) s = s + sph_tri_area(v, v + 3 * k, v + 3 * (k + 1));
    return s;
}

/* make counter-clockwise seen from outside the sphere (reverse in place if not) */
static void orient_ccw(double *v, int n) {
    double s = 0.0, nrm[3], e1[3], e2[3];
    for (int k = 1; k + 1 < n; ++k) {
        for (int d = 0; d < 3; ++d) { e1[d] = v[3 * k + d] - v[d]; e2[d] = v[3 * (k + 1) + d] - v[d]; }
        cross3(e1, e2, nrm);
        s = s + dot3(v, nrm);
    }
    if (s < 0.0)
        for (int a = 0, b = n - 1; a < b; ++a, --b)
            for (int d = 0; d < 3; ++d) { double t = v[3 * a + d]; v[3 * a + d] = v[3 * b + d]; v[3 * b + d] = t; }
}

static int clip_by_plane(const double *in, int n, const double *nrm, double *out) {
    int m = 0;
    if (n == 0) return 0;
    const double *s = in + 3 * (n - 1);
    double ds = dot3(nrm, s);
    for (int k = 0; k < n; ++k) {
        const double *e = in + 3 * k;
        double de = dot3(nrm, e);
        if ((de >= 0.0) != (ds >= 0.0)) {
            double tau = ds / (ds - de);
            double x[3] = {s[0] + tau * (e[0] - s[0]), s[1] + tau * (e[1] - s[1]), s[2] + tau * (e[2] - s[2])};
            double inv = 1.0 / sqrt(dot3(x, x));
            if (m < ORC_MAXPOLY) { out[3 * m] = x[0] * inv; out[3 * m + 1] = x[1] * inv; out[3 * m + 2] = x[2] * inv; ++m; }
        }
        if (de >= 0.0 && m < ORC_MAXPOLY) { out[3 * m] = e[0]; out[3 * m + 1] = e[1]; out[3 * m + 2] = e[2]; ++m; }
        s = e;
        ds = de;
    }
    return m;
}

/* overlap area of source polygon sp (ns vertices, CCW) with destination quad dq (CCW) */
static double overlap_area(const double *sp, int ns, const double *dq) {
    double a[3 * ORC_MAXPOLY], b[3 * ORC_MAXPOLY];
    memcpy(a, sp, sizeof(double) * 3 * (size_t)ns);
    int n = ns;
    double *cur = a, *nxt = b;
    for (int e = 0; e < 4 && n > 0; ++e) {
        double nrm[3];
        cross3(dq + 3 * e, dq + 3 * ((e + 1) % 4), nrm);
        n = clip_by_plane(cur, n, nrm, nxt);
        double *t = cur; cur = nxt; nxt = t;
    }
    if (n < 3) return 0.0;
    return sph_poly_area(cur, n);
}

typedef struct {
    const double *vxyz;
    const int32_t *voc;
    int32_t maxEdges;
    const double *dq;
    double dstArea;
    int32_t *col;   /* NULL in the counting pass */
    double *w;
    int32_t n, cap;
} cons_ctx;

static void cons_visit_cell(int32_t cell, void *vctx) {
    cons_ctx *c = (cons_ctx *)vctx;
    double sp[3 * ORC_MAXPOLY];
    int ns = 0;
    for (int32_t k = 0; k < c->maxEdges && ns < ORC_MAXPOLY - 8; ++k) {
        int32_t v = c->voc[(int64_t)cell * c->maxEdges + k];
        if (v <= 0) continue;
        memcpy(sp + 3 * ns, c->vxyz + 3 * (int64_t)(v - 1), 3 * sizeof(double));
        ++ns;
    }
    if (ns < 3) return;
    orient_ccw(sp, ns);
    double ar = overlap_area(sp, ns, c->dq);
    if (!(ar > 0.0)) return;
    if (c->col) {
        /* insertion by ascending source id */
        int32_t pos = c->n;
        while (pos > 0 && c->col[pos - 1] > cell) { c->col[pos] = c->col[pos - 1]; c->w[pos] = c->w[pos - 1]; --pos; }
        c->col[pos] = cell;
        c->w[pos] = ar / c->dstArea;
    }
    c->n++;
}

/* Two-pass CSR: call with col == NULL to fill rowcount[nDst]; build rowptr; call
 * again with col/w sized rowptr[nDst].  corner_xyz: [(nj+1)][(ni+1)][3];
 * destination cell (i,j) (0-based, t = j*ni+i) uses corners (i,j),(i+1,j),
 * (i+1,j+1),(i,j+1).  cxyz = source cell centres (search only). */
int orc_conserve(int32_t nCells, const double *cxyz, const double *vxyz, int32_t maxEdges, const int32_t *voc,
                 int32_t ni, int32_t nj, const double *corner_xyz, int32_t *rowcount, const int32_t *rowptr,
                 int32_t *col, double *w, int brute) {
    kdtree *t = NULL;
    double rs = 0.0;
    if (!brute) {
        /* largest centre->vertex distance over all source cells */
        for (int32_t c = 0; c < nCells; ++c)
            for (int32_t k = 0; k < maxEdges; ++k) {
                int32_t v = voc[(int64_t)c * maxEdges + k];
                if (v <= 0) continue;
                double d = dist2(cxyz + 3 * (int64_t)c, vxyz + 3 * (int64_t)(v - 1));
                if (d > rs) rs = d;
            }
        rs = sqrt(rs);
        t = kd_build(nCells, cxyz);
    }
    int64_t nDst = (int64_t)ni * nj;
#pragma omp parallel for schedule(dynamic, 128)
    for (int64_t n = 0; n < nDst; ++n) {
        int32_t i = (int32_t)(n % ni), j = (int32_t)(n / ni);
        double dq[12];
        const double *c00 = corner_xyz + 3 * ((int64_t)j * (ni + 1) + i);
        memcpy(dq, c00, 24);
        memcpy(dq + 3, c00 + 3, 24);
        memcpy(dq + 6, c00 + 3 * ((int64_t)ni + 2), 24);
        memcpy(dq + 9, c00 + 3 * ((int64_t)ni + 1), 24);
        orient_ccw(dq, 4);
        cons_ctx c;
        c.vxyz = vxyz; c.voc = voc; c.maxEdges = maxEdges; c.dq = dq;
        c.dstArea = sph_poly_area(dq, 4);
        c.n = 0; c.cap = 0;
        c.col = col ? col + rowptr[n] : NULL;
        c.w = w ? w + rowptr[n] : NULL;
        if (c.dstArea > 0.0) {
            if (brute) {
                for (int32_t s = 0; s < nCells; ++s) cons_visit_cell(s, &c);
            } else {
                double cen[3] = {((dq[0] + dq[3]) + dq[6]) + dq[9], ((dq[1] + dq[4]) + dq[7]) + dq[10],
                                 ((dq[2] + dq[5]) + dq[8]) + dq[11]};
                double inv = 1.0 / sqrt(dot3(cen, cen));
                cen[0] *= inv; cen[1] *= inv; cen[2] *= inv;
                double rd = 0.0;
                for (int k = 0; k < 4; ++k) { double d = dist2(cen, dq + 3 * k); if (d > rd) rd = d; }
                double r = (sqrt(rd) + rs) * 1.001 + 1e-9;
                kd_radius_rec(t, 0, nCells, cen, r * r, cons_visit_cell, &c);
            }
        }
        if (rowcount) rowcount[n] = c.n;
    }
    if (t) kd_free(t);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* (9) BILINEAR, Mesh(node) -> Grid  (vorticity bundle, interp.F90:350-366)   */
/* ------------------------------------------------------------------------ */
/* Source values sit on the Voronoi vertices; source elements are the original
 * polygons.  ESMF splits polygons with more than 4 corners into triangles
 * before bilinear mapping; the exact split is an ESMF-internal choice that
 * cannot be checked here, so this restatement (and the engine) use a fan from
 * the polygon's first vertex in verticesOnCell order, for every polygon size.
 * Element id = cell id: smallest accepting cell wins, then its first accepting
 * fan triangle.  col [nDst][3] = vertex ids (0-based), w [nDst][3]. */
typedef struct {
    const double *vxyz;
    const int32_t *voc;
    int32_t maxEdges;
    const double *p;
    int32_t best;
    int32_t c[3];
    double w[3];
} node_ctx;

static void node_visit_cell(int32_t cell, void *vctx) {
    node_ctx *c = (node_ctx *)vctx;
    if (c->best >= 0 && cell >= c->best) return;
    int32_t v0 = -1, vp = -1;
    for (int32_t k = 0; k < c->maxEdges; ++k) {
        int32_t v = c->voc[(int64_t)cell * c->maxEdges + k];
        if (v <= 0) continue;
        v -= 1;
        if (v0 < 0) { v0 = v; continue; }
        if (vp < 0) { vp = v; continue; }
        double w[3];
        if (tri_locate(c->vxyz + 3 * (int64_t)v0, c->vxyz + 3 * (int64_t)vp, c->vxyz + 3 * (int64_t)v, c->p, w)) {
            c->best = cell;
            c->c[0] = v0; c->c[1] = vp; c->c[2] = v;
            c->w[0] = w[0]; c->w[1] = w[1]; c->w[2] = w[2];
            return;
        }
        vp = v;
    }
}

int orc_bilinear_node(int32_t nCells, const double *cxyz, const double *vxyz, int32_t maxEdges, const int32_t *voc,
                      int64_t nDst, const double *dxyz, int32_t *elem, int32_t *col, double *w, int brute) {
    kdtree *t = NULL;
    double r2 = 0.0;
    if (!brute) {
        double rs = 0.0;
        for (int32_t c = 0; c < nCells; ++c)
            for (int32_t k = 0; k < maxEdges; ++k) {
                int32_t v = voc[(int64_t)c * maxEdges + k];
                if (v <= 0) continue;
                double d = dist2(cxyz + 3 * (int64_t)c, vxyz + 3 * (int64_t)(v - 1));
                if (d > rs) rs = d;
            }
        double r = sqrt(rs) * 1.01 + 1e-9;
        r2 = r * r;
        t = kd_build(nCells, cxyz);
    }
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < nDst; ++i) {
        node_ctx c;
        c.vxyz = vxyz; c.voc = voc; c.maxEdges = maxEdges; c.p = dxyz + 3 * i; c.best = -1;
        c.c[0] = c.c[1] = c.c[2] = -1; c.w[0] = c.w[1] = c.w[2] = 0.0;
        if (brute) for (int32_t s = 0; s < nCells; ++s) node_visit_cell(s, &c);
        else kd_radius_rec(t, 0, nCells, c.p, r2, node_visit_cell, &c);
        elem[i] = c.best;
        for (int k = 0; k < 3; ++k) { col[3 * i + k] = c.c[k]; w[3 * i + k] = c.w[k]; }
    }
    if (t) kd_free(t);
    return 0;
}
