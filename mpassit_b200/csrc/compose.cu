// compose.cu -- the wind chain of interp_hist_data as ONE weight matrix per staggered grid.
//
// The reference regrids the cell-centre winds to the mass points (ESMF_FieldRegrid, interp.F90:256-289), rotates
// them there (rotate_winds_cgrid, :291-293, :689-749) and regrids the rotated mass-point fields to the EDGE1 /
// EDGE2 points (:295-328).  All three steps are linear in (u, v):
//     u'(c) = (u(c) + v(c) tana) / (cosa + sina tana)                    = rden (u + tana v)
//     v'(c) = (v(c) - u'(c) sina) / cosa                                = -rca sina rden u + rca (1 - sina rden tana) v
//     U(t)  = sum_q S_U[t][q] u'(c_q),   V(t) = sum_q S_V[t][q] v'(c_q),   u(c) = sum_k W[c][k] u_src[col_k]
// so  U = A_U u_src + B_U v_src  and  V = A_V u_src + B_V v_src  with, per destination row, at most 4 x 3 entries on
// the mesh cells around the point.  The composed route holds (col, a, b) per entry; the column kernel stages the
// zonal and the meridional source columns of a tile one after the other and writes the staggered field directly:
// the mass-point intermediates (UMASS / VMASS: one write and one read of two 3-D fields) never exist, and the
// grid-source apply (apply_planes.cuh) drops out of the pass.  Differences to the three-step chain are rounding
// only: the chain rounds the mass-point values to the output type, this path does not.
//
// Composable when the target grid is regional (no periodic seam, no pole rows), the rotation angles are set and
// no composed row exceeds kCmpRow distinct cells; otherwise the caller keeps the chain.
#include "common.cuh"

namespace mprg {

void scan_counts(mprg_ctx *ctx, const int32_t *cnt, int32_t *rowptr, int64_t nPlus1);  // locate.cu

// kind 0: destination carries the rotated zonal wind (EDGE1), 1: the rotated meridional wind (EDGE2)
__global__ void __launch_bounds__(128)
k_compose_wind(int64_t nDst, const int32_t *__restrict__ srp, const int32_t *__restrict__ scol, const double *__restrict__ sw,
               const int32_t *__restrict__ wrp, const int32_t *__restrict__ wcol, const double *__restrict__ ww,
               const double *__restrict__ rotc, int kind, int32_t *__restrict__ ecol, double *__restrict__ ea,
               double *__restrict__ eb, int32_t *__restrict__ cnt, int32_t *__restrict__ overflow) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDst) return;
    int32_t c[kCmpRow];
    double a[kCmpRow], b[kCmpRow];
    int n = 0;
    bool over = false;
    for (int e = srp[t]; e < srp[t + 1]; ++e) {
        const int32_t q = scol[e];
        const double ws = sw[e];
        const double sa = rotc[4 * (int64_t)q], tana = rotc[4 * (int64_t)q + 1], rca = rotc[4 * (int64_t)q + 2], rden = rotc[4 * (int64_t)q + 3];
        double fa, fb;
        if (kind == 0) {
            fa = ws * rden;
            fb = ws * (rden * tana);
        } else {
            fa = -(ws * (rca * (sa * rden)));
            fb = ws * (rca * (1.0 - sa * (rden * tana)));
        }
        for (int k = wrp[q]; k < wrp[q + 1]; ++k) {
            const int32_t cc = wcol[k];
            const double w = ww[k];
            int pos = -1;
            for (int m = 0; m < n; ++m)
                if (c[m] == cc) pos = m;
            if (pos < 0) {
                if (n == kCmpRow) { over = true; continue; }
                pos = n++;
                c[pos] = cc; a[pos] = 0.0; b[pos] = 0.0;
            }
            a[pos] += fa * w;
            b[pos] += fb * w;
        }
    }
    // ascending cell id (the summation order of the apply)
    for (int i = 1; i < n; ++i) {
        const int32_t ci = c[i];
        const double ai = a[i], bi = b[i];
        int j = i - 1;
        while (j >= 0 && c[j] > ci) { c[j + 1] = c[j]; a[j + 1] = a[j]; b[j + 1] = b[j]; --j; }
        c[j + 1] = ci; a[j + 1] = ai; b[j + 1] = bi;
    }
    for (int m = 0; m < n; ++m) {
        ecol[kCmpRow * t + m] = c[m];
        ea[kCmpRow * t + m] = a[m];
        eb[kCmpRow * t + m] = b[m];
    }
    cnt[t] = n;
    if (over) atomicOr(overflow, 1);
}

__global__ void k_compact_cmp(int64_t nDst, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ ecol,
                              const double *__restrict__ ea, const double *__restrict__ eb, int32_t *__restrict__ col,
                              double *__restrict__ w, double *__restrict__ w2) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDst) return;
    const int b = rowptr[t], n = rowptr[t + 1] - b;
    for (int m = 0; m < n; ++m) {
        col[b + m] = ecol[kCmpRow * t + m];
        w[b + m] = ea[kCmpRow * t + m];
        w2[b + m] = eb[kCmpRow * t + m];
    }
}

bool store_wind_composed(mprg_ctx *ctx, mprg_route *r, const mprg_route *stag, const mprg_route *bil) {
    if (!ctx->haveRot || ctx->gridKind != MPRG_GRID_NOPERI) return false;
    if (!stag->srcLevelSlowest || bil->srcLevelSlowest || stag->nSrc != bil->nDst) return false;
    if (stag->maxRow > kLongRow || bil->maxRow > 3) return false;
    const int64_t n = stag->nDst;
    r->nDst = n;
    r->nSrc = bil->nSrc;
    r->composite = true;
    r->srcLevelSlowest = false;
    if (route_empty_slab(ctx, r, n, bil->nSrc)) { r->w2.alloc(1); return true; }
    const double *rotc = ctx->rotc.p + 4 * ctx->target[MPRG_CENTER_HALO].slabOffset();
    DevBuf<int32_t> ecol((size_t)kCmpRow * n), cnt(n + 1), over(1);
    DevBuf<double> ea((size_t)kCmpRow * n), eb((size_t)kCmpRow * n);
    MPRG_CUDA(cudaMemsetAsync(over.p, 0, sizeof(int32_t), ctx->stream));
    k_compose_wind<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(n, stag->rowptr.p, stag->col.p, stag->w.p, bil->rowptr.p,
                                                                         bil->col.p, bil->w.p, rotc, r->dst_stagger == MPRG_EDGE2 ? 1 : 0,
                                                                         ecol.p, ea.p, eb.p, cnt.p, over.p);
    ctx->launches++;
    MPRG_CUDA(cudaGetLastError());
    int32_t hover = 0;
    peek(ctx, &hover, over.p, sizeof hover);
    if (hover) return false;   // some point draws on more than kCmpRow cells (a mesh much finer than the grid)
    r->rowptr.alloc(n + 1);
    scan_counts(ctx, cnt.p, r->rowptr.p, n + 1);
    int32_t nnz = 0;
    peek(ctx, &nnz, r->rowptr.p + n, sizeof(int32_t));
    r->nnz = nnz;
    r->col.alloc(nnz > 0 ? nnz : 1);
    r->w.alloc(nnz > 0 ? nnz : 1);
    r->w2.alloc(nnz > 0 ? nnz : 1);
    k_compact_cmp<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, r->rowptr.p, ecol.p, ea.p, eb.p, r->col.p, r->w.p, r->w2.p);
    ctx->launches++;
    MPRG_CUDA(cudaGetLastError());
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
    return true;
}

}  // namespace mprg
