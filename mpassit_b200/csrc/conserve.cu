// conserve.cu -- weight generation for first-order CONSERVE (snow, snowh:
// interp.F90:368-407).  w_ij = area(S_j ^ D_i) / area(D_i) (normType DSTAREA):
// S_j = Voronoi polygon of source cell j, D_i = quad of the 4 CORNER-stagger
// points around destination centre i; edges are great circles.
//
// One thread per destination cell: BVH over source-polygon boxes -> candidates,
// Sutherland-Hodgman clip of each candidate by the destination quad's 4
// great-circle planes, spherical-triangle fan for areas.  Rows have variable
// length: one pass keeps up to 16 entries per cell in fixed-width rows (count,
// scan, compact); only when some cell overlaps more does a second, exact-size fill
// pass run.  Each row is emitted in ascending source id.  fp64, -fmad=false (which
// candidates overlap must be reproducible by a plain IEEE host restatement).
#include "bvh.cuh"
#include "common.cuh"

namespace mprg {

void scan_counts(mprg_ctx *ctx, const int32_t *cnt, int32_t *rowptr, int64_t nPlus1);  // locate.cu

constexpr int kEllCap = 16;  // entries per destination cell kept by the single-pass kernel

__device__ __forceinline__ double sph_tri_area(d3 a, d3 b, d3 c) {
    d3 n = cross(sub(b, a), sub(c, a));
    double num = fabs(dot(a, n));
    double den = ((1.0 + dot(a, b)) + dot(b, c)) + dot(c, a);
    return 2.0 * atan2(num, den);
}

__device__ double sph_poly_area(const d3 *v, int n) {
    double s = 0.0;
    for (int k = 1; k + 1 < n; ++k) s = s + sph_tri_area(v[0], v[k], v[k + 1]);
    return s;
}

__device__ void orient_ccw(d3 *v, int n) {
    double s = 0.0;
    for (int k = 1; k + 1 < n; ++k) s = s + dot(v[0], cross(sub(v[k], v[0]), sub(v[k + 1], v[0])));
    if (s < 0.0)
        for (int a = 0, b = n - 1; a < b; ++a, --b) { d3 t = v[a]; v[a] = v[b]; v[b] = t; }
}

template <int MAXP>
__device__ int clip_by_plane(const d3 *in, int n, d3 nrm, d3 *out) {
    int m = 0;
    if (n == 0) return 0;
    d3 s = in[n - 1];
    double ds = dot(nrm, s);
    for (int k = 0; k < n; ++k) {
        d3 e = in[k];
        double de = dot(nrm, e);
        if ((de >= 0.0) != (ds >= 0.0)) {
            double tau = ds / (ds - de);
            d3 x{s.x + tau * (e.x - s.x), s.y + tau * (e.y - s.y), s.z + tau * (e.z - s.z)};
            double inv = 1.0 / sqrt(dot(x, x));
            if (m < MAXP) { out[m] = d3{x.x * inv, x.y * inv, x.z * inv}; ++m; }
        }
        if (de >= 0.0 && m < MAXP) { out[m] = e; ++m; }
        s = e;
        ds = de;
    }
    return m;
}

// area of sp ^ dq; sp is clobbered (clipping ping-pongs between sp and tmp)
template <int MAXP>
__device__ double overlap_area(d3 *sp, int ns, const d3 *dq, d3 *tmp) {
    int n = ns;
    d3 *cur = sp, *nxt = tmp;
    for (int e = 0; e < 4 && n > 0; ++e) {
        d3 nrm = cross(dq[e], dq[(e + 1) & 3]);
        n = clip_by_plane<MAXP>(cur, n, nrm, nxt);
        d3 *t = cur; cur = nxt; nxt = t;
    }
    if (n < 3) return 0.0;
    return sph_poly_area(cur, n);
}

// MODE 0: count overlapping source cells per destination cell (first of two passes).
// MODE 1: write (col, w) sorted by ascending col into the CSR row (second pass).
// MODE 2: single pass -- count AND keep up to kEllCap sorted entries per cell in a fixed-width
//         row (rc/rw = ell + t * kEllCap); a cell with more raises *overflow and the host reruns
//         the two-pass form.
// MAXP: clip-buffer capacity; a polygon of maxEdges corners gains at most 4 by the 4 clips, so 16
// covers maxEdges <= 12 with 768 bytes of thread-local storage instead of 1920.
template <int MODE, int MAXP>
__global__ void __launch_bounds__(128)
k_conserve(BvhView bvh, int32_t maxEdges, const int32_t *__restrict__ voc, const double *__restrict__ vxyz,
           int32_t ni, int32_t j0, const double *__restrict__ cornerXyz, int64_t nDst, int32_t *__restrict__ cnt,
           const int32_t *__restrict__ rowptr, int32_t *__restrict__ col, double *__restrict__ w,
           int32_t *__restrict__ overflow) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDst) return;
    int32_t i = (int32_t)(t % ni), j = j0 + (int32_t)(t / ni);
    const double *c00 = cornerXyz + 3 * ((size_t)j * (ni + 1) + i);
    d3 dq[4] = {ld3(c00), ld3(c00 + 3), ld3(c00 + 3 * ((size_t)ni + 2)), ld3(c00 + 3 * ((size_t)ni + 1))};
    orient_ccw(dq, 4);
    const double dstArea = sph_poly_area(dq, 4);
    int n = 0;
    int32_t *rc = MODE == 1 ? col + rowptr[t] : (MODE == 2 ? col + t * kEllCap : nullptr);
    double *rw = MODE == 1 ? w + rowptr[t] : (MODE == 2 ? w + t * kEllCap : nullptr);
    if (dstArea > 0.0) {
        // query box: corners, grown by the great-circle bulge of the quad's edges
        double e2 = fmax(fmax(dist2(dq[0], dq[1]), dist2(dq[1], dq[2])), fmax(dist2(dq[2], dq[3]), dist2(dq[3], dq[0])));
        double m = (1.0 - sqrt(fmax(0.0, 1.0 - e2 / 4.0))) * 1.01 + 1e-12;
        d3 qlo{fmin(fmin(dq[0].x, dq[1].x), fmin(dq[2].x, dq[3].x)) - m, fmin(fmin(dq[0].y, dq[1].y), fmin(dq[2].y, dq[3].y)) - m,
               fmin(fmin(dq[0].z, dq[1].z), fmin(dq[2].z, dq[3].z)) - m};
        d3 qhi{fmax(fmax(dq[0].x, dq[1].x), fmax(dq[2].x, dq[3].x)) + m, fmax(fmax(dq[0].y, dq[1].y), fmax(dq[2].y, dq[3].y)) + m,
               fmax(fmax(dq[0].z, dq[1].z), fmax(dq[2].z, dq[3].z)) + m};
        bvh_overlap(bvh, qlo, qhi, [&](int s0, int s1) {
            for (int s = s0; s < s1; ++s) {
                // a polygon whose own (conservative) box misses the query box cannot overlap the cell
                if (!prim_hit(bvh, s, qlo, qhi)) continue;
                int32_t cell = __ldg(bvh.primId + s);
                d3 sp[MAXP], tmp[MAXP];
                int ns = 0;
                for (int k = 0; k < maxEdges && ns < MAXP - 4; ++k) {
                    int32_t v = __ldg(voc + (size_t)cell * maxEdges + k);
                    if (v <= 0) continue;
                    sp[ns++] = ld3(vxyz + 3 * (size_t)(v - 1));
                }
                if (ns < 3) continue;
                orient_ccw(sp, ns);
                double ar = overlap_area<MAXP>(sp, ns, dq, tmp);
                if (!(ar > 0.0)) continue;
                if (MODE == 1 || (MODE == 2 && n < kEllCap)) {
                    int pos = n;
                    while (pos > 0 && rc[pos - 1] > cell) { rc[pos] = rc[pos - 1]; rw[pos] = rw[pos - 1]; --pos; }
                    rc[pos] = cell;
                    rw[pos] = ar / dstArea;
                }
                ++n;
            }
        });
    }
    if (MODE != 1) {
        cnt[t] = n;
        if (t == 0) cnt[nDst] = 0;
        if (MODE == 2 && n > kEllCap) *overflow = 1;
    }
}

__global__ void k_compact_ell(int64_t nDst, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ ecol,
                              const double *__restrict__ ew, int32_t *__restrict__ col, double *__restrict__ w) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDst) return;
    int32_t b = rowptr[t], e = rowptr[t + 1];
    for (int k = 0; k < e - b; ++k) { col[b + k] = ecol[t * kEllCap + k]; w[b + k] = ew[t * kEllCap + k]; }
}

template <int MAXP>
static void conserve_launch(mprg_ctx *ctx, mprg_route *r, const BvhView &v, const Target &tg, const Target &cor) {
    Mesh &m = ctx->mesh;
    const int64_t n = tg.nSlab();
    const unsigned grid = (unsigned)((n + 127) / 128);
    DevBuf<int32_t> cnt(n + 1), flag(1);
    r->rowptr.alloc(n + 1);
    auto finish_sizes = [&]() {
        scan_counts(ctx, cnt.p, r->rowptr.p, n + 1);
        int32_t nnz = 0;
        peek(ctx, &nnz, r->rowptr.p + n, sizeof(int32_t));
        r->nnz = nnz;
        r->col.alloc(nnz > 0 ? nnz : 1);
        r->w.alloc(nnz > 0 ? nnz : 1);
    };
    {   // single pass into fixed-width rows
        DevBuf<int32_t> ecol((size_t)n * kEllCap);
        DevBuf<double> ew((size_t)n * kEllCap);
        MPRG_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int32_t), ctx->stream));
        k_conserve<2, MAXP><<<grid, 128, 0, ctx->stream>>>(v, m.maxEdges, m.voc.p, m.vertXyz.p, tg.ni, tg.j0, cor.x(), n,
                                                           cnt.p, nullptr, ecol.p, ew.p, flag.p);
        ctx->launches++;
        MPRG_CUDA(cudaGetLastError());
        finish_sizes();  // synchronises the stream
        int32_t over = 0;
        peek(ctx, &over, flag.p, sizeof(int32_t));
        if (!over) {
            k_compact_ell<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, r->rowptr.p, ecol.p, ew.p, r->col.p, r->w.p);
            ctx->launches++;
            MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
            return;
        }
    }
    // some destination cell overlaps more than kEllCap source cells (target much coarser than the mesh):
    // counts are already exact, fill the variable-length rows in a second pass
    k_conserve<1, MAXP><<<grid, 128, 0, ctx->stream>>>(v, m.maxEdges, m.voc.p, m.vertXyz.p, tg.ni, tg.j0, cor.x(), n,
                                                       nullptr, r->rowptr.p, r->col.p, r->w.p, nullptr);
    ctx->launches++;
    MPRG_CUDA(cudaGetLastError());
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
}

void store_conserve(mprg_ctx *ctx, mprg_route *r) {
    if (r->dst_stagger != MPRG_CENTER) fail(58, "mprg_store: CONSERVE is defined for the CENTER stagger only");
    Target &tg = ctx->target[MPRG_CENTER];
    Target &cor = ctx->target[MPRG_CORNER];
    if (!cor.set) fail(59, "mprg_store: CONSERVE needs the CORNER stagger (mprg_set_target(MPRG_CORNER, ...))");
    if (cor.ni != tg.ni + 1 || cor.nj != tg.nj + 1) fail(60, "mprg_store: CORNER stagger must be (ni+1) x (nj+1)");
    mesh_need_poly_bvh(ctx);
    Mesh &m = ctx->mesh;
    if (m.maxEdges > 36) fail(64, "mprg_store: maxEdges %d exceeds the clipping buffer", m.maxEdges);
    if (route_empty_slab(ctx, r, tg.nSlab(), m.nCells)) return;
    r->nDst = tg.nSlab();
    r->nSrc = m.nCells;
    BvhView v{m.polyBvh.nodes.p, m.polyBvh.primId.p, m.polyBvh.nLeafNodes, m.polyBvh.nPrim, m.polyBvh.primBox.p};
    if (m.maxEdges <= 12) conserve_launch<16>(ctx, r, v, tg, cor);
    else conserve_launch<40>(ctx, r, v, tg, cor);
}

}  // namespace mprg
