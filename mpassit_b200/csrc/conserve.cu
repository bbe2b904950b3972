// conserve.cu -- weight generation for first-order CONSERVE (snow, snowh:
// interp.F90:368-407).  w_ij = area(S_j ^ D_i) / area(D_i) (normType DSTAREA):
// S_j = Voronoi polygon of source cell j, D_i = quad of the 4 CORNER-stagger
// points around destination centre i; edges are great circles.
//
// One thread per destination cell: BVH over source-polygon boxes -> candidates,
// Sutherland-Hodgman clip of each candidate by the destination quad's 4
// great-circle planes, spherical-triangle fan for areas.  Rows have variable
// length, so it runs twice (count, scan, fill) and each row is emitted in
// ascending source id.  fp64, -fmad=false (which candidates overlap must be
// reproducible by a plain IEEE host restatement).
#include "bvh.cuh"
#include "common.cuh"

namespace mprg {

void scan_counts(mprg_ctx *ctx, const int32_t *cnt, int32_t *rowptr, int64_t nPlus1);  // locate.cu

constexpr int kMaxPoly = 40;

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
            if (m < kMaxPoly) { out[m] = d3{x.x * inv, x.y * inv, x.z * inv}; ++m; }
        }
        if (de >= 0.0 && m < kMaxPoly) { out[m] = e; ++m; }
        s = e;
        ds = de;
    }
    return m;
}

__device__ double overlap_area(const d3 *sp, int ns, const d3 *dq) {
    d3 a[kMaxPoly], b[kMaxPoly];
    for (int k = 0; k < ns; ++k) a[k] = sp[k];
    int n = ns;
    d3 *cur = a, *nxt = b;
    for (int e = 0; e < 4 && n > 0; ++e) {
        d3 nrm = cross(dq[e], dq[(e + 1) & 3]);
        n = clip_by_plane(cur, n, nrm, nxt);
        d3 *t = cur; cur = nxt; nxt = t;
    }
    if (n < 3) return 0.0;
    return sph_poly_area(cur, n);
}

// FILL == false: count overlapping source cells per destination cell.
// FILL == true : write (col, w) sorted by ascending col into the CSR row.
template <bool FILL>
__global__ void __launch_bounds__(128)
k_conserve(BvhView bvh, int32_t maxEdges, const int32_t *__restrict__ voc, const double *__restrict__ vxyz,
           int32_t ni, int32_t j0, const double *__restrict__ cornerXyz, int64_t nDst, int32_t *__restrict__ cnt,
           const int32_t *__restrict__ rowptr, int32_t *__restrict__ col, double *__restrict__ w) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDst) return;
    int32_t i = (int32_t)(t % ni), j = j0 + (int32_t)(t / ni);
    const double *c00 = cornerXyz + 3 * ((size_t)j * (ni + 1) + i);
    d3 dq[4] = {ld3(c00), ld3(c00 + 3), ld3(c00 + 3 * ((size_t)ni + 2)), ld3(c00 + 3 * ((size_t)ni + 1))};
    orient_ccw(dq, 4);
    const double dstArea = sph_poly_area(dq, 4);
    int n = 0;
    int32_t *rc = FILL ? col + rowptr[t] : nullptr;
    double *rw = FILL ? w + rowptr[t] : nullptr;
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
                int32_t cell = __ldg(bvh.primId + s);
                d3 sp[kMaxPoly];
                int ns = 0;
                for (int k = 0; k < maxEdges && ns < kMaxPoly - 8; ++k) {
                    int32_t v = __ldg(voc + (size_t)cell * maxEdges + k);
                    if (v <= 0) continue;
                    sp[ns++] = ld3(vxyz + 3 * (size_t)(v - 1));
                }
                if (ns < 3) continue;
                orient_ccw(sp, ns);
                double ar = overlap_area(sp, ns, dq);
                if (!(ar > 0.0)) continue;
                if (FILL) {
                    int pos = n;
                    while (pos > 0 && rc[pos - 1] > cell) { rc[pos] = rc[pos - 1]; rw[pos] = rw[pos - 1]; --pos; }
                    rc[pos] = cell;
                    rw[pos] = ar / dstArea;
                }
                ++n;
            }
        });
    }
    if (!FILL) {
        cnt[t] = n;
        if (t == 0) cnt[nDst] = 0;
    }
}

void store_conserve(mprg_ctx *ctx, mprg_route *r) {
    if (r->dst_stagger != MPRG_CENTER) fail(58, "mprg_store: CONSERVE is defined for the CENTER stagger only");
    Target &tg = ctx->target[MPRG_CENTER];
    Target &cor = ctx->target[MPRG_CORNER];
    if (!cor.set) fail(59, "mprg_store: CONSERVE needs the CORNER stagger (mprg_set_target(MPRG_CORNER, ...))");
    if (cor.ni != tg.ni + 1 || cor.nj != tg.nj + 1) fail(60, "mprg_store: CORNER stagger must be (ni+1) x (nj+1)");
    mesh_need_poly_bvh(ctx);
    Mesh &m = ctx->mesh;
    if (m.maxEdges > kMaxPoly - 8) fail(64, "mprg_store: maxEdges %d exceeds the clipping buffer", m.maxEdges);
    int64_t n = tg.nSlab();
    r->nDst = n;
    r->nSrc = m.nCells;
    DevBuf<int32_t> cnt(n + 1);
    BvhView v{m.polyBvh.nodes.p, m.polyBvh.primId.p, m.polyBvh.nLeafNodes, m.polyBvh.nPrim};
    const unsigned grid = (unsigned)((n + 127) / 128);
    k_conserve<false><<<grid, 128, 0, ctx->stream>>>(v, m.maxEdges, m.voc.p, m.vertXyz.p, tg.ni, tg.j0, cor.xyz.p, n,
                                                      cnt.p, nullptr, nullptr, nullptr);
    ctx->launches++;
    MPRG_CUDA(cudaGetLastError());
    r->rowptr.alloc(n + 1);
    scan_counts(ctx, cnt.p, r->rowptr.p, n + 1);
    int32_t nnz = 0;
    MPRG_CUDA(cudaMemcpy(&nnz, r->rowptr.p + n, sizeof(int32_t), cudaMemcpyDeviceToHost));
    r->nnz = nnz;
    r->col.alloc(nnz > 0 ? nnz : 1);
    r->w.alloc(nnz > 0 ? nnz : 1);
    k_conserve<true><<<grid, 128, 0, ctx->stream>>>(v, m.maxEdges, m.voc.p, m.vertXyz.p, tg.ni, tg.j0, cor.xyz.p, n,
                                                     nullptr, r->rowptr.p, r->col.p, r->w.p);
    ctx->launches++;
    MPRG_CUDA(cudaGetLastError());
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
}

}  // namespace mprg
