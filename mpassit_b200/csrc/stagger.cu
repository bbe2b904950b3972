// stagger.cu -- weight generation for BILINEAR with a structured source:
//   * Grid(CENTER) -> Grid(EDGE1/EDGE2)  (u/v_target_grid_nostag -> U/V,
//     interp.F90:298,316): source elements are the quads of 4 adjacent centres.
//   * Mesh(NODE) -> Grid(CENTER)        (vorticity bundle, interp.F90:353):
//     source elements are the Voronoi polygons, fan-triangulated.
// Candidates come from a BVH over the source elements; the quad mapping is the
// Newton iteration ESMF uses (start 0, |F|^2 < 1e-20, <= 100 steps).
#include "bvh.cuh"
#include "common.cuh"

namespace mprg {

void scan_counts(mprg_ctx *ctx, const int32_t *cnt, int32_t *rowptr, int64_t nPlus1);  // locate.cu

__device__ __forceinline__ bool invert3(const double *J, double *inv) {
    double c00 = J[4] * J[8] - J[5] * J[7];
    double c01 = J[5] * J[6] - J[3] * J[8];
    double c02 = J[3] * J[7] - J[4] * J[6];
    double det = (J[0] * c00 + J[1] * c01) + J[2] * c02;
    if (det == 0.0) return false;
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
    return true;
}

__device__ bool quad_locate(const double *q0, const double *q1, const double *q2, const double *q3,
                            const double *p, double tol, double *w) {
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
        if (!invert3(J, inv)) return false;
        for (int r = 0; r < 3; ++r)
            X[r] = X[r] - ((inv[3 * r] * F[0] + inv[3 * r + 1] * F[1]) + inv[3 * r + 2] * F[2]);
    }
    const double u = X[0], v = X[1], t = X[2];
    if (!(t > 0.0)) return false;
    if (!(u >= -tol && u <= 1.0 + tol && v >= -tol && v <= 1.0 + tol)) return false;
    w[0] = (1.0 - u) * (1.0 - v);
    w[1] = u * (1.0 - v);
    w[2] = u * v;
    w[3] = (1.0 - u) * v;
    return true;
}

// Topology of the source grid (mprg_set_grid_kind; model_grid.F90:684-703):
//   nqi      quads per source row: ni - 1 (ESMF_GridCreateNoPeriDim) or ni (ESMF_GridCreate1PeriDim: the last quad
//            joins column ni-1 to column 0 across the seam)
//   caps     bit 0 / bit 1: this row block holds grid row 0 / the last grid row of a MONOPOLE grid, closed by the fan
//            of triangles (pole, P_i, P_i+1) around ESMF's artificial pole node (polemethod ALLAVG: the centre of
//            the row projected onto the sphere, value = average of the row)
// Element ids (tie rule: smallest wins): quads j * nqi + i, then the south cap's triangles, then the north cap's.
struct QuadTopo {
    int32_t ni, nrows, nqi, caps;
    const double *pole;  // [8]: south pole xyz, cos(cap radius), north pole xyz, cos(cap radius)
};

__global__ void k_quad_boxes(QuadTopo tp, const double *__restrict__ sxyz, float *lo, float *hi) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t nq = (int64_t)tp.nqi * (tp.nrows - 1);
    if (q >= nq) return;
    const int32_t ni = tp.ni;
    int32_t i = (int32_t)(q % tp.nqi), j = (int32_t)(q / tp.nqi);
    int32_t i1 = (i + 1 == ni) ? 0 : i + 1;
    const double *p0 = sxyz + 3 * ((size_t)j * ni + i), *p1 = sxyz + 3 * ((size_t)j * ni + i1);
    d3 a = ld3(p0), b = ld3(p1), c = ld3(p1 + 3 * (size_t)ni), d = ld3(p0 + 3 * (size_t)ni);
    double e2 = fmax(fmax(fmax(dist2(a, b), dist2(b, c)), fmax(dist2(c, d), dist2(d, a))), fmax(dist2(a, c), dist2(b, d)));
    // points of the bilinear patch lie in the hull of the 4 corners: |x|^2 >= 1 - 3E^2/8
    double m = (1.0 - sqrt(fmax(0.0, 1.0 - 0.375 * e2))) * 1.01 + sqrt(e2) * 1e-9 + 1e-12;
    float *l = lo + 3 * q, *h = hi + 3 * q;
    l[0] = f_down(fmin(fmin(a.x, b.x), fmin(c.x, d.x)) - m);
    l[1] = f_down(fmin(fmin(a.y, b.y), fmin(c.y, d.y)) - m);
    l[2] = f_down(fmin(fmin(a.z, b.z), fmin(c.z, d.z)) - m);
    h[0] = f_up(fmax(fmax(a.x, b.x), fmax(c.x, d.x)) + m);
    h[1] = f_up(fmax(fmax(a.y, b.y), fmax(c.y, d.y)) + m);
    h[2] = f_up(fmax(fmax(a.z, b.z), fmax(c.z, d.z)) + m);
}

// artificial pole nodes of a monopole grid: sequential sums in column order (the oracle's order), one thread
__global__ void k_pole_nodes(QuadTopo tp, const double *__restrict__ sxyz, double *__restrict__ pole) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int side = 0; side < 2; ++side) {
        double *o = pole + 4 * side;
        o[0] = 0.0; o[1] = 0.0; o[2] = side ? 1.0 : -1.0; o[3] = 2.0;  // cos(radius) = 2: no cap
        if (!(tp.caps & (1 << side))) continue;
        const double *row = sxyz + 3 * (size_t)(side ? tp.nrows - 1 : 0) * tp.ni;
        double sx = 0.0, sy = 0.0, sz = 0.0;
        for (int32_t i = 0; i < tp.ni; ++i) { sx = sx + row[3 * (size_t)i]; sy = sy + row[3 * (size_t)i + 1]; sz = sz + row[3 * (size_t)i + 2]; }
        const double inv = 1.0 / sqrt((sx * sx + sy * sy) + sz * sz);
        const d3 pn{sx * inv, sy * inv, sz * inv};
        double c = 1.0;
        for (int32_t i = 0; i < tp.ni; ++i) { const double d = dot(pn, ld3(row + 3 * (size_t)i)); if (d < c) c = d; }
        o[0] = pn.x; o[1] = pn.y; o[2] = pn.z; o[3] = c - 1e-9;
    }
}

// Per destination point: ecol / ew, 4 wide.  Quad: columns (q0,q1,q2,q3), bilinear weights, cnt 4.  Cap triangle:
// columns (P_i, P_i+1, first point of the pole row, -1), weights (a, b, ws, 0), cnt ni (expanded by k_compact4).
__global__ void __launch_bounds__(128)
k_bilinear_quad(BvhView bvh, QuadTopo tp, const double *__restrict__ sxyz, const double *__restrict__ dstXyz,
                int64_t nDst, int32_t *__restrict__ ecol, double *__restrict__ ew, int32_t *__restrict__ cnt) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDst) return;
    const int32_t ni = tp.ni;
    d3 p = ld3(dstXyz + 3 * t);
    double pp[3] = {p.x, p.y, p.z};
    int32_t best = -1;
    double bw[4] = {0.0, 0.0, 0.0, 0.0};
    bvh_overlap(bvh, p, p, [&](int s0, int s1) {
        for (int s = s0; s < s1; ++s) {
            int32_t q = __ldg(bvh.primId + s);
            if (best >= 0 && q >= best) continue;  // smallest source element id wins
            int32_t i = q % tp.nqi, j = q / tp.nqi;
            int32_t i1 = (i + 1 == ni) ? 0 : i + 1;
            const double *q0 = sxyz + 3 * ((size_t)j * ni + i), *q1 = sxyz + 3 * ((size_t)j * ni + i1);
            double w[4];
            if (quad_locate(q0, q1, q1 + 3 * (size_t)ni, q0 + 3 * (size_t)ni, pp, kTol, w)) {
                best = q;
                bw[0] = w[0]; bw[1] = w[1]; bw[2] = w[2]; bw[3] = w[3];
            }
        }
    });
    if (best >= 0) {
        int32_t i = best % tp.nqi, j = best / tp.nqi;
        int32_t i1 = (i + 1 == ni) ? 0 : i + 1;
        int32_t b0 = j * ni + i, b1 = j * ni + i1;
        ecol[4 * t + 0] = b0; ecol[4 * t + 1] = b1; ecol[4 * t + 2] = b1 + ni; ecol[4 * t + 3] = b0 + ni;
        for (int k = 0; k < 4; ++k) ew[4 * t + k] = bw[k];
        cnt[t] = 4;
    } else {
        // no quad took the point: the polar caps (their ids follow the quads'), first accepting triangle wins
        int32_t cc[3] = {-1, -1, -1};
        double cw[3] = {0.0, 0.0, 0.0};
        bool got = false;
        for (int side = 0; side < 2 && !got; ++side) {
            if (!(tp.caps & (1 << side))) continue;
            const d3 pn = ld3(tp.pole + 4 * side);
            if (!(dot(p, pn) >= tp.pole[4 * side + 3])) continue;
            const int32_t row = side ? tp.nrows - 1 : 0;
            for (int32_t i = 0; i < ni && !got; ++i) {
                const int32_t i1 = (i + 1 == ni) ? 0 : i + 1;
                const int32_t a = row * ni + i, b = row * ni + i1;
                double w[3];
                if (tri_locate(pn, ld3(sxyz + 3 * (size_t)a), ld3(sxyz + 3 * (size_t)b), p, kTol, w)) {
                    got = true;
                    cc[0] = a; cc[1] = b; cc[2] = row * ni;
                    cw[0] = w[1]; cw[1] = w[2]; cw[2] = w[0];
                }
            }
        }
        if (got) {
            for (int k = 0; k < 3; ++k) { ecol[4 * t + k] = cc[k]; ew[4 * t + k] = cw[k]; }
            ecol[4 * t + 3] = -1; ew[4 * t + 3] = 0.0;
            cnt[t] = ni;
        } else {
            cnt[t] = 0;
        }
    }
    if (t == 0) cnt[nDst] = 0;
}

__global__ void k_compact4(int64_t nDst, int32_t ni, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ ecol,
                           const double *__restrict__ ew, int32_t *__restrict__ col, double *__restrict__ w) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDst) return;
    int32_t b = rowptr[t], e = rowptr[t + 1];
    if (e - b == 4 && ecol[4 * t + 3] >= 0) {
        for (int k = 0; k < 4; ++k) { col[b + k] = ecol[4 * t + k]; w[b + k] = ew[4 * t + k]; }
    } else if (e > b) {
        // cap row: the pole's value is the average of its row -> ws / ni on every point of the row, plus the
        // triangle's weights on its two row points; ascending column order
        const int32_t ca = ecol[4 * t], cb = ecol[4 * t + 1], base = ecol[4 * t + 2];
        const double share = ew[4 * t + 2] / (double)ni;
        for (int32_t i = 0; i < ni; ++i) {
            double v = share;
            if (base + i == ca) v = v + ew[4 * t];
            if (base + i == cb) v = v + ew[4 * t + 1];
            col[b + i] = base + i;
            w[b + i] = v;
        }
    }
}

void store_bilinear_grid(mprg_ctx *ctx, mprg_route *r) {
    // Source = the CENTER rows this rank holds mass-point values for (MPRG_CENTER_HALO: its own slab
    // +- one row; the whole grid on one rank).  Quads, their ids (tie rule: smallest id wins) and the
    // emitted columns are all relative to that row block, which preserves the global order.
    Target &src = ctx->target[MPRG_CENTER_HALO];
    Target &tg = ctx->target[r->dst_stagger];
    if (!src.set) fail(55, "mprg_store: GRID_CENTER source needs the CENTER stagger");
    if (r->dst_stagger != MPRG_EDGE1 && r->dst_stagger != MPRG_EDGE2)
        fail(56, "mprg_store: GRID_CENTER source feeds the EDGE1 / EDGE2 staggers only");
    const int32_t srows = src.j1 - src.j0;
    int64_t n = tg.nSlab();
    r->nDst = n;
    r->nSrc = (int64_t)src.ni * srows;
    r->srcLevelSlowest = true;
    r->srcPlane = r->nSrc;
    if (n == 0 || src.ni < 2 || srows < 2) {
        if (ctx->nranks == 1) fail(57, "mprg_store: CENTER grid too small for quads");
        // a rank with fewer than two source rows owns no mappable edge point: all rows empty
        r->nnz = 0;
        r->rowptr.alloc(n + 1); r->col.alloc(1); r->w.alloc(1);
        MPRG_CUDA(cudaMemsetAsync(r->rowptr.p, 0, (n + 1) * sizeof(int32_t), ctx->stream));
        MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
        return;
    }
    const double *sxyz = src.x() + 3 * src.slabOffset();
    const bool peri = ctx->gridKind == MPRG_GRID_1PERI_MONOPOLE;
    DevBuf<double> pole(8);
    QuadTopo tp;
    tp.ni = src.ni; tp.nrows = srows; tp.nqi = peri ? src.ni : src.ni - 1;
    tp.caps = peri ? ((src.j0 == 0 ? 1 : 0) | (src.j1 == src.nj ? 2 : 0)) : 0;
    tp.pole = pole.p;
    k_pole_nodes<<<1, 32, 0, ctx->stream>>>(tp, sxyz, pole.p);
    ctx->launches++;
    int64_t nq = (int64_t)tp.nqi * (srows - 1);
    DevBuf<float> lo(3 * nq), hi(3 * nq);
    k_quad_boxes<<<(unsigned)((nq + 255) / 256), 256, 0, ctx->stream>>>(tp, sxyz, lo.p, hi.p);
    ctx->launches++;
    Bvh bvh;
    bvh_build_boxes(ctx, lo.p, hi.p, (int32_t)nq, bvh);
    DevBuf<int32_t> ecol(4 * n), cnt(n + 1);
    DevBuf<double> ew(4 * n);
    BvhView v{bvh.nodes.p, bvh.primId.p, bvh.nLeafNodes, bvh.nPrim};
    k_bilinear_quad<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(v, tp, sxyz, tg.x() + 3 * tg.slabOffset(), n,
                                                                          ecol.p, ew.p, cnt.p);
    ctx->launches++;
    MPRG_CUDA(cudaGetLastError());
    r->rowptr.alloc(n + 1);
    scan_counts(ctx, cnt.p, r->rowptr.p, n + 1);
    int32_t nnz = 0;
    peek(ctx, &nnz, r->rowptr.p + n, sizeof(int32_t));
    r->nnz = nnz;
    r->col.alloc(nnz > 0 ? nnz : 1);
    r->w.alloc(nnz > 0 ? nnz : 1);
    k_compact4<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, src.ni, r->rowptr.p, ecol.p, ew.p, r->col.p, r->w.p);
    ctx->launches++;
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
}

// ---------------------------------------------------------------------------
// BILINEAR from mesh NODES: point in Voronoi polygon, fan-triangulated from the
// polygon's first vertex; weights on the (up to 3) fan-triangle corners.
// Element id = cell id; smallest accepting cell wins, then the first fan triangle.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_bilinear_node(BvhView bvh, int32_t maxEdges, const int32_t *__restrict__ voc, const double *__restrict__ vxyz,
                const double *__restrict__ dstXyz, int64_t nDst, int32_t *__restrict__ ecol,
                double *__restrict__ ew, int32_t *__restrict__ cnt) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDst) return;
    d3 p = ld3(dstXyz + 3 * t);
    int32_t best = -1, bc[3] = {0, 0, 0};
    double bw[3] = {0.0, 0.0, 0.0};
    bvh_overlap(bvh, p, p, [&](int s0, int s1) {
        for (int s = s0; s < s1; ++s) {
            int32_t c = __ldg(bvh.primId + s);
            if (best >= 0 && c >= best) continue;
            const int32_t *vc = voc + (size_t)c * maxEdges;
            int32_t v0 = -1, vp = -1;
            for (int k = 0; k < maxEdges; ++k) {
                int32_t v = __ldg(vc + k);
                if (v <= 0) continue;
                v -= 1;
                if (v0 < 0) { v0 = v; continue; }
                if (vp < 0) { vp = v; continue; }
                double w[3];
                if (tri_locate(ld3(vxyz + 3 * (size_t)v0), ld3(vxyz + 3 * (size_t)vp), ld3(vxyz + 3 * (size_t)v), p,
                               kTol, w)) {
                    best = c;
                    bc[0] = v0; bc[1] = vp; bc[2] = v;
                    bw[0] = w[0]; bw[1] = w[1]; bw[2] = w[2];
                    break;
                }
                vp = v;
            }
        }
    });
    if (best >= 0) {
        for (int k = 0; k < 3; ++k) { ecol[3 * t + k] = bc[k]; ew[3 * t + k] = bw[k]; }
        cnt[t] = 3;
    } else {
        cnt[t] = 0;
    }
    if (t == 0) cnt[nDst] = 0;
}

__global__ void k_compact3n(int64_t nDst, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ ecol,
                            const double *__restrict__ ew, int32_t *__restrict__ col, double *__restrict__ w) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDst) return;
    int32_t b = rowptr[t], e = rowptr[t + 1];
    for (int k = 0; k < e - b; ++k) { col[b + k] = ecol[3 * t + k]; w[b + k] = ew[3 * t + k]; }
}

void store_bilinear_node(mprg_ctx *ctx, mprg_route *r) {
    mesh_need_poly_bvh(ctx);
    Mesh &m = ctx->mesh;
    Target &tg = ctx->target[r->dst_stagger];
    int64_t n = tg.nSlab();
    if (route_empty_slab(ctx, r, n, m.nVertices)) return;
    r->nDst = n;
    r->nSrc = m.nVertices;
    DevBuf<int32_t> ecol(3 * n), cnt(n + 1);
    DevBuf<double> ew(3 * n);
    BvhView v{m.polyBvh.nodes.p, m.polyBvh.primId.p, m.polyBvh.nLeafNodes, m.polyBvh.nPrim};
    k_bilinear_node<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(
        v, m.maxEdges, m.voc.p, m.vertXyz.p, tg.x() + 3 * tg.slabOffset(), n, ecol.p, ew.p, cnt.p);
    ctx->launches++;
    MPRG_CUDA(cudaGetLastError());
    r->rowptr.alloc(n + 1);
    scan_counts(ctx, cnt.p, r->rowptr.p, n + 1);
    int32_t nnz = 0;
    peek(ctx, &nnz, r->rowptr.p + n, sizeof(int32_t));
    r->nnz = nnz;
    r->col.alloc(nnz > 0 ? nnz : 1);
    r->w.alloc(nnz > 0 ? nnz : 1);
    k_compact3n<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, r->rowptr.p, ecol.p, ew.p, r->col.p, r->w.p);
    ctx->launches++;
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
}

}  // namespace mprg
