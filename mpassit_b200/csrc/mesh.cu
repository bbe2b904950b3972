// mesh.cu -- source mesh and target grid ingestion.
//
// Replaces what the reference hands to ESMF_MeshCreate (model_grid.F90:446-497)
// and ESMF_GridAddCoord/GetCoord (model_grid.F90:707-1038): coordinates go to
// unit-sphere Cartesian fp64, the dual (Delaunay) triangles are derived by
// inverting verticesOnCell on the device, and search structures are built
// lazily the first time a regrid method needs them.
#include <algorithm>
#include <cmath>
#include <thread>

#include "bvh.cuh"
#include "common.cuh"

namespace mprg {

// ---- host-side coordinate marshalling ------------------------------------
// The trig here is deliberately done with the host libm, in the same operation
// order the reference/ESMF use, so Cartesian coordinates (and therefore every
// index / mask decision derived from them) are bit-identical to a CPU run.
template <typename F>
static void parallel_for(int64_t n, F f) {
    unsigned hw = std::thread::hardware_concurrency();
    int T = (int)(hw == 0 ? 1 : (hw > 32 ? 32 : hw));
    if (n < 65536) T = 1;
    if (T == 1) { f(0, n); return; }
    std::vector<std::thread> th;
    int64_t chunk = (n + T - 1) / T;
    for (int t = 0; t < T; ++t) {
        int64_t b = t * chunk, e = b + chunk < n ? b + chunk : n;
        if (b >= e) break;
        th.emplace_back([=] { f(b, e); });
    }
    for (auto &x : th) x.join();
}

static inline void deg_to_cart(double lon_deg, double lat_deg, double *xyz) {
    // ESMF_COORDSYS_SPH_DEG -> Cartesian
    const double DEG2RAD = 3.141592653589793238 / 180.0;
    double th = lon_deg * DEG2RAD;
    double ph = (90.0 - lat_deg) * DEG2RAD;
    double sp = sin(ph);
    xyz[0] = cos(th) * sp;
    xyz[1] = sin(th) * sp;
    xyz[2] = cos(ph);
}

static void mesh_rad_to_cart(int64_t n, const double *lon_rad, const double *lat_rad, double *xyz) {
    // model_grid.F90:450-454 / :464-468: degrees with PI = 4*atan(1), wrap > 180
    const double PI = 4.0 * atan(1.0);
    parallel_for(n, [=](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; ++i) {
            double lo = lon_rad[i] * 180.0 / PI;
            if (lo > 180.0) lo = lo - 360.0;
            double la = lat_rad[i] * 180.0 / PI;
            deg_to_cart(lo, la, xyz + 3 * i);
        }
    });
}

// ---- dual triangles --------------------------------------------------------
__global__ void k_dual_scatter(const int32_t *__restrict__ voc, int64_t nSlots, int32_t maxEdges,
                               int32_t nVertices, int32_t *cnt, int32_t *tri, int32_t *flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nSlots) return;
    int32_t v = voc[i];
    if (v <= 0) return;
    v -= 1;
    if (v >= nVertices) { atomicOr(flag, 2); return; }
    int32_t cell = (int32_t)(i / maxEdges);
    int pos = atomicAdd(cnt + v, 1);
    if (pos < 3) tri[3 * (size_t)v + pos] = cell;
    else atomicOr(flag, 1);
}

__global__ void k_dual_finish(int32_t nVertices, const int32_t *__restrict__ cnt, int32_t *tri) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nVertices) return;
    int32_t *t = tri + 3 * (size_t)v;
    if (cnt[v] != 3) { t[0] = t[1] = t[2] = -1; return; }
    int32_t a = t[0], b = t[1], c = t[2], x;
    if (a > b) { x = a; a = b; b = x; }
    if (b > c) { x = b; b = c; c = x; }
    if (a > b) { x = a; a = b; b = x; }
    t[0] = a; t[1] = b; t[2] = c;
}

__global__ void k_tri_boxes(int32_t nVertices, const int32_t *__restrict__ tri, const double *__restrict__ cxyz,
                            float *lo, float *hi) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nVertices) return;
    const int32_t *t = tri + 3 * (size_t)v;
    float *l = lo + 3 * (size_t)v, *h = hi + 3 * (size_t)v;
    if (t[0] < 0) { l[0] = l[1] = l[2] = INFINITY; h[0] = h[1] = h[2] = -INFINITY; return; }
    d3 a = ld3(cxyz + 3 * (size_t)t[0]), b = ld3(cxyz + 3 * (size_t)t[1]), c = ld3(cxyz + 3 * (size_t)t[2]);
    double e2 = fmax(dist2(a, b), fmax(dist2(b, c), dist2(c, a)));
    // any point x of the flat triangle has |x|^2 >= 1 - E^2/3, so its radial image on
    // the sphere is within 1-|x| of it; add the parametric tolerance and slack.
    double m = (1.0 - sqrt(fmax(0.0, 1.0 - e2 / 3.0))) * 1.01 + sqrt(e2) * 1e-9 + 1e-12;
    l[0] = f_down(fmin(a.x, fmin(b.x, c.x)) - m);
    l[1] = f_down(fmin(a.y, fmin(b.y, c.y)) - m);
    l[2] = f_down(fmin(a.z, fmin(b.z, c.z)) - m);
    h[0] = f_up(fmax(a.x, fmax(b.x, c.x)) + m);
    h[1] = f_up(fmax(a.y, fmax(b.y, c.y)) + m);
    h[2] = f_up(fmax(a.z, fmax(b.z, c.z)) + m);
}

__global__ void k_poly_boxes(int32_t nCells, int32_t maxEdges, const int32_t *__restrict__ voc,
                             const double *__restrict__ vxyz, float *lo, float *hi) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    float *l = lo + 3 * (size_t)c, *h = hi + 3 * (size_t)c;
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    double e2 = 0.0;
    int n = 0;
    d3 first{0, 0, 0};
    for (int k = 0; k < maxEdges; ++k) {
        int32_t v = voc[(size_t)c * maxEdges + k];
        if (v <= 0) continue;
        d3 p = ld3(vxyz + 3 * (size_t)(v - 1));
        mn[0] = fmin(mn[0], p.x); mn[1] = fmin(mn[1], p.y); mn[2] = fmin(mn[2], p.z);
        mx[0] = fmax(mx[0], p.x); mx[1] = fmax(mx[1], p.y); mx[2] = fmax(mx[2], p.z);
        if (n == 0) first = p; else e2 = fmax(e2, dist2(first, p));
        ++n;
    }
    if (n < 3) { l[0] = l[1] = l[2] = INFINITY; h[0] = h[1] = h[2] = -INFINITY; return; }
    // The polygon's diameter D is at most twice the largest distance from its first vertex.
    // Great-circle edges bulge beyond their chords by <= 1 - sqrt(1 - D^2/4) and the radial image of
    // any flat fan triangle lies within 1 - sqrt(1 - D^2/3) of it; cover both.
    double D2 = 4.0 * e2;
    double m = (1.0 - sqrt(fmax(0.0, 1.0 - D2 / 3.0))) * 1.01 + sqrt(D2) * 1e-9 + 1e-12;
    for (int a = 0; a < 3; ++a) { l[a] = f_down(mn[a] - m); h[a] = f_up(mx[a] + m); }
}

void mesh_set(mprg_ctx *ctx, int32_t nCells, int32_t nVertices, int32_t maxEdges, const double *lonC,
              const double *latC, const double *lonV, const double *latV, const int32_t *voc) {
    if (nCells <= 0 || nVertices <= 0 || maxEdges <= 0) fail(11, "mprg_set_mesh: empty mesh");
    if (!lonC || !latC || !lonV || !latV || !voc) fail(12, "mprg_set_mesh: null array");
    Mesh &m = ctx->mesh;
    m = Mesh();
    m.nCells = nCells; m.nVertices = nVertices; m.maxEdges = maxEdges;
    std::vector<double> cx(3 * (size_t)nCells), vx(3 * (size_t)nVertices);
    mesh_rad_to_cart(nCells, lonC, latC, cx.data());
    mesh_rad_to_cart(nVertices, lonV, latV, vx.data());
    m.cellXyz.alloc(cx.size());
    m.vertXyz.alloc(vx.size());
    m.voc.alloc((size_t)nCells * maxEdges);
    m.tri.alloc(3 * (size_t)nVertices);
    cudaStream_t s = ctx->stream;
    MPRG_CUDA(cudaMemcpyAsync(m.cellXyz.p, cx.data(), m.cellXyz.bytes(), cudaMemcpyHostToDevice, s));
    MPRG_CUDA(cudaMemcpyAsync(m.vertXyz.p, vx.data(), m.vertXyz.bytes(), cudaMemcpyHostToDevice, s));
    MPRG_CUDA(cudaMemcpyAsync(m.voc.p, voc, m.voc.bytes(), cudaMemcpyHostToDevice, s));
    DevBuf<int32_t> cnt(nVertices), flag(1);
    MPRG_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), s));
    MPRG_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int32_t), s));
    MPRG_CUDA(cudaMemsetAsync(m.tri.p, 0xff, m.tri.bytes(), s));
    int64_t nSlots = (int64_t)nCells * maxEdges;
    k_dual_scatter<<<(unsigned)((nSlots + 255) / 256), 256, 0, s>>>(m.voc.p, nSlots, maxEdges, nVertices, cnt.p,
                                                                    m.tri.p, flag.p);
    k_dual_finish<<<(nVertices + 255) / 256, 256, 0, s>>>(nVertices, cnt.p, m.tri.p);
    ctx->launches += 2;
    int32_t hflag = 0;
    MPRG_CUDA(cudaMemcpyAsync(&hflag, flag.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    MPRG_CUDA(cudaStreamSynchronize(s));
    if (hflag & 2) fail(13, "mprg_set_mesh: verticesOnCell entry exceeds nVertices");
    if (hflag & 1) fail(14, "mprg_set_mesh: a vertex is shared by more than 3 cells (vertexDegree > 3 unsupported)");
    // geometry changed: every memoised route is stale
    for (auto &kv : ctx->routes) { kv.second->memoised = false; if (kv.second->refcount <= 0) delete kv.second; }
    ctx->routes.clear();
    for (bool &d : ctx->windDeclined) d = false;
}

void mesh_need_cell_bvh(mprg_ctx *ctx) {
    Mesh &m = ctx->mesh;
    if (m.haveCellBvh) return;
    bvh_build_points(ctx, m.cellXyz.p, m.nCells, m.cellBvh, &m.cellSorted);
    m.haveCellBvh = true;
}

void mesh_need_tri_bvh(mprg_ctx *ctx) {
    Mesh &m = ctx->mesh;
    if (m.haveTriBvh) return;
    DevBuf<float> lo(3 * (size_t)m.nVertices), hi(3 * (size_t)m.nVertices);
    k_tri_boxes<<<(m.nVertices + 255) / 256, 256, 0, ctx->stream>>>(m.nVertices, m.tri.p, m.cellXyz.p, lo.p, hi.p);
    ctx->launches++;
    bvh_build_boxes(ctx, lo.p, hi.p, m.nVertices, m.triBvh);
    m.haveTriBvh = true;
}

void mesh_need_poly_bvh(mprg_ctx *ctx) {
    Mesh &m = ctx->mesh;
    if (m.havePolyBvh) return;
    DevBuf<float> lo(3 * (size_t)m.nCells), hi(3 * (size_t)m.nCells);
    k_poly_boxes<<<(m.nCells + 255) / 256, 256, 0, ctx->stream>>>(m.nCells, m.maxEdges, m.voc.p, m.vertXyz.p,
                                                                  lo.p, hi.p);
    ctx->launches++;
    bvh_build_boxes(ctx, lo.p, hi.p, m.nCells, m.polyBvh, /*keepPrimBoxes=*/true);
    m.havePolyBvh = true;
}

// the stagger's Cartesian coordinates are in t.xyz: slab bounds, halo view, stale routes
void target_register(mprg_ctx *ctx, int stagger, int32_t ni, int32_t nj) {
    Target &t = ctx->target[stagger];
    t.ni = ni; t.nj = nj;
    para_range(nj, ctx->nranks, ctx->rank, &t.j0, &t.j1);
    t.set = true;
    t.xyzRef = nullptr;
    if (stagger == MPRG_CENTER) {
        // CENTER rows +- one halo row: what this rank's EDGE1 / EDGE2 rows read (para_range over nj and
        // nj + 1 differ by at most one row at each slab boundary)
        Target &h = ctx->target[MPRG_CENTER_HALO];
        h.ni = ni; h.nj = nj;
        h.j0 = ctx->nranks > 1 ? std::max(0, t.j0 - 1) : t.j0;
        h.j1 = ctx->nranks > 1 ? std::min(nj, t.j1 + 1) : t.j1;
        h.xyzRef = t.xyz.p;
        h.set = true;
        ctx->haveRot = false;  // angles belong to the previous grid
    }
    // routes into this stagger (or out of CENTER) are stale
    for (bool &d : ctx->windDeclined) d = false;
    for (auto it = ctx->routes.begin(); it != ctx->routes.end();) {
        int dst = std::get<2>(it->first), srcloc = std::get<1>(it->first);
        if (dst == stagger || (stagger == MPRG_CENTER && (srcloc == MPRG_SRC_GRID_CENTER || srcloc == MPRG_SRC_MESH_WIND || dst == MPRG_CENTER_HALO))) {
            it->second->memoised = false;
            if (it->second->refcount <= 0) delete it->second;
            it = ctx->routes.erase(it);
        } else {
            ++it;
        }
    }
}

void target_set(mprg_ctx *ctx, int stagger, int32_t ni, int32_t nj, const double *lon, const double *lat) {
    if (stagger < 0 || stagger > 3) fail(15, "mprg_set_target: bad stagger %d", stagger);
    if (ni <= 0 || nj <= 0 || !lon || !lat) fail(16, "mprg_set_target: empty grid");
    Target &t = ctx->target[stagger];
    int64_t n = (int64_t)ni * nj;
    std::vector<double> x(3 * (size_t)n);
    double *xp = x.data();
    parallel_for(n, [=](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; ++i) deg_to_cart(lon[i], lat[i], xp + 3 * i);
    });
    t.xyz.alloc(x.size());
    // degrees kept too: the device-side map factors / rotation angles (target_gen.cu) read them
    t.lon.alloc(n);
    t.lat.alloc(n);
    MPRG_CUDA(cudaMemcpyAsync(t.xyz.p, x.data(), t.xyz.bytes(), cudaMemcpyHostToDevice, ctx->stream));
    MPRG_CUDA(cudaMemcpyAsync(t.lon.p, lon, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    MPRG_CUDA(cudaMemcpyAsync(t.lat.p, lat, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
    target_register(ctx, stagger, ni, nj);
}

}  // namespace mprg
