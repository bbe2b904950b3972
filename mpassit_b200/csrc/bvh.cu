// bvh.cu -- build of the implicit, Morton-ordered bounding-volume hierarchy used by
// the weight-generation ("RegridStore") kernels to find candidate source
// entities for every target point / target cell.
#include <cub/device/device_radix_sort.cuh>

#include "bvh.cuh"
#include "common.cuh"

namespace mprg {

__device__ __forceinline__ uint64_t spread21(uint64_t v) {
    v &= 0x1fffffULL;
    v = (v | v << 32) & 0x1f00000000ffffULL;
    v = (v | v << 16) & 0x1f0000ff0000ffULL;
    v = (v | v << 8) & 0x100f00f00f00f00fULL;
    v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
    v = (v | v << 2) & 0x1249249249249249ULL;
    return v;
}

__device__ __forceinline__ uint64_t morton63(double x, double y, double z) {
    // unit sphere lives in [-1,1]^3
    const double s = 1048576.0;  // 2^20 per unit => 2^21 cells over [-1,1]
    auto q = [&](double v) {
        double t = (v + 1.0) * s;
        t = fmin(fmax(t, 0.0), 2097151.0);
        return (uint64_t)t;
    };
    return spread21(q(x)) | (spread21(q(y)) << 1) | (spread21(q(z)) << 2);
}

__global__ void k_morton_points(const double *__restrict__ xyz, int32_t n, uint64_t *keys, int32_t *ids) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = morton63(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2]);
    ids[i] = i;
}

__global__ void k_morton_boxes(const float *__restrict__ lo, const float *__restrict__ hi, int32_t n,
                               uint64_t *keys, int32_t *ids) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *l = lo + 3 * (size_t)i, *h = hi + 3 * (size_t)i;
    if (l[0] > h[0]) {  // empty primitive: park at the end of the order
        keys[i] = ~0ULL;
    } else {
        keys[i] = morton63(0.5 * ((double)l[0] + h[0]), 0.5 * ((double)l[1] + h[1]), 0.5 * ((double)l[2] + h[2]));
    }
    ids[i] = i;
}

__global__ void k_gather_xyz(const double *__restrict__ xyz, const int32_t *__restrict__ ids, int32_t n,
                             double *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int id = ids[i];
    out[3 * (size_t)i] = xyz[3 * (size_t)id];
    out[3 * (size_t)i + 1] = xyz[3 * (size_t)id + 1];
    out[3 * (size_t)i + 2] = xyz[3 * (size_t)id + 2];
}

// leaf boxes from Morton-sorted points (sortedXyz) or from boxes via primId
__global__ void k_leaf_points(const double *__restrict__ sortedXyz, int32_t n, int32_t L, float4 *nodes) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= L) return;
    float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
    int s0 = k * kLeaf, s1 = min(s0 + kLeaf, n);
    for (int s = s0; s < s1; ++s) {
        double x = sortedXyz[3 * (size_t)s], y = sortedXyz[3 * (size_t)s + 1], z = sortedXyz[3 * (size_t)s + 2];
        lx = fminf(lx, f_down(x)); ly = fminf(ly, f_down(y)); lz = fminf(lz, f_down(z));
        hx = fmaxf(hx, f_up(x)); hy = fmaxf(hy, f_up(y)); hz = fmaxf(hz, f_up(z));
    }
    size_t node = (size_t)(L - 1) + k;
    nodes[2 * node] = make_float4(lx, ly, lz, hx);
    nodes[2 * node + 1] = make_float4(hy, hz, 0.f, 0.f);
}

__global__ void k_leaf_boxes(const float *__restrict__ lo, const float *__restrict__ hi,
                             const int32_t *__restrict__ primId, int32_t n, int32_t L, float4 *nodes) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= L) return;
    float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
    int s0 = k * kLeaf, s1 = min(s0 + kLeaf, n);
    for (int s = s0; s < s1; ++s) {
        int id = primId[s];
        const float *l = lo + 3 * (size_t)id, *h = hi + 3 * (size_t)id;
        if (l[0] > h[0]) continue;
        lx = fminf(lx, l[0]); ly = fminf(ly, l[1]); lz = fminf(lz, l[2]);
        hx = fmaxf(hx, h[0]); hy = fmaxf(hy, h[1]); hz = fmaxf(hz, h[2]);
    }
    size_t node = (size_t)(L - 1) + k;
    nodes[2 * node] = make_float4(lx, ly, lz, hx);
    nodes[2 * node + 1] = make_float4(hy, hz, 0.f, 0.f);
}

// per-primitive boxes in Morton order (lets a leaf visitor reject a primitive without touching its geometry)
__global__ void k_prim_boxes(const float *__restrict__ lo, const float *__restrict__ hi,
                             const int32_t *__restrict__ primId, int32_t n, float4 *__restrict__ out) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int id = primId[s];
    const float *l = lo + 3 * (size_t)id, *h = hi + 3 * (size_t)id;
    out[2 * (size_t)s] = make_float4(l[0], l[1], l[2], h[0]);
    out[2 * (size_t)s + 1] = make_float4(h[1], h[2], 0.f, 0.f);
}

__global__ void k_inner_level(float4 *nodes, int32_t first, int32_t count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    size_t node = (size_t)first + i;
    float4 a1 = nodes[2 * (2 * node + 1)], b1 = nodes[2 * (2 * node + 1) + 1];
    float4 a2 = nodes[2 * (2 * node + 2)], b2 = nodes[2 * (2 * node + 2) + 1];
    nodes[2 * node] = make_float4(fminf(a1.x, a2.x), fminf(a1.y, a2.y), fminf(a1.z, a2.z), fmaxf(a1.w, a2.w));
    nodes[2 * node + 1] = make_float4(fmaxf(b1.x, b2.x), fmaxf(b1.y, b2.y), 0.f, 0.f);
}

static int32_t next_pow2(int32_t v) {
    int32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

static void sort_by_morton(mprg_ctx *ctx, DevBuf<uint64_t> &keys, DevBuf<int32_t> &ids, int32_t n,
                           DevBuf<int32_t> &outIds) {
    DevBuf<uint64_t> keysOut(n);
    outIds.alloc(n);
    size_t tmpBytes = 0;
    MPRG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys.p, keysOut.p, ids.p, outIds.p, n, 0, 63,
                                              ctx->stream));
    DevBuf<unsigned char> tmp(tmpBytes);
    MPRG_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, keys.p, keysOut.p, ids.p, outIds.p, n, 0, 63,
                                              ctx->stream));
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
}

static void build_inner(mprg_ctx *ctx, Bvh &out) {
    int32_t L = out.nLeafNodes;
    for (int32_t count = L / 2; count >= 1; count /= 2) {
        int32_t first = count - 1;
        k_inner_level<<<(count + 255) / 256, 256, 0, ctx->stream>>>(out.nodes.p, first, count);
        ctx->launches++;
    }
    MPRG_CUDA(cudaGetLastError());
}

void bvh_build_points(mprg_ctx *ctx, const double *xyz_dev, int32_t n, Bvh &out, DevBuf<double> *sortedXyz) {
    if (n <= 0) fail(21, "bvh_build_points: no primitives");
    DevBuf<uint64_t> keys(n);
    DevBuf<int32_t> ids(n);
    k_morton_points<<<(n + 255) / 256, 256, 0, ctx->stream>>>(xyz_dev, n, keys.p, ids.p);
    ctx->launches++;
    sort_by_morton(ctx, keys, ids, n, out.primId);
    out.nPrim = n;
    out.nLeafNodes = next_pow2((n + kLeaf - 1) / kLeaf);
    out.nodes.alloc(2 * (size_t)(2 * out.nLeafNodes - 1));
    sortedXyz->alloc(3 * (size_t)n);
    k_gather_xyz<<<(n + 255) / 256, 256, 0, ctx->stream>>>(xyz_dev, out.primId.p, n, sortedXyz->p);
    k_leaf_points<<<(out.nLeafNodes + 255) / 256, 256, 0, ctx->stream>>>(sortedXyz->p, n, out.nLeafNodes,
                                                                          out.nodes.p);
    ctx->launches += 2;
    build_inner(ctx, out);
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
}

void bvh_build_boxes(mprg_ctx *ctx, const float *lo_dev, const float *hi_dev, int32_t n, Bvh &out,
                     bool keepPrimBoxes) {
    if (n <= 0) fail(21, "bvh_build_boxes: no primitives");
    DevBuf<uint64_t> keys(n);
    DevBuf<int32_t> ids(n);
    k_morton_boxes<<<(n + 255) / 256, 256, 0, ctx->stream>>>(lo_dev, hi_dev, n, keys.p, ids.p);
    ctx->launches++;
    sort_by_morton(ctx, keys, ids, n, out.primId);
    out.nPrim = n;
    out.nLeafNodes = next_pow2((n + kLeaf - 1) / kLeaf);
    out.nodes.alloc(2 * (size_t)(2 * out.nLeafNodes - 1));
    k_leaf_boxes<<<(out.nLeafNodes + 255) / 256, 256, 0, ctx->stream>>>(lo_dev, hi_dev, out.primId.p, n,
                                                                         out.nLeafNodes, out.nodes.p);
    ctx->launches++;
    if (keepPrimBoxes) {
        out.primBox.alloc(2 * (size_t)n);
        k_prim_boxes<<<(n + 255) / 256, 256, 0, ctx->stream>>>(lo_dev, hi_dev, out.primId.p, n, out.primBox.p);
        ctx->launches++;
    }
    build_inner(ctx, out);
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
}

}  // namespace mprg
