// bvh.cuh -- device-side traversal of the implicit BVH built in bvh.cu.
//
// Layout: complete binary tree in heap order over L = nLeafNodes (power of two)
// leaf nodes; node i has children 2i+1, 2i+2; leaf node k (heap index L-1+k)
// covers Morton-sorted primitives [k*kLeaf, min((k+1)*kLeaf, nPrim)).  Each node
// is two float4: (lo.x, lo.y, lo.z, hi.x) (hi.y, hi.z, -, -), boxes rounded
// outward so they contain their fp64 primitives.  Empty nodes are inverted
// (+inf, -inf) and never hit.
#pragma once
#include "geom.cuh"

namespace mprg {

constexpr int kLeaf = 8;
constexpr int kStack = 64;

struct BvhView {
    const float4 *nodes;
    const int32_t *primId;
    int32_t nLeafNodes;
    int32_t nPrim;
    const float4 *primBox = nullptr;  // optional per-primitive boxes, Morton order
};

// does primitive s (Morton order) overlap the query box?  true when no per-primitive boxes are kept
__device__ __forceinline__ bool prim_hit(const BvhView &t, int s, d3 qlo, d3 qhi) {
    if (!t.primBox) return true;
    const float4 a = __ldg(t.primBox + 2 * (size_t)s), b = __ldg(t.primBox + 2 * (size_t)s + 1);
    return (double)a.x <= qhi.x && (double)a.w >= qlo.x && (double)a.y <= qhi.y && (double)b.x >= qlo.y &&
           (double)a.z <= qhi.z && (double)b.y >= qlo.z;
}

__device__ __forceinline__ void node_box(const float4 *nodes, int i, float &lx, float &ly, float &lz,
                                         float &hx, float &hy, float &hz) {
    float4 a = __ldg(nodes + 2 * (size_t)i);
    float4 b = __ldg(nodes + 2 * (size_t)i + 1);
    lx = a.x; ly = a.y; lz = a.z; hx = a.w; hy = b.x; hz = b.y;
}

__device__ __forceinline__ double box_dist2(const float4 *nodes, int i, d3 q) {
    float lx, ly, lz, hx, hy, hz;
    node_box(nodes, i, lx, ly, lz, hx, hy, hz);
    double dx = fmax(fmax((double)lx - q.x, 0.0), q.x - (double)hx);
    double dy = fmax(fmax((double)ly - q.y, 0.0), q.y - (double)hy);
    double dz = fmax(fmax((double)lz - q.z, 0.0), q.z - (double)hz);
    return (dx * dx + dy * dy) + dz * dz;  // +inf for empty nodes
}

// Nearest point (3-D chord distance), ties -> smallest original id.
// sortedXyz: coordinates in Morton order, [nPrim][3].
__device__ inline void bvh_nearest(const BvhView &t, const double *__restrict__ sortedXyz, d3 q,
                                   double &best, int32_t &bid) {
    int stack[kStack];
    int sp = 0;
    stack[sp++] = 0;
    const int firstLeaf = t.nLeafNodes - 1;
    while (sp > 0) {
        int node = stack[--sp];
        if (node != 0) {
            // re-check against the (possibly improved) best; slack covers fp64 rounding
            if (box_dist2(t.nodes, node, q) * (1.0 - 1e-14) > best) continue;
        }
        if (node >= firstLeaf) {
            int s0 = (node - firstLeaf) * kLeaf;
            int s1 = min(s0 + kLeaf, t.nPrim);
            for (int s = s0; s < s1; ++s) {
                d3 p = ld3(sortedXyz + 3 * (size_t)s);
                double d = dist2(q, p);
                int32_t id = __ldg(t.primId + s);
                if (d < best || (d == best && id < bid)) { best = d; bid = id; }
            }
        } else {
            int c1 = 2 * node + 1, c2 = c1 + 1;
            double d1 = box_dist2(t.nodes, c1, q), d2 = box_dist2(t.nodes, c2, q);
            // push the farther child first so the nearer one is popped next
            if (d1 <= d2) {
                if (d2 * (1.0 - 1e-14) <= best) stack[sp++] = c2;
                if (d1 * (1.0 - 1e-14) <= best) stack[sp++] = c1;
            } else {
                if (d1 * (1.0 - 1e-14) <= best) stack[sp++] = c1;
                if (d2 * (1.0 - 1e-14) <= best) stack[sp++] = c2;
            }
        }
    }
}

// Visit every leaf range whose box overlaps the query box [qlo, qhi] (a point
// query passes qlo == qhi).  f(s0, s1) receives a Morton-order range.
template <typename F>
__device__ inline void bvh_overlap(const BvhView &t, d3 qlo, d3 qhi, F f) {
    int stack[kStack];
    int sp = 0;
    stack[sp++] = 0;
    const int firstLeaf = t.nLeafNodes - 1;
    while (sp > 0) {
        int node = stack[--sp];
        float lx, ly, lz, hx, hy, hz;
        node_box(t.nodes, node, lx, ly, lz, hx, hy, hz);
        bool hit = (double)lx <= qhi.x && (double)hx >= qlo.x && (double)ly <= qhi.y && (double)hy >= qlo.y &&
                   (double)lz <= qhi.z && (double)hz >= qlo.z;
        if (!hit) continue;
        if (node >= firstLeaf) {
            int s0 = (node - firstLeaf) * kLeaf;
            int s1 = min(s0 + kLeaf, t.nPrim);
            f(s0, s1);
        } else {
            stack[sp++] = 2 * node + 2;
            stack[sp++] = 2 * node + 1;
        }
    }
}

}  // namespace mprg
