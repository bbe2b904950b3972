// geom.cuh -- fp64 geometry primitives with a fixed operation order.
// Compiled with -fmad=false so that index / mask decisions do not depend on FMA
// contraction (parity with a host restatement must be bit-exact for them).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace mprg {

#define MPRG_HD __host__ __device__ __forceinline__

struct d3 {
    double x, y, z;
};

MPRG_HD d3 ld3(const double *p) { return d3{p[0], p[1], p[2]}; }
MPRG_HD d3 sub(d3 a, d3 b) { return d3{a.x - b.x, a.y - b.y, a.z - b.z}; }
MPRG_HD d3 cross(d3 a, d3 b) {
    return d3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
MPRG_HD double dot(d3 a, d3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
MPRG_HD double dist2(d3 a, d3 b) {
    double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return (dx * dx + dy * dy) + dz * dz;
}

// Ray origin->p against the flat triangle (v0,v1,v2): v0 + a e1 + b e2 = s p.
// Accept iff a,b >= -tol, a+b <= 1+tol, s > 0.  Weights (1-a-b, a, b).
MPRG_HD bool tri_locate(d3 v0, d3 v1, d3 v2, d3 p, double tol, double *w) {
    d3 e1 = sub(v1, v0), e2 = sub(v2, v0);
    d3 q = cross(e2, p);
    double det = dot(e1, q);
    if (det == 0.0) return false;
    d3 r = cross(v0, p);
    d3 n = cross(e2, v0);
    double a = -dot(v0, q) / det;
    double b = -dot(e1, r) / det;
    double s = dot(e1, n) / det;
    if (!(s > 0.0)) return false;
    if (a < -tol || b < -tol || (a + b) > 1.0 + tol) return false;
    w[0] = (1.0 - a) - b;
    w[1] = a;
    w[2] = b;
    return true;
}

// float rounded toward -inf / +inf of a double (conservative boxes)
__device__ __forceinline__ float f_down(double v) { return __double2float_rd(v); }
__device__ __forceinline__ float f_up(double v) { return __double2float_ru(v); }

}  // namespace mprg
