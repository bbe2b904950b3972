// target_gen.cu -- target-grid coordinates, map factors and wind-rotation angles generated ON THE DEVICE
// (SURVEY.md §8 row f3).  The reference computes them on the host, point by point, through the WPS map utilities:
//   coordinates   get_lat_lon_fields (model_grid.F90:2188-2219) -> xytoll (llxy_module.F90:166-216) -> ij_to_latlon
//                 -> ijll_lc / ijll_latlon (module_map_utils.F90:1160-1233, 1398-1428), 4 staggers x (nx x ny) points;
//   map factors   get_map_factor (model_grid.F90:2229-2365);
//   rotation      get_rotang (model_grid.F90:2450-2507).
// Here the per-grid scalars of the projection (set_lc, module_map_utils.F90:1083-1121: cone, polei, polej, rebydx) are
// computed once on the host by the host code that owns the namelist and handed over in mprg_projection; everything
// that is O(nx x ny) runs as one kernel per stagger in fp64 with the same formulas in the same operation order.
// CUDA's fp64 sin / cos / atan2 / pow / log / tan are within 1-2 ulp of the host libm, not bit-identical: coordinates
// agree with the host mirror to a few ulp, and a destination point that sits within an ulp of a triangle edge could in
// principle choose the neighbouring triangle.  tests/test_gpu_targetgen.py measures both (coordinate ulps; CSR
// structure identical on the BASELINE configs); the host path stays available and bit-exact with a CPU run.
// Compiled with -fmad=false like the other geometry units.
#include "common.cuh"

namespace mprg {

namespace {
constexpr double PI = 3.141592653589793;              // constants_module.F90:9
constexpr double DEG_PER_RAD = 180.0 / PI;
constexpr double RAD_PER_DEG = PI / 180.0;

// ijll_lc, module_map_utils.F90:1160-1233
__device__ void ijll_lc(double i, double j, const mprg_projection &p, double *lat, double *lon) {
    const double chi1 = (90.0 - p.hemi * p.truelat1) * RAD_PER_DEG;
    const double chi2 = (90.0 - p.hemi * p.truelat2) * RAD_PER_DEG;
    const double inew = p.hemi * i, jnew = p.hemi * j;
    const double xx = inew - p.polei, yy = p.polej - jnew;
    const double r2 = xx * xx + yy * yy;
    const double r = sqrt(r2) / p.rebydx;
    if (r2 == 0.0) {
        *lat = p.hemi * 90.0;
        *lon = p.stdlon;
    } else {
        double lo = p.stdlon + DEG_PER_RAD * atan2(p.hemi * xx, yy) / p.cone;
        lo = fmod(lo + 360.0, 360.0);
        double chi;
        if (chi1 == chi2) chi = 2.0 * atan(pow(r / tan(chi1), 1.0 / p.cone) * tan(chi1 * 0.5));
        else chi = 2.0 * atan(pow(r * p.cone / sin(chi1), 1.0 / p.cone) * tan(chi1 * 0.5));
        *lat = (90.0 - chi * DEG_PER_RAD) * p.hemi;
        *lon = lo;
    }
    if (*lon > 180.0) *lon -= 360.0;
    if (*lon < -180.0) *lon += 360.0;
}

// ijll_latlon, module_map_utils.F90:1398-1428 (longitudes are NOT wrapped)
__device__ void ijll_latlon(double i, double j, const mprg_projection &p, double *lat, double *lon) {
    double i_work = i;
    if (i < (double)p.nxmin - 0.5) i_work = i + (double)(p.nxmax - p.nxmin + 1);
    if (i >= (double)p.nxmax + 0.5) i_work = i - (double)(p.nxmax - p.nxmin + 1);
    i_work = i_work - p.knowni;
    const double j_work = j - p.knownj;
    *lat = p.lat1 + j_work * p.latinc;
    *lon = p.lon1 + i_work * p.loninc;
}

__global__ void k_target_gen(mprg_projection p, int32_t ni, int32_t nj, double ox, double oy, double *__restrict__ lon,
                             double *__restrict__ lat, double *__restrict__ xyz) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)ni * nj) return;
    const int32_t j = (int32_t)(t / ni) + 1, i = (int32_t)(t - (int64_t)(j - 1) * ni) + 1;
    // get_lat_lon_fields, model_grid.F90:2212-2217 with rx = ry = 1; xytoll stagger offsets, llxy_module.F90:182-203
    const double x = ((double)i - 0.5) / 1.0 + 0.5, y = ((double)j - 0.5) / 1.0 + 0.5;
    const double rx = x - ox, ry = y - oy;
    double la, lo;
    if (p.code == 1) ijll_lc(rx, ry, p, &la, &lo);
    else ijll_latlon(rx, ry, p, &la, &lo);
    lon[t] = lo;
    lat[t] = la;
    // ESMF_COORDSYS_SPH_DEG -> Cartesian (mesh.cu: deg_to_cart)
    const double DEG2RAD = 3.141592653589793238 / 180.0;
    const double th = lo * DEG2RAD, ph = (90.0 - la) * DEG2RAD;
    const double sp = sin(ph);
    xyz[3 * t + 0] = cos(th) * sp;
    xyz[3 * t + 1] = sin(th) * sp;
    xyz[3 * t + 2] = cos(ph);
}

// get_map_factor, model_grid.F90:2229-2365 (Saucier pp. 32-33), Lambert only; 0 for the other projections (the
// reference writes the array uninitialised there)
__global__ void k_map_factor(int proj_code, double truelat1, double truelat2, const double *__restrict__ xlat, int64_t n,
                             double *__restrict__ mapfac) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (proj_code != 1) { mapfac[k] = 0.0; return; }
    const double colat = RAD_PER_DEG * (90.0 - xlat[k]);
    if (truelat1 != truelat2) {
        const double colat1 = RAD_PER_DEG * (90.0 - truelat1), colat2 = RAD_PER_DEG * (90.0 - truelat2);
        const double nn = (log(sin(colat1)) - log(sin(colat2))) / (log(tan(colat1 / 2.0)) - log(tan(colat2 / 2.0)));
        mapfac[k] = sin(colat2) / sin(colat) * pow(tan(colat / 2.0) / tan(colat2 / 2.0), nn);
    } else {
        const double colat0 = RAD_PER_DEG * (90.0 - truelat1);
        mapfac[k] = sin(colat0) / sin(colat) * pow(tan(colat / 2.0) / tan(colat0 / 2.0), cos(colat0));
    }
}

// get_rotang, model_grid.F90:2450-2507: centred in j, one-sided on the first / last row of the (global) grid
__global__ void k_rotang(const double *__restrict__ xlat, const double *__restrict__ xlon, int32_t ni, int32_t nj,
                         double *__restrict__ cosa, double *__restrict__ sina) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)ni * nj) return;
    const int32_t j = (int32_t)(t / ni);
    int64_t a = t + ni, b = t - ni;            // j + 1, j - 1
    if (nj < 2) { cosa[t] = 1.0; sina[t] = 0.0; return; }
    if (j == 0) b = t;
    if (j == nj - 1) a = t;
    double d_lon = xlon[a] - xlon[b];
    const double d_lat = xlat[a] - xlat[b];
    if (d_lon > 180.0) d_lon -= 360.0;
    else if (d_lon < -180.0) d_lon += 360.0;
    const double alpha = atan2(-cos(xlat[t] * RAD_PER_DEG) * (d_lon * RAD_PER_DEG), d_lat * RAD_PER_DEG);
    sina[t] = sin(alpha);
    cosa[t] = cos(alpha);
}
}  // namespace

void target_generate(mprg_ctx *ctx, int stagger, int32_t ni, int32_t nj, const mprg_projection *p) {
    if (stagger < 0 || stagger > 3) fail(15, "mprg_set_target_projected: bad stagger %d", stagger);
    if (ni <= 0 || nj <= 0 || !p) fail(16, "mprg_set_target_projected: empty grid");
    if (p->code != 0 && p->code != 1) fail(17, "mprg_set_target_projected: projection %d is not supported (lat-lon, Lambert)", p->code);
    Target &t = ctx->target[stagger];
    const int64_t n = (int64_t)ni * nj;
    t.xyz.alloc(3 * (size_t)n);
    t.lon.alloc(n);
    t.lat.alloc(n);
    const double ox = (stagger == MPRG_EDGE1 || stagger == MPRG_CORNER) ? 0.5 : 0.0;
    const double oy = (stagger == MPRG_EDGE2 || stagger == MPRG_CORNER) ? 0.5 : 0.0;
    k_target_gen<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(*p, ni, nj, ox, oy, t.lon.p, t.lat.p, t.xyz.p);
    ctx->launches++;
    MPRG_CUDA(cudaGetLastError());
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
    target_register(ctx, stagger, ni, nj);
}

void target_map_factor(mprg_ctx *ctx, int stagger, int proj_code, double truelat1, double truelat2, double *out_dev) {
    const Target &t = ctx->target[stagger];
    if (!t.set || !t.lat.p) fail(18, "mprg_target_map_factor: stagger %d has no coordinates", stagger);
    const int64_t n = (int64_t)t.ni * t.nj;
    k_map_factor<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(proj_code, truelat1, truelat2, t.lat.p, n, out_dev);
    ctx->launches++;
    MPRG_CUDA(cudaGetLastError());
}

void target_rotang(mprg_ctx *ctx) {
    const Target &t = ctx->target[MPRG_CENTER];
    if (!t.set || !t.lat.p) fail(18, "mprg_set_rotation_from_target: the CENTER stagger has no coordinates");
    const int64_t n = (int64_t)t.ni * t.nj;
    ctx->cosa.alloc(n);
    ctx->sina.alloc(n);
    k_rotang<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(t.lat.p, t.lon.p, t.ni, t.nj, ctx->cosa.p, ctx->sina.p);
    ctx->launches++;
    MPRG_CUDA(cudaGetLastError());
}

}  // namespace mprg
