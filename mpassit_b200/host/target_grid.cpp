// target_grid.cpp -- host mirror of define_target_grid_params (model_grid.F90:644-1201):
// the lat/lon of every stagger of the projected target grid and the wind
// rotation angles.  Projection arithmetic restates the WPS map utilities the
// reference links (module_map_utils.F90) for the two projections that appear in
// BASELINE.json's configs: Lambert conformal and cylindrical lat-lon.
// All arithmetic is fp64, as in the reference's -r8 build.
#include <cmath>
#include <cstdio>
#include <string>

#include "../../include/mpassit_host.h"
#include "par.hpp"

namespace {

// constants_module.F90:9-27
const double PI = 3.141592653589793;
const double DEG_PER_RAD = 180.0 / PI;
const double RAD_PER_DEG = PI / 180.0;
const double EARTH_RADIUS_M = 6370000.0;

// proj_info subset, module_map_utils.F90:150-192, filled by map_set (:243-568)
struct Proj {
    int code = 0;
    double lat1 = -999.9, lon1 = -999.9, dx = -999.9, latinc = -999.9, loninc = -999.9, stdlon = -999.9;
    double truelat1 = -999.9, truelat2 = -999.9, hemi = 0.0, cone = -999.9, polei = -999.9, polej = -999.9;
    double rsw = -999.9, rebydx = -999.9, knowni = -999.9, knownj = -999.9, re_m = EARTH_RADIUS_M;
    int nxmin = 1, nxmax = 43200;
};

double wrap180(double lon) {  // map_set :419-433
    int iter = 0;
    while (std::fabs(lon) > 180.0 && iter < 10) {
        if (lon < -180.0) lon += 360.0;
        if (lon > 180.0) lon -= 360.0;
        ++iter;
    }
    return lon;
}

// lc_cone, module_map_utils.F90:1124-1157
double lc_cone(double truelat1, double truelat2) {
    if (std::fabs(truelat1 - truelat2) > 0.1) {
        double cone = std::log10(std::cos(truelat1 * RAD_PER_DEG)) - std::log10(std::cos(truelat2 * RAD_PER_DEG));
        cone = cone / (std::log10(std::tan((45.0 - std::fabs(truelat1) / 2.0) * RAD_PER_DEG)) -
                       std::log10(std::tan((45.0 - std::fabs(truelat2) / 2.0) * RAD_PER_DEG)));
        return cone;
    }
    return std::sin(std::fabs(truelat1) * RAD_PER_DEG);
}

// set_lc, module_map_utils.F90:1083-1121
void set_lc(Proj &p) {
    p.cone = lc_cone(p.truelat1, p.truelat2);
    double deltalon1 = p.lon1 - p.stdlon;
    if (deltalon1 > 180.0) deltalon1 -= 360.0;
    if (deltalon1 < -180.0) deltalon1 += 360.0;
    double tl1r = p.truelat1 * RAD_PER_DEG;
    double ctl1r = std::cos(tl1r);
    p.rsw = p.rebydx * ctl1r / p.cone *
            std::pow(std::tan((90.0 * p.hemi - p.lat1) * RAD_PER_DEG / 2.0) /
                         std::tan((90.0 * p.hemi - p.truelat1) * RAD_PER_DEG / 2.0),
                     p.cone);
    double arg = p.cone * (deltalon1 * RAD_PER_DEG);
    p.polei = p.hemi * p.knowni - p.hemi * p.rsw * std::sin(arg);
    p.polej = p.hemi * p.knownj + p.rsw * std::cos(arg);
}

// ijll_lc, module_map_utils.F90:1160-1233
void ijll_lc(double i, double j, const Proj &p, double *lat, double *lon) {
    double chi1 = (90.0 - p.hemi * p.truelat1) * RAD_PER_DEG;
    double chi2 = (90.0 - p.hemi * p.truelat2) * RAD_PER_DEG;
    double inew = p.hemi * i, jnew = p.hemi * j;
    double xx = inew - p.polei, yy = p.polej - jnew;
    double r2 = xx * xx + yy * yy;
    double r = std::sqrt(r2) / p.rebydx;
    if (r2 == 0.0) {
        *lat = p.hemi * 90.0;
        *lon = p.stdlon;
    } else {
        double lo = p.stdlon + DEG_PER_RAD * std::atan2(p.hemi * xx, yy) / p.cone;
        lo = std::fmod(lo + 360.0, 360.0);
        double chi;
        if (chi1 == chi2) chi = 2.0 * std::atan(std::pow(r / std::tan(chi1), 1.0 / p.cone) * std::tan(chi1 * 0.5));
        else chi = 2.0 * std::atan(std::pow(r * p.cone / std::sin(chi1), 1.0 / p.cone) * std::tan(chi1 * 0.5));
        *lat = (90.0 - chi * DEG_PER_RAD) * p.hemi;
        *lon = lo;
    }
    if (*lon > 180.0) *lon -= 360.0;
    if (*lon < -180.0) *lon += 360.0;
}

// ijll_latlon, module_map_utils.F90:1398-1428 (longitudes are NOT wrapped)
void ijll_latlon(double i, double j, const Proj &p, double *lat, double *lon) {
    double i_work = i;
    if (i < (double)p.nxmin - 0.5) i_work = i + (double)(p.nxmax - p.nxmin + 1);
    if (i >= (double)p.nxmax + 0.5) i_work = i - (double)(p.nxmax - p.nxmin + 1);
    i_work = i_work - p.knowni;
    double j_work = j - p.knownj;
    *lat = p.lat1 + j_work * p.latinc;
    *lon = p.lon1 + i_work * p.loninc;
}

// push_source_projection + map_set, llxy_module.F90:38-160, module_map_utils.F90:243-568
bool make_proj(const mpassit_config *c, Proj &p, std::string &why) {
    p.code = c->proj_code;
    if (c->proj_code == MPASSIT_PROJ_LATLON) {
        p.lat1 = c->known_lat;
        p.lon1 = wrap180(c->known_lon);
        p.knowni = c->known_x;
        p.knownj = c->known_y;
        p.nxmax = (int)std::lround(360.0 / c->dlondeg);
        p.latinc = c->dlatdeg;
        p.loninc = c->dlondeg;
        return true;
    }
    if (c->proj_code == MPASSIT_PROJ_LC) {
        if (!(c->dxkm > 0.0)) { why = "Require grid spacing (dx) in meters be positive!"; return false; }
        if (std::fabs(c->truelat1) > 90.0) { why = "Set true latitude 1 for all projections!"; return false; }
        p.truelat1 = c->truelat1;
        p.truelat2 = c->truelat2;
        p.stdlon = wrap180(c->stand_lon);
        p.lat1 = c->known_lat;
        p.lon1 = wrap180(c->known_lon);
        p.knowni = c->known_x;
        p.knownj = c->known_y;
        p.dx = c->dxkm;
        p.hemi = c->truelat1 < 0.0 ? -1.0 : 1.0;
        p.rebydx = p.re_m / p.dx;
        if (std::fabs(p.truelat2) > 90.0) p.truelat2 = p.truelat1;
        set_lc(p);
        return true;
    }
    why = "target projection not supported by the host mirror (only 'lambert' and 'lat-lon'; "
          "the reference's presentation lists mercator/polar as untested)";
    return false;
}

}  // namespace

extern "C" {

int mpassit_target_dims(const mpassit_config *cfg, int stagger, int32_t *ni, int32_t *nj) {
    if (!cfg || !ni || !nj) return 1;
    // interp.F90:477-520 / model_grid.F90:707-728: CENTER i_target x j_target, EDGE1 +1 in i, EDGE2 +1 in j
    *ni = cfg->i_target + ((stagger == MPRG_EDGE1 || stagger == MPRG_CORNER) ? 1 : 0);
    *nj = cfg->j_target + ((stagger == MPRG_EDGE2 || stagger == MPRG_CORNER) ? 1 : 0);
    return (stagger < 0 || stagger > 3) ? 2 : 0;
}

int mpassit_projection(const mpassit_config *cfg, mprg_projection *out, char *err, size_t errlen) {
    // push_source_projection + map_set (+ set_lc): the per-grid scalars; the O(nx x ny) part can then run on the device
    // (mprg_set_target_projected)
    if (!cfg || !out) return 1;
    Proj p;
    std::string why;
    if (!make_proj(cfg, p, why)) {
        if (err && errlen) std::snprintf(err, errlen, "%s", why.c_str());
        return 2;
    }
    out->code = p.code == MPASSIT_PROJ_LC ? 1 : 0;
    out->nxmin = p.nxmin; out->nxmax = p.nxmax;
    out->lat1 = p.lat1; out->lon1 = p.lon1; out->knowni = p.knowni; out->knownj = p.knownj;
    out->latinc = p.latinc; out->loninc = p.loninc;
    out->stdlon = p.stdlon; out->truelat1 = p.truelat1; out->truelat2 = p.truelat2; out->hemi = p.hemi;
    out->cone = p.cone; out->polei = p.polei; out->polej = p.polej; out->rebydx = p.rebydx;
    return 0;
}

int mpassit_target_coords(const mpassit_config *cfg, int stagger, double *lat, double *lon, char *err, size_t errlen) {
    int32_t ni, nj;
    if (mpassit_target_dims(cfg, stagger, &ni, &nj) || !lat || !lon) return 1;
    Proj p;
    std::string why;
    if (!make_proj(cfg, p, why)) {
        if (err && errlen) std::snprintf(err, errlen, "%s", why.c_str());
        return 2;
    }
    // xytoll stagger offsets, llxy_module.F90:182-203:  U: x-0.5   V: y-0.5   CORNER: both
    const double ox = (stagger == MPRG_EDGE1 || stagger == MPRG_CORNER) ? 0.5 : 0.0;
    const double oy = (stagger == MPRG_EDGE2 || stagger == MPRG_CORNER) ? 0.5 : 0.0;
    // rows are independent: a few host threads (1.9 M points x 4 staggers of inverse-Lambert trig on the CONUS grid)
    par::range(nj, 16, [&](int64_t jb, int64_t je) {
    for (int32_t j = (int32_t)jb + 1; j <= (int32_t)je; ++j) {
        for (int32_t i = 1; i <= ni; ++i) {
            // get_lat_lon_fields, model_grid.F90:2212-2217 with rx = ry = 1
            double x = ((double)i - 0.5) / 1.0 + 0.5, y = ((double)j - 0.5) / 1.0 + 0.5;
            double rx = x - ox, ry = y - oy;
            size_t o = (size_t)(j - 1) * ni + (i - 1);
            if (p.code == MPASSIT_PROJ_LC) ijll_lc(rx, ry, p, &lat[o], &lon[o]);
            else ijll_latlon(rx, ry, p, &lat[o], &lon[o]);
        }
    }
    });
    return 0;
}

int mpassit_xytoll(const mpassit_config *cfg, double x, double y, int stagger, double *lat, double *lon) {
    // xytoll, llxy_module.F90:166-216 (stagger: MPASSIT_M / _U / _V / _CORNER)
    if (!cfg || !lat || !lon) return 1;
    Proj p;
    std::string why;
    if (!make_proj(cfg, p, why)) return 2;
    const double rx = x - ((stagger == MPASSIT_U || stagger == MPASSIT_CORNER) ? 0.5 : 0.0);
    const double ry = y - ((stagger == MPASSIT_V || stagger == MPASSIT_CORNER) ? 0.5 : 0.0);
    if (p.code == MPASSIT_PROJ_LC) ijll_lc(rx, ry, p, lat, lon);
    else ijll_latlon(rx, ry, p, lat, lon);
    return 0;
}

int mpassit_get_map_factor(const mpassit_config *cfg, const double *xlat, int64_t n, double *mapfac) {
    // get_map_factor, model_grid.F90:2229-2365 (Saucier pp. 32-33); mapfac_x == mapfac_y for Lambert.
    // For 'lat-lon' targets no branch of the reference assigns the array (it is written uninitialised);
    // this mirror writes 0 there.
    if (!cfg || !xlat || !mapfac) return 1;
    if (cfg->proj_code == MPASSIT_PROJ_LC) {
        if (cfg->truelat1 != cfg->truelat2) {
            const double colat1 = RAD_PER_DEG * (90.0 - cfg->truelat1), colat2 = RAD_PER_DEG * (90.0 - cfg->truelat2);
            const double nn = (std::log(std::sin(colat1)) - std::log(std::sin(colat2))) /
                              (std::log(std::tan(colat1 / 2.0)) - std::log(std::tan(colat2 / 2.0)));
            par::range(n, 1 << 15, [&](int64_t kb, int64_t ke) {
                for (int64_t k = kb; k < ke; ++k) {
                    const double colat = RAD_PER_DEG * (90.0 - xlat[k]);
                    mapfac[k] = std::sin(colat2) / std::sin(colat) * std::pow(std::tan(colat / 2.0) / std::tan(colat2 / 2.0), nn);
                }
            });
        } else {
            const double colat0 = RAD_PER_DEG * (90.0 - cfg->truelat1);
            par::range(n, 1 << 15, [&](int64_t kb, int64_t ke) {
                for (int64_t k = kb; k < ke; ++k) {
                    const double colat = RAD_PER_DEG * (90.0 - xlat[k]);
                    mapfac[k] = std::sin(colat0) / std::sin(colat) * std::pow(std::tan(colat / 2.0) / std::tan(colat0 / 2.0), std::cos(colat0));
                }
            });
        }
        return 0;
    }
    for (int64_t k = 0; k < n; ++k) mapfac[k] = 0.0;
    return 0;
}

void mpassit_get_cell_corners(const double *lat, const double *lon, int32_t ni, int32_t nj, double dx, double *clat,
                              double *clon) {
    // get_cell_corners, model_grid.F90:1902-1972, as written: every interior corner (i, j) is the point at bearing
    // 135 degrees (clockwise from north, i.e. SOUTH-EAST despite the `_sw` names) and distance sqrt(dx^2 / 2) from
    // centre (i, j); the last column uses bearing 225 from centre (ni, j), the last row bearing 45 from centre (i, nj),
    // the last corner bearing 315 from centre (ni, nj).  pi is the truncated literal 3.14159265359 (R8 under -r8),
    // R = 6370000 m.  lat/lon: [nj][ni]; clat/clon: [nj+1][ni+1].
    const double pi = 3.14159265359, R = 6370000.0;
    const double d = std::sqrt((dx * dx) / 2.0);
    auto step = [&](double lat_deg, double lon_deg, double brng_deg, double *olat, double *olon) {
        const double lat1 = lat_deg * (pi / 180.0), lon1 = lon_deg * (pi / 180.0), brng = brng_deg * pi / 180.0;
        const double lat2 = std::asin(std::sin(lat1) * std::cos(d / R) + std::cos(lat1) * std::sin(d / R) * std::cos(brng));
        const double lon2 = lon1 + std::atan2(std::sin(brng) * std::sin(d / R) * std::cos(lat1), std::cos(d / R) - std::sin(lat1) * std::sin(lat2));
        *olat = lat2 * 180.0 / pi;
        *olon = lon2 * 180.0 / pi;
    };
    const int32_t nic = ni + 1;
    par::range(nj + 1, 16, [&](int64_t jb, int64_t je) {
        for (int32_t j = (int32_t)jb; j < (int32_t)je; ++j)
            for (int32_t i = 0; i <= ni; ++i) {
                double *olat = &clat[(size_t)j * nic + i], *olon = &clon[(size_t)j * nic + i];
                if (j == nj && i == ni) step(lat[(size_t)(nj - 1) * ni + ni - 1], lon[(size_t)(nj - 1) * ni + ni - 1], 315.0, olat, olon);
                else if (i == ni) step(lat[(size_t)j * ni + ni - 1], lon[(size_t)j * ni + ni - 1], 225.0, olat, olon);
                else if (j == nj) step(lat[(size_t)(nj - 1) * ni + i], lon[(size_t)(nj - 1) * ni + i], 45.0, olat, olon);
                else step(lat[(size_t)j * ni + i], lon[(size_t)j * ni + i], 135.0, olat, olon);
            }
    });
}

void mpassit_get_rotang(const double *xlat, const double *xlon, int32_t ni, int32_t nj, double *cosa, double *sina) {
    // get_rotang, model_grid.F90:2450-2507: centred in j, one-sided on the first/last row.
    // (The reference evaluates this per PET tile, so with >1 PET its one-sided rows sit
    // at tile edges; this mirror always uses the global grid == the 1-PET result.)
    auto at = [&](const double *a, int i, int j) { return a[(size_t)j * ni + i]; };
    auto rot = [&](int i, int j, double d_lon, double d_lat) {
        if (d_lon > 180.0) d_lon -= 360.0;
        else if (d_lon < -180.0) d_lon += 360.0;
        double alpha = std::atan2(-std::cos(at(xlat, i, j) * RAD_PER_DEG) * (d_lon * RAD_PER_DEG), d_lat * RAD_PER_DEG);
        sina[(size_t)j * ni + i] = std::sin(alpha);
        cosa[(size_t)j * ni + i] = std::cos(alpha);
    };
    par::range(ni, 64, [&](int64_t ib, int64_t ie) {
    for (int i = (int)ib; i < (int)ie; ++i) {
        for (int j = 1; j < nj - 1; ++j)
            rot(i, j, at(xlon, i, j + 1) - at(xlon, i, j - 1), at(xlat, i, j + 1) - at(xlat, i, j - 1));
        if (nj >= 2) {
            rot(i, 0, at(xlon, i, 1) - at(xlon, i, 0), at(xlat, i, 1) - at(xlat, i, 0));
            rot(i, nj - 1, at(xlon, i, nj - 1) - at(xlon, i, nj - 2), at(xlat, i, nj - 1) - at(xlat, i, nj - 2));
        }
    }
    });
}

}  // extern "C"
