"""numpy restatement of the WPS projection formulas the reference uses to lay out its
target grid (module_map_utils.F90: set_lc 1083-1121, lc_cone 1124-1157, ijll_lc 1160-1233,
ijll_latlon 1398-1428; llxy_module.F90 xytoll 166-216 stagger offsets).
TEST INFRASTRUCTURE ONLY (checks mpassit_b200/host/target_grid.cpp).  Parity unpinned.
"""
import numpy as np

PI = 3.141592653589793
RAD_PER_DEG = PI / 180.0
DEG_PER_RAD = 180.0 / PI
EARTH_RADIUS_M = 6370000.0


def lc_grid(nx, ny, dx, ref_lat, ref_lon, truelat1, truelat2, stand_lon, stagger="M", known_x=None, known_y=None):
    """(lat, lon) [nj][ni] for stagger in {M, U, V, CORNER}; nx, ny are the namelist (corner) counts."""
    it, jt = nx - 1, ny - 1
    kx = (it + 1) / 2.0 if known_x is None else known_x
    ky = (jt + 1) / 2.0 if known_y is None else known_y
    ni = it + (1 if stagger in ("U", "CORNER") else 0)
    nj = jt + (1 if stagger in ("V", "CORNER") else 0)
    hemi = -1.0 if truelat1 < 0 else 1.0
    rebydx = EARTH_RADIUS_M / dx
    if abs(truelat1 - truelat2) > 0.1:
        cone = (np.log10(np.cos(truelat1 * RAD_PER_DEG)) - np.log10(np.cos(truelat2 * RAD_PER_DEG))) / \
               (np.log10(np.tan((45.0 - abs(truelat1) / 2.0) * RAD_PER_DEG)) -
                np.log10(np.tan((45.0 - abs(truelat2) / 2.0) * RAD_PER_DEG)))
    else:
        cone = np.sin(abs(truelat1) * RAD_PER_DEG)
    dl = ref_lon - stand_lon
    if dl > 180:
        dl -= 360
    if dl < -180:
        dl += 360
    rsw = rebydx * np.cos(truelat1 * RAD_PER_DEG) / cone * \
        (np.tan((90.0 * hemi - ref_lat) * RAD_PER_DEG / 2.0) / np.tan((90.0 * hemi - truelat1) * RAD_PER_DEG / 2.0)) ** cone
    arg = cone * (dl * RAD_PER_DEG)
    polei = hemi * kx - hemi * rsw * np.sin(arg)
    polej = hemi * ky + rsw * np.cos(arg)
    ox = 0.5 if stagger in ("U", "CORNER") else 0.0
    oy = 0.5 if stagger in ("V", "CORNER") else 0.0
    I, J = np.meshgrid(np.arange(1, ni + 1, dtype=np.float64) - ox, np.arange(1, nj + 1, dtype=np.float64) - oy)
    chi1 = (90.0 - hemi * truelat1) * RAD_PER_DEG
    chi2 = (90.0 - hemi * truelat2) * RAD_PER_DEG
    xx = hemi * I - polei
    yy = polej - hemi * J
    r = np.sqrt(xx * xx + yy * yy) / rebydx
    lon = stand_lon + DEG_PER_RAD * np.arctan2(hemi * xx, yy) / cone
    lon = np.fmod(lon + 360.0, 360.0)
    if chi1 == chi2:
        chi = 2.0 * np.arctan((r / np.tan(chi1)) ** (1.0 / cone) * np.tan(chi1 * 0.5))
    else:
        chi = 2.0 * np.arctan((r * cone / np.sin(chi1)) ** (1.0 / cone) * np.tan(chi1 * 0.5))
    lat = (90.0 - chi * DEG_PER_RAD) * hemi
    lon = np.where(lon > 180.0, lon - 360.0, lon)
    lon = np.where(lon < -180.0, lon + 360.0, lon)
    return lat, lon


def latlon_global_grid(nx, ny, stand_lon, stagger="M"):
    """Global lat-lon (dx, dy unset): program_setup.F90:197-210 + ijll_latlon."""
    it, jt = nx - 1, ny - 1
    dlon, dlat = 360.0 / it, 180.0 / jt
    ni = it + (1 if stagger in ("U", "CORNER") else 0)
    nj = jt + (1 if stagger in ("V", "CORNER") else 0)
    ox = 0.5 if stagger in ("U", "CORNER") else 0.0
    oy = 0.5 if stagger in ("V", "CORNER") else 0.0
    I, J = np.meshgrid(np.arange(1, ni + 1, dtype=np.float64) - ox, np.arange(1, nj + 1, dtype=np.float64) - oy)
    nxmax = int(round(360.0 / dlon))
    Iw = np.where(I < 1 - 0.5, I + nxmax, I)
    Iw = np.where(I >= nxmax + 0.5, I - nxmax, Iw)
    lat = (-90.0 + dlat / 2.0) + (J - 1.0) * dlat
    lon = (stand_lon + dlon / 2.0) + (Iw - 1.0) * dlon
    return lat, lon


def rotang(lat, lon):
    """get_rotang, model_grid.F90:2450-2507."""
    nj, ni = lat.shape
    cosa, sina = np.empty_like(lat), np.empty_like(lat)

    def fix(d):
        d = np.where(d > 180.0, d - 360.0, d)
        return np.where(d < -180.0, d + 360.0, d)

    def put(j, dlon, dlat):
        a = np.arctan2(-np.cos(lat[j] * RAD_PER_DEG) * (fix(dlon) * RAD_PER_DEG), dlat * RAD_PER_DEG)
        sina[j], cosa[j] = np.sin(a), np.cos(a)

    for j in range(1, nj - 1):
        put(j, lon[j + 1] - lon[j - 1], lat[j + 1] - lat[j - 1])
    put(0, lon[1] - lon[0], lat[1] - lat[0])
    put(nj - 1, lon[nj - 1] - lon[nj - 2], lat[nj - 1] - lat[nj - 2])
    return cosa, sina
