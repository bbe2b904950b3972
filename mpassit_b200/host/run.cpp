// run.cpp -- host mirror of `program mpassit` (mpassit.F90:23-146) with files on both sides:
//   read_setup_namelist -> define_target_grid -> define_input_grid -> read_input_data -> interp_data -> write_to_file
// on top of the engine's C ABI and the classic-NetCDF container of ncio.cpp ("I/O adjacency").
//
// What differs from the reference by design:
//   * each MPAS variable is consumed where it lies in the mapped input file: [cell][level] level-fastest, big-endian.
//     No whole-variable read per PET (input_data.F90:645), no transpose (:653-655), no host byte swap: the engine
//     uploads the cell range its slab references and swaps it in HBM (mprg_set_source_byte_order);
//   * outputs stay in HBM for the writer's WRF post-ops (write_data.F90:1339-1432: T-300, Z_C, PHB = 9.81 zgrid,
//     P_TOP), are swapped to file order on the device and written by EVERY rank at the offsets of its own row slab
//     (pwrite) -- no ESMF_FieldGather to PET 0 (write_data.F90:1006-1453), no funnel through one writer;
//   * the container is NetCDF classic (CDF-2 / CDF-5) instead of NetCDF-4/HDF5 (write_data.F90:170): same data
//     model, dimension / variable / attribute names, definition order and values.
// Quirks of the reference's writer that are kept, because downstream tools see them:
//   DY is written from dx (:230); CEN_LAT / CEN_LON are the grid-centre values computed by xytoll
//   (model_grid.F90:1113); COSALPHA carries no attributes and SINALPHA carries the "COSINE ..." description (:436-442
//   put them on id_sina); Z_C is defined on bottom_top_stag but only nz levels are written (:443,1413: the top level
//   keeps the fill value); PB receives the P_HYD values (:1376); the `< 10` guard of T-300 is a no-op (:1341);
//   XTIME = (start - valid) minutes (:1198-1199).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mpassit_host.h"
#include "ncio.hpp"

namespace {

struct Fail {
    int rc;
    std::string msg;
};
[[noreturn]] void die(int rc, const std::string &m) { throw Fail{rc ? rc : 1, m}; }
void ck(mprg_ctx *ctx, int rc, const char *where) {
    if (rc != 0) die(rc, std::string("IN ") + where + ": " + mprg_last_error(ctx));
}
// netcdf_err, utils.F90:35-60: "FATAL ERROR: <string>: <nf90_strerror>"
[[noreturn]] void netcdf_err(const std::string &what, const std::string &why) { die(999, "FATAL ERROR: " + what + ": " + why); }

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

std::string rtrim(std::string s) {
    while (!s.empty() && (s.back() == ' ' || s.back() == '\0' || s.back() == '\n')) s.pop_back();
    return s;
}

struct VarList {
    int32_t n = 0;
    std::vector<char> a, b;
    const char *name(int i) const { return a.data() + (size_t)i * MPASSIT_NAMELEN; }
    const char *target(int i) const { return b.data() + (size_t)i * MPASSIT_NAMELEN; }
};
VarList read_list(const std::string &dir, const char *fname) {
    VarList v;
    const int32_t cap = 1024;
    v.a.assign((size_t)cap * MPASSIT_NAMELEN, 0);
    v.b.assign((size_t)cap * MPASSIT_NAMELEN, 0);
    char e[1024] = {0};
    const std::string path = dir.empty() ? std::string(fname) : dir + "/" + fname;
    int rc = mpassit_read_varlist(path.c_str(), cap, &v.n, v.a.data(), v.b.data(), e, sizeof e);
    if (rc) die(rc, e);
    return v;
}

// days since 1970-01-01 of a proleptic Gregorian date (datetime_module's date arithmetic, used at write_data.F90:1198)
int64_t days_from_civil(int64_t y, int m, int d) {
    y -= m <= 2;
    const int64_t era = (y >= 0 ? y : y - 399) / 400;
    const int64_t yoe = y - era * 400;
    const int64_t doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    const int64_t doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + doe - 719468;
}
// "YYYY-MM-DD_hh:mm:ss" -> seconds; substr(), write_data.F90:1535-1542
bool stamp_seconds(const std::string &s, int64_t *out) {
    if (s.size() < 19) return false;
    auto num = [&](int a, int b, int *v) {
        char *end = nullptr;
        const std::string t = s.substr(a, b - a);
        long x = std::strtol(t.c_str(), &end, 10);
        if (end == t.c_str()) return false;
        *v = (int)x;
        return true;
    };
    int y, mo, d, h, mi, se;
    if (!num(0, 4, &y) || !num(5, 7, &mo) || !num(8, 10, &d) || !num(11, 13, &h) || !num(14, 16, &mi) || !num(17, 19, &se))
        return false;
    *out = days_from_civil(y, mo, d) * 86400 + h * 3600 + mi * 60 + se;
    return true;
}

constexpr float kFillFloat = 9.9692099683868690e+36f;  // NC_FILL_FLOAT: what an unwritten NF90_FLOAT level reads as

// one regridded output variable: a device slab of this rank + where it goes in the file
struct OutVar {
    std::string name;
    int varid = -1, stagger = MPRG_CENTER;
    int32_t nlev = 1;
    void *dev = nullptr;  // [nlev][rows][ni] fp32 on the device
};

struct Source {
    const void *ptr = nullptr;  // big-endian [n][nlev] inside the mapped file
    std::string units, longname;
};

}  // namespace

extern "C" {

int mpassit_run(const char *namelist_file, const char *varlist_dir, int device, int rank, int nranks,
                mpassit_comm_fn comm, void *comm_arg, mpassit_run_stats *stats, char *err, size_t errlen) {
    mprg_ctx *ctx = nullptr;
    int rc_out = 0;
    mpassit_run_stats st;
    std::memset(&st, 0, sizeof st);
    std::vector<void *> dev_allocs;
    void *pinned[2] = {nullptr, nullptr};
    std::thread writer;  // background pwrite of the field downloaded last
    bool writer_ok = true;
    std::string writer_err;
    auto barrier = [&]() {
        if (nranks > 1 && comm) comm(comm_arg, MPASSIT_COMM_BARRIER, nullptr, 0);
    };
    try {
        if (!namelist_file) die(1, "namelist file - (null) does not exist.");
        if (nranks > 1 && !comm) die(1, "mpassit_run: more than one rank needs a communication callback");
        const double t0 = now_ms();
        char e[2048] = {0};
        mpassit_config cfg;
        int rc = mpassit_read_setup_namelist(namelist_file, &cfg, e, sizeof e);
        if (rc) die(rc, e);
        if (!cfg.interp_diag && !cfg.interp_hist)  // input_data.F90:110-112
            die(-1, " SET INTERP_DIAG AND/OR INTERP_HIST TO TRUE TO OBTAIN OUTPUT");
        const bool from_file = rtrim(cfg.target_grid_type) == "file";

        // device < 0: lay the output file out (header, grid description, times) without touching a GPU -- no field
        // is regridded or written; the regular run needs a device and fails without one (no CPU path)
        const bool dry = device < 0;
        double tq = now_ms();
        if (!dry) ck(nullptr, mprg_init(device, rank, nranks, &ctx), "mprg_init");
        // whatever MPASSIT_GPU_ASYNC says: the write stage hands every downloaded buffer to a writer thread as soon
        // as mprg_download returns, so downloads must be blocking here (interp_data switches asynchronous applies on
        // for its own host-buffer passes and restores this setting)
        if (!dry) ck(ctx, mprg_set_async(ctx, 0), "mprg_set_async");
        st.init_ms = now_ms() - tq;
        tq = now_ms();

        // ------------------------------------------------------------------ define_target_grid
        const int staggers[4] = {MPRG_CENTER, MPRG_EDGE1, MPRG_EDGE2, MPRG_CORNER};
        int32_t ni[4], nj[4];
        std::vector<double> lat[4], lon[4], mapfac[3], cosa, sina, hgt_file;
        double cen_lat = cfg.ref_lat, cen_lon = cfg.ref_lon;
        std::string why;
        // The host-only part of define_target_grid (coordinate tables, map factors, rotation angles, or the target
        // file) runs on its own thread beside define_input_grid; the engine calls follow once both are done.
        auto define_target = [&]() {
            std::string twhy;
            char te[2048] = {0};
            if (from_file) {
                // define_target_grid_file, model_grid.F90:1203-1890: sizes, projection attributes, the three staggers, map
                // factors, rotation angles and terrain come from a WRF-style file; the corners are synthesised
                ncio::Reader tf;
                if (!tf.open(cfg.file_target_grid, twhy)) netcdf_err(std::string("opening: ") + cfg.file_target_grid, twhy);
                auto tdim = [&](const char *name) -> int32_t {
                    const ncio::Dim *d = tf.dim(name);
                    if (!d) netcdf_err(std::string("reading ") + name + " id", "NetCDF: Invalid dimension ID or name");
                    return (int32_t)d->len;
                };
                auto gnum = [&](const char *name, double *out, bool required) {
                    const ncio::Att *a = tf.gatt(name);
                    if (!a) {
                        if (required) netcdf_err(std::string("reading ") + name, "NetCDF: Attribute not found");
                        return;
                    }
                    *out = a->number();
                };
                cfg.i_target = tdim("west_east");
                cfg.j_target = tdim("south_north");
                gnum("DX", &cfg.dx, true);
                double pc = cfg.proj_code;
                gnum("CEN_LAT", &cfg.ref_lat, true);
                gnum("CEN_LON", &cfg.ref_lon, true);
                gnum("TRUELAT1", &cfg.truelat1, true);
                gnum("TRUELAT2", &cfg.truelat2, true);
                gnum("MOAD_CEN_LAT", &cfg.ref_lat, true);  // overrides CEN_LAT, model_grid.F90:1273
                gnum("STAND_LON", &cfg.stand_lon, true);
                gnum("POLE_LAT", &cfg.pole_lat, true);
                gnum("POLE_LON", &cfg.pole_lon, true);
                gnum("MAP_PROJ", &pc, true);
                cfg.proj_code = (int32_t)pc;
                if (const ncio::Att *a = tf.gatt("MAP_PROJ_CHAR")) std::snprintf(cfg.map_proj_char, MPASSIT_NAMELEN, "%s", a->text().c_str());
                else std::snprintf(cfg.map_proj_char, MPASSIT_NAMELEN, "%s", cfg.proj_code == 1 ? "Lambert Conformal" : "Lat/Lon");
                cen_lat = cfg.ref_lat;
                cen_lon = cfg.ref_lon;
                for (int s = 0; s < 4; ++s) {
                    mpassit_target_dims(&cfg, staggers[s], &ni[s], &nj[s]);
                    lat[s].resize((size_t)ni[s] * nj[s]);
                    lon[s].resize((size_t)ni[s] * nj[s]);
                }
                auto tvar = [&](const char *name, const char *alt, size_t n, double *out) {
                    const ncio::Var *v = tf.var(name);
                    if (!v && alt) v = tf.var(alt);
                    if (!v) netcdf_err(std::string("reading ") + name + " id", "NetCDF: Variable not found");
                    if (tf.count(*v) != n) netcdf_err(std::string("reading ") + name, "NetCDF: Start+count exceeds dimension bound");
                    if (!tf.read_doubles(*v, 0, n, out, twhy)) netcdf_err(std::string("reading ") + name, twhy);
                };
                tvar("XLONG", "XLONG_M", lon[0].size(), lon[0].data());
                tvar("XLAT", "XLAT_M", lat[0].size(), lat[0].data());
                tvar("XLONG_U", nullptr, lon[1].size(), lon[1].data());
                tvar("XLAT_U", nullptr, lat[1].size(), lat[1].data());
                tvar("XLONG_V", nullptr, lon[2].size(), lon[2].data());
                tvar("XLAT_V", nullptr, lat[2].size(), lat[2].data());
                mpassit_get_cell_corners(lat[0].data(), lon[0].data(), ni[0], nj[0], cfg.dx, lat[3].data(), lon[3].data());
                const char *mf[3] = {"MAPFAC_M", "MAPFAC_U", "MAPFAC_V"};
                for (int s = 0; s < 3; ++s) {
                    mapfac[s].resize(lat[s].size());
                    tvar(mf[s], nullptr, mapfac[s].size(), mapfac[s].data());
                }
                if (cfg.proj_code == MPASSIT_PROJ_LC) {
                    sina.resize(lat[0].size());
                    cosa.resize(lat[0].size());
                    tvar("SINALPHA", nullptr, sina.size(), sina.data());
                    tvar("COSALPHA", nullptr, cosa.size(), cosa.data());
                }
                hgt_file.resize(lat[0].size());
                tvar("HGT", "HGT_M", hgt_file.size(), hgt_file.data());
            } else {
                for (int s = 0; s < 4; ++s) {
                    mpassit_target_dims(&cfg, staggers[s], &ni[s], &nj[s]);
                    lat[s].resize((size_t)ni[s] * nj[s]);
                    lon[s].resize((size_t)ni[s] * nj[s]);
                    const int trc = mpassit_target_coords(&cfg, staggers[s], lat[s].data(), lon[s].data(), te, sizeof te);
                    if (trc) die(trc, te);
                }
                for (int s = 0; s < 3; ++s) {  // get_map_factor on the M / U / V latitudes, model_grid.F90:1022-1036
                    mapfac[s].resize(lat[s].size());
                    mpassit_get_map_factor(&cfg, lat[s].data(), (int64_t)lat[s].size(), mapfac[s].data());
                }
                // model_grid.F90:1113: ref_lat / ref_lon become the grid-centre coordinates (they end up in CEN_LAT / CEN_LON)
                mpassit_xytoll(&cfg, cfg.i_target / 2.0, cfg.j_target / 2.0, MPASSIT_M, &cen_lat, &cen_lon);
                if (cfg.proj_code == MPASSIT_PROJ_LC) {  // model_grid.F90:1114-1185
                    cosa.resize(lat[0].size());
                    sina.resize(lat[0].size());
                    mpassit_get_rotang(lat[0].data(), lon[0].data(), cfg.i_target, cfg.j_target, cosa.data(), sina.data());
                }
            }
        };
        std::future<void> target_job = std::async(std::launch::async, define_target);
        st.target_ms = now_ms() - tq;
        tq = now_ms();

        // ------------------------------------------------------------------ define_input_grid, model_grid.F90:263-640
        ncio::Reader grid;
        if (!grid.open(cfg.grid_file_input_grid, why)) netcdf_err(std::string("opening: ") + cfg.grid_file_input_grid, why);
        auto dimlen = [&](const ncio::Reader &r, const char *name, const char *msg) -> int64_t {
            const ncio::Dim *d = r.dim(name);
            if (!d) netcdf_err(msg, "NetCDF: Invalid dimension ID or name");
            return d->len ? (int64_t)d->len : (int64_t)r.numrecs;
        };
        auto getvar = [&](const ncio::Reader &r, const std::string &name, const char *msg) -> const ncio::Var & {
            const ncio::Var *v = r.var(name);
            if (!v) netcdf_err(msg, "NetCDF: Variable not found");
            return *v;
        };
        const int64_t nCells = dimlen(grid, "nCells", "reading nCells id");
        const int64_t nVertices = dimlen(grid, "nVertices", "reading nVertices id");
        const int32_t nz = (int32_t)dimlen(grid, "nVertLevels", "reading nVertLevels id");
        const int32_t nzp1 = (int32_t)dimlen(grid, "nVertLevelsP1", "reading nVertLevelsP1 id");
        const int32_t maxEdges = (int32_t)dimlen(grid, "maxEdges", "reading maxEdges id");
        const int32_t nsoil = (int32_t)dimlen(grid, "nSoilLevels", "reading nSoilLevels id");
        std::vector<double> lonCell(nCells), latCell(nCells), lonVert(nVertices), latVert(nVertices), zs(nsoil), ter(nCells);
        std::vector<int32_t> voc((size_t)nCells * maxEdges);
        auto need_d = [&](const char *name, const char *msg, double *out, uint64_t n) {
            if (!grid.read_doubles(getvar(grid, name, msg), 0, n, out, why)) netcdf_err(msg, why);
        };
        need_d("lonCell", "reading lonCell", lonCell.data(), nCells);
        need_d("latCell", "reading latCell", latCell.data(), nCells);
        need_d("lonVertex", "reading lonVertex", lonVert.data(), nVertices);
        need_d("latVertex", "reading latVertex", latVert.data(), nVertices);
        need_d("zs", "reading ZS", zs.data(), nsoil);  // start=(1,1), count=(nsoil,1): the first cell's depths
        need_d("ter", "reading ter", ter.data(), nCells);
        if (!grid.read_ints(getvar(grid, "verticesOnCell", "reading verticesOnCell id"), 0, voc.size(), voc.data(), why))
            netcdf_err("reading verticesOnCell", why);
        st.gridfile_ms = now_ms() - tq;
        tq = now_ms();
        if (!dry)
            ck(ctx, mprg_set_mesh(ctx, (int32_t)nCells, (int32_t)nVertices, maxEdges, lonCell.data(), latCell.data(),
                                  lonVert.data(), latVert.data(), voc.data()), "MeshCreate");
        st.mesh_ms = now_ms() - tq;
        st.n_cells = nCells;
        tq = now_ms();
        target_job.get();  // rethrows what define_target threw
        if (!dry)  // GridCreate1PeriDim (global) or GridCreateNoPeriDim (regional), model_grid.F90:684-703
            ck(ctx, mprg_set_grid_kind(ctx, cfg.is_regional ? MPRG_GRID_NOPERI : MPRG_GRID_1PERI_MONOPOLE), "GridCreate");
        if (!dry)
            for (int s = 0; s < 4; ++s)
                ck(ctx, mprg_set_target(ctx, staggers[s], ni[s], nj[s], lon[s].data(), lat[s].data()), "GridAddCoord");
        const int32_t it = cfg.i_target, jt = cfg.j_target;
        const bool lc = cfg.proj_code == MPASSIT_PROJ_LC;
        if (lc && !dry) ck(ctx, mprg_set_rotation(ctx, cosa.data(), sina.data()), "get_rotang");
        st.target_ms += now_ms() - tq;  // what the main thread still waited for, plus the uploads
        st.setup_ms = now_ms() - t0;

        // ------------------------------------------------------------------ read_input_data, input_data.F90:97-812
        const double t1 = now_ms();
        const std::string ldir = varlist_dir ? varlist_dir : "";
        VarList ldiag, l2d, l3d, lsoil;
        if (cfg.interp_diag) ldiag = read_list(ldir, "diaglist");  // input_data.F90:276
        if (cfg.interp_hist) {                                      // input_data.F90:851-855
            l2d = read_list(ldir, "histlist_2d");
            l3d = read_list(ldir, "histlist_3d");
            lsoil = read_list(ldir, "histlist_soil");
        }
        ncio::Reader fdiag, fhist;
        std::string start_time, valid_time;
        double config_dt = 0.0;
        int32_t diag_out_interval = 0, lsm_scheme = 0, mp_scheme = 0, conv_scheme = 0;
        int src_type = 0;  // NC_FLOAT or NC_DOUBLE, the same for every regridded variable
        auto source = [&](const ncio::Reader &r, const char *name, int64_t n, int32_t nlev) {
            const ncio::Var &v = getvar(r, name, (std::string("reading field id - ") + name).c_str());
            if (v.type != ncio::NC_FLOAT && v.type != ncio::NC_DOUBLE) netcdf_err(std::string("reading field ") + name, "not a real variable");
            if (src_type && v.type != src_type)
                netcdf_err(std::string("reading field ") + name, "single and double precision variables cannot be mixed in one run");
            src_type = v.type;
            if ((int64_t)r.count(v) != n * nlev)
                netcdf_err(std::string("reading field ") + name, "NetCDF: Start+count exceeds dimension bound");
            Source s;
            s.ptr = r.data(v, 0);
            if (!s.ptr) netcdf_err(std::string("reading field ") + name, "no record in file");
            const ncio::Att *u = v.att("units"), *ln = v.att("long_name");
            if (!u) netcdf_err("reading field units", "NetCDF: Attribute not found");
            if (!ln) netcdf_err("reading field long_name", "NetCDF: Attribute not found");
            s.units = u->text();
            s.longname = ln->text();
            st.bytes_in += (int64_t)r.count(v) * (int64_t)ncio::type_size(v.type);
            return s;
        };
        auto read_xtime = [&](const ncio::Reader &r) {  // input_data.F90:243-253 / 394-405
            if (!r.dim("StrLen")) netcdf_err("reading strlen dim id", "NetCDF: Invalid dimension ID or name");
            const ncio::Var &v = getvar(r, "xtime", "reading xtime id");
            const uint8_t *p = r.data(v, 0);
            if (!p) netcdf_err("getting xtime", "no record in file");
            valid_time = rtrim(std::string((const char *)p, r.count(v)));
        };
        std::vector<mpassit_field> fd, f2, f3, fs;
        std::vector<Source> sd, s2, s3, ss;
        auto mkfield = [](const VarList &l, int i, int32_t nlev) {
            mpassit_field f;
            std::memset(&f, 0, sizeof f);
            std::snprintf(f.name, MPASSIT_NAMELEN, "%s", l.name(i));
            std::snprintf(f.target_name, MPASSIT_NAMELEN, "%s", l.target(i));
            f.nlev = nlev;
            return f;
        };
        if (cfg.interp_diag) {
            if (!fdiag.open(cfg.diag_file_input_grid, why)) netcdf_err(std::string("opening: ") + cfg.diag_file_input_grid, why);
            for (int i = 0; i < ldiag.n; ++i) {
                // the reference takes the level count from the file variable's rank (input_data.F90:185-212)
                const ncio::Var &v = getvar(fdiag, ldiag.name(i), (std::string("reading field id - ") + ldiag.name(i)).c_str());
                const int32_t nlev = fdiag.shape(v).size() >= 2 ? nz : 1;
                if ((nlev > 1) != (mpassit_classify_diag(ldiag.name(i)) == MPASSIT_CLASS_DIAG_3D))
                    die(3, std::string("IN FieldGet: diag variable ") + ldiag.name(i) + " does not have the rank its field was created with");
                fd.push_back(mkfield(ldiag, i, nlev));
                sd.push_back(source(fdiag, ldiag.name(i), nCells, nlev));
            }
            if (const ncio::Att *a = fdiag.gatt("config_start_time")) start_time = rtrim(a->text());
            else if (!cfg.interp_hist) netcdf_err("reading config_start_time", "NetCDF: Attribute not found");
            if (const ncio::Att *a = fdiag.gatt("config_dt")) config_dt = a->number();
            if (const ncio::Att *a = fdiag.gatt("output_interval")) diag_out_interval = (int32_t)a->number();
            read_xtime(fdiag);
        }
        if (cfg.interp_hist) {
            if (!fhist.open(cfg.hist_file_input_grid, why)) netcdf_err(std::string("opening: ") + cfg.hist_file_input_grid, why);
            if (const ncio::Att *a = fhist.gatt("config_lsm_scheme")) {  // input_data.F90:343-350
                const std::string t = rtrim(a->text());
                lsm_scheme = t == "noah" ? 2 : t == "ruc" ? 3 : 0;
            }
            if (const ncio::Att *a = fhist.gatt("config_start_time")) start_time = rtrim(a->text());
            else netcdf_err("reading config_start_time", "NetCDF: Attribute not found");
            if (const ncio::Att *a = fhist.gatt("config_microp_scheme")) {  // :357-364
                const std::string t = rtrim(a->text());
                mp_scheme = t == "mp_thompson" ? 8 : t == "mp_nssl2m" ? 18 : 0;
            }
            if (const ncio::Att *a = fhist.gatt("config_convection_scheme")) {  // :368-377
                const std::string t = rtrim(a->text());
                conv_scheme = t == "cu_ntiedke" ? 16 : t == "cu_kain_fritsch" ? 1 : t == "cu_grell_freitas" ? 3 : 0;
            }
            if (const ncio::Att *a = fhist.gatt("config_dt")) config_dt = a->number();
            else config_dt = 0.0;
            read_xtime(fhist);
            for (int i = 0; i < l2d.n; ++i) {
                f2.push_back(mkfield(l2d, i, 1));
                s2.push_back(source(fhist, l2d.name(i), nCells, 1));
            }
            for (int i = 0; i < lsoil.n; ++i) {
                fs.push_back(mkfield(lsoil, i, nsoil));
                ss.push_back(source(fhist, lsoil.name(i), nCells, nsoil));
            }
            for (int i = 0; i < l3d.n; ++i) {
                const int k = mpassit_classify_hist_3d(l3d.name(i), cfg.wrf_mod_vars);
                const int32_t nlev = k == MPASSIT_CLASS_3D_NZP1 ? nzp1 : nz;
                f3.push_back(mkfield(l3d, i, nlev));
                s3.push_back(source(fhist, l3d.name(i), k == MPASSIT_CLASS_3D_VERT ? nVertices : nCells, nlev));
            }
        }
        if (src_type == 0) src_type = ncio::NC_FLOAT;
        const int sdt = src_type == ncio::NC_DOUBLE ? MPRG_F64 : MPRG_F32;
        // hgt_input_grid (model_grid.F90:602-616): `ter` of the grid file, in the precision and byte order of the
        // other sources (nCells values: converted on the host)
        std::vector<uint8_t> ter_be((size_t)nCells * (sdt == MPRG_F64 ? 8 : 4));
        if (sdt == MPRG_F64) {
            std::memcpy(ter_be.data(), ter.data(), ter_be.size());
            ncio::to_big_endian(ter_be.data(), 8, nCells);
        } else {
            float *f = (float *)ter_be.data();
            for (int64_t k = 0; k < nCells; ++k) f[k] = (float)ter[k];
            ncio::to_big_endian(ter_be.data(), 4, nCells);
        }
        st.read_ms = now_ms() - t1;

        // ------------------------------------------------------------------ outputs: device slabs of this rank
        mpassit_interp_io io;
        std::memset(&io, 0, sizeof io);
        io.src_dtype = sdt;
        io.dst_dtype = MPRG_F32;  // the file holds NF90_FLOAT (write_data.F90:587)
        io.mem = MPRG_HOST;
        io.dst_device = 1;
        io.nz = nz;
        io.n_diag = (int32_t)fd.size(); io.diag = fd.data();
        io.n_hist_2d = (int32_t)f2.size(); io.hist_2d = f2.data();
        io.n_hist_3d = (int32_t)f3.size(); io.hist_3d = f3.data();
        io.n_soil = (int32_t)fs.size(); io.soil = fs.data();
        int32_t do_u = 0, do_v = 0, u10 = -1, v10 = -1;
        mpassit_classify_fields(&cfg, &io, &do_u, &do_v, &u10, &v10);
        int32_t j0[3], j1[3];
        for (int s = 0; s < 3; ++s) {
            j0[s] = j1[s] = 0;
            if (!dry) ck(ctx, mprg_get_slab(ctx, staggers[s], &j0[s], &j1[s]), "get_slab");
        }
        // One device arena for every output slab (plus Z_C's scratch): a single allocation instead of ~45 -- on a
        // shared host each cudaMalloc can stall behind other tenants' driver calls.  Pass 0 sizes it, pass 1 carves it.
        unsigned char *arena = nullptr;
        size_t arena_used = 0;
        auto dalloc = [&](int s, int32_t nlev) -> void * {
            const size_t bytes = (std::max<size_t>((size_t)(j1[s] - j0[s]) * ni[s] * nlev * 4, 4) + 255) & ~(size_t)255;
            void *p = arena ? arena + arena_used : nullptr;
            arena_used += bytes;
            return p;
        };
        void *d_mid = nullptr;
        for (int pass = 0; pass < 2; ++pass) {
            arena_used = 0;
            for (size_t i = 0; i < fd.size(); ++i) { fd[i].src = sd[i].ptr; fd[i].dst = dalloc(0, fd[i].nlev); }
            for (size_t i = 0; i < f2.size(); ++i) { f2[i].src = s2[i].ptr; f2[i].dst = dalloc(0, 1); }
            for (size_t i = 0; i < fs.size(); ++i) { fs[i].src = ss[i].ptr; fs[i].dst = dalloc(0, fs[i].nlev); }
            for (size_t i = 0; i < f3.size(); ++i) {
                f3[i].src = s3[i].ptr;
                // U / V of the wrf_mod_vars wind chain leave through u_stag / v_stag (interp.F90:295-328)
                if (f3[i].klass != MPASSIT_CLASS_U && f3[i].klass != MPASSIT_CLASS_V) f3[i].dst = dalloc(0, f3[i].nlev);
            }
            if (cfg.interp_hist) {
                io.ter = ter_be.data();
                io.hgt = dalloc(0, 1);
                if (do_u) io.u_stag = dalloc(1, nz);
                if (do_v) io.v_stag = dalloc(2, nz);
                d_mid = dalloc(0, nzp1 - 1);  // Z_C (write_data.F90:1406-1413)
            }
            if (pass == 0 && !dry) {
                void *p = nullptr;
                ck(ctx, mprg_device_alloc(ctx, arena_used, &p), "FieldCreate");
                dev_allocs.push_back(p);
                arena = (unsigned char *)p;
            }
        }

        // ------------------------------------------------------------------ interp_data
        const double t2 = now_ms();
        st.alloc_ms = t2 - (t1 + st.read_ms);
        if (!dry) {
            ck(ctx, mprg_set_source_byte_order(ctx, 1), "set_source_byte_order");
            rc = mpassit_interp_data(ctx, &cfg, &io, e, sizeof e);
            mprg_set_source_byte_order(ctx, 0);
            if (rc) die(rc, e);
            ck(ctx, mprg_synchronize(ctx), "interp_data");
        }
        st.interp_ms = now_ms() - t2;

        // ------------------------------------------------------------------ write_to_file, write_data.F90:96-1496
        const double t3 = now_ms();
        ncio::Writer w(0);
        struct JoinOnExit {  // declared after `w`: the background writer is joined before the Writer goes away
            std::thread &t;
            ~JoinOnExit() {
                if (t.joinable()) t.join();
            }
        } join_on_exit{writer};
        const int dTime = w.def_dim("Time", 0), dWE = w.def_dim("west_east", it), dWEs = w.def_dim("west_east_stag", it + 1),
                  dSN = w.def_dim("south_north", jt), dSNs = w.def_dim("south_north_stag", jt + 1),
                  dZ = w.def_dim("bottom_top", nz), dZs = w.def_dim("bottom_top_stag", nzp1),
                  dSoil = w.def_dim("soil_layers_stag", nsoil), dStr = w.def_dim("StrLen", 19);
        w.att_int(-1, "WEST-EAST_GRID_DIMENSION", it + 1);
        w.att_int(-1, "SOUTH-NORTH_GRID_DIMENSION", jt + 1);
        w.att_int(-1, "BOTTOM-TOP_GRID_DIMENSION", nz + 1);
        w.att_text(-1, "SIMULATION_START_DATE", start_time);
        w.att_text(-1, "START_DATE", start_time);
        w.att_double(-1, "DX", cfg.dx);
        w.att_double(-1, "DY", cfg.dx);  // sic, write_data.F90:230
        w.att_double(-1, "DT", config_dt);
        w.att_int(-1, "SF_SURFACE_PHYSICS", lsm_scheme);
        w.att_int(-1, "MP_PHYSICS", mp_scheme);
        w.att_int(-1, "CU_PHYSICS", conv_scheme);
        w.att_double(-1, "CEN_LAT", cen_lat);
        w.att_double(-1, "CEN_LON", cen_lon);
        w.att_double(-1, "TRUELAT1", cfg.truelat1);
        w.att_double(-1, "TRUELAT2", cfg.truelat2);
        w.att_double(-1, "MOAD_CEN_LAT", cen_lat);
        w.att_double(-1, "STAND_LON", cfg.stand_lon);
        w.att_double(-1, "POLE_LAT", cfg.pole_lat);
        w.att_double(-1, "POLE_LON", cfg.pole_lon);
        w.att_double(-1, "POL_ELAT", cfg.pole_lat);
        w.att_int(-1, "MAP_PROJ", cfg.proj_code);
        w.att_text(-1, "MAP_PROJ_CHAR", rtrim(cfg.map_proj_char));
        if (cfg.interp_diag) w.att_int(-1, "PREC_ACC_DT", diag_out_interval);
        w.att_int(-1, "I_PARENT_START", 1);
        w.att_int(-1, "J_PARENT_START", 1);
        w.att_int(-1, "WEST-EAST_PATCH_START_UNSTAG", 1);
        w.att_int(-1, "WEST-EAST_PATCH_START_STAG", 1);
        w.att_int(-1, "SOUTH-NORTH_PATCH_START_UNSTAG", 1);
        w.att_int(-1, "SOUTH-NORTH_PATCH_START_STAG", 1);
        w.att_int(-1, "BOTTOM-TOP_PATCH_START_UNSTAG", 1);
        w.att_int(-1, "BOTTOM-TOP_PATCH_START_STAG", 1);
        w.att_int(-1, "WEST-EAST_PATCH_END_UNSTAG", it);
        w.att_int(-1, "WEST-EAST_PATCH_END_STAG", it + 1);
        w.att_int(-1, "SOUTH-NORTH_PATCH_END_UNSTAG", jt);
        w.att_int(-1, "SOUTH-NORTH_PATCH_END_STAG", jt + 1);
        w.att_int(-1, "BOTTOM-TOP_PATCH_END_UNSTAG", nz);
        w.att_int(-1, "BOTTOM-TOP_PATCH_END_STAG", nz + 1);

        // NetCDF dimension order = the Fortran list reversed: (/dim_lon, dim_lat, dim_time/) -> [Time][south_north][west_east]
        const std::vector<int> d2M{dTime, dSN, dWE}, d2U{dTime, dSN, dWEs}, d2V{dTime, dSNs, dWE};
        auto def = [&](const char *name, const std::vector<int> &dims, const char *desc, const char *units, const char *mo,
                       const char *coords, const char *stag, int type = ncio::NC_FLOAT) {
            const int id = w.def_var(name, type, dims);
            if (desc) w.att_text(id, "description", desc);
            if (units) w.att_text(id, "units", units);
            if (mo) w.att_text(id, "MemoryOrder", mo);
            if (coords) w.att_text(id, "coordinates", coords);
            if (stag) w.att_text(id, "stagger", stag);
            w.att_int(id, "FieldType", 104);
            return id;
        };
        const int id_lon = def("XLONG", d2M, "LONGITUDE, WEST IS NEGATIVE", "degree_east", "XY ", "XLONG XLAT", "");
        const int id_lonu = def("XLONG_U", d2U, "LONGITUDE, WEST IS NEGATIVE", "degree_east", "XY ", "XLONG_U XLAT_U", "X");
        const int id_lonv = def("XLONG_V", d2V, "LONGITUDE, WEST IS NEGATIVE", "degree_east", "XY ", "XLONG_V XLAT_V", "Y");
        const int id_lat = def("XLAT", d2M, "LATITUDE, SOUTH IS NEGATIVE", "degree_north", "XY ", "XLONG XLAT", "");
        const int id_latu = def("XLAT_U", d2U, "LATITUDE, SOUTH IS NEGATIVE", "degree_north", "XY ", "XLONG_U XLAT_U", "X");
        const int id_latv = def("XLAT_V", d2V, "LATITUDE, SOUTH IS NEGATIVE", "degree_north", "XY ", "XLONG_V XLAT_V", "Y");
        // (the map factors carry the latitude descriptions in the reference, write_data.F90:400-420)
        const int id_mfm = def("MAPFAC_M", d2M, "LATITUDE, SOUTH IS NEGATIVE", "degree_north", "XY ", "XLONG XLAT", " ");
        const int id_mfu = def("MAPFAC_U", d2U, "LATITUDE, SOUTH IS NEGATIVE", "degree_north", "XY ", "XLONG_U XLAT_U", "X");
        const int id_mfv = def("MAPFAC_V", d2V, "LATITUDE, SOUTH IS NEGATIVE", "degree_north", "XY ", "XLONG_V XLAT_V", "Y");
        int id_sina = -1, id_cosa = -1;
        if (lc) {
            // the second block of put_att calls targets id_sina again (write_data.F90:436-442): SINALPHA ends up with
            // the COSINE description, COSALPHA with no attributes at all
            id_sina = def("SINALPHA", d2M, "COSINE OF GRID ROTATION ANGLE ALPHA", " ", "XY ", "XLONG XLAT", " ");
            id_cosa = w.def_var("COSALPHA", ncio::NC_FLOAT, d2M);
        }
        const int id_z = def("Z_C", {dTime, dZs, dSN, dWE}, "Layer center height above mean sea level", "m AMSL", "XYZ ",
                             "XLAT XLONG Z_C", "");
        const int id_zs = def("ZS", {dTime, dSoil}, "DEPTHS OF CENTERS OF SOIL LAYERS", "m", "X", "ZS XTIME", "");
        const int id_hgt = def("HGT", d2M, "TERRAIN HEIGHT ", "m AMSL", "XY ", "XLAT XLONG ", "");
        const int id_times = def("Times", {dTime, dStr}, "Times", "m", nullptr, "Time", "", ncio::NC_CHAR);
        const int id_itime = w.def_var("ITIMESTEP", ncio::NC_INT, {dTime});
        w.att_text(id_itime, "description", "");
        w.att_text(id_itime, "units", "");
        w.att_text(id_itime, "stagger", "");
        w.att_int(id_itime, "FieldType", 106);
        w.att_text(id_itime, "MemoryOrder", "O ");
        const int id_xtime = w.def_var("XTIME", ncio::NC_FLOAT, {dTime});
        w.att_text(id_xtime, "description", "minutes since " + start_time);
        w.att_text(id_xtime, "units", "minutes since " + start_time);
        w.att_text(id_xtime, "stagger", "");
        w.att_int(id_xtime, "FieldType", 104);
        w.att_text(id_xtime, "MemoryOrder", "O ");

        std::vector<OutVar> outs;
        auto def_field = [&](const char *name, int stagger, int32_t nlev, int zdim, const char *mo, const char *coords,
                             const std::string &units, const std::string &desc, const char *stag, void *dev) {
            // order of the attributes as in write_data.F90:600-606
            const std::vector<int> &h = stagger == MPRG_EDGE1 ? d2U : stagger == MPRG_EDGE2 ? d2V : d2M;
            std::vector<int> dims = zdim >= 0 ? std::vector<int>{dTime, zdim, h[1], h[2]} : h;
            const int id = w.def_var(name, ncio::NC_FLOAT, dims);
            w.att_text(id, "MemoryOrder", mo);
            w.att_text(id, "coordinates", coords);
            w.att_text(id, "units", units);
            w.att_text(id, "description", desc);
            w.att_text(id, "stagger", stag);
            w.att_int(id, "FieldType", 104);
            OutVar o;
            o.name = name; o.varid = id; o.stagger = stagger; o.nlev = nlev; o.dev = dev;
            outs.push_back(o);
            return (int)outs.size() - 1;
        };
        const char *cM = "XLONG XLAT XTIME";
        // MU, PH and P are defined but never written: the file is created zero-filled, which is their content
        int o_T = -1, o_phyd = -1, o_phb = -1, id_ptop = -1, id_pb = -1;
        // diag fields, defined in list order whatever their rank (write_data.F90:571-613)
        for (size_t i = 0; i < fd.size(); ++i) {
            if (fd[i].nlev == 1) def_field(fd[i].target_name, MPRG_CENTER, 1, -1, "XY ", cM, sd[i].units, sd[i].longname, "", fd[i].dst);
            else def_field(fd[i].target_name, MPRG_CENTER, fd[i].nlev, dZ, "XYZ ", cM, sd[i].units, sd[i].longname, "", fd[i].dst);
        }
        if (cfg.interp_hist) {
            for (int pass = 0; pass < 3; ++pass) {  // 2d_cons, 2d_patch, 2d_nstd (write_data.F90:618-698)
                const int want = pass == 0 ? MPASSIT_CLASS_2D_CONS : pass == 1 ? MPASSIT_CLASS_2D_PATCH : MPASSIT_CLASS_2D_NSTD;
                for (size_t i = 0; i < f2.size(); ++i)
                    if (f2[i].klass == want)
                        def_field(f2[i].target_name, MPRG_CENTER, 1, -1, "XY ", cM, s2[i].units, s2[i].longname, "", f2[i].dst);
            }
            for (size_t i = 0; i < fs.size(); ++i)  // :700-725
                def_field(fs[i].target_name, MPRG_CENTER, fs[i].nlev, dSoil, "XYZ ", cM, ss[i].units, ss[i].longname, "", fs[i].dst);
            for (size_t i = 0; i < f3.size(); ++i) {  // 3d_nz, :729-776
                if (f3[i].klass != MPASSIT_CLASS_3D_NZ) continue;
                const int o = def_field(f3[i].target_name, MPRG_CENTER, nz, dZ, "XYZ ", cM, s3[i].units, s3[i].longname, "", f3[i].dst);
                const std::string tn = f3[i].target_name;
                if (cfg.wrf_mod_vars && tn == "T") o_T = o;
                if (cfg.wrf_mod_vars && tn == "MUB") {
                    const int id_mu = w.def_var("MU", ncio::NC_FLOAT, {dTime, dZ, dSN, dWE});
                    w.att_text(id_mu, "MemoryOrder", "XYZ ");
                    w.att_text(id_mu, "coordinates", cM);
                    w.att_text(id_mu, "units", s3[i].units);
                    w.att_text(id_mu, "description", "Perturbation " + s3[i].longname);
                    w.att_text(id_mu, "stagger", "");
                    w.att_int(id_mu, "FieldType", 104);
                }
                if (cfg.wrf_mod_vars && tn == "P_HYD") {
                    o_phyd = o;
                    id_ptop = w.def_var("P_TOP", ncio::NC_FLOAT, {dTime});
                    w.att_text(id_ptop, "MemoryOrder", "0 ");
                    w.att_text(id_ptop, "units", s3[i].units);
                    w.att_text(id_ptop, "description", "PRESSURE TOP OF THE MODEL");
                    w.att_text(id_ptop, "stagger", "");
                    w.att_int(id_ptop, "FieldType", 104);
                }
            }
            if (do_u) def_field("U", MPRG_EDGE1, nz, dZ, "XYZ ", "XLONG_U XLAT_U XTIME", "m s^{-1}", "", "X", io.u_stag);  // :779-789
            if (do_v) def_field("V", MPRG_EDGE2, nz, dZ, "XYZ ", "XLONG_V XLAT_V XTIME", "m s^{-1}", "", "Y", io.v_stag);  // :790-800
            for (size_t i = 0; i < f3.size(); ++i) {  // 3d_nzp1, :803-845
                if (f3[i].klass != MPASSIT_CLASS_3D_NZP1) continue;
                const bool phb = cfg.wrf_mod_vars && std::string(f3[i].target_name) == "PHB";
                const int o = def_field(f3[i].target_name, MPRG_CENTER, nzp1, dZs, "XYZ ", cM, phb ? "gpm" : s3[i].units,
                                        phb ? "Base Geopotential Height" : s3[i].longname, "Z", f3[i].dst);
                // Z_C and the x 9.81 are keyed on the name alone (write_data.F90:1403), PH on wrf_mod_vars too
                if (std::string(f3[i].target_name) == "PHB") o_phb = o;
                if (phb) {
                    const int id_ph = w.def_var("PH", ncio::NC_FLOAT, {dTime, dZs, dSN, dWE});
                    w.att_text(id_ph, "MemoryOrder", "XYZ ");
                    w.att_text(id_ph, "coordinates", cM);
                    w.att_text(id_ph, "units", "gpm");
                    w.att_text(id_ph, "description", "Perturbation Geopotential Height");
                    w.att_text(id_ph, "stagger", "Z");
                    w.att_int(id_ph, "FieldType", 104);
                }
            }
            for (size_t i = 0; i < f3.size(); ++i)  // 3d_vert, :847-869 (MemoryOrder "XYZ" without the blank)
                if (f3[i].klass == MPASSIT_CLASS_3D_VERT)
                    def_field(f3[i].target_name, MPRG_CENTER, nz, dZ, "XYZ", cM, s3[i].units, s3[i].longname, "", f3[i].dst);
        }
        if (cfg.wrf_mod_vars) {  // :872-893
            auto def_dummy = [&](const char *name, const char *desc) {
                const int id = w.def_var(name, ncio::NC_FLOAT, {dTime, dZ, dSN, dWE});
                w.att_text(id, "MemoryOrder", "XYZ ");
                w.att_text(id, "coordinates", cM);
                w.att_text(id, "units", "Pa");
                w.att_text(id, "description", desc);
                w.att_text(id, "stagger", "");
                w.att_int(id, "FieldType", 104);
                return id;
            };
            def_dummy("P", "perturbation pressure (0.0)");
            id_pb = def_dummy("PB", "BASE STATE PRESSURE (pfull)");
        }

        // every rank lays out the same header; rank 0 creates the file, the others open it
        if (rank == 0 && !w.enddef(cfg.output_file, 1, true, why)) netcdf_err(std::string("opening") + cfg.output_file, why);
        barrier();
        if (rank != 0 && !w.enddef(cfg.output_file, 1, false, why)) netcdf_err(std::string("opening") + cfg.output_file, why);

        auto wr = [&](bool ok) {
            if (!ok) netcdf_err("writing to " + std::string(cfg.output_file), why);
        };
        if (rank == 0) {  // grid description + times: small arrays, written where the reference writes them (PET 0)
            wr(w.put_doubles(id_lon, lon[0].data(), lon[0].size(), why));
            wr(w.put_doubles(id_lat, lat[0].data(), lat[0].size(), why));
            wr(w.put_doubles(id_lonu, lon[1].data(), lon[1].size(), why));
            wr(w.put_doubles(id_latu, lat[1].data(), lat[1].size(), why));
            wr(w.put_doubles(id_latv, lat[2].data(), lat[2].size(), why));
            wr(w.put_doubles(id_lonv, lon[2].data(), lon[2].size(), why));
            wr(w.put_doubles(id_mfm, mapfac[0].data(), mapfac[0].size(), why));
            wr(w.put_doubles(id_mfu, mapfac[1].data(), mapfac[1].size(), why));
            wr(w.put_doubles(id_mfv, mapfac[2].data(), mapfac[2].size(), why));
            if (lc) {
                wr(w.put_doubles(id_sina, sina.data(), sina.size(), why));
                wr(w.put_doubles(id_cosa, cosa.data(), cosa.size(), why));
            }
            wr(w.put_doubles(id_zs, zs.data(), zs.size(), why));
            wr(w.put_text(id_times, valid_time.substr(0, std::min<size_t>(19, valid_time.size())), why));
            int64_t ssec = 0, vsec = 0;
            if (!stamp_seconds(start_time, &ssec)) die(4, "IN write_to_file: cannot parse config_start_time '" + start_time + "'");
            if (!stamp_seconds(valid_time, &vsec)) die(4, "IN write_to_file: cannot parse xtime '" + valid_time + "'");
            const double total_seconds = (double)(ssec - vsec);  // datetime(start) - datetime(valid), write_data.F90:1198
            const double xt = total_seconds / 60.0;
            wr(w.put_doubles(id_xtime, &xt, 1, why));
            const int32_t itimestep = config_dt > 0.0 ? (int32_t)(total_seconds / config_dt) : 0;  // :1205-1209
            wr(w.put_ints(id_itime, &itimestep, 1, why));
        }

        // this rank's rows of every regridded field: post-op + byte swap in HBM, one download, pwrite per level run
        // Two pinned buffers of a few levels each (page-locking memory costs ~0.3 s per GB: whole-field buffers would
        // cost more than they save): while a background thread writes one chunk, the next is downloaded into the other.
        size_t pin_bytes = (size_t)64 << 20;
        for (int s = 0; s < 3; ++s) pin_bytes = std::max(pin_bytes, (size_t)(j1[s] - j0[s]) * ni[s] * 4);
        if (!dry)
            for (void *&pb : pinned) ck(ctx, mprg_host_alloc(ctx, pin_bytes, &pb), "host_alloc");
        struct Job {
            int varid;
            uint64_t off;
            const uint8_t *p;
            size_t n;
        };
        int cur = 0;
        auto flush = [&]() {
            const double tw = now_ms();
            if (writer.joinable()) writer.join();
            st.writer_wait_ms += now_ms() - tw;
            if (!writer_ok) netcdf_err("writing to " + std::string(cfg.output_file), writer_err);
        };
        auto submit = [&](std::vector<Job> jobs) {
            flush();  // at most one field in flight: it owns the other buffer
            writer = std::thread([&w, &writer_ok, &writer_err, jobs]() {
                for (const Job &j : jobs)
                    if (!w.write_raw(j.varid, 0, j.off, j.p, j.n, writer_err)) {
                        writer_ok = false;
                        return;
                    }
            });
        };
        // also_varid: a second variable that receives the same bytes (PB <- P_HYD)
        auto write_slab = [&](int varid, int s, int32_t nlev, const void *dev, int also_varid = -1) {
            const size_t rows = (size_t)(j1[s] - j0[s]), slab = rows * ni[s];
            if (slab == 0 || nlev == 0) return;
            const size_t plane = (size_t)nj[s] * ni[s];
            const int32_t per = (int32_t)std::max<size_t>(1, pin_bytes / (slab * 4));  // levels per chunk
            for (int32_t l0 = 0; l0 < nlev; l0 += per) {
                const int32_t n = std::min(per, nlev - l0);
                const uint8_t *buf = (const uint8_t *)pinned[cur];
                cur ^= 1;  // the writer in flight (at most one) reads the other buffer
                const double td = now_ms();
                ck(ctx, mprg_download(ctx, (const uint8_t *)dev + (size_t)l0 * slab * 4, (void *)buf, slab * n * 4), "download");
                st.download_ms += now_ms() - td;
                std::vector<Job> jobs;
                for (int id : {varid, also_varid}) {
                    if (id < 0) continue;
                    if (rows == (size_t)nj[s]) {  // whole rows of the grid: the levels are one run in the file
                        jobs.push_back(Job{id, (size_t)l0 * plane * 4, buf, slab * n * 4});
                    } else {
                        for (int32_t l = 0; l < n; ++l)
                            jobs.push_back(Job{id, ((size_t)(l0 + l) * plane + (size_t)j0[s] * ni[s]) * 4, buf + l * slab * 4, slab * 4});
                    }
                    st.bytes_out += (int64_t)(slab * n * 4);
                }
                submit(std::move(jobs));
            }
        };
        auto swap_dev = [&](void *dev, int s, int32_t nlev) {
            if (!dry) ck(ctx, mprg_bswap(ctx, dev, (size_t)(j1[s] - j0[s]) * ni[s] * nlev, MPRG_F32), "bswap");
        };
        const size_t slabM = (size_t)(j1[0] - j0[0]) * ni[0];
        if (cfg.interp_hist) {  // HGT, :1128-1139
            swap_dev(io.hgt, 0, 1);
            write_slab(id_hgt, 0, 1, io.hgt);
        } else if (from_file && rank == 0) {  // hgt_target_grid as read from the target file (model_grid.F90:1868-1884)
            wr(w.put_doubles(id_hgt, hgt_file.data(), hgt_file.size(), why));
        }
        for (size_t k = 0; k < outs.size(); ++k) {
            OutVar &o = outs[k];
            const int s = o.stagger == MPRG_EDGE1 ? 1 : o.stagger == MPRG_EDGE2 ? 2 : 0;
            if ((int)k == o_phyd && !dry) {
                // P_TOP (write_data.F90:1364-1373) and PB, which receives the same values (:1376-1378)
                double part[2] = {0.0, std::numeric_limits<double>::infinity()};
                if (slabM) ck(ctx, mprg_post_ptop(ctx, MPRG_CENTER, nz, MPRG_F32, MPRG_DEVICE, o.dev, &part[0], &part[1]), "post_ptop");
                else part[0] = -std::numeric_limits<double>::infinity();
                if (nranks > 1) {
                    comm(comm_arg, MPASSIT_COMM_MAX, &part[0], 1);
                    comm(comm_arg, MPASSIT_COMM_MIN, &part[1], 1);
                }
                const double ptop = std::min(part[0], part[1]);
                if (rank == 0) wr(w.put_doubles(id_ptop, &ptop, 1, why));
                st.p_top = ptop;
            }
            if ((int)k == o_T && !dry) ck(ctx, mprg_post_affine(ctx, o.dev, slabM * nz, MPRG_F32, 1.0, -300.0), "post_affine");
            if ((int)k == o_phb && !dry) {
                // Z_C = mid-level average of the regridded zgrid (:1406-1413), THEN zgrid x 9.81 (:1414)
                void *mid = d_mid;
                ck(ctx, mprg_post_midlevels(ctx, MPRG_CENTER, nzp1, MPRG_F32, MPRG_DEVICE, o.dev, mid), "post_midlevels");
                swap_dev(mid, 0, nzp1 - 1);
                write_slab(id_z, 0, nzp1 - 1, mid);
                if (slabM) {  // the level the reference never writes reads back as the fill value
                    std::vector<float> fill(slabM, kFillFloat);
                    ncio::to_big_endian(fill.data(), 4, fill.size());
                    wr(w.write_raw(id_z, 0, ((size_t)(nzp1 - 1) * nj[0] * ni[0] + (size_t)j0[0] * ni[0]) * 4, fill.data(),
                                   fill.size() * 4, why));
                }
                ck(ctx, mprg_post_affine(ctx, o.dev, slabM * nzp1, MPRG_F32, 9.81, 0.0), "post_affine");
            }
            swap_dev(o.dev, s, o.nlev);
            write_slab(o.varid, s, o.nlev, o.dev, (int)k == o_phyd ? id_pb : -1);  // PB: the same bytes again (:1376)
        }
        flush();
        st.n_vars_written = (int32_t)outs.size();
        wr(w.close(why));
        barrier();
        st.write_ms = now_ms() - t3;
        st.total_ms = now_ms() - t0;
        st.output_version = w.version();
    } catch (const Fail &f) {
        if (err && errlen) std::snprintf(err, errlen, "%s", f.msg.c_str());
        rc_out = f.rc;
        // error_handler (utils.F90:16-33) ends with mpi_abort: the other ranks must not be left waiting in a barrier
        // or in the P_TOP reductions.  The host's communicator decides how (MPI_Abort in a Fortran host).
        if (nranks > 1 && comm) {
            double code = (double)(f.rc ? f.rc : 1);
            comm(comm_arg, MPASSIT_COMM_ABORT, &code, 1);
        }
    }
    if (writer.joinable()) writer.join();
    if (ctx) {
        for (void *pb : pinned)
            if (pb) mprg_host_free(ctx, pb);
        for (void *p : dev_allocs) mprg_device_free(ctx, p);
        mprg_finalize(ctx);  // cleanup_input_target_grid_data + ESMF_finalize, mpassit.F90:137-142
    }
    if (stats) *stats = st;
    return rc_out;
}

}  // extern "C"
