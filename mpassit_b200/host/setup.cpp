// setup.cpp -- host mirror of program_setup.F90 (namelist), input_data.F90's
// var-list reader / regrid-class tables, and model_grid.F90's decomposition helpers.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/mpassit_host.h"

namespace {

void set_err(char *err, size_t n, const std::string &m) {
    if (err && n) {
        std::snprintf(err, n, "%s", m.c_str());
    }
}

std::string upper(std::string s) {  // to_upper, utils.F90:67
    for (auto &c : s) c = (char)std::toupper((unsigned char)c);
    return s;
}
std::string lower(std::string s) {
    for (auto &c : s) c = (char)std::tolower((unsigned char)c);
    return s;
}
std::string trim(const std::string &s) {
    size_t b = s.find_first_not_of(" \t\r\n"), e = s.find_last_not_of(" \t\r\n");
    return b == std::string::npos ? "" : s.substr(b, e - b + 1);
}

bool parse_logical(const std::string &v, int32_t *out) {
    std::string u = lower(trim(v));
    if (u == ".true." || u == "t" || u == ".t." || u == "true") { *out = 1; return true; }
    if (u == ".false." || u == "f" || u == ".f." || u == "false") { *out = 0; return true; }
    return false;
}
bool parse_real(const std::string &v, double *out) {
    std::string u = trim(v);
    for (auto &c : u) if (c == 'd' || c == 'D') c = 'e';
    char *end = nullptr;
    double x = std::strtod(u.c_str(), &end);
    if (end == u.c_str() || *end != '\0') return false;
    *out = x;
    return true;
}
bool parse_int(const std::string &v, int32_t *out) {
    std::string u = trim(v);
    char *end = nullptr;
    long x = std::strtol(u.c_str(), &end, 10);
    if (end == u.c_str() || *end != '\0') return false;
    *out = (int32_t)x;
    return true;
}
bool parse_string(const std::string &v, char *out) {
    std::string u = trim(v);
    if (u.size() >= 2 && (u.front() == '\'' || u.front() == '"') && u.back() == u.front())
        u = u.substr(1, u.size() - 2);
    else if (!u.empty() && (u.front() == '\'' || u.front() == '"'))
        return false;
    std::snprintf(out, MPASSIT_STRLEN, "%s", u.c_str());
    return true;
}

// Fortran namelist group -> (key, value) pairs; values keep their quotes.
bool namelist_pairs(const std::string &text, const std::string &group,
                    std::vector<std::pair<std::string, std::string>> &out) {
    // strip comments outside strings
    std::string s;
    char q = 0;
    for (size_t i = 0; i < text.size(); ++i) {
        char c = text[i];
        if (q) { s += c; if (c == q) q = 0; continue; }
        if (c == '\'' || c == '"') { q = c; s += c; continue; }
        if (c == '!') { while (i < text.size() && text[i] != '\n') ++i; s += '\n'; continue; }
        s += c;
    }
    std::string low = lower(s);
    size_t g = low.find("&" + lower(group));
    if (g == std::string::npos) return false;
    size_t i = g + 1 + group.size();
    // scan to terminating '/' or '&end' outside strings
    size_t end = std::string::npos;
    q = 0;
    for (size_t k = i; k < s.size(); ++k) {
        char c = s[k];
        if (q) { if (c == q) q = 0; continue; }
        if (c == '\'' || c == '"') { q = c; continue; }
        if (c == '/') { end = k; break; }
        if (c == '&' && low.compare(k, 4, "&end") == 0) { end = k; break; }
    }
    if (end == std::string::npos) return false;
    std::string body = s.substr(i, end - i);
    // split on '=' : key is the identifier before it, value runs to the next key
    std::vector<size_t> eqs;
    q = 0;
    for (size_t k = 0; k < body.size(); ++k) {
        char c = body[k];
        if (q) { if (c == q) q = 0; continue; }
        if (c == '\'' || c == '"') { q = c; continue; }
        if (c == '=') eqs.push_back(k);
    }
    if (eqs.empty()) return trim(body).empty();
    std::vector<std::pair<size_t, size_t>> keys;  // [begin,end) of identifier before each '='
    for (size_t e : eqs) {
        size_t b = e;
        while (b > 0 && std::isspace((unsigned char)body[b - 1])) --b;
        size_t kend = b;
        while (b > 0 && (std::isalnum((unsigned char)body[b - 1]) || body[b - 1] == '_' || body[b - 1] == '%')) --b;
        if (b == kend) return false;
        keys.push_back({b, kend});
    }
    if (!trim(body.substr(0, keys[0].first)).empty()) return false;
    for (size_t n = 0; n < eqs.size(); ++n) {
        size_t vb = eqs[n] + 1, ve = n + 1 < eqs.size() ? keys[n + 1].first : body.size();
        std::string v = trim(body.substr(vb, ve - vb));
        while (!v.empty() && (v.back() == ',' )) v = trim(v.substr(0, v.size() - 1));
        out.push_back({lower(body.substr(keys[n].first, keys[n].second - keys[n].first)), v});
    }
    return true;
}

}  // namespace

extern "C" {

int mpassit_read_setup_namelist(const char *filename, mpassit_config *cfg, char *err, size_t errlen) {
    if (!cfg) return 1;
    std::memset(cfg, 0, sizeof *cfg);
    // defaults: program_setup.F90:23-75 and :108-117
    for (char *s : {cfg->grid_file_input_grid, cfg->diag_file_input_grid, cfg->hist_file_input_grid,
                    cfg->file_target_grid, cfg->output_file, cfg->block_decomp_file})
        std::snprintf(s, MPASSIT_STRLEN, "NULL");
    cfg->is_regional = 1;
    cfg->interp_as_bundle = 1;
    cfg->truelat1 = cfg->truelat2 = cfg->stand_lon = MPASSIT_NAN;
    cfg->ref_x = cfg->ref_y = cfg->ref_lat = cfg->ref_lon = cfg->dx = cfg->dy = MPASSIT_NAN;
    cfg->pole_lat = 90.0;
    cfg->pole_lon = 0.0;

    std::ifstream in(filename ? filename : "./fort.41");  // program_setup.F90:124-125
    if (!in) { set_err(err, errlen, "OPENING SETUP NAMELIST."); return 2; }
    std::stringstream ss;
    ss << in.rdbuf();
    std::vector<std::pair<std::string, std::string>> kv;
    if (!namelist_pairs(ss.str(), "config", kv)) { set_err(err, errlen, "READING SETUP NAMELIST."); return 3; }
    for (auto &p : kv) {
        const std::string &k = p.first, &v = p.second;
        bool ok = true;
        if (k == "grid_file_input_grid") ok = parse_string(v, cfg->grid_file_input_grid);
        else if (k == "diag_file_input_grid") ok = parse_string(v, cfg->diag_file_input_grid);
        else if (k == "hist_file_input_grid") ok = parse_string(v, cfg->hist_file_input_grid);
        else if (k == "file_target_grid") ok = parse_string(v, cfg->file_target_grid);
        else if (k == "output_file") ok = parse_string(v, cfg->output_file);
        else if (k == "block_decomp_file") ok = parse_string(v, cfg->block_decomp_file);
        else if (k == "target_grid_type") ok = parse_string(v, cfg->target_grid_type);
        else if (k == "interp_diag") ok = parse_logical(v, &cfg->interp_diag);
        else if (k == "interp_hist") ok = parse_logical(v, &cfg->interp_hist);
        else if (k == "wrf_mod_vars") ok = parse_logical(v, &cfg->wrf_mod_vars);
        else if (k == "esmf_log") ok = parse_logical(v, &cfg->esmf_log);
        else if (k == "is_regional") ok = parse_logical(v, &cfg->is_regional);
        else if (k == "interp_as_bundle") ok = parse_logical(v, &cfg->interp_as_bundle);
        else if (k == "nx") ok = parse_int(v, &cfg->nx);
        else if (k == "ny") ok = parse_int(v, &cfg->ny);
        else if (k == "dx") ok = parse_real(v, &cfg->dx);
        else if (k == "dy") ok = parse_real(v, &cfg->dy);
        else if (k == "ref_lat") ok = parse_real(v, &cfg->ref_lat);
        else if (k == "ref_lon") ok = parse_real(v, &cfg->ref_lon);
        else if (k == "ref_x") ok = parse_real(v, &cfg->ref_x);
        else if (k == "ref_y") ok = parse_real(v, &cfg->ref_y);
        else if (k == "truelat1") ok = parse_real(v, &cfg->truelat1);
        else if (k == "truelat2") ok = parse_real(v, &cfg->truelat2);
        else if (k == "stand_lon") ok = parse_real(v, &cfg->stand_lon);
        else if (k == "pole_lat") ok = parse_real(v, &cfg->pole_lat);
        else if (k == "pole_lon") ok = parse_real(v, &cfg->pole_lon);
        else ok = false;  // unknown namelist member: Fortran's read fails too
        if (!ok) { set_err(err, errlen, "READING SETUP NAMELIST."); return 3; }
    }
    if (std::strcmp(cfg->block_decomp_file, "NULL") != 0) {  // program_setup.F90:146-152
        std::ifstream f(cfg->block_decomp_file);
        if (!f) { set_err(err, errlen, "block_decomp_file DOES NOT EXIST."); return 4; }
    }
    if (trim(cfg->target_grid_type) == "file") return 0;  // derived values come from the WRF file

    // program_setup.F90:155-243
    const double PI = 3.141592653589793, EARTH_RADIUS_M = 6370000.0;
    cfg->dxkm = cfg->dx;
    cfg->dykm = cfg->dy;
    cfg->known_lat = cfg->ref_lat;
    cfg->known_lon = cfg->ref_lon;
    cfg->known_x = cfg->ref_x;
    cfg->known_y = cfg->ref_y;
    cfg->i_target = cfg->nx - 1;
    cfg->j_target = cfg->ny - 1;
    std::string mp = upper(trim(cfg->target_grid_type));
    if (mp == "LAMBERT") { cfg->proj_code = MPASSIT_PROJ_LC; std::snprintf(cfg->map_proj_char, MPASSIT_NAMELEN, "Lambert Conformal"); }
    else if (mp == "MERCATOR") { cfg->proj_code = MPASSIT_PROJ_MERC; std::snprintf(cfg->map_proj_char, MPASSIT_NAMELEN, "Mercator"); }
    else if (mp == "POLAR") { cfg->proj_code = MPASSIT_PROJ_PS; std::snprintf(cfg->map_proj_char, MPASSIT_NAMELEN, "Polar Stereographic"); }
    else if (mp == "LAT-LON") { cfg->proj_code = MPASSIT_PROJ_LATLON; std::snprintf(cfg->map_proj_char, MPASSIT_NAMELEN, "Lat/Lon"); }
    else {
        set_err(err, errlen, "In namelist, invalid target_grid_type specified. Valid projections are "
                             "\"lambert\", \"mercator\", \"polar\", and \"lat-lon\".");
        return 5;
    }
    if (cfg->proj_code == MPASSIT_PROJ_LATLON) {
        if (cfg->dx == MPASSIT_NAN && cfg->dy == MPASSIT_NAN) {
            if (cfg->is_regional) {
                set_err(err, errlen, "For lat-lon projection, if dx/dy are not specified a global grid is assumed. "
                                     "Please set dx/dy if a regional grid is desired, or change is_regional to "
                                     ".false. if a global grid is desired.");
                return 6;
            }
            cfg->dlondeg = 360.0 / cfg->i_target;
            cfg->dlatdeg = 180.0 / cfg->j_target;
            cfg->known_x = 1.0;
            cfg->known_y = 1.0;
            cfg->known_lon = cfg->stand_lon + cfg->dlondeg / 2.0;
            cfg->known_lat = -90.0 + cfg->dlatdeg / 2.0;
            cfg->dxkm = EARTH_RADIUS_M * PI * 2.0 / cfg->i_target;
            cfg->dykm = EARTH_RADIUS_M * PI / cfg->j_target;
        } else {
            if (!cfg->is_regional) {
                set_err(err, errlen, "For lat-lon projection, if dx/dy are specified a regional grid is assumed. "
                                     "Please unset dx/dy if a global grid is desired, or change is_regional to "
                                     ".true. if a regional grid is desired.");
                return 7;
            }
            cfg->dlatdeg = cfg->dy;
            cfg->dlondeg = cfg->dx;
            cfg->dxkm = cfg->dlondeg * EARTH_RADIUS_M * PI * 2.0 / 360.0;
            cfg->dykm = cfg->dlatdeg * EARTH_RADIUS_M * PI * 2.0 / 360.0;
            if (cfg->known_lat == MPASSIT_NAN || cfg->known_lon == MPASSIT_NAN) {
                set_err(err, errlen, "For lat-lon projection, if dx/dy are specified, a regional domain is assumed, "
                                     "and a ref_lat,ref_lon must also be specified");
                return 8;
            }
        }
    }
    if (cfg->proj_code == MPASSIT_PROJ_LC && cfg->truelat2 == MPASSIT_NAN) {
        if (cfg->truelat1 == MPASSIT_NAN) {
            set_err(err, errlen, "No TRUELAT1 specified for Lambert conformal projection.");
            return 9;
        }
        cfg->truelat2 = cfg->truelat1;
    }
    if (cfg->known_x == MPASSIT_NAN && cfg->known_y == MPASSIT_NAN) {
        cfg->known_x = (double)(cfg->i_target + 1) / 2.0;
        cfg->known_y = (double)(cfg->j_target + 1) / 2.0;
    } else if (cfg->known_x == MPASSIT_NAN || cfg->known_y == MPASSIT_NAN) {
        set_err(err, errlen, "In namelist, neither or both of ref_x, ref_y must be specified.");
        return 10;
    }
    return 0;
}

int mpassit_read_varlist(const char *file, int32_t max, int32_t *nfields, char *field_names,
                         char *field_names_target, char *err, size_t errlen) {
    if (!file || !nfields) return 1;
    std::ifstream in(file);
    if (!in) { set_err(err, errlen, std::string("VARLIST FILE ") + file + " not exist"); return 1; }
    std::string line;
    int32_t n = 0;
    while (std::getline(in, line)) {
        if (trim(line).empty()) continue;  // blank lines skipped, input_data.F90:1177
        // list-directed read of two character items: blanks, tabs or commas separate; quotes optional
        std::vector<std::string> items;
        std::string cur;
        char q = 0;
        for (char c : line + " ") {
            if (q) { if (c == q) { q = 0; items.push_back(cur); cur.clear(); } else cur += c; continue; }
            if (c == '\'' || c == '"') { q = c; continue; }
            if (std::isspace((unsigned char)c) || c == ',') { if (!cur.empty()) { items.push_back(cur); cur.clear(); } continue; }
            cur += c;
        }
        if (items.size() < 2) { set_err(err, errlen, "READING VARLIST FILE"); return 2; }
        if (n < max && field_names && field_names_target) {
            std::snprintf(field_names + (size_t)n * MPASSIT_NAMELEN, MPASSIT_NAMELEN, "%s", items[0].c_str());
            std::snprintf(field_names_target + (size_t)n * MPASSIT_NAMELEN, MPASSIT_NAMELEN, "%s", items[1].c_str());
        }
        ++n;
    }
    *nfields = n;
    if (n > max && field_names) { set_err(err, errlen, "READING VARLIST FILE: more fields than buffer"); return 3; }
    return 0;
}

int mpassit_classify_hist_2d(const char *name) {
    // cons_vars / nstd_vars, input_data.F90:840-841
    static const char *cons[] = {"snow", "snowh"};
    static const char *nstd[] = {"ivgtyp", "isltyp", "xland", "landmask"};
    for (auto c : cons) if (!std::strcmp(name, c)) return MPASSIT_CLASS_2D_CONS;
    for (auto c : nstd) if (!std::strcmp(name, c)) return MPASSIT_CLASS_2D_NSTD;
    return MPASSIT_CLASS_2D_PATCH;
}

int mpassit_classify_hist_3d(const char *name, int wrf_mod_vars) {
    // input_data.F90:896-911 (nzp1_vars / vert_vars :842-843)
    if (wrf_mod_vars && !std::strcmp(name, "uReconstructZonal")) return MPASSIT_CLASS_U;
    if (wrf_mod_vars && !std::strcmp(name, "uReconstructMeridional")) return MPASSIT_CLASS_V;
    if (!std::strcmp(name, "zgrid") || !std::strcmp(name, "w")) return MPASSIT_CLASS_3D_NZP1;
    if (!std::strcmp(name, "vorticity")) return MPASSIT_CLASS_3D_VERT;
    return MPASSIT_CLASS_3D_NZ;
}

int mpassit_classify_diag(const char *name) {
    return !std::strcmp(name, "refl10cm") ? MPASSIT_CLASS_DIAG_3D : MPASSIT_CLASS_DIAG_2D;  // input_data.F90:283
}

void mpassit_para_range(int32_t n1, int32_t n2, int32_t nprocs, int32_t irank, int32_t *ista, int32_t *iend) {
    int32_t iwork1 = (n2 - n1 + 1) / nprocs;
    int32_t iwork2 = (n2 - n1 + 1) % nprocs;
    *ista = irank * iwork1 + n1 + std::min(irank, iwork2);
    *iend = *ista + iwork1 - 1;
    if (iwork2 > irank) *iend = *iend + 1;
}

int mpassit_read_block_decomp_file(const char *file, int32_t ncells, int32_t npets, int32_t *owner, char *err,
                                   size_t errlen) {
    std::ifstream in(file ? file : "");
    if (!in) { set_err(err, errlen, "BLOCK DECOMP FILE DOES NOT EXIST"); return 1; }
    std::string line;
    int32_t n = 0, pmax = 0;
    while (std::getline(in, line)) {
        if (trim(line).empty()) continue;
        int32_t p;
        if (!parse_int(line, &p)) { set_err(err, errlen, "READING BLOCK DECOMPOSITION FILE"); return 2; }
        if (n < ncells && owner) owner[n] = p;
        pmax = std::max(pmax, p);
        ++n;
    }
    if (n != ncells) {  // model_grid.F90:2401
        set_err(err, errlen, "BLOCK DECOMPOSITION FILE CONTAINS MORE CELLS THAN INPUT GRID");
        return 3;
    }
    if (pmax + 1 != npets) {  // model_grid.F90:2418-2421
        char b[256];
        std::snprintf(b, sizeof b, "BLOCK DECOMPOSITION FILE GENERATED FOR %10d PROCESSES BUT \n%10d PROCESSORS USED.",
                      pmax + 1, npets);
        set_err(err, errlen, b);
        return 4;
    }
    return 0;
}

}  // extern "C"
