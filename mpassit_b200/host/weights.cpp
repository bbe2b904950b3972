// weights.cpp -- ESMF regrid weight files (the NetCDF file `ESMF_RegridWeightGen -w` and
// ESMF_FieldRegridStore + ESMF_SparseMatrixWrite produce): dimensions n_a (source size), n_b (destination size),
// n_s (stored weights); variables col(n_s), row(n_s) (1-based source / destination sequence indices) and S(n_s).
//
// Why it is here: the reference's arithmetic lives in ESMF, which this image does not have (parity
// is pinned to the oracle, not to ESMF).  A maintainer with ESMF can dump the matrix the reference builds -- or run
// ESMF_RegridWeightGen on the mesh / grid files written by tools/esmf_kit.py -- and compare it with the engine's
// route entry by entry (mprg_route_export_csr), or run the engine ON ESMF's own weights (mprg_route_import_csr).
// Classic-format files only (ESMF's default output; `--netcdf4` files must be converted with nccopy).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/mpassit_host.h"
#include "ncio.hpp"

namespace {
void put_err(char *err, size_t n, const std::string &m) {
    if (err && n) std::snprintf(err, n, "%s", m.c_str());
}
bool open_weights(const char *path, ncio::Reader &r, int64_t *n_a, int64_t *n_b, int64_t *n_s, std::string &why) {
    if (!path || !r.open(path, why)) return false;
    const ncio::Dim *a = r.dim("n_a"), *b = r.dim("n_b"), *s = r.dim("n_s");
    if (!a || !b || !s) {
        why = std::string(path) + ": not an ESMF weight file (dimensions n_a, n_b, n_s)";
        return false;
    }
    for (const char *v : {"col", "row", "S"})
        if (!r.var(v) || r.count(*r.var(v)) != s->len) {
            why = std::string(path) + ": not an ESMF weight file (variable " + v + "(n_s))";
            return false;
        }
    *n_a = (int64_t)a->len;
    *n_b = (int64_t)b->len;
    *n_s = (int64_t)s->len;
    return true;
}
}  // namespace

extern "C" {

int mpassit_weights_sizes(const char *path, int64_t *n_a, int64_t *n_b, int64_t *n_s, char *err, size_t errlen) {
    ncio::Reader r;
    std::string why;
    if (!n_a || !n_b || !n_s || !open_weights(path, r, n_a, n_b, n_s, why)) {
        put_err(err, errlen, why);
        return 1;
    }
    return 0;
}

int mpassit_weights_read_csr(const char *path, int64_t n_b, int64_t n_s, int32_t *rowptr, int32_t *col, double *w,
                             char *err, size_t errlen) {
    ncio::Reader r;
    std::string why;
    int64_t na = 0, nb = 0, ns = 0;
    if (!rowptr || !open_weights(path, r, &na, &nb, &ns, why)) {
        put_err(err, errlen, why);
        return 1;
    }
    if (nb != n_b || ns != n_s) {
        put_err(err, errlen, std::string(path) + ": sizes differ from the caller's (call mpassit_weights_sizes first)");
        return 2;
    }
    std::vector<int32_t> rr(ns), cc(ns);
    std::vector<double> ss(ns);
    if (ns > 0 && (!r.read_ints(*r.var("row"), 0, ns, rr.data(), why) || !r.read_ints(*r.var("col"), 0, ns, cc.data(), why) ||
                   !r.read_doubles(*r.var("S"), 0, ns, ss.data(), why))) {
        put_err(err, errlen, why);
        return 3;
    }
    // ESMF lists the entries in no promised order: counting sort by destination row, stable (keeps ESMF's order
    // inside a row, which is the order its apply sums in)
    std::fill(rowptr, rowptr + nb + 1, 0);
    for (int64_t k = 0; k < ns; ++k) {
        if (rr[k] < 1 || rr[k] > nb || cc[k] < 1 || cc[k] > na) {
            put_err(err, errlen, std::string(path) + ": row / col index out of range at entry " + std::to_string(k));
            return 4;
        }
        rowptr[rr[k]]++;
    }
    for (int64_t i = 0; i < nb; ++i) rowptr[i + 1] += rowptr[i];
    std::vector<int32_t> fill(rowptr, rowptr + nb);
    for (int64_t k = 0; k < ns; ++k) {
        const int32_t at = fill[rr[k] - 1]++;
        col[at] = cc[k] - 1;  // 0-based, like mprg_route_export_csr / mprg_route_import_csr
        w[at] = ss[k];
    }
    return 0;
}

int mpassit_weights_write(const char *path, int64_t n_a, int64_t n_b, const int32_t *rowptr, const int32_t *col,
                          const double *w, const char *method, char *err, size_t errlen) {
    if (!path || !rowptr || n_b < 0) {
        put_err(err, errlen, "mpassit_weights_write: null argument");
        return 1;
    }
    const int64_t ns = rowptr[n_b];
    ncio::Writer wr(0);
    const int dA = wr.def_dim("n_a", (uint64_t)n_a), dB = wr.def_dim("n_b", (uint64_t)n_b);
    (void)dA;
    (void)dB;
    const int dS = wr.def_dim("n_s", (uint64_t)std::max<int64_t>(ns, 1));
    wr.att_text(-1, "title", "ESMF Offline Regridding Weight Generator");
    wr.att_text(-1, "normalization", "destarea");
    wr.att_text(-1, "map_method", method ? method : "Bilinear remapping");
    wr.att_text(-1, "ESMF_regrid_method", method ? method : "Bilinear");
    wr.att_text(-1, "conventions", "NCAR-CSM");
    wr.att_text(-1, "source", "mpassit-b200 (mprg_route_export_csr)");
    const int vcol = wr.def_var("col", ncio::NC_INT, {dS}), vrow = wr.def_var("row", ncio::NC_INT, {dS}),
              vS = wr.def_var("S", ncio::NC_DOUBLE, {dS});
    std::string why;
    if (!wr.enddef(path, 0, true, why)) {
        put_err(err, errlen, why);
        return 2;
    }
    if (ns == 0) {  // no zero-length fixed dimension in the classic format: one zero weight
        const int32_t one = 1;
        const double zero = 0.0;
        if (!wr.put_ints(vcol, &one, 1, why) || !wr.put_ints(vrow, &one, 1, why) || !wr.put_doubles(vS, &zero, 1, why)) {
            put_err(err, errlen, why);
            return 3;
        }
    } else {
        std::vector<int32_t> rr(ns), cc(ns);
        for (int64_t i = 0; i < n_b; ++i)
            for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
                rr[k] = (int32_t)(i + 1);
                cc[k] = col[k] + 1;
            }
        if (!wr.put_ints(vcol, cc.data(), ns, why) || !wr.put_ints(vrow, rr.data(), ns, why) || !wr.put_doubles(vS, w, ns, why)) {
            put_err(err, errlen, why);
            return 3;
        }
    }
    if (!wr.close(why)) {
        put_err(err, errlen, why);
        return 4;
    }
    return 0;
}

}  // extern "C"
