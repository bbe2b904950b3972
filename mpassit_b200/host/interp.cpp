// interp.cpp -- host mirror of the reference's hot path, interp.F90:92-465
// (interp_data -> interp_diag_data / interp_hist_data) on top of the engine's
// C ABI.  Same call order, same regrid classes, same quirks:
//   * `method` is a carried variable: set to BILINEAR only when a 2-D patch
//     variable exists (interp.F90:203-204), to CONSERVE by the cons bundle (:370)
//     and NEAREST_STOD by the nstd bundle (:420); the soil bundle uses whatever
//     it holds at that point (:436-443).
//   * winds: cell-centre u,v -> mass points (:256-289), rotation at mass points
//     (:291-293, LC only), then mass -> EDGE1 / EDGE2 by a grid-to-grid bilinear
//     (:295-328).
// What differs by design: fields that share a route are stacked into ONE batched
// apply, and each distinct weight matrix is generated once (mprg_store memoises),
// instead of the reference's 12 RegridStore calls.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mpassit_host.h"

namespace {

struct Fail {
    int rc;
    std::string msg;
};

void ck(mprg_ctx *ctx, int rc, const char *where) {
    // reference pattern: if (ESMF_logFoundError(rc)) call error_handler("IN <where>", rc)
    if (rc != 0) throw Fail{rc, std::string("IN ") + where + ": " + mprg_last_error(ctx)};
}

struct Batch {
    std::vector<const void *> src;
    std::vector<void *> dst;
    std::vector<int32_t> nlev, epi;
    std::vector<double> earg;
    bool any_epi = false;
    void add(const void *s, void *d, int32_t n, int32_t op = MPRG_EPI_NONE) {
        src.push_back(s); dst.push_back(d); nlev.push_back(n); epi.push_back(op); earg.push_back(0.0);
        any_epi = any_epi || op != MPRG_EPI_NONE;
    }
    bool empty() const { return src.empty(); }
};

void run(mprg_ctx *ctx, mprg_route *rh, Batch &b, int sdt, int smem, int ddt, int dmem, const char *where,
         bool into_full = false) {
    if (b.empty()) return;
    const int32_t n = (int32_t)b.src.size();
    if (into_full) {
        ck(ctx, mprg_apply_into(ctx, rh, n, b.src.data(), b.nlev.data(), sdt, smem, b.dst.data(), ddt,
                                b.any_epi ? b.epi.data() : nullptr, b.any_epi ? b.earg.data() : nullptr), where);
        b = Batch();
        return;
    }
    ck(ctx, b.any_epi ? mprg_apply_ex(ctx, rh, n, b.src.data(), b.nlev.data(), sdt, smem, b.dst.data(), ddt, dmem,
                                      b.epi.data(), b.earg.data())
                      : mprg_apply(ctx, rh, n, b.src.data(), b.nlev.data(), sdt, smem, b.dst.data(), ddt, dmem),
       where);
    b = Batch();
}

// ESMF_REGRIDMETHOD flags carried in `method`; -1 = never assigned (Fortran: undefined)
constexpr int kUnset = -1;

}  // namespace

extern "C" {

int mpassit_classify_fields(const mpassit_config *cfg, mpassit_interp_io *io, int32_t *do_u, int32_t *do_v,
                            int32_t *u10_ind, int32_t *v10_ind) {
    if (!cfg || !io) return 1;
    int32_t du = 0, dv = 0, u10 = -1, v10 = -1;
    for (int i = 0; i < io->n_diag; ++i) {
        io->diag[i].klass = mpassit_classify_diag(io->diag[i].name);
        if (!std::strcmp(io->diag[i].name, "u10")) u10 = i;  // input_data.F90:173-180
        if (!std::strcmp(io->diag[i].name, "v10")) v10 = i;
    }
    for (int i = 0; i < io->n_hist_2d; ++i) io->hist_2d[i].klass = mpassit_classify_hist_2d(io->hist_2d[i].name);
    for (int i = 0; i < io->n_hist_3d; ++i) {
        io->hist_3d[i].klass = mpassit_classify_hist_3d(io->hist_3d[i].name, cfg->wrf_mod_vars);
        if (io->hist_3d[i].klass == MPASSIT_CLASS_U) du = 1;
        if (io->hist_3d[i].klass == MPASSIT_CLASS_V) dv = 1;
    }
    for (int i = 0; i < io->n_soil; ++i) io->soil[i].klass = MPASSIT_CLASS_SOIL;
    if (do_u) *do_u = du;
    if (do_v) *do_v = dv;
    if (u10_ind) *u10_ind = u10;
    if (v10_ind) *v10_ind = v10;
    return 0;
}

int mpassit_interp_data(mprg_ctx *ctx, const mpassit_config *cfg, mpassit_interp_io *io, char *err, size_t errlen) {
    if (!ctx || !cfg || !io) return 1;
    void *d_um = nullptr, *d_vm = nullptr;
    std::vector<mprg_route *> held;
    int rc_out = 0;
    // Host-buffer passes run with asynchronous applies: every Store/Regrid pair below is only QUEUED
    // (H2D, kernels and D2H of consecutive bundles overlap, and weight generation runs beside them on
    // its own stream); interp_data as a whole stays blocking, like the reference's.
    const int was_async = mprg_get_async(ctx);
    const bool host_pass = io->mem == MPRG_HOST;
    const int dmem = io->dst_device ? MPRG_DEVICE : io->mem;  // where the outputs live (sources: io->mem)
    const bool into = io->dst_full != 0;  // outputs are full-grid fields (own or mapped): gather fused into the store
    if (into && dmem == MPRG_HOST) {
        if (err && errlen) std::snprintf(err, errlen, "IN interp_data: dst_full needs device buffers");
        return 1;
    }
    if (host_pass) mprg_set_async(ctx, 1);
    try {
        // the target grid's topology decides how its centres act as a regrid SOURCE (centre -> edge staggering):
        // ESMF_GridCreate1PeriDim + MONOPOLE for global targets, GridCreateNoPeriDim otherwise (model_grid.F90:684-703)
        ck(ctx, mprg_set_grid_kind(ctx, cfg->is_regional ? MPRG_GRID_NOPERI : MPRG_GRID_1PERI_MONOPOLE), "GridCreate");
        int32_t do_u = 0, do_v = 0, u10 = -1, v10 = -1;
        mpassit_classify_fields(cfg, io, &do_u, &do_v, &u10, &v10);
        const int sdt = io->src_dtype, ddt = io->dst_dtype, mem = io->mem;
        auto store = [&](int method, int src_loc, int stagger, const char *where) {
            mprg_route *rh = nullptr;
            ck(ctx, mprg_store(ctx, method, src_loc, stagger, &rh), where);
            held.push_back(rh);
            return rh;
        };
        // rotation applies on Lambert targets only (interp.F90:138,291: proj_code==PROJ_LC)
        const bool rotate = cfg->proj_code == MPASSIT_PROJ_LC;
        if (rotate && !mprg_has_rotation(ctx))
            throw Fail{41, "IN rotate_winds_cgrid: cosa/sina not registered (mprg_set_rotation)"};

        // ---------------- interp_diag_data, interp.F90:107-141 ----------------
        // The diag bundle uses the same bilinear element->CENTER matrix as the hist 2d_patch / hgt /
        // 3d bundles.  When both stages run and the hist stage's `method` is BILINEAR, their fields are
        // stacked into ONE apply (independent fields: order of evaluation does not change any value).
        Batch diagBatch;
        const bool have_diag = cfg->interp_diag && io->n_diag > 0;
        // 10-m winds are rotated after the bundle regrid (interp.F90:138-139).  With host buffers they are
        // regridded into device scratch, rotated there and downloaded, so nothing waits on a round trip.
        const bool rot10 = have_diag && u10 >= 0 && v10 >= 0 && rotate;
        const bool rot10_dev = rot10 && (dmem == MPRG_HOST || into);
        if (have_diag)
            for (int i = 0; i < io->n_diag; ++i)
                if (!(rot10_dev && (i == u10 || i == v10)))
                    diagBatch.add(io->diag[i].src, io->diag[i].dst, io->diag[i].nlev);
        auto finish_diag = [&](mprg_route *rh) {
            if (!rot10) return;
            if (!rot10_dev) {
                ck(ctx, mprg_rotate_winds(ctx, io->diag[u10].dst, io->diag[v10].dst, 1, ddt, dmem), "rotate_winds_cgrid");
                return;
            }
            int32_t j0 = 0, j1 = 0, ni = 0, nj = 0;
            mpassit_target_dims(cfg, MPRG_CENTER, &ni, &nj);
            ck(ctx, mprg_get_slab(ctx, MPRG_CENTER, &j0, &j1), "get_slab");
            const size_t bytes = (size_t)(j1 - j0) * ni * (ddt == MPRG_F64 ? 8 : 4);
            void *du = nullptr, *dv = nullptr;
            ck(ctx, mprg_scratch(ctx, 2, bytes, &du), "scratch");
            ck(ctx, mprg_scratch(ctx, 3, bytes, &dv), "scratch");
            const void *s2[2] = {io->diag[u10].src, io->diag[v10].src};
            void *d2[2] = {du, dv};
            int32_t n2[2] = {1, 1};
            ck(ctx, mprg_apply(ctx, rh, 2, s2, n2, sdt, mem, d2, ddt, MPRG_DEVICE), "FieldBundleRegrid");
            ck(ctx, mprg_rotate_winds(ctx, du, dv, 1, ddt, MPRG_DEVICE), "rotate_winds_cgrid");
            if (into) {
                ck(ctx, mprg_put_slab(ctx, MPRG_CENTER, 1, ddt, du, io->diag[u10].dst), "put_slab");
                ck(ctx, mprg_put_slab(ctx, MPRG_CENTER, 1, ddt, dv, io->diag[v10].dst), "put_slab");
            } else {
                ck(ctx, mprg_download(ctx, du, io->diag[u10].dst, bytes), "download");
                ck(ctx, mprg_download(ctx, dv, io->diag[v10].dst, bytes), "download");
            }
        };
        if (have_diag && !cfg->interp_hist) {
            mprg_route *rh = store(MPRG_BILINEAR, MPRG_SRC_MESH_ELEMENT, MPRG_CENTER, "FieldBundleRegridStore");
            run(ctx, rh, diagBatch, sdt, mem, ddt, dmem, "FieldBundleRegrid", into);
            finish_diag(rh);
        }

        // ---------------- interp_hist_data, interp.F90:183-465 ----------------
        if (cfg->interp_hist) {
            int method = kUnset;
            int n2p = 0, n2c = 0, n2n = 0, n3 = 0, n3p1 = 0, n3v = 0;
            for (int i = 0; i < io->n_hist_2d; ++i) {
                n2p += io->hist_2d[i].klass == MPASSIT_CLASS_2D_PATCH;
                n2c += io->hist_2d[i].klass == MPASSIT_CLASS_2D_CONS;
                n2n += io->hist_2d[i].klass == MPASSIT_CLASS_2D_NSTD;
            }
            for (int i = 0; i < io->n_hist_3d; ++i) {
                n3 += io->hist_3d[i].klass == MPASSIT_CLASS_3D_NZ;
                n3p1 += io->hist_3d[i].klass == MPASSIT_CLASS_3D_NZP1;
                n3v += io->hist_3d[i].klass == MPASSIT_CLASS_3D_VERT;
            }
            if (n2p > 0) method = MPRG_BILINEAR;  // interp.F90:203-204
            // interp.F90:226-358 use `method` unconditionally; Fortran leaves it undefined when no
            // 2-D patch variable is listed.  ESMF_REGRIDMETHOD_BILINEAR is the zero flag, which is
            // what zero-initialised storage yields, so that is what the mirror assumes.
            const int m_bil = method == kUnset ? MPRG_BILINEAR : method;

            // mass-point winds (UMASS / VMASS, interp.F90:256-289) are intermediates that never leave the
            // device.  Default: kept in the output precision (the rotation itself runs in fp64 registers);
            // MPASSIT_GPU_ACC=f64 keeps them as the reference's R8 fields.  Either is far inside the 1e-5
            // contract; the fp32 chain moves half the bytes.
            const mpassit_field *fu = nullptr, *fv = nullptr;
            for (int i = 0; i < io->n_hist_3d; ++i) {
                if (io->hist_3d[i].klass == MPASSIT_CLASS_U) fu = &io->hist_3d[i];
                if (io->hist_3d[i].klass == MPASSIT_CLASS_V) fv = &io->hist_3d[i];
            }
            char accv[16] = "";
            mprg_get_option(ctx, "accumulate", accv, sizeof accv);
            const bool chain64 = !std::strcmp(accv, "f64");
            const int chain_dt = chain64 ? MPRG_F64 : ddt;
            const size_t chain_sz = chain_dt == MPRG_F64 ? 8 : 4;
            // Mass-point winds live on MPRG_CENTER_HALO rows: this rank's CENTER slab plus the one row
            // either side that its EDGE1 / EDGE2 points interpolate from (the reference gets those rows
            // through ESMF's halo exchange; recomputing them from the replicated mesh needs no exchange).
            // The whole wind chain as one matrix per staggered grid (mprg_store_wind): cell-centre (u, v) -> rotated U on
            // EDGE1 / V on EDGE2 directly, no mass-point intermediates.  Taken when the sources are device buffers with
            // 16-byte-aligned columns, the grid is composable (regional, rotation set) and both staggered outputs are
            // wanted; otherwise the reference's three steps run one after the other below.
            mprg_route *cmp_u = nullptr, *cmp_v = nullptr;
            if (fu && fv && rotate && fu->nlev == fv->nlev && io->u_stag && io->v_stag && mem == MPRG_DEVICE && dmem == MPRG_DEVICE &&
                ((size_t)fu->nlev * (sdt == MPRG_F64 ? 8 : 4)) % 16 == 0 && (uintptr_t)fu->src % 16 == 0 && (uintptr_t)fv->src % 16 == 0) {
                ck(ctx, mprg_store_wind(ctx, MPRG_EDGE1, &cmp_u), "FieldRegridStore");
                if (cmp_u) {
                    held.push_back(cmp_u);
                    ck(ctx, mprg_store_wind(ctx, MPRG_EDGE2, &cmp_v), "FieldRegridStore");
                    if (cmp_v) held.push_back(cmp_v);
                }
            }
            const bool composed = cmp_u && cmp_v;
            bool halo_is_center = true;
            if ((fu || fv) && !composed) {
                int32_t j0 = 0, j1 = 0, c0 = 0, c1 = 0, ni = 0, nj = 0;
                mpassit_target_dims(cfg, MPRG_CENTER, &ni, &nj);
                ck(ctx, mprg_get_slab(ctx, MPRG_CENTER_HALO, &j0, &j1), "get_slab");
                ck(ctx, mprg_get_slab(ctx, MPRG_CENTER, &c0, &c1), "get_slab");
                halo_is_center = (j0 == c0 && j1 == c1);
                const size_t nslab = (size_t)(j1 - j0) * ni;
                if (fu) ck(ctx, mprg_scratch(ctx, 0, nslab * fu->nlev * chain_sz, &d_um), "scratch");
                if (fv) ck(ctx, mprg_scratch(ctx, 1, nslab * fv->nlev * chain_sz, &d_vm), "scratch");
            }
            // the winds join the main stacked apply when its destinations are device buffers of the same type
            // rotate_winds_cgrid (interp.F90:291-293) is fused into the store of the (u, v) pair
            const bool fuse_rot = fu && fv && rotate && fu->nlev == fv->nlev;
            const int op_u = fuse_rot ? MPRG_EPI_ROT_U : MPRG_EPI_NONE, op_v = fuse_rot ? MPRG_EPI_ROT_V : MPRG_EPI_NONE;
            const bool winds_in_batch = !composed && (fu || fv) && mem == MPRG_DEVICE && chain_dt == ddt && m_bil == MPRG_BILINEAR && halo_is_center && !into;

            // one stacked apply for everything on the bilinear element->CENTER route:
            // 2d_patch (:207-221), hgt (:226-238), 3d_nz (:240-254), 3d_nzp1 (:331-347)
            {
                Batch b;
                if (have_diag && m_bil == MPRG_BILINEAR) {
                    b = diagBatch;
                } else if (have_diag) {  // cannot happen with the reference's method sequence; kept for safety
                    mprg_route *rd = store(MPRG_BILINEAR, MPRG_SRC_MESH_ELEMENT, MPRG_CENTER, "FieldBundleRegridStore");
                    run(ctx, rd, diagBatch, sdt, mem, ddt, dmem, "FieldBundleRegrid", into);
                    finish_diag(rd);
                }
                for (int i = 0; i < io->n_hist_2d; ++i)
                    if (io->hist_2d[i].klass == MPASSIT_CLASS_2D_PATCH) b.add(io->hist_2d[i].src, io->hist_2d[i].dst, 1);
                if (io->ter && io->hgt) b.add(io->ter, io->hgt, 1);
                if (winds_in_batch) {
                    if (fu) b.add(fu->src, d_um, fu->nlev, op_u);
                    if (fv) b.add(fv->src, d_vm, fv->nlev, op_v);
                }
                for (int i = 0; i < io->n_hist_3d; ++i)
                    if (io->hist_3d[i].klass == MPASSIT_CLASS_3D_NZ || io->hist_3d[i].klass == MPASSIT_CLASS_3D_NZP1)
                        b.add(io->hist_3d[i].src, io->hist_3d[i].dst, io->hist_3d[i].nlev);
                if (!b.empty() || (have_diag && m_bil == MPRG_BILINEAR)) {
                    mprg_route *rh = store(m_bil, MPRG_SRC_MESH_ELEMENT, MPRG_CENTER, "FieldBundleRegridStore");
                    run(ctx, rh, b, sdt, mem, ddt, dmem, "FieldBundleRegrid", into);
                    if (have_diag && m_bil == MPRG_BILINEAR) finish_diag(rh);
                }
            }

            // winds, interp.F90:256-328
            if (composed) {
                ck(ctx, mprg_apply_wind(ctx, cmp_u, fu->src, fv->src, fu->nlev, sdt, io->u_stag, ddt, into ? 1 : 0), "FieldRegrid");
                ck(ctx, mprg_apply_wind(ctx, cmp_v, fu->src, fv->src, fu->nlev, sdt, io->v_stag, ddt, into ? 1 : 0), "FieldRegrid");
            } else if (do_u || do_v) {
                if (!winds_in_batch) {
                    mprg_route *rh = store(m_bil, MPRG_SRC_MESH_ELEMENT, MPRG_CENTER_HALO, "FieldRegridStore");
                    Batch b;
                    if (fu) b.add(fu->src, d_um, fu->nlev, op_u);
                    if (fv) b.add(fv->src, d_vm, fv->nlev, op_v);
                    run(ctx, rh, b, sdt, mem, chain_dt, MPRG_DEVICE, "FieldRegrid");
                }
                if (fu && fv && rotate && !fuse_rot)  // interp.F90:291-293
                    ck(ctx, mprg_rotate_winds_on(ctx, MPRG_CENTER_HALO, d_um, d_vm, fu->nlev, chain_dt, MPRG_DEVICE), "rotate_winds_cgrid");
                if (fu && io->u_stag) {  // interp.F90:295-311
                    mprg_route *ru = store(m_bil, MPRG_SRC_GRID_CENTER, MPRG_EDGE1, "FieldRegridStore");
                    const void *s = d_um; void *d = io->u_stag; int32_t nl = fu->nlev;
                    ck(ctx, into ? mprg_apply_into(ctx, ru, 1, &s, &nl, chain_dt, MPRG_DEVICE, &d, ddt, nullptr, nullptr)
                                 : mprg_apply(ctx, ru, 1, &s, &nl, chain_dt, MPRG_DEVICE, &d, ddt, dmem), "FieldRegrid");
                }
                if (fv && io->v_stag) {  // interp.F90:313-328
                    mprg_route *rv = store(m_bil, MPRG_SRC_GRID_CENTER, MPRG_EDGE2, "FieldRegridStore");
                    const void *s = d_vm; void *d = io->v_stag; int32_t nl = fv->nlev;
                    ck(ctx, into ? mprg_apply_into(ctx, rv, 1, &s, &nl, chain_dt, MPRG_DEVICE, &d, ddt, nullptr, nullptr)
                                 : mprg_apply(ctx, rv, 1, &s, &nl, chain_dt, MPRG_DEVICE, &d, ddt, dmem), "FieldRegrid");
                }
            }

            // 3d_vert bundle (source on mesh nodes), interp.F90:350-366
            if (n3v > 0) {
                mprg_route *rh = store(m_bil, MPRG_SRC_MESH_NODE, MPRG_CENTER, "FieldBundleRegridStore");
                Batch b;
                for (int i = 0; i < io->n_hist_3d; ++i)
                    if (io->hist_3d[i].klass == MPASSIT_CLASS_3D_VERT)
                        b.add(io->hist_3d[i].src, io->hist_3d[i].dst, io->hist_3d[i].nlev);
                run(ctx, rh, b, sdt, mem, ddt, dmem, "FieldBundleRegrid", into);
            }
            // 2d_cons bundle, interp.F90:368-416 (bundle and per-field paths apply the same matrix)
            if (n2c > 0) {
                method = MPRG_CONSERVE;
                mprg_route *rh = store(method, MPRG_SRC_MESH_ELEMENT, MPRG_CENTER, "FieldBundleRegridStore");
                Batch b;
                for (int i = 0; i < io->n_hist_2d; ++i)
                    if (io->hist_2d[i].klass == MPASSIT_CLASS_2D_CONS) b.add(io->hist_2d[i].src, io->hist_2d[i].dst, 1);
                run(ctx, rh, b, sdt, mem, ddt, dmem, "FieldBundleRegrid", into);
            }
            // 2d_nstd bundle, interp.F90:418-434
            bool soil_done = false;
            if (n2n > 0) {
                method = MPRG_NEAREST_STOD;
                mprg_route *rh = store(method, MPRG_SRC_MESH_ELEMENT, MPRG_CENTER, "FieldBundleRegridStore");
                Batch b;
                for (int i = 0; i < io->n_hist_2d; ++i)
                    if (io->hist_2d[i].klass == MPASSIT_CLASS_2D_NSTD) b.add(io->hist_2d[i].src, io->hist_2d[i].dst, 1);
                // the soil bundle inherits this `method` (:436-443): same matrix, so its fields ride the same apply
                for (int i = 0; i < io->n_soil; ++i) b.add(io->soil[i].src, io->soil[i].dst, io->soil[i].nlev);
                soil_done = true;
                run(ctx, rh, b, sdt, mem, ddt, dmem, "FieldBundleRegrid", into);
            }
            // soil bundle: whatever `method` holds now, interp.F90:436-447
            if (io->n_soil > 0 && !soil_done) {
                const int m_soil = method == kUnset ? MPRG_BILINEAR : method;
                mprg_route *rh = store(m_soil, MPRG_SRC_MESH_ELEMENT, MPRG_CENTER, "FieldBundleRegridStore");
                Batch b;
                for (int i = 0; i < io->n_soil; ++i) b.add(io->soil[i].src, io->soil[i].dst, io->soil[i].nlev);
                run(ctx, rh, b, sdt, mem, ddt, dmem, "FieldBundleRegrid", into);
            }
        }
    } catch (const Fail &f) {
        if (err && errlen) std::snprintf(err, errlen, "%s", f.msg.c_str());
        rc_out = f.rc ? f.rc : 1;
    }
    // interp_data returns with every output in place
    if (host_pass) {
        if (mprg_synchronize(ctx) != 0 && rc_out == 0) {
            if (err && errlen) std::snprintf(err, errlen, "IN interp_data: %s", mprg_last_error(ctx));
            rc_out = 1;
        }
        mprg_set_async(ctx, was_async);
    }
    // interp.F90:449-464 FieldBundleRegridRelease (here: every handle taken above)
    for (mprg_route *rh : held) mprg_release(ctx, rh);
    return rc_out;
}

}  // extern "C"
