// gather.cu -- collect every rank's destination slab on the writing rank.
// Replaces the 20 ESMF_FieldGather(rootPet=0) calls of write_data.F90:1006-1453.
//
// The regridding itself needs no exchange (each rank owns whole destination
// rows); this is the path's only collective.  A slab of a [lev][nj][ni] field
// is nlev contiguous runs of the full field, so the gather is nlev grouped
// ncclSend/ncclRecv pairs per peer written straight into place (no re-tiling
// pass).  NCCL is resolved with dlopen so single-GPU hosts need no libnccl.
#include <dlfcn.h>

#include "common.cuh"

namespace mprg {
namespace {

struct NcclId { char bytes[128]; };  // ncclUniqueId
typedef int (*fn_getid)(NcclId *);
typedef int (*fn_init)(void **, int, NcclId, int);
typedef int (*fn_send)(const void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_recv)(void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_void)(void);
typedef int (*fn_destroy)(void *);
typedef const char *(*fn_errstr)(int);
constexpr int kNcclInt8 = 0;  // ncclInt8 == ncclChar: counts below are bytes

void *lib(mprg_ctx *ctx) {
    if (ctx->ncclLib) return ctx->ncclLib;
    for (const char *n : {"libnccl.so.2", "libnccl.so"}) {
        ctx->ncclLib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (ctx->ncclLib) return ctx->ncclLib;
    }
    fail(81, "NCCL library not found: %s", dlerror());
}

template <typename T>
T sym(mprg_ctx *ctx, const char *name) {
    T f = (T)dlsym(lib(ctx), name);
    if (!f) fail(82, "NCCL symbol %s missing", name);
    return f;
}

void check(mprg_ctx *ctx, int rc, const char *what) {
    if (rc == 0) return;
    fn_errstr es = (fn_errstr)dlsym(lib(ctx), "ncclGetErrorString");
    fail(800 + rc, "%s failed: %s", what, es ? es(rc) : "unknown NCCL error");
}

}  // namespace

void comm_destroy(mprg_ctx *ctx) {
    if (ctx->nccl && ctx->ncclLib) {
        fn_destroy d = (fn_destroy)dlsym(ctx->ncclLib, "ncclCommDestroy");
        if (d) d(ctx->nccl);
    }
    ctx->nccl = nullptr;
}

void comm_id(mprg_ctx *ctx, void *id128) {
    check(ctx, sym<fn_getid>(ctx, "ncclGetUniqueId")((NcclId *)id128), "ncclGetUniqueId");
}

void comm_init(mprg_ctx *ctx, const void *id128) {
    if (ctx->nranks == 1) return;
    NcclId id;
    memcpy(&id, id128, sizeof id);
    check(ctx, sym<fn_init>(ctx, "ncclCommInitRank")(&ctx->nccl, ctx->nranks, id, ctx->rank), "ncclCommInitRank");
}

void gather_slabs(mprg_ctx *ctx, int nfields, const int *stagger, const int32_t *nlev, int dtype,
                  const void *const *slab_dev, int root, void *const *full_dev) {
    if (root < 0 || root >= ctx->nranks) fail(86, "mprg_gather: bad root %d", root);
    if (nfields <= 0) return;
    if (!stagger || !nlev || !slab_dev) fail(1, "mprg_gather: null argument");
    const size_t esz = dtype == MPRG_F32 ? 4 : 8;
    for (int f = 0; f < nfields; ++f) {
        if (stagger[f] < 0 || stagger[f] > 3 || !ctx->target[stagger[f]].set)
            fail(83, "mprg_gather: stagger %d not set", stagger[f]);
        const Target &tg = ctx->target[stagger[f]];
        if (ctx->rank == root && nlev[f] > 0 && (!full_dev || !full_dev[f])) fail(1, "mprg_gather: root needs a destination buffer");
        if (tg.nSlab() > 0 && nlev[f] > 0 && !slab_dev[f]) fail(1, "mprg_gather: null slab");
    }
    if (ctx->nranks > 1 && !ctx->nccl) fail(84, "mprg_gather: communicator not initialised (call mprg_comm_init)");

    // own slab: strided device copy (nlev runs of nMine elements)
    if (ctx->rank == root) {
        for (int f = 0; f < nfields; ++f) {
            const Target &tg = ctx->target[stagger[f]];
            const int64_t nFull = (int64_t)tg.ni * tg.nj, nMine = tg.nSlab();
            if (nMine > 0 && nlev[f] > 0)
                MPRG_CUDA(cudaMemcpy2DAsync((unsigned char *)full_dev[f] + (size_t)tg.slabOffset() * esz, (size_t)nFull * esz,
                                            slab_dev[f], (size_t)nMine * esz, (size_t)nMine * esz, (size_t)nlev[f],
                                            cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    if (ctx->nranks == 1) return;
    fn_void gstart = sym<fn_void>(ctx, "ncclGroupStart"), gend = sym<fn_void>(ctx, "ncclGroupEnd");
    fn_send send = sym<fn_send>(ctx, "ncclSend");
    fn_recv recv = sym<fn_recv>(ctx, "ncclRecv");
    check(ctx, gstart(), "ncclGroupStart");
    // inside the group nothing may throw: an open NCCL group would poison every later collective on the
    // communicator.  The first failure is remembered, the group is always closed, then the failure is reported.
    int bad = 0;
    const char *badWhat = nullptr;
    auto note = [&](int rc, const char *what) {
        if (rc != 0 && !bad) { bad = rc; badWhat = what; }
    };
    for (int f = 0; f < nfields && !bad; ++f) {
        const Target &tg = ctx->target[stagger[f]];
        const int64_t nFull = (int64_t)tg.ni * tg.nj, nMine = tg.nSlab();
        if (ctx->rank != root) {
            for (int l = 0; l < nlev[f] && nMine > 0 && !bad; ++l)
                note(send((const unsigned char *)slab_dev[f] + (size_t)l * nMine * esz, (size_t)nMine * esz, kNcclInt8,
                          root, ctx->nccl, ctx->stream), "ncclSend");
        } else {
            for (int p = 0; p < ctx->nranks; ++p) {
                if (p == root) continue;
                int32_t j0, j1;
                para_range(tg.nj, ctx->nranks, p, &j0, &j1);
                const int64_t ns = (int64_t)(j1 - j0) * tg.ni, off = (int64_t)j0 * tg.ni;
                for (int l = 0; l < nlev[f] && ns > 0 && !bad; ++l)
                    note(recv((unsigned char *)full_dev[f] + ((size_t)l * nFull + off) * esz, (size_t)ns * esz,
                              kNcclInt8, p, ctx->nccl, ctx->stream), "ncclRecv");
            }
        }
    }
    const int endrc = gend();
    if (bad) check(ctx, bad, badWhat);
    check(ctx, endrc, "ncclGroupEnd");
}

}  // namespace mprg
