// capi.cu -- extern "C" entry points declared in include/mpassit_rg.h.
// No exception crosses the boundary: every entry returns an rc and records a
// message retrievable with mprg_last_error().
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <chrono>
#include <cmath>
#include <thread>

#include "common.cuh"

using namespace mprg;

static std::string g_init_error;

// MPASSIT_TRACE=1 / mprg_set_option("trace", "1"): host wall time of every store / apply call on stderr, with the time
// since the trace was switched on (where an end-to-end pass goes)
static std::atomic<int> g_trace{-1};
static std::chrono::steady_clock::time_point g_trace_t0;
static bool trace_on() {
    int v = g_trace.load(std::memory_order_relaxed);
    if (v < 0) {
        v = getenv("MPASSIT_TRACE") != nullptr ? 1 : 0;
        g_trace_t0 = std::chrono::steady_clock::now();
        g_trace.store(v);
    }
    return v > 0;
}
struct Trace {
    const char *what;
    int a, b;
    std::chrono::steady_clock::time_point t0;
    bool on;
    Trace(const char *w, int a_ = 0, int b_ = 0) : what(w), a(a_), b(b_) {
        on = trace_on();
        if (on) t0 = std::chrono::steady_clock::now();
    }
    ~Trace() {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        const double at = std::chrono::duration<double, std::milli>(t0 - g_trace_t0).count();
        fprintf(stderr, "[mprg] +%9.3f ms  %-14s %3d %3d  %9.3f ms\n", at, what, a, b, ms);
    }
};

#define MPRG_ENTER(ctx)                                     \
    if (!(ctx)) return 1;                                   \
    try {                                                   \
        MPRG_CUDA(cudaSetDevice((ctx)->device));           \
        ::mprg::tl_stream = (ctx)->stream;

#define MPRG_LEAVE(ctx)                                     \
        return 0;                                           \
    } catch (const Error &e) {                              \
        (ctx)->err = e.msg;                                 \
        return e.rc ? e.rc : 1;                             \
    } catch (const std::exception &e) {                     \
        (ctx)->err = e.what();                              \
        return 2;                                           \
    } catch (...) {                                         \
        (ctx)->err = "unknown error";                       \
        return 3;                                           \
    }

// returns false for an unknown key / value
static bool set_option(mprg_ctx *c, const char *key, const char *val) {
    const std::string k = key ? key : "", v = val ? val : "";
    mprg_tuning &t = c->tune;
    if (k == "accumulate") {
        if (v == "f64" || v == "fp64") t.acc64 = true;
        else if (v == "f32" || v == "fp32" || v.empty()) t.acc64 = false;
        else return false;
    } else if (k == "apply") {
        t.pipeOff = v == "direct";
        if (v != "direct" && v != "pipe" && !v.empty()) return false;
    } else if (k == "trace") {
        g_trace_t0 = std::chrono::steady_clock::now();
        g_trace.store(atoi(v.c_str()) != 0 ? 1 : 0);
    } else if (k == "pipe_split") {
        t.pipeSplit = v.empty() || atoi(v.c_str()) != 0;
    } else if (k == "pipe_minb") {
        t.pipeMinb = atoi(v.c_str());
    } else if (k == "planes_shape") {
        t.planesShape = atoi(v.c_str());
    } else if (k == "cols_minb") {
        t.colsMinb = v.empty() ? 3 : atoi(v.c_str());
    } else if (k == "wind") {
        if (v == "chain") t.windChain = true;
        else if (v == "composed" || v.empty()) t.windChain = false;
        else return false;
    } else if (k == "upload_threads") {
        t.uploadThreads = std::max(0, atoi(v.c_str()));
    } else {
        return false;
    }
    return true;
}

extern "C" {

const char *mprg_version(void) { return "mpassit-rg 0.1 (sm_100a)"; }

int mprg_init(int device, int rank, int nranks, mprg_ctx **out) {
    if (!out) return 1;
    *out = nullptr;
    if (nranks < 1 || rank < 0 || rank >= nranks) { g_init_error = "mprg_init: bad rank/nranks"; return 4; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        // no CPU fallback: the engine is CUDA only
        g_init_error = std::string("mprg_init: no CUDA device (") + cudaGetErrorString(e) + ")";
        return 5;
    }
    if (device < 0 || device >= ndev) { g_init_error = "mprg_init: bad device ordinal"; return 6; }
    mprg_ctx *c = new mprg_ctx();
    c->device = device; c->rank = rank; c->nranks = nranks;
    try {
        MPRG_CUDA(cudaSetDevice(device));
        MPRG_CUDA(cudaDeviceGetAttribute(&c->numSM, cudaDevAttrMultiProcessorCount, device));
        // engine temporaries come from the stream-ordered pool and stay cached in it (common.cuh: DevBuf)
        cudaMemPool_t pool;
        MPRG_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long keep = ~0ULL;
        MPRG_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        MPRG_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
        MPRG_CUDA(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
        MPRG_CUDA(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
        MPRG_CUDA(cudaStreamCreateWithFlags(&c->store_stream, cudaStreamNonBlocking));
        c->stream = c->own_stream;
        MPRG_CUDA(cudaEventCreate(&c->ev0));
        MPRG_CUDA(cudaEventCreate(&c->ev1));
        MPRG_CUDA(cudaEventCreateWithFlags(&c->evDl, cudaEventDisableTiming));
        MPRG_CUDA(cudaMallocHost(&c->peekBuf, 256));
        if (const char *e = getenv("MPASSIT_GPU_ASYNC")) c->async = atoi(e) != 0;
        if (const char *e = getenv("MPASSIT_WEIGHT_CACHE")) c->cacheDir = e;
        // tuning knobs: the environment is read here, once; mprg_set_option changes them afterwards
        for (const char *const *kv = (const char *const[]){"MPASSIT_GPU_ACC", "accumulate", "MPASSIT_GPU_APPLY", "apply",
                                                           "MPASSIT_GPU_PIPE_MINB", "pipe_minb", "MPASSIT_GPU_PIPE_SPLIT", "pipe_split",
                                                           "MPASSIT_GPU_MINB", "cols_minb", "MPASSIT_GPU_PLANES", "planes_shape", "MPASSIT_GPU_WIND", "wind", "MPASSIT_UPLOAD_THREADS",
                                                           "upload_threads", nullptr};
             *kv; kv += 2)
            if (const char *e = getenv(kv[0])) set_option(c, kv[1], e);
        for (int i = 0; i < mprg_ctx::kSlots; ++i) {
            MPRG_CUDA(cudaEventCreateWithFlags(&c->evIn[i], cudaEventDisableTiming));
            MPRG_CUDA(cudaEventCreateWithFlags(&c->evK[i], cudaEventDisableTiming));
            MPRG_CUDA(cudaEventCreateWithFlags(&c->evOut[i], cudaEventDisableTiming));
        }
    } catch (const Error &er) {
        g_init_error = er.msg;
        delete c;
        return er.rc;
    }
    *out = c;
    return 0;
}

int mprg_host_bind_to_device(int device, int *node) {
    if (node) *node = -1;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return 6; }
    for (char *p = bus; *p; ++p) *p = (char)tolower(*p);
    int nd = -1;
    {
        std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
        FILE *f = fopen(path.c_str(), "r");
        if (f) { if (fscanf(f, "%d", &nd) != 1) nd = -1; fclose(f); }
    }
    if (nd < 0) return 0;   // unknown / single node: leave the process alone
    // the node's CPUs: /sys/devices/system/node/node<N>/cpulist, e.g. "0-31,64-95"
    cpu_set_t set;
    CPU_ZERO(&set);
    int ncpu = 0;
    {
        std::string path = "/sys/devices/system/node/node" + std::to_string(nd) + "/cpulist";
        FILE *f = fopen(path.c_str(), "r");
        if (f) {
            int a, b;
            char sep;
            while (fscanf(f, "%d", &a) == 1) {
                b = a;
                int c = fgetc(f);
                if (c == '-') { if (fscanf(f, "%d", &b) != 1) b = a; c = fgetc(f); }
                for (int k = a; k <= b && k < CPU_SETSIZE; ++k) { CPU_SET(k, &set); ++ncpu; }
                if (c != ',') break;
            }
            (void)sep;
            fclose(f);
        }
    }
    // keep only CPUs this process is allowed to use (containers, cgroups)
    cpu_set_t cur;
    if (sched_getaffinity(0, sizeof cur, &cur) == 0) {
        int left = 0;
        for (int k = 0; k < CPU_SETSIZE; ++k) {
            if (CPU_ISSET(k, &set) && !CPU_ISSET(k, &cur)) CPU_CLR(k, &set);
            if (CPU_ISSET(k, &set)) ++left;
        }
        ncpu = left;
    }
    if (ncpu > 0) sched_setaffinity(0, sizeof set, &set);
    // MPOL_PREFERRED (1) for this node: pages of later allocations (page-locked ones included) come from it when it has room
    unsigned long mask[16] = {0};
    if (nd < (int)(sizeof mask * 8)) {
        mask[nd / (8 * sizeof(unsigned long))] |= 1UL << (nd % (8 * sizeof(unsigned long)));
        syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, sizeof mask * 8);
    }
    if (node) *node = nd;
    return 0;
}

int mprg_finalize(mprg_ctx *ctx) {
    if (!ctx) return 1;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    mprg::tl_stream = nullptr;  // frees below are ordered on the default stream: the context's streams are going away
    for (auto &kv : ctx->routes) delete kv.second;
    for (auto *r : ctx->imported) delete r;
    ctx->routes.clear();
    ctx->imported.clear();
    mprg::comm_destroy(ctx);
    for (auto &kv : ctx->ipcOpen) cudaIpcCloseMemHandle(kv.second);
    ctx->ipcOpen.clear();
    if (ctx->evDl) cudaEventDestroy(ctx->evDl);
    if (ctx->bounce) cudaFreeHost(ctx->bounce);
    for (auto &e : ctx->evBounce)
        if (e) cudaEventDestroy(e);
    if (ctx->peekBuf) cudaFreeHost(ctx->peekBuf);
    if (ctx->store_stream) cudaStreamDestroy(ctx->store_stream);
    for (int i = 0; i < mprg_ctx::kSlots; ++i) {
        if (ctx->evIn[i]) cudaEventDestroy(ctx->evIn[i]);
        if (ctx->evK[i]) cudaEventDestroy(ctx->evK[i]);
        if (ctx->evOut[i]) cudaEventDestroy(ctx->evOut[i]);
    }
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    delete ctx;
    return 0;
}

const char *mprg_last_error(const mprg_ctx *ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

int mprg_set_stream(mprg_ctx *ctx, void *cuda_stream) {
    MPRG_ENTER(ctx)
    // pool allocations made so far become valid on any stream once the old one has drained
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    mprg::tl_stream = ctx->stream;
    MPRG_LEAVE(ctx)
}

int mprg_synchronize(mprg_ctx *ctx) {
    Trace tr("synchronize");
    MPRG_ENTER(ctx)
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
    MPRG_CUDA(cudaStreamSynchronize(ctx->h2d_stream));
    MPRG_CUDA(cudaStreamSynchronize(ctx->d2h_stream));
    MPRG_CUDA(cudaStreamSynchronize(ctx->store_stream));
    MPRG_LEAVE(ctx)
}

int mprg_host_alloc(mprg_ctx *ctx, size_t bytes, void **ptr) {
    MPRG_ENTER(ctx)
    if (!ptr) fail(1, "mprg_host_alloc: null out pointer");
    MPRG_CUDA(cudaMallocHost(ptr, bytes ? bytes : 1));
    MPRG_LEAVE(ctx)
}

int mprg_host_free(mprg_ctx *ctx, void *ptr) {
    MPRG_ENTER(ctx)
    if (ptr) MPRG_CUDA(cudaFreeHost(ptr));
    MPRG_LEAVE(ctx)
}

int mprg_device_alloc(mprg_ctx *ctx, size_t bytes, void **ptr) {
    MPRG_ENTER(ctx)
    if (!ptr) fail(1, "mprg_device_alloc: null out pointer");
    MPRG_CUDA(cudaMalloc(ptr, bytes ? bytes : 1));
    MPRG_LEAVE(ctx)
}

int mprg_device_free(mprg_ctx *ctx, void *ptr) {
    MPRG_ENTER(ctx)
    if (ptr) MPRG_CUDA(cudaFree(ptr));
    MPRG_LEAVE(ctx)
}

int mprg_scratch(mprg_ctx *ctx, int slot, size_t bytes, void **ptr) {
    MPRG_ENTER(ctx)
    if (!ptr || slot < 0 || slot >= 8) fail(1, "mprg_scratch: bad slot/pointer");
    if (bytes > ctx->userScratch[slot].n) {
        if (ctx->capturing) fail(45, "mprg_scratch: a slot cannot grow while capturing a graph");
        MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->userScratch[slot].alloc(bytes);
        MPRG_CUDA(cudaStreamSynchronize(ctx->stream));  // callers may use the block on their own streams
    }
    *ptr = ctx->userScratch[slot].p;
    MPRG_LEAVE(ctx)
}

int mprg_has_rotation(const mprg_ctx *ctx) { return ctx && ctx->haveRot ? 1 : 0; }

int mprg_set_weight_cache(mprg_ctx *ctx, const char *dir) {
    MPRG_ENTER(ctx)
    ctx->cacheDir = dir ? dir : "";
    MPRG_LEAVE(ctx)
}

int mprg_weight_cache_stats(const mprg_ctx *ctx, int64_t *hits, int64_t *stores) {
    if (!ctx) return 1;
    if (hits) *hits = ctx->cacheHits;
    if (stores) *stores = ctx->cacheStores;
    return 0;
}

int mprg_set_option(mprg_ctx *ctx, const char *key, const char *value) {
    MPRG_ENTER(ctx)
    if (ctx->capturing) fail(45, "mprg_set_option: not while capturing a graph");
    if (!set_option(ctx, key, value)) fail(59, "mprg_set_option: unknown option %s=%s", key ? key : "(null)", value ? value : "(null)");
    MPRG_LEAVE(ctx)
}

int mprg_get_option(const mprg_ctx *ctx, const char *key, char *value, size_t len) {
    if (!ctx || !key || !value || !len) return 1;
    const std::string k = key;
    const mprg_tuning &t = ctx->tune;
    std::string v;
    if (k == "accumulate") v = t.acc64 ? "f64" : "f32";
    else if (k == "apply") v = t.pipeOff ? "direct" : "pipe";
    else if (k == "pipe_split") v = t.pipeSplit ? "1" : "0";
    else if (k == "pipe_minb") v = std::to_string(t.pipeMinb);
    else if (k == "planes_shape") v = std::to_string(t.planesShape);
    else if (k == "cols_minb") v = std::to_string(t.colsMinb);
    else if (k == "wind") v = t.windChain ? "chain" : "composed";
    else if (k == "upload_threads") v = std::to_string(t.uploadThreads);
    else return 59;
    snprintf(value, len, "%s", v.c_str());
    return 0;
}

int mprg_set_async(mprg_ctx *ctx, int on) {
    MPRG_ENTER(ctx)
    if (ctx->async && !on) {  // leaving asynchronous mode: drain what is in flight
        MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
        MPRG_CUDA(cudaStreamSynchronize(ctx->h2d_stream));
        MPRG_CUDA(cudaStreamSynchronize(ctx->d2h_stream));
    }
    ctx->async = on != 0;
    MPRG_LEAVE(ctx)
}

int mprg_get_async(const mprg_ctx *ctx) { return ctx && ctx->async ? 1 : 0; }

int mprg_download(mprg_ctx *ctx, const void *dev, void *host, size_t bytes) {
    Trace tr("download", (int)(bytes >> 20));
    MPRG_ENTER(ctx)
    if (!dev || !host) fail(1, "mprg_download: null argument");
    // after everything queued so far on the context's stream; on the D2H stream so that it overlaps later kernels
    MPRG_CUDA(cudaEventRecord(ctx->evDl, ctx->stream));
    MPRG_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, ctx->evDl, 0));
    MPRG_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->d2h_stream));
    ctx->d2hBytes += bytes;
    if (!ctx->async) MPRG_CUDA(cudaStreamSynchronize(ctx->d2h_stream));
    MPRG_LEAVE(ctx)
}

int mprg_set_source_byte_order(mprg_ctx *ctx, int big_endian) {
    MPRG_ENTER(ctx)
    ctx->srcBigEndian = big_endian != 0;
    MPRG_LEAVE(ctx)
}

int mprg_bswap(mprg_ctx *ctx, void *dev, size_t count, int dtype) {
    MPRG_ENTER(ctx)
    if (!dev && count) fail(1, "mprg_bswap: null argument");
    bswap_device(ctx, dev, count, dtype == MPRG_F32 ? 4 : 8, ctx->stream);
    MPRG_LEAVE(ctx)
}

int mprg_post_affine(mprg_ctx *ctx, void *dev, size_t count, int dtype, double scale, double offset) {
    MPRG_ENTER(ctx)
    if (!dev && count) fail(1, "mprg_post_affine: null argument");
    post_affine_device(ctx, dev, count, dtype, scale, offset);
    MPRG_LEAVE(ctx)
}

int mprg_set_mesh(mprg_ctx *ctx, int32_t nCells, int32_t nVertices, int32_t maxEdges, const double *lonCell_rad,
                  const double *latCell_rad, const double *lonVertex_rad, const double *latVertex_rad,
                  const int32_t *verticesOnCell) {
    MPRG_ENTER(ctx)
    mesh_set(ctx, nCells, nVertices, maxEdges, lonCell_rad, latCell_rad, lonVertex_rad, latVertex_rad,
             verticesOnCell);
    MPRG_LEAVE(ctx)
}

int mprg_set_target(mprg_ctx *ctx, int stagger, int32_t ni, int32_t nj, const double *lon_deg,
                    const double *lat_deg) {
    MPRG_ENTER(ctx)
    target_set(ctx, stagger, ni, nj, lon_deg, lat_deg);
    MPRG_LEAVE(ctx)
}

int mprg_set_target_projected(mprg_ctx *ctx, int stagger, int32_t ni, int32_t nj, const mprg_projection *proj) {
    MPRG_ENTER(ctx)
    target_generate(ctx, stagger, ni, nj, proj);
    MPRG_LEAVE(ctx)
}

int mprg_get_target_lonlat(mprg_ctx *ctx, int stagger, double *lon_deg, double *lat_deg) {
    MPRG_ENTER(ctx)
    if (stagger < 0 || stagger > 3 || !ctx->target[stagger].set || !ctx->target[stagger].lon.p) fail(18, "mprg_get_target_lonlat: stagger %d not set", stagger);
    const Target &t = ctx->target[stagger];
    const size_t b = (size_t)t.ni * t.nj * sizeof(double);
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (lon_deg) MPRG_CUDA(cudaMemcpy(lon_deg, t.lon.p, b, cudaMemcpyDeviceToHost));
    if (lat_deg) MPRG_CUDA(cudaMemcpy(lat_deg, t.lat.p, b, cudaMemcpyDeviceToHost));
    MPRG_LEAVE(ctx)
}

int mprg_target_map_factor(mprg_ctx *ctx, int stagger, int proj_code, double truelat1, double truelat2, double *mapfac) {
    MPRG_ENTER(ctx)
    if (stagger < 0 || stagger > 3 || !mapfac) fail(1, "mprg_target_map_factor: bad argument");
    const Target &t = ctx->target[stagger];
    DevBuf<double> out((size_t)t.ni * t.nj);
    target_map_factor(ctx, stagger, proj_code, truelat1, truelat2, out.p);
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
    MPRG_CUDA(cudaMemcpy(mapfac, out.p, out.bytes(), cudaMemcpyDeviceToHost));
    MPRG_LEAVE(ctx)
}

int mprg_set_rotation_from_target(mprg_ctx *ctx, double *cosa, double *sina) {
    MPRG_ENTER(ctx)
    target_rotang(ctx);
    const Target &tg = ctx->target[MPRG_CENTER];
    const int64_t n = (int64_t)tg.ni * tg.nj;
    rotation_constants(ctx, n);
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (cosa) MPRG_CUDA(cudaMemcpy(cosa, ctx->cosa.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (sina) MPRG_CUDA(cudaMemcpy(sina, ctx->sina.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    ctx->haveRot = true;
    MPRG_LEAVE(ctx)
}

int mprg_set_grid_kind(mprg_ctx *ctx, int kind) {
    MPRG_ENTER(ctx)
    if (kind != MPRG_GRID_NOPERI && kind != MPRG_GRID_1PERI_MONOPOLE) fail(58, "mprg_set_grid_kind: bad kind %d", kind);
    if (kind != ctx->gridKind) {
        if (ctx->capturing) fail(45, "mprg_set_grid_kind: not while capturing a graph");
        MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
        for (auto it = ctx->routes.begin(); it != ctx->routes.end();) {  // grid-source routes depend on the topology
            if (std::get<1>(it->first) == MPRG_SRC_GRID_CENTER || std::get<1>(it->first) == MPRG_SRC_MESH_WIND) {
                it->second->memoised = false;
                if (it->second->refcount <= 0) delete it->second;
                else ctx->imported.push_back(it->second);
                it = ctx->routes.erase(it);
            } else {
                ++it;
            }
        }
        ctx->gridKind = kind;
        for (bool &d : ctx->windDeclined) d = false;
    }
    MPRG_LEAVE(ctx)
}

int mprg_get_slab(const mprg_ctx *ctx, int stagger, int32_t *j0, int32_t *j1) {
    if (!ctx || stagger < 0 || stagger > MPRG_CENTER_HALO || !ctx->target[stagger].set) return 1;
    if (j0) *j0 = ctx->target[stagger].j0;
    if (j1) *j1 = ctx->target[stagger].j1;
    return 0;
}

int mprg_store(mprg_ctx *ctx, int method, int src_loc, int dst_stagger, mprg_route **rh) {
    Trace tr("store", method, dst_stagger);
    MPRG_ENTER(ctx)
    if (!rh) fail(1, "mprg_store: null route pointer");
    *rh = nullptr;
    if (dst_stagger < 0 || dst_stagger > MPRG_CENTER_HALO) fail(51, "mprg_store: bad destination stagger %d", dst_stagger);
    if (dst_stagger == MPRG_CENTER_HALO && ctx->nranks == 1) dst_stagger = MPRG_CENTER;  // same rows: share the route
    if (!ctx->target[dst_stagger].set) fail(52, "mprg_store: target stagger %d not set", dst_stagger);
    if (src_loc != MPRG_SRC_GRID_CENTER && ctx->mesh.nCells == 0) fail(53, "mprg_store: mesh not set");
    auto key = std::make_tuple(method, src_loc, dst_stagger);
    auto it = ctx->routes.find(key);
    if (it != ctx->routes.end()) {
        it->second->refcount++;
        *rh = it->second;
        ctx->last_ms = 0.0;
        return 0;
    }
    // Weights depend on geometry only, so they are generated on their own stream: queued (asynchronous)
    // applies keep streaming while the host waits here.  The call returns with the route complete.
    struct StreamSwap {
        mprg_ctx *c; cudaStream_t saved;
        StreamSwap(mprg_ctx *c_) : c(c_), saved(c_->stream) { c->stream = c->store_stream; mprg::tl_stream = c->stream; }
        ~StreamSwap() { c->stream = saved; mprg::tl_stream = saved; }
    } swap(ctx);
    if (ctx->capturing) fail(46, "mprg_store: this route is not memoised yet; build it before mprg_capture_begin");
    std::unique_ptr<mprg_route> r(new mprg_route());  // after `swap`: on failure its buffers are freed on the store stream
    r->method = method; r->src_loc = src_loc; r->dst_stagger = dst_stagger;
    MPRG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    const bool known = (src_loc == MPRG_SRC_MESH_ELEMENT && (method == MPRG_NEAREST_STOD || method == MPRG_BILINEAR || method == MPRG_CONSERVE)) ||
                       (src_loc == MPRG_SRC_MESH_NODE && method == MPRG_BILINEAR) || (src_loc == MPRG_SRC_GRID_CENTER && method == MPRG_BILINEAR);
    if (!known) fail(54, "mprg_store: unsupported (method %d, src_loc %d)", method, src_loc);
    unsigned long long ckey[2] = {0, 0};
    const bool cached = !ctx->cacheDir.empty() && wcache_load(ctx, r.get(), ckey);   // cross-run weight cache (wcache.cu)
    if (!cached) {
        if (src_loc == MPRG_SRC_MESH_ELEMENT && method == MPRG_NEAREST_STOD) store_nearest(ctx, r.get());
        else if (src_loc == MPRG_SRC_MESH_ELEMENT && method == MPRG_BILINEAR) store_bilinear_element(ctx, r.get());
        else if (src_loc == MPRG_SRC_MESH_ELEMENT && method == MPRG_CONSERVE) store_conserve(ctx, r.get());
        else if (src_loc == MPRG_SRC_MESH_NODE && method == MPRG_BILINEAR) store_bilinear_node(ctx, r.get());
        else store_bilinear_grid(ctx, r.get());
    }
    r->dstNi = ctx->target[dst_stagger].ni;
    if (r->srcLevelSlowest) r->srcNi = ctx->target[MPRG_CENTER_HALO].ni;
    route_finish(ctx, r.get());
    if (!cached && !ctx->cacheDir.empty()) wcache_save(ctx, r.get(), ckey);
    MPRG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));  // every allocation and kernel of the route is complete
    float ms = 0.f;
    MPRG_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    r->refcount = 1;
    r->memoised = true;
    *rh = r.get();
    ctx->routes[key] = r.release();
    MPRG_LEAVE(ctx)
}

// drop the memoised composed wind routes (they embed the rotation angles)
static void drop_wind_routes(mprg_ctx *ctx) {
    for (bool &d : ctx->windDeclined) d = false;
    for (auto it = ctx->routes.begin(); it != ctx->routes.end();) {
        if (std::get<1>(it->first) == MPRG_SRC_MESH_WIND) {
            it->second->memoised = false;
            if (it->second->refcount <= 0) delete it->second;
            else ctx->imported.push_back(it->second);
            it = ctx->routes.erase(it);
        } else {
            ++it;
        }
    }
}

int mprg_store_wind(mprg_ctx *ctx, int dst_stagger, mprg_route **rh) {
    Trace tr("store_wind", dst_stagger);
    if (!ctx) return 1;
    if (!rh) { ctx->err = "mprg_store_wind: null route pointer"; return 1; }
    *rh = nullptr;
    if (dst_stagger != MPRG_EDGE1 && dst_stagger != MPRG_EDGE2) { ctx->err = "mprg_store_wind: destination must be MPRG_EDGE1 or MPRG_EDGE2"; return 51; }
    auto key = std::make_tuple((int)MPRG_BILINEAR, (int)MPRG_SRC_MESH_WIND, dst_stagger);
    auto it = ctx->routes.find(key);
    if (it != ctx->routes.end()) {
        it->second->refcount++;
        *rh = it->second;
        ctx->last_ms = 0.0;
        return 0;
    }
    if (ctx->tune.windChain || !ctx->haveRot || ctx->gridKind != MPRG_GRID_NOPERI || ctx->windDeclined[dst_stagger]) return 0;   // not composable: *rh stays NULL
    // the two matrices of the chain (memoised like any other route)
    mprg_route *bil = nullptr, *stag = nullptr;
    int rc = mprg_store(ctx, MPRG_BILINEAR, MPRG_SRC_MESH_ELEMENT, MPRG_CENTER_HALO, &bil);
    if (rc) return rc;
    rc = mprg_store(ctx, MPRG_BILINEAR, MPRG_SRC_GRID_CENTER, dst_stagger, &stag);
    if (rc) { mprg_release(ctx, bil); return rc; }
    int out = 0;
    {
        struct Guard { mprg_ctx *c; mprg_route *a, *b; ~Guard() { mprg_release(c, a); mprg_release(c, b); } } guard{ctx, bil, stag};
        MPRG_ENTER(ctx)
        if (ctx->capturing) fail(46, "mprg_store_wind: this route is not memoised yet; build it before mprg_capture_begin");
        struct StreamSwap {
            mprg_ctx *c; cudaStream_t saved;
            StreamSwap(mprg_ctx *c_) : c(c_), saved(c_->stream) { c->stream = c->store_stream; mprg::tl_stream = c->stream; }
            ~StreamSwap() { c->stream = saved; mprg::tl_stream = saved; }
        } swap(ctx);
        std::unique_ptr<mprg_route> r(new mprg_route());
        r->method = MPRG_BILINEAR; r->src_loc = MPRG_SRC_MESH_WIND; r->dst_stagger = dst_stagger;
        MPRG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
        if (!store_wind_composed(ctx, r.get(), stag, bil)) {
            MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
            ctx->windDeclined[dst_stagger] = true;
            return 0;   // not composable on this grid / mesh: the caller keeps the three-step chain
        }
        r->dstNi = ctx->target[dst_stagger].ni;
        route_finish(ctx, r.get());
        MPRG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
        MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        MPRG_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->last_ms = ms;
        if (r->tileEntriesMax <= 0 || r->tileEntriesMax > 512 || !r->rec32.p) {   // no tile schedule: chain
            ctx->windDeclined[dst_stagger] = true;
            return 0;
        }
        r->refcount = 1;
        r->memoised = true;
        *rh = r.get();
        ctx->routes[key] = r.release();
        MPRG_LEAVE(ctx)
    }
    return out;
}

int mprg_apply_wind(mprg_ctx *ctx, mprg_route *rh, const void *u_src, const void *v_src, int32_t nlev, int src_dtype,
                    void *dst, int dst_dtype, int into_full) {
    Trace tr("apply_wind", nlev, into_full);
    MPRG_ENTER(ctx)
    if (!rh) fail(1, "mprg_apply_wind: null route");
    if (!ctx->capturing) MPRG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    apply_wind_device(ctx, rh, u_src, v_src, nlev, src_dtype, dst, dst_dtype, into_full != 0);
    if (!ctx->capturing) MPRG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    MPRG_LEAVE(ctx)
}

int mprg_release(mprg_ctx *ctx, mprg_route *rh) {
    MPRG_ENTER(ctx)
    if (!rh) fail(1, "mprg_release: null route");
    rh->refcount--;
    if (rh->refcount <= 0 && !rh->memoised) {
        auto it = std::find(ctx->imported.begin(), ctx->imported.end(), rh);
        if (it != ctx->imported.end()) ctx->imported.erase(it);
        MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
        delete rh;
    }
    MPRG_LEAVE(ctx)
}

int mprg_clear_routes(mprg_ctx *ctx) {
    Trace tr("clear_routes");
    MPRG_ENTER(ctx)
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto &kv : ctx->routes) {
        kv.second->memoised = false;
        if (kv.second->refcount <= 0) delete kv.second;
        else ctx->imported.push_back(kv.second);  // still held by a caller: freed on its release
    }
    ctx->routes.clear();
    for (bool &d : ctx->windDeclined) d = false;
    MPRG_LEAVE(ctx)
}

int mprg_route_info(const mprg_route *rh, int64_t *nDst, int64_t *nnz, int64_t *nUnmapped, int64_t *nSrc) {
    if (!rh) return 1;
    if (nDst) *nDst = rh->nDst;
    if (nnz) *nnz = rh->nnz;
    if (nUnmapped) *nUnmapped = rh->nUnmapped;
    if (nSrc) *nSrc = rh->nSrc;
    return 0;
}

int mprg_route_export_csr(mprg_ctx *ctx, const mprg_route *rh, int32_t *rowptr, int32_t *col, double *w) {
    MPRG_ENTER(ctx)
    if (!rh) fail(1, "mprg_route_export_csr: null route");
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (rowptr) MPRG_CUDA(cudaMemcpy(rowptr, rh->rowptr.p, (rh->nDst + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (col && rh->nnz) MPRG_CUDA(cudaMemcpy(col, rh->col.p, rh->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (w && rh->nnz) MPRG_CUDA(cudaMemcpy(w, rh->w.p, rh->nnz * sizeof(double), cudaMemcpyDeviceToHost));
    MPRG_LEAVE(ctx)
}

int mprg_route_export_w2(mprg_ctx *ctx, const mprg_route *rh, double *w2) {
    MPRG_ENTER(ctx)
    if (!rh || !w2) fail(1, "mprg_route_export_w2: null argument");
    if (!rh->composite) fail(47, "mprg_route_export_w2: not a composed wind route");
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (rh->nnz) MPRG_CUDA(cudaMemcpy(w2, rh->w2.p, rh->nnz * sizeof(double), cudaMemcpyDeviceToHost));
    MPRG_LEAVE(ctx)
}

int mprg_route_import_csr(mprg_ctx *ctx, int64_t nSrc, int64_t nDst, const int32_t *rowptr, const int32_t *col,
                          const double *w, mprg_route **rh) {
    MPRG_ENTER(ctx)
    if (!rh || !rowptr || nDst < 0 || nSrc <= 0) fail(1, "mprg_route_import_csr: bad arguments");
    *rh = nullptr;
    if (rowptr[0] != 0) fail(61, "mprg_route_import_csr: rowptr[0] != 0");
    for (int64_t t = 0; t < nDst; ++t)
        if (rowptr[t + 1] < rowptr[t]) fail(62, "mprg_route_import_csr: rowptr not monotone at row %lld", (long long)t);
    int64_t nnz = rowptr[nDst];
    if (nnz > 0 && (!col || !w)) fail(1, "mprg_route_import_csr: null col/w");
    for (int64_t k = 0; k < nnz; ++k)
        if (col[k] < 0 || col[k] >= nSrc) fail(63, "mprg_route_import_csr: column %d out of range at %lld", col[k], (long long)k);
    std::unique_ptr<mprg_route> r(new mprg_route());
    r->method = -1; r->src_loc = MPRG_SRC_MESH_ELEMENT; r->dst_stagger = -1;
    r->nDst = nDst; r->nnz = nnz; r->nSrc = nSrc;
    r->rowptr.alloc(nDst + 1); r->col.alloc(nnz > 0 ? nnz : 1); r->w.alloc(nnz > 0 ? nnz : 1);
    MPRG_CUDA(cudaMemcpyAsync(r->rowptr.p, rowptr, (nDst + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    if (nnz) {
        MPRG_CUDA(cudaMemcpyAsync(r->col.p, col, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        MPRG_CUDA(cudaMemcpyAsync(r->w.p, w, nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    route_finish(ctx, r.get());
    r->refcount = 1;
    *rh = r.get();
    ctx->imported.push_back(r.release());
    MPRG_LEAVE(ctx)
}

// ---------------------------------------------------------------------------
// apply
// ---------------------------------------------------------------------------
// Host -> device copy of a source that is not page-locked (typically a variable inside a memory-mapped input
// file).  cudaMemcpyAsync from pageable memory is staged by the driver on ONE host thread (page faults of the
// mapping included): ~10 GB/s measured on the 3-km history file.  Here a few host threads copy chunks into a
// ring of pinned slots and the copy engine drains the slots in order on the H2D stream, so the page-cache reads
// run in parallel and overlap the DMA (4-MiB chunks, 32 slots: 4 in flight on the copy engine, the rest being filled).  Returns when the last chunk has been handed to the copy engine and its
// slot is free again (the device-side copy is then complete); kernels queued behind it stay asynchronous.
static void upload_unpinned(mprg_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes) {
    constexpr int R = mprg_ctx::kBounce;
    constexpr size_t C = mprg_ctx::kBounceBytes;
    constexpr int64_t kInflight = 4;  // chunks handed to the copy engine and not yet known complete (< R)
    if (!ctx->bounce) {
        MPRG_CUDA(cudaHostAlloc((void **)&ctx->bounce, (size_t)R * C, cudaHostAllocDefault));
        for (auto &e : ctx->evBounce) MPRG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const int64_t nChunks = (int64_t)((bytes + C - 1) / C);
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    // three quarters of the cores, shared between the ranks of the box
    unsigned want = std::max(2u, hw * 3 / 4 / (unsigned)std::max(1, ctx->nranks));
    if (ctx->tune.uploadThreads > 0) want = (unsigned)ctx->tune.uploadThreads;
    const unsigned nt = (unsigned)std::min<int64_t>(std::min<unsigned>(want, 16), nChunks);
    std::atomic<int64_t> next{0}, released{0};  // next chunk to claim; chunks whose slot is free again
    std::vector<std::atomic<int>> done(nChunks);
    for (auto &d : done) d.store(0, std::memory_order_relaxed);
    std::atomic<bool> abort{false};
    const unsigned char *src = (const unsigned char *)src_host;
    unsigned char *ring = ctx->bounce;
    std::vector<std::thread> workers;
    for (unsigned t = 0; t < nt; ++t)
        workers.emplace_back([&] {
            for (;;) {
                const int64_t i = next.fetch_add(1);
                if (i >= nChunks) return;
                while (i - released.load(std::memory_order_acquire) >= R) {  // slot still owned by chunk i - R
                    if (abort.load()) return;
                    std::this_thread::yield();
                }
                const size_t off = (size_t)i * C, len = std::min(C, bytes - off);
                std::memcpy(ring + (size_t)(i % R) * C, src + off, len);
                done[i].store(1, std::memory_order_release);
            }
        });
    cudaError_t bad = cudaSuccess;
    int64_t freed = 0;
    for (int64_t i = 0; i < nChunks && bad == cudaSuccess; ++i) {
        while (!done[i].load(std::memory_order_acquire)) std::this_thread::yield();
        const size_t off = (size_t)i * C, len = std::min(C, bytes - off);
        bad = cudaMemcpyAsync((unsigned char *)dst_dev + off, ring + (size_t)(i % R) * C, len, cudaMemcpyHostToDevice, ctx->h2d_stream);
        if (bad == cudaSuccess) bad = cudaEventRecord(ctx->evBounce[i % R], ctx->h2d_stream);
        // keep a few transfers in flight; the slots of the completed ones go back to the workers
        while (bad == cudaSuccess && freed <= i - kInflight) {
            bad = cudaEventSynchronize(ctx->evBounce[freed % R]);
            released.store(++freed, std::memory_order_release);
        }
    }
    if (bad != cudaSuccess) abort.store(true);
    released.store(nChunks + R, std::memory_order_release);  // unblock (after an error: let the workers run out)
    for (auto &w : workers) w.join();
    MPRG_CUDA(bad);
    // the last few slots: the next upload reuses the ring from slot 0
    for (; freed < nChunks; ++freed) MPRG_CUDA(cudaEventSynchronize(ctx->evBounce[freed % R]));
}

static bool is_pinned_host(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type != cudaMemoryTypeUnregistered;
}

static void apply_impl(mprg_ctx *ctx, mprg_route *rh, int32_t nfields, const void *const *src, const int32_t *nlev,
                       int src_dtype, int src_mem, void *const *dst, int dst_dtype, int dst_mem,
                       const int32_t *epi_op, const double *epi_arg, bool into_full = false) {
    if (!rh) fail(1, "mprg_apply: null route");
    if (nfields <= 0) return;
    if (!src || !dst || !nlev) fail(1, "mprg_apply: null argument");
    if (rh->nDst == 0) return;  // this rank owns no destination rows (nranks > nj): nothing to regrid
    const size_t isz = src_dtype == MPRG_F32 ? 4 : 8, osz = dst_dtype == MPRG_F32 ? 4 : 8;
    const int64_t nSrcPts = rh->srcLevelSlowest ? rh->srcPlane : rh->nSrc;
    if (ctx->capturing && (src_mem != MPRG_DEVICE || dst_mem != MPRG_DEVICE))
        fail(45, "mprg_apply: host buffers cannot be used while capturing a graph");
    if (!ctx->capturing) MPRG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    if (src_mem == MPRG_DEVICE && dst_mem == MPRG_DEVICE) {
        std::vector<ApplyField> fl(nfields);
        for (int f = 0; f < nfields; ++f)
            fl[f] = ApplyField{src[f], dst[f], nlev[f], epi_op ? epi_op[f] : 0, epi_arg ? epi_arg[f] : 0.0};
        apply_device(ctx, rh, fl.data(), nfields, src_dtype, dst_dtype, into_full);
    } else {
        // Pipelined staging: fields are cut into batches; a batch's H2D overlaps the kernels of the batch
        // before it and the D2H of the ones before that.  The slot ring persists across calls, so with
        // mprg_set_async(1) consecutive applies overlap the same way.
        const size_t budget = (size_t)768 << 20;
        // Source halo-sharding: only the id range [srcLo, srcHi) that this rank's weights reference is
        // uploaded (file order keeps it contiguous per field); kernels see the field through a virtual
        // base pointer `staged - srcLo * column`.  With a row-slab decomposition and a mesh numbered with
        // any spatial locality this is ~1/nranks of the field plus a halo; worst case it is the whole field.
        const bool range = src_mem == MPRG_HOST && !rh->srcLevelSlowest && rh->srcHi > rh->srcLo;
        const int64_t lo = range ? rh->srcLo : 0, hi = range ? rh->srcHi : nSrcPts;
        constexpr size_t kPad = 512;  // slack either side of a staged field: aligned-window reads may start / end up to 15 bytes outside
        auto in_bytes = [&](int k) { return (size_t)(hi - lo) * nlev[k] * isz; };
        auto in_slot = [&](int k) { return (in_bytes(k) + 2 * kPad + 255) & ~(size_t)255; };
        int f = 0;
        while (f < nfields) {
            int g = f;
            size_t inB = 0, outB = 0;
            while (g < nfields) {
                size_t a = in_slot(g), b = ((size_t)rh->nDst * nlev[g] * osz + 255) & ~(size_t)255;
                const bool pair_tail = epi_op && epi_op[g] == MPRG_EPI_ROT_V;  // never cut a wind pair in two
                if (g > f && !pair_tail && (inB + a > budget || outB + b > budget)) break;
                inB += a; outB += b; ++g;
            }
            const int slot = (int)(ctx->slotCursor++ % mprg_ctx::kSlots);
            // a slot that must grow grows every slot of the ring to the same size: the ring rotates across
            // calls, so otherwise the same growth (a device-wide sync plus a large allocation) would be paid
            // again on each of the next passes
            if (src_mem == MPRG_HOST && inB > ctx->stageIn[slot].n) {
                // size for the largest batch a pass can bring, not just this one: a wind pair is never cut in two, so
                // leave room for two of the largest field seen (growing 4 x 1.2 GB a second time cost 0.6 - 2.3 s of
                // physical allocation in the file driver, where every run starts with a fresh context)
                size_t big = 0;
                for (int k = 0; k < nfields; ++k) big = std::max(big, in_slot(k));
                const size_t want = std::max(inB, 2 * big);
                for (auto &b : ctx->stageIn) b.ensure_shared(want);
            }
            if (dst_mem == MPRG_HOST && outB > ctx->stageOut[slot].n)
                for (auto &b : ctx->stageOut) b.ensure_shared(outB);
            // slot reuse: the kernel that last read stageIn[slot] / the D2H that last read stageOut[slot]
            if (ctx->slotUsed[slot]) {
                MPRG_CUDA(cudaStreamWaitEvent(ctx->h2d_stream, ctx->evK[slot], 0));
                MPRG_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->evOut[slot], 0));
            }
            std::vector<ApplyField> fl;
            std::vector<std::pair<unsigned char *, size_t>> staged;  // (device address, elements) of each uploaded field
            size_t io = 0, oo = 0;
            for (int k = f; k < g; ++k) {
                size_t b = (size_t)rh->nDst * nlev[k] * osz;
                const void *s = src[k];
                void *d = dst[k];
                if (src_mem == MPRG_HOST) {
                    const size_t col = (size_t)nlev[k] * isz, skip = (size_t)lo * col;
                    // keep the 16-byte phase of the original layout so aligned fields stay aligned
                    unsigned char *data = ctx->stageIn[slot].p + io + kPad + (skip & 15);
                    const unsigned char *from = (const unsigned char *)src[k] + skip;
                    // only the id ranges the weights reference cross PCIe (one range when the mesh is numbered along
                    // the slab; a short list otherwise); the staged field keeps the layout of [lo, hi)
                    std::pair<int64_t, int64_t> whole(lo, hi);
                    const std::pair<int64_t, int64_t> *rb = &whole, *re = rb + 1;
                    if (range && !rh->srcRanges.empty()) { rb = rh->srcRanges.data(); re = rb + rh->srcRanges.size(); }
                    const bool pinned = is_pinned_host(from);
                    for (const std::pair<int64_t, int64_t> *rg = rb; rg != re; ++rg) {
                        const size_t off = (size_t)(rg->first - lo) * col, nb = (size_t)(rg->second - rg->first) * col;
                        if (nb == 0) continue;
                        if (nb >= ((size_t)16 << 20) && !pinned)
                            upload_unpinned(ctx, data + off, from + off, nb);
                        else
                            MPRG_CUDA(cudaMemcpyAsync(data + off, from + off, nb, cudaMemcpyHostToDevice, ctx->h2d_stream));
                        ctx->h2dBytes += nb;
                        staged.emplace_back(data + off, nb / isz);
                    }
                    s = data - skip;
                }
                if (dst_mem == MPRG_HOST) d = ctx->stageOut[slot].p + oo;
                fl.push_back(ApplyField{s, d, nlev[k], epi_op ? epi_op[k] : 0, epi_arg ? epi_arg[k] : 0.0});
                io += in_slot(k);
                oo += (b + 255) & ~(size_t)255;
            }
            if (src_mem == MPRG_HOST) {
                MPRG_CUDA(cudaEventRecord(ctx->evIn[slot], ctx->h2d_stream));
                MPRG_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->evIn[slot], 0));
                if (ctx->srcBigEndian)  // file bytes crossed PCIe untouched: swap them where bandwidth is cheap
                    for (auto &sf : staged) bswap_device(ctx, sf.first, sf.second, isz, ctx->stream);
            }
            apply_device(ctx, rh, fl.data(), (int)fl.size(), src_dtype, dst_dtype, into_full);
            MPRG_CUDA(cudaEventRecord(ctx->evK[slot], ctx->stream));
            if (dst_mem == MPRG_HOST) {
                MPRG_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, ctx->evK[slot], 0));
                oo = 0;
                for (int k = f; k < g; ++k) {
                    size_t b = (size_t)rh->nDst * nlev[k] * osz;
                    MPRG_CUDA(cudaMemcpyAsync(dst[k], ctx->stageOut[slot].p + oo, b, cudaMemcpyDeviceToHost, ctx->d2h_stream));
                    ctx->d2hBytes += b;
                    oo += (b + 255) & ~(size_t)255;
                }
            }
            MPRG_CUDA(cudaEventRecord(ctx->evOut[slot], ctx->d2h_stream));
            ctx->slotUsed[slot] = true;
            f = g;
        }
        // host-buffer applies complete before returning (reference semantics: Regrid is blocking) unless
        // the caller asked for asynchronous applies and synchronises itself (mprg_synchronize)
        if (!ctx->async) {
            MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
            if (dst_mem == MPRG_HOST) MPRG_CUDA(cudaStreamSynchronize(ctx->d2h_stream));
        }
    }
    if (!ctx->capturing) MPRG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
}

int mprg_apply_ex(mprg_ctx *ctx, mprg_route *rh, int32_t nfields, const void *const *src, const int32_t *nlev,
                  int src_dtype, int src_mem, void *const *dst, int dst_dtype, int dst_mem, const int32_t *epi_op,
                  const double *epi_arg) {
    Trace tr("apply", nfields, src_mem * 2 + dst_mem);
    MPRG_ENTER(ctx)
    apply_impl(ctx, rh, nfields, src, nlev, src_dtype, src_mem, dst, dst_dtype, dst_mem, epi_op, epi_arg);
    MPRG_LEAVE(ctx)
}

int mprg_apply_into(mprg_ctx *ctx, mprg_route *rh, int32_t nfields, const void *const *src, const int32_t *nlev,
                    int src_dtype, int src_mem, void *const *dst_full, int dst_dtype, const int32_t *epi_op,
                    const double *epi_arg) {
    Trace tr("apply_into", nfields, src_mem);
    MPRG_ENTER(ctx)
    apply_impl(ctx, rh, nfields, src, nlev, src_dtype, src_mem, dst_full, dst_dtype, MPRG_DEVICE, epi_op, epi_arg, true);
    MPRG_LEAVE(ctx)
}

// ---- CUDA IPC: the writing rank exports its full-grid output buffers, the others map them
int mprg_ipc_export(mprg_ctx *ctx, const void *dev_ptr, void *handle64, size_t *offset) {
    MPRG_ENTER(ctx)
    if (!dev_ptr || !handle64 || !offset) fail(1, "mprg_ipc_export: null argument");
    // the handle names the whole allocation: find its base (driver entry point, no libcuda link)
    typedef int (*fn_range)(unsigned long long *, size_t *, unsigned long long);
    static fn_range range = nullptr;
    if (!range) {
        cudaDriverEntryPointQueryResult q;
        void *f = nullptr;
        MPRG_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q));
        if (!f || q != cudaDriverEntryPointSuccess) fail(87, "mprg_ipc_export: cuMemGetAddressRange unavailable");
        range = (fn_range)f;
    }
    unsigned long long base = 0;
    size_t size = 0;
    if (range(&base, &size, (unsigned long long)(uintptr_t)dev_ptr) != 0) fail(88, "mprg_ipc_export: not a device allocation");
    cudaIpcMemHandle_t h;
    MPRG_CUDA(cudaIpcGetMemHandle(&h, (void *)(uintptr_t)base));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    *offset = (size_t)((uintptr_t)dev_ptr - (uintptr_t)base);
    MPRG_LEAVE(ctx)
}

int mprg_ipc_open(mprg_ctx *ctx, const void *handle64, size_t offset, void **peer_ptr) {
    MPRG_ENTER(ctx)
    if (!handle64 || !peer_ptr) fail(1, "mprg_ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    // one mapping per exported allocation: several buffers usually share an allocation
    std::string key((const char *)handle64, 64);
    auto it = ctx->ipcOpen.find(key);
    void *base = nullptr;
    if (it != ctx->ipcOpen.end()) {
        base = it->second;
    } else {
        MPRG_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->ipcOpen[key] = base;
    }
    *peer_ptr = (unsigned char *)base + offset;
    MPRG_LEAVE(ctx)
}

int mprg_put_slab(mprg_ctx *ctx, int stagger, int32_t nlev, int dtype, const void *slab_dev, void *full_dev) {
    MPRG_ENTER(ctx)
    if (stagger < 0 || stagger > MPRG_CORNER || !ctx->target[stagger].set) fail(83, "mprg_put_slab: stagger %d not set", stagger);
    const Target &tg = ctx->target[stagger];
    const size_t esz = dtype == MPRG_F32 ? 4 : 8;
    const int64_t nFull = (int64_t)tg.ni * tg.nj, nMine = tg.nSlab();
    if (nMine > 0 && nlev > 0) {
        if (!slab_dev || !full_dev) fail(1, "mprg_put_slab: null argument");
        MPRG_CUDA(cudaMemcpy2DAsync((unsigned char *)full_dev + (size_t)tg.slabOffset() * esz, (size_t)nFull * esz, slab_dev,
                                    (size_t)nMine * esz, (size_t)nMine * esz, (size_t)nlev, cudaMemcpyDeviceToDevice,
                                    ctx->stream));
    }
    MPRG_LEAVE(ctx)
}

int mprg_ipc_close_all(mprg_ctx *ctx) {
    MPRG_ENTER(ctx)
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto &kv : ctx->ipcOpen) cudaIpcCloseMemHandle(kv.second);
    ctx->ipcOpen.clear();
    MPRG_LEAVE(ctx)
}

int mprg_apply(mprg_ctx *ctx, mprg_route *rh, int32_t nfields, const void *const *src, const int32_t *nlev,
               int src_dtype, int src_mem, void *const *dst, int dst_dtype, int dst_mem) {
    return mprg_apply_ex(ctx, rh, nfields, src, nlev, src_dtype, src_mem, dst, dst_dtype, dst_mem, nullptr, nullptr);
}

// ---------------------------------------------------------------------------
// wind rotation
// ---------------------------------------------------------------------------
int mprg_set_rotation(mprg_ctx *ctx, const double *cosa, const double *sina) {
    MPRG_ENTER(ctx)
    const Target &tg = ctx->target[MPRG_CENTER];
    if (!tg.set) fail(42, "mprg_set_rotation: CENTER target not set");
    if (!cosa || !sina) fail(1, "mprg_set_rotation: null argument");
    const int64_t n = (int64_t)tg.ni * tg.nj;  // full grid: both CENTER and CENTER_HALO rows index into it
    ctx->cosa.alloc(n);
    ctx->sina.alloc(n);
    MPRG_CUDA(cudaMemcpyAsync(ctx->cosa.p, cosa, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    MPRG_CUDA(cudaMemcpyAsync(ctx->sina.p, sina, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    rotation_constants(ctx, n);
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->haveRot = true;
    drop_wind_routes(ctx);   // composed wind routes embed the previous angles
    MPRG_LEAVE(ctx)
}

int mprg_rotate_winds_on(mprg_ctx *ctx, int stagger, void *u, void *v, int32_t nlev, int dtype, int mem) {
    MPRG_ENTER(ctx)
    if (!u || !v) fail(1, "mprg_rotate_winds: null argument");
    if (stagger != MPRG_CENTER && stagger != MPRG_CENTER_HALO) fail(43, "mprg_rotate_winds: winds live on CENTER rows");
    const Target &tg = ctx->target[stagger];
    if (!tg.set) fail(42, "mprg_rotate_winds: CENTER target not set");
    const size_t bytes = (size_t)tg.nSlab() * nlev * (dtype == MPRG_F32 ? 4 : 8);
    if (mem == MPRG_DEVICE) {
        rotate_device(ctx, stagger, u, v, nlev, dtype);
    } else {
        ctx->stageIn[0].ensure_shared(2 * bytes);
        unsigned char *du = ctx->stageIn[0].p, *dv = du + bytes;
        MPRG_CUDA(cudaMemcpyAsync(du, u, bytes, cudaMemcpyHostToDevice, ctx->stream));
        MPRG_CUDA(cudaMemcpyAsync(dv, v, bytes, cudaMemcpyHostToDevice, ctx->stream));
        rotate_device(ctx, stagger, du, dv, nlev, dtype);
        MPRG_CUDA(cudaMemcpyAsync(u, du, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        MPRG_CUDA(cudaMemcpyAsync(v, dv, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    MPRG_LEAVE(ctx)
}

int mprg_rotate_winds(mprg_ctx *ctx, void *u, void *v, int32_t nlev, int dtype, int mem) {
    return mprg_rotate_winds_on(ctx, MPRG_CENTER, u, v, nlev, dtype, mem);
}

// ---------------------------------------------------------------------------
// WRF-compatibility post-ops
// ---------------------------------------------------------------------------
static const Target &post_target(mprg_ctx *ctx, int stagger) {
    if (stagger < 0 || stagger > MPRG_CENTER_HALO || !ctx->target[stagger].set) fail(44, "mprg_post: stagger %d not set", stagger);
    return ctx->target[stagger];
}

int mprg_post_midlevels(mprg_ctx *ctx, int stagger, int32_t nlev, int dtype, int mem, const void *x, void *mid) {
    MPRG_ENTER(ctx)
    const Target &tg = post_target(ctx, stagger);
    const int64_t n = tg.nSlab();
    if (n == 0 || nlev < 2) return 0;
    if (!x || !mid) fail(1, "mprg_post_midlevels: null argument");
    const size_t esz = dtype == MPRG_F32 ? 4 : 8, inB = (size_t)n * nlev * esz, outB = (size_t)n * (nlev - 1) * esz;
    if (mem == MPRG_DEVICE) {
        post_midlevels_device(ctx, n, nlev, dtype, x, mid);
    } else {
        MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
        DevBuf<unsigned char> din(inB), dout(outB);
        MPRG_CUDA(cudaMemcpyAsync(din.p, x, inB, cudaMemcpyHostToDevice, ctx->stream));
        post_midlevels_device(ctx, n, nlev, dtype, din.p, dout.p);
        MPRG_CUDA(cudaMemcpyAsync(mid, dout.p, outB, cudaMemcpyDeviceToHost, ctx->stream));
        MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    MPRG_LEAVE(ctx)
}

int mprg_post_ptop(mprg_ctx *ctx, int stagger, int32_t nlev, int dtype, int mem, const void *x, double *maxval,
                   double *mincand) {
    MPRG_ENTER(ctx)
    const Target &tg = post_target(ctx, stagger);
    const int64_t n = tg.nSlab();
    if (!maxval || !mincand) fail(1, "mprg_post_ptop: null argument");
    if (n > 0 && nlev > 0 && !x) fail(1, "mprg_post_ptop: null field");
    DevBuf<double> res(2);
    if (mem == MPRG_DEVICE || n == 0) {
        post_ptop_device(ctx, n, nlev, dtype, x, res.p);
    } else {
        const size_t inB = (size_t)n * nlev * (dtype == MPRG_F32 ? 4 : 8);
        DevBuf<unsigned char> din(inB);
        MPRG_CUDA(cudaMemcpyAsync(din.p, x, inB, cudaMemcpyHostToDevice, ctx->stream));
        post_ptop_device(ctx, n, nlev, dtype, din.p, res.p);
    }
    double h[2];
    MPRG_CUDA(cudaMemcpy(h, res.p, sizeof h, cudaMemcpyDeviceToHost));
    *maxval = h[0];
    *mincand = h[1] >= 1.0e300 ? (double)INFINITY : h[1];
    MPRG_LEAVE(ctx)
}

// ---------------------------------------------------------------------------
// gather (gather.cu)
// ---------------------------------------------------------------------------
int mprg_comm_id(mprg_ctx *ctx, void *id128) {
    MPRG_ENTER(ctx)
    if (!id128) fail(1, "mprg_comm_id: null buffer");
    comm_id(ctx, id128);
    MPRG_LEAVE(ctx)
}

int mprg_comm_init(mprg_ctx *ctx, const void *id128) {
    MPRG_ENTER(ctx)
    if (ctx->nranks > 1 && !id128) fail(1, "mprg_comm_init: null id");
    comm_init(ctx, id128);
    MPRG_LEAVE(ctx)
}

int mprg_gather(mprg_ctx *ctx, int stagger, int32_t nlev, int dtype, const void *slab_dev, int root, void *full_dev) {
    MPRG_ENTER(ctx)
    gather_slabs(ctx, 1, &stagger, &nlev, dtype, &slab_dev, root, &full_dev);
    MPRG_LEAVE(ctx)
}

int mprg_gather_v(mprg_ctx *ctx, int32_t nfields, const int *stagger, const int32_t *nlev, int dtype,
                  const void *const *slab_dev, int root, void *const *full_dev) {
    MPRG_ENTER(ctx)
    gather_slabs(ctx, nfields, stagger, nlev, dtype, slab_dev, root, full_dev);
    MPRG_LEAVE(ctx)
}

int mprg_profile_enable(mprg_ctx *ctx, int on) {
    MPRG_ENTER(ctx)
    ctx->profile = on != 0;
    MPRG_LEAVE(ctx)
}

int mprg_profile_reset(mprg_ctx *ctx) {
    MPRG_ENTER(ctx)
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto &r : ctx->prof) { ctx->evPool.push_back(r.a); ctx->evPool.push_back(r.b); }
    ctx->prof.clear();
    MPRG_LEAVE(ctx)
}

int mprg_profile_read(mprg_ctx *ctx, int32_t max, int32_t *kind, double *ms, double *alg_bytes, double *units) {
    if (!ctx) return -1;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    int32_t n = (int32_t)ctx->prof.size();
    for (int32_t i = 0; i < n && i < max; ++i) {
        float t = 0.f;
        cudaEventElapsedTime(&t, ctx->prof[i].a, ctx->prof[i].b);
        if (kind) kind[i] = ctx->prof[i].kind;
        if (ms) ms[i] = t;
        if (alg_bytes) alg_bytes[i] = ctx->prof[i].algBytes;
        if (units) units[i] = ctx->prof[i].units;
    }
    return n;
}

int64_t mprg_route_src_referenced(const mprg_route *rh) { return rh ? rh->nSrcRef : 0; }

int mprg_route_schedule_info(const mprg_route *rh, int64_t *tiles, int64_t *columns, int64_t *runs) {
    if (!rh) return 1;
    if (tiles) *tiles = rh->schedTiles;
    if (columns) *columns = rh->schedCols;
    if (runs) *runs = rh->schedRuns;
    return 0;
}

int64_t mprg_kernel_launches(const mprg_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------
// CUDA graphs
// ---------------------------------------------------------------------------
int mprg_capture_begin(mprg_ctx *ctx) {
    MPRG_ENTER(ctx)
    if (ctx->capturing) fail(47, "mprg_capture_begin: already capturing");
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
    MPRG_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    ctx->capturing = true;
    ctx->captureLaunches0 = ctx->launches;
    MPRG_LEAVE(ctx)
}

int mprg_capture_end(mprg_ctx *ctx, mprg_graph **graph) {
    if (!ctx) return 1;
    cudaSetDevice(ctx->device);
    cudaGraph_t g = nullptr;
    const bool was = ctx->capturing;
    ctx->capturing = false;
    cudaError_t e = was ? cudaStreamEndCapture(ctx->stream, &g) : cudaErrorInvalidValue;
    if (graph) *graph = nullptr;
    if (e != cudaSuccess || !g || !graph) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        ctx->err = std::string("mprg_capture_end: ") + (was ? cudaGetErrorString(e) : "no capture in progress");
        return 48;
    }
    cudaGraphExec_t ex = nullptr;
    e = cudaGraphInstantiate(&ex, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) {
        ctx->err = std::string("mprg_capture_end: cudaGraphInstantiate: ") + cudaGetErrorString(e);
        return 49;
    }
    mprg_graph *out = new mprg_graph();
    out->exec = ex;
    out->launches = ctx->launches - ctx->captureLaunches0;
    ctx->launches = ctx->captureLaunches0;  // recorded, not run
    *graph = out;
    return 0;
}

int mprg_graph_launch(mprg_ctx *ctx, mprg_graph *graph) {
    MPRG_ENTER(ctx)
    if (!graph || !graph->exec) fail(1, "mprg_graph_launch: null graph");
    if (ctx->capturing) fail(47, "mprg_graph_launch: a capture is in progress");
    MPRG_CUDA(cudaGraphLaunch(graph->exec, ctx->stream));
    ctx->launches += graph->launches;
    MPRG_LEAVE(ctx)
}

int mprg_graph_release(mprg_ctx *ctx, mprg_graph *graph) {
    MPRG_ENTER(ctx)
    if (graph) {
        MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
        if (graph->exec) cudaGraphExecDestroy(graph->exec);
        delete graph;
    }
    MPRG_LEAVE(ctx)
}

int mprg_io_bytes(const mprg_ctx *ctx, uint64_t *h2d, uint64_t *d2h) {
    if (!ctx) return 1;
    if (h2d) *h2d = ctx->h2dBytes;
    if (d2h) *d2h = ctx->d2hBytes;
    return 0;
}

double mprg_last_ms(const mprg_ctx *ctx) {
    if (!ctx) return 0.0;
    if (cudaEventQuery(ctx->ev1) == cudaSuccess && cudaEventQuery(ctx->ev0) == cudaSuccess) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) return ms;
    }
    return ctx->last_ms;
}

}  // extern "C"

// ---- not-yet-implemented stores (filled in by conserve.cu / stagger.cu) -------
namespace mprg {
#ifndef MPRG_HAVE_CONSERVE
void store_conserve(mprg_ctx *, mprg_route *) { fail(91, "CONSERVE store not built into this library"); }
#endif
#ifndef MPRG_HAVE_STAGGER
void store_bilinear_grid(mprg_ctx *, mprg_route *) { fail(92, "GRID_CENTER bilinear store not built into this library"); }
#endif
#ifndef MPRG_HAVE_NODE
void store_bilinear_node(mprg_ctx *, mprg_route *) { fail(93, "MESH_NODE bilinear store not built into this library"); }
#endif
}  // namespace mprg

// test / tuning hook (not part of the reference-facing surface): per-tile maxima that size the
// pipelined apply kernel's staging
extern "C" int mprg_debug_route_tiles(const mprg_route *rh, int32_t *entriesMax, int32_t *uniqMax) {
    if (!rh) return 1;
    if (entriesMax) *entriesMax = rh->tileEntriesMax;
    if (uniqMax) *uniqMax = rh->tileUniqMax;
    return 0;
}
