// common.cuh -- context, device buffers and error plumbing of libmpassit_rg.so
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/mpassit_rg.h"

namespace mprg {

constexpr double kTol = 1e-10;  // ESMF parametric point-in-element tolerance
constexpr int kNumSM = 148;     // B200

struct Error {
    int rc;
    std::string msg;
};

[[noreturn]] inline void fail(int rc, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Error{rc, buf};
}

#define MPRG_CUDA(expr)                                                                      \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess)                                                              \
            ::mprg::fail(700 + (int)e__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                         __FILE__, __LINE__);                                                \
    } while (0)

// Stream on which this thread's DevBuf allocations and frees are ordered: every C-ABI entry sets
// it to the context's stream (MPRG_ENTER in capi.cu).
inline thread_local cudaStream_t tl_stream = nullptr;

// RAII device allocation from the device's stream-ordered memory pool (cudaMallocAsync, release
// threshold = keep everything: mprg_init).  Weight generation allocates ~15 temporaries per route;
// with plain cudaMalloc/cudaFree each of those is a driver call behind a host-wide lock shared with
// every other tenant of the machine (measured: 30 ms .. 1.3 s per store on a shared B200 host for
// 10-20 ms of kernels).  Pool allocations are valid for work ordered after them on tl_stream; buffers
// that other streams touch are allocated with ensure_shared().
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) MPRG_CUDA(cudaMallocAsync((void **)&p, count * sizeof(T), tl_stream));
    }
    void ensure(size_t count) { if (count > n) alloc(count); }
    // for buffers used on more than one stream (staging): the old block may still be in flight on
    // another stream and the new one must be visible to all of them
    void ensure_shared(size_t count) {
        if (count <= n) return;
        MPRG_CUDA(cudaDeviceSynchronize());
        alloc(count);
        MPRG_CUDA(cudaStreamSynchronize(tl_stream));
    }
    void release() {
        if (p) cudaFreeAsync(p, tl_stream);
        p = nullptr;
        n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
};

struct PinnedBuf {
    void *p = nullptr;
    size_t n = 0;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
    void ensure(size_t bytes) {
        if (bytes <= n) return;
        if (p) cudaFreeHost(p);
        p = nullptr;
        n = 0;
        MPRG_CUDA(cudaMallocHost(&p, bytes));
        n = bytes;
    }
};

// Implicit bounding-volume hierarchy over Morton-sorted primitives (bvh.cu)
struct Bvh {
    int32_t nPrim = 0;
    int32_t nLeafNodes = 0;    // power of two
    DevBuf<float4> nodes;      // [2*nLeafNodes-1][2]: (lox,loy,loz,hix) (hiy,hiz,-,-)
    DevBuf<int32_t> primId;    // [nPrim] Morton order -> original id
    DevBuf<float4> primBox;    // optional [nPrim][2], Morton order: per-primitive boxes (same packing as nodes)
};

struct Mesh {
    int32_t nCells = 0, nVertices = 0, maxEdges = 0;
    DevBuf<double> cellXyz;    // [nCells][3]   unit sphere
    DevBuf<double> vertXyz;    // [nVertices][3]
    DevBuf<int32_t> voc;       // [nCells][maxEdges] 1-based, 0 = pad (as given)
    DevBuf<int32_t> tri;       // [nVertices][3] dual triangles, ascending cell ids, -1 = none
    DevBuf<double> cellSorted; // [nCells][3] in cellBvh order (leaf-contiguous loads)
    Bvh cellBvh;               // points: cell centres              (nearest)
    Bvh triBvh;                // boxes: dual triangles              (bilinear, element)
    Bvh polyBvh;               // boxes: Voronoi polygons            (conserve; bilinear, node)
    bool haveCellBvh = false, haveTriBvh = false, havePolyBvh = false;
    double maxTriEdge2 = 0.0;
};

struct Target {
    int32_t ni = 0, nj = 0;    // full grid
    int32_t j0 = 0, j1 = 0;    // slab rows owned by this rank
    DevBuf<double> xyz;        // full grid [nj][ni][3]
    DevBuf<double> lon, lat;   // full grid [nj][ni] degrees (as given, or generated by mprg_set_target_projected)
    const double *xyzRef = nullptr;  // MPRG_CENTER_HALO shares CENTER's coordinates
    const double *x() const { return xyzRef ? xyzRef : xyz.p; }
    bool set = false;
    int64_t nSlab() const { return (int64_t)(j1 - j0) * ni; }
    int64_t slabOffset() const { return (int64_t)j0 * ni; }
};

}  // namespace mprg

// Route handle: CSR weights of this rank's destination slab.
namespace mprg { constexpr int kLongRow = 4; constexpr int kCmpRow = 12; }

struct mprg_route {
    int method = 0, src_loc = 0, dst_stagger = 0;
    int64_t nDst = 0, nnz = 0, nUnmapped = 0, nSrc = 0;
    int64_t nSrcRef = 0;       // distinct source entities referenced by the weights
    int64_t srcLo = 0, srcHi = 0;  // [srcLo, srcHi): smallest id range containing all of them
    // the referenced ids as a short list of ranges inside [srcLo, srcHi) (256-cell granularity, small gaps merged):
    // what a host-buffer apply uploads.  One range on a mesh numbered along the slab; a few dozen on a Z-order or
    // generator-ordered mesh, where [srcLo, srcHi) alone would be most of the field on every rank.
    std::vector<std::pair<int64_t, int64_t>> srcRanges;
    int32_t tileEntriesMax = 0, tileUniqMax = 0;  // per 32-target tile: CSR entries / distinct columns
    int32_t tileRunsMax = 0;                      // per tile: runs of consecutive column ids
    int32_t dstNi = 0;         // destination row length (tiles of the apply kernel never straddle rows)
    // tile schedule for the pipelined apply kernel (apply_pipe.cuh): one record per 32-target tile, for fp32 weights
    // (built with the route) and for fp64 weights (built on first use)
    mprg::DevBuf<unsigned char> rec32, rec64;
    int32_t tileRowMax = 0;                       // longest row of any tile (> 3: the records keep the CSR slices)
    int64_t schedTiles = 0, schedCols = 0, schedRuns = 0;  // tiles, distinct columns and id-runs summed over tiles
    int32_t maxRow = 0;        // longest row
    bool uniform = false;      // every mapped row has exactly `maxRow` entries, stored ELL-like
    mprg::DevBuf<int32_t> rowptr;  // [nDst+1]
    mprg::DevBuf<int32_t> col;     // [nnz]
    mprg::DevBuf<double> w;        // [nnz]
    mprg::DevBuf<float> w32;       // [nnz] fp32 copy for the all-fp32 path
    int refcount = 0;
    bool memoised = false;
    // source is a structured grid (MPRG_SRC_GRID_CENTER): level-slowest source layout
    bool srcLevelSlowest = false;
    int64_t srcPlane = 0;      // points per source level plane
    int32_t srcNi = 0;         // row length of the source grid
    int32_t planeWin = 0;      // source rows per 256-target tile (apply_planes.cuh); 0: register-gather kernel only
    mprg::DevBuf<int32_t> planeTiles;  // [tiles][8] per-tile headers of the staged grid-source kernel (k_plane_stats)
    // rows of a grid-source route longer than kLongRow entries (the pole rows of a periodic grid: ni + 2 entries):
    // applied by one warp per (row, level) instead of one thread per row
    mprg::DevBuf<int32_t> longRows;
    int64_t nLong = 0;
    // composed wind route (mprg_store_wind, compose.cu): every entry carries TWO weights -- w multiplies the zonal
    // source column, w2 the meridional one -- and rows have at most kCmpRow entries
    bool composite = false;
    mprg::DevBuf<double> w2;
    mprg::DevBuf<float> w2_32;
};

struct mprg_graph {
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;   // engine kernels inside
};

// Tuning knobs: read once from the environment at mprg_init, changed at run time with mprg_set_option
// (never getenv on the launch path).
struct mprg_tuning {
    bool acc64 = false;        // MPASSIT_GPU_ACC=f64 | option "accumulate": fp32 fields accumulate (and rotate) in fp64, the reference's R8
    bool pipeOff = false;      // MPASSIT_GPU_APPLY=direct | option "apply": register-gather kernels only
    int pipeMinb = 0;          // MPASSIT_GPU_PIPE_MINB | option "pipe_minb": 4 / 5 resident CTAs per SM (0 = by shared memory)
    bool pipeSplit = false;    // MPASSIT_GPU_PIPE_SPLIT=0|1 | option "pipe_split": plain aligned units and the rest in separate launches
    int planesShape = 0;       // option "planes_shape": levels per stage * 10 + stages of the staged grid-source kernel (0 = default)
    int colsMinb = 3;          // MPASSIT_GPU_MINB | option "cols_minb": register cap of the fallback kernel
    bool windChain = true;     // MPASSIT_GPU_WIND=composed|chain | option "wind": chain = mprg_store_wind declines, hosts keep the three-step wind chain
    int uploadThreads = 0;     // MPASSIT_UPLOAD_THREADS | option "upload_threads" (0 = 3/4 of the cores / ranks)
};

struct mprg_ctx {
    int device = 0, rank = 0, nranks = 1;
    int numSM = mprg::kNumSM;             // multiprocessors of the device (persistent grids)
    mprg_tuning tune;
    int gridKind = MPRG_GRID_NOPERI;      // mprg_set_grid_kind: topology of the target grid as a regrid source
    std::string cacheDir;                 // mprg_set_weight_cache: directory of the cross-run weight cache ("" = off)
    int64_t cacheHits = 0, cacheStores = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaStream_t store_stream = nullptr;  // weight generation runs here, beside the copy-bound apply pipeline
    bool async = false;                   // host-buffer applies return once enqueued (mprg_set_async)
    bool capturing = false;               // between mprg_capture_begin / _end
    bool srcBigEndian = false;            // host sources hold file-order (big-endian) words (mprg_set_source_byte_order)
    int64_t captureLaunches0 = 0;
    std::string err;
    mprg::Mesh mesh;
    mprg::Target target[5];           // + MPRG_CENTER_HALO (derived from CENTER)
    std::map<std::tuple<int, int, int>, mprg_route *> routes;
    std::vector<mprg_route *> imported;
    mprg::DevBuf<double> cosa, sina;  // CENTER, full grid
    mprg::DevBuf<double> rotc;        // [n][4] per-point rotation constants (sina, tana, 1/cosa, 1/(cosa + sina tana))
    mprg::DevBuf<float> rotc32;       // the same rounded to fp32 (all-fp32 applies rotate in fp32)
    bool haveRot = false;
    bool windDeclined[4] = {};        // mprg_store_wind found this stagger not composable (reset with the routes)
    int64_t launches = 0;
    double last_ms = 0.0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // staging for host-buffer applies: a ring of slots that persists across calls, so consecutive
    // (asynchronous) applies keep H2D, kernels and D2H of different batches in flight together
    static constexpr int kSlots = 4;
    mprg::DevBuf<unsigned char> stageIn[kSlots], stageOut[kSlots];
    cudaEvent_t evIn[kSlots] = {}, evK[kSlots] = {}, evOut[kSlots] = {};
    bool slotUsed[kSlots] = {};
    // bounce ring for host sources that are NOT page-locked (a variable mapped from a file): host threads copy
    // chunks into pinned slots, the copy engine takes them from there (capi.cu: upload_unpinned)
    static constexpr int kBounce = 32;
    static constexpr size_t kBounceBytes = (size_t)4 << 20;
    unsigned char *bounce = nullptr;
    cudaEvent_t evBounce[kBounce] = {};
    unsigned long long h2dBytes = 0, d2hBytes = 0;  // field bytes moved by host-buffer applies / downloads since init
    unsigned slotCursor = 0;
    cudaEvent_t evDl = nullptr;           // mprg_download ordering
    mprg::DevBuf<unsigned char> userScratch[8];  // mprg_scratch slots
    void *peekBuf = nullptr;              // pinned, device-visible: scalar readbacks without a copy engine (peek)
    mprg::PinnedBuf blockMarks;           // pinned, device-visible: per-256-cell "referenced" flags of a route (route_finish)
    std::map<std::string, void *> ipcOpen;  // peer allocations mapped with mprg_ipc_open (handle bytes -> base)
    void *nccl = nullptr;                 // ncclComm_t
    void *ncclLib = nullptr;
    // optional per-launch profiling of the apply kernels (mprg_profile_*)
    struct ProfRec { int kind; double algBytes; double units; cudaEvent_t a, b; };
    bool profile = false;
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> evPool;
};

namespace mprg {

// reference's balanced block splitter, model_grid.F90:2428-2441 (0-based, half-open)
inline void para_range(int32_t n, int nprocs, int irank, int32_t *begin, int32_t *end) {
    int32_t w1 = n / nprocs, w2 = n % nprocs;
    int32_t ista = irank * w1 + (irank < w2 ? irank : w2);
    int32_t iend = ista + w1 + (w2 > irank ? 1 : 0);
    *begin = ista;
    *end = iend;
}

// Scalars move between host and device WITHOUT a copy engine: during a host-buffer pass the DMA queues
// are busy with 100-MB field transfers and a 4-byte cudaMemcpy would wait behind them (milliseconds per
// readback, a dozen readbacks per weight generation).  peek: a one-block kernel stores the words into
// pinned, device-visible memory, then the stream is synchronised.  poke: the bytes travel as kernel
// parameters.  bytes must be a multiple of 4, at most 64.  (locate.cu)
void peek(mprg_ctx *ctx, void *host, const void *dev, size_t bytes);
void poke(mprg_ctx *ctx, void *dev, const void *host, size_t bytes);

// bvh.cu
void bvh_build_points(mprg_ctx *ctx, const double *xyz_dev, int32_t n, Bvh &out, DevBuf<double> *sortedXyz);
void bvh_build_boxes(mprg_ctx *ctx, const float *lo_dev, const float *hi_dev, int32_t n, Bvh &out,
                     bool keepPrimBoxes = false);

// mesh.cu
void mesh_set(mprg_ctx *ctx, int32_t nCells, int32_t nVertices, int32_t maxEdges, const double *lonC,
              const double *latC, const double *lonV, const double *latV, const int32_t *voc);
void target_set(mprg_ctx *ctx, int stagger, int32_t ni, int32_t nj, const double *lon, const double *lat);
void target_register(mprg_ctx *ctx, int stagger, int32_t ni, int32_t nj);
// target_gen.cu: coordinates, map factors and rotation angles of a projected target grid, on the device
void target_generate(mprg_ctx *ctx, int stagger, int32_t ni, int32_t nj, const mprg_projection *p);
void target_map_factor(mprg_ctx *ctx, int stagger, int proj_code, double truelat1, double truelat2, double *out_dev);
void target_rotang(mprg_ctx *ctx);   // fills ctx->cosa / ctx->sina from the CENTER stagger's lon / lat
void mesh_need_cell_bvh(mprg_ctx *ctx);
void mesh_need_tri_bvh(mprg_ctx *ctx);
void mesh_need_poly_bvh(mprg_ctx *ctx);

// a rank that owns no destination rows (nranks > nj): an empty, valid CSR
inline bool route_empty_slab(mprg_ctx *ctx, mprg_route *r, int64_t nDst, int64_t nSrc) {
    if (nDst > 0) return false;
    r->nDst = 0; r->nnz = 0; r->nSrc = nSrc;
    r->rowptr.alloc(1); r->col.alloc(1); r->w.alloc(1);
    MPRG_CUDA(cudaMemsetAsync(r->rowptr.p, 0, sizeof(int32_t), ctx->stream));
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
    return true;
}

// locate.cu
void store_nearest(mprg_ctx *ctx, mprg_route *r);
void store_bilinear_element(mprg_ctx *ctx, mprg_route *r);
// conserve.cu / stagger.cu / polygon
void store_conserve(mprg_ctx *ctx, mprg_route *r);
void store_bilinear_grid(mprg_ctx *ctx, mprg_route *r);
void store_bilinear_node(mprg_ctx *ctx, mprg_route *r);
// compose.cu: stagger x rotation x bilinear in one matrix; false = not composable (the caller keeps the chain)
bool store_wind_composed(mprg_ctx *ctx, mprg_route *r, const mprg_route *stag, const mprg_route *bil);
void route_finish(mprg_ctx *ctx, mprg_route *r);  // stats + fp32 weight copy
void route_tile_stats(mprg_ctx *ctx, mprg_route *r);  // apply.cu
void route_plane_stats(mprg_ctx *ctx, mprg_route *r); // apply.cu

// apply.cu
struct ApplyField {
    const void *src;
    void *dst;
    int32_t nlev;
    int32_t epi_op;
    double epi_arg;
};
void apply_device(mprg_ctx *ctx, const mprg_route *r, const ApplyField *fields, int nfields, int src_dtype,
                  int dst_dtype, bool into_full = false);
void apply_wind_device(mprg_ctx *ctx, const mprg_route *r, const void *u, const void *v, int32_t nlev, int src_dtype,
                       void *dst, int dst_dtype, bool into_full);
void rotate_device(mprg_ctx *ctx, int stagger, void *u, void *v, int32_t nlev, int dtype);
void rotation_constants(mprg_ctx *ctx, int64_t n);  // fills ctx->rotc from ctx->cosa / ctx->sina

void bswap_device(mprg_ctx *ctx, void *x, size_t count, size_t elem, cudaStream_t st);
void post_affine_device(mprg_ctx *ctx, void *x, size_t count, int dtype, double scale, double offset);
void post_midlevels_device(mprg_ctx *ctx, int64_t n, int32_t nlev, int dtype, const void *x, void *mid);
void post_ptop_device(mprg_ctx *ctx, int64_t n, int32_t nlev, int dtype, const void *x, double *out2_dev);

// wcache.cu: cross-run weight cache
bool wcache_load(mprg_ctx *ctx, mprg_route *r, unsigned long long key[2]);
void wcache_save(mprg_ctx *ctx, const mprg_route *r, const unsigned long long key[2]);

// gather.cu
void comm_destroy(mprg_ctx *ctx);
void comm_id(mprg_ctx *ctx, void *id128);
void comm_init(mprg_ctx *ctx, const void *id128);
void gather_slabs(mprg_ctx *ctx, int nfields, const int *stagger, const int32_t *nlev, int dtype,
                  const void *const *slab_dev, int root, void *const *full_dev);

}  // namespace mprg
