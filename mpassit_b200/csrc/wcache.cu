// wcache.cu -- cross-run weight cache.
//
// The reference regenerates every regridding matrix in every run (12 RegridStore calls per run,
// interp.F90:123-437; weights never kept, program_setup.F90:72-75).  Weights depend on geometry only, so a
// route is stored under a key derived from everything that determines it:
//     source mesh (cell + vertex Cartesian coordinates, verticesOnCell), the destination points of this rank's
//     slab (and, for grid sources, the source rows), grid topology, method, source location, destination
//     stagger, slab bounds, and the engine's weight-format version.
// The key is computed ON THE DEVICE from the arrays the weight generation itself reads (a 128-bit
// order-independent sum of per-element mixes: two passes over data that is already in HBM), so equal keys mean
// bit-identical inputs to the store kernels.  A cache file holds the CSR of the slab; the tile schedule and the
// fp32 weight copy are rebuilt by route_finish (a millisecond).  A file that is truncated, of another version or
// of another key is ignored and rewritten.
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <string>

#include "common.cuh"

namespace mprg {

constexpr uint32_t kCacheMagic = 0x5747504du;   // "MPGW"
constexpr uint32_t kCacheVersion = 2;           // bump when a store kernel's output may change

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x += 0x9e3779b97f4a7c15ULL;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}

// h[0] += sum mix(i, word_i), h[1] += sum mix'(i, word_i): position-dependent, order-independent accumulation
__global__ void k_hash_words(const uint32_t *__restrict__ w, size_t n, unsigned long long salt, unsigned long long *h) {
    unsigned long long a = 0, b = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long v = ((unsigned long long)w[i] << 32) ^ (unsigned long long)i;
        a += mix64(v ^ salt);
        b += mix64(~v + salt * 0x2545f4914f6cdd1dULL);
    }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(h, a);
        atomicAdd(h + 1, b);
    }
}

static void hash_dev(mprg_ctx *ctx, const void *p, size_t bytes, unsigned long long salt, unsigned long long *h_dev) {
    const size_t n = bytes / 4;
    if (!p || n == 0) return;
    const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((n + 1023) / 1024, (size_t)kNumSM * 8));
    k_hash_words<<<grid, 256, 0, ctx->stream>>>((const uint32_t *)p, n, salt, h_dev);
    ctx->launches++;
}

// key of the route `r` is about to become (method / src_loc / dst_stagger already set)
static void route_key(mprg_ctx *ctx, const mprg_route *r, unsigned long long key[2]) {
    DevBuf<unsigned long long> h(2);
    MPRG_CUDA(cudaMemsetAsync(h.p, 0, 2 * sizeof(unsigned long long), ctx->stream));
    const Mesh &m = ctx->mesh;
    const Target &tg = ctx->target[r->dst_stagger];
    unsigned long long salt = 1;
    if (r->src_loc != MPRG_SRC_GRID_CENTER) {
        hash_dev(ctx, m.cellXyz.p, (size_t)m.nCells * 3 * sizeof(double), salt++, h.p);
        hash_dev(ctx, m.vertXyz.p, (size_t)m.nVertices * 3 * sizeof(double), salt++, h.p);
        hash_dev(ctx, m.voc.p, (size_t)m.nCells * m.maxEdges * sizeof(int32_t), salt++, h.p);
    } else {
        const Target &src = ctx->target[MPRG_CENTER_HALO];
        salt = 10;
        hash_dev(ctx, src.x() + 3 * src.slabOffset(), (size_t)src.nSlab() * 3 * sizeof(double), salt++, h.p);
    }
    salt = 20;
    hash_dev(ctx, tg.x() + 3 * tg.slabOffset(), (size_t)tg.nSlab() * 3 * sizeof(double), salt++, h.p);
    if (r->method == MPRG_CONSERVE) {   // destination cells are the quads of CORNER points around the centres
        const Target &co = ctx->target[MPRG_CORNER];
        if (co.set) hash_dev(ctx, co.x() + 3 * (size_t)tg.j0 * co.ni, (size_t)(tg.j1 - tg.j0 + 1) * co.ni * 3 * sizeof(double), salt++, h.p);
    }
    unsigned long long hk[2] = {0, 0};
    peek(ctx, hk, h.p, sizeof hk);
    // scalars that shape the route
    const unsigned long long sc[] = {(unsigned long long)r->method, (unsigned long long)r->src_loc, (unsigned long long)r->dst_stagger,
                                     (unsigned long long)tg.ni, (unsigned long long)tg.nj, (unsigned long long)tg.j0,
                                     (unsigned long long)tg.j1, (unsigned long long)ctx->gridKind, (unsigned long long)m.nCells,
                                     (unsigned long long)m.nVertices, (unsigned long long)m.maxEdges, (unsigned long long)kCacheVersion};
    for (size_t i = 0; i < sizeof sc / sizeof sc[0]; ++i) {
        unsigned long long x = sc[i] + 0x9e3779b97f4a7c15ULL * (i + 1);
        x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
        x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
        hk[0] = (hk[0] ^ (x ^ (x >> 31))) * 0x100000001b3ULL;
        hk[1] += (x ^ (x >> 29)) + (hk[0] << 7);
    }
    key[0] = hk[0];
    key[1] = hk[1];
}

struct CacheHeader {
    uint32_t magic, version;
    unsigned long long key[2];
    int64_t nDst, nnz, nSrc, srcPlane;
    int32_t method, src_loc, dst_stagger, srcLevelSlowest;
};

static std::string cache_path(const mprg_ctx *ctx, const unsigned long long key[2]) {
    char name[96];
    snprintf(name, sizeof name, "/mprg_%016llx%016llx.w", key[0], key[1]);
    return ctx->cacheDir + name;
}

// true: `r` now holds the cached CSR (route_finish still to be called by the caller)
bool wcache_load(mprg_ctx *ctx, mprg_route *r, unsigned long long key[2]) {
    route_key(ctx, r, key);
    const std::string path = cache_path(ctx, key);
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    CacheHeader h;
    bool ok = fread(&h, sizeof h, 1, f) == 1 && h.magic == kCacheMagic && h.version == kCacheVersion && h.key[0] == key[0] &&
              h.key[1] == key[1] && h.method == r->method && h.src_loc == r->src_loc && h.dst_stagger == r->dst_stagger &&
              h.nDst >= 0 && h.nnz >= 0 && h.nDst == ctx->target[r->dst_stagger].nSlab();
    std::vector<int32_t> rowptr, col;
    std::vector<double> w;
    if (ok) {
        rowptr.resize(h.nDst + 1);
        col.resize(h.nnz);
        w.resize(h.nnz);
        ok = fread(rowptr.data(), 4, rowptr.size(), f) == rowptr.size() &&
             (h.nnz == 0 || (fread(col.data(), 4, col.size(), f) == col.size() && fread(w.data(), 8, w.size(), f) == w.size()));
        ok = ok && rowptr[0] == 0 && rowptr[h.nDst] == h.nnz;
    }
    fclose(f);
    if (!ok) return false;
    r->nDst = h.nDst; r->nnz = h.nnz; r->nSrc = h.nSrc;
    r->srcLevelSlowest = h.srcLevelSlowest != 0;
    r->srcPlane = h.srcPlane;
    r->rowptr.alloc(h.nDst + 1);
    r->col.alloc(h.nnz > 0 ? h.nnz : 1);
    r->w.alloc(h.nnz > 0 ? h.nnz : 1);
    MPRG_CUDA(cudaMemcpyAsync(r->rowptr.p, rowptr.data(), rowptr.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (h.nnz) {
        MPRG_CUDA(cudaMemcpyAsync(r->col.p, col.data(), col.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        MPRG_CUDA(cudaMemcpyAsync(r->w.p, w.data(), w.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));   // the host vectors go out of scope
    ctx->cacheHits++;
    return true;
}

// best effort: a cache that cannot be written is not an error
void wcache_save(mprg_ctx *ctx, const mprg_route *r, const unsigned long long key[2]) {
    CacheHeader h;
    memset(&h, 0, sizeof h);
    h.magic = kCacheMagic; h.version = kCacheVersion;
    h.key[0] = key[0]; h.key[1] = key[1];
    h.nDst = r->nDst; h.nnz = r->nnz; h.nSrc = r->nSrc; h.srcPlane = r->srcPlane;
    h.method = r->method; h.src_loc = r->src_loc; h.dst_stagger = r->dst_stagger; h.srcLevelSlowest = r->srcLevelSlowest ? 1 : 0;
    std::vector<int32_t> rowptr(r->nDst + 1), col(r->nnz);
    std::vector<double> w(r->nnz);
    MPRG_CUDA(cudaMemcpy(rowptr.data(), r->rowptr.p, rowptr.size() * 4, cudaMemcpyDeviceToHost));
    if (r->nnz) {
        MPRG_CUDA(cudaMemcpy(col.data(), r->col.p, col.size() * 4, cudaMemcpyDeviceToHost));
        MPRG_CUDA(cudaMemcpy(w.data(), r->w.p, w.size() * 8, cudaMemcpyDeviceToHost));
    }
    mkdir(ctx->cacheDir.c_str(), 0777);
    const std::string path = cache_path(ctx, key), tmp = path + ".tmp" + std::to_string((long)getpid()) + "." + std::to_string(ctx->rank);
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    bool ok = fwrite(&h, sizeof h, 1, f) == 1 && fwrite(rowptr.data(), 4, rowptr.size(), f) == rowptr.size() &&
              (r->nnz == 0 || (fwrite(col.data(), 4, col.size(), f) == col.size() && fwrite(w.data(), 8, w.size(), f) == w.size()));
    ok = (fclose(f) == 0) && ok;
    if (ok && rename(tmp.c_str(), path.c_str()) == 0) ctx->cacheStores++;   // atomic: readers never see a partial file
    else remove(tmp.c_str());
}

}  // namespace mprg
