// locate.cu -- weight generation ("RegridStore") for the point-location methods:
// NEAREST_STOD and BILINEAR from mesh elements (dual / Delaunay triangles).
//
// One thread per destination point of this rank's slab.  Candidates come from
// the implicit BVH (bvh.cuh); every decision is fp64 with -fmad=false.
// Output is CSR over the slab (rows in slab order).
#include <cub/device/device_scan.cuh>

#include "bvh.cuh"
#include "common.cuh"

namespace mprg {

// ---- scalar traffic without a copy engine (common.cuh) -------------------------
struct PokeWords { uint32_t w[16]; };
__global__ void k_peek(const uint32_t *__restrict__ dev, uint32_t *host_visible, int nwords) {
    if ((int)threadIdx.x < nwords) host_visible[threadIdx.x] = dev[threadIdx.x];
}
__global__ void k_poke(uint32_t *dev, PokeWords p, int nwords) {
    if ((int)threadIdx.x < nwords) dev[threadIdx.x] = p.w[threadIdx.x];
}
void peek(mprg_ctx *ctx, void *host, const void *dev, size_t bytes) {
    if (bytes == 0) return;
    if (bytes % 4 || bytes > 64) fail(95, "peek: %zu bytes", bytes);
    k_peek<<<1, 32, 0, ctx->stream>>>((const uint32_t *)dev, (uint32_t *)ctx->peekBuf, (int)(bytes / 4));
    MPRG_CUDA(cudaGetLastError());
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(host, ctx->peekBuf, bytes);
}
void poke(mprg_ctx *ctx, void *dev, const void *host, size_t bytes) {
    if (bytes == 0) return;
    if (bytes % 4 || bytes > 64) fail(95, "poke: %zu bytes", bytes);
    PokeWords p;
    memcpy(p.w, host, bytes);
    k_poke<<<1, 32, 0, ctx->stream>>>((uint32_t *)dev, p, (int)(bytes / 4));
    MPRG_CUDA(cudaGetLastError());
}

// ---- NEAREST_STOD (interp.F90:420-431; soil bundle :436-443) ----------------
__global__ void __launch_bounds__(128)
k_nearest(BvhView bvh, const double *__restrict__ sortedXyz, const double *__restrict__ dstXyz, int64_t nDst,
          int32_t *__restrict__ rowptr, int32_t *__restrict__ col, double *__restrict__ w) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > nDst) return;
    if (t == nDst) { rowptr[t] = (int32_t)nDst; return; }
    d3 q = ld3(dstXyz + 3 * t);
    double best = 1e300;
    int32_t bid = 0x7fffffff;
    bvh_nearest(bvh, sortedXyz, q, best, bid);
    rowptr[t] = (int32_t)t;
    col[t] = bid;
    w[t] = 1.0;
}

void store_nearest(mprg_ctx *ctx, mprg_route *r) {
    mesh_need_cell_bvh(ctx);
    Mesh &m = ctx->mesh;
    Target &tg = ctx->target[r->dst_stagger];
    int64_t n = tg.nSlab();
    if (route_empty_slab(ctx, r, n, m.nCells)) return;
    r->nDst = n; r->nnz = n; r->nSrc = m.nCells;
    r->rowptr.alloc(n + 1); r->col.alloc(n); r->w.alloc(n);
    BvhView v{m.cellBvh.nodes.p, m.cellBvh.primId.p, m.cellBvh.nLeafNodes, m.cellBvh.nPrim};
    k_nearest<<<(unsigned)((n + 1 + 127) / 128), 128, 0, ctx->stream>>>(
        v, m.cellSorted.p, tg.x() + 3 * tg.slabOffset(), n, r->rowptr.p, r->col.p, r->w.p);
    ctx->launches++;
    MPRG_CUDA(cudaGetLastError());
}

// ---- BILINEAR, Mesh(element) -> Grid (interp.F90:123,207,226,241,259,277,334) --
__global__ void __launch_bounds__(128)
k_bilinear_tri(BvhView bvh, const int32_t *__restrict__ tri, const double *__restrict__ cxyz,
               const double *__restrict__ dstXyz, int64_t nDst, int32_t *__restrict__ elem,
               int32_t *__restrict__ ecol, double *__restrict__ ew, int32_t *__restrict__ cnt) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDst) return;
    d3 p = ld3(dstXyz + 3 * t);
    int32_t best = -1;
    double bw[3] = {0.0, 0.0, 0.0};
    bvh_overlap(bvh, p, p, [&](int s0, int s1) {
        for (int s = s0; s < s1; ++s) {
            int32_t v = __ldg(bvh.primId + s);
            if (best >= 0 && v >= best) continue;  // smallest dual-element id wins
            const int32_t *tv = tri + 3 * (size_t)v;
            int32_t c0 = __ldg(tv);
            if (c0 < 0) continue;
            int32_t c1 = __ldg(tv + 1), c2 = __ldg(tv + 2);
            double w[3];
            if (tri_locate(ld3(cxyz + 3 * (size_t)c0), ld3(cxyz + 3 * (size_t)c1), ld3(cxyz + 3 * (size_t)c2), p,
                           kTol, w)) {
                best = v;
                bw[0] = w[0]; bw[1] = w[1]; bw[2] = w[2];
            }
        }
    });
    elem[t] = best;
    if (best >= 0) {
        const int32_t *tv = tri + 3 * (size_t)best;
        for (int k = 0; k < 3; ++k) { ecol[3 * t + k] = tv[k]; ew[3 * t + k] = bw[k]; }
        cnt[t] = 3;
    } else {
        cnt[t] = 0;
    }
    if (t == 0) cnt[nDst] = 0;
}

__global__ void k_compact3(int64_t nDst, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ ecol,
                           const double *__restrict__ ew, int32_t *__restrict__ col, double *__restrict__ w) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDst) return;
    int32_t b = rowptr[t], e = rowptr[t + 1];
    for (int k = 0; k < e - b; ++k) { col[b + k] = ecol[3 * t + k]; w[b + k] = ew[3 * t + k]; }
}

// exclusive scan of per-row counts (cnt has nDst+1 entries, last = 0) into rowptr
void scan_counts(mprg_ctx *ctx, const int32_t *cnt, int32_t *rowptr, int64_t nPlus1) {
    size_t tmpBytes = 0;
    MPRG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, cnt, rowptr, (int)nPlus1, ctx->stream));
    DevBuf<unsigned char> tmp(tmpBytes);
    MPRG_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmpBytes, cnt, rowptr, (int)nPlus1, ctx->stream));
    ctx->launches++;
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
}

void store_bilinear_element(mprg_ctx *ctx, mprg_route *r) {
    mesh_need_tri_bvh(ctx);
    Mesh &m = ctx->mesh;
    Target &tg = ctx->target[r->dst_stagger];
    int64_t n = tg.nSlab();
    if (route_empty_slab(ctx, r, n, m.nCells)) return;
    r->nDst = n; r->nSrc = m.nCells;
    DevBuf<int32_t> elem(n), ecol(3 * n), cnt(n + 1);
    DevBuf<double> ew(3 * n);
    BvhView v{m.triBvh.nodes.p, m.triBvh.primId.p, m.triBvh.nLeafNodes, m.triBvh.nPrim};
    k_bilinear_tri<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(
        v, m.tri.p, m.cellXyz.p, tg.x() + 3 * tg.slabOffset(), n, elem.p, ecol.p, ew.p, cnt.p);
    ctx->launches++;
    MPRG_CUDA(cudaGetLastError());
    r->rowptr.alloc(n + 1);
    scan_counts(ctx, cnt.p, r->rowptr.p, n + 1);
    int32_t nnz = 0;
    peek(ctx, &nnz, r->rowptr.p + n, sizeof(int32_t));
    r->nnz = nnz;
    r->col.alloc(nnz > 0 ? nnz : 1);
    r->w.alloc(nnz > 0 ? nnz : 1);
    k_compact3<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, r->rowptr.p, ecol.p, ew.p, r->col.p, r->w.p);
    ctx->launches++;
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));
}

// ---- route statistics + fp32 weight copy ------------------------------------
__global__ void k_route_stats(int64_t nDst, const int32_t *__restrict__ rowptr, unsigned long long *nUnmapped,
                              int32_t *maxRow, int32_t *minMappedRow) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDst) return;
    int32_t len = rowptr[t + 1] - rowptr[t];
    if (len == 0) atomicAdd(nUnmapped, 1ULL);
    else atomicMin(minMappedRow, len);
    atomicMax(maxRow, len);
}

__global__ void k_mark_cols(int64_t nnz, const int32_t *__restrict__ col, unsigned char *__restrict__ mark) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) mark[col[i]] = 1;
}
// out[0] += referenced sources; out[1] = lowest, out[2] = highest referenced source id
__global__ void k_count_marks(int64_t n, const unsigned char *__restrict__ mark, unsigned long long *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned v = (i < n && mark[i]) ? 1u : 0u;
    unsigned b = __ballot_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && b) {
        atomicAdd(out, (unsigned long long)__popc(b));
        const int64_t w0 = i;  // lane 0 of the warp
        atomicMin(out + 1, (unsigned long long)(w0 + __ffs(b) - 1));
        atomicMax(out + 2, (unsigned long long)(w0 + 31 - __clz(b)));
    }
}

// one flag per kMarkBlock source ids: any of them referenced?  Written straight into pinned host memory.
constexpr int kMarkBlock = 256;
__global__ void __launch_bounds__(kMarkBlock)
k_block_marks(int64_t n, const unsigned char *__restrict__ mark, unsigned char *host_visible) {
    const int64_t i = (int64_t)blockIdx.x * kMarkBlock + threadIdx.x;
    const int any = __syncthreads_or(i < n && mark[i]);
    if (threadIdx.x == 0) host_visible[blockIdx.x] = any ? 1 : 0;
}

__global__ void k_w32(int64_t nnz, const double *__restrict__ w, float *__restrict__ w32) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) w32[i] = (float)w[i];
}

// rows longer than kLongRow: counted (out == nullptr) or listed
__global__ void k_long_rows(int64_t nDst, const int32_t *__restrict__ rowptr, int32_t *__restrict__ out,
                            unsigned long long *count) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDst || rowptr[t + 1] - rowptr[t] <= kLongRow) return;
    const unsigned long long i = atomicAdd(count, 1ULL);
    if (out) out[i] = (int32_t)t;
}

void route_finish(mprg_ctx *ctx, mprg_route *r) {
    DevBuf<unsigned long long> un(1);
    DevBuf<int32_t> mm(2);
    int32_t init[2] = {0, 0x7fffffff};
    MPRG_CUDA(cudaMemsetAsync(un.p, 0, sizeof(unsigned long long), ctx->stream));
    poke(ctx, mm.p, init, sizeof init);
    if (r->nDst > 0) {
        k_route_stats<<<(unsigned)((r->nDst + 255) / 256), 256, 0, ctx->stream>>>(r->nDst, r->rowptr.p, un.p, mm.p,
                                                                                 mm.p + 1);
        ctx->launches++;
    }
    r->w32.alloc(r->nnz > 0 ? r->nnz : 1);
    if (r->nnz > 0) {
        k_w32<<<(unsigned)((r->nnz + 255) / 256), 256, 0, ctx->stream>>>(r->nnz, r->w.p, r->w32.p);
        ctx->launches++;
    }
    if (r->composite) {
        r->w2_32.alloc(r->nnz > 0 ? r->nnz : 1);
        if (r->nnz > 0) {
            k_w32<<<(unsigned)((r->nnz + 255) / 256), 256, 0, ctx->stream>>>(r->nnz, r->w2.p, r->w2_32.p);
            ctx->launches++;
        }
    }
    DevBuf<unsigned long long> nref(3);
    const unsigned long long nref0[3] = {0ULL, ~0ULL, 0ULL};
    poke(ctx, nref.p, nref0, sizeof nref0);
    DevBuf<unsigned char> mark;
    int64_t nblk = 0;
    if (r->nnz > 0 && r->nSrc > 0) {
        mark.alloc(r->nSrc);
        MPRG_CUDA(cudaMemsetAsync(mark.p, 0, r->nSrc, ctx->stream));
        k_mark_cols<<<(unsigned)((r->nnz + 255) / 256), 256, 0, ctx->stream>>>(r->nnz, r->col.p, mark.p);
        k_count_marks<<<(unsigned)((r->nSrc + 255) / 256), 256, 0, ctx->stream>>>(r->nSrc, mark.p, nref.p);
        ctx->launches += 2;
        if (!r->srcLevelSlowest) {
            nblk = (r->nSrc + kMarkBlock - 1) / kMarkBlock;
            ctx->blockMarks.ensure((size_t)nblk);
            k_block_marks<<<(unsigned)nblk, kMarkBlock, 0, ctx->stream>>>(r->nSrc, mark.p, (unsigned char *)ctx->blockMarks.p);
            ctx->launches++;
        }
    }
    unsigned long long hun = 0, href[3] = {0, 0, 0};
    int32_t hmm[2] = {0, 0};
    peek(ctx, href, nref.p, sizeof href);
    peek(ctx, &hun, un.p, sizeof hun);
    peek(ctx, hmm, mm.p, sizeof hmm);
    r->nUnmapped = (int64_t)hun;
    r->nSrcRef = (int64_t)href[0];
    // contiguous id range that holds every referenced source: host-buffer applies upload only this range
    r->srcLo = href[0] ? (int64_t)href[1] : 0;
    r->srcHi = href[0] ? (int64_t)href[2] + 1 : 0;
    // (the peeks above synchronised the stream: the block flags are in host memory)
    r->srcRanges.clear();
    if (nblk > 0 && href[0]) {
        const unsigned char *bm = (const unsigned char *)ctx->blockMarks.p;
        for (int64_t gap = 4;; gap *= 2) {   // merge gaps of up to `gap` blocks; widen until the list is short
            r->srcRanges.clear();
            int64_t b = 0;
            while (b < nblk) {
                if (!bm[b]) { ++b; continue; }
                int64_t e = b + 1, last = b;   // last marked block of the range
                while (e < nblk && e - last <= gap) {
                    if (bm[e]) last = e;
                    ++e;
                }
                r->srcRanges.emplace_back(std::max(b * kMarkBlock, r->srcLo), std::min((last + 1) * kMarkBlock, r->srcHi));
                b = last + 1;
            }
            if (r->srcRanges.size() <= 256) break;
        }
    }
    if (!r->srcLevelSlowest) route_tile_stats(ctx, r);
    r->maxRow = hmm[0];
    r->uniform = (r->nnz > 0 && hmm[0] == hmm[1]);
    route_plane_stats(ctx, r);
    r->nLong = 0;
    if (r->srcLevelSlowest && r->maxRow > kLongRow) {
        const unsigned g = (unsigned)((r->nDst + 255) / 256);
        MPRG_CUDA(cudaMemsetAsync(un.p, 0, sizeof(unsigned long long), ctx->stream));
        k_long_rows<<<g, 256, 0, ctx->stream>>>(r->nDst, r->rowptr.p, nullptr, un.p);
        peek(ctx, &hun, un.p, sizeof hun);
        r->longRows.alloc(hun);
        MPRG_CUDA(cudaMemsetAsync(un.p, 0, sizeof(unsigned long long), ctx->stream));
        k_long_rows<<<g, 256, 0, ctx->stream>>>(r->nDst, r->rowptr.p, r->longRows.p, un.p);
        ctx->launches += 2;
        r->nLong = (int64_t)hun;
    }
}

}  // namespace mprg
