// Weight application for grid-source routes (centre -> EDGE1 / EDGE2 staggering, interp.F90:298-325), staged with
// TMA bulk copies.  (Included by apply.cu after apply_pipe.cuh.)
//
// The source is a [lev][row][srcNi] field.  A tile of <= 256 consecutive targets of one destination row references,
// per level, a short window of two or three consecutive source rows.  The register-gather kernel (k_apply_planes)
// fetches them with four misaligned 128-byte gathers per warp per level and is bound by L1 wavefronts; here each
// (level, source row) window is one cp.async.bulk into shared memory -- no LSU traffic for the loads -- kPlLev
// levels per stage, kPlStages stages in flight, completion on one mbarrier per stage.  Lanes then read their <= 4
// entries with 4-byte shared loads (consecutive lanes, consecutive words) and store coalesced, streaming.
// Windows are fetched as the 16-byte-aligned span around them, read where they lie (the same absolute-address rule as
// the column kernel's unaligned units).  A tile whose windows do not fit (more than kPlWin source rows, or wider than
// the tile: the seam of a periodic grid) runs the register-gather code instead, CTA-uniformly.
#pragma once

namespace mprg {

constexpr int kPlTile = 256;    // threads = targets per tile (upper bound; tiles split a row evenly)
constexpr int kPlWin = 3;       // source rows a tile may reference (the launch sizes shared memory for what the route needs)

// one (level, source row) window slot of a stage
template <typename TIN>
__host__ __device__ constexpr unsigned pl_win_bytes() { return (unsigned)(((kPlTile + 4) * sizeof(TIN) + 32 + 15) & ~15u); }

// one thread, one target, every level of one field, straight from global memory
template <typename TIN, typename TOUT, typename TACC>
__device__ __forceinline__ void planes_direct(const ApplyArgs<TACC> &a, const FieldDev &fd, int64_t t, int ne, const int (&c)[kFlatRow],
                                              const TACC (&w)[kFlatRow]) {
    const TIN *__restrict__ src = (const TIN *)fd.src;
    TOUT *__restrict__ dst = (TOUT *)fd.dst;
    int lev = 0;
    if (ne > 0) {
        // kPlaneBatch levels x 4 gathers in flight per thread
        for (; lev + kPlaneBatch <= fd.nlev; lev += kPlaneBatch) {
            TIN x[kPlaneBatch][kFlatRow];
#pragma unroll
            for (int q = 0; q < kPlaneBatch; ++q) {
                const TIN *pl = src + (size_t)(lev + q) * a.srcPlane;
#pragma unroll
                for (int k = 0; k < kFlatRow; ++k) x[q][k] = (k < ne) ? __ldg(pl + c[k]) : (TIN)0;
            }
#pragma unroll
            for (int q = 0; q < kPlaneBatch; ++q) {
                TACC acc = 0;
#pragma unroll
                for (int k = 0; k < kFlatRow; ++k)
                    if (k < ne) acc += w[k] * (TACC)x[q][k];
                st_stream(dst + (size_t)(lev + q) * a.dstLev + a.dstOff + t, (TOUT)epilogue(acc, fd.epi_op, fd.epi_arg));
            }
        }
    }
    for (; lev < fd.nlev; ++lev) {
        const TIN *pl = src + (size_t)lev * a.srcPlane;
        TACC acc = 0;
#pragma unroll
        for (int k = 0; k < kFlatRow; ++k)
            if (k < ne) acc += w[k] * (TACC)__ldg(pl + c[k]);
        st_stream(dst + (size_t)lev * a.dstLev + a.dstOff + t, (TOUT)epilogue(acc, fd.epi_op, fd.epi_arg));
    }
}

struct PlaneGeom {
    int32_t tilesPerRow, tw;   // tiles per destination row, targets per tile
    int32_t dstNi, srcNi;
    int32_t nWin;              // source rows per tile the stages have room for (tiles needing more take the gather path)
    // per-tile headers written by k_plane_stats when the route is built: first source row, window extents per source
    // row and whether the windows fit -- [tile][8] = y0, xmin[3], xmax[3], fits.  The producer warp starts streaming
    // one memory latency after the CTA starts instead of after the consumers' three dependent CSR loads.
    const int32_t *tiles;
};

// Route statistics for the launch: the largest number of source rows any tile references whose windows fit
// (out[0]), and how many tiles do not fit at all (out[1]).
__global__ void __launch_bounds__(kPlTile)
k_plane_stats(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, PlaneGeom g, int32_t *__restrict__ out,
              int32_t *__restrict__ tiles) {
    __shared__ int s_y0, s_xmin[kPlWin + 1], s_xmax[kPlWin + 1], s_bad;
    const int tid = threadIdx.x;
    const int j = blockIdx.x / g.tilesPerRow, tx = blockIdx.x - j * g.tilesPerRow;
    const int i0 = tx * g.tw, cnt = min(g.tw, g.dstNi - i0);
    const int64_t t = (int64_t)j * g.dstNi + i0 + tid;
    int b = 0, ne = 0;
    if (tid < cnt) { b = rowptr[t]; ne = rowptr[t + 1] - b; }
    if (ne > kLongRow) ne = 0;
    const int nMapped = __syncthreads_count(tid < cnt && ne > 0);
    if (tid == 0) {
        s_y0 = 0x7fffffff; s_bad = 0;
        for (int q = 0; q <= kPlWin; ++q) { s_xmin[q] = 0x7fffffff; s_xmax[q] = -1; }
    }
    __syncthreads();
    for (int k = 0; k < ne; ++k) atomicMin(&s_y0, col[b + k] / g.srcNi);
    __syncthreads();
    const int y0 = s_y0;
    for (int k = 0; k < ne; ++k) {
        const int c = col[b + k], y = c / g.srcNi, x = c - y * g.srcNi, q = y - y0;
        if (q >= kPlWin) s_bad = 1;
        else { atomicMin(&s_xmin[q], x); atomicMax(&s_xmax[q], x); }
    }
    __syncthreads();
    if (tid == 0) {
        bool bad = s_bad != 0 || y0 == 0x7fffffff;
        int nw = 0;
        for (int q = 0; q < kPlWin; ++q) {
            if (s_xmax[q] >= 0) nw = q + 1;
            if (s_xmax[q] - s_xmin[q] + 1 > kPlTile + 4) bad = true;
        }
        if (y0 != 0x7fffffff) {
            if (bad) atomicAdd(out + 1, 1);
            else atomicMax(out, nw);
        }
        int32_t *h = tiles + 8 * (size_t)blockIdx.x;
        h[0] = y0;
        for (int q = 0; q < kPlWin; ++q) { h[1 + q] = s_xmin[q]; h[4 + q] = s_xmax[q]; }
        // bit 0: the windows fit (a tile with nothing mapped does not: it only writes zeros); bit 1: every target mapped
        h[7] = bad ? 0 : (1 | (nMapped == cnt ? 2 : 0));
    }
}

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

constexpr int kPlThreads = kPlTile + 32;   // 8 consumer warps (one target per lane) + 1 producer warp

// kPlLev levels per stage, kPlStages stages in flight; a stage is laid out [window][level][WCB]
template <typename TIN, typename TOUT, typename TACC, int kPlLev, int kPlStages>
__global__ void __launch_bounds__(kPlThreads, sizeof(TIN) == 4 ? 4 : 2)
k_apply_planes_pipe(ApplyArgs<TACC> a, const __grid_constant__ FieldPack fp, PlaneGeom g) {
    extern __shared__ __align__(128) unsigned char pl_smem[];
    __shared__ unsigned long long s_full[kPlStages], s_empty[kPlStages];
    constexpr unsigned ESZ = sizeof(TIN), WCB = pl_win_bytes<TIN>();
    const unsigned STB = (unsigned)g.nWin * kPlLev * WCB;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp == kPlTile / 32;
    const int j = blockIdx.x / g.tilesPerRow, tx = blockIdx.x - j * g.tilesPerRow;
    const int i0 = tx * g.tw, cnt = min(g.tw, g.dstNi - i0);
    const int64_t t = (int64_t)j * g.dstNi + i0 + tid;
    const bool live = tid < cnt;      // (producer lanes: tid >= kPlTile >= cnt)

    // the tile's header (k_plane_stats): every thread reads the same 32 bytes
    const int4 h0 = __ldg((const int4 *)(g.tiles + 8 * (size_t)blockIdx.x)), h1 = __ldg((const int4 *)(g.tiles + 8 * (size_t)blockIdx.x) + 1);
    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < kPlStages; ++q) { mbar_init(s_full + q, 1); mbar_init(s_empty + q, kPlTile / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const int y0 = h0.x;
    const int s_xmin[kPlWin] = {h0.y, h0.z, h0.w}, s_xmax[kPlWin] = {h1.x, h1.y, h1.z};
    const bool fits = (h1.w & 1) != 0;
    const bool fullTile = (h1.w & 2) != 0;     // every target of the tile is mapped: no per-level zero selects

    int ne = 0, b = 0;
    int c[kFlatRow] = {};
    TACC w[kFlatRow] = {};
    bool store = false;
    if (!producer || !fits) {     // (the producer warp of a staged tile needs nothing from the CSR)
        if (live) {
            b = __ldg(a.rowptr + t);
            ne = __ldg(a.rowptr + t + 1) - b;
        }
        store = live && ne <= kLongRow;   // longer rows: k_apply_planes_long
        if (!store) ne = 0;
#pragma unroll
        for (int k = 0; k < kFlatRow; ++k) {
            const bool h = k < ne;
            c[k] = h ? __ldg(a.col + b + k) : 0;
            w[k] = h ? __ldg(a.w + b + k) : (TACC)0;
        }
    }
    if (!fits) {
        // windows too wide / too many source rows (or nothing mapped: zeros): register-gather code, whole CTA
        if (store)
            for (int f = 0; f < a.nfields; ++f) planes_direct<TIN, TOUT, TACC>(a, fp.f[f], t, ne, c, w);
        return;
    }
    const size_t planeB = (size_t)a.srcPlane * ESZ;     // a multiple of 16 (checked by the launch)

    if (producer) {
        // lane = (level slot l, window q) of a stage; everything but the level is fixed per field
        const int l = lane / g.nWin, q = lane - l * g.nWin;
        int wd = 0, xm = 0;
        if (l < kPlLev) {
            const int hi = q == 0 ? s_xmax[0] : (q == 1 ? s_xmax[1] : s_xmax[2]);
            xm = q == 0 ? s_xmin[0] : (q == 1 ? s_xmin[1] : s_xmin[2]);
            wd = hi - xm + 1;
            if (hi < 0 || wd <= 0) wd = 0;  // no entry in this source row
        }
        const unsigned sdst0 = (unsigned)__cvta_generic_to_shared(pl_smem) + (unsigned)(q * kPlLev + l) * WCB;
        int n = 0;
        for (int f = 0; f < a.nfields; ++f) {
            const FieldDev &fd = fp.f[f];
            const uintptr_t aend = (uintptr_t)fd.src + (size_t)fd.nlev * planeB;
            const uintptr_t a00 = (uintptr_t)fd.src + ((size_t)(y0 + q) * g.srcNi + xm) * ESZ + (size_t)l * planeB;
            const uintptr_t ga0 = a00 & ~(uintptr_t)15;
            const unsigned nbytes = wd > 0 ? (unsigned)(((a00 + (size_t)wd * ESZ + 15) & ~(uintptr_t)15) - ga0) : 0u;
            for (int l0 = 0; l0 < fd.nlev; l0 += kPlLev, ++n) {
                const int slot = n % kPlStages;
                if (n >= kPlStages) mbar_wait(s_empty + slot, (unsigned)((n / kPlStages - 1) & 1));
                unsigned nb = (l0 + l < fd.nlev) ? nbytes : 0u;
                const uintptr_t ga = ga0 + (size_t)l0 * planeB;
                const unsigned sdst = sdst0 + (unsigned)slot * STB;
                if (nb && ga + nb > aend) {
                    // last window of the array: bulk-copy the whole 16-byte chunks, hand-copy the tail words
                    const size_t full = (aend - ga) & ~(size_t)15;
                    for (size_t bb = full; ga + bb < aend; bb += 4) {
                        const int32_t v = *(const int32_t *)(ga + bb);
                        asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(sdst + (unsigned)bb), "r"(v) : "memory");
                    }
                    nb = (unsigned)full;
                }
                const unsigned wb = __reduce_add_sync(0xffffffffu, nb);   // (also orders the hand-copied tail before the arrival)
                if (lane == 0) mbar_arrive_tx(s_full + slot, wb);
                if (nb) bulk_g2s(sdst, (const void *)ga, nb, s_full + slot);
            }
        }
        return;
    }

    // ---- consumers ----
    int yk[kFlatRow], xk[kFlatRow];
#pragma unroll
    for (int k = 0; k < kFlatRow; ++k) {
        yk[k] = c[k] / g.srcNi;
        xk[k] = c[k] - yk[k] * g.srcNi;
    }
    // byte offset of each entry inside a level's slot of a stage, without the window's alignment shift.
    // Absent entries (k >= ne) carry weight 0 and re-read the lane's first entry (always staged: finite x 0 adds
    // nothing); a lane with no entry at all reads the start of the stage and its result is replaced by 0.
    const bool mapped = ne > 0;
    unsigned off[kFlatRow];
    size_t gx[kFlatRow];    // element offset of the entry's window start inside a plane
#pragma unroll
    for (int k = 0; k < kFlatRow; ++k) {
        const int kk = (k < ne) ? k : 0;
        int q = 0, xm = 0, xx = 0;
#pragma unroll
        for (int m = 0; m < kFlatRow; ++m)
            if (m == kk) { q = mapped ? yk[m] - y0 : 0; xx = xk[m]; }
#pragma unroll
        for (int p = 0; p < kPlWin; ++p)
            if (q == p) xm = s_xmin[p];
        off[k] = (unsigned)q * (kPlLev * WCB) + (unsigned)(mapped ? xx - xm : 0) * ESZ;
        gx[k] = (size_t)(y0 + q) * g.srcNi + (mapped ? xm : s_xmin[0]);
    }
    // + the window's alignment shift: the same at every level and for every field (planes are a multiple of 16 bytes
    // apart and the launch takes this kernel only when every source base is 16-byte aligned)
    unsigned ad0[kFlatRow];
#pragma unroll
    for (int k = 0; k < kFlatRow; ++k) ad0[k] = off[k] + ((unsigned)(gx[k] * ESZ) & 15u);
    int slot = 0;
    unsigned par = 0, so = 0;
    const size_t dl = a.dstLev32;
    for (int f = 0; f < a.nfields; ++f) {
        const FieldDev &fd = fp.f[f];
        TOUT *__restrict__ d = (TOUT *)fd.dst + a.dstOff + t;
        const int eop = fd.epi_op;
        const TACC earg = (TACC)fd.epi_arg;
        const bool lean = fullTile && eop == 0;
        for (int l0 = 0; l0 < fd.nlev; l0 += kPlLev) {
            const int nl = fd.nlev - l0;
            mbar_wait(s_full + slot, par);
            TIN x[kPlLev][kFlatRow];
#pragma unroll
            for (int l = 0; l < kPlLev; ++l) {
#pragma unroll
                for (int k = 0; k < kFlatRow; ++k)   // (levels beyond nl read stale but in-bounds shared memory; not stored)
                    x[l][k] = *(const TIN *)(pl_smem + (ad0[k] + so) + l * WCB);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty + slot);   // this warp holds the stage in registers: refill it
            if (++slot == kPlStages) { slot = 0; so = 0; par ^= 1u; } else so += STB;
            if (lean && nl >= kPlLev) {
                // the common case, straight line: every lane of the tile mapped, a full stage, no epilogue
                TACC acc[kPlLev];
#pragma unroll
                for (int l = 0; l < kPlLev; ++l) {
                    acc[l] = 0;
#pragma unroll
                    for (int k = 0; k < kFlatRow; ++k) acc[l] += w[k] * (TACC)x[l][k];
                }
                if (store) {
#pragma unroll
                    for (int l = 0; l < kPlLev; ++l) st_stream(d + l * dl, (TOUT)acc[l]);
                }
                d += kPlLev * dl;
            } else if (store) {
                TACC acc[kPlLev];
#pragma unroll
                for (int l = 0; l < kPlLev; ++l) {
                    acc[l] = 0;
#pragma unroll
                    for (int k = 0; k < kFlatRow; ++k) acc[l] += w[k] * (TACC)x[l][k];
                    if (!mapped) acc[l] = 0;
                }
                if (eop) {
#pragma unroll
                    for (int l = 0; l < kPlLev; ++l) acc[l] = epilogue(acc[l], eop, (double)earg);
                }
#pragma unroll
                for (int l = 0; l < kPlLev; ++l) {
                    if (l < nl) st_stream(d, (TOUT)acc[l]);
                    d += dl;
                }
            } else {
                d += (size_t)kPlLev * dl;
            }
        }
    }
}

}  // namespace mprg
