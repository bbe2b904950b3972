// apply_pipe.cuh -- TMA-staged, pipelined column apply kernel (the hot kernel).
//
// One CTA owns a tile of up to 32 consecutive destination points of one grid row and
// sweeps EVERY stacked 3-D field of the launch for it:
//   prologue  the tile's CSR slice is cached in shared memory and its source columns are
//             de-duplicated (neighbouring targets share most of their sources), giving `nu`
//             unique columns and, per row entry, the shared-memory offset of its column;
//             rows with <= 3 entries (bilinear, nearest) keep weights + offsets in registers
//             for the whole field sweep;
//   pipeline  for unit u (= one field x one 64-level chunk) every unique column is fetched
//             by ONE bulk asynchronous copy (cp.async.bulk -> UBLKCP, completion counted on an
//             mbarrier) STAGES-1 units ahead of the math: no registers or LSU wavefronts are
//             spent on the gather and HBM latency is covered by the depth of the pipeline;
//             the copy moves the 16-byte-aligned window enclosing the column chunk, so any
//             level count works (60, 55, 61, ...);
//   phase A   lanes along levels: each half-warp reduces one target's row from the staged
//             columns (16-byte shared loads when the unit's columns are 16-byte aligned,
//             4-byte loads plus the in-window offset otherwise) and writes the result
//             transposed into a rotation-swizzled, conflict-free out tile;
//   phase B   one coalesced 128-byte streaming store per level into [lev][j][i].
// A per-thread cp.async (LDGSTS) staging path is kept behind MPASSIT_GPU_FILL=ldgsts for
// A/B measurements.
#pragma once
#include "common.cuh"

namespace mprg {

constexpr int kPipeThreads = 256;
constexpr int kPipeWarps = kPipeThreads / 32;
constexpr int kPipeTile = 32;
constexpr int kPipeCap = 256;      // CSR entries per tile accepted (== threads: one entry per thread in the dedup)
constexpr int kPipeLev = 64;       // levels per unit
constexpr int kPipeMaxUnits = 64;  // units (field x 64-level chunk) per launch; the host splits longer stacks

struct UnitDev {
    const void *src;
    void *dst;
    size_t srcBytes;   // size of the source array (guards the last aligned window)
    int32_t nlev;      // column stride of the field, in elements
    int32_t L0, Ln;    // level chunk [L0, L0+Ln)
    int32_t epi_op;    // bits 0-7: MPRG_EPI_*;  bit 8: columns are 16-byte aligned (vector phase A)
    double epi_arg;
};
constexpr int kUnitAligned = 0x100;

template <typename TW>
struct PipeArgs {
    const int32_t *rowptr;
    const int32_t *col;
    const TW *w;
    // tile schedule built once per route (k_tile_schedule): unique source columns of every tile
    // and, per CSR entry, the index of its column in that list
    const int32_t *tileUPtr;        // [nTiles + 1]
    const int32_t *tileUCols;       // [tileUPtr[nTiles]]
    const unsigned char *entrySlot; // [nnz]
    int64_t nDst;
    int32_t ni;         // destination row length (tiles never straddle rows)
    int32_t tilesPerRow;
    const UnitDev *units;
    int32_t nunits;
    int32_t maxU;       // slot capacity of one stage (>= max unique columns of any tile)
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
            (unsigned)__cvta_generic_to_shared(bar)),
        "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned smemDst, const void *gmem, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smemDst),
                 "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

template <typename TIN>
__host__ __device__ constexpr int pipe_slot_bytes() { return kPipeLev * (int)sizeof(TIN) + 16; }

// fixed part of the dynamic shared memory (bytes); STAGES * maxU * slotBytes of staging follow
template <typename TOUT, typename TACC>
__host__ __device__ constexpr size_t pipe_fixed_bytes() {
    return 64 * 4                                   // s_rowptr (33 used) + misc + mbarriers
           + kPipeCap * 4                           // s_col, later s_off (slot byte offset | column id mod EPV)
           + kPipeCap * 4                           // s_uniq
           + kPipeCap * sizeof(TACC)                // s_w
           + kPipeLev * kPipeTile * sizeof(TOUT)    // s_out (rotation-swizzled, no padding)
           + kPipeMaxUnits * sizeof(UnitDev);       // s_units
}

template <typename TACC>
__device__ __forceinline__ TACC pipe_epi(TACC v, int op, TACC arg) {
    return op == MPRG_EPI_ADD ? v + arg : (op == MPRG_EPI_MUL ? v * arg : v);
}

// ALLVEC: every unit of the launch has 16-byte-aligned columns (compile-time specialisation
// without the aligned-window arithmetic and without the 4-byte phase A).
template <typename TIN, typename TOUT, typename TACC, int STAGES, bool BULK, bool ALLVEC>
__global__ void __launch_bounds__(kPipeThreads, 4)
k_apply_pipe(PipeArgs<TACC> a) {
    constexpr int SLOTB = pipe_slot_bytes<TIN>();
    constexpr int QN = kPipeLev * (int)sizeof(TIN) / 16;  // 16-byte chunks per full column: 16 (f32) / 32 (f64)
    constexpr int SPP = kPipeThreads / QN;                // slots copied per pass of the whole CTA (LDGSTS path)
    constexpr int EPV = 16 / (int)sizeof(TIN);            // elements per 16-byte chunk

    extern __shared__ __align__(16) unsigned char smem[];
    int32_t *s_rowptr = (int32_t *)smem;            // [33]; s_misc at [36..43]; mbarriers at [48..55]
    int32_t *s_misc = s_rowptr + 36;
    unsigned long long *s_mbar = (unsigned long long *)(s_rowptr + 48);  // one per stage (BULK)
    int32_t *s_col = s_rowptr + 64;                 // column ids, then per-entry (slot byte offset | c mod EPV)
    int32_t *s_uniq = s_col + kPipeCap;
    TACC *s_w = (TACC *)(s_uniq + kPipeCap);
    TOUT *s_out = (TOUT *)(s_w + kPipeCap);         // [64][32], rotation-swizzled
    UnitDev *s_units = (UnitDev *)(s_out + kPipeLev * kPipeTile);
    unsigned char *s_stage = smem + pipe_fixed_bytes<TOUT, TACC>();

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = blockIdx.x / a.tilesPerRow;
    const int i0 = (blockIdx.x - row * a.tilesPerRow) * kPipeTile;
    const int64_t t0 = (int64_t)row * a.ni + i0;
    const int ntile = (int)min((int64_t)min(kPipeTile, a.ni - i0), a.nDst - t0);

    // ---- prologue: CSR slice + de-duplication of the tile's source columns --------
    if (tid <= kPipeTile) s_rowptr[tid] = a.rowptr[min(t0 + min(tid, ntile), a.nDst)];
    for (int i = tid; i < a.nunits * (int)(sizeof(UnitDev) / 4); i += kPipeThreads)
        ((int32_t *)s_units)[i] = __ldg((const int32_t *)a.units + i);
    if (BULK && tid == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; ++i) mbar_init(s_mbar + i, ALLVEC ? 1 : kPipeWarps);  // arrivals per unit
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const int base = s_rowptr[0];
    const int cnt = s_rowptr[ntile] - base;  // host guarantees cnt <= kPipeCap
    const int ub = __ldg(a.tileUPtr + blockIdx.x);
    const int nu = __ldg(a.tileUPtr + blockIdx.x + 1) - ub;
    if (tid < nu) s_uniq[tid] = __ldg(a.tileUCols + ub + tid);
    int eslot = 0;
    if (tid < cnt) {
        s_w[tid] = __ldg(a.w + base + tid);
        eslot = __ldg(a.entrySlot + base + tid);
    }
    __syncthreads();
    // SLOTB is a multiple of 16, so the low 4 bits of the offset are free for (column id mod EPV),
    // which phase A needs to find a non-16-byte-aligned column inside its staged window
    if (tid < cnt) s_col[tid] = eslot * SLOTB | (s_uniq[eslot] & (EPV - 1));
    __syncthreads();
    const int32_t *s_off = s_col;

    // this lane's two targets (half-warp per target); rows with <= 3 entries stay in registers
    const int l16 = lane & 15;
    TACC rw[2][3];
    int ro[2][3];
    int rlen[2], rbeg[2], tt[2];
    bool fast = true;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        tt[it] = warp * 4 + it * 2 + (lane >> 4);
        rbeg[it] = 0;
        rlen[it] = 0;
        if (tt[it] < ntile) { rbeg[it] = s_rowptr[tt[it]] - base; rlen[it] = s_rowptr[tt[it] + 1] - s_rowptr[tt[it]]; }
        fast = fast && rlen[it] <= 3;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const bool h = j < rlen[it] && rlen[it] <= 3;
            rw[it][j] = h ? s_w[rbeg[it] + j] : (TACC)0;
            ro[it][j] = h ? s_off[rbeg[it] + j] : 0;
        }
    }
    fast = __all_sync(0xffffffffu, fast);
    // whole tile made of 3-entry rows (bilinear, fully mapped, full tile): no per-entry predicates at all
    const bool all3 = __syncthreads_and(ntile == kPipeTile && cnt == 3 * kPipeTile &&
                                        (tid >= kPipeTile || s_rowptr[tid + 1] - s_rowptr[tid] == 3));

    const int stageBytes = a.maxU * SLOTB;
    // BULK: slot s = lane * warps + warp is copied by that lane, so the (warp-serialised) bulk-copy
    // issue is spread evenly over all warps instead of queuing behind the first two
    const int bslot = lane * kPipeWarps + warp;
    const int bcol = (BULK && bslot < nu) ? s_uniq[bslot] : -1;
    const unsigned bstage = (unsigned)__cvta_generic_to_shared(s_stage) + bslot * SLOTB;

    auto issue = [&](int u) {
        if (u < a.nunits) {
            const UnitDev &ud = s_units[u];
            if (BULK) {
                if (ALLVEC) {
                    // equal, exact column chunks: thread 0 posts the unit's byte count, slot owners copy
                    unsigned long long *bar = s_mbar + (u % STAGES);
                    const unsigned colB = (unsigned)ud.Ln * (unsigned)sizeof(TIN);
                    if (tid == 0) mbar_arrive_tx(bar, colB * (unsigned)nu);
                    if (bcol >= 0)
                        bulk_g2s(bstage + (u % STAGES) * stageBytes,
                                 (const char *)ud.src + ((size_t)bcol * ud.nlev + ud.L0) * sizeof(TIN), colB, bar);
                } else {
                // lane 0 of every warp arrives with the warp's byte count; the owner of a slot then
                // launches one bulk copy of the 16-byte-aligned window enclosing the column chunk
                unsigned long long *bar = s_mbar + (u % STAGES);
                unsigned nb = 0;
                size_t al = 0;
                if (bcol >= 0) {
                    const size_t byte0 = ((size_t)bcol * ud.nlev + ud.L0) * sizeof(TIN);
                    al = byte0 & ~(size_t)15;
                    size_t n = ((byte0 + (size_t)ud.Ln * sizeof(TIN) + 15) & ~(size_t)15) - al;
                    if (al + n > ud.srcBytes) {
                        // last window of the allocation: bulk-copy the whole chunks, hand-copy the tail words
                        const size_t full = (ud.srcBytes - al) & ~(size_t)15;
                        for (size_t b = full; al + b < ud.srcBytes; b += 4)
                            *(int32_t *)(s_stage + (size_t)(u % STAGES) * stageBytes + bslot * SLOTB + b) =
                                *(const int32_t *)((const char *)ud.src + al + b);
                        n = full;
                    }
                    nb = (unsigned)n;
                }
                const unsigned wb = __reduce_add_sync(0xffffffffu, nb);
                if (lane == 0) mbar_arrive_tx(bar, wb);
                __syncwarp();  // the expected byte count is posted before any of this warp's copies can complete
                if (nb) bulk_g2s(bstage + (u % STAGES) * stageBytes, (const char *)ud.src + al, nb, bar);
                }
            } else {
                const char *src = (const char *)ud.src;
                const int q = tid % QN;
                for (int s = tid / QN; s < nu; s += SPP) {
                    const size_t byte0 = ((size_t)s_uniq[s] * ud.nlev + ud.L0) * sizeof(TIN);
                    const size_t al = byte0 & ~(size_t)15;
                    const size_t end = byte0 + (size_t)ud.Ln * sizeof(TIN);
                    unsigned char *dstc = s_stage + (size_t)(u % STAGES) * stageBytes + s * SLOTB;
                    // the enclosing aligned window has at most QN + 1 chunks: lane q takes chunk q, lane 0 also chunk QN
                    for (int qq = q; qq <= QN; qq += QN) {
                        const size_t g = al + (size_t)qq * 16;
                        if (g < end && (qq < QN || q == 0)) {
                            if (g + 16 <= ud.srcBytes) {
                                cp_async16(dstc + qq * 16, src + g);
                            } else {  // last window of the allocation: copy only what exists
                                for (int b = 0; b < 16 && g + b < ud.srcBytes; b += 4) cp_async4(dstc + qq * 16 + b, src + g + b);
                            }
                        }
                    }
                }
            }
        }
        if (!BULK) cp_async_commit();  // always commit (possibly empty) so group counting stays uniform
    };

#pragma unroll
    for (int u = 0; u < STAGES - 1; ++u) issue(u);

    // out tile [lev][32]; the column of target t in row lev is (t + rot(lev)) & 31 so that both the
    // transposed writes of phase A and the row reads of phase B are bank-conflict free:
    //   aligned units   lane l16 holds levels 4*l16..4*l16+3  -> rot = 2*(lev/4)
    //   unaligned units lane l16 holds levels l16 + 16k       -> rot = 2*(lev%16)
    for (int u = 0; u < a.nunits; ++u) {
        issue(u + STAGES - 1);          // refills the buffer read in unit u-1 (reads done: A/B barrier of u-1)
        if (BULK) {
            __syncthreads();            // orders phase B(u-1) before phase A(u) on s_out
            mbar_wait(s_mbar + (u % STAGES), (unsigned)((u / STAGES) & 1));  // unit u's bytes have landed
        } else {
            cp_async_wait<STAGES - 1>();    // this thread's copies of unit u have landed
            __syncthreads();                // ... and everyone's; also orders phase B(u-1) before phase A(u) on s_out
        }
        const UnitDev &ud = s_units[u];
        const unsigned char *st = s_stage + (u % STAGES) * stageBytes;
        const int Ln = ud.Ln;
        const int eop = ud.epi_op & 0xff;
        const bool aligned = ALLVEC || (ud.epi_op & kUnitAligned) != 0;
        const TACC earg = (TACC)ud.epi_arg;
        // ---- phase A ---------------------------------------------------------
        if (aligned) {
            const bool act = 4 * l16 < Ln;
            const unsigned char *lp = st + l16 * 16 * (int)(sizeof(TIN) / 4);  // 4 levels = 16 B (f32) / 32 B (f64)
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                if (tt[it] < ntile && act) {
                    TACC acc[4] = {0, 0, 0, 0};
                    if (fast) {
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            if (!all3 && j >= rlen[it]) continue;  // never touch staging for absent entries (0 x garbage = NaN)
                            const unsigned char *p = lp + (ro[it][j] & ~15);
                            if (sizeof(TIN) == 4) {
                                const float4 x = *(const float4 *)p;
                                acc[0] += rw[it][j] * (TACC)x.x; acc[1] += rw[it][j] * (TACC)x.y;
                                acc[2] += rw[it][j] * (TACC)x.z; acc[3] += rw[it][j] * (TACC)x.w;
                            } else {
                                const double2 x = *(const double2 *)p, y = *((const double2 *)p + 1);
                                acc[0] += rw[it][j] * (TACC)x.x; acc[1] += rw[it][j] * (TACC)x.y;
                                acc[2] += rw[it][j] * (TACC)y.x; acc[3] += rw[it][j] * (TACC)y.y;
                            }
                        }
                    } else {
                        for (int k = rbeg[it]; k < rbeg[it] + rlen[it]; ++k) {
                            const TACC wt = s_w[k];
                            const unsigned char *p = lp + (s_off[k] & ~15);
                            if (sizeof(TIN) == 4) {
                                const float4 x = *(const float4 *)p;
                                acc[0] += wt * (TACC)x.x; acc[1] += wt * (TACC)x.y; acc[2] += wt * (TACC)x.z; acc[3] += wt * (TACC)x.w;
                            } else {
                                const double2 x = *(const double2 *)p, y = *((const double2 *)p + 1);
                                acc[0] += wt * (TACC)x.x; acc[1] += wt * (TACC)x.y; acc[2] += wt * (TACC)y.x; acc[3] += wt * (TACC)y.y;
                            }
                        }
                    }
                    if (eop != MPRG_EPI_NONE) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) acc[k] = pipe_epi(acc[k], eop, earg);
                    }
                    TOUT *o = s_out + (4 * l16) * kPipeTile + ((tt[it] + 2 * l16) & 31);
#pragma unroll
                    for (int k = 0; k < 4; ++k) o[k * kPipeTile] = (TOUT)acc[k];
                }
            }
        } else {
            // element offset of a column's first wanted level inside its staged window:
            //   (c * nlev + L0) mod EPV = ((c mod EPV) * (nlev mod EPV) + L0 mod EPV) mod EPV
            const int nm = ud.nlev & (EPV - 1), lm = ud.L0 & (EPV - 1);
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                if (tt[it] < ntile) {
                    TACC acc[4] = {0, 0, 0, 0};
                    const int nk = rlen[it];
                    for (int j = 0; j < nk; ++j) {
                        const TACC wt = fast ? rw[it][j] : s_w[rbeg[it] + j];
                        const int so = fast ? ro[it][j] : s_off[rbeg[it] + j];
                        const int eo = ((so & 15) * nm + lm) & (EPV - 1);
                        const TIN *p = (const TIN *)(st + (so & ~15)) + eo + l16;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (l16 + 16 * k < Ln) acc[k] += wt * (TACC)p[16 * k];
                    }
                    const int cc = (tt[it] + 2 * l16) & 31;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (l16 + 16 * k < Ln) s_out[(l16 + 16 * k) * kPipeTile + cc] = (TOUT)pipe_epi(acc[k], eop, earg);
                }
            }
        }
        __syncthreads();
        // ---- phase B: transposed, coalesced streaming store ---------------------
        if (lane < ntile) {
            TOUT *dst = (TOUT *)ud.dst + ((size_t)(ud.L0 + warp) * a.nDst + t0 + lane);
            const size_t step = (size_t)kPipeWarps * a.nDst;
            constexpr int NB = kPipeLev / kPipeWarps;  // levels per warp
            TOUT v[NB];
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                const int lev = warp + k * kPipeWarps;
                const int rot = aligned ? 2 * (lev >> 2) : 2 * (lev & 15);
                v[k] = s_out[lev * kPipeTile + ((lane + rot) & 31)];  // rows >= Ln hold stale data, never stored
            }
#pragma unroll
            for (int k = 0; k < NB; ++k)
                if (warp + k * kPipeWarps < Ln) __stcs(dst + k * step, v[k]);
        }
    }
    if (!BULK) cp_async_wait<0>();
}

// Tile schedule of a route: for every row-aligned 32-target tile the list of distinct source
// columns (first-occurrence order) and, per CSR entry, the index of its column in that list.
// FILL == false: per-tile unique counts (+ global maxima);  FILL == true: write the lists.
template <bool FILL>
__global__ void __launch_bounds__(kPipeThreads)
k_tile_schedule(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, int64_t nDst, int32_t ni,
                int32_t tilesPerRow, int32_t *maxEntries, int32_t *maxUniq, int32_t *tileCount,
                const int32_t *__restrict__ tileUPtr, int32_t *__restrict__ tileUCols,
                unsigned char *__restrict__ entrySlot) {
    __shared__ int32_t s_col[kPipeCap];
    __shared__ int32_t s_slot[kPipeCap];
    __shared__ int32_t s_cnt[kPipeWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = blockIdx.x / tilesPerRow;
    const int i0 = (blockIdx.x - row * tilesPerRow) * kPipeTile;
    const int64_t t0 = (int64_t)row * ni + i0;
    const int64_t t1 = min(t0 + min(kPipeTile, ni - i0), nDst);
    int base = 0, cnt = 0;
    if (t0 < nDst) { base = rowptr[t0]; cnt = rowptr[t1] - base; }
    if (!FILL && tid == 0) atomicMax(maxEntries, cnt);
    if (cnt > kPipeCap) {  // tile too fat for the pipelined kernel: the route falls back to register gathers
        if (!FILL && tid == 0) { atomicMax(maxUniq, cnt); tileCount[blockIdx.x] = 0; }
        return;
    }
    int c = -1;
    if (tid < cnt) { c = col[base + tid]; s_col[tid] = c; }
    __syncthreads();
    int first = tid;
    bool uniq = false;
    if (tid < cnt) {
        for (int i = 0; i < tid; ++i)
            if (s_col[i] == c) { first = i; break; }
        uniq = first == tid;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, uniq);
    if (lane == 0) s_cnt[warp] = __popc(bal);
    __syncthreads();
    int slot = __popc(bal & ((1u << lane) - 1u)), nu = 0;
#pragma unroll
    for (int w = 0; w < kPipeWarps; ++w) {
        const int v = s_cnt[w];
        if (w < warp) slot += v;
        nu += v;
    }
    if (!FILL) {
        if (tid == 0) { atomicMax(maxUniq, nu); tileCount[blockIdx.x] = nu; }
        return;
    }
    if (uniq) { s_slot[tid] = slot; tileUCols[tileUPtr[blockIdx.x] + slot] = c; }
    __syncthreads();
    if (tid < cnt) entrySlot[base + tid] = (unsigned char)s_slot[first];
}

}  // namespace mprg
