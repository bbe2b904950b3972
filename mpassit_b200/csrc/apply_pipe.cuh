// apply_pipe.cuh -- asynchronous-copy pipelined column apply kernel (the hot kernel).
//
// One CTA owns a tile of up to 32 consecutive destination points of one grid row and
// sweeps EVERY stacked field of the launch for it:
//   prologue  the tile's CSR slice is cached in shared memory and its source columns are
//             de-duplicated (neighbouring targets share most of their sources), giving `nu`
//             unique columns and, per row entry, the shared-memory offset of its column;
//             rows with <= 3 entries (bilinear, nearest) keep weights + offsets in registers
//             for the whole field sweep;
//   pipeline  for unit u (= one field x one 64-level chunk) the unique columns stream
//             global -> shared with cp.async (16 B, L1-allocating) STAGES-1 units ahead of the
//             math: no registers are tied up by loads and HBM latency is covered by the
//             depth of the pipeline rather than by warp occupancy;
//   phase A   lanes along levels: a half-warp (16-byte path) or warp (4-byte path) reduces one
//             target's row from the staged columns, result written transposed to the out tile;
//   phase B   one coalesced 128-byte streaming store per level into [lev][j][i].
// Columns of any level count work: the generic path copies the 16-byte-aligned window that
// encloses the column and phase A adds the in-window element offset.
#pragma once
#include "common.cuh"

namespace mprg {

constexpr int kPipeThreads = 256;
constexpr int kPipeWarps = kPipeThreads / 32;
constexpr int kPipeTile = 32;
constexpr int kPipeCap = 256;   // CSR entries per tile accepted (== threads: one entry per thread in the dedup)
constexpr int kPipeLev = 64;    // levels per unit
constexpr int kPipeMaxUnits = 64;  // units (field x 64-level chunk) per launch; the host splits longer stacks

struct UnitDev {
    const void *src;
    void *dst;
    size_t srcBytes;   // size of the source array (guards the last aligned window, generic path)
    int32_t nlev;      // column stride of the field, in elements
    int32_t L0, Ln;    // level chunk [L0, L0+Ln)
    int32_t epi_op;
    double epi_arg;
};

template <typename TW>
struct PipeArgs {
    const int32_t *rowptr;
    const int32_t *col;
    const TW *w;
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
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
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
    return 64 * 4                       // s_rowptr (33 used) + misc + mbarriers
           + kPipeCap * 4               // s_col, later reused as s_off (byte offset of the entry's slot)
           + kPipeCap * 4               // s_uniq
           + kPipeCap * sizeof(TACC)    // s_w
           + kPipeLev * kPipeTile * sizeof(TOUT)                        // s_out (rotation-swizzled, no padding)
           + kPipeMaxUnits * sizeof(UnitDev);                           // s_units
}

template <typename TACC>
__device__ __forceinline__ TACC pipe_epi(TACC v, int op, TACC arg) {
    return op == MPRG_EPI_ADD ? v + arg : (op == MPRG_EPI_MUL ? v * arg : v);
}

// BULK: stage columns with the TMA bulk-copy engine (cp.async.bulk + mbarrier complete_tx), one
// copy per source column, instead of per-thread 16-byte cp.async (VEC layouts only).
template <typename TIN, typename TOUT, typename TACC, bool VEC, int STAGES, bool BULK>
__global__ void __launch_bounds__(kPipeThreads, 4)
k_apply_pipe(PipeArgs<TACC> a) {
    constexpr int SLOTB = pipe_slot_bytes<TIN>();
    constexpr int QN = kPipeLev * (int)sizeof(TIN) / 16;  // 16-byte chunks per full column: 16 (f32) / 32 (f64)
    constexpr int SPP = kPipeThreads / QN;                // slots copied per pass of the whole CTA
    constexpr int EPV = 16 / (int)sizeof(TIN);            // elements per 16-byte chunk

    extern __shared__ __align__(16) unsigned char smem[];
    int32_t *s_rowptr = (int32_t *)smem;            // [33]; s_misc at [36..43]; mbarriers at [48..55]
    int32_t *s_misc = s_rowptr + 36;
    unsigned long long *s_mbar = (unsigned long long *)(s_rowptr + 48);  // one per stage (BULK only)
    int32_t *s_col = s_rowptr + 64;                 // column ids, then per-entry slot byte offsets
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
    __syncthreads();
    const int base = s_rowptr[0];
    const int cnt = s_rowptr[ntile] - base;  // host guarantees cnt <= kPipeCap
    int c = -1;
    if (tid < cnt) {
        c = __ldg(a.col + base + tid);
        s_col[tid] = c;
        s_w[tid] = __ldg(a.w + base + tid);
    }
    __syncthreads();
    int first = tid;
    bool uniq = false;
    if (tid < cnt) {
        for (int i = 0; i < tid; ++i)
            if (s_col[i] == c) { first = i; break; }
        uniq = first == tid;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, uniq);
    if (lane == 0) s_misc[warp] = __popc(bal);
    __syncthreads();
    int slot = __popc(bal & ((1u << lane) - 1u)), nu = 0;
#pragma unroll
    for (int w = 0; w < kPipeWarps; ++w) {
        const int v = s_misc[w];
        if (w < warp) slot += v;
        nu += v;
    }
    if (uniq) { s_uniq[slot] = c; s_col[tid] = slot * SLOTB; }  // s_col[first] now holds the slot offset ...
    __syncthreads();
    if (tid < cnt && !uniq) s_col[tid] = s_col[first];
    // ... (a duplicate only ever reads the entry of its FIRST occurrence, which is unique and
    // therefore already final) -- from here on s_col[k] = byte offset of entry k's column slot
    __syncthreads();
    const int32_t *s_off = s_col;

    // rows with <= 3 entries keep their (weight, slot offset) in registers for the whole sweep
    constexpr int NT = VEC ? 2 : 4;  // targets handled by this lane's (half-)warp
    TACC rw[NT][3];
    int ro[NT][3];
    int rlen[NT], rbeg[NT], tt[NT];
    bool fast = true;
#pragma unroll
    for (int it = 0; it < NT; ++it) {
        tt[it] = VEC ? warp * 4 + it * 2 + (lane >> 4) : warp * 4 + it;
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

    // unit descriptors -> shared memory (they sit on every unit's critical path otherwise)
    for (int i = tid; i < a.nunits * (int)(sizeof(UnitDev) / 4); i += kPipeThreads)
        ((int32_t *)s_units)[i] = __ldg((const int32_t *)a.units + i);

    // copy assignment: lane q = tid % QN moves 16-byte chunk q of slots s0, s0+SPP, ... ; the first
    // KREG slots' column ids live in registers for the whole sweep
    constexpr int KREG = 4;
    const int q = tid % QN, s0 = tid / QN;
    int cs[KREG];
#pragma unroll
    for (int i = 0; i < KREG; ++i) cs[i] = s0 + i * SPP < nu ? s_uniq[s0 + i * SPP] : -1;
    const unsigned stage0 = (unsigned)__cvta_generic_to_shared(s_stage) + s0 * SLOTB + q * 16;
    if (BULK && tid == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; ++i) mbar_init(s_mbar + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    // BULK: slot s = lane * warps + warp is copied by that lane, so the (warp-serialised) bulk-copy
    // issue is spread evenly over all warps instead of queuing behind the first two
    const int bslot = lane * kPipeWarps + warp;
    const int bcol = (BULK && bslot < nu) ? s_uniq[bslot] : -1;
    const unsigned bstage = (unsigned)__cvta_generic_to_shared(s_stage) + bslot * SLOTB;
    __syncthreads();  // s_units (and the mbarriers) visible

    auto issue = [&](int u) {
        if (u < a.nunits) {
            const UnitDev &ud = s_units[u];
            const unsigned st = stage0 + (u % STAGES) * stageBytes;
            if (BULK) {
                const unsigned colB = (unsigned)ud.Ln * (unsigned)sizeof(TIN);   // multiple of 16 on VEC layouts
                if (tid == 0) mbar_expect_tx(s_mbar + (u % STAGES), colB * (unsigned)nu);
                if (bcol >= 0)
                    bulk_g2s(bstage + (u % STAGES) * stageBytes,
                             (const char *)ud.src + ((size_t)bcol * ud.nlev + ud.L0) * sizeof(TIN), colB, s_mbar + (u % STAGES));
            } else if (VEC) {
                // columns are 16-byte aligned multiples of 16 bytes: chunk q of column s, no guards
                if (q * EPV < ud.Ln) {
                    const char *g0 = (const char *)ud.src + ((size_t)ud.L0 * sizeof(TIN) + (size_t)q * 16);
                    const size_t colBytes = (size_t)ud.nlev * sizeof(TIN);
#pragma unroll
                    for (int i = 0; i < KREG; ++i)
                        if (cs[i] >= 0)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(st + i * SPP * SLOTB),
                                         "l"(g0 + (size_t)cs[i] * colBytes) : "memory");
                    for (int s = s0 + KREG * SPP; s < nu; s += SPP)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(st + (s - s0) * SLOTB),
                                     "l"(g0 + (size_t)s_uniq[s] * colBytes) : "memory");
                }
            } else {
                const char *src = (const char *)ud.src;
                for (int s = s0; s < nu; s += SPP) {
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

    // out tile [lev][32], column rotated by f(lev) so that both the transposed writes of phase A
    // and the row reads of phase B are bank-conflict free:  VEC f = 2*(lev/4),  generic f = lev
    const int l16 = lane & 15;
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
        const int eop = ud.epi_op;
        const TACC earg = (TACC)ud.epi_arg;
        // ---- phase A ---------------------------------------------------------
        if (VEC) {
            const bool act = 4 * l16 < Ln;
            const unsigned char *lp = st + l16 * 16 * (int)(sizeof(TIN) / 4);  // 4 levels = 16 B (f32) / 32 B (f64)
#pragma unroll
            for (int it = 0; it < NT; ++it) {
                if (tt[it] < ntile && act) {
                    TACC acc[4] = {0, 0, 0, 0};
                    if (fast) {
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            if (!all3 && j >= rlen[it]) continue;  // never touch staging for absent entries (0 x garbage = NaN)
                            const unsigned char *p = lp + ro[it][j];
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
                            const unsigned char *p = lp + s_off[k];
                            if (sizeof(TIN) == 4) {
                                const float4 x = *(const float4 *)p;
                                acc[0] += wt * (TACC)x.x; acc[1] += wt * (TACC)x.y; acc[2] += wt * (TACC)x.z; acc[3] += wt * (TACC)x.w;
                            } else {
                                const double2 x = *(const double2 *)p, y = *((const double2 *)p + 1);
                                acc[0] += wt * (TACC)x.x; acc[1] += wt * (TACC)x.y; acc[2] += wt * (TACC)y.x; acc[3] += wt * (TACC)y.y;
                            }
                        }
                    }
                    TOUT *o = s_out + (4 * l16) * kPipeTile + ((tt[it] + 2 * l16) & 31);
                    if (eop != MPRG_EPI_NONE) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) acc[k] = pipe_epi(acc[k], eop, earg);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) o[k * kPipeTile] = (TOUT)acc[k];
                }
            }
        } else {
            const bool a0 = lane < Ln, a1 = lane + 32 < Ln;
#pragma unroll
            for (int it = 0; it < NT; ++it) {
                if (tt[it] < ntile) {
                    TACC acc0 = 0, acc1 = 0;
                    const int nk = rlen[it];
                    for (int j = 0; j < nk; ++j) {
                        const TACC wt = fast ? rw[it][j] : s_w[rbeg[it] + j];
                        const int so = fast ? ro[it][j] : s_off[rbeg[it] + j];
                        // in-window element offset of this column's first wanted level
                        const int eo = (int)(((size_t)s_uniq[so / SLOTB] * ud.nlev + ud.L0) & (size_t)(EPV - 1));
                        const TIN *p = (const TIN *)(st + so) + eo + lane;
                        if (a0) acc0 += wt * (TACC)p[0];
                        if (a1) acc1 += wt * (TACC)p[32];
                    }
                    const int cc = (tt[it] + lane) & 31;  // rows lane and lane+32 rotate by the same amount mod 32
                    if (a0) s_out[lane * kPipeTile + cc] = (TOUT)pipe_epi(acc0, eop, earg);
                    if (a1) s_out[(lane + 32) * kPipeTile + cc] = (TOUT)pipe_epi(acc1, eop, earg);
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
                const int rot = VEC ? 2 * (lev >> 2) : lev;
                v[k] = s_out[lev * kPipeTile + ((lane + rot) & 31)];  // rows >= Ln hold stale data, never stored
            }
#pragma unroll
            for (int k = 0; k < NB; ++k)
                if (warp + k * kPipeWarps < Ln) __stcs(dst + k * step, v[k]);
        }
    }
    if (!BULK) cp_async_wait<0>();
}

// per-route tile statistics: largest CSR slice and largest number of distinct columns of any
// row-aligned 32-target tile (decides whether / with how many stages the pipelined kernel runs)
__global__ void __launch_bounds__(kPipeThreads)
k_tile_stats(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, int64_t nDst, int32_t ni,
             int32_t tilesPerRow, int32_t *maxEntries, int32_t *maxUniq) {
    __shared__ int32_t s_col[kPipeCap];
    __shared__ int32_t s_cnt[kPipeWarps];
    const int tid = threadIdx.x;
    const int row = blockIdx.x / tilesPerRow;
    const int i0 = (blockIdx.x - row * tilesPerRow) * kPipeTile;
    const int64_t t0 = (int64_t)row * ni + i0;
    const int64_t t1 = min(t0 + min(kPipeTile, ni - i0), nDst);
    if (t0 >= nDst) return;
    const int base = rowptr[t0], cnt = rowptr[t1] - base;
    if (tid == 0) atomicMax(maxEntries, cnt);
    if (cnt > kPipeCap) { if (tid == 0) atomicMax(maxUniq, cnt); return; }
    if (tid < cnt) s_col[tid] = col[base + tid];
    __syncthreads();
    bool uniq = false;
    if (tid < cnt) {
        uniq = true;
        const int c = s_col[tid];
        for (int i = 0; i < tid; ++i)
            if (s_col[i] == c) { uniq = false; break; }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, uniq);
    if ((tid & 31) == 0) s_cnt[tid >> 5] = __popc(bal);
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int w = 0; w < kPipeWarps; ++w) n += s_cnt[w];
        atomicMax(maxUniq, n);
    }
}

}  // namespace mprg
