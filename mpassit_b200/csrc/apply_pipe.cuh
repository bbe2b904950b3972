// apply_pipe.cuh -- TMA-staged, pipelined column apply kernel (the hot kernel).
//
// One CTA owns a tile of up to 32 consecutive destination points of one grid row and
// sweeps EVERY stacked 3-D field of the launch for it:
//   prologue  the tile's CSR slice and its tile schedule (distinct source columns of the tile,
//             built once per route by k_tile_schedule) go to shared memory; every lane keeps the
//             (weight, staged-column offset) pairs of ITS target in registers for the whole sweep
//             when the row has <= 3 entries (bilinear, nearest);
//   pipeline  for unit u (= one field x one 64-level chunk) every distinct column -- every RUN of
//             consecutively numbered columns, which file order keeps contiguous -- is fetched by
//             ONE bulk asynchronous copy (cp.async.bulk -> UBLKCP, completion counted on an
//             mbarrier) STAGES-1 units ahead of the math: no registers or LSU wavefronts are
//             spent on the gather and HBM latency is covered by the depth of the pipeline; the
//             copy moves the 16-byte-aligned window enclosing the column chunk, so any level
//             count works (60, 55, 61, ...);
//   math      lanes run along TARGETS: warp w takes level groups w, w+8 (4 levels each); lane t
//             reads 4 levels of each of its row's columns with one 16-byte shared load (units whose
//             columns are not 16-byte aligned are first shifted into place inside their slots), and
//             the 32 lanes' results for one level leave as one coalesced 128-byte streaming store
//             into [lev][j][i].  No transpose through shared memory and one CTA barrier per unit.
#pragma once
#include "common.cuh"

namespace mprg {

constexpr int kPipeThreads = 256;
constexpr int kPipeWarps = kPipeThreads / 32;
constexpr int kPipeTile = 32;
constexpr int kPipeCap = 256;      // CSR entries per tile accepted (== threads: one entry per thread in the prologue)
constexpr int kPipeLev = 64;       // levels per unit
constexpr int kPipeMaxUnits = 64;  // units (field x 64-level chunk) per launch; the host splits longer stacks

struct UnitDev {
    const void *src;
    void *dst;
    size_t srcBytes;   // size of the source array (guards the last aligned window)
    int32_t nlev;      // column stride of the field, in elements
    int32_t L0, Ln;    // level chunk [L0, L0+Ln)
    int32_t epi_op;    // bits 0-7: MPRG_EPI_*;  bit 8: columns are 16-byte aligned (vector loads)
    double epi_arg;
};
// the unit descriptors of a launch travel as a kernel parameter (3 KB of constant bank): no
// descriptor copy sits between consecutive launches on the stream
struct UnitPack {
    UnitDev u[kPipeMaxUnits];
};
constexpr int kUnitAligned = 0x100;
constexpr int kUnitRotU = 0x200, kUnitRotV = 0x400;  // wind pair: this unit is the zonal / meridional chunk

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
    // destination addressing: point t of level l of a field goes to dst[l * dstLev + dstOff + t].  A slab
    // buffer has dstLev = nDst, dstOff = 0; writing straight into a full-grid [lev][nj][ni] field (this
    // rank's own, or the writing rank's mapped over NVLink: the gather fused into the store) has
    // dstLev = ni * nj, dstOff = first owned point.
    int64_t dstLev, dstOff;
    int32_t ni;         // destination row length (tiles never straddle rows)
    int32_t tilesPerRow;
    int32_t nunits;
    int32_t maxU;       // slot capacity of one stage (>= max unique columns of any tile)
    // ROT launches only: rotation angles of this rank's destination rows (rotate_winds_cgrid fused into the store)
    const double *rotc;  // [nDst][4]: sina, tana, 1/cosa, 1/(cosa + sina tana)
};

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
template <typename TACC>
__host__ __device__ constexpr size_t pipe_fixed_bytes() {
    return 64 * 4                                   // s_rowptr (33 used) + mbarriers
           + kPipeCap * 4                           // s_off (slot index of every entry's column)
           + kPipeCap * 4                           // s_uniq
           + kPipeCap * sizeof(TACC)                // s_w
           + kPipeMaxUnits * sizeof(UnitDev);       // s_units
}

template <typename TACC>
__device__ __forceinline__ TACC pipe_epi(TACC v, int op, TACC arg) {
    return op == MPRG_EPI_ADD ? v + arg : (op == MPRG_EPI_MUL ? v * arg : v);
}

// Arithmetic type of the fused / stand-alone wind rotation: the reference rotates R8 fields in R8
// (interp.F90:737-748); an all-fp32 apply (fp32 output, fp32 accumulation) rotates in fp32 -- <= 3e-7
// relative, inside the 1e-5 contract -- because B200's fp64 / conversion throughput would otherwise make
// the rotation cost as much as regridding the two fields.  MPASSIT_GPU_ACC=f64 selects fp64 throughout.
template <typename TOUT, typename TACC> struct RotMath { using type = double; };
template <> struct RotMath<float, float> { using type = float; };

// 4 consecutive levels of one staged column, addressed in the shared window (explicit ld.shared:
// no generic->shared conversion in the inner loop)
template <typename TIN, typename TACC>
__device__ __forceinline__ void fma4(TACC (&acc)[4], TACC wt, unsigned saddr) {
    if (sizeof(TIN) == 4) {
        float x, y, z, w;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(saddr));
        acc[0] += wt * (TACC)x; acc[1] += wt * (TACC)y; acc[2] += wt * (TACC)z; acc[3] += wt * (TACC)w;
    } else {
        double x, y, z, w;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(x), "=d"(y) : "r"(saddr));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(z), "=d"(w) : "r"(saddr + 16));
        acc[0] += wt * (TACC)x; acc[1] += wt * (TACC)y; acc[2] += wt * (TACC)z; acc[3] += wt * (TACC)w;
    }
}
template <typename TIN>
__device__ __forceinline__ void sts1(unsigned saddr, TIN v) {
    if (sizeof(TIN) == 4) asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(saddr), "f"(*(float *)&v) : "memory");
    else asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(saddr), "d"(*(double *)&v) : "memory");
}
template <typename TIN>
__device__ __forceinline__ TIN lds1(unsigned saddr) {
    TIN v;
    if (sizeof(TIN) == 4) asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(*(float *)&v) : "r"(saddr));
    else asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(*(double *)&v) : "r"(saddr));
    return v;
}

// ALLVEC: every unit of the launch has 16-byte-aligned columns (compile-time specialisation
// without the aligned-window arithmetic and without the 4-byte load path).
// MINB: resident CTAs per SM the kernel is compiled for (register cap 64 at 4, 48 at 5).
// ROT: some units are (zonal wind, meridional wind) chunk pairs (kUnitRotU then kUnitRotV, same levels);
//      the thread that reduces u(t, lev) also reduces v(t, lev) one unit later, so rotate_winds_cgrid
//      (interp.F90:737-748) runs in registers between the two and both are stored rotated -- no separate
//      pass over the fields.
template <typename TIN, typename TOUT, typename TACC, int STAGES, bool ALLVEC, int MINB, bool ROT = false>
__global__ void __launch_bounds__(kPipeThreads, MINB)
k_apply_pipe(PipeArgs<TACC> a, const __grid_constant__ UnitPack up) {
    constexpr int SLOTB = pipe_slot_bytes<TIN>();
    constexpr int EPV = 16 / (int)sizeof(TIN);            // elements per 16-byte chunk
    constexpr int GB = 4 * (int)sizeof(TIN);              // bytes of one 4-level group in a staged column

    extern __shared__ __align__(16) unsigned char smem[];
    int32_t *s_rowptr = (int32_t *)smem;            // [33]; mbarriers at [48..55]
    unsigned long long *s_mbar = (unsigned long long *)(s_rowptr + 48);  // one per stage
    int32_t *s_off = s_rowptr + 64;                 // per entry: slot index of its column in the tile's list
    int32_t *s_uniq = s_off + kPipeCap;
    TACC *s_w = (TACC *)(s_uniq + kPipeCap);
    UnitDev *s_units = (UnitDev *)(s_w + kPipeCap);
    unsigned char *s_stage = smem + pipe_fixed_bytes<TACC>();

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = blockIdx.x / a.tilesPerRow;
    const int i0 = (blockIdx.x - row * a.tilesPerRow) * kPipeTile;
    const int64_t t0 = (int64_t)row * a.ni + i0;
    const int ntile = (int)min((int64_t)min(kPipeTile, a.ni - i0), a.nDst - t0);

    // ---- prologue: CSR slice + the tile's schedule ---------------------------------
    if (tid <= kPipeTile) s_rowptr[tid] = a.rowptr[min(t0 + min(tid, ntile), a.nDst)];
    for (int i = tid; i < a.nunits * (int)(sizeof(UnitDev) / 4); i += kPipeThreads)
        ((int32_t *)s_units)[i] = ((const int32_t *)&up)[i];
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; ++i) mbar_init(s_mbar + i, ALLVEC ? 1 : kPipeWarps);  // arrivals per unit
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    const int ub = __ldg(a.tileUPtr + blockIdx.x);
    const int nu = __ldg(a.tileUPtr + blockIdx.x + 1) - ub;
    if (tid < nu) s_uniq[tid] = __ldg(a.tileUCols + ub + tid);
    __syncthreads();
    const int base = s_rowptr[0];
    const int cnt = s_rowptr[ntile] - base;  // host guarantees cnt <= kPipeCap
    if (tid < cnt) {
        s_w[tid] = __ldg(a.w + base + tid);
        s_off[tid] = __ldg(a.entrySlot + base + tid);
    }
    __syncthreads();

    // this lane's target; rows with <= 3 entries stay in registers
    const bool live = lane < ntile;
    int rbeg = 0, rlen = 0;
    if (live) { rbeg = s_rowptr[lane] - base; rlen = s_rowptr[lane + 1] - s_rowptr[lane]; }
    TACC rw[3];
    int ro[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const bool h = j < rlen && rlen <= 3;
        rw[j] = h ? s_w[rbeg + j] : (TACC)0;
        ro[j] = h ? s_off[rbeg + j] : 0;
    }
    // wind rotation: angles of this lane's target (same operation order as k_rotate)
    // (fp32 arithmetic when the whole apply is fp32 -- RotMath -- : the fp64 pipe would otherwise bound the launch)
    using TR = typename RotMath<TOUT, TACC>::type;
    TR rsa = 0, rtana = 0, rcai = 1, rdeni = 1;
    if (ROT && live) {  // per-point constants prepared once by mprg_set_rotation (k_rot_consts)
        const double2 c0 = __ldg((const double2 *)(a.rotc + 4 * (t0 + lane)));
        const double2 c1 = __ldg((const double2 *)(a.rotc + 4 * (t0 + lane)) + 1);
        rsa = (TR)c0.x; rtana = (TR)c0.y; rcai = (TR)c1.x; rdeni = (TR)c1.y;
    }
    TOUT hold[kPipeLev / 4 / kPipeWarps][4];  // ROT: the zonal unit's results, held until the meridional unit
    const bool fast = __all_sync(0xffffffffu, rlen <= 3);
    // whole tile made of 3-entry rows (bilinear, fully mapped, full tile): no per-entry predicates at all
    const bool all3 = __all_sync(0xffffffffu, live && rlen == 3);

    const int stageBytes = a.maxU * SLOTB;
    // slot s = lane * warps + warp is copied by that lane, so the (warp-serialised) bulk-copy
    // issue is spread evenly over all warps instead of queuing behind the first two
    const int bslot = lane * kPipeWarps + warp;
    const int bcol = bslot < nu ? s_uniq[bslot] : -1;
    // ALLVEC: the list is in ascending id order and columns of consecutive ids are contiguous in memory, so
    // the owner of the first slot of a run fetches the whole run with one bulk copy (brun = its length in
    // columns, 0 for the other slots of the run).  The TMA unit accepts a request every ~19 cycles per SM,
    // which is what bounds this kernel: fewer, larger requests are the lever.
    int brun = bcol >= 0 ? 1 : 0;
    if (ALLVEC && bcol >= 0) {
        if (bslot > 0 && s_uniq[bslot - 1] == bcol - 1) {
            brun = 0;
        } else {
            while (bslot + brun < nu && s_uniq[bslot + brun] == bcol + brun) ++brun;
        }
    }
    const unsigned stage0 = (unsigned)__cvta_generic_to_shared(s_stage);
    const unsigned bstage = stage0 + bslot * SLOTB;

    auto issue = [&](int u) {
        if (u >= a.nunits) return;
        const UnitDev &ud = s_units[u];
        unsigned long long *bar = s_mbar + (u % STAGES);
        if (ALLVEC) {
            // equal, exact column chunks: thread 0 posts the unit's byte count, run owners copy.  Slots are
            // packed at the column size so that a run is contiguous in shared memory too -- unless that size
            // is a multiple of 128 bytes (every slot would start on bank 0): then slots keep 16 bytes of
            // padding and every column is its own copy.  Whole-field units only: a run of columns is
            // contiguous in memory only when the chunk is the entire column.
            const unsigned colB = (unsigned)ud.Ln * (unsigned)sizeof(TIN);
            const bool packed = (colB & 127u) != 0 && ud.Ln == ud.nlev;
            const unsigned stride = packed ? colB : colB + 16;
            if (tid == 0) mbar_arrive_tx(bar, colB * (unsigned)nu);
            const char *g = (const char *)ud.src + ((size_t)bcol * ud.nlev + ud.L0) * sizeof(TIN);
            const unsigned sdst = stage0 + (u % STAGES) * stageBytes + bslot * stride;
            if (packed) {
                if (brun > 0) bulk_g2s(sdst, g, colB * (unsigned)brun, bar);
            } else if (bcol >= 0) {
                bulk_g2s(sdst, g, colB, bar);
            }
        } else if (ud.epi_op & kUnitAligned) {
            // aligned unit of a mixed launch: exact chunks again, but this barrier counts one arrival per
            // warp (lane 0 posts the bytes of the warp's own copies; an mbarrier's transaction count may
            // run ahead of the expectation, so no ordering between the two is needed)
            const unsigned colB = (unsigned)ud.Ln * (unsigned)sizeof(TIN);
            const unsigned mine = __popc(__ballot_sync(0xffffffffu, bcol >= 0));
            if (lane == 0) mbar_arrive_tx(bar, colB * mine);
            if (bcol >= 0)
                bulk_g2s(bstage + (u % STAGES) * stageBytes,
                         (const char *)ud.src + ((size_t)bcol * ud.nlev + ud.L0) * sizeof(TIN), colB, bar);
        } else {
            // lane 0 of every warp arrives with the warp's byte count; the owner of a slot then
            // launches one bulk copy of the 16-byte-aligned window enclosing the column chunk
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
            __syncwarp();  // hand-copied tail words are ordered before the arrival that publishes them
            if (lane == 0) mbar_arrive_tx(bar, wb);
            if (nb) bulk_g2s(bstage + (u % STAGES) * stageBytes, (const char *)ud.src + al, nb, bar);
        }
    };

#pragma unroll
    for (int u = 0; u < STAGES - 1; ++u) issue(u);

    const size_t grp8 = (size_t)(4 * kPipeWarps) * (size_t)a.dstLev;  // elements between a warp's consecutive level groups
    for (int u = 0; u < a.nunits; ++u) {
        if (u > 0) __syncthreads();     // every warp has finished reading unit u-1: its buffer may be refilled
        issue(u + STAGES - 1);
        mbar_wait(s_mbar + (u % STAGES), (unsigned)((u / STAGES) & 1));  // unit u's bytes have landed
        const UnitDev &ud = s_units[u];
        const unsigned st = stage0 + (u % STAGES) * stageBytes;  // shared-window address of unit u's staging
        const int Ln = ud.Ln;
        const int eop = ud.epi_op & 0xff;
        const bool aligned = ALLVEC || (ud.epi_op & kUnitAligned) != 0;
        const TACC earg = (TACC)ud.epi_arg;
        const int ngroups = (Ln + 3) >> 2;
        if (!ALLVEC && !aligned) {
            // Columns whose start is not 16-byte aligned were copied as the aligned window around them, so
            // the wanted levels begin eo elements into the slot, (c * nlev + L0) mod EPV =
            // ((c mod EPV) * (nlev mod EPV) + L0 mod EPV) mod EPV.  Each warp shifts its share of the
            // slots down by eo (conflict-free consecutive words) so that the math below reads every unit
            // with 16-byte loads; per-lane 4-byte reads at eo would conflict 4-5 way on the banks.
            const int nm = ud.nlev & (EPV - 1), lm = ud.L0 & (EPV - 1);
            for (int s = warp; s < nu; s += kPipeWarps) {
                const int eo = ((s_uniq[s] & (EPV - 1)) * nm + lm) & (EPV - 1);
                if (eo) {
                    const unsigned sb = st + s * SLOTB;
                    const TIN x0 = lds1<TIN>(sb + (eo + lane) * (int)sizeof(TIN));
                    const TIN x1 = lds1<TIN>(sb + (eo + lane + 32) * (int)sizeof(TIN));  // the window holds EPV-1 elements of slack
                    __syncwarp();
                    sts1<TIN>(sb + lane * (int)sizeof(TIN), x0);
                    sts1<TIN>(sb + (lane + 32) * (int)sizeof(TIN), x1);
                }
            }
            __syncthreads();
        }
        if (!live) continue;
        // byte offsets of this lane's columns in the unit's staging (slot index x the unit's slot stride)
        const int colBu = Ln * (int)sizeof(TIN);
        const int ustride = !ALLVEC ? SLOTB : (((colBu & 127) != 0 && Ln == ud.nlev) ? colBu : colBu + 16);
        // this lane's output column: level L0 + 4 * warp of target t0 + lane; groups are 8 * 4 levels apart
        const size_t dcol = (size_t)(ud.L0 + 4 * warp) * a.dstLev + a.dstOff + t0 + lane;
        TOUT *d = (TOUT *)ud.dst + dcol;
        const bool rotU = ROT && (ud.epi_op & kUnitRotU), rotV = ROT && (ud.epi_op & kUnitRotV);
        TOUT *du = (TOUT *)s_units[rotV ? u - 1 : u].dst + dcol;        // ROT: where the held zonal values go
#pragma unroll
        for (int gi = 0; gi < kPipeLev / 4 / kPipeWarps; ++gi, d += grp8, du += grp8) {
            const int g = warp + gi * kPipeWarps;
            if (g >= ngroups) break;
            TACC acc[4] = {0, 0, 0, 0};
            const unsigned lp = st + g * GB;
            if (all3) {             // straight line: 3 x (LDS.128 + 4 FFMA)
                fma4<TIN, TACC>(acc, rw[0], lp + ro[0] * ustride);
                fma4<TIN, TACC>(acc, rw[1], lp + ro[1] * ustride);
                fma4<TIN, TACC>(acc, rw[2], lp + ro[2] * ustride);
            } else if (fast) {
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    if (j < rlen) fma4<TIN, TACC>(acc, rw[j], lp + ro[j] * ustride);  // absent entries never touch staging (0 x garbage = NaN)
            } else {
                for (int k = rbeg; k < rbeg + rlen; ++k) fma4<TIN, TACC>(acc, s_w[k], lp + s_off[k] * ustride);
            }
            if (ROT && rotU) {          // zonal unit: keep, rounded to the output type exactly as a store would
#pragma unroll
                for (int k = 0; k < 4; ++k) hold[gi][k] = (TOUT)acc[k];
                continue;
            }
            if (ROT && rotV) {
                // meridional unit: u' = (u + v tana) / (cosa + sina tana); v' = (v - u' sina) / cosa  (v' uses u');
                // the two divisors are per-point constants, applied as reciprocals (same in k_rotate)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    TR uu = (TR)hold[gi][k], vv = (TR)(TOUT)acc[k];
                    uu = (uu + vv * rtana) * rdeni;
                    vv = (vv - uu * rsa) * rcai;
                    if (4 * g + k < Ln) {
                        __stcs(du + (size_t)k * a.dstLev, (TOUT)uu);
                        __stcs(d + (size_t)k * a.dstLev, (TOUT)vv);
                    }
                }
                continue;
            }
            if (eop != MPRG_EPI_NONE) {
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k] = pipe_epi(acc[k], eop, earg);
            }
            // one coalesced 128-byte (fp32) streaming store per level
            if (4 * g + 3 < Ln) {
                TOUT *d1 = d + a.dstLev, *d2 = d1 + a.dstLev, *d3 = d2 + a.dstLev;
                __stcs(d, (TOUT)acc[0]);
                __stcs(d1, (TOUT)acc[1]);
                __stcs(d2, (TOUT)acc[2]);
                __stcs(d3, (TOUT)acc[3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (4 * g + k < Ln) __stcs(d + (size_t)k * a.dstLev, (TOUT)acc[k]);
            }
        }
    }
}

// Tile schedule of a route: for every row-aligned 32-target tile the list of distinct source
// columns in ASCENDING id order and, per CSR entry, the index of its column in that list.  Ascending
// order makes columns of consecutively numbered cells neighbours in the list; they are also
// neighbours in memory (file order), so the apply kernel fetches each such run with ONE bulk copy.
// FILL == false: per-tile unique counts (+ global maxima);  FILL == true: write the lists.
template <bool FILL>
__global__ void __launch_bounds__(kPipeThreads)
k_tile_schedule(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, int64_t nDst, int32_t ni,
                int32_t tilesPerRow, int32_t *maxEntries, int32_t *maxUniq, int32_t *tileCount,
                const int32_t *__restrict__ tileUPtr, int32_t *__restrict__ tileUCols,
                unsigned char *__restrict__ entrySlot, unsigned long long *runsTotal) {
    __shared__ int32_t s_col[kPipeCap];   // sort keys (column ids; INT_MAX padding)
    __shared__ int32_t s_idx[kPipeCap];   // entry index within the tile that the key came from
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
    s_col[tid] = tid < cnt ? col[base + tid] : 0x7fffffff;
    s_idx[tid] = tid;
    __syncthreads();
    // bitonic sort of the kPipeCap (key, index) pairs
    for (int k = 2; k <= kPipeCap; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int p = tid ^ j;
            if (p > tid) {
                const bool up = (tid & k) == 0;
                const int a = s_col[tid], b = s_col[p];
                if ((a > b) == up) {
                    s_col[tid] = b; s_col[p] = a;
                    const int ia = s_idx[tid]; s_idx[tid] = s_idx[p]; s_idx[p] = ia;
                }
            }
            __syncthreads();
        }
    const bool uniq = tid < cnt && (tid == 0 || s_col[tid] != s_col[tid - 1]);
    const unsigned bal = __ballot_sync(0xffffffffu, uniq);
    if (lane == 0) s_cnt[warp] = __popc(bal);
    __syncthreads();
    int incl = __popc(bal & ((2u << lane) - 1u)), nu = 0;  // unique keys up to and including this position
#pragma unroll
    for (int w = 0; w < kPipeWarps; ++w) {
        const int v = s_cnt[w];
        if (w < warp) incl += v;
        nu += v;
    }
    if (!FILL) {
        if (tid == 0) { atomicMax(maxUniq, nu); tileCount[blockIdx.x] = nu; }
        return;
    }
    if (uniq) tileUCols[tileUPtr[blockIdx.x] + incl - 1] = s_col[tid];
    if (tid < cnt) entrySlot[base + s_idx[tid]] = (unsigned char)(incl - 1);
    // statistics: runs of consecutive ids (= bulk copies per unit of an all-aligned launch)
    const bool runStart = uniq && (tid == 0 || s_col[tid - 1] != s_col[tid] - 1);
    const unsigned rb = __ballot_sync(0xffffffffu, runStart);
    if (lane == 0 && rb && runsTotal) atomicAdd(runsTotal, (unsigned long long)__popc(rb));
}

}  // namespace mprg
