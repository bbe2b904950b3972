// apply_pipe.cuh -- TMA-staged, pipelined column apply kernel (the hot kernel), v7.
//
// One CTA owns a tile of up to 32 consecutive destination points of one grid row and sweeps EVERY
// stacked 3-D field of the launch for it:
//   prologue  the tile's CSR slice and its tile schedule (distinct source columns of the tile in ascending
//             id order, their runs of consecutive ids; built once per route by k_tile_schedule) go to shared
//             memory; every lane keeps the (weight, staged-column slot) pairs of ITS target in registers for
//             the whole sweep when the row has <= 3 entries (bilinear, nearest);
//   staging   unit u + 1 (= one field x one 64-level chunk) is fetched while unit u is reduced: ONE
//             cp.async.bulk (UBLKCP, completion counted on an mbarrier) per RUN of consecutively numbered
//             columns, which file order keeps contiguous in memory -- the TMA unit accepts one request per
//             ~17 cycles per SM, so fewer, larger requests are the lever.  Columns whose byte size is a
//             multiple of 16 land packed, 16-byte aligned; others (61 levels: 244 B) are fetched as the
//             16-byte-aligned window around the run and read in place with 4-byte shared loads at stride 61
//             words (odd: conflict-free) -- no realignment pass, no second barrier;
//   math      lanes run along TARGETS: warp w takes level groups w, w + 8 (4 levels each); lane t reads 4 levels
//             of each of its row's columns with one 16-byte shared load and the 32 lanes' results for one level
//             leave as one coalesced 128-byte streaming store into [lev][j][i].  No transpose through shared
//             memory, one CTA barrier per unit;
//   rotation  units of a wind pair alternate zonal / meridional chunks of the same levels; the zonal results wait
//             in shared memory (8 KB) until the meridional unit, rotate_winds_cgrid (interp.F90:737-748) runs
//             between the two, both are stored rotated.
// The kernel is compiled per launch content (MODE): only launches that hold unaligned units / wind pairs carry
// the code and registers for them, so the main stack of aligned fields keeps its 48-register, 5-CTA/SM shape.
// (Negative result, profiles/r02: staging with per-thread cp.async (LDGSTS) instead of bulk copies -- meant for
// meshes numbered without locality -- ran at 60-68 % of the HBM peak against 71-88 % for bulk copies on every
// numbering, and was removed.)
#pragma once
#include <type_traits>

#include "common.cuh"

namespace mprg {

constexpr int kPipeThreads = 256;
constexpr int kPipeWarps = kPipeThreads / 32;
constexpr int kPipeTile = 32;
constexpr int kPipeCap = 256;      // CSR entries per tile accepted (== threads: one entry per thread in the prologue)
constexpr int kPipeLev = 64;       // levels per unit
constexpr int kPipeMaxUnits = 64;  // units (field x 64-level chunk) per launch; the host splits longer stacks
constexpr int kPipeStages = 2;     // resident CTAs beat pipeline depth (profiles/r01: 2 x 5 CTAs > 3 x 4 > 4 x 3)
constexpr int kRunPad = 40;        // unaligned units: slack per run so that 16-byte-aligned windows never overlap
// launch content the kernel is compiled for
constexpr int kModeUnal = 1;       // some unit's column chunks are not 16-byte aligned
constexpr int kModeRot = 2;        // some units are wind pairs (fused rotation)
constexpr int kModeCmp = 4;        // composed wind route: every two units are the (zonal, meridional) sources of one output

struct UnitDev {
    const void *src;
    void *dst;
    size_t srcBytes;   // size of the source array (guards the last aligned window)
    int32_t nlev;      // column stride of the field, in elements
    int32_t L0, Ln;    // level chunk [L0, L0+Ln)
    int32_t flags;     // bits 0-7: MPRG_EPI_*; kUnit* below
    double epi_arg;
};
// the unit descriptors of a launch travel as a kernel parameter (3 KB of constant bank): no
// descriptor copy sits between consecutive launches on the stream
struct UnitPack {
    UnitDev u[kPipeMaxUnits];
};
constexpr int kUnitAligned = 0x100;                  // column chunks start 16-byte aligned and are a multiple of 16 bytes long
constexpr int kUnitRotU = 0x200, kUnitRotV = 0x400;  // wind pair: this unit is the zonal / meridional chunk
constexpr int kUnitMerged = 0x800;                   // the chunk is the whole column: runs of consecutive ids are contiguous in memory
// composed wind route: A = the zonal source (partial sums parked in shared memory), B = the meridional source (adds its
// terms with the entries' second weights and stores)
constexpr int kUnitCmpA = 0x1000, kUnitCmpB = 0x2000;

// Tile record: everything the kernel needs to know about one tile, precomputed when the route is built
// (k_tile_schedule) and laid out so that ONE bulk copy brings it into shared memory -- the prologue of a tile is a
// single memory latency instead of three dependent ones (row pointers -> weights / slots; tile pointer -> columns),
// which cost 17 % of the stall samples of the main launch and 38 % of a 4-unit launch.  Fixed stride per route
// (sized by the route's tile maxima), sections at fixed offsets:
//   header   16 B   uint16 nu, cnt, nruns, ntile; uint32 flags (bit 0: every row <= 3 entries, bit 1: every row == 3
//                   entries and the tile is full); int32 reserved
//   rows     32 x 8 words of the weight type: per target  w0 w1 w2 | slot0 | slot1 << 8 | slot2 << 16 | len << 24 |
//                   run0 | run1 << 8 | run2 << 16   (rows of <= 3 entries; longer rows: len only)
//   uniq     int32 [nuCap]   distinct source columns of the tile, ascending
//   urun     uint8 [nuCap]   run (maximal sequence of consecutive ids) each of them belongs to
//   runFirst uint8 [runCap + 1]  first slot of every run; runFirst[nruns] = nu (low byte), so a run's length is a difference
//   (routes with rows of more than 3 entries only -- conservative:)
//   rowoff   uint16 [34]     entry offset of every target's row
//   eoff     uint16 [entCap] per entry: slot | run << 8
//   ew       weight [entCap] per entry
// Composed wind routes (compose.cu) replace the rows section by 32 x cmpStride bytes: per target
//   a[12] b[12] (weight type) | uint32 slots 0-3 | slots 4-7 | slots 8-11 | uint32 len
struct RecLayout {
    int32_t stride, offUniq, offUrun, offRunFirst, offRowoff, offEoff, offEw;
    int32_t cmpStride;   // bytes of one composed row; 0: ordinary route
};
__host__ __device__ inline RecLayout rec_layout(int wsize, int nuMax, int runsMax, int entMax, bool generic, bool composite = false) {
    RecLayout L;
    L.cmpStride = composite ? 2 * kCmpRow * wsize + 16 : 0;
    int o = 16 + kPipeTile * (composite ? L.cmpStride : 8 * wsize);
    if (composite) generic = false;
    L.offUniq = o; o += ((nuMax + 3) & ~3) * 4;
    L.offUrun = o; o += (nuMax + 15) & ~15;
    L.offRunFirst = o; o += (runsMax + 1 + 15) & ~15;   // + the sentinel runFirst[nruns] = nu
    L.offRowoff = L.offEoff = L.offEw = 0;
    if (generic) {
        L.offRowoff = o; o += 80;
        L.offEoff = o; o += ((entMax + 7) & ~7) * 2;
        L.offEw = o; o += ((entMax + 1) & ~1) * wsize;
    }
    L.stride = (o + 15) & ~15;
    return L;
}
constexpr unsigned kRecFast = 1u, kRecAll3 = 2u;

template <typename TW>
struct PipeArgs {
    const unsigned char *rec;  // tile records of the route for weight type TW (k_tile_schedule)
    RecLayout lay;
    int64_t nDst;
    // destination addressing: point t of level l of a field goes to dst[l * dstLev + dstOff + t].  A slab
    // buffer has dstLev = nDst, dstOff = 0; writing straight into a full-grid [lev][nj][ni] field (this
    // rank's own, or the writing rank's mapped over NVLink: the gather fused into the store) has
    // dstLev = ni * nj, dstOff = first owned point.
    int64_t dstLev, dstOff;
    uint32_t dstLev32;  // = dstLev (always < 2^31: rows are indexed with int32): level offsets are one 32 x 32 -> 64 multiply
    int32_t ni;         // destination row length (tiles never straddle rows)
    int32_t tilesPerRow;
    int32_t nunits;
    int32_t nPlain;     // the first nPlain units are plain aligned fields (phase A of the kernel)
    int32_t nRotA;      // the next nRotA units are aligned wind pairs: phase A's copy protocol, rotation in the math
    int32_t stageOff;   // byte offset of the first stage in dynamic shared memory (after the record and the unit descriptors)
    int32_t stageBytes; // bytes of one stage (host: the largest unit's need at the route's tile maxima)
    int32_t holdOff;    // byte offset of the wind-pair hold buffer (kModeRot launches)
    int32_t rotOff;     // byte offset of the tile's rotation constants (kModeRot launches): fetched with the record
    // kModeRot launches only: per-point rotation constants of this rank's destination rows, in the arithmetic type
    // of the rotation (RotMath): [nDst][4] = sina, tana, 1/cosa, 1/(cosa + sina tana)
    const void *rotc;
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

// dynamic shared memory: [mbarriers 64 B][tile record][unit descriptors][stages][hold buffer (kModeRot)]
constexpr int kPipeSmemHead = 64;
// bytes of one stage that a unit needs for a tile of nu columns in nruns runs.  Aligned units: slots packed at the
// column size so that a run is contiguous in shared memory too -- unless that size is a multiple of 128 bytes (every
// slot would start on bank 0) or the chunk is not the whole column: 16 bytes of padding and one copy per column then.
__host__ __device__ inline size_t pipe_unit_stage_bytes(bool aligned, bool merged, unsigned chunkB, int nu, int nruns) {
    if (aligned) return (size_t)nu * ((merged && (chunkB & 127u)) ? chunkB : chunkB + 16u);
    return (size_t)nu * chunkB + (size_t)kRunPad * (merged ? nruns : nu) + 32;
}

template <typename TACC>
__device__ __forceinline__ TACC pipe_epi(TACC v, int op, TACC arg) {
    return op == MPRG_EPI_ADD ? v + arg : (op == MPRG_EPI_MUL ? v * arg : v);
}

// Arithmetic type of the fused / stand-alone wind rotation: the reference rotates R8 fields in R8
// (interp.F90:737-748); an all-fp32 apply (fp32 output, fp32 accumulation) rotates in fp32 -- <= 3e-7
// relative, inside the 1e-5 contract -- because B200's fp64 / conversion throughput would otherwise make
// the rotation cost as much as regridding the two fields.  accumulate = f64 selects fp64 throughout.
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
// the same from a column that is only element-aligned in shared memory (unaligned units)
template <typename TIN, typename TACC>
__device__ __forceinline__ void fma4u(TACC (&acc)[4], TACC wt, unsigned saddr) {
    if (sizeof(TIN) == 4) {
        float x, y, z, w;
        asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(x) : "r"(saddr));
        asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(y) : "r"(saddr + 4));
        asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(z) : "r"(saddr + 8));
        asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(w) : "r"(saddr + 12));
        acc[0] += wt * (TACC)x; acc[1] += wt * (TACC)y; acc[2] += wt * (TACC)z; acc[3] += wt * (TACC)w;
    } else {
        double x, y, z, w;
        asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(x) : "r"(saddr));
        asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(y) : "r"(saddr + 8));
        asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(z) : "r"(saddr + 16));
        asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(w) : "r"(saddr + 24));
        acc[0] += wt * (TACC)x; acc[1] += wt * (TACC)y; acc[2] += wt * (TACC)z; acc[3] += wt * (TACC)w;
    }
}
template <typename T>
__device__ __forceinline__ void sts4(unsigned saddr, const T (&v)[4]) {
    if (sizeof(T) == 4) {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(saddr), "f"(*(const float *)&v[0]), "f"(*(const float *)&v[1]),
                     "f"(*(const float *)&v[2]), "f"(*(const float *)&v[3]) : "memory");
    } else {
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(saddr), "d"(*(const double *)&v[0]), "d"(*(const double *)&v[1]) : "memory");
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(saddr + 16), "d"(*(const double *)&v[2]), "d"(*(const double *)&v[3]) : "memory");
    }
}
template <typename T>
__device__ __forceinline__ void lds4(unsigned saddr, T (&v)[4]) {
    if (sizeof(T) == 4) {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(*(float *)&v[0]), "=f"(*(float *)&v[1]), "=f"(*(float *)&v[2]),
                     "=f"(*(float *)&v[3]) : "r"(saddr));
    } else {
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(*(double *)&v[0]), "=d"(*(double *)&v[1]) : "r"(saddr));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(*(double *)&v[2]), "=d"(*(double *)&v[3]) : "r"(saddr + 16));
    }
}

// MODE: launch content (kModeUnal | kModeRot).  MINB: resident CTAs per SM the kernel is compiled for
// (register cap 48 at 5, 64 at 4).
template <typename TIN, typename TOUT, typename TACC, int MODE, int MINB>
__global__ void __launch_bounds__(kPipeThreads, MINB)
k_apply_pipe(PipeArgs<TACC> a, const __grid_constant__ UnitPack up) {
    constexpr bool UNAL = (MODE & kModeUnal) != 0, ROT = (MODE & kModeRot) != 0, CMP = (MODE & kModeCmp) != 0;
    constexpr int ESZ = (int)sizeof(TIN);
    constexpr int GB = 4 * ESZ;                           // bytes of one 4-level group in a staged column
    using TR = typename RotMath<TOUT, TACC>::type;

    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long *s_mbar = (unsigned long long *)smem;        // [0..1] phase A stages, [2] the tile record, [3..4] phase B stages
    unsigned char *s_rec = smem + kPipeSmemHead;                    // the tile record (RecLayout)
    UnitDev *s_units = (UnitDev *)(s_rec + a.lay.stride);
    unsigned char *s_stage = smem + a.stageOff;
    const int32_t *s_uniq = (const int32_t *)(s_rec + a.lay.offUniq);
    const unsigned char *s_urun = s_rec + a.lay.offUrun;
    const unsigned char *s_runFirst = s_rec + a.lay.offRunFirst;
    const unsigned short *s_rowoff = (const unsigned short *)(s_rec + a.lay.offRowoff);   // (generic routes only)
    const unsigned short *s_off = (const unsigned short *)(s_rec + a.lay.offEoff);
    const TACC *s_w = (const TACC *)(s_rec + a.lay.offEw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = blockIdx.x / a.tilesPerRow;
    const int i0 = (blockIdx.x - row * a.tilesPerRow) * kPipeTile;
    const int64_t t0 = (int64_t)row * a.ni + i0;

    // ---- prologue: ONE bulk copy brings the tile's record (schedule, per-target rows, weights) ---------------
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kPipeStages; ++i) {
            mbar_init(s_mbar + i, 1);                    // phase A: one arrival per unit
            mbar_init(s_mbar + 3 + i, kPipeWarps);       // phase B: one arrival per warp per unit
        }
        mbar_init(s_mbar + 2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        // kModeRot: the tile's rotation constants ride the same barrier (one latency, instead of a global load in the
        // middle of every meridional unit: 0.13 ms of a 0.6-ms pair launch, profiles/r02)
        unsigned rotB = 0;
        if (ROT) rotB = (unsigned)min((int64_t)min(kPipeTile, a.ni - i0), a.nDst - t0) * 4u * (unsigned)sizeof(TR);
        mbar_arrive_tx(s_mbar + 2, (unsigned)a.lay.stride + rotB);
        bulk_g2s((unsigned)__cvta_generic_to_shared(s_rec), a.rec + (size_t)blockIdx.x * a.lay.stride, (unsigned)a.lay.stride, s_mbar + 2);
        if (ROT && rotB) bulk_g2s((unsigned)__cvta_generic_to_shared(smem) + (unsigned)a.rotOff, (const TR *)a.rotc + 4 * t0, rotB, s_mbar + 2);
    }
    for (int i = tid; i < a.nunits * (int)(sizeof(UnitDev) / 4); i += kPipeThreads)
        ((int32_t *)s_units)[i] = ((const int32_t *)&up)[i];
    __syncthreads();             // barriers initialised, unit descriptors in place
    mbar_wait(s_mbar + 2, 0);    // the record has landed
    const int nu = ((const unsigned short *)s_rec)[0];
    const int ntile = ((const unsigned short *)s_rec)[3];
    const unsigned rflags = ((const unsigned *)s_rec)[2];
    const bool live = lane < ntile;
    const bool fast = (rflags & kRecFast) != 0;   // every row of the tile has <= 3 entries
    // whole tile made of 3-entry rows (bilinear, fully mapped, full tile): no per-entry predicates at all
    const bool all3 = (rflags & kRecAll3) != 0;
    // this lane's row (3 weights, slots, runs, length): re-read every unit with one or two 16-byte shared loads
    // instead of living in 6-8 registers across the copy issue and the barrier
    const unsigned row0 = (unsigned)__cvta_generic_to_shared(s_rec + 16 + lane * (CMP ? a.lay.cmpStride : 8 * (int)sizeof(TACC)));

    const unsigned stage0 = (unsigned)__cvta_generic_to_shared(s_stage);
    const unsigned hold0 = (unsigned)__cvta_generic_to_shared(smem) + (unsigned)a.holdOff;
    const unsigned rot0 = (unsigned)__cvta_generic_to_shared(smem) + (unsigned)a.rotOff;

    // slot s = lane * warps + warp is copied by that lane, so the (warp-serialised) bulk-copy issue is spread
    // evenly over all warps instead of queuing behind the first two.  The list is in ascending id order and columns
    // of consecutive ids are contiguous in memory, so the owner of the first slot of a run fetches the whole run
    // with one bulk copy (brun = its length in columns, 0 for the other slots of the run).
    const int bslot = lane * kPipeWarps + warp;
    const int bcol = bslot < nu ? s_uniq[bslot] : -1;
    int brun = 0;
    if (bcol >= 0) {
        const int r = s_urun[bslot];
        if ((int)s_runFirst[r] == bslot) brun = (((int)s_runFirst[r + 1] - bslot - 1) & 0xff) + 1;   // (1..256: the sentinel is nu mod 256)
    }

    // Two phases in one launch.  Phase A: the first a.nPlain units are plain aligned fields -- the bulk of every
    // pass -- and run the leanest code (one barrier arrival per unit, 16-byte loads only).  Phase B: wind pairs and
    // unaligned columns (per-warp arrivals, in-place element loads, rotation).  One launch pays the tile prologue
    // once; the lean loop is not slowed by the code and registers the general one needs.
    const int nA = a.nPlain, nA2 = a.nPlain + (ROT ? a.nRotA : 0);
    auto issue = [&](int u) {
        if (u >= a.nunits) return;
        const UnitDev &ud = s_units[u];
        const unsigned chunkB = (unsigned)ud.Ln * ESZ;
        const unsigned sbase = stage0 + (u % kPipeStages) * a.stageBytes;
        const bool merged = (ud.flags & kUnitMerged) != 0;
        // exact column chunks (aligned units).  Whole-column units whose column size is not a multiple of 128 bytes
        // pack their slots at the column size, so a run is contiguous in shared memory too and its owner fetches it whole
        const bool packed = merged && (chunkB & 127u);
        if (u < nA2) {
            unsigned long long *bar = s_mbar + (u % kPipeStages);
            const char *g = (const char *)ud.src + ((size_t)bcol * ud.nlev + ud.L0) * ESZ;
            if (tid == 0) mbar_arrive_tx(bar, chunkB * (unsigned)nu);   // one arrival posts the unit's bytes
            if (packed) {
                if (brun > 0) bulk_g2s(sbase + bslot * chunkB, g, chunkB * (unsigned)brun, bar);
            } else if (bcol >= 0) {
                bulk_g2s(sbase + bslot * (chunkB + 16u), g, chunkB, bar);
            }
            return;
        }
        if (MODE == 0) return;
        // phase B: the barrier counts one arrival per warp (lane 0 posts the bytes of the warp's copies)
        unsigned long long *bar = s_mbar + 3 + ((u - nA2) % kPipeStages);
        unsigned nb = 0, sdst = 0;
        uintptr_t ga = 0;
        if (!UNAL || (ud.flags & kUnitAligned)) {
            if (packed ? brun > 0 : bcol >= 0) {
                nb = packed ? chunkB * (unsigned)brun : chunkB;
                sdst = sbase + bslot * (packed ? chunkB : chunkB + 16u);
                ga = (uintptr_t)ud.src + ((size_t)bcol * ud.nlev + ud.L0) * ESZ;
            }
        } else if (merged ? brun > 0 : bcol >= 0) {
            // unaligned columns: the 16-byte-aligned window around the run (or the single column chunk); it lands at
            // a 16-byte-aligned address chosen so that windows never overlap, and the math reads it where it lies.
            // (absolute addresses: the source base itself need only be element-aligned; device allocations are
            // 256-byte aligned, so the window's first 16-byte chunk always lies inside the caller's allocation)
            const int ncol = merged ? brun : 1, r = merged ? (int)s_urun[bslot] : bslot;
            const uintptr_t a0 = (uintptr_t)ud.src + ((size_t)bcol * ud.nlev + ud.L0) * ESZ;
            const uintptr_t aend = (uintptr_t)ud.src + ud.srcBytes;
            ga = a0 & ~(uintptr_t)15;
            size_t n = ((a0 + (size_t)(ncol - 1) * ud.nlev * ESZ + chunkB + 15) & ~(uintptr_t)15) - ga;
            sdst = sbase + (((unsigned)bslot * chunkB + (unsigned)(kRunPad * r) + 15u) & ~15u);
            if (ga + n > aend) {
                // last window of the array: bulk-copy the whole 16-byte chunks, hand-copy the tail words
                const size_t full = (aend - ga) & ~(size_t)15;
                for (size_t b = full; ga + b < aend; b += 4) {
                    const int32_t v = *(const int32_t *)(ga + b);
                    asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(sdst + (unsigned)b), "r"(v) : "memory");
                }
                n = full;
            }
            nb = (unsigned)n;
        }
        const unsigned wb = __reduce_add_sync(0xffffffffu, nb);   // (also orders the hand-copied tail before the arrival)
        if (lane == 0) mbar_arrive_tx(bar, wb);
        if (nb) bulk_g2s(sdst, (const void *)ga, nb, bar);
    };

    const size_t dlev = a.dstLev32;                              // elements between consecutive levels of a destination field
    const size_t grp8 = (size_t)(4 * kPipeWarps) * dlev;          // elements between a warp's consecutive level groups

    // the reduction of one unit; UN / RT: compile-time content of the phase the unit belongs to
    auto math = [&](int u, auto UN_c, auto RT_c) {
        constexpr bool UN = decltype(UN_c)::value, RT = decltype(RT_c)::value;
        const UnitDev &ud = s_units[u];
        const unsigned st = stage0 + (u % kPipeStages) * a.stageBytes;  // shared-window address of unit u's staging
        const int Ln = ud.Ln;
        const int eop = ud.flags & 0xff;
        const TACC earg = (TACC)ud.epi_arg;
        const int ngroups = (Ln + 3) >> 2;
        if constexpr (CMP) {
            // composed wind route: this lane's <= 8 entries; the A unit (zonal source) parks its partial sums, the B unit
            // (meridional source, the entries' second weights) starts from them and stores
            const bool isB = (ud.flags & kUnitCmpB) != 0;
            const unsigned wrow = row0 + (isB ? (unsigned)(kCmpRow * sizeof(TACC)) : 0u);
            TACC cw0[4], cw1[4], cw2[4];
            lds4<TACC>(wrow, cw0);
            lds4<TACC>(wrow + 4 * (unsigned)sizeof(TACC), cw1);
            lds4<TACC>(wrow + 8 * (unsigned)sizeof(TACC), cw2);
            unsigned sl[4];   // slots 0-3, 4-7, 8-11, len
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(sl[0]), "=r"(sl[1]), "=r"(sl[2]), "=r"(sl[3])
                         : "r"(row0 + (unsigned)(2 * kCmpRow * sizeof(TACC))));
            const int clen = (int)sl[3];
            const unsigned cstride = ((ud.flags & kUnitMerged) && (((unsigned)Ln * ESZ) & 127u)) ? (unsigned)Ln * ESZ : (unsigned)Ln * ESZ + 16u;
            TOUT *d = (TOUT *)ud.dst + ((size_t)(unsigned)(ud.L0 + 4 * warp) * dlev + a.dstOff + t0 + lane);
#pragma unroll
            for (int gi = 0; gi < kPipeLev / 4 / kPipeWarps; ++gi, d += grp8) {
                const int g = warp + gi * kPipeWarps;
                if (g >= ngroups) break;
                const unsigned hp = hold0 + (unsigned)((gi * kPipeThreads + tid) * 4 * (int)sizeof(TACC));
                TACC acc[4] = {0, 0, 0, 0};
                if (isB) lds4<TACC>(hp, acc);
                const unsigned lp = st + g * GB;
#pragma unroll
                for (int j = 0; j < kCmpRow; ++j)   // absent entries never touch staging
                    if (j < clen) fma4<TIN, TACC>(acc, j < 4 ? cw0[j & 3] : (j < 8 ? cw1[j & 3] : cw2[j & 3]), lp + ((sl[j >> 2] >> (8 * (j & 3))) & 0xffu) * cstride);
                if (!isB) { sts4<TACC>(hp, acc); continue; }
                if (eop != MPRG_EPI_NONE) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[k] = pipe_epi(acc[k], eop, earg);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (4 * g + k < Ln) __stcs(d + (size_t)k * dlev, (TOUT)acc[k]);
            }
            return;
        }
        // this lane's row (weights, slots, runs): one or two 16-byte shared loads per unit
        TACC rw[3];
        unsigned pk, pr = 0;
        if (sizeof(TACC) == 4) {
            float4 q;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(row0));
            rw[0] = (TACC)q.x; rw[1] = (TACC)q.y; rw[2] = (TACC)q.z; pk = __float_as_uint(q.w);
            if (UN) asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(pr) : "r"(row0 + 16));
        } else {
            double x, y, z;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(x), "=d"(y) : "r"(row0));
            asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(z) : "r"(row0 + 16));
            asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(pk) : "r"(row0 + 24));
            if (UN) asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(pr) : "r"(row0 + 32));
            rw[0] = (TACC)x; rw[1] = (TACC)y; rw[2] = (TACC)z;
        }
        const int rlen = (int)(pk >> 24);                                       // (fast paths: <= 3)
        const int rbeg = fast ? 0 : (int)s_rowoff[lane], rend = fast ? 0 : (int)s_rowoff[lane + 1];   // (generic rows only)
        const unsigned chunkB = (unsigned)Ln * ESZ;
        const bool direct = !UN || (ud.flags & kUnitAligned);     // columns sit 16-byte aligned at slot * ustride
        const unsigned ustride = ((ud.flags & kUnitMerged) && (chunkB & 127u)) ? chunkB : chunkB + 16u;
        // byte offsets of this lane's columns in an unaligned unit's staging: the column of slot s in run r lies at
        // window(r) + (its global byte offset - the window's)
        unsigned co[3] = {0, 0, 0};
        if (UN && !direct) {
            const bool merged = (ud.flags & kUnitMerged) != 0;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int sl = (int)((pk >> (8 * j)) & 0xffu), r = merged ? (int)((pr >> (8 * j)) & 0xffu) : sl;
                const int f = merged ? (int)s_runFirst[r] : sl;
                const uintptr_t a0 = (uintptr_t)ud.src + ((size_t)s_uniq[f] * ud.nlev + ud.L0) * ESZ;
                co[j] = (((unsigned)f * chunkB + (unsigned)(kRunPad * r) + 15u) & ~15u) + (unsigned)(a0 & 15) +
                        (unsigned)(sl - f) * (unsigned)(ud.nlev * ESZ);
            }
        }
        // this lane's output column: level L0 + 4 * warp of target t0 + lane; groups are 8 * 4 levels apart
        const size_t dcol = (size_t)(unsigned)(ud.L0 + 4 * warp) * dlev + a.dstOff + t0 + lane;
        TOUT *d = (TOUT *)ud.dst + dcol;
        const bool rotU = RT && (ud.flags & kUnitRotU), rotV = RT && (ud.flags & kUnitRotV);
#pragma unroll
        for (int gi = 0; gi < kPipeLev / 4 / kPipeWarps; ++gi, d += grp8) {
            const int g = warp + gi * kPipeWarps;
            if (g >= ngroups) break;
            TACC acc[4] = {0, 0, 0, 0};
            const unsigned lp = st + g * GB;
            if (direct) {
                if (all3) {             // straight line: 3 x (LDS.128 + 4 FFMA)
                    fma4<TIN, TACC>(acc, rw[0], lp + (pk & 0xffu) * ustride);
                    fma4<TIN, TACC>(acc, rw[1], lp + ((pk >> 8) & 0xffu) * ustride);
                    fma4<TIN, TACC>(acc, rw[2], lp + ((pk >> 16) & 0xffu) * ustride);
                } else if (fast) {
#pragma unroll
                    for (int j = 0; j < 3; ++j)   // absent entries never touch staging (0 x garbage = NaN)
                        if (j < rlen) fma4<TIN, TACC>(acc, rw[j], lp + ((pk >> (8 * j)) & 0xffu) * ustride);
                } else {
                    for (int k = rbeg; k < rend; ++k) fma4<TIN, TACC>(acc, s_w[k], lp + (unsigned)(s_off[k] & 0xff) * ustride);
                }
            } else if (fast) {
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    if (j < rlen) fma4u<TIN, TACC>(acc, rw[j], lp + co[j]);
            } else {
                const bool merged = (ud.flags & kUnitMerged) != 0;
                for (int k = rbeg; k < rend; ++k) {
                    const int sl = s_off[k] & 0xff, r = merged ? (s_off[k] >> 8) : sl;
                    const int f = merged ? (int)s_runFirst[r] : sl;
                    const uintptr_t a0 = (uintptr_t)ud.src + ((size_t)s_uniq[f] * ud.nlev + ud.L0) * ESZ;
                    fma4u<TIN, TACC>(acc, s_w[k], lp + (((unsigned)f * chunkB + (unsigned)(kRunPad * r) + 15u) & ~15u) +
                                                      (unsigned)(a0 & 15) + (unsigned)(sl - f) * (unsigned)(ud.nlev * ESZ));
                }
            }
            if (rotU) {          // zonal unit: park the results, rounded to the output type exactly as a store would
                TOUT h[4] = {(TOUT)acc[0], (TOUT)acc[1], (TOUT)acc[2], (TOUT)acc[3]};
                sts4<TOUT>(hold0 + (unsigned)((gi * kPipeThreads + tid) * 4 * (int)sizeof(TOUT)), h);
                continue;
            }
            if (rotV) {
                // meridional unit: u' = (u + v tana) / (cosa + sina tana); v' = (v - u' sina) / cosa  (v' uses u');
                // the two divisors are per-point constants, applied as reciprocals (same in k_rotate)
                TOUT h[4];
                lds4<TOUT>(hold0 + (unsigned)((gi * kPipeThreads + tid) * 4 * (int)sizeof(TOUT)), h);
                TR c[4];   // sina, tana, 1/cosa, 1/(cosa + sina tana) of this lane's target (staged with the record)
                lds4<TR>(rot0 + (unsigned)(lane * 4 * (int)sizeof(TR)), c);
                TOUT *du = (TOUT *)s_units[u - 1].dst + dcol + (size_t)gi * grp8;   // where the held zonal values go
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    TR uu = (TR)h[k], vv = (TR)(TOUT)acc[k];
                    uu = (uu + vv * c[1]) * c[3];
                    vv = (vv - uu * c[0]) * c[2];
                    if (4 * g + k < Ln) {
                        __stcs(du + (size_t)k * dlev, (TOUT)uu);
                        __stcs(d + (size_t)k * dlev, (TOUT)vv);
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
                TOUT *d1 = d + dlev, *d2 = d1 + dlev, *d3 = d2 + dlev;
                __stcs(d, (TOUT)acc[0]);
                __stcs(d1, (TOUT)acc[1]);
                __stcs(d2, (TOUT)acc[2]);
                __stcs(d3, (TOUT)acc[3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (4 * g + k < Ln) __stcs(d + (size_t)k * dlev, (TOUT)acc[k]);
            }
        }
    };

    issue(0);
    for (int u = 0; u < nA; ++u) {
        if (u > 0) __syncthreads();     // every warp has finished reading unit u - 1: its buffer may be refilled
        issue(u + 1);
        mbar_wait(s_mbar + (u % kPipeStages), (unsigned)((u / kPipeStages) & 1));  // unit u's bytes have landed
        if (live) math(u, std::false_type{}, std::false_type{});
    }
    if (ROT) {     // aligned wind pairs: the same single-arrival copy protocol, rotation fused into the math
        for (int u = nA; u < nA2; ++u) {
            if (u > 0) __syncthreads();
            issue(u + 1);
            mbar_wait(s_mbar + (u % kPipeStages), (unsigned)((u / kPipeStages) & 1));
            if (live) math(u, std::false_type{}, std::true_type{});
        }
    }
    if (MODE != 0) {
        for (int u = nA2; u < a.nunits; ++u) {
            if (u > 0) __syncthreads();
            issue(u + 1);
            const int k = u - nA2;
            mbar_wait(s_mbar + 3 + (k % kPipeStages), (unsigned)((k / kPipeStages) & 1));
            if (live) math(u, std::integral_constant<bool, UNAL>{}, std::integral_constant<bool, ROT>{});
        }
    }
}

// Tile schedule of a route: for every row-aligned 32-target tile the list of distinct source
// columns in ASCENDING id order, the run (maximal sequence of consecutive ids) each belongs to and, per CSR
// entry, the index of its column in that list -- written as the tile's RECORD (RecLayout).  Ascending order makes
// columns of consecutively numbered cells neighbours in the list; they are also neighbours in memory (file order),
// so the apply kernel fetches each such run with ONE bulk copy.
// FILL == false: per-tile maxima (entries, distinct columns, runs, longest row) and totals;  FILL == true: the records.
template <bool FILL, typename TW, int CAP>
__global__ void __launch_bounds__(CAP)
k_tile_schedule(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, const TW *__restrict__ w,
                const TW *__restrict__ w2 /* composed routes: the entries' second weights */, int64_t nDst,
                int32_t ni, int32_t tilesPerRow, int32_t *maxima /* entries, uniq, runs, row */,
                unsigned long long *totals /* columns, runs */, unsigned char *__restrict__ rec, RecLayout lay) {
    // CAP = threads = CSR entries a tile may hold: kPipeCap for ordinary routes, 2 kPipeCap for composed wind routes
    __shared__ int32_t s_col[CAP];   // sort keys (column ids; INT_MAX padding)
    __shared__ int32_t s_idx[CAP];   // entry index within the tile that the key came from
    __shared__ int32_t s_cnt[CAP / 32], s_rcnt[CAP / 32];
    __shared__ unsigned short s_eoff[CAP];   // per entry (original order): slot | run << 8
    __shared__ int32_t s_rp[kPipeTile + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = blockIdx.x / tilesPerRow;
    const int i0 = (blockIdx.x - row * tilesPerRow) * kPipeTile;
    const int64_t t0 = (int64_t)row * ni + i0;
    const int64_t t1 = min(t0 + min(kPipeTile, ni - i0), nDst);
    const int ntile = t0 < nDst ? (int)(t1 - t0) : 0;
    int base = 0, cnt = 0;
    if (t0 < nDst) { base = rowptr[t0]; cnt = rowptr[t1] - base; }
    if (tid <= kPipeTile) s_rp[tid] = (t0 < nDst ? rowptr[min(t0 + min(tid, ntile), nDst)] : 0) - base;
    if (!FILL && tid == 0) atomicMax(maxima, cnt);
    if (cnt > CAP) {  // tile too fat for the pipelined kernel: the route falls back to register gathers
        if (!FILL && tid == 0) atomicMax(maxima + 1, cnt);
        return;
    }
    s_col[tid] = tid < cnt ? col[base + tid] : 0x7fffffff;
    s_idx[tid] = tid;
    __syncthreads();
    // bitonic sort of the CAP (key, index) pairs
    for (int k = 2; k <= CAP; k <<= 1)
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
    // a run starts at a distinct column whose predecessor in the sorted list is not id - 1 (equal ids were skipped above)
    const bool runStart = uniq && (tid == 0 || s_col[tid - 1] != s_col[tid] - 1);
    const unsigned bal = __ballot_sync(0xffffffffu, uniq), rb = __ballot_sync(0xffffffffu, runStart);
    if (lane == 0) { s_cnt[warp] = __popc(bal); s_rcnt[warp] = __popc(rb); }
    __syncthreads();
    int incl = __popc(bal & ((2u << lane) - 1u)), nu = 0;   // unique keys up to and including this position
    int rincl = __popc(rb & ((2u << lane) - 1u)), nr = 0;   // run starts up to and including this position
#pragma unroll
    for (int wq = 0; wq < CAP / 32; ++wq) {
        const int v = s_cnt[wq], rv = s_rcnt[wq];
        if (wq < warp) { incl += v; rincl += rv; }
        nu += v; nr += rv;
    }
    int rowMax = 0;
    if (tid < ntile) rowMax = s_rp[tid + 1] - s_rp[tid];
    if (!FILL) {
        for (int o = 16; o > 0; o >>= 1) rowMax = max(rowMax, __shfl_xor_sync(0xffffffffu, rowMax, o));
        if (tid == 0) {
            atomicMax(maxima + 1, nu); atomicMax(maxima + 2, nr); atomicMax(maxima + 3, rowMax);
            atomicAdd(totals, (unsigned long long)nu); atomicAdd(totals + 1, (unsigned long long)nr);
        }
        return;
    }
    unsigned char *R = rec + (size_t)blockIdx.x * lay.stride;
    if (uniq) {
        ((int32_t *)(R + lay.offUniq))[incl - 1] = s_col[tid];
        R[lay.offUrun + incl - 1] = (unsigned char)(rincl - 1);
        if (runStart) R[lay.offRunFirst + rincl - 1] = (unsigned char)(incl - 1);
    }
    if (tid == 0) R[lay.offRunFirst + nr] = (unsigned char)nu;   // sentinel
    if (tid < cnt) s_eoff[s_idx[tid]] = (unsigned short)((incl - 1) | ((rincl - 1) << 8));
    __syncthreads();
    // per-target rows
    const bool shortRow = rowMax <= 3;
    const unsigned allShort = __ballot_sync(0xffffffffu, tid >= ntile || shortRow);
    const unsigned allThree = __ballot_sync(0xffffffffu, tid < ntile && rowMax == 3);
    if (lay.cmpStride) {
        if (tid < kPipeTile) {   // composed wind route: <= kCmpRow entries, two weights each
            TW *wa = (TW *)(R + 16 + tid * lay.cmpStride), *wb = wa + kCmpRow;
            unsigned *sl = (unsigned *)(wb + kCmpRow);
            const int len = tid < ntile ? min(rowMax, kCmpRow) : 0;
            unsigned s[3] = {0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < kCmpRow; ++j) {
                const bool h = j < len;
                wa[j] = h ? w[base + s_rp[tid] + j] : (TW)0;
                wb[j] = h ? w2[base + s_rp[tid] + j] : (TW)0;
                s[j >> 2] |= (h ? (unsigned)(s_eoff[s_rp[tid] + j] & 0xffu) : 0u) << (8 * (j & 3));
            }
            sl[0] = s[0]; sl[1] = s[1]; sl[2] = s[2]; sl[3] = (unsigned)len;
        }
    } else if (tid < kPipeTile) {
        TW *rw = (TW *)(R + 16 + tid * 8 * (int)sizeof(TW));
        unsigned pk = (unsigned)min(rowMax, 255) << 24, pr = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const bool h = tid < ntile && j < rowMax && shortRow;
            rw[j] = h ? w[base + s_rp[tid] + j] : (TW)0;
            const unsigned o = h ? (unsigned)s_eoff[s_rp[tid] + j] : 0u;
            pk |= (o & 0xffu) << (8 * j);
            pr |= (o >> 8) << (8 * j);
        }
        ((unsigned *)(rw + 3))[0] = pk;
        ((unsigned *)(rw + 3))[sizeof(TW) / 4] = pr;
    }
    if (tid == 0) {
        unsigned short *h = (unsigned short *)R;
        h[0] = (unsigned short)nu; h[1] = (unsigned short)cnt; h[2] = (unsigned short)nr; h[3] = (unsigned short)ntile;
        ((unsigned *)R)[2] = (allShort == 0xffffffffu ? kRecFast : 0u) | ((allThree == 0xffffffffu && ntile == kPipeTile) ? kRecAll3 : 0u);
        ((unsigned *)R)[3] = 0u;
    }
    if (lay.offEw) {   // routes with longer rows keep the whole CSR slice of the tile
        if (tid <= kPipeTile) ((unsigned short *)(R + lay.offRowoff))[tid] = (unsigned short)s_rp[min(tid, ntile)];
        if (tid < cnt) {
            ((unsigned short *)(R + lay.offEoff))[tid] = s_eoff[tid];
            ((TW *)(R + lay.offEw))[tid] = w[base + tid];
        }
    }
}

}  // namespace mprg
