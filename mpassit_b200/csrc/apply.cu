// apply.cu -- batched sparse-weight application ("FieldBundleRegrid"):
//     dst[f][lev][t] = sum_k W[t][k] * src[f][col[t][k]][lev]      (0 for empty rows)
// for every stacked field f and level lev, over this rank's destination slab.
// Replaces ESMF_FieldRegrid / ESMF_FieldBundleRegrid, interp.F90:134,219,236,251,
// 268,286,307,325,344,363,382,404,431,443.
//
// Memory-bound irregular SpMM (no tensor cores: < 1 flop/byte).  Three kernels:
//   k_apply_cols    3-D fields, source level-fastest [cell][nlev] (MPAS file order):
//                   lanes run along levels so every gathered source column is one
//                   contiguous 128..256-byte request; the finished 32-target x
//                   64-level tile is transposed through shared memory so each level
//                   row leaves as one coalesced 128-byte store into [lev][j][i].
//   k_apply_flat    2-D fields and short columns (nlev <= 8: soil): lanes run along targets, the row's
//                   (col, w) stay in registers across the stacked fields.
//   k_apply_planes  source is itself a [lev][j][i] grid field (centre -> edge
//                   staggering, interp.F90:298,316): lanes along targets per level.
// Grid order is (tiles fastest, field-chunks slowest) so one chunk's source
// columns are swept once while they are L2-resident; outputs use streaming
// stores so they do not evict them.
#include <algorithm>

#include "common.cuh"

namespace mprg {

constexpr int kTile = 32;        // targets per CTA tile
constexpr int kLevChunk = 64;    // levels per pass through shared memory
constexpr int kCsrCap = 512;     // CSR entries cached per tile (falls back to global beyond)
constexpr int kFieldsPerCta = 4; // stacked fields handled by one CTA for one tile
constexpr int kThreads = 256;

struct FieldDev {
    const void *src;
    void *dst;
    int32_t nlev;
    int32_t epi_op;
    double epi_arg;
};

template <typename TW>
struct ApplyArgs {
    const int32_t *rowptr;
    const int32_t *col;
    const TW *w;
    int64_t nDst;
    int64_t dstLev, dstOff;  // dst[l * dstLev + dstOff + t]: slab buffer (nDst, 0) or full-grid field (see PipeArgs)
    uint32_t dstLev32;       // = dstLev (< 2^31): level offsets are one 32 x 32 -> 64 multiply
    int32_t uniformRow;      // > 0: every row has exactly this many entries (row t starts at t * uniformRow): no row-pointer load
    int32_t nfields;
    int64_t srcPlane;  // k_apply_planes only
};

// field descriptors of a launch, passed as a kernel parameter (constant bank) instead of through a
// device-side copy that would sit between launches on the stream
struct DstLayout {
    int64_t lev, off;  // dst[l * lev + off + t]
};
constexpr int kPackFields = 96;
struct FieldPack {
    FieldDev f[kPackFields];
};

template <typename T>
__device__ __forceinline__ void st_stream(T *p, T v) { __stcs(p, v); }

constexpr int kRowChunk = 3;  // row entries gathered per batch of loads (bilinear rows are exactly 3)

// four consecutive levels of one source column
template <typename T> struct Vec4;
template <> struct Vec4<float> {
    float x, y, z, w;
    __device__ __forceinline__ void zero() { x = y = z = w = 0.f; }
    __device__ __forceinline__ void load(const float *p) {
        const float4 v = __ldg((const float4 *)p);
        x = v.x; y = v.y; z = v.z; w = v.w;
    }
};
template <> struct Vec4<double> {
    double x, y, z, w;
    __device__ __forceinline__ void zero() { x = y = z = w = 0.0; }
    __device__ __forceinline__ void load(const double *p) {
        const double2 a = __ldg((const double2 *)p), b = __ldg((const double2 *)p + 1);
        x = a.x; y = a.y; z = b.x; w = b.y;
    }
};

template <typename TACC>
__device__ __forceinline__ TACC epilogue(TACC v, int op, double arg) {
    if (op == MPRG_EPI_ADD) return v + (TACC)arg;
    if (op == MPRG_EPI_MUL) return v * (TACC)arg;
    return v;
}

}  // namespace mprg
#include "apply_pipe.cuh"
namespace mprg {

// ---------------------------------------------------------------------------
// 3-D fields, register-gather variant (fallback when a tile does not fit the pipeline)
// ---------------------------------------------------------------------------
template <typename TIN, typename TOUT, typename TACC, bool VEC, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
k_apply_cols(ApplyArgs<TACC> a, const __grid_constant__ FieldPack fp) {
    __shared__ int32_t s_rowptr[kTile + 1];
    __shared__ int32_t s_col[kCsrCap];
    __shared__ TACC s_w[kCsrCap];
    __shared__ TOUT s_out[2][kLevChunk][kTile + 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t t0 = (int64_t)blockIdx.x * kTile;
    const int ntile = (int)min((int64_t)kTile, a.nDst - t0);

    if (tid <= kTile) s_rowptr[tid] = a.rowptr[min(t0 + tid, a.nDst)];
    __syncthreads();
    const int base = s_rowptr[0];
    const int cnt = s_rowptr[ntile] - base;
    const bool cached = cnt <= kCsrCap;
    if (cached) {
        for (int k = tid; k < cnt; k += kThreads) {
            s_col[k] = __ldg(a.col + base + k);
            s_w[k] = __ldg(a.w + base + k);
        }
    }
    __syncthreads();

    const int f0 = blockIdx.y * kFieldsPerCta;
    const int f1 = min(f0 + kFieldsPerCta, a.nfields);
    int buf = 0;
    for (int f = f0; f < f1; ++f) {
        const FieldDev fd = fp.f[f];
        const TIN *__restrict__ src = (const TIN *)fd.src;
        TOUT *__restrict__ dst = (TOUT *)fd.dst;
        const int nlev = fd.nlev;
        for (int L0 = 0; L0 < nlev; L0 += kLevChunk, buf ^= 1) {
            const int Ln = min(kLevChunk, nlev - L0);
            // ---- phase A: gather + reduce, lanes along levels ----------------
            // Every load of a row chunk (kRowChunk entries x 2 targets) is issued before the
            // first FMA that consumes one, so each lane keeps 6 (vector) / 12 (scalar) requests
            // in flight: the kernel is latency-bound otherwise.
            if (VEC) {
                // half-warp per target, 4 consecutive levels per lane (16/32-byte loads)
                const int l16 = lane & 15, hw = lane >> 4;
                const bool act = 4 * l16 < Ln;
                const TIN *sbase = src + L0 + 4 * l16;
                int tt[2], rb[2], re[2];
#pragma unroll
                for (int it = 0; it < 2; ++it) {
                    tt[it] = warp * 4 + it * 2 + hw;
                    rb[it] = re[it] = 0;
                    if (tt[it] < ntile && act) { rb[it] = s_rowptr[tt[it]] - base; re[it] = s_rowptr[tt[it] + 1] - base; }
                }
                TACC acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
                const int nk = max(re[0] - rb[0], re[1] - rb[1]);
                for (int k0 = 0; k0 < nk; k0 += kRowChunk) {
                    Vec4<TIN> v[2][kRowChunk];
                    TACC wt[2][kRowChunk];
#pragma unroll
                    for (int it = 0; it < 2; ++it)
#pragma unroll
                        for (int j = 0; j < kRowChunk; ++j) {
                            const int k = rb[it] + k0 + j;
                            wt[it][j] = 0;
                            v[it][j].zero();
                            if (k < re[it]) {
                                const int c = cached ? s_col[k] : __ldg(a.col + base + k);
                                wt[it][j] = cached ? s_w[k] : __ldg(a.w + base + k);
                                v[it][j].load(sbase + (size_t)c * nlev);
                            }
                        }
#pragma unroll
                    for (int it = 0; it < 2; ++it)
#pragma unroll
                        for (int j = 0; j < kRowChunk; ++j) {
                            acc[it][0] += wt[it][j] * (TACC)v[it][j].x; acc[it][1] += wt[it][j] * (TACC)v[it][j].y;
                            acc[it][2] += wt[it][j] * (TACC)v[it][j].z; acc[it][3] += wt[it][j] * (TACC)v[it][j].w;
                        }
                }
#pragma unroll
                for (int it = 0; it < 2; ++it)
                    if (tt[it] < ntile && act) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            s_out[buf][4 * l16 + q][tt[it]] = (TOUT)epilogue(acc[it][q], fd.epi_op, fd.epi_arg);
                    }
            } else {
                // warp per target, levels lane and lane+32; two targets per step
                const bool a0 = lane < Ln, a1 = lane + 32 < Ln;
                const TIN *sbase = src + L0 + lane;
#pragma unroll
                for (int pr = 0; pr < 2; ++pr) {
                    int tt[2], rb[2], re[2];
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        tt[it] = warp * 4 + pr * 2 + it;
                        rb[it] = re[it] = 0;
                        if (tt[it] < ntile) { rb[it] = s_rowptr[tt[it]] - base; re[it] = s_rowptr[tt[it] + 1] - base; }
                    }
                    TACC acc[2][2] = {{0, 0}, {0, 0}};
                    const int nk = max(re[0] - rb[0], re[1] - rb[1]);
                    for (int k0 = 0; k0 < nk; k0 += kRowChunk) {
                        TIN x[2][kRowChunk][2];
                        TACC wt[2][kRowChunk];
#pragma unroll
                        for (int it = 0; it < 2; ++it)
#pragma unroll
                            for (int j = 0; j < kRowChunk; ++j) {
                                const int k = rb[it] + k0 + j;
                                wt[it][j] = 0;
                                x[it][j][0] = x[it][j][1] = 0;
                                if (k < re[it]) {
                                    const int c = cached ? s_col[k] : __ldg(a.col + base + k);
                                    wt[it][j] = cached ? s_w[k] : __ldg(a.w + base + k);
                                    const TIN *p = sbase + (size_t)c * nlev;
                                    if (a0) x[it][j][0] = __ldg(p);
                                    if (a1) x[it][j][1] = __ldg(p + 32);
                                }
                            }
#pragma unroll
                        for (int it = 0; it < 2; ++it)
#pragma unroll
                            for (int j = 0; j < kRowChunk; ++j) {
                                acc[it][0] += wt[it][j] * (TACC)x[it][j][0];
                                acc[it][1] += wt[it][j] * (TACC)x[it][j][1];
                            }
                    }
#pragma unroll
                    for (int it = 0; it < 2; ++it)
                        if (tt[it] < ntile) {
                            if (a0) s_out[buf][lane][tt[it]] = (TOUT)epilogue(acc[it][0], fd.epi_op, fd.epi_arg);
                            if (a1) s_out[buf][lane + 32][tt[it]] = (TOUT)epilogue(acc[it][1], fd.epi_op, fd.epi_arg);
                        }
                }
            }
            __syncthreads();
            // ---- phase B: transposed, coalesced streaming store --------------
            if (lane < ntile) {
                for (int lev = warp; lev < Ln; lev += kThreads / 32)
                    st_stream(dst + (size_t)(L0 + lev) * a.dstLev + a.dstOff + t0 + lane, s_out[buf][lev][lane]);
            }
            // s_out is double buffered: the next pass writes the other buffer, and the
            // barrier of that pass orders this pass's reads before the buffer is reused.
        }
    }
}

// ---------------------------------------------------------------------------
// 2-D fields (nlev == 1): thread per target, row kept in registers
// ---------------------------------------------------------------------------
constexpr int kFlatRow = 4;  // row entries held in registers; longer rows stream from global

constexpr int kShortLev = 8;  // fields with at most this many levels (2-D fields, soil) take the flat kernel

// (latency-bound: 6 resident CTAs per SM instead of the 4 that 64 registers allowed -- ncu: 72-77 % of the stall samples
// were long-scoreboard at 46 % occupancy)
// ROW: row entries held in registers (3 when no row of the route is longer: bilinear, nearest; else kFlatRow);
// BATCH: 2-D fields whose gathers are all issued before the first FMA (ROW x BATCH loads in flight per thread)
template <typename TIN, typename TOUT, typename TACC, int ROW, int BATCH>
__global__ void __launch_bounds__(256, sizeof(TACC) == 4 ? 6 : 4)
k_apply_flat(ApplyArgs<TACC> a, const __grid_constant__ FieldPack fp) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nDst) return;
    // (a fully mapped bilinear / nearest route: the row pointer is t * row length -- one dependent load less in a
    // latency-bound kernel)
    const int b = a.uniformRow ? (int)t * a.uniformRow : __ldg(a.rowptr + t);
    const int e = a.uniformRow ? b + a.uniformRow : __ldg(a.rowptr + t + 1);
    int c[kFlatRow];
    TACC w[kFlatRow];
#pragma unroll
    for (int k = 0; k < kFlatRow; ++k) {
        const bool h = k < ROW && b + k < e;
        c[k] = h ? __ldg(a.col + b + k) : 0;
        w[k] = h ? __ldg(a.w + b + k) : (TACC)0;
    }
    int f = 0;
    if (e - b <= ROW) {
        // 2-D fields, short rows (bilinear, nearest): BATCH fields' gathers in flight per thread
        while (f < a.nfields) {
            const int nb = min(BATCH, a.nfields - f);     // (the last batch may be partial)
            bool flat2d = true;
#pragma unroll
            for (int q = 0; q < BATCH; ++q) flat2d = flat2d && (q >= nb || fp.f[f + q].nlev == 1);
            if (!flat2d) break;
            TIN x[BATCH][ROW];
#pragma unroll
            for (int q = 0; q < BATCH; ++q) {
                const TIN *__restrict__ src = (const TIN *)fp.f[q < nb ? f + q : f].src;
#pragma unroll
                for (int k = 0; k < ROW; ++k) x[q][k] = (q < nb && b + k < e) ? __ldg(src + c[k]) : (TIN)0;
            }
#pragma unroll
            for (int q = 0; q < BATCH; ++q) {
                if (q >= nb) break;
                TACC acc = 0;
#pragma unroll
                for (int k = 0; k < ROW; ++k)
                    if (b + k < e) acc += w[k] * (TACC)x[q][k];
                st_stream((TOUT *)fp.f[f + q].dst + a.dstOff + t, (TOUT)epilogue(acc, fp.f[f + q].epi_op, fp.f[f + q].epi_arg));
            }
            f += nb;
        }
    }
    for (; f < a.nfields; ++f) {
        const FieldDev fd = fp.f[f];
        const TIN *__restrict__ src = (const TIN *)fd.src;
        const int nlev = fd.nlev;
        if (nlev == 1) {
            TACC acc = 0;
#pragma unroll
            for (int k = 0; k < kFlatRow; ++k)
                if (b + k < e) acc += w[k] * (TACC)__ldg(src + c[k]);
            for (int k = b + kFlatRow; k < e; ++k) acc += __ldg(a.w + k) * (TACC)__ldg(src + __ldg(a.col + k));
            st_stream((TOUT *)fd.dst + a.dstOff + t, (TOUT)epilogue(acc, fd.epi_op, fd.epi_arg));
        } else {
            // short columns (soil layers): the whole column of every row entry, one coalesced store per level
            TACC acc[kShortLev];
#pragma unroll
            for (int l = 0; l < kShortLev; ++l) acc[l] = 0;
            for (int k = b; k < e; ++k) {
                const bool inreg = k - b < kFlatRow;
                TACC wk = (TACC)0;
                int ck = 0;
#pragma unroll
                for (int q = 0; q < kFlatRow; ++q)
                    if (k - b == q) { wk = w[q]; ck = c[q]; }
                if (!inreg) { wk = __ldg(a.w + k); ck = __ldg(a.col + k); }
                const TIN *p = src + (size_t)ck * nlev;
#pragma unroll
                for (int l = 0; l < kShortLev; ++l)
                    if (l < nlev) acc[l] += wk * (TACC)__ldg(p + l);
            }
#pragma unroll
            for (int l = 0; l < kShortLev; ++l)
                if (l < nlev) st_stream((TOUT *)fd.dst + (size_t)l * a.dstLev + a.dstOff + t, (TOUT)epilogue(acc[l], fd.epi_op, fd.epi_arg));
        }
    }
}

// ---------------------------------------------------------------------------
// source is a level-slowest grid field [lev][srcPlane] (centre -> edge stagger)
// ---------------------------------------------------------------------------
constexpr int kPlaneBatch = 8;  // levels whose gathers are all issued before the first FMA (memory-level parallelism)

}  // namespace mprg
#include "apply_planes.cuh"
namespace mprg {

template <typename TIN, typename TOUT, typename TACC>
__global__ void __launch_bounds__(256)
k_apply_planes(ApplyArgs<TACC> a, const __grid_constant__ FieldPack fp) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nDst) return;
    const int b = __ldg(a.rowptr + t), e = __ldg(a.rowptr + t + 1);
    if (e - b > kLongRow) return;  // pole rows of a periodic source grid: k_apply_planes_long
    int c[kFlatRow];
    TACC w[kFlatRow];
#pragma unroll
    for (int k = 0; k < kFlatRow; ++k) {
        const bool h = b + k < e;
        c[k] = h ? __ldg(a.col + b + k) : 0;   // absent entries: weight 0 on a valid address (index 0)
        w[k] = h ? __ldg(a.w + b + k) : (TACC)0;
    }
    planes_direct<TIN, TOUT, TACC>(a, fp.f[blockIdx.y], t, e - b, c, w);
}

// Long rows of a grid-source route (a pole-row point of a periodic grid averages the whole end row: ni + 2 entries,
// consecutive source columns).  One warp per (row, level): lanes stride over the entries (coalesced), fp64 partial
// sums, fixed shuffle tree -- deterministic, and spread over the whole GPU instead of 2 x ni threads.
template <typename TIN, typename TOUT, typename TACC>
__global__ void __launch_bounds__(256)
k_apply_planes_long(ApplyArgs<TACC> a, const __grid_constant__ FieldPack fp, const int32_t *__restrict__ longRows,
                    int64_t nLong) {
    const FieldDev fd = fp.f[blockIdx.y];
    const int64_t id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (id >= nLong * fd.nlev) return;
    const int lane = threadIdx.x & 31;
    const int64_t t = __ldg(longRows + id / fd.nlev);
    const int lev = (int)(id % fd.nlev);
    const int b = __ldg(a.rowptr + t), e = __ldg(a.rowptr + t + 1);
    const TIN *__restrict__ pl = (const TIN *)fd.src + (size_t)lev * a.srcPlane;
    double acc = 0.0;
    for (int k = b + lane; k < e; k += 32) acc += (double)__ldg(a.w + k) * (double)__ldg(pl + __ldg(a.col + k));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0)
        st_stream((TOUT *)fd.dst + (size_t)lev * a.dstLev + a.dstOff + t, (TOUT)epilogue((TACC)acc, fd.epi_op, fd.epi_arg));
}

// ---------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------
// Accumulation type for fp32-in / fp32-out fields.  Default fp32: every weight set on this path is a
// convex combination (bilinear, conservative) or a single 1.0 (nearest), so fp32 FMA accumulation
// stays within ~2e-7 of the fp64 result -- 50x inside the 1e-5 contract -- and nearest-neighbour
// output remains bit-exact (1.0f * x).  mprg_set_option("accumulate", "f64") / MPASSIT_GPU_ACC=f64 restores
// the reference's R8 arithmetic (one rounding on store).
static bool acc_fp32_requested(const mprg_ctx *ctx) { return !ctx->tune.acc64; }

// ---- per-launch profiling (bench.py's live roofline measurement) ---------------
static cudaEvent_t prof_event(mprg_ctx *ctx) {
    cudaEvent_t e;
    if (!ctx->evPool.empty()) { e = ctx->evPool.back(); ctx->evPool.pop_back(); return e; }
    MPRG_CUDA(cudaEventCreate(&e));
    return e;
}
// Algorithmic bytes of one launch (SURVEY.md §8d / DESIGN.md):
//   K nSrcRef b_in + K nDst b_out + nnz (4 + b_w) + (nDst + 1) 4
static double alg_bytes(const mprg_route *r, double K, size_t bin, size_t bout, size_t bw) {
    return K * (double)r->nSrcRef * bin + K * (double)r->nDst * bout + (double)r->nnz * (4.0 + bw) +
           ((double)r->nDst + 1.0) * 4.0;
}
struct ProfScope {
    mprg_ctx *ctx;
    bool on;
    mprg_ctx::ProfRec rec;
    ProfScope(mprg_ctx *c, int kind, double bytes, double units) : ctx(c), on(c->profile && !c->capturing) {
        if (!on) return;
        rec.kind = kind; rec.algBytes = bytes; rec.units = units;
        rec.a = prof_event(c); rec.b = prof_event(c);
        cudaEventRecord(rec.a, c->stream);
    }
    void cancel() {
        if (!on) return;
        on = false;
        ctx->evPool.push_back(rec.a); ctx->evPool.push_back(rec.b);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(rec.b, ctx->stream);
        ctx->prof.push_back(rec);
    }
};

// ---- pipelined column kernel (apply_pipe.cuh) -------------------------------------
template <typename KERN, typename TACC>
static void launch_pipe_k(mprg_ctx *ctx, KERN kern, const PipeArgs<TACC> &pa, const UnitPack &up, size_t smemBytes,
                          unsigned tiles) {
    MPRG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
    kern<<<tiles, kPipeThreads, smemBytes, ctx->stream>>>(pa, up);
    ctx->launches++;
}

static RecLayout route_rec_layout(const mprg_route *r, int wsize);
static const unsigned char *route_rec64(mprg_ctx *ctx, mprg_route *r);

// Every 3-D field of an apply in ONE launch of the pipelined kernel (kPipeMaxUnits units per launch).
// Fields flagged MPRG_EPI_ROT_U / ROT_V are wind pairs whose rotation is fused into the store.
// Returns false if this route / field set does not fit the kernel (the register-gather kernel takes over).
template <typename TIN, typename TOUT, typename TACC>
static bool launch_pipe(mprg_ctx *ctx, const mprg_route *r, const std::vector<FieldDev> &fields, DstLayout dl) {
    if (ctx->tune.pipeOff || r->tileEntriesMax <= 0 || r->tileEntriesMax > kPipeCap || !r->rec32.p) return false;
    const RecLayout lay = route_rec_layout(r, (int)sizeof(TACC));
    const unsigned char *rec = sizeof(TACC) == 8 ? route_rec64(ctx, const_cast<mprg_route *>(r)) : r->rec32.p;
    std::vector<UnitDev> units;
    auto unit_of = [&](const FieldDev &f, int L0) {
        UnitDev u;
        u.src = f.src; u.dst = f.dst; u.srcBytes = (size_t)r->nSrc * f.nlev * sizeof(TIN);
        u.nlev = f.nlev; u.L0 = L0; u.Ln = std::min(kPipeLev, f.nlev - L0);
        // aligned: every column chunk starts on a 16-byte boundary and is a multiple of 16 bytes long
        const bool aligned = ((size_t)f.nlev * sizeof(TIN)) % 16 == 0 && ((uintptr_t)f.src % 16) == 0 &&
                             ((size_t)L0 * sizeof(TIN)) % 16 == 0 && ((size_t)u.Ln * sizeof(TIN)) % 16 == 0;
        u.flags = (f.epi_op & 0xff) | (aligned ? kUnitAligned : 0) | (u.Ln == f.nlev ? kUnitMerged : 0);
        u.epi_arg = f.epi_arg;
        return u;
    };
    // a wind pair (ROT_U field followed by its ROT_V field) contributes alternating zonal / meridional chunks
    bool rot = false;
    for (size_t f = 0; f < fields.size(); ++f) {
        if ((fields[f].epi_op & 0xff) == MPRG_EPI_ROT_U && f + 1 < fields.size()) {
            rot = true;
            for (int L0 = 0; L0 < fields[f].nlev; L0 += kPipeLev) {
                UnitDev uu = unit_of(fields[f], L0), uv = unit_of(fields[f + 1], L0);
                uu.flags = (uu.flags & ~0xff) | kUnitRotU;
                uv.flags = (uv.flags & ~0xff) | kUnitRotV;
                units.push_back(uu);
                units.push_back(uv);
            }
            ++f;
        } else {
            for (int L0 = 0; L0 < fields[f].nlev; L0 += kPipeLev) units.push_back(unit_of(fields[f], L0));
        }
    }
    // One launch, two phases: the plain aligned units first (the kernel's lean phase A), then wind pairs and
    // unaligned columns (phase B).  pipe_split = 1 puts the two groups into separate launches instead.
    // (aligned wind pairs come right after the plain units: they keep phase A's copy protocol)
    std::vector<UnitDev> plain, rotal, rest;
    for (size_t k = 0; k < units.size(); ++k) {
        const UnitDev &u = units[k];
        const bool isrot = (u.flags & (kUnitRotU | kUnitRotV)) != 0;
        // a pair is aligned when both of its chunks are (same level count; the two sources may differ in alignment)
        bool pairAligned = false;
        if (isrot) {
            const size_t k0 = (u.flags & kUnitRotU) ? k : k - 1;
            pairAligned = (units[k0].flags & kUnitAligned) && (units[k0 + 1].flags & kUnitAligned);
        }
        if (!isrot && (u.flags & kUnitAligned)) plain.push_back(u);
        else if (isrot && pairAligned) rotal.push_back(u);
        else rest.push_back(u);
    }
    rest.insert(rest.begin(), rotal.begin(), rotal.end());
    std::vector<std::vector<UnitDev>> groups;
    std::vector<size_t> groupPlain;
    if (ctx->tune.pipeSplit) {
        if (!plain.empty()) { groupPlain.push_back(plain.size()); groups.push_back(plain); }
        if (!rest.empty()) { groupPlain.push_back(0); groups.push_back(rest); }
    } else {
        groupPlain.push_back(plain.size());
        plain.insert(plain.end(), rest.begin(), rest.end());
        groups.push_back(std::move(plain));
    }
    PipeArgs<TACC> pa;
    pa.rec = rec;
    pa.lay = lay;
    pa.nDst = r->nDst;
    pa.dstLev = dl.lev; pa.dstOff = dl.off;
    pa.dstLev32 = (uint32_t)dl.lev;
    pa.ni = r->dstNi;
    pa.tilesPerRow = (r->dstNi + kPipeTile - 1) / kPipeTile;
    pa.rotc = nullptr;
    if (rot) {  // checked by apply_device: rotation registered, destination on CENTER / CENTER_HALO rows
        using TR = typename RotMath<TOUT, TACC>::type;
        const int64_t off = 4 * ctx->target[r->dst_stagger].slabOffset();
        pa.rotc = sizeof(TR) == 4 ? (const void *)(ctx->rotc32.p + off) : (const void *)(ctx->rotc.p + off);
    }
    const unsigned tiles = (unsigned)(((r->nDst + r->dstNi - 1) / r->dstNi) * pa.tilesPerRow);
    struct Launch { size_t g, u0, nu, smem; int mode, minb, nPlain, nRotA; int32_t stageOff, stageBytes, holdOff; };
    std::vector<Launch> plan;
    for (size_t g = 0; g < groups.size(); ++g) {
        const std::vector<UnitDev> &us = groups[g];
        for (size_t u0 = 0; u0 < us.size();) {
            size_t nu = std::min<size_t>(kPipeMaxUnits, us.size() - u0);
            if (u0 + nu < us.size() && (us[u0 + nu - 1].flags & kUnitRotU)) --nu;  // keep a wind pair in one launch
            int mode = 0;
            size_t stage = 0;   // one stage holds the largest unit of the launch at the route's tile maxima
            const int nPlain = (int)std::min<size_t>(nu, groupPlain[g] > u0 ? groupPlain[g] - u0 : 0);
            int nRotA = 0;   // aligned wind-pair units directly after the plain ones
            for (size_t k = (size_t)nPlain; k < nu; ++k) {
                const UnitDev &u = us[u0 + k];
                if (!(u.flags & (kUnitRotU | kUnitRotV)) || !(u.flags & kUnitAligned)) break;
                ++nRotA;
            }
            nRotA &= ~1;
            for (size_t k = 0; k < nu; ++k) {
                const UnitDev &u = us[u0 + k];
                if ((int)k >= nPlain + nRotA && !(u.flags & kUnitAligned)) mode |= kModeUnal;
                if ((int)k >= nPlain && (u.flags & (kUnitRotU | kUnitRotV))) mode |= kModeRot;
                stage = std::max(stage, pipe_unit_stage_bytes((u.flags & kUnitAligned) != 0, (u.flags & kUnitMerged) != 0,
                                                              (unsigned)(u.Ln * sizeof(TIN)), r->tileUniqMax, r->tileRunsMax));
            }
            stage = (stage + 15) & ~(size_t)15;
            const size_t fixed = ((size_t)kPipeSmemHead + lay.stride + nu * sizeof(UnitDev) + 15) & ~(size_t)15;
            using TRot = typename RotMath<TOUT, TACC>::type;
            const size_t hold = (mode & kModeRot) ? (size_t)(kPipeLev / 4 / kPipeWarps) * kPipeThreads * 4 * sizeof(TOUT) : 0;
            const size_t rotc = (mode & kModeRot) ? (size_t)kPipeTile * 4 * sizeof(TRot) : 0;   // the tile's rotation constants
            const size_t smemBytes = fixed + kPipeStages * stage + hold + rotc;
            if (smemBytes + 1024 > (size_t)227 * 1024) return false;
            // 5 resident CTAs per SM (48 registers) when their shared memory fits, else 4 (64 registers)
            int minb = (smemBytes + 1024) * 5 <= (size_t)228 * 1024 ? 5 : 4;
            if (ctx->tune.pipeMinb) minb = ctx->tune.pipeMinb >= 5 ? 5 : 4;
            plan.push_back(Launch{g, u0, nu, smemBytes, mode, minb, nPlain, nRotA, (int32_t)fixed, (int32_t)stage,
                                  (int32_t)(fixed + kPipeStages * stage)});
            u0 += nu;
        }
    }
    for (const Launch &l : plan) {   // nothing is launched unless every launch of the apply fits
        UnitPack up;
        memcpy(up.u, groups[l.g].data() + l.u0, l.nu * sizeof(UnitDev));
        pa.nunits = (int)l.nu;
        pa.nPlain = l.nPlain; pa.nRotA = l.nRotA;
        pa.stageOff = l.stageOff; pa.stageBytes = l.stageBytes; pa.holdOff = l.holdOff;
        pa.rotOff = (int32_t)(l.holdOff + ((l.mode & kModeRot) ? (size_t)(kPipeLev / 4 / kPipeWarps) * kPipeThreads * 4 * sizeof(TOUT) : 0));
        const size_t smemBytes = l.smem;
        const int minb = l.minb;
#define MPRG_PIPE_LAUNCH(M)                                                                                        \
    if (minb >= 5) launch_pipe_k(ctx, k_apply_pipe<TIN, TOUT, TACC, M, 5>, pa, up, smemBytes, tiles);             \
    else launch_pipe_k(ctx, k_apply_pipe<TIN, TOUT, TACC, M, 4>, pa, up, smemBytes, tiles)
        switch (l.mode) {
            case 0: MPRG_PIPE_LAUNCH(0); break;
            case kModeUnal: MPRG_PIPE_LAUNCH(kModeUnal); break;
            case kModeRot: MPRG_PIPE_LAUNCH(kModeRot); break;
            default: MPRG_PIPE_LAUNCH(kModeUnal | kModeRot); break;
        }
#undef MPRG_PIPE_LAUNCH
    }
    MPRG_CUDA(cudaGetLastError());
    return true;
}

void scan_counts(mprg_ctx *ctx, const int32_t *cnt, int32_t *rowptr, int64_t nPlus1);  // locate.cu

static RecLayout route_rec_layout(const mprg_route *r, int wsize) {
    return rec_layout(wsize, r->tileUniqMax, r->tileRunsMax, r->tileEntriesMax, r->tileRowMax > 3, r->composite);
}

template <typename TW>
static void route_build_records(mprg_ctx *ctx, mprg_route *r, const TW *w, const TW *w2, DevBuf<unsigned char> &out) {
    const int tilesPerRow = (r->dstNi + kPipeTile - 1) / kPipeTile;
    const unsigned tiles = (unsigned)(((r->nDst + r->dstNi - 1) / r->dstNi) * tilesPerRow);
    const RecLayout lay = route_rec_layout(r, (int)sizeof(TW));
    out.alloc((size_t)tiles * lay.stride);
    MPRG_CUDA(cudaMemsetAsync(out.p, 0, (size_t)tiles * lay.stride, ctx->stream));
    if (r->composite)
        k_tile_schedule<true, TW, 2 * kPipeCap><<<tiles, 2 * kPipeCap, 0, ctx->stream>>>(r->rowptr.p, r->col.p, w, w2, r->nDst, r->dstNi,
                                                                                         tilesPerRow, nullptr, nullptr, out.p, lay);
    else
        k_tile_schedule<true, TW, kPipeCap><<<tiles, kPipeCap, 0, ctx->stream>>>(r->rowptr.p, r->col.p, w, w2, r->nDst, r->dstNi, tilesPerRow,
                                                                                 nullptr, nullptr, out.p, lay);
    ctx->launches++;
    MPRG_CUDA(cudaGetLastError());
}

// records with fp64 weights (fp64 accumulation / fp64 fields): built the first time an apply needs them
static const unsigned char *route_rec64(mprg_ctx *ctx, mprg_route *r) {
    if (!r->rec64.p) {
        if (ctx->capturing) fail(46, "mprg_apply: this route's fp64 schedule is not built yet; run the pass once before mprg_capture_begin");
        route_build_records<double>(ctx, r, r->w.p, r->composite ? r->w2.p : nullptr, r->rec64);
    }
    return r->rec64.p;
}

// builds the route's tile schedule (the "communication schedule" half of an ESMF route handle)
void route_tile_stats(mprg_ctx *ctx, mprg_route *r) {
    r->tileEntriesMax = r->tileUniqMax = r->tileRunsMax = r->tileRowMax = 0;
    if (r->nDst <= 0 || r->nnz <= 0) return;
    if (r->dstNi <= 0) r->dstNi = (int32_t)std::min<int64_t>(r->nDst, 0x7fffffff);
    const int tilesPerRow = (r->dstNi + kPipeTile - 1) / kPipeTile;
    const unsigned tiles = (unsigned)(((r->nDst + r->dstNi - 1) / r->dstNi) * tilesPerRow);
    DevBuf<int32_t> mm(4);
    DevBuf<unsigned long long> tot(2);
    MPRG_CUDA(cudaMemsetAsync(mm.p, 0, 4 * sizeof(int32_t), ctx->stream));
    MPRG_CUDA(cudaMemsetAsync(tot.p, 0, 2 * sizeof(unsigned long long), ctx->stream));
    if (r->composite)
        k_tile_schedule<false, float, 2 * kPipeCap><<<tiles, 2 * kPipeCap, 0, ctx->stream>>>(r->rowptr.p, r->col.p, r->w32.p, nullptr, r->nDst,
                                                                                             r->dstNi, tilesPerRow, mm.p, tot.p, nullptr, RecLayout{});
    else
        k_tile_schedule<false, float, kPipeCap><<<tiles, kPipeCap, 0, ctx->stream>>>(r->rowptr.p, r->col.p, r->w32.p, nullptr, r->nDst, r->dstNi,
                                                                                     tilesPerRow, mm.p, tot.p, nullptr, RecLayout{});
    ctx->launches++;
    int32_t h[4] = {0, 0, 0, 0};
    unsigned long long ht[2] = {0, 0};
    peek(ctx, h, mm.p, sizeof h);
    peek(ctx, ht, tot.p, sizeof ht);
    r->tileEntriesMax = h[0];
    r->tileUniqMax = h[1];
    r->tileRunsMax = h[2];
    r->tileRowMax = h[3];
    if (h[0] > (r->composite ? 2 * kPipeCap : kPipeCap) || h[1] > 255 || h[2] > 255) return;  // no schedule: register-gather kernels only
    route_build_records<float>(ctx, r, r->w32.p, r->composite ? r->w2_32.p : nullptr, r->rec32);
    r->rec64 = DevBuf<unsigned char>();
    r->schedTiles = tiles; r->schedCols = (int64_t)ht[0]; r->schedRuns = (int64_t)ht[1];
}

// tile geometry of the TMA-staged planes kernel for a grid-source route
static PlaneGeom plane_geom(const mprg_route *r) {
    PlaneGeom pg;
    pg.tilesPerRow = (r->dstNi + kPlTile - 1) / kPlTile;
    pg.tw = (r->dstNi + pg.tilesPerRow - 1) / pg.tilesPerRow;
    pg.dstNi = r->dstNi; pg.srcNi = r->srcNi;
    pg.nWin = r->planeWin;
    pg.tiles = r->planeTiles.p;
    return pg;
}

// route_finish: how many source rows the tiles of a grid-source route reference (0: the route cannot be staged)
void route_plane_stats(mprg_ctx *ctx, mprg_route *r) {
    r->planeWin = 0;
    if (!r->srcLevelSlowest || r->srcNi <= 0 || r->dstNi <= 0 || r->nDst <= 0 || r->nDst % r->dstNi != 0 || r->nnz <= 0) return;
    PlaneGeom pg = plane_geom(r);
    DevBuf<int32_t> st(2);
    MPRG_CUDA(cudaMemsetAsync(st.p, 0, 2 * sizeof(int32_t), ctx->stream));
    const unsigned tiles = (unsigned)((r->nDst / r->dstNi) * pg.tilesPerRow);
    r->planeTiles.alloc((size_t)tiles * 8);
    k_plane_stats<<<tiles, kPlTile, 0, ctx->stream>>>(r->rowptr.p, r->col.p, pg, st.p, r->planeTiles.p);
    ctx->launches++;
    int32_t h[2] = {0, 0};
    peek(ctx, h, st.p, sizeof h);
    // staged when most tiles fit (the rest -- the seam tiles of a periodic grid -- take the gather path inside the kernel)
    if (h[0] > 0 && (int64_t)h[1] * 4 <= (int64_t)tiles) r->planeWin = h[0];
}

// cols: every 3-D field of the apply (wind pairs adjacent, ROT_U then ROT_V); flat: 2-D fields and short columns;
// planes: grid-source fields.  Returns false when wind pairs (fused rotation) are present but the pipelined
// kernel could not take the route.
template <typename TIN, typename TOUT, typename TACC>
static bool launch_all(mprg_ctx *ctx, const mprg_route *r, const std::vector<FieldDev> &cols,
                       const std::vector<FieldDev> &flat, const std::vector<FieldDev> &planes, DstLayout dl) {
    auto has_rot = [](const std::vector<FieldDev> &v) {
        for (auto &f : v) if (f.epi_op == MPRG_EPI_ROT_U) return true;
        return false;
    };
    const size_t total = cols.size() + flat.size() + planes.size();
    if (total == 0 || r->nDst == 0) return true;
    auto ksum = [](const std::vector<FieldDev> &v) { double k = 0; for (auto &f : v) k += f.nlev; return k; };
    bool piped = false;
    if (!cols.empty()) {
        ProfScope ps(ctx, 0, alg_bytes(r, ksum(cols), sizeof(TIN), sizeof(TOUT), sizeof(TACC)), ksum(cols) * r->nDst);
        piped = launch_pipe<TIN, TOUT, TACC>(ctx, r, cols, dl);
        if (!piped) ps.cancel();
        if (!piped && has_rot(cols)) return false;
    }
    ApplyArgs<TACC> a;
    a.rowptr = r->rowptr.p;
    a.col = r->col.p;
    if (sizeof(TACC) == 8) a.w = (const TACC *)r->w.p; else a.w = (const TACC *)r->w32.p;
    a.nDst = r->nDst;
    a.dstLev = dl.lev; a.dstOff = dl.off;
    a.dstLev32 = (uint32_t)dl.lev;
    a.uniformRow = (r->uniform && r->nUnmapped == 0 && r->maxRow > 0 && r->nnz == r->nDst * r->maxRow) ? r->maxRow : 0;
    a.srcPlane = r->srcPlane;
    const unsigned tiles = (unsigned)((r->nDst + kTile - 1) / kTile);
    // field-descriptor kernels take their fields kPackFields at a time, as a kernel parameter
    auto packs = [&](const std::vector<FieldDev> &v, auto &&launch) {
        for (size_t f0 = 0; f0 < v.size(); f0 += kPackFields) {
            const size_t n = std::min<size_t>(kPackFields, v.size() - f0);
            FieldPack fp;
            memcpy(fp.f, v.data() + f0, n * sizeof(FieldDev));
            a.nfields = (int)n;
            launch(fp, n);
            ctx->launches++;
        }
    };
    if (!cols.empty() && !piped) {
        // register-gather fallback for routes whose tiles exceed the pipelined kernel's caps
        ProfScope ps(ctx, 1, alg_bytes(r, ksum(cols), sizeof(TIN), sizeof(TOUT), sizeof(TACC)), ksum(cols) * r->nDst);
        std::vector<FieldDev> vec, sca;
        for (auto &f : cols)
            (((size_t)f.nlev * sizeof(TIN)) % 16 == 0 && ((uintptr_t)f.src % 16) == 0 ? vec : sca).push_back(f);
        const int minb = ctx->tune.colsMinb;
        packs(vec, [&](const FieldPack &fp, size_t n) {
            dim3 g(tiles, (unsigned)((n + kFieldsPerCta - 1) / kFieldsPerCta));
            if (minb == 2) k_apply_cols<TIN, TOUT, TACC, true, 2><<<g, kThreads, 0, ctx->stream>>>(a, fp);
            else if (minb == 4) k_apply_cols<TIN, TOUT, TACC, true, 4><<<g, kThreads, 0, ctx->stream>>>(a, fp);
            else k_apply_cols<TIN, TOUT, TACC, true, 3><<<g, kThreads, 0, ctx->stream>>>(a, fp);
        });
        packs(sca, [&](const FieldPack &fp, size_t n) {
            dim3 g(tiles, (unsigned)((n + kFieldsPerCta - 1) / kFieldsPerCta));
            k_apply_cols<TIN, TOUT, TACC, false, 3><<<g, kThreads, 0, ctx->stream>>>(a, fp);
        });
    }
    if (!flat.empty()) {
        ProfScope ps(ctx, 2, alg_bytes(r, ksum(flat), sizeof(TIN), sizeof(TOUT), sizeof(TACC)), ksum(flat) * r->nDst);
        packs(flat, [&](const FieldPack &fp, size_t) {
            const unsigned g = (unsigned)((r->nDst + 255) / 256);
            if (r->maxRow <= 3) k_apply_flat<TIN, TOUT, TACC, 3, (sizeof(TACC) == 4 ? 6 : 4)><<<g, 256, 0, ctx->stream>>>(a, fp);
            else k_apply_flat<TIN, TOUT, TACC, kFlatRow, 4><<<g, 256, 0, ctx->stream>>>(a, fp);
        });
    }
    if (!planes.empty()) {
        ProfScope ps(ctx, 3, alg_bytes(r, ksum(planes), sizeof(TIN), sizeof(TOUT), sizeof(TACC)), ksum(planes) * r->nDst);
        // TMA-staged kernel when the slab is whole destination rows of a known source grid whose tiles' windows fit
        // (route_finish: planeWin); "apply" = "direct" keeps the register-gather kernel
        bool staged = !ctx->tune.pipeOff && r->planeWin > 0 && ((size_t)r->srcPlane * sizeof(TIN)) % 16 == 0;   // windows keep their 16-byte phase from level to level
        for (const FieldDev &f : planes) staged = staged && ((uintptr_t)f.src % 16) == 0;                          // ... and from field to field
        packs(planes, [&](const FieldPack &fp, size_t n) {
            if (staged) {
                PlaneGeom pg = plane_geom(r);
                const unsigned tiles = (unsigned)((r->nDst / r->dstNi) * pg.tilesPerRow);
                auto go = [&](auto lev_c, auto st_c) {
                    constexpr int LEV = decltype(lev_c)::value, ST = decltype(st_c)::value;
                    const unsigned smem = (unsigned)ST * pg.nWin * LEV * pl_win_bytes<TIN>();
                    static unsigned attr = 0;
                    if (smem > attr) {
                        MPRG_CUDA(cudaFuncSetAttribute(k_apply_planes_pipe<TIN, TOUT, TACC, LEV, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        attr = smem;
                    }
                    k_apply_planes_pipe<TIN, TOUT, TACC, LEV, ST><<<tiles, kPlThreads, smem, ctx->stream>>>(a, fp, pg);
                };
                using std::integral_constant;
                switch (ctx->tune.planesShape) {
                // measured on the 3-km case (profiles/r02/README.md): 4 levels x 4 stages 0.44 ms per pass, 4 x 6 0.55,
                // 8 x 3 0.50, 8 x 4 0.53 -- deeper pipelines are slower, as for the column kernel
                case 83: go(integral_constant<int, 8>{}, integral_constant<int, 3>{}); break;
                default: go(integral_constant<int, 4>{}, integral_constant<int, 4>{}); break;
                }
            } else {
                dim3 g((unsigned)((r->nDst + 255) / 256), (unsigned)n);
                k_apply_planes<TIN, TOUT, TACC><<<g, 256, 0, ctx->stream>>>(a, fp);
            }
            if (r->nLong > 0) {
                int maxLev = 0;
                for (size_t i = 0; i < n; ++i) maxLev = std::max(maxLev, (int)fp.f[i].nlev);
                dim3 gl((unsigned)((r->nLong * maxLev + 7) / 8), (unsigned)n);
                k_apply_planes_long<TIN, TOUT, TACC><<<gl, 256, 0, ctx->stream>>>(a, fp, r->longRows.p, r->nLong);
                ctx->launches++;
            }
        });
    }
    MPRG_CUDA(cudaGetLastError());
    return true;
}

void apply_device(mprg_ctx *ctx, const mprg_route *r, const ApplyField *fields, int nfields, int src_dtype,
                  int dst_dtype, bool into_full) {
    if (r->composite) fail(47, "mprg_apply: a composed wind route takes mprg_apply_wind");
    // destination layout: this rank's slab buffer, or (into_full) its rows inside a full-grid field
    DstLayout dl{r->nDst, 0};
    if (into_full) {
        if (r->dst_stagger < MPRG_CENTER || r->dst_stagger > MPRG_CORNER) fail(38, "mprg_apply_into: the route has no full-grid destination");
        const Target &tg = ctx->target[r->dst_stagger];
        dl.lev = (int64_t)tg.ni * tg.nj;
        dl.off = tg.slabOffset();
    }
    if (dl.lev >= ((int64_t)1 << 32)) fail(40, "mprg_apply: destination level stride %lld does not fit 32 bits", (long long)dl.lev);
    std::vector<FieldDev> cols, rots, flat, planes;   // cols: plain 3-D fields; rots: wind pairs (appended to cols)
    std::vector<std::pair<void *, void *>> pairs;  // every (u, v) destination pair with fused rotation requested
    std::vector<int32_t> pair_nlev;
    bool late_rot = false;                          // some pair is too short for the column kernel: rotate afterwards
    for (int f = 0; f < nfields; ++f) {
        if (fields[f].nlev <= 0) fail(31, "mprg_apply: field %d has nlev %d", f, fields[f].nlev);
        if (!fields[f].src || !fields[f].dst) fail(32, "mprg_apply: field %d has a null buffer", f);
        const size_t esz = src_dtype == MPRG_F32 ? 4 : 8;
        if ((uintptr_t)fields[f].src % esz || (uintptr_t)fields[f].dst % (dst_dtype == MPRG_F32 ? 4 : 8))
            fail(34, "mprg_apply: field %d is not aligned to its element size", f);
        FieldDev d{fields[f].src, fields[f].dst, fields[f].nlev, fields[f].epi_op, fields[f].epi_arg};
        if (d.epi_op == MPRG_EPI_ROT_V) fail(35, "mprg_apply: MPRG_EPI_ROT_V field %d has no MPRG_EPI_ROT_U before it", f);
        if (d.epi_op == MPRG_EPI_ROT_U) {
            if (into_full) fail(39, "mprg_apply_into: wind pairs (fused rotation) are intermediates of the wind chain, not gathered fields");
            if (f + 1 >= nfields || fields[f + 1].epi_op != MPRG_EPI_ROT_V || fields[f + 1].nlev != d.nlev)
                fail(35, "mprg_apply: MPRG_EPI_ROT_U field %d needs a following MPRG_EPI_ROT_V field of the same level count", f);
            if (!fields[f + 1].src || !fields[f + 1].dst) fail(32, "mprg_apply: field %d has a null buffer", f + 1);
            if (r->dst_stagger != MPRG_CENTER && r->dst_stagger != MPRG_CENTER_HALO)
                fail(36, "mprg_apply: wind rotation is defined on CENTER rows");
            if (!ctx->haveRot) fail(41, "mprg_apply: mprg_set_rotation was not called");
            if (r->srcLevelSlowest) fail(37, "mprg_apply: wind rotation needs a mesh source");
            FieldDev v{fields[f + 1].src, fields[f + 1].dst, fields[f + 1].nlev, fields[f + 1].epi_op, fields[f + 1].epi_arg};
            pairs.emplace_back(d.dst, v.dst);
            pair_nlev.push_back(d.nlev);
            if (d.nlev <= kShortLev) {  // flat kernel, rotated in a second (tiny) pass
                d.epi_op = v.epi_op = MPRG_EPI_NONE;
                flat.push_back(d); flat.push_back(v);
                late_rot = true;
            } else {
                rots.push_back(d); rots.push_back(v);  // adjacent: the unit builder interleaves their chunks
            }
            ++f;
            continue;
        }
        if (r->srcLevelSlowest) planes.push_back(d);
        else if (d.nlev <= kShortLev) flat.push_back(d);
        else cols.push_back(d);
    }
    // one column launch: aligned level counts first (the LDG staging caches the first unit's chunk offsets), then
    // the others, then the wind pairs
    std::stable_sort(cols.begin(), cols.end(), [&](const FieldDev &x, const FieldDev &y) {
        const size_t esz = src_dtype == MPRG_F32 ? 4 : 8;
        return ((x.nlev * esz) % 16 == 0) > ((y.nlev * esz) % 16 == 0);
    });
    const size_t nplain = cols.size();
    cols.insert(cols.end(), rots.begin(), rots.end());
    auto run = [&]() {
        if (src_dtype == MPRG_F32 && dst_dtype == MPRG_F32) {
            return acc_fp32_requested(ctx) ? launch_all<float, float, float>(ctx, r, cols, flat, planes, dl)
                                           : launch_all<float, float, double>(ctx, r, cols, flat, planes, dl);
        } else if (src_dtype == MPRG_F32 && dst_dtype == MPRG_F64) {
            return launch_all<float, double, double>(ctx, r, cols, flat, planes, dl);
        } else if (src_dtype == MPRG_F64 && dst_dtype == MPRG_F32) {
            return launch_all<double, float, double>(ctx, r, cols, flat, planes, dl);
        } else if (src_dtype == MPRG_F64 && dst_dtype == MPRG_F64) {
            return launch_all<double, double, double>(ctx, r, cols, flat, planes, dl);
        }
        fail(33, "mprg_apply: bad dtype %d/%d", src_dtype, dst_dtype);
    };
    bool all_late = false;
    if (!run()) {
        // the route does not fit the pipelined kernel: regrid the wind pairs like any field (register-gather
        // kernels), then rotate them in a second pass
        for (size_t k = nplain; k < cols.size(); ++k) cols[k].epi_op = MPRG_EPI_NONE;
        run();
        all_late = true;
    }
    if (late_rot || all_late)
        for (size_t p = 0; p < pairs.size(); ++p)
            if (all_late || pair_nlev[p] <= kShortLev)
                rotate_device(ctx, r->dst_stagger, pairs[p].first, pairs[p].second, pair_nlev[p], dst_dtype);
}

// ---------------------------------------------------------------------------
// composed wind route (compose.cu): dst = A u_src + B v_src in one launch of the column kernel
// ---------------------------------------------------------------------------
template <typename TIN, typename TOUT, typename TACC>
static void launch_wind(mprg_ctx *ctx, const mprg_route *r, const void *u, const void *v, int32_t nlev, void *dst, DstLayout dl) {
    if (r->tileEntriesMax <= 0 || r->tileEntriesMax > 2 * kPipeCap || !r->rec32.p)
        fail(47, "mprg_apply_wind: the composed route has no tile schedule");
    if (((size_t)nlev * sizeof(TIN)) % 16 || (uintptr_t)u % 16 || (uintptr_t)v % 16)
        fail(48, "mprg_apply_wind: source columns must be 16-byte aligned (level count x element size a multiple of 16)");
    const RecLayout lay = route_rec_layout(r, (int)sizeof(TACC));
    const unsigned char *rec = sizeof(TACC) == 8 ? route_rec64(ctx, const_cast<mprg_route *>(r)) : r->rec32.p;
    UnitPack up;
    int nu = 0;
    size_t stage = 0;
    for (int L0 = 0; L0 < nlev; L0 += kPipeLev) {
        for (int s = 0; s < 2; ++s) {
            if (nu >= kPipeMaxUnits) fail(49, "mprg_apply_wind: too many levels");
            UnitDev &d = up.u[nu++];
            d.src = s ? v : u; d.dst = dst; d.srcBytes = (size_t)r->nSrc * nlev * sizeof(TIN);
            d.nlev = nlev; d.L0 = L0; d.Ln = std::min(kPipeLev, nlev - L0);
            const bool aligned = ((size_t)L0 * sizeof(TIN)) % 16 == 0 && ((size_t)d.Ln * sizeof(TIN)) % 16 == 0;
            if (!aligned) fail(48, "mprg_apply_wind: level chunk not 16-byte aligned");
            d.flags = kUnitAligned | (d.Ln == nlev ? kUnitMerged : 0) | (s ? kUnitCmpB : kUnitCmpA);
            d.epi_arg = 0.0;
            stage = std::max(stage, pipe_unit_stage_bytes(true, (d.flags & kUnitMerged) != 0, (unsigned)(d.Ln * sizeof(TIN)),
                                                          r->tileUniqMax, r->tileRunsMax));
        }
    }
    stage = (stage + 15) & ~(size_t)15;
    PipeArgs<TACC> pa;
    pa.rec = rec; pa.lay = lay; pa.nDst = r->nDst;
    pa.dstLev = dl.lev; pa.dstOff = dl.off;
    pa.dstLev32 = (uint32_t)dl.lev;
    pa.ni = r->dstNi;
    pa.tilesPerRow = (r->dstNi + kPipeTile - 1) / kPipeTile;
    pa.rotc = nullptr;
    pa.nunits = nu; pa.nPlain = 0; pa.nRotA = 0; pa.rotOff = 0;
    const size_t fixed = ((size_t)kPipeSmemHead + lay.stride + nu * sizeof(UnitDev) + 15) & ~(size_t)15;
    const size_t hold = (size_t)(kPipeLev / 4 / kPipeWarps) * kPipeThreads * 4 * sizeof(TACC);
    const size_t smemBytes = fixed + kPipeStages * stage + hold;
    if (smemBytes + 1024 > (size_t)227 * 1024) fail(47, "mprg_apply_wind: tile too large for shared memory");
    pa.stageOff = (int32_t)fixed; pa.stageBytes = (int32_t)stage; pa.holdOff = (int32_t)(fixed + kPipeStages * stage);
    const unsigned tiles = (unsigned)(((r->nDst + r->dstNi - 1) / r->dstNi) * pa.tilesPerRow);
    const int minb = (smemBytes + 1024) * 5 <= (size_t)228 * 1024 ? 5 : 4;
    const double K = nlev;
    ProfScope ps(ctx, 4,
                 2.0 * K * (double)r->nSrcRef * sizeof(TIN) + K * (double)r->nDst * sizeof(TOUT) +
                     (double)r->nnz * (4.0 + 2.0 * sizeof(TACC)) + ((double)r->nDst + 1.0) * 4.0,
                 K * (double)r->nDst);
    if (minb >= 5) launch_pipe_k(ctx, k_apply_pipe<TIN, TOUT, TACC, kModeCmp, 5>, pa, up, smemBytes, tiles);
    else launch_pipe_k(ctx, k_apply_pipe<TIN, TOUT, TACC, kModeCmp, 4>, pa, up, smemBytes, tiles);
    MPRG_CUDA(cudaGetLastError());
}

void apply_wind_device(mprg_ctx *ctx, const mprg_route *r, const void *u, const void *v, int32_t nlev, int src_dtype,
                       void *dst, int dst_dtype, bool into_full) {
    if (!r->composite) fail(47, "mprg_apply_wind: not a composed wind route (mprg_store_wind)");
    if (nlev <= 0) fail(31, "mprg_apply_wind: nlev %d", nlev);
    if (!u || !v || !dst) fail(32, "mprg_apply_wind: null buffer");
    if (r->nDst == 0) return;
    DstLayout dl{r->nDst, 0};
    if (into_full) {
        const Target &tg = ctx->target[r->dst_stagger];
        dl.lev = (int64_t)tg.ni * tg.nj;
        dl.off = tg.slabOffset();
    }
    if (src_dtype == MPRG_F32 && dst_dtype == MPRG_F32) {
        if (acc_fp32_requested(ctx)) launch_wind<float, float, float>(ctx, r, u, v, nlev, dst, dl);
        else launch_wind<float, float, double>(ctx, r, u, v, nlev, dst, dl);
    } else if (src_dtype == MPRG_F32 && dst_dtype == MPRG_F64) {
        launch_wind<float, double, double>(ctx, r, u, v, nlev, dst, dl);
    } else if (src_dtype == MPRG_F64 && dst_dtype == MPRG_F32) {
        launch_wind<double, float, double>(ctx, r, u, v, nlev, dst, dl);
    } else if (src_dtype == MPRG_F64 && dst_dtype == MPRG_F64) {
        launch_wind<double, double, double>(ctx, r, u, v, nlev, dst, dl);
    } else {
        fail(33, "mprg_apply_wind: bad dtype %d/%d", src_dtype, dst_dtype);
    }
}

// ---------------------------------------------------------------------------
// rotate_winds_cgrid, interp.F90:689-749 (v' uses the already rotated u')
// ---------------------------------------------------------------------------
// per-point constants of rotate_winds_cgrid: u' = (u + v tana) / (cosa + sina tana); v' = (v - u' sina) / cosa.
// The divisors do not depend on the level, so they are applied as reciprocals prepared once per grid
// (<= 1 ulp of fp64 from the divisions); both the stand-alone pass and the rotation fused into
// k_apply_pipe read them, which keeps the two bit-identical.
__global__ void k_rot_consts(const double *__restrict__ cosa, const double *__restrict__ sina, int64_t n,
                             double *__restrict__ rotc, float *__restrict__ rotc32) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double ca = cosa[i], sa = sina[i];
    const double tana = sa / ca;
    rotc[4 * i + 0] = sa;
    rotc[4 * i + 1] = tana;
    rotc[4 * i + 2] = 1.0 / ca;
    rotc[4 * i + 3] = 1.0 / (ca + sa * tana);
    for (int k = 0; k < 4; ++k) rotc32[4 * i + k] = (float)rotc[4 * i + k];   // what an all-fp32 apply rotates with
}

void rotation_constants(mprg_ctx *ctx, int64_t n) {
    ctx->rotc.alloc(4 * (size_t)n);
    ctx->rotc32.alloc(4 * (size_t)n);
    k_rot_consts<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->cosa.p, ctx->sina.p, n, ctx->rotc.p, ctx->rotc32.p);
    ctx->launches++;
    MPRG_CUDA(cudaGetLastError());
}

// TR = the arithmetic type (RotMath, apply_pipe.cuh): the same expressions as the fused rotation
template <typename T, typename TR>
__global__ void k_rotate(T *__restrict__ u, T *__restrict__ v, const double *__restrict__ rotc, int64_t n, int32_t nlev) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const TR rsa = (TR)rotc[4 * i], rtana = (TR)rotc[4 * i + 1], rcai = (TR)rotc[4 * i + 2], rdeni = (TR)rotc[4 * i + 3];
    for (int l = blockIdx.y; l < nlev; l += gridDim.y) {
        const size_t o = (size_t)l * n + i;
        TR uu = (TR)u[o], vv = (TR)v[o];
        uu = (uu + vv * rtana) * rdeni;
        vv = (vv - uu * rsa) * rcai;
        u[o] = (T)uu;
        v[o] = (T)vv;
    }
}

void rotate_device(mprg_ctx *ctx, int stagger, void *u, void *v, int32_t nlev, int dtype) {
    if (!ctx->haveRot) fail(41, "mprg_rotate_winds: mprg_set_rotation was not called");
    const Target &tg = ctx->target[stagger];
    const int64_t n = tg.nSlab();
    if (n == 0 || nlev <= 0) return;
    const double *rotc = ctx->rotc.p + 4 * tg.slabOffset();
    dim3 g((unsigned)((n + 255) / 256), (unsigned)min(nlev, 4));  // >= 15 levels per thread: tana, den once per point
    if (dtype == MPRG_F32 && acc_fp32_requested(ctx))
        k_rotate<float, float><<<g, 256, 0, ctx->stream>>>((float *)u, (float *)v, rotc, n, nlev);
    else if (dtype == MPRG_F32)
        k_rotate<float, double><<<g, 256, 0, ctx->stream>>>((float *)u, (float *)v, rotc, n, nlev);
    else
        k_rotate<double, double><<<g, 256, 0, ctx->stream>>>((double *)u, (double *)v, rotc, n, nlev);
    ctx->launches++;
    MPRG_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// WRF-compatibility post-ops (write_data.F90:1364-1373, 1406-1412)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void k_midlevels(const T *__restrict__ x, T *__restrict__ mid, int64_t n, int32_t nlev) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double lo = (double)x[i];
    for (int k = 1; k < nlev; ++k) {
        const double hi = (double)x[(size_t)k * n + i];
        mid[(size_t)(k - 1) * n + i] = (T)(0.5 * (hi + lo));  // 0.5*(dum3dp1(k) + dum3dp1(k-1)), R8 arithmetic
        lo = hi;
    }
}

// out[0] = max over the whole field, out[1] = min of 0.8 * top-level value over points with top >= 10
// (doubles kept as ordered integers for atomicMax / atomicMin; all values of interest are finite)
__device__ __forceinline__ long long dbl_key(double v) {
    long long b = __double_as_longlong(v);
    return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
}
template <typename T>
__global__ void k_ptop(const T *__restrict__ x, int64_t n, int32_t nlev, long long *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double mx = -1.0e300, mn = 1.0e300;
    if (i < n) {
        for (int k = 0; k < nlev; ++k) mx = fmax(mx, (double)x[(size_t)k * n + i]);
        const double top = (double)x[(size_t)(nlev - 1) * n + i];
        if (top >= 10.0) mn = top * 0.80;
    }
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out, dbl_key(mx));
        atomicMin(out + 1, dbl_key(mn));
    }
}
__global__ void k_ptop_finish(const long long *in, double *out) {
    for (int k = 0; k < 2; ++k) {
        long long b = in[k];
        out[k] = __longlong_as_double(b >= 0 ? b : (b ^ 0x7fffffffffffffffLL));
    }
}

// Byte order of file data: NetCDF classic / CDF-5 variables are big-endian.  The reference lets the NetCDF
// library swap on the host (nf90_get_var / nf90_put_var, input_data.F90:186, write_data.F90:1010); here file
// bytes cross PCIe untouched and are swapped in HBM (2 x bytes of traffic at TB/s instead of a host pass).
template <typename U>
__device__ __forceinline__ U bswap_word(U v);
template <>
__device__ __forceinline__ uint32_t bswap_word<uint32_t>(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
template <>
__device__ __forceinline__ unsigned long long bswap_word<unsigned long long>(unsigned long long v) {
    const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    return ((unsigned long long)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
}
template <typename U>
__global__ void k_bswap(U *__restrict__ x, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = bswap_word<U>(x[i]);
}
template <typename T>
__global__ void k_affine(T *__restrict__ x, size_t n, T scale, T offset) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = x[i] * scale + offset;
}
static unsigned stream_grid(size_t n) {  // enough CTAs to fill the device, grid-stride beyond
    const size_t want = (n + 1023) / 1024;
    return (unsigned)std::max<size_t>(1, std::min<size_t>(want, (size_t)148 * 16));
}

void bswap_device(mprg_ctx *ctx, void *x, size_t count, size_t elem, cudaStream_t st) {
    if (count == 0) return;
    if (elem == 4) k_bswap<uint32_t><<<stream_grid(count), 256, 0, st>>>((uint32_t *)x, count);
    else k_bswap<unsigned long long><<<stream_grid(count), 256, 0, st>>>((unsigned long long *)x, count);
    ctx->launches++;
    MPRG_CUDA(cudaGetLastError());
}

void post_affine_device(mprg_ctx *ctx, void *x, size_t count, int dtype, double scale, double offset) {
    if (count == 0) return;
    // fp32 fields: the reference computes `dum3d - 300.0` / `dum3dp1*9.81` in R8 and the NetCDF layer rounds to
    // NF90_FLOAT; one fp32 multiply-add of fp32 data differs from that by at most one rounding (<= 6e-8 relative)
    if (dtype == MPRG_F32) k_affine<float><<<stream_grid(count), 256, 0, ctx->stream>>>((float *)x, count, (float)scale, (float)offset);
    else k_affine<double><<<stream_grid(count), 256, 0, ctx->stream>>>((double *)x, count, scale, offset);
    ctx->launches++;
    MPRG_CUDA(cudaGetLastError());
}

void post_midlevels_device(mprg_ctx *ctx, int64_t n, int32_t nlev, int dtype, const void *x, void *mid) {
    if (n <= 0 || nlev < 2) return;
    const unsigned g = (unsigned)((n + 255) / 256);
    if (dtype == MPRG_F32) k_midlevels<float><<<g, 256, 0, ctx->stream>>>((const float *)x, (float *)mid, n, nlev);
    else k_midlevels<double><<<g, 256, 0, ctx->stream>>>((const double *)x, (double *)mid, n, nlev);
    ctx->launches++;
    MPRG_CUDA(cudaGetLastError());
}

void post_ptop_device(mprg_ctx *ctx, int64_t n, int32_t nlev, int dtype, const void *x, double *out2_dev) {
    DevBuf<long long> keys(2);
    const double init[2] = {-1.0e300, 1.0e300};
    long long hk[2];
    for (int k = 0; k < 2; ++k) {
        long long b;
        memcpy(&b, &init[k], 8);
        hk[k] = b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
    }
    poke(ctx, keys.p, hk, sizeof hk);
    if (n > 0 && nlev > 0) {
        const unsigned g = (unsigned)((n + 255) / 256);
        if (dtype == MPRG_F32) k_ptop<float><<<g, 256, 0, ctx->stream>>>((const float *)x, n, nlev, keys.p);
        else k_ptop<double><<<g, 256, 0, ctx->stream>>>((const double *)x, n, nlev, keys.p);
        ctx->launches++;
    }
    k_ptop_finish<<<1, 1, 0, ctx->stream>>>(keys.p, out2_dev);
    ctx->launches++;
    MPRG_CUDA(cudaGetLastError());
    MPRG_CUDA(cudaStreamSynchronize(ctx->stream));  // keys is freed on return; results are read by the caller
}

}  // namespace mprg
