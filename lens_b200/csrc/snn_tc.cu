// K3 (tensor-core path): output layer  x2 = W_out . s1  on tcgen05 with EXACT integer arithmetic,
// fused with the IAF#2 scan, spike counting and the per-query similarity rows.
//
// Replaces the second F.linear + IAFSqueeze + spikes.sum(0) of lens/run_model.py:145,238-239
// (sinabs forward) for many independent streams.
//
// Exactness on tensor cores ("digit planes"): every output weight is a signed fixed-point integer m (snn.cu,
// |m| < 2^46 before the row's common trailing zero bits are shifted into its scale) = sum_j d_j 256^j with
// balanced digits d_j in [-128, 127], j < 6.  For each digit plane j one int8 x int8 -> int32 MMA
// (tcgen05.mma kind::i8) computes
//     P_j[place][step] = sum_k d_j[place][k] * s1[step][k]          (|P_j| < 2^22, exact)
// and the epilogue recombines X = sum_j P_j 256^j in int64 and rounds ONCE to fp32 (cvt.rn.f32.s64); the
// row's 2^q is applied by the scan's first FMA (exact product): bit-identical to the event-driven kernel and
// to the CPU oracle.  Place tiles whose top plane is all zero (npl = 5: rows whose weights span <= 15 binary
// orders of magnitude) skip that plane's operand loads, MMAs and drains.
//
// Mapping: M = 128 places (TMEM lanes), N = 64 columns = up to 32 consecutive timesteps of TWO streams
// (column 2n + s = step n of stream s; chunks never straddle a query), K = F padded to 32.  Time runs along
// the columns, so a thread owns one place and scans the columns serially with the membrane potentials and
// spike counts in registers.  The digit planes of the current place tile (<= 6 x 128 x Fp bytes, canonical
// no-swizzle K-major core-matrix layout) sit in shared memory; hidden-spike pair tiles (64 x Fp bytes, written
// by the hidden-layer variant of this kernel directly in the canonical layout) stream in through a
// cp.async.bulk (TMA) ring.  TMEM holds the 6 x 64-column int32 accumulator set plus two 64-column fp32
// exchange buffers (512 columns in total).
// Work = (block of <= 16 stream pairs, place tile) super-items, tile fastest, dealt round-robin to the persistent
// CTAs: CTAs running together share pair blocks, so spike tiles come from HBM once and then from L2.
//
// Four warpgroups (512 threads), registers redistributed with setmaxnreg (56 / 168 / 168 / 120):
//   control  warp 0 = TMA producer; warps 1-3 = MMA issuers, one per plane pair (warp 1 also allocates
//            TMEM): the ~450 cycles a warp spends per pair in tcgen05.commit and in waiting for the pair's
//            accumulators to come back overlap the other two warps' MMAs,
//   drain    warps 4-7 (columns 0..31 = steps 0..15 of both streams) and 8-11 (columns 32..63): tcgen05.ld the
//            accumulators plane pair by plane pair with register double buffering (each pair goes back to
//            its MMA warp as soon as its last load has landed), recombine X = sum_j P_j 256^j in int64,
//            round once to fp32, tcgen05.st the results into an exchange buffer,
//   scan     warps 12-15: tcgen05.ld the 64 results of every place and run the serial IAF recurrence of BOTH
//            streams of the pair as one chain of packed f32x2 operations (output layer: FFMA2 -> {FSET, FMNMX}
//            -> FADD2 -> FADD2 per step, spikes counted as steps minus sum of (a < 1)); the hidden-layer variant
//            (7-operation chain with exact multi-spike counts) scans the tiles of two stream pairs together.
// Config 3 (P = 10 000, 5-plane tiles): ~1970 cycles per tile against 35 MMAs x 48 cycles = 1680 of
// shared-memory operand fetch (ncu: that pipe 88 % busy); drain and scan are equally close to their limits
// (profiles/r02_tc_phase_profile.md).  -DLENS_TC_PROFILE builds an instrumented variant (phase clocks + a
// Gantt chart of three tiles).
#include "snn.cuh"

#include <algorithm>
#include <cstdio>
#include <vector>

namespace lens {

namespace tc {

constexpr int kM = 128;                 // places per CTA tile
constexpr int kN = kTileRows;           // TMEM columns per plane: 2 streams x 32 timesteps
constexpr int kStages = 3;              // hidden-spike pair tiles in flight
constexpr int kTmemCols = 512;          // 6 planes x 64 accumulator columns + 2 x 64 exchange columns
constexpr int kXCol = kPlanes * kN;     // first exchange column
constexpr int kDrainGroups = 2;         // drain warpgroups; each owns kN / kDrainGroups accumulator columns
constexpr int kNd = kN / kDrainGroups;  // (= the 32 steps of one stream)
constexpr int kDrainWarp0 = 4, kScanWarp0 = 4 + 4 * kDrainGroups;   // warps 0-3 control, then drain, drain, scan
constexpr int kThreads = 128 * (2 + kDrainGroups);   // control + drain warpgroups + scan
constexpr int kRegsControl = 56;        // setmaxnreg budgets: 128 * (56 + 2 * 168 + 120) = 512 * 128
constexpr int kRegsDrain = 168;
constexpr int kRegsScan = 120;
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
// One lane of a fully converged warp (elect.sync): single-thread instructions (TMA, tcgen05.mma,
// tcgen05.commit) are issued under this predicate from warp-uniform code so that their operands
// stay in uniform registers.
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// Non-blocking probe.  A barrier poll takes ~150 cycles to return even when the phase is already
// complete; issued early, its latency hides behind independent work and mbar_wait_probed() is free
// in the common case.
__device__ __forceinline__ uint32_t mbar_probe(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait_probed(uint32_t probed, uint64_t *bar, uint32_t parity)
{
    if (!probed) mbar_wait(bar, parity);
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], int8 x int8 -> int32, M = 128, N = kN, K = 32
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
        "%14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// store 16 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&r)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
        "%14, %15, %16};"
        ::"r"(taddr), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]),
          "f"(r[8]), "f"(r[9]), "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16f(uint32_t taddr, float (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
        "%14, %15}, [%16];"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
          "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle shared-memory operand descriptor (cute::UMMA::SmemDescriptor, version 1):
// core matrix = 8 rows x 16 bytes stored contiguously (128 B); `sbo` = byte distance between
// 8-row groups, `lbo` = byte distance between the two 16-byte K chunks of one K=32 instruction.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}

// cute::UMMA::InstrDescriptor for kind::i8: D = S32, A = B = signed int8, both K-major.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N)
{
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Params {
    const int8_t *planes;   // [n_tiles][kPlanes][Fp/16][128][16]
    const int *npl;         // [n_tiles] digit planes that are not all zero in the tile (5 or 6)
    const int8_t *S1;       // [pairs][chunks][Fp/16][64][16]
    const float *scale;     // [P]
    float *v2;              // [nb][P] (offset to the first stream of the launch)
    float *counts;          // [nb][Q][P]
    uint8_t *out_steps;     // nullable [nb][steps][P] (hidden layer: [nb][steps][P] hidden spikes)
    int P, Fp, T, steps, chunks, nb, n_pairs, n_tiles;
    int pb_size, n_super;   // work = (pair block, place tile) super-items, tile fastest (see snn_tc_output)
    float thr, vmin;
    // hidden-layer variant (kHidden): the "places" are feature neurons and the result is their spike
    // raster, written as pair tiles for the output layer
    int8_t *S1_out;         // [pairs][chunks][out_Fp/16][64][16]
    int out_Fp;
    int64_t *overflow;      // spike counts above LENS_MAX_SPIKE
#ifdef LENS_TC_PROFILE
    long long *prof;        // [gridDim.x][24] phase clocks of warps 1, 4 and 12 (profiling builds only)
#endif
};

#ifdef LENS_TC_PROFILE
#define PROF_DECL long long prof_t = clock64(), prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PROF(i) do { const long long now_ = clock64(); prof_acc[i] += now_ - prof_t; prof_t = now_; } while (0)
#define GANTT(k) do { if (blockIdx.x == 0 && lane == 0 && (int)it >= 1000 && (int)it < 1003) p.prof[1024 * 24 + ((int)it - 1000) * 48 + (k)] = clock64(); } while (0)
#define PROF_FLUSH(base) do { if (lane == 0) for (int i_ = 0; i_ < 8; ++i_) p.prof[blockIdx.x * 24 + (base) + i_] = prof_acc[i_]; } while (0)
#else
#define PROF_DECL
#define PROF(i)
#define GANTT(k)
#define PROF_FLUSH(base)
#endif

__device__ __forceinline__ void bulk_s2g(void *gmem_dst, const void *smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ float fmax3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// One IAF#2 step on the exact contraction result.
// kUnit: thr == 1 and v_min == -1 (the reference's fixed configuration, lens/run_model.py:151-156).
// Then v / thr == v and s * thr == s exactly, (v > 0) * trunc(v / thr) is 0 below 1, 1 in [1, 2) and
// trunc(v) above, and t = v - s is exact (s is the integer part of v), so the clip
// relu(t - v_min) + v_min == relu(fl(v + (1 - s))) - 1 rounds exactly like the reference's three ops.
template <bool kUnit>
__device__ __forceinline__ float iaf_out(float &v, float x, float thr, float vmin)
{
    if (kUnit) {
        const float vv = __fadd_rn(v, x);
        float s = (vv >= 1.0f) ? 1.0f : 0.0f;
        float c = (vv >= 1.0f) ? 0.0f : 1.0f;          // 1 - s
        if (vv >= 2.0f) { s = truncf(vv); c = 1.0f - s; }   // rare: several spikes in one step
        v = __fadd_rn(fmaxf(__fadd_rn(vv, c), 0.0f), -1.0f);
        return s;
    }
    return iaf_step(v, x, thr, vmin);
}

// Fast IAF#2 scan of one tile for BOTH streams of the pair (x[2n + s] = step n of stream s; v, csum:
// .x = even stream, .y = odd stream), valid when thr == 1, v_min == -1 and no step fires twice.
// Per step and stream the reference computes a = v + x; s = (a >= 1); v' = relu((a - s) + 1) - 1.
// With c = 1 - s:  relu(fl(a + c)) == fl(max(a, -1) + c)  (both are 0 exactly when a < -1, because then
// c = 1), so the chain per step is FADD2 -> {FSET, FMNMX} -> FADD2 -> FADD2 on packed f32x2 pairs.
// csum accumulates c (steps WITHOUT a spike); amax = largest pre-spike potential of the tile, which
// decides afterwards whether some step left the fast path's domain (a >= 2).
template <bool kRagged>
__device__ __forceinline__ void scan_tile_unit(const float (&x)[kN], float scale, int nvalid, float2 &v, float2 &csum,
                                               float &amax)
{
    const float2 m1 = make_float2(-1.0f, -1.0f), sc = make_float2(scale, scale);
#pragma unroll
    for (int n = 0; n < kTileSteps; ++n) {
        if (!kRagged || n < nvalid) {
            // x arrives as the rounded integer sum; scale = 2^q, so x * scale is exact and the FMA rounds once,
            // exactly like fl(v + fl(x * scale))
            const float2 a = __ffma2_rn(make_float2(x[2 * n], x[2 * n + 1]), sc, v);
            const float2 c = make_float2(a.x >= 1.0f ? 0.0f : 1.0f, a.y >= 1.0f ? 0.0f : 1.0f);
            const float2 lo = make_float2(fmaxf(a.x, -1.0f), fmaxf(a.y, -1.0f));
            amax = fmax3(amax, a.x, a.y);
            csum = __fadd2_rn(csum, c);
            v = __fadd2_rn(__fadd2_rn(lo, c), m1);
        }
    }
}

// Fast IAF#1 scan (hidden layer): several spikes per step do occur (~3e-4 of neuron-steps), so the count
// is computed exactly on the ALU in every step: for 0 <= m < 2^23, t = fl_rz(m + 2^23) is 2^23 + trunc(m),
// its low byte is the spike count s itself, and (2^23 + 1) - t = 1 - s.  v - s is exact, so
// relu(fl(a + (1 - s))) - 1 rounds like the reference's three operations.  The spike bytes go to the
// staging tile rows 2n (even stream) and 2n + 1 (odd stream) of this lane's k-chunk.
template <bool kRagged>
__device__ __forceinline__ void scan_tile_hidden(const float (&x)[kN], float scale, int nvalid, float2 &v, float &amax,
                                                 uint8_t *stage_out)
{
    const float2 m1 = make_float2(-1.0f, -1.0f), big = make_float2(8388608.0f, 8388608.0f);
    const float2 big1 = make_float2(8388609.0f, 8388609.0f), sc = make_float2(scale, scale);
#pragma unroll
    for (int n = 0; n < kTileSteps; ++n) {
        if (!kRagged || n < nvalid) {
            const float2 a = __ffma2_rn(make_float2(x[2 * n], x[2 * n + 1]), sc, v);   // exact product, one rounding
            const float2 ta = __fadd2_rz(make_float2(fmaxf(a.x, 0.0f), fmaxf(a.y, 0.0f)), big);
            amax = fmax3(amax, a.x, a.y);
            const float2 r = __fadd2_rn(a, __ffma2_rn(ta, m1, big1));       // a + (1 - s)
            v = __fadd2_rn(make_float2(fmaxf(r.x, 0.0f), fmaxf(r.y, 0.0f)), m1);
            stage_out[(2 * n) * 16] = (uint8_t)__float_as_uint(ta.x);
            stage_out[(2 * n + 1) * 16] = (uint8_t)__float_as_uint(ta.y);
        }
    }
}

// The hidden-layer scan for TWO stream pairs at once (tiles of pairs A and B that cover the same chunk): the scan
// warp then advances two independent packed chains, which hides the latency of the 7-operation chain step that a
// single pair leaves exposed (the recurrence is serial in time, so the parallelism has to come from streams).
// One call covers 16 steps: xa / xb = 32 exchange columns of A / B (column 2n + s = step n0 + n of stream s).
template <bool kRagged>
__device__ __forceinline__ void scan_half2_hidden(const float (&xa)[32], const float (&xb)[32], float scale, int n0,
                                                  int nvalid, float2 &va, float2 &vb, float &amax, uint8_t *out_a,
                                                  uint8_t *out_b)
{
    const float2 m1 = make_float2(-1.0f, -1.0f), big = make_float2(8388608.0f, 8388608.0f);
    const float2 big1 = make_float2(8388609.0f, 8388609.0f), sc = make_float2(scale, scale);
#pragma unroll
    for (int n = 0; n < kTileSteps / 2; ++n) {
        if (!kRagged || n0 + n < nvalid) {
            const float2 a = __ffma2_rn(make_float2(xa[2 * n], xa[2 * n + 1]), sc, va);
            const float2 b = __ffma2_rn(make_float2(xb[2 * n], xb[2 * n + 1]), sc, vb);
            const float2 ta = __fadd2_rz(make_float2(fmaxf(a.x, 0.0f), fmaxf(a.y, 0.0f)), big);
            const float2 tb = __fadd2_rz(make_float2(fmaxf(b.x, 0.0f), fmaxf(b.y, 0.0f)), big);
            amax = fmax3(fmax3(amax, a.x, a.y), b.x, b.y);
            const float2 ra = __fadd2_rn(a, __ffma2_rn(ta, m1, big1));
            const float2 rb = __fadd2_rn(b, __ffma2_rn(tb, m1, big1));
            va = __fadd2_rn(make_float2(fmaxf(ra.x, 0.0f), fmaxf(ra.y, 0.0f)), m1);
            vb = __fadd2_rn(make_float2(fmaxf(rb.x, 0.0f), fmaxf(rb.y, 0.0f)), m1);
            const int row = 2 * (n0 + n);
            out_a[row * 16] = (uint8_t)__float_as_uint(ta.x);
            out_a[(row + 1) * 16] = (uint8_t)__float_as_uint(ta.y);
            out_b[row * 16] = (uint8_t)__float_as_uint(tb.x);
            out_b[(row + 1) * 16] = (uint8_t)__float_as_uint(tb.y);
        }
    }
}

// kKSteps = Fp / 32 as a compile-time constant (0 = generic runtime loop): with it the MMAs of a tile
// are straight-line code whose descriptors are constant offsets from two uniform bases.
// kHidden: feature layer (IAF#1) instead of output layer (IAF#2): spikes leave as pair tiles, no counts.
template <bool kUnitThr, bool kDebug, int kKSteps, bool kHidden>
__global__ void __launch_bounds__(kThreads, 1) output_tc_kernel(Params p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int Fp = p.Fp;
    const int ksteps = kKSteps > 0 ? kKSteps : Fp / 32;
    const uint32_t plane_bytes = (uint32_t)kM * Fp;          // 128 rows x Fp bytes
    const uint32_t tile_bytes = (uint32_t)kN * Fp;           // 64 rows x Fp bytes
    uint8_t *sA = smem;                                      // [kPlanes][plane_bytes]
    uint8_t *sB = sA + kPlanes * plane_bytes;                // [kStages][tile_bytes]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + kStages * tile_bytes);
    uint64_t *a_full = bars;                                 // [1]
    uint64_t *b_full = bars + 1;                             // [kStages]
    uint64_t *b_empty = b_full + kStages;                    // [kStages]
    uint64_t *acc_full = b_empty + kStages;                  // [3] one per plane pair
    uint64_t *acc_empty = acc_full + 3;                      // [3]
    uint64_t *x_full = acc_empty + 3;                        // [2] exchange buffers drain -> scan
    uint64_t *x_empty = x_full + 2;                          // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(x_empty + 2);
    uint8_t *sOut = reinterpret_cast<uint8_t *>(bars) + 256;  // [2][8][64][16] spike staging (kHidden only)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // The hidden layer takes the stream pairs of a block two at a time with their chunks interleaved (A0 B0 A1 B1 ...)
    // and scans the two tiles of a chunk together: its IAF#1 chain is 7 dependent operations per step and the scan
    // warpgroup is what bounds that kernel, so two independent chains per thread pay.  The output layer does not:
    // there the drain warpgroups bound the tile and a faster, burstier scan only takes issue slots from them
    // (measured: -7 %, profiles/r02_tc_phase_profile.md).
    constexpr int kPairStep = kHidden ? 2 : 1;
    // Work = super-items (pair block, place tile), tile fastest, dealt round-robin to the persistent CTAs:
    // the CTAs running at the same time work on the same few pair blocks with different place tiles, so a
    // hidden-spike tile is fetched from HBM once and then served from L2 to the other place tiles; a CTA
    // keeps its digit planes for the pb_size stream pairs of the block.

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        for (int i = 0; i < kStages; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
        for (int i = 0; i < 3; ++i) {
            mbar_init(acc_full + i, 1);
            mbar_init(acc_empty + i, 4 * kDrainGroups);  // one arrival per drain warp
        }
        for (int i = 0; i < 2; ++i) { mbar_init(x_full + i, 4 * kDrainGroups); mbar_init(x_empty + i, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM allocation (whole warp), address lands in shared memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsControl));
    if (warp == 0) {
        // ===================== TMA producer (warp-uniform, one elected lane issues) ==============
        uint32_t it = 0;
        bool first = true;
        for (int u = blockIdx.x; u < p.n_super; u += gridDim.x) {
            const int pb = u / p.n_tiles, tile = u - pb * p.n_tiles;
            const int pr0 = pb * p.pb_size, pr1 = min(pr0 + p.pb_size, p.n_pairs);
            const uint32_t a_bytes = (uint32_t)__ldg(p.npl + tile) * plane_bytes;
            // new place tile: every MMA that reads the old planes must have retired
            if (!first)
                for (uint32_t j = it > (uint32_t)kStages ? it - kStages : 0; j < it; ++j)
                    mbar_wait(b_empty + j % kStages, (j / kStages) & 1);
            first = false;
            if (elect_one()) {
                mbar_expect_tx(a_full, a_bytes);
                const int8_t *src = p.planes + (size_t)tile * kPlanes * plane_bytes;
                for (uint32_t o = 0; o < a_bytes; o += plane_bytes)
                    bulk_g2s(sA + o, src + o, plane_bytes, a_full);
            }
            __syncwarp();
            for (int pa = pr0; pa < pr1; pa += kPairStep) {
                const int npair = min(kPairStep, pr1 - pa);
                for (int c = 0; c < p.chunks; ++c) {
                    for (int j = 0; j < npair; ++j, ++it) {
                        const int8_t *src = p.S1 + ((size_t)(pa + j) * p.chunks + c) * tile_bytes;
                        const uint32_t stage = it % kStages, phase = (it / kStages) & 1;
                        mbar_wait(b_empty + stage, phase ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx(b_full + stage, tile_bytes);
                            bulk_g2s(sB + stage * tile_bytes, src, tile_bytes, b_full + stage);
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else {
        // ===================== MMA issuers: warp 1 + g owns plane pair g (one elected lane issues) =========
        // Three issuing warps instead of one: the commit of a pair and the wait for its accumulators to be
        // drained cost the issuing warp ~450 cycles per pair, during which a single issuer cannot feed the
        // tensor pipe; with one warp per pair those latencies overlap the other pairs' MMAs.
        const int g = warp - 1;
        const uint32_t idesc = make_idesc(kM, kN);
        // descriptors differ only in the 14-bit start-address field: precompute, then add offsets
        const uint64_t da0 = make_desc(smem_u32(sA), kM * 16, 128);
        const uint64_t db0 = make_desc(smem_u32(sB), kN * 16, 128);
        const uint32_t a_plane = plane_bytes >> 4, a_kstep = (2 * kM * 16) >> 4;
        const uint32_t b_stage = tile_bytes >> 4, b_kstep = (2 * kN * 16) >> 4;
        // the two planes of the pair are interleaved along K so that consecutive MMAs accumulate into
        // different TMEM tiles (no back-to-back accumulator dependency)
        const uint64_t da_a = da0 + (uint32_t)(2 * g) * a_plane, da_b = da_a + a_plane;
        const uint32_t d_a = tmem_base + (2 * g) * kN, d_b = d_a + kN;
        uint32_t it = 0, a_phase = 0;
        PROF_DECL;
        for (int u = blockIdx.x; u < p.n_super; u += gridDim.x) {
            const int pb = u / p.n_tiles, tile = u - pb * p.n_tiles;
            const int n_it = (min(pb * p.pb_size + p.pb_size, p.n_pairs) - pb * p.pb_size) * p.chunks;
            const bool both = 2 * g + 1 < __ldg(p.npl + tile);    // the pair's upper plane is not all zero
            mbar_wait(a_full, a_phase); a_phase ^= 1;               // planes landed
            PROF(0);
            for (int i = 0; i < n_it; ++i, ++it) {
                const uint32_t stage = it % kStages, phase = (it / kStages) & 1;
                GANTT(0 + 5 * g);
                mbar_wait(b_full + stage, phase);           // spikes landed
                PROF(1);
                GANTT(1 + 5 * g);
                const uint64_t db_s = db0 + stage * b_stage;
                // The accumulator set is handed over in plane pairs: the MMAs of pair g of this tile start
                // as soon as the epilogue has drained pair g of the previous tile, and the epilogue starts
                // draining pair g while the other pairs are still being computed.
                mbar_wait(acc_empty + g, (it & 1) ^ 1);
                tc_fence_after();
                PROF(2);
                GANTT(2 + 5 * g);
                if (elect_one()) {
                    if (both) {
                        if (kKSteps > 0) {
#pragma unroll
                            for (int ks = 0; ks < kKSteps; ++ks) {
                                mma_i8(d_a, da_a + ks * a_kstep, db_s + ks * b_kstep, idesc, ks > 0 ? 1u : 0u);
                                mma_i8(d_b, da_b + ks * a_kstep, db_s + ks * b_kstep, idesc, ks > 0 ? 1u : 0u);
                            }
                        } else {
                            for (int ks = 0; ks < ksteps; ++ks) {
                                mma_i8(d_a, da_a + ks * a_kstep, db_s + ks * b_kstep, idesc, ks > 0 ? 1u : 0u);
                                mma_i8(d_b, da_b + ks * a_kstep, db_s + ks * b_kstep, idesc, ks > 0 ? 1u : 0u);
                            }
                        }
                    } else {
                        if (kKSteps > 0) {
#pragma unroll
                            for (int ks = 0; ks < kKSteps; ++ks)
                                mma_i8(d_a, da_a + ks * a_kstep, db_s + ks * b_kstep, idesc, ks > 0 ? 1u : 0u);
                        } else {
                            for (int ks = 0; ks < ksteps; ++ks)
                                mma_i8(d_a, da_a + ks * a_kstep, db_s + ks * b_kstep, idesc, ks > 0 ? 1u : 0u);
                        }
                    }
                    GANTT(3 + 5 * g);
                    tc_commit(acc_full + g);               // this pair is ready for the drain warps
                    // (the shared-memory slot of the spike tile is released by the drain warpgroup when it
                    //  has seen all three pairs complete)
                }
                __syncwarp();
                PROF(3);
                GANTT(4 + 5 * g);
            }
        }
        if (warp == 1) PROF_FLUSH(0);
    }
    } else if (warp >= kDrainWarp0 && warp < kDrainWarp0 + 4 * kDrainGroups) {
        // ===================== drain warpgroups: TMEM accumulators -> exact fp32 contraction results =========
        // warpgroup dw owns columns [dw * kNd, (dw + 1) * kNd) of every plane (16 steps of both streams)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsDrain));
        const int quarter = warp & 3;                          // TMEM lanes this warp may touch
        const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((warp - kDrainWarp0) >> 2) * kNd;
        int it = 0;
        uint32_t probe0 = 0;
        PROF_DECL;
        for (int u = blockIdx.x; u < p.n_super; u += gridDim.x) {
        const int pb = u / p.n_tiles, tile = u - pb * p.n_tiles;
        const int n_it = (min(pb * p.pb_size + p.pb_size, p.n_pairs) - pb * p.pb_size) * p.chunks;
        const bool top2 = __ldg(p.npl + tile) > 5;             // plane 5 present
        for (int i = 0; i < n_it; ++i, ++it) {
            // X = sum_j P_j 256^j as (xh:xl) for the kNd columns of this lane.  The (pair, 16-column) loads
            // are double-buffered in registers: the tcgen05.ld of step s + 1 is in flight while step s is
            // folded into X, so only the first load's latency is exposed.
            int32_t xl[kNd], xh[kNd];
            int32_t buf[2][32];
            uint32_t probe = 0;
            const uint32_t xb = it & 1;
            constexpr int kH = kNd / 16, kSteps = 3 * kH;
            static_assert(kH == 2, "the drain schedule below assumes two 16-column halves per warpgroup");
            PROF(1);
            GANTT(16);
            mbar_wait_probed(probe0, acc_full + 0, it & 1);
            tc_fence_after();
            // in steady state the MMA warps are a tile ahead: poll the other two pairs now, so that the
            // polls' latency hides behind the first pair's loads and folding
            const uint32_t probe1 = mbar_probe(acc_full + 1, it & 1), probe2 = mbar_probe(acc_full + 2, it & 1);
            PROF(0);
            GANTT(17);
            tmem_ld16(tlane + 0 * kN, reinterpret_cast<int32_t(&)[16]>(buf[0][0]));
            tmem_ld16(tlane + 1 * kN, reinterpret_cast<int32_t(&)[16]>(buf[0][16]));
#pragma unroll
            for (int s = 0; s < kSteps; ++s) {
                const int g = s / kH, h = s % kH;
                PROF(1);
                tmem_ld_wait();                                  // buf[s & 1] has landed
                PROF(5);
                if (h == kH - 1) {
                    // every load of pair g is complete: hand the pair back to its MMA warp
                    tc_fence_before();
                    __syncwarp();
                    GANTT(18 + 2 * g);
                    if (lane == 0) {
                        mbar_arrive(acc_empty + g);
                        // all MMAs of this tile have retired: its spike tile may be overwritten by the TMA
                        if (g == 2 && warp == kDrainWarp0) mbar_arrive(b_empty + (uint32_t)it % kStages);
                    }
                }
                if (s + 1 < kSteps) {
                    const int g2 = (s + 1) / kH, h2 = (s + 1) % kH;
                    if (h2 == 0) {
                        PROF(1);
                        mbar_wait_probed(g2 == 1 ? probe1 : probe2, acc_full + g2, it & 1);
                        tc_fence_after();
                        PROF(4);
                        GANTT(17 + 2 * g2);
                    }
                    tmem_ld16(tlane + (2 * g2) * kN + h2 * 16, reinterpret_cast<int32_t(&)[16]>(buf[(s + 1) & 1][0]));
                    if (g2 < 2 || top2)
                        tmem_ld16(tlane + (2 * g2 + 1) * kN + h2 * 16, reinterpret_cast<int32_t(&)[16]>(buf[(s + 1) & 1][16]));
                    if (s + 1 == 2 * kH) probe = mbar_probe(x_empty + (it & 1), ((it >> 1) & 1) ^ 1);
                }
                if (g == 2 && !top2) {
#pragma unroll
                    for (int n = 0; n < 16; ++n) xh[h * 16 + n] += buf[s & 1][n];              // plane 4 alone
                } else {
#pragma unroll
                    for (int n = 0; n < 16; ++n) {
                        const int m = h * 16 + n;
                        const int32_t qq = buf[s & 1][16 + n] * 256 + buf[s & 1][n];          // |qq| < 2^30
                        if (g == 0) {
                            xl[m] = qq;
                        } else if (g == 1) {
                            const int64_t t = (int64_t)xl[m] + ((int64_t)qq << 16);
                            xl[m] = (int32_t)(uint32_t)t;
                            xh[m] = (int32_t)(t >> 32);
                        } else {
                            xh[m] += qq;
                        }
                    }
                }
            }
            // one rounding to fp32 (cvt.rn.f32.s64), hand over through TMEM; the exact power-of-two scale of the
            // row is applied by the scan's first operation (an FMA whose product is exact)
            PROF(1);
            mbar_wait_probed(probe, x_empty + xb, ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            PROF(2);
            GANTT(23);
            probe0 = mbar_probe(acc_full + 0, (it + 1) & 1);        // next tile's first pair
#pragma unroll
            for (int h = 0; h < kH; ++h) {
                float xf[16];
#pragma unroll
                for (int n = 0; n < 16; ++n) {
                    const int m = h * 16 + n;
                    xf[n] = __ll2float_rn((int64_t)(((uint64_t)(uint32_t)xh[m] << 32) | (uint32_t)xl[m]));
                }
                PROF(6);
                tmem_st16(tlane + kXCol + xb * kN + h * 16, xf);
            }
            tmem_st_wait();
            PROF(7);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(x_full + xb);
            PROF(3);
            GANTT(24);
        }
        }
        if (warp == kDrainWarp0) PROF_FLUSH(8);
    } else {
        // ===================== scan warpgroup: IAF recurrence of the streams, spikes out ===================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsScan));
        const int quarter = warp & 3;
        const float thr = p.thr, vmin = p.vmin;
        const int cpq = chunks_per_query(p.T);
        const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        constexpr bool kFast = kUnitThr && !kDebug;
        int it = 0;
        PROF_DECL;
        if (kHidden) {
        // ---- hidden layer (IAF#1): spike bytes leave as pair tiles for the output layer.  Tiles arrive
        // chunk-interleaved from two independent stream pairs A and B (kPairStep); on the fast path the two tiles of
        // a chunk are scanned TOGETHER, 16 steps at a time, as two independent packed chains.  A block's odd last
        // pair, debug output and a half tile in which a step left the fast path's domain go tile by tile / step by step.
        long long n_over = 0;
        for (int u = blockIdx.x; u < p.n_super; u += gridDim.x) {
        const int pb = u / p.n_tiles, tile = u - pb * p.n_tiles;
        const int pr0 = pb * p.pb_size, pr1 = min(pr0 + p.pb_size, p.n_pairs);
        const int place = tile * kM + quarter * 32 + lane;
        const bool live_place = place < p.P;
        const float scale = live_place ? p.scale[place] : 0.0f;
        // a finished tile: generic-proxy writes -> async proxy, then one lane ships the warp's two k-chunks
        // (contiguous in the pair-tile layout) with one bulk store
        auto ship = [&](int pr, int c, int itx) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                const int kcw = tile * 8 + quarter * 2, nkc = min(2, p.out_Fp / 16 - kcw);
                if (nkc > 0)
                    bulk_s2g(p.S1_out + ((size_t)pr * p.chunks + c) * ((size_t)kN * p.out_Fp) + (size_t)kcw * 1024,
                             sOut + (itx & 1) * 8192 + quarter * 2048, (uint32_t)nkc * 1024u);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        };
        for (int pa = pr0; pa < pr1; pa += kPairStep) {
            const int npair = min(kPairStep, pr1 - pa);
            // state of the (up to) two pairs: index 0 = pair A, 1 = pair B; .x = even stream, .y = odd stream
            float2 v[2];
            bool live[2][2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int b0 = 2 * (pa + j), b1 = b0 + 1;
                live[j][0] = live_place && j < npair && b0 < p.nb;
                live[j][1] = live_place && j < npair && b1 < p.nb;
                v[j] = make_float2(live[j][0] ? p.v2[(size_t)b0 * p.P + place] : 0.0f,
                                   live[j][1] ? p.v2[(size_t)b1 * p.P + place] : 0.0f);
            }
            int cq = 0, q = 0;                                   // chunk within the query, query index
            for (int c = 0; c < p.chunks; ++c) {
                // this chunk: steps [t_base, t_base + nvalid) of query q (chunks never straddle a query)
                const int t_base = q * p.T + cq * kTileSteps;
                const int nvalid = min(kTileSteps, min(p.T - cq * kTileSteps, p.steps - t_base));
                if (++cq == cpq) { cq = 0; ++q; }
                if (kFast && npair == 2) {
                    // ---------------- two tiles together ----------------
                    const int ita = it, itb = it + 1;
                    it += 2;
                    const uint32_t xa_col = tlane + kXCol + (uint32_t)(ita & 1) * kN, xb_col = tlane + kXCol + (uint32_t)(itb & 1) * kN;
                    // staging tiles [kc][row = 2n + stream][16] of A and B
                    uint8_t *out_a = sOut + (ita & 1) * 8192 + (quarter * 2 + (lane >> 4)) * 1024 + (lane & 15);
                    uint8_t *out_b = sOut + (itb & 1) * 8192 + (quarter * 2 + (lane >> 4)) * 1024 + (lane & 15);
                    PROF(1);
                    mbar_wait(x_full + (ita & 1), (ita >> 1) & 1);
                    mbar_wait(x_full + (itb & 1), (itb >> 1) & 1);
                    tc_fence_after();
                    PROF(0);
                    // both staging buffers are rewritten: this warp's two stores of the previous chunk must have been read
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        float xa[32], xb[32];
                        tmem_ld16f(xa_col + half * 32, reinterpret_cast<float(&)[16]>(xa[0]));
                        tmem_ld16f(xa_col + half * 32 + 16, reinterpret_cast<float(&)[16]>(xa[16]));
                        tmem_ld16f(xb_col + half * 32, reinterpret_cast<float(&)[16]>(xb[0]));
                        tmem_ld16f(xb_col + half * 32 + 16, reinterpret_cast<float(&)[16]>(xb[16]));
                        tmem_ld_wait();
                        if (half == 1) {
                            // every column of both tiles is in registers: hand the exchange buffers back
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) { mbar_arrive(x_empty + (ita & 1)); mbar_arrive(x_empty + (itb & 1)); }
                        }
                        PROF(2);
                        const int n0 = half * (kTileSteps / 2);
                        if (n0 < nvalid) {
                            const float2 sva = v[0], svb = v[1];
                            float amax = -1.0f;
                            if (nvalid >= n0 + kTileSteps / 2) scan_half2_hidden<false>(xa, xb, scale, n0, nvalid, v[0], v[1], amax, out_a, out_b);
                            else scan_half2_hidden<true>(xa, xb, scale, n0, nvalid, v[0], v[1], amax, out_a, out_b);
                            if (amax >= 128.0f) {
                                // beyond LENS_MAX_SPIKE somewhere in these 16 steps: redo them one by one (clipped, counted)
                                v[0] = sva; v[1] = svb;
#pragma unroll
                                for (int j = 0; j < 2; ++j) {
#pragma unroll
                                    for (int sp = 0; sp < 2; ++sp) {
                                        float vv = sp ? v[j].y : v[j].x;
                                        uint8_t *out = j ? out_b : out_a;
#pragma unroll
                                        for (int n = 0; n < kTileSteps / 2; ++n) {
                                            if (n0 + n < nvalid) {
                                                float sk = iaf_out<kUnitThr>(vv, __fmul_rn(j ? xb[2 * n + sp] : xa[2 * n + sp], scale), thr, vmin);
                                                if (sk > (float)LENS_MAX_SPIKE) { sk = (float)LENS_MAX_SPIKE; if (live[j][sp]) ++n_over; }
                                                out[(2 * (n0 + n) + sp) * 16] = (uint8_t)sk;
                                            }
                                        }
                                        if (sp) v[j].y = vv; else v[j].x = vv;
                                    }
                                }
                            }
                        }
                        PROF(3);
                    }
                    for (int n = nvalid; n < kTileSteps; ++n) {            // rows of a ragged chunk that hold no step
                        out_a[(2 * n) * 16] = 0; out_a[(2 * n + 1) * 16] = 0;
                        out_b[(2 * n) * 16] = 0; out_b[(2 * n + 1) * 16] = 0;
                    }
                    ship(pa, c, ita);
                    ship(pa + 1, c, itb);
                } else {
                    // ---------------- tile by tile ----------------
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        if (j >= npair) break;
                        const int pr = pa + j;
                        const uint32_t xb = it & 1;
                        float x[kN];
                        PROF(1);
                        mbar_wait(x_full + xb, (it >> 1) & 1);
                        tc_fence_after();
                        PROF(0);
#pragma unroll
                        for (int h = 0; h < kN / 16; ++h)
                            tmem_ld16f(tlane + kXCol + xb * kN + h * 16, reinterpret_cast<float(&)[16]>(x[h * 16]));
                        tmem_ld_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(x_empty + xb);
                        PROF(2);
                        const int itx = it;
                        ++it;
                        uint8_t *stage_out = sOut + (itx & 1) * 8192 + (quarter * 2 + (lane >> 4)) * 1024 + (lane & 15);
                        // every scan warp ships its own 32 neurons, so the staging buffers need warp-level
                        // synchronisation only: this warp's store of two tiles ago has been read
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        __syncwarp();
                        bool redo = true;
                        if (kFast) {
                            const float2 saved = v[j];
                            float amax = -1.0f;
                            if (nvalid == kTileSteps) scan_tile_hidden<false>(x, scale, nvalid, v[j], amax, stage_out);
                            else scan_tile_hidden<true>(x, scale, nvalid, v[j], amax, stage_out);
                            redo = amax >= 128.0f;                       // beyond LENS_MAX_SPIKE: generic path
                            if (redo) v[j] = saved;
                        }
                        if (redo) {
#pragma unroll
                            for (int sp = 0; sp < 2; ++sp) {
                                const bool lv = live[j][sp];
                                const int b = 2 * pr + sp;
                                float vv = sp ? v[j].y : v[j].x;
#pragma unroll
                                for (int n = 0; n < kTileSteps; ++n) {
                                    if (n < nvalid) {
                                        float sk = iaf_out<kUnitThr>(vv, __fmul_rn(x[2 * n + sp], scale), thr, vmin);
                                        if (sk > (float)LENS_MAX_SPIKE) { sk = (float)LENS_MAX_SPIKE; if (lv) ++n_over; }
                                        stage_out[(2 * n + sp) * 16] = (uint8_t)sk;
                                        if (kDebug && lv) p.out_steps[((size_t)b * p.steps + t_base + n) * p.P + place] = (uint8_t)sk;
                                    }
                                }
                                if (sp) v[j].y = vv; else v[j].x = vv;
                            }
                        }
                        PROF(3);
                        for (int n = nvalid; n < kTileSteps; ++n) { stage_out[(2 * n) * 16] = 0; stage_out[(2 * n + 1) * 16] = 0; }
                        ship(pr, c, itx);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (live[j][0]) p.v2[(size_t)(2 * (pa + j)) * p.P + place] = v[j].x;
                if (live[j][1]) p.v2[(size_t)(2 * (pa + j) + 1) * p.P + place] = v[j].y;
            }
        }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (n_over) atomicAdd((unsigned long long *)p.overflow, (unsigned long long)n_over);
        } else {
        // ---- output layer (IAF#2): spike counts per query = the similarity rows.  One stream pair at a time: here
        // the drain warpgroups bound the tile, and a scan that runs two pairs together only takes issue slots from
        // them (measured: -7 %, profiles/r02_tc_phase_profile.md).
        const int Q = p.steps / p.T;
        for (int u = blockIdx.x; u < p.n_super; u += gridDim.x) {
        const int pb = u / p.n_tiles, tile = u - pb * p.n_tiles;
        const int pr0 = pb * p.pb_size, pr1 = min(pr0 + p.pb_size, p.n_pairs);
        const int place = tile * kM + quarter * 32 + lane;
        const bool live_place = place < p.P;
        const float scale = live_place ? p.scale[place] : 0.0f;
        for (int pr = pr0; pr < pr1; ++pr) {
            const int b0 = 2 * pr, b1 = 2 * pr + 1;
            const bool live0 = live_place && b0 < p.nb, live1 = live_place && b1 < p.nb;
            float2 v = make_float2(live0 ? p.v2[(size_t)b0 * p.P + place] : 0.0f,
                                   live1 ? p.v2[(size_t)b1 * p.P + place] : 0.0f);
            float2 count = make_float2(0.0f, 0.0f);
            int cq = 0, q = 0;                                   // chunk within the query, query index
            for (int c = 0; c < p.chunks; ++c, ++it) {
                const uint32_t xb = it & 1;
                float x[kN];
                PROF(1);
                GANTT(32);
                mbar_wait(x_full + xb, (it >> 1) & 1);
                tc_fence_after();
                PROF(0);
                GANTT(33);
#pragma unroll
                for (int h = 0; h < kN / 16; ++h)
                    tmem_ld16f(tlane + kXCol + xb * kN + h * 16, reinterpret_cast<float(&)[16]>(x[h * 16]));
                tmem_ld_wait();                                  // all four loads in flight together
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(x_empty + xb);
                PROF(2);
                GANTT(34);
                // this chunk: steps [t_base, t_base + nvalid) of query q (chunks never straddle a query)
                const int t_base = q * p.T + cq * kTileSteps;
                const int nvalid = min(kTileSteps, min(p.T - cq * kTileSteps, p.steps - t_base));
                const bool q_done = (cq + 1) * kTileSteps >= p.T;
                if (++cq == cpq) { cq = 0; }
                if (!live0 && !live1) { if (q_done) ++q; continue; }
                bool redo = true;
                if (kFast) {
                    // Fast path: both streams as one packed f32x2 chain (scan_tile_unit); a tile in which some step
                    // leaves the fast path's domain is redone step by step below.
                    const float2 saved = v;
                    float amax = -1.0f;
                    float2 csum = make_float2(0.0f, 0.0f);
                    if (nvalid == kTileSteps) scan_tile_unit<false>(x, scale, nvalid, v, csum, amax);
                    else scan_tile_unit<true>(x, scale, nvalid, v, csum, amax);
                    redo = amax >= 2.0f;                         // several spikes in one step
                    if (!redo) {
                        count.x += (float)nvalid - csum.x;       // small integers: exact
                        count.y += (float)nvalid - csum.y;
                    } else {
                        v = saved;
                    }
                }
                // generic: multi-spike steps, debug output, other thresholds
                if (redo) {
#pragma unroll
                    for (int sp = 0; sp < 2; ++sp) {
                        const bool live = sp ? live1 : live0;
                        const int b = sp ? b1 : b0;
                        float vv = sp ? v.y : v.x, cnt = 0.0f;
                        if (live) {
#pragma unroll
                            for (int n = 0; n < kTileSteps; ++n) {
                                if (n < nvalid) {
                                    const float sk = iaf_out<kUnitThr>(vv, __fmul_rn(x[2 * n + sp], scale), thr, vmin);
                                    if (kDebug) p.out_steps[((size_t)b * p.steps + t_base + n) * p.P + place] = (uint8_t)fminf(sk, 255.0f);
                                    cnt += sk;
                                }
                            }
                        }
                        if (sp) { v.y = vv; count.y += cnt; } else { v.x = vv; count.x += cnt; }
                    }
                }
                if (q_done) {
                    // last chunk of query q: its similarity row entries (lens/run_model.py:239)
                    if (q < Q) {
                        if (live0) p.counts[((size_t)b0 * Q + q) * p.P + place] = count.x;
                        if (live1) p.counts[((size_t)b1 * Q + q) * p.P + place] = count.y;
                    }
                    count = make_float2(0.0f, 0.0f);
                    ++q;
                }
                PROF(3);
                GANTT(35);
            }
            if (live0) p.v2[(size_t)b0 * p.P + place] = v.x;
            if (live1) p.v2[(size_t)b1 * p.P + place] = v.y;
        }
        }
        }
        if (warp == kScanWarp0) PROF_FLUSH(16);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

// Wo_fx [F][P] int64 fixed point -> balanced radix-256 digit planes in the canonical UMMA layout.
// grid = (P_tiles, Fp/16); block = 128 (one thread per row of the tile).
// npl[tile] (initialised to the minimum the kernel supports, 5) is raised to the number of planes that are
// not entirely zero in the tile: the MMAs, the operand loads and the accumulator drains of the missing
// top plane are skipped (after weights_to_fixed_kernel removed the row's common trailing zeros, rows whose
// weights span <= 15 binary orders of magnitude fit 5 planes).
__global__ void __launch_bounds__(128) planes_kernel(const int64_t *__restrict__ Wo_fx, int F, int P, int Fp,
                                                     int8_t *__restrict__ planes, int *__restrict__ npl)
{
    const int tile = blockIdx.x, kc = blockIdx.y, row = threadIdx.x;
    const int place = tile * kM + row;
    const size_t plane_bytes = (size_t)kM * Fp;
    int8_t dig[kPlanes][16];
    int top = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int k = kc * 16 + i;
        int64_t m = (place < P && k < F) ? Wo_fx[(size_t)k * P + place] : 0;
#pragma unroll
        for (int j = 0; j < kPlanes; ++j) {
            const int64_t d = ((m + 128) & 255) - 128;      // balanced digit in [-128, 127]
            dig[j][i] = (int8_t)d;
            if (d != 0) top = max(top, j + 1);
            m = (m - d) >> 8;                                // exact: m - d is a multiple of 256
        }
    }
    top = __reduce_max_sync(0xffffffffu, top);
    if ((threadIdx.x & 31) == 0 && top > 5) atomicMax(npl + tile, top);
#pragma unroll
    for (int j = 0; j < kPlanes; ++j) {
        int8_t *dst = planes + ((size_t)tile * kPlanes + j) * plane_bytes + ((size_t)kc * kM + row) * 16;
        *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(dig[j]);
    }
}

__global__ void fill_int_kernel(int *dst, int n, int value)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = value;
}

// Work decomposition of a launch: super-items = (block of pb_size stream pairs, place tile), tile fastest,
// dealt round-robin to `grid` persistent CTAs.  pb_size balances three things: the makespan (rounds x
// pb_size pair-items, every pair-item costs the same), the digit-plane reloads (once per super-item), and
// the L2 footprint of the pair blocks in flight (their spike tiles should be fetched from HBM only once).
static void schedule(Params &p, int sms, size_t pair_bytes, unsigned &grid, int min_m = 1)
{
    long long cost[17], best_cost = -1;
    for (int m = 1; m <= 16; ++m) {
        const long long n_super = (long long)ceil_div(p.n_pairs, m) * p.n_tiles;
        cost[m] = ceil_div64(n_super, std::min<long long>(sms, n_super)) * m;
        if (best_cost < 0 || cost[m] < best_cost) best_cost = cost[m];
    }
    int best_m = 0, smallest = 0;
    for (int m = 1; m <= 16; ++m) {
        if (cost[m] * 200 > best_cost * 201) continue;               // within 0.5 % of the best makespan
        if (!smallest) smallest = m;
        const long long n_super = (long long)ceil_div(p.n_pairs, m) * p.n_tiles;
        const double window = (double)ceil_div64(std::min<long long>(sms, n_super), p.n_tiles) * m * (double)pair_bytes;
        if (window <= 64.0 * 1024 * 1024) best_m = m;                // largest block whose window fits L2
    }
    if (!best_m) best_m = smallest;
    if (best_m < min_m) best_m = std::max(1, std::min(min_m, p.n_pairs));   // the hidden layer pairs up stream pairs
    p.pb_size = best_m;
    p.n_super = ceil_div(p.n_pairs, best_m) * p.n_tiles;
    grid = (unsigned)std::min<long long>(sms, p.n_super);
}

static size_t smem_bytes(int Fp, bool hidden = false)
{
    return (size_t)kPlanes * kM * Fp + (size_t)kStages * kN * Fp + 256 + (hidden ? 2 * 8192 : 0);
}

}  // namespace tc

// Input-spike pair tiles for the hidden-layer kernel: S0[pair][chunk][Ip/16][64][16] int8 with
// byte = (pixel > Uq[t][i]) -- the raster of lens/src/dataset.py:121 (see raster_thresholds_kernel).
// One thread per 16-byte row segment (16 inputs of one step of one stream).
__global__ void __launch_bounds__(256) raster_tiles_kernel(const uint8_t *__restrict__ pooled,
                                                           const uint8_t *__restrict__ Uq, int I, int Ip, int T, int Q,
                                                           int steps, int chunks, int nb, int n_pairs,
                                                           int8_t *__restrict__ S0)
{
    const int cpq = chunks_per_query(T);
    const int vec_per_tile = Ip / 16 * kTileRows;
    const long long total = (long long)n_pairs * chunks * vec_per_tile;
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < total;
         v += (long long)gridDim.x * blockDim.x) {
        const long long tile_id = v / vec_per_tile;
        const int r = (int)(v - tile_id * vec_per_tile);
        const int kc = r / kTileRows, row = r - kc * kTileRows;
        const int sp = row & 1, n = row >> 1;                      // row = 2 * step + stream
        const int pr = (int)(tile_id / chunks), c = (int)(tile_id - (long long)pr * chunks);
        const int b = 2 * pr + sp;
        const int q = c / cpq, t = (c - q * cpq) * kTileSteps + n; // chunks never straddle a query
        uint32_t out[4] = {0u, 0u, 0u, 0u};
        if (b < nb && t < T && q < Q) {
            const uint8_t *px = pooled + ((size_t)b * Q + q) * I, *uq = Uq + (size_t)t * I;
            if ((I & 3) == 0) {     // rows are 4-byte aligned: word loads + byte-wise SIMD compare
                const uint32_t *px4 = reinterpret_cast<const uint32_t *>(px) + kc * 4;
                const uint32_t *uq4 = reinterpret_cast<const uint32_t *>(uq) + kc * 4;
#pragma unroll
                for (int w = 0; w < 4; ++w)
                    if (kc * 16 + 4 * w < I) out[w] = __vcmpgtu4(__ldg(px4 + w), __ldg(uq4 + w)) & 0x01010101u;
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int i = kc * 16 + e;
                    const uint32_t spike = (i < I && __ldg(px + i) > __ldg(uq + i)) ? 1u : 0u;
                    out[e >> 2] |= spike << (8 * (e & 3));
                }
            }
        }
        reinterpret_cast<uint4 *>(S0)[v] = make_uint4(out[0], out[1], out[2], out[3]);
    }
}

bool snn_tc_supported(const SnnHandle *h)
{
    return tc::smem_bytes(h->Fp) <= 227 * 1024;
}

// hidden layer on tensor cores: binary raster, IAF#0 at rest, reference configuration (thr 1, v_min -1)
bool snn_tc_hidden_supported(const SnnHandle *h)
{
    const int Ip = (h->I + 31) & ~31;
    return h->Uq && h->thr == 1.0f && h->vmin == -1.0f && !h->v0_dirty && Ip <= 256 &&
           tc::smem_bytes(Ip, true) <= 227 * 1024;
}

int snn_tc_prepare(SnnHandle *h, cudaStream_t st)
{
    if (!h->Wo_planes) {
        h->P_tiles = ceil_div(h->P, tc::kM);
        const size_t bytes = (size_t)h->P_tiles * kPlanes * tc::kM * h->Fp;
        LENS_CUDA(cudaMalloc(&h->Wo_planes, bytes));
        LENS_CUDA(cudaMalloc(&h->Wo_npl, (size_t)h->P_tiles * sizeof(int)));
        tc::fill_int_kernel<<<ceil_div(h->P_tiles, 256), 256, 0, st>>>(h->Wo_npl, h->P_tiles, 5);
        LENS_LAUNCH_CHECK();
        dim3 grid(h->P_tiles, h->Fp / 16);
        tc::planes_kernel<<<grid, 128, 0, st>>>(h->Wo_fx, h->F, h->P, h->Fp, h->Wo_planes, h->Wo_npl);
        LENS_LAUNCH_CHECK();
    }
    if (!h->Wf_planes && snn_tc_hidden_supported(h)) {
        h->Ip = (h->I + 31) & ~31;
        h->F_tiles = ceil_div(h->F, tc::kM);
        const size_t bytes = (size_t)h->F_tiles * kPlanes * tc::kM * h->Ip;
        LENS_CUDA(cudaMalloc(&h->Wf_planes, bytes));
        LENS_CUDA(cudaMalloc(&h->Wf_npl, (size_t)h->F_tiles * sizeof(int)));
        tc::fill_int_kernel<<<ceil_div(h->F_tiles, 256), 256, 0, st>>>(h->Wf_npl, h->F_tiles, 5);
        LENS_LAUNCH_CHECK();
        dim3 grid(h->F_tiles, h->Ip / 16);
        tc::planes_kernel<<<grid, 128, 0, st>>>(h->Wf_fx, h->I, h->F, h->Ip, h->Wf_planes, h->Wf_npl);
        LENS_LAUNCH_CHECK();
    }
    return 0;
}

void snn_tc_release(SnnHandle *h)
{
    if (h->Wo_planes) cudaFree(h->Wo_planes);
    if (h->Wf_planes) cudaFree(h->Wf_planes);
    if (h->Wo_npl) cudaFree(h->Wo_npl);
    if (h->Wf_npl) cudaFree(h->Wf_npl);
    if (h->S0) cudaFree(h->S0);
    h->Wo_planes = nullptr; h->Wf_planes = nullptr; h->Wo_npl = nullptr; h->Wf_npl = nullptr;
    h->S0 = nullptr; h->S0_cap = 0;
}

#define LENS_TC_LAUNCH_K(U, D, K, H)                                                                                 \
    do {                                                                                                             \
        LENS_CUDA(cudaFuncSetAttribute(tc::output_tc_kernel<U, D, K, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       (int)smem));                                                                  \
        tc::output_tc_kernel<U, D, K, H><<<grid, tc::kThreads, smem, st>>>(p);                                       \
    } while (0)

#ifdef LENS_TC_PROFILE
struct ProfDump {
    long long *d; unsigned n; cudaStream_t st; long long iters; const char *name;
    static long long *buffer()
    {
        static long long *dev = nullptr;
        if (!dev) cudaMalloc(&dev, (1024 * 24 + 256) * sizeof(long long));
        return dev;
    }
    ProfDump(unsigned n_, cudaStream_t st_, long long iters_, const char *name_) : d(buffer()), n(n_), st(st_), iters(iters_), name(name_)
    {
        cudaMemsetAsync(d, 0, (1024 * 24 + 256) * sizeof(long long), st);
    }
    ~ProfDump()
    {
        std::vector<long long> hbuf(n * 24);
        cudaStreamSynchronize(st);
        cudaMemcpy(hbuf.data(), d, hbuf.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg[24] = {0};
        for (unsigned b = 0; b < n; ++b) for (int i = 0; i < 24; ++i) avg[i] += (double)hbuf[b * 24 + i] / n / iters;
        fprintf(stderr, "[tc-prof %s] clk per tile-iteration (avg over %u CTAs, %lld iters/CTA)\n", name, n, iters);
        fprintf(stderr, "  mma  : planes-wait %.0f  b_full-wait %.0f  acc_empty-wait %.0f  issue %.0f\n", avg[0], avg[1], avg[2], avg[3]);
        fprintf(stderr, "  drain: acc_full-wait %.0f  fold %.0f  x_empty-wait %.0f  fence+arrive %.0f  pair 1/2 wait %.0f  tcgen05.wait::ld %.0f\n",
                avg[8], avg[9], avg[10], avg[11], avg[12], avg[13]);
        fprintf(stderr, "  drain: conversions %.0f  tcgen05.st + wait::st %.0f\n", avg[14], avg[15]);
        fprintf(stderr, "  scan : x_full-wait %.0f  other %.0f  load %.0f  chain %.0f\n", avg[16], avg[17], avg[18], avg[19]);
        long long g_[144];
        cudaMemcpy(g_, d + 1024 * 24, sizeof(g_), cudaMemcpyDeviceToHost);
        const long long t0_ = g_[0];
        for (int r = 0; r < 3; ++r) {
            fprintf(stderr, "  gantt it=%d  mma:", 1000 + r);
            for (int k = 0; k <= 14; ++k) fprintf(stderr, " %lld", g_[r * 48 + k] - t0_);
            fprintf(stderr, "  | drain:");
            for (int k = 16; k <= 24; ++k) fprintf(stderr, " %lld", g_[r * 48 + k] - t0_);
            fprintf(stderr, "  | scan:");
            for (int k = 32; k <= 35; ++k) fprintf(stderr, " %lld", g_[r * 48 + k] - t0_);
            fprintf(stderr, "\n");
        }
    }
};
#endif

int snn_tc_output(SnnHandle *h, const int8_t *S1, int nb, int b0, int steps, float *counts,
                  uint8_t *out_steps, cudaStream_t st)
{
    tc::Params p{};
    p.planes = h->Wo_planes; p.S1 = S1; p.scale = h->Wo_scale;
    p.v2 = h->v2 + (size_t)b0 * h->P; p.counts = counts; p.out_steps = out_steps;
    p.P = h->P; p.Fp = h->Fp; p.T = h->T; p.steps = steps;
    p.chunks = n_chunks_of(steps, h->T); p.nb = nb; p.n_pairs = (nb + 1) / 2;
    p.npl = h->Wo_npl;
    p.thr = h->thr; p.vmin = h->vmin;
    const int sms = std::max(sm_count(), 1);
    p.n_tiles = h->P_tiles;
    const size_t smem = tc::smem_bytes(h->Fp);
    unsigned grid_x = 1;
    tc::schedule(p, sms, (size_t)p.chunks * s1_tile_bytes(h->Fp), grid_x);
    dim3 grid(grid_x);
#ifdef LENS_TC_PROFILE
    ProfDump prof_dump(grid.x, st, (long long)h->P_tiles * p.n_pairs * p.chunks / grid.x, "output");
    p.prof = prof_dump.d;
#endif
    LaunchTimer timer(h, st, 1);
    const bool unit = h->thr == 1.0f && h->vmin == -1.0f, dbg = out_steps != nullptr;
#define LENS_TC_LAUNCH(U, D)                                     \
    do {                                                         \
        if (h->Fp == 224) LENS_TC_LAUNCH_K(U, D, 7, false);      \
        else if (h->Fp == 64) LENS_TC_LAUNCH_K(U, D, 2, false);  \
        else LENS_TC_LAUNCH_K(U, D, 0, false);                   \
    } while (0)
    if (unit && !dbg) LENS_TC_LAUNCH(true, false);
    else if (unit && dbg) LENS_TC_LAUNCH(true, true);
    else if (!unit && !dbg) LENS_TC_LAUNCH(false, false);
    else LENS_TC_LAUNCH(false, true);
#undef LENS_TC_LAUNCH
    LENS_LAUNCH_CHECK();
    return 0;
}

// Feature layer on the tensor cores: raster tiles -> (W_feat digit planes) -> IAF#1 -> hidden-spike pair
// tiles in h->S1.  Same kernel as the output layer with the kHidden epilogue.
int snn_tc_hidden(SnnHandle *h, const uint8_t *pooled, int nb, int b0, int steps, uint8_t *hidden_steps,
                  cudaStream_t st)
{
    const int chunks = n_chunks_of(steps, h->T), n_pairs = (nb + 1) / 2;
    const size_t s0_bytes = (size_t)n_pairs * chunks * kTileRows * h->Ip;
    if (s0_bytes > h->S0_cap) {
        if (h->S0) LENS_CUDA(cudaFree(h->S0));
        h->S0 = nullptr; h->S0_cap = 0;
        LENS_CUDA(cudaMalloc(&h->S0, s0_bytes));
        h->S0_cap = s0_bytes;
    }
    {
        const long long vecs = (long long)(s0_bytes / 16);
        const int blocks = (int)std::min<long long>((vecs + 255) / 256, (long long)std::max(sm_count(), 1) * 16);
        LaunchTimer timer(h, st, 0);
        raster_tiles_kernel<<<blocks, 256, 0, st>>>(pooled, h->Uq, h->I, h->Ip, h->T, steps / h->T, steps, chunks, nb,
                                                    n_pairs, h->S0);
        LENS_LAUNCH_CHECK();
    }
    tc::Params p{};
    p.planes = h->Wf_planes; p.S1 = h->S0; p.scale = h->Wf_scale;
    p.v2 = h->v1 + (size_t)b0 * h->F; p.counts = nullptr; p.out_steps = hidden_steps;
    p.P = h->F; p.Fp = h->Ip; p.T = h->T; p.steps = steps; p.chunks = chunks; p.nb = nb; p.n_pairs = n_pairs;
    p.thr = h->thr; p.vmin = h->vmin;
    p.S1_out = h->S1; p.out_Fp = h->Fp; p.overflow = h->counters;
    const int sms = std::max(sm_count(), 1);
    p.n_tiles = h->F_tiles;
    p.npl = h->Wf_npl;
    const size_t smem = tc::smem_bytes(h->Ip, true);
    unsigned grid_x = 1;
    tc::schedule(p, sms, (size_t)chunks * s1_tile_bytes(h->Ip), grid_x, 2);
    dim3 grid(grid_x);
#ifdef LENS_TC_PROFILE
    ProfDump prof_dump(grid.x, st, (long long)h->F_tiles * p.n_pairs * p.chunks / grid.x, "hidden");
    p.prof = prof_dump.d;
#endif
    LaunchTimer timer(h, st, 0);
    const bool dbg = hidden_steps != nullptr;
    if (h->Ip == 128) { if (dbg) LENS_TC_LAUNCH_K(true, true, 4, true); else LENS_TC_LAUNCH_K(true, false, 4, true); }
    else if (h->Ip == 64) { if (dbg) LENS_TC_LAUNCH_K(true, true, 2, true); else LENS_TC_LAUNCH_K(true, false, 2, true); }
    else { if (dbg) LENS_TC_LAUNCH_K(true, true, 0, true); else LENS_TC_LAUNCH_K(true, false, 0, true); }
    LENS_LAUNCH_CHECK();
    return 0;
}
#undef LENS_TC_LAUNCH_K

}  // namespace lens
