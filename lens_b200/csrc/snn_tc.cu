// K3 (tensor-core path): output layer  x2 = W_out . s1  on tcgen05 with EXACT integer arithmetic,
// fused with the IAF#2 scan, spike counting and the per-query similarity rows.
//
// Replaces the second F.linear + IAFSqueeze + spikes.sum(0) of lens/run_model.py:145,238-239
// (sinabs forward) for many independent streams.
//
// Exactness on tensor cores ("digit planes"): every output weight is a 47-bit signed fixed-point
// integer m (snn.cu) = sum_j d_j 256^j with balanced digits d_j in [-128, 127], j < 6.  For each
// digit plane j one int8 x int8 -> int32 MMA (tcgen05.mma kind::i8) computes
//     P_j[place][step] = sum_k d_j[place][k] * s1[step][k]          (|P_j| < 2^22, exact)
// and the epilogue recombines X = sum_j P_j 256^j in int64, rounds ONCE to fp32 (cvt.rn.f32.s64)
// and scales by the row's 2^q: bit-identical to the event-driven kernel and to the CPU oracle.
//
// Mapping: M = 128 places (TMEM lanes), N = 32 consecutive timesteps of one stream (TMEM columns),
// K = F padded to 32.  Time runs along the columns, so each epilogue thread owns one place and scans
// its 32 columns serially with the membrane potential and spike count in registers -- the recurrence
// never leaves the register file for a whole stream.  A CTA owns one place tile for the whole
// launch: its 6 digit planes (6 x 128 x Fp bytes, canonical no-swizzle K-major core-matrix layout)
// stay resident in shared memory; hidden-spike tiles (32 x Fp bytes, written by feature_kernel
// directly in the canonical layout) stream in through a 4-stage cp.async.bulk (TMA) ring; two
// 6 x 32-column int32 accumulator sets in TMEM ping-pong between the MMA warp and the 4 epilogue warps.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
#include "snn.cuh"

#include <algorithm>

namespace lens {

namespace tc {

constexpr int kM = 128;                 // places per CTA tile
constexpr int kN = kTileSteps;          // timesteps per MMA (TMEM columns per plane)
constexpr int kStages = 4;              // hidden-spike tiles in flight
constexpr int kAccBufs = 2;             // accumulator sets in TMEM
constexpr int kTmemCols = 512;          // allocation (power of two >= kAccBufs * kPlanes * kN = 384)
constexpr int kThreads = 192;

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
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
    const int8_t *planes;   // [P_tiles][kPlanes][Fp/16][128][16]
    const int8_t *S1;       // [nb][chunks][Fp/16][32][16]
    const float *scale;     // [P]
    float *v2;              // [nb][P] (offset to the first stream of the launch)
    float *counts;          // [nb][Q][P]
    uint8_t *out_steps;     // nullable [nb][steps][P]
    int P, Fp, T, steps, chunks, nb, n_groups;
    float thr, vmin;
};

// One IAF#2 step on the exact contraction result.
template <bool kUnitThr>
__device__ __forceinline__ float iaf_out(float &v, float x, float thr, float vmin)
{
    if (kUnitThr) {
        float vv = __fadd_rn(v, x);
        float s = (vv > 0.0f) ? truncf(vv) : 0.0f;     // v / 1.0f == v exactly
        vv = __fsub_rn(vv, s);                         // s * 1.0f == s exactly
        v = __fadd_rn(fmaxf(__fsub_rn(vv, vmin), 0.0f), vmin);
        return s;
    }
    return iaf_step(v, x, thr, vmin);
}

template <bool kUnitThr>
__global__ void __launch_bounds__(kThreads, 1) output_tc_kernel(Params p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int Fp = p.Fp;
    const int ksteps = Fp / 32;
    const uint32_t plane_bytes = (uint32_t)kM * Fp;          // 128 rows x Fp bytes
    const uint32_t tile_bytes = (uint32_t)kN * Fp;           // 32 steps x Fp bytes
    uint8_t *sA = smem;                                      // [kPlanes][plane_bytes]
    uint8_t *sB = sA + kPlanes * plane_bytes;                // [kStages][tile_bytes]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + kStages * tile_bytes);
    uint64_t *a_full = bars;                                 // [1]
    uint64_t *b_full = bars + 1;                             // [kStages]
    uint64_t *b_empty = b_full + kStages;                    // [kStages]
    uint64_t *acc_full = b_empty + kStages;                  // [kAccBufs]
    uint64_t *acc_empty = acc_full + kAccBufs;               // [kAccBufs]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + kAccBufs);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;          // place tile
    const int group = blockIdx.y;         // stream group: streams group, group + n_groups, ...

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        for (int i = 0; i < kStages; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
        for (int i = 0; i < kAccBufs; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 4); }
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

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(a_full, kPlanes * plane_bytes);
            const int8_t *src = p.planes + (size_t)tile * kPlanes * plane_bytes;
            for (int j = 0; j < kPlanes; ++j)
                bulk_g2s(sA + j * plane_bytes, src + (size_t)j * plane_bytes, plane_bytes, a_full);
            uint32_t it = 0;
            for (int b = group; b < p.nb; b += p.n_groups) {
                const int8_t *sb = p.S1 + (size_t)b * p.chunks * tile_bytes;
                for (int c = 0; c < p.chunks; ++c, ++it) {
                    const uint32_t stage = it % kStages, phase = (it / kStages) & 1;
                    mbar_wait(b_empty + stage, phase ^ 1);
                    mbar_expect_tx(b_full + stage, tile_bytes);
                    bulk_g2s(sB + stage * tile_bytes, sb + (size_t)c * tile_bytes, tile_bytes, b_full + stage);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(kM, kN);
            const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
            mbar_wait(a_full, 0);
            uint32_t it = 0;
            for (int b = group; b < p.nb; b += p.n_groups) {
                for (int c = 0; c < p.chunks; ++c, ++it) {
                    const uint32_t stage = it % kStages, phase = (it / kStages) & 1;
                    const uint32_t buf = it % kAccBufs, aphase = (it / kAccBufs) & 1;
                    mbar_wait(acc_empty + buf, aphase ^ 1);     // epilogue drained this accumulator set
                    mbar_wait(b_full + stage, phase);           // spikes landed
                    tc_fence_after();
                    for (int j = 0; j < kPlanes; ++j) {
                        const uint32_t d = tmem_base + buf * (kPlanes * kN) + j * kN;
                        for (int ks = 0; ks < ksteps; ++ks) {
                            // K step = two 16-byte chunks: chunk stride = rows * 16 bytes
                            const uint64_t da = make_desc(a_base + j * plane_bytes + ks * 2 * (kM * 16), kM * 16, 128);
                            const uint64_t db = make_desc(b_base + stage * tile_bytes + ks * 2 * (kN * 16), kN * 16, 128);
                            mma_i8(d, da, db, idesc, ks > 0 ? 1u : 0u);
                        }
                    }
                    tc_commit(b_empty + stage);    // smem slot reusable once these MMAs retire
                    tc_commit(acc_full + buf);     // accumulators ready for the epilogue
                }
            }
        }
    } else {
        // ===================== epilogue: IAF#2 scan, spike counts =====================
        const int quarter = warp & 3;                          // TMEM lanes this warp may touch
        const int row = quarter * 32 + lane;
        const int place = tile * kM + row;
        const bool live = place < p.P;
        const float scale = live ? p.scale[place] : 0.0f;
        const float thr = p.thr, vmin = p.vmin;
        const int Q = p.steps / p.T;
        uint32_t it = 0;
        for (int b = group; b < p.nb; b += p.n_groups) {
            float v = live ? p.v2[(size_t)b * p.P + place] : 0.0f;
            float count = 0.0f;
            int t_in_q = 0, q = 0;
            for (int c = 0; c < p.chunks; ++c, ++it) {
                const uint32_t buf = it % kAccBufs, aphase = (it / kAccBufs) & 1;
                mbar_wait(acc_full + buf, aphase);
                tc_fence_after();
                const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * (kPlanes * kN);
                const int nvalid = min(kN, p.steps - c * kN);
#pragma unroll 1
                for (int half = 0; half < kN / 16; ++half) {
                    int32_t r[kPlanes][16];
#pragma unroll
                    for (int j = 0; j < kPlanes; ++j) tmem_ld16(tbase + j * kN + half * 16, r[j]);
                    tmem_ld_wait();
#pragma unroll
                    for (int n = 0; n < 16; ++n) {
                        const int step_in_chunk = half * 16 + n;
                        if (step_in_chunk < nvalid) {
                            // X = sum_j P_j 256^j, exactly, in 64 bits
                            const int32_t q0 = r[1][n] * 256 + r[0][n];
                            const int32_t q1 = r[3][n] * 256 + r[2][n];
                            const int32_t q2 = r[5][n] * 256 + r[4][n];
                            const int64_t X = (int64_t)q0 + ((int64_t)q1 << 16) + ((int64_t)q2 << 32);
                            const float x = __fmul_rn(__ll2float_rn(X), scale);
                            const float s = iaf_out<kUnitThr>(v, x, thr, vmin);
                            if (live) {
                                if (p.out_steps)
                                    p.out_steps[((size_t)b * p.steps + c * kN + step_in_chunk) * p.P + place] =
                                        (uint8_t)fminf(s, 255.0f);
                                count += s;
                                if (++t_in_q == p.T) {
                                    p.counts[((size_t)b * Q + q) * p.P + place] = count;
                                    count = 0.0f; t_in_q = 0; ++q;
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + buf);
            }
            if (live) p.v2[(size_t)b * p.P + place] = v;
        }
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
__global__ void __launch_bounds__(128) planes_kernel(const int64_t *__restrict__ Wo_fx, int F, int P, int Fp,
                                                     int8_t *__restrict__ planes)
{
    const int tile = blockIdx.x, kc = blockIdx.y, row = threadIdx.x;
    const int place = tile * kM + row;
    const size_t plane_bytes = (size_t)kM * Fp;
    int8_t dig[kPlanes][16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int k = kc * 16 + i;
        int64_t m = (place < P && k < F) ? Wo_fx[(size_t)k * P + place] : 0;
#pragma unroll
        for (int j = 0; j < kPlanes; ++j) {
            const int64_t d = ((m + 128) & 255) - 128;      // balanced digit in [-128, 127]
            dig[j][i] = (int8_t)d;
            m = (m - d) >> 8;                                // exact: m - d is a multiple of 256
        }
    }
#pragma unroll
    for (int j = 0; j < kPlanes; ++j) {
        int8_t *dst = planes + ((size_t)tile * kPlanes + j) * plane_bytes + ((size_t)kc * kM + row) * 16;
        *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(dig[j]);
    }
}

static size_t smem_bytes(int Fp)
{
    return (size_t)kPlanes * kM * Fp + (size_t)kStages * kN * Fp + (1 + 2 * kStages + 2 * kAccBufs) * 8 + 16;
}

}  // namespace tc

bool snn_tc_supported(const SnnHandle *h)
{
    return tc::smem_bytes(h->Fp) <= 227 * 1024;
}

int snn_tc_prepare(SnnHandle *h, cudaStream_t st)
{
    if (h->Wo_planes) return 0;
    h->P_tiles = ceil_div(h->P, tc::kM);
    const size_t bytes = (size_t)h->P_tiles * kPlanes * tc::kM * h->Fp;
    LENS_CUDA(cudaMalloc(&h->Wo_planes, bytes));
    dim3 grid(h->P_tiles, h->Fp / 16);
    tc::planes_kernel<<<grid, 128, 0, st>>>(h->Wo_fx, h->F, h->P, h->Fp, h->Wo_planes);
    LENS_LAUNCH_CHECK();
    return 0;
}

void snn_tc_release(SnnHandle *h)
{
    if (h->Wo_planes) cudaFree(h->Wo_planes);
    h->Wo_planes = nullptr;
}

int snn_tc_output(SnnHandle *h, const int8_t *S1, int nb, int b0, int steps, float *counts,
                  uint8_t *out_steps, cudaStream_t st)
{
    tc::Params p;
    p.planes = h->Wo_planes; p.S1 = S1; p.scale = h->Wo_scale;
    p.v2 = h->v2 + (size_t)b0 * h->P; p.counts = counts; p.out_steps = out_steps;
    p.P = h->P; p.Fp = h->Fp; p.T = h->T; p.steps = steps;
    p.chunks = ceil_div(steps, kTileSteps); p.nb = nb;
    p.thr = h->thr; p.vmin = h->vmin;
    const int sms = std::max(sm_count(), 1);
    p.n_groups = std::max(1, std::min(nb, sms / std::max(h->P_tiles, 1)));
    const size_t smem = tc::smem_bytes(h->Fp);
    dim3 grid(h->P_tiles, p.n_groups);
    LaunchTimer timer(h, st, 1);
    if (h->thr == 1.0f) {
        LENS_CUDA(cudaFuncSetAttribute(tc::output_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc::output_tc_kernel<true><<<grid, tc::kThreads, smem, st>>>(p);
    } else {
        LENS_CUDA(cudaFuncSetAttribute(tc::output_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc::output_tc_kernel<false><<<grid, tc::kThreads, smem, st>>>(p);
    }
    LENS_LAUNCH_CHECK();
    return 0;
}

}  // namespace lens
