// Microbenchmark 2 (not part of the product): N=64, 6 planes, pair-interleaved MMAs, with
//   MODE 0: MMA warp alone
//   MODE 1: + 4 drain warps free-running tcgen05.ld over the accumulator columns (TMEM port contention?)
//   MODE 2: + plane-pair full/empty barriers between MMA warp and drain warps (LDTM + arrive only)
//   MODE 3: MODE 2 + int combine (LEA/IMAD) per element like the real drain
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N)
{
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *b) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void ld16(uint32_t taddr, int32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}

template <int MODE>
__global__ void __launch_bounds__(192, 1) bench(int tiles, long long *out, int *sink)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[8];   // full[3], empty[3], done
    __shared__ uint32_t slot;
    constexpr int KS = 7, Fp = 224, N = 64;
    uint8_t *sA = smem, *sB = smem + 6 * 128 * Fp;
    uint64_t *full = bars, *empty = bars + 3, *done = bars + 6;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 3; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 4); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    if (warp == 1) {
        const uint32_t idesc = make_idesc(128, N);
        const uint64_t da0 = make_desc(smem_u32(sA), 128 * 16, 128), db0 = make_desc(smem_u32(sB), N * 16, 128);
        const uint32_t a_plane = (128 * Fp) >> 4, a_k = (2 * 128 * 16) >> 4, b_k = (2 * N * 16) >> 4;
        long long t0 = clock64();
        for (int t = 0; t < tiles; ++t) {
#pragma unroll 1
            for (int g = 0; g < 3; ++g) {
                if (MODE >= 2) { mbar_wait(empty + g, (t & 1) ^ 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        mma_i8(tm + (2 * g) * N, da0 + (2 * g) * a_plane + ks * a_k, db0 + ks * b_k, idesc, ks > 0);
                        mma_i8(tm + (2 * g + 1) * N, da0 + (2 * g + 1) * a_plane + ks * a_k, db0 + ks * b_k, idesc, ks > 0);
                    }
                    if (MODE >= 2) commit(full + g);
                }
                __syncwarp();
            }
        }
        if (elect_one()) commit(done);
        __syncwarp();
        mbar_wait(done, 0);
        long long t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
    } else if (warp >= 2 && MODE >= 1) {
        const uint32_t tl = tm + ((uint32_t)((warp & 3) * 32) << 16);
        int acc = 0;
        for (int t = 0; t < tiles; ++t) {
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                if (MODE >= 2) { mbar_wait(full + g, t & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    int32_t lo[16], hi[16];
                    ld16(tl + (2 * g) * N + h * 16, lo);
                    ld16(tl + (2 * g + 1) * N + h * 16, hi);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (MODE >= 3) {
#pragma unroll
                        for (int n = 0; n < 16; ++n) acc += hi[n] * 256 + lo[n];
                    } else {
                        acc += lo[0] + hi[15];
                    }
                }
                if (MODE >= 2) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty + g);
                }
            }
        }
        if (acc == 0x12345678) sink[0] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

template <int MODE>
void run(const char *name, long long *d_out, int *sink)
{
    const int tiles = 2000;
    size_t smem = 6 * 128 * 224 + 64 * 224 + 1024;
    cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) bench<MODE><<<148, 192, smem>>>(tiles, d_out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, d_out, sizeof(cyc), cudaMemcpyDeviceToHost);
    printf("mode %d %-46s %8.1f clk/tile err=%d\n", MODE, name, (double)cyc / tiles, (int)e);
}

int main()
{
    long long *d_out; int *sink;
    cudaMalloc(&d_out, 64); cudaMalloc(&sink, 64);
    run<0>("MMA warp alone", d_out, sink);
    run<1>("+ free-running LDTM (4 warps)", d_out, sink);
    run<2>("+ pair full/empty barriers", d_out, sink);
    run<3>("+ integer combine per element", d_out, sink);
    return 0;
}
