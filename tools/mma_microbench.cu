// Microbenchmark (not part of the product): issue rate / duration of tcgen05.mma kind::i8 with
// shared-memory operands in the no-swizzle K-major layout, as a function of N and issue order.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_microbench tools/mma_microbench.cu
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

// variant: N columns per MMA, PL planes (accumulators), order 0 = plane-major, 1 = k-major over all planes
template <int N, int PL, int ORDER>
__global__ void __launch_bounds__(64, 1) bench(int tiles, long long *out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    constexpr int KS = 7, Fp = 224;
    uint8_t *sA = smem;                  // [6][128*Fp]
    uint8_t *sB = smem + 6 * 128 * Fp;   // [N*Fp]
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
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
    long long t0 = 0, t1 = 0;
    if (warp == 1) {
        const uint32_t idesc = make_idesc(128, N);
        const uint64_t da0 = make_desc(smem_u32(sA), 128 * 16, 128), db0 = make_desc(smem_u32(sB), N * 16, 128);
        const uint32_t a_plane = (128 * Fp) >> 4, a_k = (2 * 128 * 16) >> 4, b_k = (2 * N * 16) >> 4;
        t0 = clock64();
        for (int t = 0; t < tiles; ++t) {
            if (elect_one()) {
                if (ORDER == 0) {
#pragma unroll
                    for (int j = 0; j < PL; ++j)
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks)
                            mma_i8(tm + j * N, da0 + j * a_plane + ks * a_k, db0 + ks * b_k, idesc, ks > 0);
                } else {
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                        for (int j = 0; j < PL; ++j)
                            mma_i8(tm + j * N, da0 + j * a_plane + ks * a_k, db0 + ks * b_k, idesc, ks > 0);
                }
            }
            __syncwarp();
        }
        if (elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        t1 = clock64();
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

template <int N, int PL, int ORDER>
void run(const char *name, long long *d_out)
{
    const int tiles = 2000;
    size_t smem = 6 * 128 * 224 + 256 * 224 + 1024;
    cudaFuncSetAttribute(bench<N, PL, ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) bench<N, PL, ORDER><<<148, 64, smem>>>(tiles, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, d_out, sizeof(cyc), cudaMemcpyDeviceToHost);
    const double per_tile = (double)cyc / tiles, mmas = PL * 7.0;
    printf("%-28s N=%3d planes=%d: %8.1f clk/tile  %6.1f clk/MMA  %6.3f clk/(128 x column x plane-group) err=%d\n", name, N, PL,
           per_tile, per_tile / mmas, per_tile / N, (int)e);
}

int main()
{
    long long *d_out;
    cudaMalloc(&d_out, 64);
    run<64, 6, 0>("plane-major", d_out);
    run<64, 6, 1>("k-major (6 planes)", d_out);
    run<64, 2, 1>("k-major (2 planes)", d_out);
    run<32, 6, 0>("plane-major", d_out);
    run<32, 6, 1>("k-major", d_out);
    run<128, 3, 0>("plane-major", d_out);
    run<128, 3, 1>("k-major", d_out);
    run<256, 1, 0>("single plane", d_out);
    run<256, 2, 1>("k-major", d_out);
    run<16, 6, 1>("k-major", d_out);
    return 0;
}
