// Microbenchmark (not part of the product): dependent-issue latency of the ops in the IAF scan chain on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/lat_microbench tools/lat_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void lat(float *out, long long *cyc, float seed)
{
    float2 v = make_float2(seed, seed * 0.5f);
    const float2 x = make_float2(1e-3f, 2e-3f), m1 = make_float2(-1.f, -1.f);
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < 256; ++r) {
#pragma unroll
        for (int n = 0; n < 32; ++n) {
            if (MODE == 0) {            // packed chain: FADD2 -> {FSET, FMNMX} -> FADD2 -> FADD2
                const float2 a = __fadd2_rn(v, x);
                const float2 c = make_float2(a.x >= 1.0f ? 0.0f : 1.0f, a.y >= 1.0f ? 0.0f : 1.0f);
                const float2 lo = make_float2(fmaxf(a.x, -1.0f), fmaxf(a.y, -1.0f));
                v = __fadd2_rn(__fadd2_rn(lo, c), m1);
            } else if (MODE == 1) {     // scalar, two interleaved chains
                const float a0 = __fadd_rn(v.x, x.x), a1 = __fadd_rn(v.y, x.y);
                const float c0 = a0 >= 1.0f ? 0.0f : 1.0f, c1 = a1 >= 1.0f ? 0.0f : 1.0f;
                const float l0 = fmaxf(a0, -1.0f), l1 = fmaxf(a1, -1.0f);
                v.x = __fadd_rn(__fadd_rn(l0, c0), -1.0f);
                v.y = __fadd_rn(__fadd_rn(l1, c1), -1.0f);
            } else if (MODE == 2) {     // FADD2 only, 3 dependent
                v = __fadd2_rn(__fadd2_rn(__fadd2_rn(v, x), x), m1);
            } else if (MODE == 3) {     // FADD only, 3 dependent (one chain)
                v.x = __fadd_rn(__fadd_rn(__fadd_rn(v.x, x.x), x.x), -1.0f);
            } else if (MODE == 4) {     // I2F.S64 dependent chain
                long long q = (long long)__float_as_int(v.x) * 77777LL;
                v.x = __ll2float_rn(q) * 1e-9f;
            } else if (MODE == 5) {     // I2F.S64 throughput: 8 independent
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) acc += __ll2float_rn(((long long)__float_as_int(v.x) << 8) + k * 1234567LL);
                v.x = acc * 1e-12f;
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = v.x + v.y;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main()
{
    float *out; long long *cyc, h[4];
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
    const char *names[] = {"packed IAF step (FADD2,FSET|FMNMX,FADD2,FADD2)", "scalar IAF step x2 chains", "3 dependent FADD2", "3 dependent FADD",
                           "I2F.S64 dependent (+IMAD,FMUL)", "8 independent I2F.S64 (+adds)"};
    for (int warps = 1; warps <= 4; warps *= 2) {
        printf("warps per SMSP = %d\n", warps);
#define RUN(M) lat<M><<<1, 128 * warps>>>(out, cyc, 0.25f); cudaDeviceSynchronize(); lat<M><<<1, 128 * warps>>>(out, cyc, 0.25f); \
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); printf("  %-52s %.2f clk per inner iteration\n", names[M], (double)h[0] / (256.0 * 32));
        RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
