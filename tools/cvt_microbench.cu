// Microbenchmark (not part of the product): throughput of int64->fp32 conversion variants per warp.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void bench(long long *out, const int *in, float *sink)
{
    int32_t xl[32], xh[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { xl[i] = in[i] + threadIdx.x; xh[i] = in[32 + i]; }
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < 256; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            float f;
            if (MODE == 0) f = __ll2float_rn((int64_t)(((uint64_t)(uint32_t)xh[i] << 32) | (uint32_t)xl[i]));       // I2F.S64
            else if (MODE == 1) f = __int2float_rn(xl[i]);                                                          // I2F.S32
            else if (MODE == 2) {                                                                                   // FP64 path
                double d = __hiloint2double(0x43300000 ^ 0x80000000, 0) ;
                double hi = __hiloint2double(0x43300000, (uint32_t)xh[i] ^ 0x80000000u) - 4503601774854144.0;  // 2^52 + 2^31
                double lo = __hiloint2double(0x43300000, (uint32_t)xl[i]) - 4503599627370496.0;                // 2^52
                f = __double2float_rn(fma(hi, 4294967296.0, lo)); (void)d;
            } else {                                                                                                // magic fp32 (only |x| < 2^22)
                f = __int_as_float(0x4B400000 + xl[i]) - 12582912.0f;
            }
            acc += f;
            xl[i] += it;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == 1.2345f) sink[0] = acc;
}

template <int MODE>
void run(const char *name, int warps_per_sm)
{
    long long *d_out; int *d_in; float *sink;
    cudaMalloc(&d_out, 8); cudaMalloc(&d_in, 64 * 4); cudaMalloc(&sink, 4);
    cudaMemset(d_in, 1, 64 * 4);
    bench<MODE><<<148, 32 * warps_per_sm>>>(d_out, d_in, sink);
    bench<MODE><<<148, 32 * warps_per_sm>>>(d_out, d_in, sink);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
    printf("%-22s warps/SM=%2d: %6.2f clk per conversion-warp-instruction (per warp), %6.2f clk/conv per SMSP\n", name, warps_per_sm,
           (double)c / (256 * 32), (double)c / (256 * 32) / ((warps_per_sm + 3) / 4));
}

int main()
{
    for (int w : {4, 8, 16}) {
        if (w == 4) { run<0>("I2F.S64", 4); run<1>("I2F.S32", 4); run<2>("fp64 magic + F2F", 4); run<3>("fp32 magic (ALU)", 4); }
        if (w == 8) { run<0>("I2F.S64", 8); run<1>("I2F.S32", 8); run<2>("fp64 magic + F2F", 8); run<3>("fp32 magic (ALU)", 8); }
        if (w == 16) { run<0>("I2F.S64", 16); run<1>("I2F.S32", 16); run<2>("fp64 magic + F2F", 16); run<3>("fp32 magic (ALU)", 16); }
    }
    return 0;
}
