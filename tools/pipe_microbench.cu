// Microbenchmark (not part of the product): issue cost of the integer / float operations the tensor-core kernel's
// epilogue is made of, per warp with 1 / 2 / 4 warps per SM sub-partition (8 independent chains per thread).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/pipe_microbench tools/pipe_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CH 8
template <int OP>
__global__ void bench(int *out, long long *cyc, int seed)
{
    int a[CH], b[CH];
    long long w[CH];
    float f[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = seed + i * 77 + threadIdx.x; b[i] = seed * 3 + i; w[i] = a[i]; f[i] = (float)a[i] * 1e-3f; }
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < 512; ++r) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (OP == 0) asm volatile("mad.lo.s32 %0, %0, 256, %1;" : "+r"(a[i]) : "r"(b[i]));                  // IMAD
                if (OP == 1) asm volatile("mad.wide.s32 %0, %1, 65536, %0;" : "+l"(w[i]) : "r"(a[i]));             // IMAD.WIDE
                if (OP == 2) asm volatile("shf.r.clamp.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));              // SHF
                if (OP == 3) asm volatile("lop3.b32 %0, %0, %1, 0x55, 0x96;" : "+r"(a[i]) : "r"(b[i]));            // LOP3
                if (OP == 4) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));                         // IADD3 / IMAD.IADD
                if (OP == 5) asm volatile("{ .reg .s32 t; shl.b32 t, %0, 8; add.s32 %0, t, %1; }" : "+r"(a[i]) : "r"(b[i]));   // LEA?
                if (OP == 6) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[i]) : "r"(a[i] ^ __float_as_int(f[i])));       // I2FP
                if (OP == 7) asm volatile("fma.rn.f32 %0, %0, 1.0001, 0.5;" : "+f"(f[i]));                         // FFMA
                if (OP == 8) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"((float)b[i]));                   // FMNMX
                if (OP == 9) asm volatile("cvt.rn.f32.s64 %0, %1;" : "=f"(f[i]) : "l"(w[i] + __float_as_int(f[i])));        // I2F.S64
                if (OP == 10) asm volatile("max.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));                         // VIMNMX
                if (OP == 11) asm volatile("{ .reg .pred q; setp.lt.f32 q, %0, 1.0; selp.f32 %0, 1.0, 0.0, q; }" : "+f"(f[i]));   // FSET
            }
        }
    }
    long long t1 = clock64();
    int acc = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) acc += a[i] + (int)w[i] + __float_as_int(f[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main()
{
    int *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
    const char *names[] = {"IMAD (mad.lo.s32)", "IMAD.WIDE (mad.wide.s32)", "SHF", "LOP3", "add.s32", "shl+add (LEA?)", "cvt.rn.f32.s32 (I2FP)",
                           "FFMA", "max.f32 (FMNMX)", "cvt.rn.f32.s64 (I2F.S64)", "max.s32 (VIMNMX)", "setp+selp (FSET)"};
    printf("cycles per warp-instruction (8 independent chains per thread)\n%-28s %8s %8s %8s\n", "op", "1 warp", "2 warps", "4 warps");
#define RUN(M) { double r[3]; int k = 0; for (int wps = 1; wps <= 4; wps *= 2) { bench<M><<<1, 128 * wps>>>(out, cyc, 5); cudaDeviceSynchronize(); \
        bench<M><<<1, 128 * wps>>>(out, cyc, 5); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); r[k++] = (double)h / (512.0 * 4 * CH); } \
        printf("%-28s %8.2f %8.2f %8.2f\n", names[M], r[0], r[1], r[2]); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
