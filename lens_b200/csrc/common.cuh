// Shared helpers for liblens_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/lens_b200.h"

namespace lens {

// thread-local last-error string (lens_last_error)
char *err_buf();
void set_err(const char *fmt, ...);

#define LENS_CHECK_ARG(cond, ...)                    \
    do {                                             \
        if (!(cond)) {                               \
            ::lens::set_err(__VA_ARGS__);            \
            return -1;                               \
        }                                            \
    } while (0)

#define LENS_CUDA(call)                                                              \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            ::lens::set_err("%s:%d %s -> %s", __FILE__, __LINE__, #call,             \
                            cudaGetErrorString(e__));                                \
            return (int)e__;                                                         \
        }                                                                            \
    } while (0)

// every kernel launch of the library goes through this (counts launches, surfaces launch errors)
void count_launch();
#define LENS_LAUNCH_CHECK()          \
    do {                             \
        ::lens::count_launch();      \
        LENS_CUDA(cudaGetLastError()); \
    } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

int sm_count();

// One timestep of sinabs IAFSqueeze (alpha = 1, MultiSpike, MembraneSubtract,
// min_v_mem clip written as relu(v - v_min) + v_min): every op is a separately
// rounded IEEE fp32 op, exactly as torch's elementwise kernels evaluate it
// (no FMA contraction: the intrinsics below are never fused by nvcc).
__device__ __forceinline__ float iaf_step(float &v, float x, float thr, float vmin)
{
    float vv = __fadd_rn(v, x);                       // 1.0 * v + x
    float s = 0.0f;
    if (vv > 0.0f) s = truncf(__fdiv_rn(vv, thr));    // (v > 0) * trunc(v / thr)
    vv = __fsub_rn(vv, __fmul_rn(s, thr));            // v - s * thr
    float r = fmaxf(__fsub_rn(vv, vmin), 0.0f);       // relu(v - v_min)
    v = __fadd_rn(r, vmin);                           //   + v_min
    return s;
}

}  // namespace lens
