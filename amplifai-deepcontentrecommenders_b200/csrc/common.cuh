// Shared helpers for the DCUE sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dcue_b200.h"

extern thread_local char g_dcue_err[256];
extern long g_dcue_launches;  // kernels launched through this library (bench.py reports it)

#define DCUE_FAIL(code, ...)                                   \
    do {                                                       \
        snprintf(g_dcue_err, sizeof(g_dcue_err), __VA_ARGS__); \
        return (code);                                         \
    } while (0)

#define DCUE_CHECK_ARG(cond)                                                                  \
    do {                                                                                      \
        if (!(cond)) DCUE_FAIL(DCUE_E_BADARG, "%s:%d: bad argument: %s", __func__, __LINE__, #cond); \
    } while (0)

#define DCUE_LAUNCH_CHECK()                                                                     \
    do {                                                                                        \
        ++g_dcue_launches;                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ != cudaSuccess)                                                                 \
            DCUE_FAIL((int)e__, "%s:%d: CUDA error: %s", __func__, __LINE__, cudaGetErrorString(e__)); \
    } while (0)

#define DCUE_CUDA(call)                                                                         \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            DCUE_FAIL((int)e__, "%s:%d: CUDA error: %s", __func__, __LINE__, cudaGetErrorString(e__)); \
    } while (0)

static inline int dcue_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

__host__ __device__ static inline long round_up_l(long a, long b) { return (a + b - 1) / b * b; }
__host__ __device__ static inline int ceil_div_i(long a, long b) { return (int)((a + b - 1) / b); }

// 16-bit storage helpers (raw ushort so one kernel serves both formats)
__device__ __forceinline__ float cvt16_to_f32(unsigned short v, int fmt) {
    if (fmt == DCUE_FMT_F16) return __half2float(__ushort_as_half(v));
    return __uint_as_float(((unsigned)v) << 16);
}
__device__ __forceinline__ unsigned short cvt_f32_to16(float x, int fmt) {
    if (fmt == DCUE_FMT_F16) return __half_as_ushort(__float2half_rn(x));
    return __bfloat16_as_ushort(__float2bfloat16_rn(x));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// panel addressing: element (row r, channel c)
__device__ __forceinline__ long panel_off(long panel_rows, long r, int c) {
    return ((long)(c >> 3) * panel_rows + r) * 8 + (c & 7);
}
