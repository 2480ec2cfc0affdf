// Geometry of one conv stage in the flat padded row space (see dcue_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct ConvGeom {
    int S;            // spectrograms
    int Lp;           // rows per spectrogram in the flat layout
    int Lin;          // data rows per spectrogram (dgrad epilogue)
    int pad;          // zero rows in front of each spectrogram's data
    int k;            // taps
    int pool;         // max-pool width (forward epilogue)
    int P;            // pooled outputs per spectrogram
    int Cin;          // contraction channels (<=128)
    int Cout;         // output channels (<=128)
    long rows_total;  // S * Lp
};

// Sum of the per-tap constants tb[j] whose tap falls outside the data rows for conv output t
// (zero padding): used when an input-side BatchNorm shift is folded into the conv bias.
__device__ __forceinline__ float missing_taps(const float (&tb)[4], int t, int k, int pad, int Lin) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int tt = t + j - pad;
        if (j < k && (tt < 0 || tt >= Lin)) a += tb[j];
    }
    return a;
}

// `sums` value that asks the forward kernels to leave their per-block statistic partials ([nparts][2*Cout] doubles) at the
// start of the workspace instead of reducing them (dcue_conv_pool_fwd_parts + dcue_bn_stats_finalize)
#define DCUE_STATS_PARTIALS (reinterpret_cast<double*>(uintptr_t(8)))
int dcue_tc_conv_fwd_nparts(long rows_total);
int dcue_simt_conv_fwd_nparts(long rows_total);

// CUDA-core path (conv_simt.cu)
int dcue_simt_conv_fwd(const void* panel, long panel_rows, int fmt, const void* w_packed, const float* bias,
                       const float* tap_bias, const ConvGeom& g, float* z, uint8_t* code, double* sums, void* ws, size_t ws_bytes,
                       cudaStream_t st);
int dcue_simt_conv_dgrad(const void* dy_panel_shifted, long panel_rows, int fmt_dy, const void* w_packed, int fmt_w,
                         const ConvGeom& g, const float* gscale, float* dx, cudaStream_t st);
int dcue_simt_conv_wgrad(const void* dy_panel, long dy_rows, int fmt_dy, const void* x_panel, long x_rows, int fmt_x,
                         long rows_total, int k, int Cin, int Cout, const float* gscale, float* dW, void* ws,
                         size_t ws_bytes, cudaStream_t st);
// tcgen05 path (conv_tc.cu)
int dcue_tc_conv_fwd(const void* panel, long panel_rows, int fmt, const void* w_packed, const float* bias,
                     const float* tap_bias, const ConvGeom& g, float* z, uint8_t* code, double* sums, void* ws, size_t ws_bytes,
                     cudaStream_t st);
int dcue_tc_conv_dgrad(const void* dy_panel_shifted, long panel_rows, int fmt_dy, const void* w_packed, int fmt_w,
                       const ConvGeom& g, const float* gscale, float* dx, void* ws, size_t ws_bytes, cudaStream_t st);
int dcue_tc_conv_wgrad(const void* dy_panel, long dy_rows, int fmt_dy, const void* x_panel, long x_rows, int fmt_x,
                       long rows_total, int k, int Cin, int Cout, const float* gscale, float* dW, void* ws,
                       size_t ws_bytes, cudaStream_t st);
int dcue_tc_conv_dgrad_stats(const void* dy_panel_shifted, long panel_rows, int fmt_dy, const void* w_packed, int fmt_w,
                             const ConvGeom& g, const float* gscale, float* dx, const float* z, const float* mean, const float* rstd,
                             const float* dtp, int lddtp, void* ws, size_t ws_bytes, cudaStream_t st);
size_t dcue_tc_ws_bytes(int k);
size_t dcue_tc_wgrad_unpool_ws_bytes(int k);
int dcue_tc_conv_wgrad_unpool(const float* dy, int lddy, const float* dtp, int lddtp, const float* z, const uint8_t* code,
                              const float* scale, const float* mean, const float* rstd, const double* sums, double count, int S,
                              int P, int Lp, const void* x_panel, long x_rows, int fmt, int k, const float* gscale, float* dW,
                              double* bias_sums, float* bias_out, void* ws, size_t ws_bytes, cudaStream_t st);

__global__ void dcue_wgrad_reduce_kernel(const float* __restrict__ part, int nparts, int Cout, int Cin, int k,
                                         const float* __restrict__ gscale, float* __restrict__ dW);
__global__ void dcue_reduce_partials_d(const double* __restrict__ partial, int nblk, int n, double* __restrict__ out);
