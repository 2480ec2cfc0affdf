// CUDA-core implicit-GEMM Conv1d kernels over the 16-bit panel layout (see dcue_b200.h).
// They consume exactly the operands the tcgen05 kernels consume (same packed weights, same
// panels, fp32 accumulation), so they serve as the on-device validator of conv_tc.cu and as
// the path for shapes the tensor-core kernels do not cover.
//
// Conv as shifted-row GEMM:  out[r, m] = sum_{j<k} sum_{c<128} A[m][j*128+c] * In[r+j, c]
// over flat rows r = s*Lp + t (reference: nn.Conv1d in truedcuemel1dbn.py:25-27,33-35,41-43,49-51).
#include "common.cuh"
#include "conv_common.cuh"

namespace {

constexpr int TRW = 32;   // flat rows per tile (multiple of every pool width)
constexpr int CH = 128;   // channels are padded to 128 in packed weights

// load rows [r0, r0+nrows) x 128 channels of a panel into smem as fp32 (row stride CH)
__device__ __forceinline__ void load_panel_tile(float* __restrict__ xs, const uint4* __restrict__ panel,
                                                long panel_rows, long r0, int nrows, int nch, int fmt, int tid,
                                                int nthreads) {
    const int chunks = nrows * (CH / 8);
    for (int e = tid; e < chunks; e += nthreads) {
        const int rl = e % nrows, q = e / nrows;  // consecutive threads walk rows of one panel
        uint4 v = make_uint4(0, 0, 0, 0);
        if (q * 8 < nch) v = __ldg(panel + (long)q * panel_rows + r0 + rl);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
        float* dst = xs + rl * CH + q * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            dst[2 * j] = cvt16_to_f32((unsigned short)(w[j] & 0xffffu), fmt);
            dst[2 * j + 1] = cvt16_to_f32((unsigned short)(w[j] >> 16), fmt);
        }
    }
}

template <int POOL>
__device__ __forceinline__ void epi_pool(const float (&acc)[TRW], long r0, int m, float bv, const ConvGeom& g,
                                         float* __restrict__ out, uint8_t* __restrict__ code, double& st1, double& st2,
                                         bool has_tb, const float (&tb)[4]) {
#pragma unroll
    for (int t0 = 0; t0 < TRW; t0 += POOL) {
        const long r = r0 + t0;
        const long s = r / g.Lp;
        const int q = (int)(r - s * g.Lp);
        const int p = q / POOL;
        float w[POOL];
#pragma unroll
        for (int i = 0; i < POOL; ++i) w[i] = acc[t0 + i];
        if (has_tb && (q < g.pad || q + POOL - 1 > g.Lin + g.pad - g.k)) {
#pragma unroll
            for (int i = 0; i < POOL; ++i) w[i] -= missing_taps(tb, q + i, g.k, g.pad, g.Lin);
        }
        float best = w[0];
        int bi = 0;
#pragma unroll
        for (int i = 1; i < POOL; ++i)
            if (w[i] > best) { best = w[i]; bi = i; }   // first maximum wins, like ATen
        if (s < g.S && p < g.P) {
            const float v = fmaxf(best + bv, 0.f);
            const long o = (s * g.P + p) * g.Cout + m;
            out[o] = v;
            if (code) code[o] = (uint8_t)bi;
            st1 += v;
            st2 += (double)v * v;
        }
    }
}

// EPI 0: bias + maxpool + relu + code + BN partial sums.   EPI 1: dgrad store of data rows.
template <int EPI>
__global__ void __launch_bounds__(128)
conv_rows_kernel(const uint4* __restrict__ panel, long panel_rows, int fmt_in, const uint4* __restrict__ wp, int fmt_w,
                 const float* __restrict__ bias, ConvGeom g, float* __restrict__ out, uint8_t* __restrict__ code,
                 double* __restrict__ partial, const float* __restrict__ gscale, const float* __restrict__ tap_bias) {
    extern __shared__ float xs[];  // [TRW + k - 1][CH]
    const float oscale = (EPI == 1 && gscale) ? gscale[1] : 1.f;
    const int m = threadIdx.x;
    const long ntiles = (g.rows_total + TRW - 1) / TRW;
    const int kpan = g.k * (CH / 8);  // 8-wide K chunks of the packed weight
    double st1 = 0.0, st2 = 0.0;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long r0 = tile * TRW;
        __syncthreads();
        load_panel_tile(xs, panel, panel_rows, r0, TRW + g.k - 1, g.Cin, fmt_in, m, 128);
        __syncthreads();
        float acc[TRW];
#pragma unroll
        for (int t = 0; t < TRW; ++t) acc[t] = 0.f;
        for (int kk8 = 0; kk8 < kpan; ++kk8) {
            const int j = kk8 / (CH / 8), c0 = (kk8 % (CH / 8)) * 8;
            if (c0 >= g.Cin) continue;
            const uint4 wv = __ldg(wp + (long)kk8 * CH + m);
            const unsigned ww[4] = {wv.x, wv.y, wv.z, wv.w};
            float w[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                w[2 * q] = cvt16_to_f32((unsigned short)(ww[q] & 0xffffu), fmt_w);
                w[2 * q + 1] = cvt16_to_f32((unsigned short)(ww[q] >> 16), fmt_w);
            }
#pragma unroll
            for (int t = 0; t < TRW; ++t) {
                const float4 a = *reinterpret_cast<const float4*>(xs + (t + j) * CH + c0);
                const float4 b = *reinterpret_cast<const float4*>(xs + (t + j) * CH + c0 + 4);
                acc[t] = fmaf(w[0], a.x, acc[t]); acc[t] = fmaf(w[1], a.y, acc[t]);
                acc[t] = fmaf(w[2], a.z, acc[t]); acc[t] = fmaf(w[3], a.w, acc[t]);
                acc[t] = fmaf(w[4], b.x, acc[t]); acc[t] = fmaf(w[5], b.y, acc[t]);
                acc[t] = fmaf(w[6], b.z, acc[t]); acc[t] = fmaf(w[7], b.w, acc[t]);
            }
        }
        if (EPI == 0) {
            if (m < g.Cout) {
                float bv = bias ? bias[m] : 0.f;
                float tb[4] = {0.f, 0.f, 0.f, 0.f};
                const bool has_tb = tap_bias != nullptr;
                if (has_tb)
                    for (int j = 0; j < g.k; ++j) { tb[j] = tap_bias[j * g.Cout + m]; bv += tb[j]; }
                if (g.pool == 4) epi_pool<4>(acc, r0, m, bv, g, out, code, st1, st2, has_tb, tb);
                else if (g.pool == 2) epi_pool<2>(acc, r0, m, bv, g, out, code, st1, st2, has_tb, tb);
                else epi_pool<1>(acc, r0, m, bv, g, out, code, st1, st2, has_tb, tb);
            }
        } else {
            if (m < g.Cout) {
#pragma unroll
                for (int t = 0; t < TRW; ++t) {
                    const long r = r0 + t;
                    const long s = r / g.Lp;
                    const int tt = (int)(r - s * g.Lp) - g.pad;
                    if (s < g.S && tt >= 0 && tt < g.Lin) out[(s * g.Lin + tt) * g.Cout + m] = acc[t] * oscale;
                }
            }
        }
    }
    if (EPI == 0 && partial && m < g.Cout) {
        partial[((long)blockIdx.x * 2 + 0) * g.Cout + m] = st1;
        partial[((long)blockIdx.x * 2 + 1) * g.Cout + m] = st2;
    }
}

// wgrad partials: block (chunk, tap j, ci tile of 32): part[chunk][co][j][ci] += dY[r,co]*X[r+j,ci]
__global__ void __launch_bounds__(128)
wgrad_rows_kernel(const uint4* __restrict__ dyp, long dy_rows, int fmt_dy, const uint4* __restrict__ xp, long x_rows,
                  int fmt_x, long rows_total, int k, int Cin, int Cout, float* __restrict__ part) {
    __shared__ float dys[TRW * CH];
    __shared__ float xs[(TRW + 3) * CH];
    const int co = threadIdx.x, j = blockIdx.y, c0 = blockIdx.z * 32;
    const long ntiles = (rows_total + TRW - 1) / TRW;
    const long per = (ntiles + gridDim.x - 1) / gridDim.x;
    const long tbeg = blockIdx.x * per, tend = min(ntiles, tbeg + per);
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.f;
    for (long tile = tbeg; tile < tend; ++tile) {
        const long r0 = tile * TRW;
        __syncthreads();
        load_panel_tile(dys, dyp, dy_rows, r0, TRW, Cout, fmt_dy, co, 128);
        load_panel_tile(xs, xp, x_rows, r0, TRW + k - 1, Cin, fmt_x, co, 128);
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < TRW; ++r) {
            const float d = dys[r * CH + co];
            const float* xr = xs + (r + j) * CH + c0;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 x = *reinterpret_cast<const float4*>(xr + c);
                acc[c] = fmaf(d, x.x, acc[c]);
                acc[c + 1] = fmaf(d, x.y, acc[c + 1]);
                acc[c + 2] = fmaf(d, x.z, acc[c + 2]);
                acc[c + 3] = fmaf(d, x.w, acc[c + 3]);
            }
        }
    }
    float* dst = part + (((long)blockIdx.x * CH + co) * k + j) * CH + c0;
#pragma unroll
    for (int c = 0; c < 32; ++c) dst[c] = acc[c];
}

}  // namespace

// dW[co][ci][j] = sum_parts part[p][co][j][ci]   (shared with the tcgen05 wgrad)
__global__ void dcue_wgrad_reduce_kernel(const float* __restrict__ part, int nparts, int Cout, int Cin, int k,
                                         const float* __restrict__ gscale, float* __restrict__ dW) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Cout * Cin * k) return;
    const int j = i % k, ci = (i / k) % Cin, co = i / (k * Cin);
    // fixed summation order (deterministic); eight independent loads in flight (the plain loop was L2-latency bound: 17 us)
    const float* src = part + ((long)co * k + j) * 128 + ci;
    const long stride = 128L * k * 128;
    // (in-graph timeline: 10 us for 38 MB of L2-resident partials = 19 dependent batches of 8 loads; 32 in flight -> 5 batches)
    float s = 0.f;
    int p = 0;
    for (; p + 32 <= nparts; p += 32) {
        float v[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) v[u] = __ldcg(src + (long)(p + u) * stride);
#pragma unroll
        for (int u = 0; u < 32; ++u) s += v[u];
    }
    for (; p + 8 <= nparts; p += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + (long)(p + u) * stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; p < nparts; ++p) s += __ldcg(src + (long)p * stride);
    dW[i] = gscale ? s * gscale[1] : s;
}

__global__ void __launch_bounds__(256)
dcue_reduce_partials_d(const double* __restrict__ partial, int nblk, int n, double* __restrict__ out) {
    __shared__ double red[32][9];
    const int cl = threadIdx.x & 7, rl = threadIdx.x >> 3;
    const int j = blockIdx.x * 8 + cl;
    double s = 0.0;
    if (j < n)
        for (int b = rl; b < nblk; b += 32) s += partial[(long)b * n + j];
    red[rl][cl] = s;
    __syncthreads();
    if (rl == 0 && j < n) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 32; ++k) t += red[k][cl];
        out[j] = t;
    }
}

__global__ void pack_conv_weight_kernel(const float* __restrict__ W, int Cout, int Cin, int k, int mode, int fmt,
                                        const float* __restrict__ col_scale, const float* __restrict__ col_scale2,
                                        unsigned short* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int K = k * 128;
    if (i >= 128 * K) return;
    const int kk = i / 128, m = i % 128;          // consecutive threads -> consecutive m
    const int j = kk / 128, c = kk % 128;
    float v = 0.f;
    if (mode == 0) {  // A[co=m][j*128+ci=c] = W[co][ci][j]
        if (m < Cout && c < Cin)
            v = W[((long)m * Cin + c) * k + j] * (col_scale ? col_scale[c] : 1.f) * (col_scale2 ? col_scale2[c] : 1.f);
    } else {          // A[ci=m][jj*128+co=c] = W[co][ci][k-1-jj]
        if (m < Cin && c < Cout) v = W[((long)c * Cin + m) * k + (k - 1 - j)];
    }
    out[((long)(kk >> 3) * 128 + m) * 8 + (kk & 7)] = cvt_f32_to16(v, fmt);
}

int dcue_simt_conv_fwd(const void* panel, long panel_rows, int fmt, const void* w_packed, const float* bias,
                       const float* tap_bias, const ConvGeom& g, float* z, uint8_t* code, double* sums, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
    const long ntiles = (g.rows_total + TRW - 1) / TRW;
    const long cap = (long)dcue_num_sms() * 4;
    const int grid = (int)(ntiles < cap ? (ntiles > 0 ? ntiles : 1) : cap);
    if (sums && (!ws || ws_bytes < (size_t)grid * 2 * g.Cout * sizeof(double)))
        DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_conv_pool_fwd(simt): workspace too small");
    const size_t smem = (size_t)(TRW + g.k - 1) * CH * sizeof(float);
    conv_rows_kernel<0><<<grid, 128, smem, st>>>((const uint4*)panel, panel_rows, fmt, (const uint4*)w_packed, fmt, bias,
                                                 g, z, code, sums ? (double*)ws : nullptr, nullptr, tap_bias);
    DCUE_LAUNCH_CHECK();
    if (sums && sums != DCUE_STATS_PARTIALS) {
        dcue_reduce_partials_d<<<ceil_div_i(2 * g.Cout, 8), 256, 0, st>>>((const double*)ws, grid, 2 * g.Cout, sums);
        DCUE_LAUNCH_CHECK();
    }
    return 0;
}

int dcue_simt_conv_fwd_nparts(long rows_total) {
    const long ntiles = (rows_total + TRW - 1) / TRW;
    const long cap = (long)dcue_num_sms() * 4;
    return (int)(ntiles < cap ? (ntiles > 0 ? ntiles : 1) : cap);
}

int dcue_simt_conv_dgrad(const void* dy_panel_shifted, long panel_rows, int fmt_dy, const void* w_packed, int fmt_w,
                         const ConvGeom& g, const float* gscale, float* dx, cudaStream_t st) {
    const long ntiles = (g.rows_total + TRW - 1) / TRW;
    const long cap = (long)dcue_num_sms() * 4;
    const int grid = (int)(ntiles < cap ? (ntiles > 0 ? ntiles : 1) : cap);
    const size_t smem = (size_t)(TRW + g.k - 1) * CH * sizeof(float);
    conv_rows_kernel<1><<<grid, 128, smem, st>>>((const uint4*)dy_panel_shifted, panel_rows, fmt_dy,
                                                 (const uint4*)w_packed, fmt_w, nullptr, g, dx, nullptr, nullptr, gscale, nullptr);
    DCUE_LAUNCH_CHECK();
    return 0;
}

int dcue_simt_conv_wgrad(const void* dy_panel, long dy_rows, int fmt_dy, const void* x_panel, long x_rows, int fmt_x,
                         long rows_total, int k, int Cin, int Cout, const float* gscale, float* dW, void* ws,
                         size_t ws_bytes, cudaStream_t st) {
    const int nchunks = 32;
    if (!ws || ws_bytes < (size_t)nchunks * 128 * k * 128 * sizeof(float))
        DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_conv_wgrad(simt): workspace too small");
    dim3 grid(nchunks, k, 4);
    wgrad_rows_kernel<<<grid, 128, 0, st>>>((const uint4*)dy_panel, dy_rows, fmt_dy, (const uint4*)x_panel, x_rows,
                                            fmt_x, rows_total, k, Cin, Cout, (float*)ws);
    DCUE_LAUNCH_CHECK();
    dcue_wgrad_reduce_kernel<<<ceil_div_i((long)Cout * Cin * k, 256), 256, 0, st>>>((const float*)ws, nchunks, Cout, Cin,
                                                                                    k, gscale, dW);
    DCUE_LAUNCH_CHECK();
    return 0;
}

extern "C" int dcue_pack_conv_weight(const float* W, int Cout, int Cin, int k, int mode, int fmt, const float* col_scale,
                                     const float* col_scale2, void* out, void* stream) {
    DCUE_CHECK_ARG(W && out && Cout > 0 && Cout <= 128 && Cin > 0 && Cin <= 128 && k >= 1 && k <= 4 &&
                   (mode == 0 || mode == 1));
    pack_conv_weight_kernel<<<ceil_div_i(128L * k * 128, 256), 256, 0, (cudaStream_t)stream>>>(
        W, Cout, Cin, k, mode, fmt, col_scale, col_scale2, (unsigned short*)out);
    DCUE_LAUNCH_CHECK();
    return 0;
}


// tapB[j][co] = sum_ci W[co][ci][j] * beta[ci]   (input-BatchNorm shift folded into the conv bias)
// the constant each input channel carries is beta[c] + gamma[c]*shift[c] (shift: x-hat = rstd*u + shift)
// one warp per output channel: lanes stride the contiguous [Cin][k] weight row (coalesced), shuffle reduction
__global__ void __launch_bounds__(128)
tap_bias_kernel(const float* __restrict__ W, int Cout, int Cin, int k, const float* __restrict__ beta,
                const float* __restrict__ gamma, const float* __restrict__ shift, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int co = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (co >= Cout) return;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    const float* w = W + (long)co * Cin * k;
    for (int c = lane; c < Cin; c += 32) {
        const float b = beta[c] + (shift ? (gamma ? gamma[c] : 1.f) * shift[c] : 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < k) s[j] = fmaf(w[c * k + j], b, s[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float t = warp_sum(s[j]);
        if (lane == 0 && j < k) out[j * Cout + co] = t;
    }
}

extern "C" int dcue_conv_tap_bias(const float* W, int Cout, int Cin, int k, const float* beta, const float* gamma,
                                  const float* shift, float* tap_bias, void* stream) {
    DCUE_CHECK_ARG(W && beta && tap_bias && Cout > 0 && Cin > 0 && k >= 1 && k <= 4);
    tap_bias_kernel<<<ceil_div_i(Cout, 4), 128, 0, (cudaStream_t)stream>>>(W, Cout, Cin, k, beta, gamma, shift, tap_bias);
    DCUE_LAUNCH_CHECK();
    return 0;
}

// out[z][i][c] = inv_scale * (sum over the z-th slice of the spectrograms) panel[s*Lp + rows[i]][c] for up to 4 rows per
// spectrogram; the PRS_SLICES partial sums are added in fixed order by the consumer (dcue_bn_fold_grads).
// (The one-block-per-(panel,row) version ran 64 blocks of 84 strided loads each: 60 us for 22 MB.)
constexpr int PRS_SLICES = 148;     // one slice per SM for the border-row kernels
__global__ void __launch_bounds__(256)
panel_row_sums_kernel(const uint4* __restrict__ panel, long panel_rows, int fmt, int S, int Lp, int4 rows,
                      const float* __restrict__ gscale, float* __restrict__ out /* [PRS_SLICES][4][128] */) {
    __shared__ float red[8][8];
    const int q = blockIdx.x, ri = blockIdx.y, z = blockIdx.z;
    const int row = ri == 0 ? rows.x : ri == 1 ? rows.y : ri == 2 ? rows.z : rows.w;
    const int per = (S + PRS_SLICES - 1) / PRS_SLICES;
    const int s0 = z * per, s1 = min(S, s0 + per);
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (row >= 0)
        for (int s = s0 + threadIdx.x; s < s1; s += 256) {
            const uint4 v = __ldg(panel + (long)q * panel_rows + (long)s * Lp + row);
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a[2 * j] += cvt16_to_f32((unsigned short)(w[j] & 0xffffu), fmt);
                a[2 * j + 1] += cvt16_to_f32((unsigned short)(w[j] >> 16), fmt);
            }
        }
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float v = warp_sum(a[j]);
        if (lane == 0) red[wp][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        float t = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < 8; ++w2) t += red[w2][threadIdx.x];
        out[(z * 4 + ri) * 128 + q * 8 + threadIdx.x] = t * (gscale ? gscale[1] : 1.f);
    }
}

extern "C" int dcue_panel_row_sums_parts(void) { return PRS_SLICES; }

extern "C" int dcue_panel_row_sums(const void* panel, long panel_rows, int fmt, int S, int Lp, int r0, int r1, int r2, int r3,
                                   const float* gscale, float* out, void* stream) {
    DCUE_CHECK_ARG(panel && out && S >= 0 && Lp > 0 && r0 < Lp && r1 < Lp && r2 < Lp && r3 < Lp);
    dim3 grid(16, 4, PRS_SLICES);
    panel_row_sums_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)panel, panel_rows, fmt, S, Lp,
                                                                  make_int4(r0, r1, r2, r3), gscale, out);
    DCUE_LAUNCH_CHECK();
    return 0;
}

// The same border row sums taken from the POOLED inputs (no dY panel exists on the fused unpool+wgrad path):
// out[z][e][c] = sum over slice z of [code[s*P+p_e][c] == c_e] * dz(s*P+p_e, c),  flat border row r_e = p_e*pool + c_e,
// dz = relu'(z) * (a*dy + b*z + c) as in bn_relu_unpool_rows_kernel (unscaled: true gradient units).
__global__ void __launch_bounds__(512)
border_row_sums_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ dtp, int lddtp,
                       const float* __restrict__ z, const uint8_t* __restrict__ code, const float* __restrict__ scale,
                       const float* __restrict__ mean, const float* __restrict__ rstd, const double* __restrict__ sums,
                       double count, int S, int P, int pool, int4 rows, float* __restrict__ out /* [PRS_SLICES][4][128] */) {
    // block (pair, slice): border rows come in pairs that share a pooling window (rows 0,1 -> window 0; the two trailing
    // rows -> the last window), so one load of (dy, z, code) serves both rows of the pair
    __shared__ float red[2][4][128];
    const int c = threadIdx.x & 127, ph = threadIdx.x >> 7;
    const int pr = blockIdx.x, zs = blockIdx.y;
    const int row_a = pr == 0 ? rows.x : rows.z, row_b = pr == 0 ? rows.y : rows.w;
    float acc_a = 0.f, acc_b = 0.f;
    const bool va = row_a >= 0 && row_a / pool < P, vb = row_b >= 0 && row_b / pool < P;
    if (va || vb) {
        const int pe_a = va ? row_a / pool : -1, pe_b = vb ? row_b / pool : -1;
        const unsigned ce_a = va ? (unsigned)(row_a - pe_a * pool) : 255u, ce_b = vb ? (unsigned)(row_b - pe_b * pool) : 255u;
        const float sc = scale ? scale[c] : 1.f;
        float kb = 0.f, kc = 0.f;
        if (sums) {
            const double inv_n = 1.0 / count;
            const double rs2 = (double)rstd[c] * sums[128 + c] * inv_n;
            kb = (float)(-(double)sc * rs2);
            kc = (float)((double)sc * (rs2 * (double)mean[c] - sums[c] * inv_n));
        }
        const float invP = 1.f / (float)P;
        const int per = (S + PRS_SLICES - 1) / PRS_SLICES;
        const int s0 = zs * per, s1 = min(S, s0 + per);
        const bool shared_window = va && vb && pe_a == pe_b;
#pragma unroll 4
        for (int sp = s0 + ph; sp < s1; sp += 4) {
            const float tpv = dtp ? dtp[(long)sp * lddtp + c] : 0.f;
            if (va) {
                const long r = (long)sp * P + pe_a;
                const float g = fmaf(tpv, invP, dy[r * lddy + c]);
                const float zz = z[r * 128 + c];
                const unsigned cd = code[r * 128 + c];
                const float v = zz > 0.f ? fmaf(sc, g, fmaf(kb, zz, kc)) : 0.f;
                acc_a += cd == ce_a ? v : 0.f;
                if (shared_window) acc_b += cd == ce_b ? v : 0.f;
            }
            if (vb && !shared_window) {
                const long r = (long)sp * P + pe_b;
                const float g = fmaf(tpv, invP, dy[r * lddy + c]);
                const float zz = z[r * 128 + c];
                const unsigned cd = code[r * 128 + c];
                const float v = zz > 0.f ? fmaf(sc, g, fmaf(kb, zz, kc)) : 0.f;
                acc_b += cd == ce_b ? v : 0.f;
            }
        }
    }
    red[0][ph][c] = acc_a;
    red[1][ph][c] = acc_b;
    __syncthreads();
    if (ph < 2) out[(zs * 4 + pr * 2 + ph) * 128 + c] = (red[ph][0][c] + red[ph][1][c]) + (red[ph][2][c] + red[ph][3][c]);
}

extern "C" int dcue_border_row_sums(const float* dy, int lddy, const float* dtp, int lddtp, const float* z, const uint8_t* code,
                                    const float* scale, const float* mean, const float* rstd, const double* sums, double count,
                                    int S, int P, int C, int pool, int r0, int r1, int r2, int r3, float* out, void* stream) {
    DCUE_CHECK_ARG(dy && z && code && out && S >= 0 && P > 0 && C == 128 && pool >= 1 && lddy >= C);
    DCUE_CHECK_ARG(!sums || (mean && rstd && count > 0));
    dim3 grid(2, PRS_SLICES);
    border_row_sums_kernel<<<grid, 512, 0, (cudaStream_t)stream>>>(dy, lddy, dtp, lddtp, z, code, scale, mean, rstd, sums,
                                                                   count > 0 ? count : 1.0, S, P, pool, make_int4(r0, r1, r2, r3), out);
    DCUE_LAUNCH_CHECK();
    return 0;
}

// Gradients of a conv whose input BatchNorm (gamma, beta) was folded into it:
//   conv input = gamma*xhat + beta inside the data rows, zero padding outside;  G = sum dY * xhat (wgrad on xhat)
//   dW[co,ci,j] = gamma[ci]*G + beta[ci]*T[co,j],  dgamma[ci] = sum W*G,  dbeta[ci] = sum W*T,
//   T[co,j] = sum over conv outputs t whose tap j reads a data row of dY[.,co,t]
//           = Tall[co] - sum_{border rows t with tap j outside} E[t][co]
__global__ void __launch_bounds__(128)
bn_fold_grads_kernel(const float* __restrict__ G, const float* __restrict__ W, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ xs_scale, const float* __restrict__ xs_shift,
                     const float* __restrict__ Tall, const float* __restrict__ E /* [PRS_SLICES][4][128] */,
                     int4 erows, int Cout, int Cin, int k, int pad, int Lin, float* __restrict__ dW,
                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
    __shared__ float r1[128], r2[128];
    const int ci = blockIdx.x, co = threadIdx.x;
    float sg = 0.f, sb = 0.f;
    if (co < Cout) {
        const int er[4] = {erows.x, erows.y, erows.z, erows.w};
        float Esum[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {   // fixed-order sum of the row-sum slices
            float t = 0.f;
            for (int z = 0; z < PRS_SLICES; ++z) t += E[(z * 4 + e) * 128 + co];
            Esum[e] = t;
        }
        const float ga = gamma[ci], be = beta[ci];
        // the operand was u with x-hat = xs_scale*u + xs_shift: G(x-hat) = xs_scale*G(u) + xs_shift*T
        const float xr = xs_scale ? xs_scale[ci] : 1.f, xsft = xs_shift ? xs_shift[ci] : 0.f;
        for (int j = 0; j < k; ++j) {
            float T = Tall[co];
            for (int e = 0; e < 4; ++e) {
                if (er[e] < 0) continue;
                const int tt = er[e] + j - pad;
                if (tt < 0 || tt >= Lin) T -= Esum[e];
            }
            const long o = ((long)co * Cin + ci) * k + j;
            const float g = xr * G[o] + xsft * T, w = W[o];
            dW[o] = ga * g + be * T;
            sg = fmaf(w, g, sg);
            sb = fmaf(w, T, sb);
        }
    }
    r1[co] = sg;
    r2[co] = sb;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (co < o) { r1[co] += r1[co + o]; r2[co] += r2[co + o]; }
        __syncthreads();
    }
    if (co == 0) { dgamma[ci] = r1[0]; dbeta[ci] = r2[0]; }
}

extern "C" int dcue_bn_fold_grads(const float* G, const float* W, const float* gamma, const float* beta,
                                  const float* xs_scale, const float* xs_shift, const float* Tall, const float* E, int r0, int r1, int r2, int r3, int Cout, int Cin, int k, int pad, int Lin,
                                  float* dW, float* dgamma, float* dbeta, void* stream) {
    DCUE_CHECK_ARG(G && W && gamma && beta && Tall && E && dW && dgamma && dbeta && Cout > 0 && Cout <= 128 && Cin > 0 &&
                   k >= 1 && k <= 4);
    bn_fold_grads_kernel<<<Cin, 128, 0, (cudaStream_t)stream>>>(G, W, gamma, beta, xs_scale, xs_shift, Tall, E, make_int4(r0, r1, r2, r3), Cout,
                                                                Cin, k, pad, Lin, dW, dgamma, dbeta);
    DCUE_LAUNCH_CHECK();
    return 0;
}
