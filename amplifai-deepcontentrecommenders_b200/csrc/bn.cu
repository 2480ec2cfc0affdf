// HBM-bound glue of the song tower: BatchNorm statistics / finalize / apply, the fused
// transpose+convert of the NCL fp32 input into 16-bit panels (this also replaces the
// reference's torch.cat([pos,neg]) copy, dcrecommend/dcue/dcue.py:90), and the backward pass
// BatchNorm-backward + ReLU mask + MaxPool unpooling in one sweep.
// Reference semantics: truedcuemel1dbn.py:77-101 (order conv -> pool -> relu -> bn).
#include "common.cuh"
#include "peer.cuh"

namespace {

// Where spectrogram s lives: either the s-th row of `pos` / `neg` (dense API, row stride L) or a crop of a
// resident song pool, pool[idx[s], :, off[s] : off[s]+L] (index API, row stride T).  Out-of-range entries
// raise the error flag and read song 0 / offset 0.
struct SpecSrc {
    const float* pos; const float* neg; int S_pos;
    const int64_t* idx; const int32_t* off; long T; long n_songs; int* err;
    __device__ __forceinline__ const float* base(long s, int C, int L, long& row_stride) const {
        if (idx) {
            long i = idx[s];
            long o = off ? off[s] : 0;
            if (i < 0 || i >= n_songs || o < 0 || o + L > T) {
                if (err) atomicExch(err, 1);
                i = 0;
                o = 0;
            }
            row_stride = T;
            return pos + i * (long)C * T + o;
        }
        row_stride = L;
        return s < S_pos ? pos + s * (long)C * L : neg + (s - S_pos) * (long)C * L;
    }
};

// ------------------------------------------------------------------ NCL input statistics
// block = 8 warps; warp w owns channels w, w+8, ... (<=16 per warp for C=128); lanes stride
// the L frames of one (s,c) row; per-lane fp32 partials, reduced across lanes/blocks in fp64.
constexpr int STAT_MAXC_PER_WARP = 16;

__global__ void __launch_bounds__(256)
ncl_stats_kernel(SpecSrc src, int S, int C, int L, double* __restrict__ partial /* [grid][2][C] */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float s1[STAT_MAXC_PER_WARP], s2[STAT_MAXC_PER_WARP];
#pragma unroll
    for (int i = 0; i < STAT_MAXC_PER_WARP; ++i) s1[i] = s2[i] = 0.f;
    for (int s = blockIdx.x; s < S; s += gridDim.x) {
        long rs;
        const float* base = src.base(s, C, L, rs);
        if (L <= 160) {
            // 4 channel rows x 5 strided frames = 20 independent loads in flight per lane
#pragma unroll
            for (int i0 = 0; i0 < STAT_MAXC_PER_WARP; i0 += 4) {
                float v[4][5];
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    const int c = w + 8 * (i0 + ii);
                    const float* row = base + (long)c * rs;
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const int t = lane + 32 * k;
                        v[ii][k] = (c < C && t < L) ? __ldg(row + t) : 0.f;
                    }
                }
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    float a = 0.f, b = 0.f;
#pragma unroll
                    for (int k = 0; k < 5; ++k) { a += v[ii][k]; b = fmaf(v[ii][k], v[ii][k], b); }
                    s1[i0 + ii] += a;
                    s2[i0 + ii] += b;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < STAT_MAXC_PER_WARP; ++i) {
                const int c = w + 8 * i;
                if (c < C) {
                    const float* row = base + (long)c * rs;
                    float a = 0.f, b = 0.f;
                    for (int t = lane; t < L; t += 32) {
                        const float v = __ldg(row + t);
                        a += v;
                        b = fmaf(v, v, b);
                    }
                    s1[i] += a;
                    s2[i] += b;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < STAT_MAXC_PER_WARP; ++i) {
        const int c = w + 8 * i;
        const double a = warp_sum_d((double)s1[i]), b = warp_sum_d((double)s2[i]);
        if (lane == 0 && c < C) {
            partial[((long)blockIdx.x * 2 + 0) * C + c] = a;
            partial[((long)blockIdx.x * 2 + 1) * C + c] = b;
        }
    }
}

// out[j] = sum_b partial[b][j]; block = 8 columns x 32 row lanes, fixed summation order (deterministic)
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const double* __restrict__ partial, int nblk, int n, double* __restrict__ out,
                       float* __restrict__ fout0 = nullptr, float* __restrict__ fout1 = nullptr, int half = 0) {
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
        // optional fp32 copies of the two halves (e.g. dbeta = sums[0:C], dgamma = sums[C:2C])
        if (fout0 && j < half) fout0[j] = (float)t;
        if (fout1 && j >= half) fout1[j - half] = (float)t;
    }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ rmean, float* __restrict__ rvar,
                                   int64_t* __restrict__ nbt, float momentum, float eps, int training,
                                   const float* __restrict__ center, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_o, float* __restrict__ rstd_o) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && training && nbt) *nbt += 1;
    if (c >= C) return;
    // `center` (nullable): the statistics in `sums` are those of x - center[c] and the returned mean / shift
    // refer to that centred operand; running_mean is still updated with the mean of x itself.
    const double ctr = center ? (double)center[c] : 0.0;   // read before running_mean is updated (may alias)
    double mean, var;
    if (training) {
        mean = sums[c] / count;
        var = sums[C + c] / count - mean * mean;
        if (var < 0.0) var = 0.0;
        if (rmean) {
            const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
            rmean[c] = (float)((1.0 - momentum) * (double)rmean[c] + momentum * (mean + ctr));
            rvar[c] = (float)((1.0 - momentum) * (double)rvar[c] + momentum * unb);
        }
    } else {
        mean = (double)rmean[c] - ctr;
        var = rvar[c];
    }
    const double rstd = 1.0 / sqrt(var + (double)eps);
    const double g = gamma ? (double)gamma[c] : 1.0, b = beta ? (double)beta[c] : 0.0;
    scale[c] = (float)(g * rstd);
    shift[c] = (float)(b - mean * g * rstd);
    mean_o[c] = (float)mean;
    rstd_o[c] = (float)rstd;
}

// ------------------------------------------------------------------ fused statistic finalisers
// ONE launch instead of reduce_partials (+ peer all-reduce) + finalize: 2C/8 blocks each sum 8 columns of the producers'
// per-block partials (32 row lanes, fixed order -> deterministic), the LAST block to finish (self-resetting ticket) gathers
// the 2C sums, (data parallel) all-reduces them over NVLink peer memory inside the same kernel, and finishes the per-channel
// arithmetic.  A single-block version of the reduction took 20 us (one SM pulling 1.2 MB of partials through its L2 port).
constexpr int FIN_THREADS = 256;

// phase 1: this block's 8 columns -> sums_g; returns true in the last block to finish, with all 2C sums in sums_sh
__device__ __forceinline__ bool finalize_phase1(const double* __restrict__ partial, int nparts, int n, double* __restrict__ sums_g,
                                                unsigned* __restrict__ ticket, double* sums_sh) {
    __shared__ double red[32][9];
    __shared__ unsigned is_last;
    const int cl = threadIdx.x & 7, rl = threadIdx.x >> 3;
    const int j = blockIdx.x * 8 + cl;
    double s = 0.0;
    if (j < n) {
        int b = rl;
        for (; b + 96 < nparts; b += 128) {      // four independent loads in flight per thread
            const double v0 = partial[(long)b * n + j], v1 = partial[(long)(b + 32) * n + j];
            const double v2 = partial[(long)(b + 64) * n + j], v3 = partial[(long)(b + 96) * n + j];
            s += (v0 + v1) + (v2 + v3);
        }
        for (; b < nparts; b += 32) s += partial[(long)b * n + j];
    }
    red[rl][cl] = s;
    __syncthreads();
    if (rl == 0 && j < n) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 32; ++k) t += red[k][cl];
        sums_g[j] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(ticket, 1u);
        is_last = (prev == gridDim.x - 1) ? 1u : 0u;
        if (is_last) *ticket = 0u;              // self-resetting: the next launch on this stream starts from zero
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    for (int i = threadIdx.x; i < n; i += blockDim.x) sums_sh[i] = __ldcg(sums_g + i);
    __syncthreads();
    return true;
}

__global__ void __launch_bounds__(FIN_THREADS)
bn_stats_finalize_kernel(const double* __restrict__ partial, int nparts, double count, int C, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float* __restrict__ rmean, float* __restrict__ rvar,
                         int64_t* __restrict__ nbt, float momentum, float eps, const float* __restrict__ center, PeerCtx pc,
                         unsigned* __restrict__ ticket, double* __restrict__ sums_out, float* __restrict__ scale,
                         float* __restrict__ shift, float* __restrict__ mean_o, float* __restrict__ rstd_o) {
    __shared__ double sums_sh[256];
    __shared__ unsigned ep;
    if (!finalize_phase1(partial, nparts, 2 * C, sums_out, ticket, sums_sh)) return;
    if (pc.world > 1) {
        peer_allreduce_block(pc, sums_sh, 2 * C, &ep);
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sums_out[i] = sums_sh[i];
    }
    const int c = threadIdx.x;
    if (c == 0 && nbt) *nbt += 1;
    if (c >= C) return;
    const double ctr = center ? (double)center[c] : 0.0;   // read before running_mean is updated (may alias)
    const double mean = sums_sh[c] / count;
    double var = sums_sh[C + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    if (rmean) {
        const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
        rmean[c] = (float)((1.0 - momentum) * (double)rmean[c] + momentum * (mean + ctr));
        rvar[c] = (float)((1.0 - momentum) * (double)rvar[c] + momentum * unb);
    }
    const double rstd = 1.0 / sqrt(var + (double)eps);
    const double g = gamma ? (double)gamma[c] : 1.0, b = beta ? (double)beta[c] : 0.0;
    scale[c] = (float)(g * rstd);
    shift[c] = (float)(b - mean * g * rstd);
    mean_o[c] = (float)mean;
    rstd_o[c] = (float)rstd;
}

// BatchNorm-backward sums: partial = [nparts][2C] (sum dy, sum dy*xhat) followed by [nparts] per-block max|dy| (doubles)
__global__ void __launch_bounds__(FIN_THREADS)
bn_bwd_finalize_kernel(const double* __restrict__ partial, int nparts, int C, const float* __restrict__ scale, double count, PeerCtx pc,
                       unsigned* __restrict__ ticket, double* __restrict__ sums_out, float* __restrict__ dbeta,
                       float* __restrict__ dgamma, float* __restrict__ absmax_out, float* __restrict__ gscale_out) {
    __shared__ double sums_sh[256];
    __shared__ float mx[FIN_THREADS / 32], mx2[FIN_THREADS / 32];
    __shared__ unsigned ep;
    if (!finalize_phase1(partial, nparts, 2 * C, sums_out, ticket, sums_sh)) return;
    if (pc.world > 1) {
        peer_allreduce_block(pc, sums_sh, 2 * C, &ep);
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sums_out[i] = sums_sh[i];
    }
    const int c = threadIdx.x;
    if (c < 2 * C) {
        if (dbeta && c < C) dbeta[c] = (float)sums_sh[c];
        if (dgamma && c >= C) dgamma[c - C] = (float)sums_sh[c];
    }
    if (!absmax_out && !gscale_out) return;
    // max|dy| over the blocks and max_c|scale_c| (both local: the operand scale only has to be undone by the same rank)
    const double* pmax = partial + (size_t)nparts * 2 * C;
    float m = 0.f, sm = 0.f;
    for (int i = threadIdx.x; i < nparts; i += FIN_THREADS) m = fmaxf(m, (float)pmax[i]);
    for (int i = threadIdx.x; i < C; i += FIN_THREADS) sm = fmaxf(sm, scale ? fabsf(scale[i]) : 1.f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        sm = fmaxf(sm, __shfl_xor_sync(0xffffffffu, sm, o));
    }
    if ((threadIdx.x & 31) == 0) { mx[threadIdx.x >> 5] = m; mx2[threadIdx.x >> 5] = sm; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float am = 0.f, smax = 0.f;
        for (int w = 0; w < FIN_THREADS / 32; ++w) { am = fmaxf(am, mx[w]); smax = fmaxf(smax, mx2[w]); }
        if (absmax_out) *absmax_out = am;
        if (gscale_out) {
            const double bound = (double)smax * (double)am * (count > 0.0 ? 2.0 + sqrt(count) : 1.0);
            int e = 0;
            if (bound > 0.0 && isfinite(bound)) {
                e = (int)floor(log2(16384.0 / bound));
                e = max(-60, min(60, e));
            }
            gscale_out[0] = (float)ldexp(1.0, e);
            gscale_out[1] = (float)ldexp(1.0, -e);
        }
    }
}

// ------------------------------------------------------------------ NCL -> panel pack
// thread <-> (spectrogram s, panel q, frame t): 8 coalesced loads (one per channel of the
// panel), one 16-byte store; consecutive threads walk t so both sides are coalesced.
__global__ void __launch_bounds__(256)
ncl_pack_kernel(SpecSrc src, int S, int C, int L, const float* __restrict__ scale, const float* __restrict__ shift,
                uint4* __restrict__ panel, long panel_rows, int Lp, int pad, int fmt) {
    const long total = (long)S * (C / 8) * L;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int t = (int)(i % L);
        const long sq = i / L;
        const int q = (int)(sq % (C / 8));
        const long s = sq / (C / 8);
        long rs;
        const float* base = src.base(s, C, L, rs);   // sets rs: keep it a separate statement
        base += (long)(q * 8) * rs + t;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(base + (long)j * rs);
        unsigned short h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = q * 8 + j;
            const float y = scale ? fmaf(v[j], scale[c], shift[c]) : v[j];
            h[j] = cvt_f32_to16(y, fmt);
        }
        uint4 o;
        o.x = h[0] | ((unsigned)h[1] << 16);
        o.y = h[2] | ((unsigned)h[3] << 16);
        o.z = h[4] | ((unsigned)h[5] << 16);
        o.w = h[6] | ((unsigned)h[7] << 16);
        panel[(long)q * panel_rows + s * Lp + pad + t] = o;
    }
}

// ------------------------------------------------------------------ single-pass input kernel
// u = x - center[c] (fp32) -> 16-bit panel, and per-channel sum / sum of squares of u in the same pass.
// block (q, g): panel q (8 channels), spectrograms g, g+G, ...; warps split the spectrograms, lanes walk
// the frames, so every lane keeps private partial sums for its panel's 8 channels (no atomics).
// 3 blocks per SM: at 101 registers (2 blocks) only 16 warps x 64 B per lane were in flight -- 74 % of the copy bandwidth
__global__ void __launch_bounds__(256, 3)
ncl_center_pack_stats_kernel(SpecSrc src, int S, int C, int L, const float* __restrict__ center, uint4* __restrict__ panel,
                             long panel_rows, int Lp, int pad, int fmt, double* __restrict__ partial /* [G][2][C] */) {
    __shared__ double red[8][16];
    const int npan = C / 8;
    const int q = blockIdx.x % npan, gi = blockIdx.x / npan, G = gridDim.x / npan;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float ctr[8], a1[8], a2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { ctr[j] = center ? center[q * 8 + j] : 0.f; a1[j] = a2[j] = 0.f; }
    for (int s = gi + w * G; s < S; s += 8 * G) {
        long rs;
        const float* base = src.base(s, C, L, rs);   // sets rs: keep it a separate statement
        base += (long)(q * 8) * rs;
        uint4* prow = panel + (long)q * panel_rows + (long)s * Lp + pad;
        for (int t0 = 0; t0 < L; t0 += 64) {   // two frames per lane in flight: 16 independent loads
            float v[2][8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int t = t0 + h * 32 + lane;
#pragma unroll
                for (int j = 0; j < 8; ++j) v[h][j] = t < L ? __ldg(base + (long)j * rs + t) : 0.f;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int t = t0 + h * 32 + lane;
                if (t >= L) continue;
                unsigned short hh[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float u = v[h][j] - ctr[j];
                    a1[j] += u;
                    a2[j] = fmaf(u, u, a2[j]);
                    hh[j] = cvt_f32_to16(u, fmt);
                }
                uint4 o;
                o.x = hh[0] | ((unsigned)hh[1] << 16);
                o.y = hh[2] | ((unsigned)hh[3] << 16);
                o.z = hh[4] | ((unsigned)hh[5] << 16);
                o.w = hh[6] | ((unsigned)hh[7] << 16);
                prow[t] = o;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double s1 = warp_sum_d((double)a1[j]), s2 = warp_sum_d((double)a2[j]);
        if (lane == 0) { red[w][j] = s1; red[w][8 + j] = s2; }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
        const int which = threadIdx.x >> 3, j = threadIdx.x & 7;
        partial[((long)gi * 2 + which) * C + q * 8 + j] = t;
    }
}

// ------------------------------------------------------------------ [rows, C] tile kernels
constexpr int TR = 32;          // rows per tile
constexpr int TLD = 129;        // padded smem row stride (C <= 128)

// y = scale*z + shift -> panel rows (s*Lp + pad + p) and/or fp32 y
__global__ void __launch_bounds__(256)
affine_pack_kernel(const float* __restrict__ z, long rows, int P, int C, const float* __restrict__ scale,
                   const float* __restrict__ shift, uint4* __restrict__ panel, long panel_rows, int Lp, int pad,
                   int fmt, float* __restrict__ y) {
    __shared__ float tile[TR][TLD];
    const int tid = threadIdx.x;
    for (long r0 = (long)blockIdx.x * TR; r0 < rows; r0 += (long)gridDim.x * TR) {
        const int C4 = C >> 2;  // C % 4 == 0 (checked by the host wrapper)
        constexpr int NIT = TR * 32 / 256;
        float4 v4[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {   // issue every load of the tile first
            const int e = tid + it * 256;
            const int rl = e / C4, c = (e - rl * C4) * 4;
            const long r = r0 + rl;
            v4[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < TR * C4 && r < rows) v4[it] = __ldg(reinterpret_cast<const float4*>(z + r * C + c));
        }
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int e = tid + it * 256;
            if (e >= TR * C4) continue;
            const int rl = e / C4, c = (e - rl * C4) * 4;
            const long r = r0 + rl;
            float4 v = v4[it];
            if (r < rows) {
                if (scale) {
                    const float4 sc = *reinterpret_cast<const float4*>(scale + c), sh = *reinterpret_cast<const float4*>(shift + c);
                    v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                }
                if (y) *reinterpret_cast<float4*>(y + r * C + c) = v;
            }
            tile[rl][c] = v.x; tile[rl][c + 1] = v.y; tile[rl][c + 2] = v.z; tile[rl][c + 3] = v.w;
        }
        __syncthreads();
        if (panel) {
            // thread <-> (row lane, panel q): lanes walk rows so each warp store is contiguous
            const int rl = tid & 31;
            const long r = r0 + rl;
            if (r < rows) {
                const long s = r / P;
                const int p = (int)(r - s * P);
                for (int q = tid >> 5; q < C / 8; q += 8) {
                    unsigned short h[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) h[j] = cvt_f32_to16(tile[rl][q * 8 + j], fmt);
                    uint4 o;
                    o.x = h[0] | ((unsigned)h[1] << 16);
                    o.y = h[2] | ((unsigned)h[3] << 16);
                    o.z = h[4] | ((unsigned)h[5] << 16);
                    o.w = h[6] | ((unsigned)h[7] << 16);
                    panel[(long)q * panel_rows + s * Lp + pad + p] = o;
                }
            }
        }
        __syncthreads();
    }
}

// tp[s, c] = mean_p (scale*z[s,p,c] + shift)
__global__ void time_mean_kernel(const float* __restrict__ z, int S, int P, int C, const float* __restrict__ scale,
                                 const float* __restrict__ shift, float* __restrict__ tp, int ldtp) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= (long)S * C) return;
    const int c = (int)(i % C);
    const long s = i / C;
    float a = 0.f;
    for (int p = 0; p < P; ++p) a += z[(s * P + p) * C + c];
    a /= (float)P;
    tp[s * ldtp + c] = scale ? fmaf(a, scale[c], shift[c]) : a;
}

// the same, four channels per thread (16-byte accesses; the scalar kernel spent 13 us on 8.6 MB at S = 21 504)
__global__ void __launch_bounds__(256)
time_mean4_kernel(const float* __restrict__ z, int S, int P, int C, const float* __restrict__ scale,
                  const float* __restrict__ shift, float* __restrict__ tp, int ldtp) {
    const int C4 = C >> 2;
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= (long)S * C4) return;
    const int c = (int)(i % C4) * 4;
    const long s = i / C4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < P; ++p) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(z + (s * P + p) * C + c));
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    const float inv = (float)P;
    a.x /= inv; a.y /= inv; a.z /= inv; a.w /= inv;          // same rounding as the scalar kernel (a / P)
    if (scale) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c)), sh = __ldg(reinterpret_cast<const float4*>(shift + c));
        a.x = fmaf(a.x, sc.x, sh.x); a.y = fmaf(a.y, sc.y, sh.y); a.z = fmaf(a.z, sc.z, sh.z); a.w = fmaf(a.w, sc.w, sh.w);
    }
    *reinterpret_cast<float4*>(tp + s * ldtp + c) = a;
}

// BN backward reductions: per block partial of sum dy, sum dy*xhat  (C <= 128, C % 4 == 0)
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ dtp, int lddtp, const float* __restrict__ z,
                     const float* __restrict__ mean, const float* __restrict__ rstd, long rows, int P, int C,
                     double* __restrict__ partial /* [grid][2][C] */) {
    __shared__ float red[2][8][128];
    __shared__ float redm[8];
    const int cl = (threadIdx.x & 31) * 4, rg = threadIdx.x >> 5;
    float a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
    float amax = 0.f;
    if (cl < C) {
        float m[4], rs[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { m[j] = mean[cl + j]; rs[j] = rstd[cl + j]; }
        const float invP = 1.f / (float)P;
        // rows are walked from the END: dy was written front to back by the data-gradient kernel just before, so its last
        // ~100 MB are still in the 126 MB L2 when this sweep starts
        for (long rf = (long)blockIdx.x * 8 + rg; rf < rows; rf += (long)gridDim.x * 8) {
            const long r = rows - 1 - rf;
            const float4 g = *reinterpret_cast<const float4*>(dy + r * lddy + cl);
            const float4 zz = *reinterpret_cast<const float4*>(z + r * C + cl);
            float gv[4] = {g.x, g.y, g.z, g.w};
            const float zv[4] = {zz.x, zz.y, zz.z, zz.w};
            if (dtp) {
                const long s = r / P;
#pragma unroll
                for (int j = 0; j < 4; ++j) gv[j] += dtp[s * lddtp + cl + j] * invP;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a1[j] += gv[j];
                a2[j] = fmaf(gv[j], (zv[j] - m[j]) * rs[j], a2[j]);
                amax = fmaxf(amax, fabsf(gv[j]));
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (cl + j < 128) { red[0][rg][cl + j] = a1[j]; red[1][rg][cl + j] = a2[j]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0) redm[rg] = amax;
    __syncthreads();
    if (threadIdx.x == 0) {
        float mm = redm[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) mm = fmaxf(mm, redm[k]);
        partial[(long)gridDim.x * 2 * C + blockIdx.x] = (double)mm;   // after the [grid][2][C] sums
    }
    const int c = threadIdx.x & 127, which = threadIdx.x >> 7;
    if (c < C) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += (double)red[which][k][c];
        partial[((long)blockIdx.x * 2 + which) * C + c] = s;
    }
}

// dz = relu'(z) * scale * (dy - s1/n - xhat*s2/n); unpool into the dY panel or dense output
__global__ void __launch_bounds__(256)
bn_relu_unpool_bwd_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ dtp, int lddtp,
                          const float* __restrict__ z, const uint8_t* __restrict__ code, const float* __restrict__ scale,
                          const float* __restrict__ mean, const float* __restrict__ rstd, const double* __restrict__ sums,
                          double count, long rows, int P, int C, int pool, int Lp, uint4* __restrict__ panel,
                          long panel_rows, int fmt, const float* __restrict__ gscale, float* __restrict__ dz_out,
                          double* __restrict__ bias_partial) {
    __shared__ float tile[TR][TLD];
    __shared__ uint8_t ctile[TR][TLD + 3];
    __shared__ long prow_s[TR];
    __shared__ float k_s1[128], k_mean[128], k_rs2[128], k_scale[128];
    const int tid = threadIdx.x;
    const float gs = gscale ? gscale[0] : 1.f;
    if (tid < 128) {
        // dz = scale * (dy - s1/n - (z - mean) * rstd * s2/n): per-channel constants once per block
        const bool on = tid < C;
        const double inv_n = 1.0 / count;
        k_s1[tid] = (on && sums) ? (float)(sums[tid] * inv_n) : 0.f;
        k_mean[tid] = (on && sums) ? mean[tid] : 0.f;
        k_rs2[tid] = (on && sums) ? (float)((double)rstd[tid] * sums[C + tid] * inv_n) : 0.f;
        k_scale[tid] = (on && scale) ? scale[tid] : 1.f;
    }
    __syncthreads();
    double bias_acc = 0.0;  // thread tid<C accumulates channel tid
    const float invP = 1.f / (float)P;
    for (long r0 = (long)blockIdx.x * TR; r0 < rows; r0 += (long)gridDim.x * TR) {
        const int C4 = C >> 2;  // C % 4 == 0 (checked by the host wrapper)
        if (tid < TR) {  // first panel row of each pooled row of this tile (-1 = beyond the end)
            const long r = r0 + tid;
            const long s = r / P;
            prow_s[tid] = r < rows ? s * Lp + (r - s * P) * pool : -1;
        }
        // all global loads of the tile are issued before any is consumed: 12 independent requests per thread
        // in flight (Little's law: the one-load-at-a-time version sat at 2.5 TB/s)
        constexpr int NIT = TR * 32 / 256;   // 4 iterations cover 32 rows x (C/4 <= 32) float4 columns
        float4 g4[NIT], z4[NIT];
        uchar4 c4[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int e = tid + it * 256;
            const int rl = e / C4, c = (e - rl * C4) * 4;
            const long r = r0 + rl;
            g4[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            z4[it] = g4[it];
            c4[it] = make_uchar4(0, 0, 0, 0);
            if (e < TR * C4 && r < rows) {
                g4[it] = __ldg(reinterpret_cast<const float4*>(dy + r * lddy + c));
                z4[it] = __ldg(reinterpret_cast<const float4*>(z + r * C + c));
                if (code) c4[it] = *reinterpret_cast<const uchar4*>(code + r * C + c);
            }
        }
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int e = tid + it * 256;
            if (e >= TR * C4) continue;
            const int rl = e / C4, c = (e - rl * C4) * 4;
            const long r = r0 + rl;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            const uint8_t cd[4] = {c4[it].x, c4[it].y, c4[it].z, c4[it].w};
            if (r < rows) {
                float g[4] = {g4[it].x, g4[it].y, g4[it].z, g4[it].w};
                const float zz[4] = {z4[it].x, z4[it].y, z4[it].z, z4[it].w};
                if (dtp) {
                    const float* dp = dtp + (r / P) * lddtp + c;
#pragma unroll
                    for (int j = 0; j < 4; ++j) g[j] += dp[j] * invP;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float gg = g[j];
                    if (sums) gg = gg - k_s1[c + j] - (zz[j] - k_mean[c + j]) * k_rs2[c + j];
                    gg *= k_scale[c + j];
                    v[j] = zz[j] > 0.f ? gg : 0.f;
                }
                if (dz_out) *reinterpret_cast<float4*>(dz_out + r * C + c) = make_float4(v[0], v[1], v[2], v[3]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) { tile[rl][c + j] = v[j]; ctile[rl][c + j] = cd[j]; }
        }
        __syncthreads();
        if (panel) {
            // consecutive lanes write consecutive 16-byte chunks of one panel: chunk = (pooled row, pool slot),
            // so every store instruction covers 512 contiguous bytes (rows of a tile are consecutive in the panel)
            const int lane = tid & 31;
            const int chunks = TR * pool;
            const int sh = pool == 4 ? 2 : (pool == 2 ? 1 : 0);
            for (int q = tid >> 5; q < C / 8; q += 8) {
                for (int cidx = lane; cidx < chunks; cidx += 32) {
                    const int rl = pool == 8 ? cidx >> 3 : cidx >> sh;
                    const int pi = cidx - (pool == 8 ? rl << 3 : rl << sh);
                    const long prow = prow_s[rl];
                    if (prow < 0) continue;
                    unsigned short w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        w[j] = ctile[rl][q * 8 + j] == pi ? cvt_f32_to16(tile[rl][q * 8 + j] * gs, fmt) : (unsigned short)0;
                    uint4 o;
                    o.x = w[0] | ((unsigned)w[1] << 16);
                    o.y = w[2] | ((unsigned)w[3] << 16);
                    o.z = w[4] | ((unsigned)w[5] << 16);
                    o.w = w[6] | ((unsigned)w[7] << 16);
                    panel[(long)q * panel_rows + prow + pi] = o;
                }
            }
        }
        if (bias_partial && tid < C) {
            float s = 0.f;
#pragma unroll 8
            for (int k = 0; k < TR; ++k) s += tile[k][tid];
            bias_acc += (double)s;
        }
        __syncthreads();
    }
    if (bias_partial && tid < C) bias_partial[(long)blockIdx.x * C + tid] = bias_acc;
}

__global__ void reduce_max_kernel(const double* __restrict__ pmax, int nblk, float* __restrict__ out) {
    __shared__ float sm[256];
    float m = 0.f;
    for (int i = threadIdx.x; i < nblk; i += 256) m = fmaxf(m, (float)pmax[i]);
    sm[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] = fmaxf(sm[threadIdx.x], sm[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sm[0];
}

// gradient scale for the 16-bit conv-backward operand: the largest power of two s with
// s * bound <= 2^14, bound >= max|dz| from |xhat| <= sqrt(n), mean|xhat| <= 1:
//   |dz| <= max_c|scale_c| * max|dy| * (2 + sqrt(n))   (batch statistics)   or   max_c|scale_c| * max|dy|
__global__ void grad_scale_kernel(const float* __restrict__ absmax, const float* __restrict__ scale, int C, double count,
                                  float* __restrict__ out /* [2]: s, 1/s */) {
    __shared__ float sm[128];
    float m = 0.f;
    for (int c = threadIdx.x; c < C; c += 128) m = fmaxf(m, scale ? fabsf(scale[c]) : 1.f);
    sm[threadIdx.x] = m;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] = fmaxf(sm[threadIdx.x], sm[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double bound = (double)sm[0] * (double)(*absmax) * (count > 0.0 ? 2.0 + sqrt(count) : 1.0);
        int e = 0;
        if (bound > 0.0 && isfinite(bound)) {
            e = (int)floor(log2(16384.0 / bound));
            e = max(-60, min(60, e));
        }
        out[0] = (float)ldexp(1.0, e);
        out[1] = (float)ldexp(1.0, -e);
    }
}

__global__ void cvt_f64_f32_kernel(const double* __restrict__ in, int n, double mul, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)(in[i] * mul);
}

// bn0 backward reductions: dx is channels-last [S*L, C] (L1 dgrad), x is the NCL fp32 input.
// block (g, q) owns the 8 channels of panel q for spectrograms g, g+G, ...
__global__ void __launch_bounds__(128)
ncl_bn_bwd_reduce_kernel(const float* __restrict__ dx, const float* __restrict__ pos, int S_pos,
                         const float* __restrict__ neg, int S_neg, int C, int L, const float* __restrict__ mean,
                         const float* __restrict__ rstd, double* __restrict__ partial /* [G][2][C] */) {
    __shared__ float red[4][16];
    const int q = blockIdx.x % (C / 8), gi = blockIdx.x / (C / 8), G = gridDim.x / (C / 8);
    const int S = S_pos + S_neg;
    float m[8], rs[8], a1[8], a2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { m[j] = mean[q * 8 + j]; rs[j] = rstd[q * 8 + j]; a1[j] = a2[j] = 0.f; }
    double d1[8], d2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) d1[j] = d2[j] = 0.0;
    for (int s = gi; s < S; s += G) {
        const float* xb = (s < S_pos ? pos + (long)s * C * L : neg + (long)(s - S_pos) * C * L) + (long)(q * 8) * L;
        for (int t = threadIdx.x; t < L; t += 128) {
            const float4 g0 = *reinterpret_cast<const float4*>(dx + ((long)s * L + t) * C + q * 8);
            const float4 g1 = *reinterpret_cast<const float4*>(dx + ((long)s * L + t) * C + q * 8 + 4);
            const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xv = __ldg(xb + (long)j * L + t);
                a1[j] += gv[j];
                a2[j] = fmaf(gv[j], (xv - m[j]) * rs[j], a2[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { d1[j] += (double)a1[j]; d2[j] += (double)a2[j]; a1[j] = a2[j] = 0.f; }
    }
    // block reduce the 16 doubles (as two float-free passes through smem of warp sums)
    __shared__ double dred[4][16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double v1 = warp_sum_d(d1[j]), v2 = warp_sum_d(d2[j]);
        if ((threadIdx.x & 31) == 0) { dred[threadIdx.x >> 5][j] = v1; dred[threadIdx.x >> 5][8 + j] = v2; }
    }
    (void)red;
    __syncthreads();
    if (threadIdx.x < 16) {
        const double v = dred[0][threadIdx.x] + dred[1][threadIdx.x] + dred[2][threadIdx.x] + dred[3][threadIdx.x];
        const int which = threadIdx.x >> 3, j = threadIdx.x & 7;
        partial[((long)gi * 2 + which) * C + q * 8 + j] = v;
    }
}

// ------------------------------------------------------------------ row-lane kernels (C == 128, panel output)
// lane = row, warp = 8-channel group (panel): a lane reads its 32-byte sector of the [rows,128] fp32 row and
// writes 16-byte panel chunks; consecutive lanes hold consecutive rows, so every warp store is one contiguous
// run of the panel and no shared-memory transpose is needed (the tiled kernels above were l1tex/smem bound:
// ncu 47 % l1tex, 37 % DRAM at layer-1 size).
constexpr int RL_THREADS = 512;     // 16 warps = the 16 panels of a 128-channel row

__device__ __forceinline__ uint4 pack8_16(const float (&v)[8], int fmt) {
    unsigned short h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = cvt_f32_to16(v[j], fmt);
    uint4 o;
    o.x = h[0] | ((unsigned)h[1] << 16);
    o.y = h[2] | ((unsigned)h[3] << 16);
    o.z = h[4] | ((unsigned)h[5] << 16);
    o.w = h[6] | ((unsigned)h[7] << 16);
    return o;
}

// y = scale*z + shift -> panel rows (s*Lp + pad + p)
template <int UNROLL>
__global__ void __launch_bounds__(RL_THREADS)
affine_pack_rows_kernel(const float* __restrict__ z, long rows, int P, const float* __restrict__ scale,
                        const float* __restrict__ shift, uint4* __restrict__ panel, long panel_rows, int Lp, int pad, int fmt) {
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = scale ? scale[q * 8 + j] : 1.f; sh[j] = shift ? shift[q * 8 + j] : 0.f; }
    uint4* dst = panel + (long)q * panel_rows + pad;
    for (long r0 = (long)blockIdx.x * (32 * UNROLL); r0 < rows; r0 += (long)gridDim.x * (32 * UNROLL)) {
        float4 a[UNROLL], b[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long r = r0 + u * 32 + lane;
            a[u] = b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < rows) {
                const float4* src = reinterpret_cast<const float4*>(z + r * 128 + q * 8);
                a[u] = __ldg(src);
                b[u] = __ldg(src + 1);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long r = r0 + u * 32 + lane;
            if (r >= rows) continue;
            const float x[8] = {a[u].x, a[u].y, a[u].z, a[u].w, b[u].x, b[u].y, b[u].z, b[u].w};
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(x[j], sc[j], sh[j]);
            const long s = (long)((unsigned)r / (unsigned)P);   // rows < 2^31 (checked by the host wrapper)
            dst[s * Lp + (r - s * P)] = pack8_16(v, fmt);
        }
    }
}

// dz = relu'(z) * scale * (dy - s1/n - xhat*s2/n) as one FMA pair per element (a*dy + b*z + c with per-channel
// a = scale, b = -scale*rstd*s2/n, c = scale*(rstd*s2/n*mean - s1/n)); the POOL rows of each window are written
// as POOL 16-byte chunks (value where the argmax code matches, zero elsewhere) = 16*POOL contiguous bytes per lane.
template <int POOL, int UNROLL>
__global__ void __launch_bounds__(RL_THREADS)
bn_relu_unpool_rows_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ dtp, int lddtp,
                           const float* __restrict__ z, const uint8_t* __restrict__ code, const float* __restrict__ scale,
                           const float* __restrict__ mean, const float* __restrict__ rstd, const double* __restrict__ sums,
                           double count, long rows, int P, int Lp, uint4* __restrict__ panel, long panel_rows, int fmt,
                           const float* __restrict__ gscale, double* __restrict__ bias_partial) {
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float gs = gscale ? gscale[0] : 1.f;
    float ka[8], kb[8], kc[8], acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = q * 8 + j;
        const float sc = scale ? scale[c] : 1.f;
        ka[j] = sc; kb[j] = 0.f; kc[j] = 0.f; acc[j] = 0.f;
        if (sums) {
            const double inv_n = 1.0 / count;
            const double rs2 = (double)rstd[c] * sums[128 + c] * inv_n;
            kb[j] = (float)(-(double)sc * rs2);
            kc[j] = (float)((double)sc * (rs2 * (double)mean[c] - sums[c] * inv_n));
        }
    }
    const float invP = 1.f / (float)P;
    uint4* dst = panel + (long)q * panel_rows;
    for (long r0 = (long)blockIdx.x * (32 * UNROLL); r0 < rows; r0 += (long)gridDim.x * (32 * UNROLL)) {
        float4 g0[UNROLL], g1[UNROLL], z0[UNROLL], z1[UNROLL];
        uint2 cd[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {   // every load of the iteration is issued before any is consumed
            const long r = r0 + u * 32 + lane;
            g0[u] = g1[u] = z0[u] = z1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            cd[u] = make_uint2(0u, 0u);
            if (r < rows) {
                const float4* gp = reinterpret_cast<const float4*>(dy + r * lddy + q * 8);
                const float4* zp = reinterpret_cast<const float4*>(z + r * 128 + q * 8);
                g0[u] = __ldg(gp); g1[u] = __ldg(gp + 1);
                z0[u] = __ldg(zp); z1[u] = __ldg(zp + 1);
                cd[u] = __ldg(reinterpret_cast<const uint2*>(code + r * 128 + q * 8));
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long r = r0 + u * 32 + lane;
            if (r >= rows) continue;
            float g[8] = {g0[u].x, g0[u].y, g0[u].z, g0[u].w, g1[u].x, g1[u].y, g1[u].z, g1[u].w};
            const float zz[8] = {z0[u].x, z0[u].y, z0[u].z, z0[u].w, z1[u].x, z1[u].y, z1[u].z, z1[u].w};
            const long s = (long)((unsigned)r / (unsigned)P);   // rows < 2^31 (checked by the host wrapper)
            if (dtp) {
                const float4* tp = reinterpret_cast<const float4*>(dtp + s * lddtp + q * 8);
                const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1);
                g[0] = fmaf(t0.x, invP, g[0]); g[1] = fmaf(t0.y, invP, g[1]); g[2] = fmaf(t0.z, invP, g[2]); g[3] = fmaf(t0.w, invP, g[3]);
                g[4] = fmaf(t1.x, invP, g[4]); g[5] = fmaf(t1.y, invP, g[5]); g[6] = fmaf(t1.z, invP, g[6]); g[7] = fmaf(t1.w, invP, g[7]);
            }
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float gg = fmaf(ka[j], g[j], fmaf(kb[j], zz[j], kc[j]));
                v[j] = zz[j] > 0.f ? gg : 0.f;
                acc[j] += v[j];
            }
            unsigned short h[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) h[j] = cvt_f32_to16(v[j] * gs, fmt);
            const unsigned cw[2] = {cd[u].x, cd[u].y};
            uint4* o = dst + s * Lp + (r - s * P) * POOL;
#pragma unroll
            for (int pi = 0; pi < POOL; ++pi) {
                unsigned w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) w[j] = ((cw[j >> 2] >> (8 * (j & 3))) & 0xffu) == (unsigned)pi ? (unsigned)h[j] : 0u;
                o[pi] = make_uint4(w[0] | (w[1] << 16), w[2] | (w[3] << 16), w[4] | (w[5] << 16), w[6] | (w[7] << 16));
            }
        }
    }
    if (bias_partial) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float t = warp_sum(acc[j]);
            if (lane == 0) bias_partial[(long)blockIdx.x * 128 + q * 8 + j] = (double)t;
        }
    }
}

int rows_grid(long rows, int rows_per_block, int blocks_per_sm) {
    const long tiles = (rows + rows_per_block - 1) / rows_per_block;
    const long cap = (long)dcue_num_sms() * blocks_per_sm;
    return (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
}

int tile_grid(long rows) {
    long tiles = (rows + TR - 1) / TR;
    long cap = (long)dcue_num_sms() * 6;
    return (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
}
constexpr int STAT_BLOCKS_PER_SM = 4;

}  // namespace

extern "C" size_t dcue_ncl_stats_ws_bytes(int C) {
    return (size_t)dcue_num_sms() * STAT_BLOCKS_PER_SM * 2 * (size_t)C * sizeof(double) + 256;
}

static int ncl_stats_impl(const SpecSrc& src, int S, int C, int L, double* sums, void* ws, size_t ws_bytes, cudaStream_t st) {
    int grid = dcue_num_sms() * STAT_BLOCKS_PER_SM;
    if (grid > S) grid = S > 0 ? S : 1;
    if (ws_bytes < (size_t)grid * 2 * C * sizeof(double)) DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_ncl_stats: workspace too small");
    ncl_stats_kernel<<<grid, 256, 0, st>>>(src, S, C, L, (double*)ws);
    DCUE_LAUNCH_CHECK();
    reduce_partials_kernel<<<ceil_div_i(2 * C, 8), 256, 0, st>>>((const double*)ws, grid, 2 * C, sums);
    DCUE_LAUNCH_CHECK();
    return 0;
}

extern "C" int dcue_ncl_stats(const float* pos, int S_pos, const float* neg, int S_neg, int C, int L, double* sums,
                              void* ws, size_t ws_bytes, void* stream) {
    DCUE_CHECK_ARG(sums && ws && S_pos >= 0 && S_neg >= 0 && (pos || S_pos == 0) && (neg || S_neg == 0) && L > 0);
    DCUE_CHECK_ARG(C > 0 && C <= 8 * STAT_MAXC_PER_WARP);
    SpecSrc src{pos, neg, S_pos, nullptr, nullptr, 0, 0, nullptr};
    return ncl_stats_impl(src, S_pos + S_neg, C, L, sums, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int dcue_ncl_stats_indexed(const float* pool, long n_songs, long T, const int64_t* idx, const int32_t* off, int S,
                                      int C, int L, int* err_flag, double* sums, void* ws, size_t ws_bytes, void* stream) {
    DCUE_CHECK_ARG(pool && idx && sums && ws && err_flag && S >= 0 && n_songs > 0 && L > 0 && T >= L);
    DCUE_CHECK_ARG(C > 0 && C <= 8 * STAT_MAXC_PER_WARP);
    SpecSrc src{pool, nullptr, 0, idx, off, T, n_songs, err_flag};
    return ncl_stats_impl(src, S, C, L, sums, ws, ws_bytes, (cudaStream_t)stream);
}

static int center_pack_groups(int S, int C) {
    const int npan = C / 8;
    int G = dcue_num_sms() * 3 / npan;     // 3 resident blocks per SM (80 registers): one wave
    if (G > (S + 7) / 8) G = (S + 7) / 8;
    return G < 1 ? 1 : G;
}

static int center_pack_impl(const SpecSrc& src, int S, int C, int L, const float* center, void* panel, long panel_rows,
                            int Lp, int pad, int fmt, double* sums, void* ws, size_t ws_bytes, cudaStream_t st) {
    const int npan = C / 8;
    const int G = center_pack_groups(S, C);
    if (ws_bytes < (size_t)G * 2 * C * sizeof(double)) DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_ncl_center_pack_stats: workspace too small");
    ncl_center_pack_stats_kernel<<<G * npan, 256, 0, st>>>(src, S, C, L, center, (uint4*)panel, panel_rows, Lp, pad, fmt,
                                                           (double*)ws);
    DCUE_LAUNCH_CHECK();
    if (sums) {   // sums == NULL: the G partial rows stay at the start of ws for dcue_bn_stats_finalize
        reduce_partials_kernel<<<ceil_div_i(2 * C, 8), 256, 0, st>>>((const double*)ws, G, 2 * C, sums);
        DCUE_LAUNCH_CHECK();
    }
    return 0;
}


extern "C" int dcue_ncl_center_pack_stats(const float* pos, int S_pos, const float* neg, int S_neg, int C, int L,
                                          const float* center, void* panel, long panel_rows, int Lp, int pad, int fmt,
                                          double* sums, void* ws, size_t ws_bytes, void* stream) {
    DCUE_CHECK_ARG(panel && ws && S_pos >= 0 && S_neg >= 0 && (pos || S_pos == 0) && (neg || S_neg == 0));
    DCUE_CHECK_ARG(C > 0 && C % 8 == 0 && C <= 128 && L > 0 && Lp >= L + pad && pad >= 0 && panel_rows >= (long)(S_pos + S_neg) * Lp);
    SpecSrc src{pos, neg, S_pos, nullptr, nullptr, 0, 0, nullptr};
    return center_pack_impl(src, S_pos + S_neg, C, L, center, panel, panel_rows, Lp, pad, fmt, sums, ws, ws_bytes,
                            (cudaStream_t)stream);
}

extern "C" int dcue_ncl_center_pack_stats_indexed(const float* pool, long n_songs, long T, const int64_t* idx,
                                                  const int32_t* off, int S, int C, int L, int* err_flag, const float* center,
                                                  void* panel, long panel_rows, int Lp, int pad, int fmt, double* sums,
                                                  void* ws, size_t ws_bytes, void* stream) {
    DCUE_CHECK_ARG(pool && idx && panel && ws && err_flag && S >= 0 && n_songs > 0 && T >= L);
    DCUE_CHECK_ARG(C > 0 && C % 8 == 0 && C <= 128 && L > 0 && Lp >= L + pad && pad >= 0 && panel_rows >= (long)S * Lp);
    SpecSrc src{pool, nullptr, 0, idx, off, T, n_songs, err_flag};
    return center_pack_impl(src, S, C, L, center, panel, panel_rows, Lp, pad, fmt, sums, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int dcue_bn_finalize(const double* sums, double count, int C, const float* gamma, const float* beta,
                                float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                                float eps, int training, const float* center, float* scale, float* shift, float* mean,
                                float* rstd, void* stream) {
    DCUE_CHECK_ARG(C > 0 && scale && shift && mean && rstd);
    DCUE_CHECK_ARG(training ? (sums != nullptr && count > 0) : (running_mean && running_var));
    bn_finalize_kernel<<<ceil_div_i(C, 128), 128, 0, (cudaStream_t)stream>>>(
        sums, count, C, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, training, center,
        scale, shift, mean, rstd);
    DCUE_LAUNCH_CHECK();
    return 0;
}

static int ncl_pack_impl(const SpecSrc& src, int S, int C, int L, const float* scale, const float* shift, void* panel,
                         long panel_rows, int Lp, int pad, int fmt, cudaStream_t st) {
    const long total = (long)S * (C / 8) * L;
    if (total == 0) return 0;
    long blocks = (total + 255) / 256;
    const long cap = (long)dcue_num_sms() * 32;
    if (blocks > cap) blocks = cap;
    ncl_pack_kernel<<<(int)blocks, 256, 0, st>>>(src, S, C, L, scale, shift, (uint4*)panel, panel_rows, Lp, pad, fmt);
    DCUE_LAUNCH_CHECK();
    return 0;
}

extern "C" int dcue_ncl_pack(const float* pos, int S_pos, const float* neg, int S_neg, int C, int L, const float* scale,
                             const float* shift, void* panel, long panel_rows, int Lp, int pad, int fmt, void* stream) {
    DCUE_CHECK_ARG(panel && S_pos >= 0 && S_neg >= 0 && (pos || S_pos == 0) && (neg || S_neg == 0));
    DCUE_CHECK_ARG(C > 0 && C % 8 == 0 && L > 0 && Lp >= L + pad && pad >= 0 && (!scale == !shift));
    DCUE_CHECK_ARG(panel_rows >= (long)(S_pos + S_neg) * Lp);
    SpecSrc src{pos, neg, S_pos, nullptr, nullptr, 0, 0, nullptr};
    return ncl_pack_impl(src, S_pos + S_neg, C, L, scale, shift, panel, panel_rows, Lp, pad, fmt, (cudaStream_t)stream);
}

extern "C" int dcue_ncl_pack_indexed(const float* pool, long n_songs, long T, const int64_t* idx, const int32_t* off, int S,
                                     int C, int L, int* err_flag, const float* scale, const float* shift, void* panel,
                                     long panel_rows, int Lp, int pad, int fmt, void* stream) {
    DCUE_CHECK_ARG(pool && idx && panel && err_flag && S >= 0 && n_songs > 0 && T >= L);
    DCUE_CHECK_ARG(C > 0 && C % 8 == 0 && L > 0 && Lp >= L + pad && pad >= 0 && (!scale == !shift));
    DCUE_CHECK_ARG(panel_rows >= (long)S * Lp);
    SpecSrc src{pool, nullptr, 0, idx, off, T, n_songs, err_flag};
    return ncl_pack_impl(src, S, C, L, scale, shift, panel, panel_rows, Lp, pad, fmt, (cudaStream_t)stream);
}

extern "C" int dcue_affine_pack(const float* z, int S, int P, int C, const float* scale, const float* shift, void* panel,
                                long panel_rows, int Lp, int pad, int fmt, float* y, float* tp, int ldtp, void* stream) {
    DCUE_CHECK_ARG(z && S >= 0 && P > 0 && C > 0 && C <= 128 && C % 4 == 0 && (!scale == !shift));
    DCUE_CHECK_ARG(!panel || (C % 8 == 0 && Lp >= P + pad && panel_rows >= (long)S * Lp));
    cudaStream_t st = (cudaStream_t)stream;
    const long rows = (long)S * P;
    if (rows == 0) return 0;
    if (panel && !y && C == 128 && ((uintptr_t)z & 15) == 0 && rows < (1L << 31)) {
        affine_pack_rows_kernel<2><<<rows_grid(rows, 64, 2), RL_THREADS, 0, st>>>(z, rows, P, scale, shift, (uint4*)panel,
                                                                                  panel_rows, Lp, pad, fmt);
        DCUE_LAUNCH_CHECK();
    } else if (panel || y) {
        affine_pack_kernel<<<tile_grid(rows), 256, 0, st>>>(z, rows, P, C, scale, shift, (uint4*)panel, panel_rows, Lp,
                                                            pad, fmt, y);
        DCUE_LAUNCH_CHECK();
    }
    if (tp) {
        DCUE_CHECK_ARG(ldtp >= C);
        if ((C & 3) == 0 && (ldtp & 3) == 0 && (((uintptr_t)z | (uintptr_t)tp) & 15) == 0 &&
            (!scale || (((uintptr_t)scale | (uintptr_t)shift) & 15) == 0))
            time_mean4_kernel<<<ceil_div_i((long)S * (C / 4), 256), 256, 0, st>>>(z, S, P, C, scale, shift, tp, ldtp);
        else
            time_mean_kernel<<<ceil_div_i((long)S * C, 256), 256, 0, st>>>(z, S, P, C, scale, shift, tp, ldtp);
        DCUE_LAUNCH_CHECK();
    }
    return 0;
}

static int bn_bwd_grid(long rows) {
    const long g = (rows + 7) / 8;
    const long cap = (long)dcue_num_sms() * 4;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

extern "C" size_t dcue_bn_bwd_reduce_nparts(int S, int P) { return (size_t)bn_bwd_grid((long)S * P); }
extern "C" size_t dcue_ncl_center_pack_stats_nparts(int S, int C) { return (size_t)center_pack_groups(S, C); }

extern "C" int dcue_bn_stats_finalize(const double* partial, int nparts, double count, int C, const float* gamma, const float* beta,
                                      float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                                      float eps, const float* center, const void* peer_bufs_dev, const void* peer_signals_dev,
                                      void* peer_counter, int rank, int world, void* ticket, double* sums_out, float* scale,
                                      float* shift, float* mean, float* rstd, void* stream) {
    DCUE_CHECK_ARG(partial && nparts > 0 && count > 0 && C > 0 && 2 * C <= 256 && scale && shift && mean && rstd && ticket && sums_out);
    DCUE_CHECK_ARG(world >= 1 && (world == 1 || (peer_bufs_dev && peer_signals_dev && peer_counter && rank >= 0 && rank < world)));
    DCUE_CHECK_ARG(2 * C <= PEER_SLOT_DOUBLES);
    PeerCtx pc{(double* const*)peer_bufs_dev, (unsigned* const*)peer_signals_dev, (unsigned*)peer_counter, rank, world};
    bn_stats_finalize_kernel<<<ceil_div_i(2 * C, 8), FIN_THREADS, 0, (cudaStream_t)stream>>>(
        partial, nparts, count, C, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, center, pc,
        (unsigned*)ticket, sums_out, scale, shift, mean, rstd);
    DCUE_LAUNCH_CHECK();
    return 0;
}

extern "C" int dcue_bn_bwd_finalize(const double* partial, int nparts, int C, const float* scale, double count,
                                    const void* peer_bufs_dev, const void* peer_signals_dev, void* peer_counter, int rank, int world,
                                    void* ticket, double* sums_out, float* dbeta, float* dgamma, float* absmax_out, float* gscale_out,
                                    void* stream) {
    DCUE_CHECK_ARG(partial && nparts > 0 && C > 0 && 2 * C <= 256 && sums_out && ticket);
    DCUE_CHECK_ARG(world >= 1 && (world == 1 || (peer_bufs_dev && peer_signals_dev && peer_counter && rank >= 0 && rank < world)));
    PeerCtx pc{(double* const*)peer_bufs_dev, (unsigned* const*)peer_signals_dev, (unsigned*)peer_counter, rank, world};
    bn_bwd_finalize_kernel<<<ceil_div_i(2 * C, 8), FIN_THREADS, 0, (cudaStream_t)stream>>>(partial, nparts, C, scale, count, pc,
                                                                                         (unsigned*)ticket, sums_out, dbeta, dgamma,
                                                                                         absmax_out, gscale_out);
    DCUE_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t dcue_bn_bwd_ws_bytes(int C) {
    return (size_t)dcue_num_sms() * 8 * (2 * (size_t)C + 1) * sizeof(double) + 256;
}

extern "C" int dcue_bn_bwd_reduce(const float* dy, int lddy, const float* dtp, int lddtp, const float* z, const float* mean,
                                  const float* rstd, int S, int P, int C, double* sums, float* absmax, float* dbeta,
                                  float* dgamma, void* ws, size_t ws_bytes, void* stream) {
    DCUE_CHECK_ARG(dy && z && mean && rstd && ws && S >= 0 && P > 0 && C > 0 && C <= 128 && C % 4 == 0);
    DCUE_CHECK_ARG(lddy >= C && lddy % 4 == 0 && ((uintptr_t)dy & 15) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const long rows = (long)S * P;
    const int grid = bn_bwd_grid(rows);
    if (ws_bytes < (size_t)grid * (2 * C + 1) * sizeof(double)) DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_bn_bwd_reduce: workspace too small");
    bn_bwd_reduce_kernel<<<grid, 256, 0, st>>>(dy, lddy, dtp, lddtp, z, mean, rstd, rows, P, C, (double*)ws);
    DCUE_LAUNCH_CHECK();
    if (!sums) return 0;      // partial mode: [grid][2C] sums + [grid] max|dy| stay in ws for dcue_bn_bwd_finalize
    reduce_partials_kernel<<<ceil_div_i(2 * C, 8), 256, 0, st>>>((const double*)ws, grid, 2 * C, sums, dbeta, dgamma, C);
    DCUE_LAUNCH_CHECK();
    if (absmax) {
        reduce_max_kernel<<<1, 256, 0, st>>>((const double*)ws + (size_t)grid * 2 * C, grid, absmax);
        DCUE_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int dcue_bn_relu_unpool_bwd(const float* dy, int lddy, const float* dtp, int lddtp, const float* z, const uint8_t* code,
                                       const float* scale, const float* mean, const float* rstd, const double* sums,
                                       double count, int S, int P, int C, int pool, int Lp, void* dy_panel,
                                       long panel_rows, int fmt, const float* gscale, float* dz_out, double* bias_sums,
                                       float* bias_out, void* ws, size_t ws_bytes, void* stream) {
    DCUE_CHECK_ARG(dy && z && S >= 0 && P > 0 && C > 0 && C <= 128 && C % 4 == 0 && (dy_panel || dz_out) && lddy >= C &&
                   lddy % 4 == 0 && ((uintptr_t)dy & 15) == 0);
    DCUE_CHECK_ARG(!sums || (mean && rstd && count > 0));
    DCUE_CHECK_ARG(!dy_panel || (code && C % 8 == 0 && (pool == 1 || pool == 2 || pool == 4 || pool == 8) && Lp >= P * pool && panel_rows >= (long)S * Lp));
    cudaStream_t st = (cudaStream_t)stream;
    const long rows = (long)S * P;
    if (rows == 0) return 0;
    const bool fast = dy_panel && !dz_out && C == 128 && (pool == 2 || pool == 4) && ((uintptr_t)z & 15) == 0 && rows < (1L << 31) &&
                      ((uintptr_t)code & 7) == 0 && (!dtp || (lddtp % 4 == 0 && ((uintptr_t)dtp & 15) == 0));
    const int grid = fast ? rows_grid(rows, 64, 1) : tile_grid(rows);
    if (bias_sums && (!ws || ws_bytes < (size_t)grid * C * sizeof(double)))
        DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_bn_relu_unpool_bwd: workspace too small");
    if (fast) {
        double* bp = bias_sums ? (double*)ws : nullptr;
        const double cnt = count > 0 ? count : 1.0;
        if (pool == 4)
            bn_relu_unpool_rows_kernel<4, 2><<<grid, RL_THREADS, 0, st>>>(dy, lddy, dtp, lddtp, z, code, scale, mean, rstd, sums, cnt,
                                                                         rows, P, Lp, (uint4*)dy_panel, panel_rows, fmt, gscale, bp);
        else
            bn_relu_unpool_rows_kernel<2, 2><<<grid, RL_THREADS, 0, st>>>(dy, lddy, dtp, lddtp, z, code, scale, mean, rstd, sums, cnt,
                                                                         rows, P, Lp, (uint4*)dy_panel, panel_rows, fmt, gscale, bp);
    } else {
        bn_relu_unpool_bwd_kernel<<<grid, 256, 0, st>>>(dy, lddy, dtp, lddtp, z, code, scale, mean, rstd, sums,
                                                        count > 0 ? count : 1.0, rows, P, C, pool, Lp, (uint4*)dy_panel,
                                                        panel_rows, fmt, gscale, dz_out, bias_sums ? (double*)ws : nullptr);
    }
    DCUE_LAUNCH_CHECK();
    if (bias_sums) {
        reduce_partials_kernel<<<ceil_div_i(C, 8), 256, 0, st>>>((const double*)ws, grid, C, bias_sums, bias_out, nullptr, C);
        DCUE_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int dcue_ncl_bn_bwd_reduce(const float* dx, const float* pos, int S_pos, const float* neg, int S_neg, int C,
                                      int L, const float* mean, const float* rstd, double* sums, void* ws,
                                      size_t ws_bytes, void* stream) {
    DCUE_CHECK_ARG(dx && mean && rstd && sums && ws && S_pos >= 0 && S_neg >= 0 && (pos || S_pos == 0) &&
                   (neg || S_neg == 0) && C > 0 && C % 8 == 0 && L > 0);
    cudaStream_t st = (cudaStream_t)stream;
    const int S = S_pos + S_neg;
    int G = dcue_num_sms() * 8 / (C / 8);
    if (G > S) G = S > 0 ? S : 1;
    if (G < 1) G = 1;
    if (ws_bytes < (size_t)G * 2 * C * sizeof(double)) DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_ncl_bn_bwd_reduce: workspace too small");
    ncl_bn_bwd_reduce_kernel<<<G * (C / 8), 128, 0, st>>>(dx, pos, S_pos, neg, S_neg, C, L, mean, rstd, (double*)ws);
    DCUE_LAUNCH_CHECK();
    reduce_partials_kernel<<<ceil_div_i(2 * C, 8), 256, 0, st>>>((const double*)ws, G, 2 * C, sums);
    DCUE_LAUNCH_CHECK();
    return 0;
}

extern "C" int dcue_grad_scale(const float* absmax, const float* scale, int C, double count, float* out, void* stream) {
    DCUE_CHECK_ARG(absmax && out && C > 0);
    grad_scale_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(absmax, scale, C, count, out);
    DCUE_LAUNCH_CHECK();
    return 0;
}

extern "C" int dcue_cvt_f64_f32(const double* in, int n, double mul, float* out, void* stream) {
    DCUE_CHECK_ARG(in && out && n >= 0);
    if (n == 0) return 0;
    cvt_f64_f32_kernel<<<ceil_div_i(n, 256), 256, 0, (cudaStream_t)stream>>>(in, n, mul, out);
    DCUE_LAUNCH_CHECK();
    return 0;
}
