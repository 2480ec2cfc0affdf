// Cosine scoring of each user against 1 positive + N negatives, fused with the max-margin hinge
// loss and its analytic backward.  One warp per triplet, one pass over the negatives, all
// reductions by warp shuffle.  HBM-bound: reads 4F(2+N) B, writes 4F(2+N)+4N B per triplet.
//
// Semantics follow the reference call sites:
//   nn.CosineSimilarity(dim=1)         dcrecommend/dcue/dcue.py:68,93-100   (per-norm eps clamp)
//   scores = pos - neg                 dcrecommend/dcue/dcue.py:106
//   max(0, margin - s).sum(1).mean()   dcrecommend/nn/dcue.py:167-170       (tie -> half gradient)
#include "common.cuh"

namespace {

constexpr int WARPS = 4;

template <int VPT>
struct Row {
    float v[VPT];
};

template <int VPT>
__device__ __forceinline__ void load_row(Row<VPT>& r, const float* __restrict__ p, int F, int lane) {
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        int e = lane + 32 * i;
        r.v[i] = e < F ? __ldg(p + e) : 0.f;
    }
}
template <int VPT>
__device__ __forceinline__ void store_row(const Row<VPT>& r, float* __restrict__ p, int F, int lane) {
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        int e = lane + 32 * i;
        if (e < F) p[e] = r.v[i];
    }
}
template <int VPT>
__device__ __forceinline__ float dot(const Row<VPT>& a, const Row<VPT>& b) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) s = fmaf(a.v[i], b.v[i], s);
    return warp_sum(s);
}

// d cos(x,y) / dx  given xh = x/max(|x|,eps), yh, c = xh.yh:
//   |x| >= eps : (yh - c*xh)/|x|      |x| < eps : yh/eps   (clamp blocks the norm's gradient)
__device__ __forceinline__ float dcos(float yh, float xh, float c, float nx, float eps) {
    return nx >= eps ? (yh - c * xh) / nx : yh / eps;
}

// MODE 0: scores only.  MODE 1: backward from gscores.  MODE 2: fused hinge forward+backward.
template <int VPT, int MODE>
__global__ void __launch_bounds__(WARPS * 32)
score_kernel(const float* __restrict__ u, const float* __restrict__ feats, const float* __restrict__ gscores,
             int B, int N, int F, float eps, float margin, float inv_batch, float* __restrict__ scores,
             float* __restrict__ loss_rows, float* __restrict__ du, float* __restrict__ dfeats) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (b >= B) return;

    Row<VPT> ur, pr, uh, ph;
    load_row(ur, u + (long)b * F, F, lane);
    load_row(pr, feats + (long)b * F, F, lane);
    const float nu = sqrtf(dot(ur, ur)), np = sqrtf(dot(pr, pr));
    const float iu = 1.f / fmaxf(nu, eps), ip = 1.f / fmaxf(np, eps);
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        uh.v[i] = ur.v[i] * iu;
        ph.v[i] = pr.v[i] * ip;
    }
    const float cpos = dot(uh, ph);

    Row<VPT> acc_du;  // sum over rows of dL/dc_row * dc_row/du
#pragma unroll
    for (int i = 0; i < VPT; ++i) acc_du.v[i] = 0.f;
    float G = 0.f, loss = 0.f;

    const float* negp = feats + ((long)B + (long)b * N) * F;
    float* dnegp = MODE ? dfeats + ((long)B + (long)b * N) * F : nullptr;
    Row<VPT> nr, nxt;
    if (N > 0) load_row(nxt, negp, F, lane);
    for (int n = 0; n < N; ++n) {
        nr = nxt;
        if (n + 1 < N) load_row(nxt, negp + (long)(n + 1) * F, F, lane);  // prefetch next row
        const float nn = sqrtf(dot(nr, nr));
        const float in_ = 1.f / fmaxf(nn, eps);
        Row<VPT> nh;
#pragma unroll
        for (int i = 0; i < VPT; ++i) nh.v[i] = nr.v[i] * in_;
        const float cn = dot(uh, nh);
        const float s = cpos - cn;
        if (MODE != 1 && lane == 0) scores[(long)b * N + n] = s;
        if (MODE == 0) continue;
        float g;  // dL/ds_n
        if (MODE == 2) {
            const float h = margin - s;
            // torch.max(0, NaN) is NaN (fmaxf would drop it): a NaN score -- e.g. from a flagged out-of-range user row --
            // poisons the loss and the gradients instead of looking like a satisfied margin
            loss += (h != h) ? h : fmaxf(h, 0.f);
            g = (h != h) ? h : -(h > 0.f ? 1.f : (h == 0.f ? 0.5f : 0.f)) * inv_batch;
        } else {
            g = gscores[(long)b * N + n];
        }
        G += g;
        // dL/dc_n = -g
        Row<VPT> dn;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            dn.v[i] = -g * dcos(uh.v[i], nh.v[i], cn, nn, eps);
            acc_du.v[i] += -g * dcos(nh.v[i], uh.v[i], cn, nu, eps);
        }
        store_row(dn, dnegp + (long)n * F, F, lane);
    }
    if (MODE == 0) return;
    Row<VPT> dp;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        dp.v[i] = G * dcos(uh.v[i], ph.v[i], cpos, np, eps);
        acc_du.v[i] += G * dcos(ph.v[i], uh.v[i], cpos, nu, eps);
    }
    store_row(dp, dfeats + (long)b * F, F, lane);
    store_row(acc_du, du + (long)b * F, F, lane);
    if (MODE == 2 && lane == 0) loss_rows[b] = loss;
}

// ---- F % 4 == 0, F <= 128: 8 lanes per feature row, 16-byte accesses, FOUR negatives of the triplet in flight per
// warp (the one-row-at-a-time kernel above is a chain of load -> two 5-step warp reductions -> store per negative:
// 2.6 TB/s at N = 20; rows of 100 floats also leave a quarter of its lanes idle).
__device__ __forceinline__ float group8_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}
struct Row4 {
    float4 v[4];
};
__device__ __forceinline__ void load_row4(Row4& r, const float* __restrict__ p, int n4, int sub, bool on) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int e = sub + 8 * i;
        r.v[i] = (on && e < n4) ? __ldg(reinterpret_cast<const float4*>(p) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
__device__ __forceinline__ void store_row4(const Row4& r, float* __restrict__ p, int n4, int sub, bool on) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int e = sub + 8 * i;
        if (on && e < n4) reinterpret_cast<float4*>(p)[e] = r.v[i];
    }
}
__device__ __forceinline__ float dot4(const Row4& a, const Row4& b) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s = fmaf(a.v[i].x, b.v[i].x, s); s = fmaf(a.v[i].y, b.v[i].y, s);
        s = fmaf(a.v[i].z, b.v[i].z, s); s = fmaf(a.v[i].w, b.v[i].w, s);
    }
    return group8_sum(s);
}
__device__ __forceinline__ void scale4(Row4& o, const Row4& a, float k) {
#pragma unroll
    for (int i = 0; i < 4; ++i) o.v[i] = make_float4(a.v[i].x * k, a.v[i].y * k, a.v[i].z * k, a.v[i].w * k);
}

// d cos / dx with the reciprocal norms hoisted out of the element loop (the division per element made the kernel
// instruction-bound): k = 1/|x| if |x| >= eps (and c kept), else k = 1/eps with c = 0
__device__ __forceinline__ float dcos_k(float yh, float xh, float c_eff, float k) { return (yh - c_eff * xh) * k; }

template <int MODE>
__global__ void __launch_bounds__(WARPS * 32, 4)
score_kernel_g8(const float* __restrict__ u, const float* __restrict__ feats, const float* __restrict__ gscores,
                int B, int N, int F, float eps, float margin, float inv_batch, float* __restrict__ scores,
                float* __restrict__ loss_rows, float* __restrict__ du, float* __restrict__ dfeats) {
    const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
    const int b = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (b >= B) return;
    const int n4 = F >> 2;
    const float inv_eps = 1.f / eps;
    Row4 uh;
    float nu, np, cpos;
    {
        Row4 ur, pr;
        load_row4(ur, u + (long)b * F, n4, sub, true);      // every 8-lane group holds the user row
        load_row4(pr, feats + (long)b * F, n4, sub, true);
        nu = sqrtf(dot4(ur, ur));
        np = sqrtf(dot4(pr, pr));
        scale4(uh, ur, 1.f / fmaxf(nu, eps));
        scale4(pr, pr, 1.f / fmaxf(np, eps));
        cpos = dot4(uh, pr);
    }
    const float ku = nu >= eps ? 1.f / nu : inv_eps;        // d/du factors
    const bool u_ok = nu >= eps;

    Row4 acc;   // this group's share of sum_n dL/dc_n * dc_n/du
#pragma unroll
    for (int i = 0; i < 4; ++i) acc.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    float G = 0.f, loss = 0.f;
    const float* negp = feats + ((long)B + (long)b * N) * F;
    float* dnegp = MODE ? dfeats + ((long)B + (long)b * N) * F : nullptr;
    Row4 nxt;
    load_row4(nxt, negp + (long)grp * F, n4, sub, grp < N);
    for (int n0 = 0; n0 < N; n0 += 4) {
        const int n = n0 + grp;
        const bool on = n < N;
        Row4 nh = nxt;
        load_row4(nxt, negp + (long)(n + 4) * F, n4, sub, n + 4 < N);   // the next four rows are in flight
        const float nn = sqrtf(dot4(nh, nh));
        scale4(nh, nh, 1.f / fmaxf(nn, eps));
        const float cn = dot4(uh, nh);
        const float s = cpos - cn;
        if (MODE != 1 && on && sub == 0) scores[(long)b * N + n] = s;
        if (MODE == 0) continue;
        float g = 0.f;  // dL/ds_n
        if (MODE == 2) {
            const float h = margin - s;
            if (on) {
                loss += (h != h) ? h : fmaxf(h, 0.f);     // NaN propagates like torch.max (see score_kernel)
                g = (h != h) ? h : -(h > 0.f ? 1.f : (h == 0.f ? 0.5f : 0.f)) * inv_batch;
            }
        } else if (on) {
            g = gscores[(long)b * N + n];
        }
        G += g;
        // dL/dc_n = -g
        const float kn = -g * (nn >= eps ? 1.f / nn : inv_eps), cn_n = nn >= eps ? cn : 0.f;
        const float kun = -g * ku, cn_u = u_ok ? cn : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 uu = uh.v[i], hh = nh.v[i];
            acc.v[i].x += dcos_k(hh.x, uu.x, cn_u, kun); acc.v[i].y += dcos_k(hh.y, uu.y, cn_u, kun);
            acc.v[i].z += dcos_k(hh.z, uu.z, cn_u, kun); acc.v[i].w += dcos_k(hh.w, uu.w, cn_u, kun);
            nh.v[i] = make_float4(dcos_k(uu.x, hh.x, cn_n, kn), dcos_k(uu.y, hh.y, cn_n, kn), dcos_k(uu.z, hh.z, cn_n, kn),
                                  dcos_k(uu.w, hh.w, cn_n, kn));
        }
        store_row4(nh, dnegp + (long)n * F, n4, sub, on);
    }
    if (MODE == 0) return;
    // combine the four groups (fixed order: deterministic)
    G += __shfl_xor_sync(0xffffffffu, G, 8);
    G += __shfl_xor_sync(0xffffffffu, G, 16);
    loss += __shfl_xor_sync(0xffffffffu, loss, 8);
    loss += __shfl_xor_sync(0xffffffffu, loss, 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float a[4] = {acc.v[i].x, acc.v[i].y, acc.v[i].z, acc.v[i].w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            a[t] += __shfl_xor_sync(0xffffffffu, a[t], 8);
            a[t] += __shfl_xor_sync(0xffffffffu, a[t], 16);
        }
        acc.v[i] = make_float4(a[0], a[1], a[2], a[3]);
    }
    // positive row: dL/dc_pos = +G (re-read: keeping its normalised copy live across the loop costs 16 registers)
    Row4 ph;
    load_row4(ph, feats + (long)b * F, n4, sub, true);
    scale4(ph, ph, 1.f / fmaxf(np, eps));
    const float kp = G * (np >= eps ? 1.f / np : inv_eps), cp_p = np >= eps ? cpos : 0.f;
    const float kup = G * ku, cp_u = u_ok ? cpos : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 uu = uh.v[i], hh = ph.v[i];
        acc.v[i].x += dcos_k(hh.x, uu.x, cp_u, kup); acc.v[i].y += dcos_k(hh.y, uu.y, cp_u, kup);
        acc.v[i].z += dcos_k(hh.z, uu.z, cp_u, kup); acc.v[i].w += dcos_k(hh.w, uu.w, cp_u, kup);
        ph.v[i] = make_float4(dcos_k(uu.x, hh.x, cp_p, kp), dcos_k(uu.y, hh.y, cp_p, kp), dcos_k(uu.z, hh.z, cp_p, kp),
                              dcos_k(uu.w, hh.w, cp_p, kp));
    }
    store_row4(ph, dfeats + (long)b * F, n4, sub, grp == 0);
    store_row4(acc, du + (long)b * F, n4, sub, grp == 0);
    if (MODE == 2 && lane == 0) loss_rows[b] = loss;
}

template <int MODE>
int launch(const float* u, const float* feats, const float* gs, int B, int N, int F, float eps, float margin,
           float inv_batch, float* scores, float* loss_rows, float* du, float* dfeats, cudaStream_t st) {
    if (B == 0) return 0;
    dim3 grid(ceil_div_i(B, WARPS)), block(WARPS * 32);
    const bool al16 = ((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(feats) | reinterpret_cast<uintptr_t>(du) |
                        reinterpret_cast<uintptr_t>(dfeats)) & 15) == 0;
    if (F <= 128 && (F & 3) == 0 && al16 && N > 0)
        score_kernel_g8<MODE><<<grid, block, 0, st>>>(u, feats, gs, B, N, F, eps, margin, inv_batch, scores, loss_rows, du,
                                                      dfeats);
    else if (F <= 128)
        score_kernel<4, MODE><<<grid, block, 0, st>>>(u, feats, gs, B, N, F, eps, margin, inv_batch, scores,
                                                      loss_rows, du, dfeats);
    else
        score_kernel<8, MODE><<<grid, block, 0, st>>>(u, feats, gs, B, N, F, eps, margin, inv_batch, scores,
                                                      loss_rows, du, dfeats);
    DCUE_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" int dcue_score_fwd(const float* u, const float* feats, int B, int N, int F, float eps, float* scores,
                              void* stream) {
    DCUE_CHECK_ARG(u && feats && scores && B >= 0 && N >= 0 && F > 0 && F <= 256);
    return launch<0>(u, feats, nullptr, B, N, F, eps, 0.f, 0.f, scores, nullptr, nullptr, nullptr,
                     (cudaStream_t)stream);
}

extern "C" int dcue_score_bwd(const float* u, const float* feats, const float* gscores, int B, int N, int F,
                              float eps, float* du, float* dfeats, void* stream) {
    DCUE_CHECK_ARG(u && feats && gscores && du && dfeats && B >= 0 && N >= 0 && F > 0 && F <= 256);
    return launch<1>(u, feats, gscores, B, N, F, eps, 0.f, 0.f, nullptr, nullptr, du, dfeats,
                     (cudaStream_t)stream);
}

extern "C" int dcue_score_hinge_fwdbwd(const float* u, const float* feats, int B, int N, int F, float eps,
                                       float margin, int batch_total, float* scores, float* loss_rows,
                                       float* du, float* dfeats, void* stream) {
    DCUE_CHECK_ARG(u && feats && scores && loss_rows && du && dfeats && B >= 0 && N >= 0 && F > 0 &&
                   F <= 256 && batch_total > 0);
    return launch<2>(u, feats, nullptr, B, N, F, eps, margin, 1.f / (float)batch_total, scores, loss_rows, du,
                     dfeats, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ loss scalar + gradient rescale (no ATen glue in the step)
namespace {
// loss = sum_b loss_rows[b] / batch_total, one block, fixed order (deterministic)
__global__ void __launch_bounds__(256) loss_mean_kernel(const float* __restrict__ rows, int B, float inv_total, float* __restrict__ out) {
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < B; i += 256) s += (double)rows[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = (float)(red[0] * (double)inv_total);
}
// a_out = a * g, b_out = b * g with g a device scalar (the incoming gradient of the loss): both in one launch
__global__ void __launch_bounds__(256) scale_pair_kernel(const float* __restrict__ a, long na, const float* __restrict__ b, long nb,
                                                         const float* __restrict__ g, float* __restrict__ ao, float* __restrict__ bo) {
    const float gs = __ldg(g);
    const long n = na + nb;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        if (i < na) ao[i] = a[i] * gs;
        else bo[i - na] = b[i - na] * gs;
    }
}
}  // namespace

extern "C" int dcue_loss_mean(const float* loss_rows, int B, int batch_total, float* loss_out, void* stream) {
    DCUE_CHECK_ARG(loss_rows && loss_out && B >= 0 && batch_total > 0);
    loss_mean_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(loss_rows, B, 1.f / (float)batch_total, loss_out);
    DCUE_LAUNCH_CHECK();
    return 0;
}

extern "C" int dcue_scale_pair(const float* a, long na, const float* b, long nb, const float* g_dev, float* a_out, float* b_out,
                               void* stream) {
    DCUE_CHECK_ARG(g_dev && na >= 0 && nb >= 0 && (na == 0 || (a && a_out)) && (nb == 0 || (b && b_out)));
    const long n = na + nb;
    if (n == 0) return 0;
    long blocks = (n + 255) / 256;
    if (blocks > 148L * 8) blocks = 148L * 8;
    scale_pair_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(a, na, b, nb, g_dev, a_out, b_out);
    DCUE_LAUNCH_CHECK();
    return 0;
}
