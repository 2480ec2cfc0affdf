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
            loss += fmaxf(h, 0.f);
            g = -(h > 0.f ? 1.f : (h == 0.f ? 0.5f : 0.f)) * inv_batch;
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

template <int MODE>
int launch(const float* u, const float* feats, const float* gs, int B, int N, int F, float eps, float margin,
           float inv_batch, float* scores, float* loss_rows, float* du, float* dfeats, cudaStream_t st) {
    if (B == 0) return 0;
    dim3 grid(ceil_div_i(B, WARPS)), block(WARPS * 32);
    if (F <= 128)
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
