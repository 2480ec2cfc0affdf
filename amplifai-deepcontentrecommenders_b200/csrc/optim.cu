// Multi-tensor Adam: every parameter tensor of the model in ONE launch.
// Replaces torch.optim.Adam's per-step foreach passes (dcrecommend/nn/dcue.py:143-147, :209): the dense [U,300] user
// table makes the optimizer an HBM pass over 4 arrays of the table's size; one fused read-modify-write of (p, g, m, v)
// moves 28 B per element instead of ~10 separate elementwise passes.  Semantics = torch.optim.Adam (amsgrad=False,
// maximize=False, L2 weight decay folded into the gradient, bias-corrected step size, eps added after the sqrt).
#include "common.cuh"

namespace {

struct AdamTensor {   // one row of the device-side table (int64 x 5 on the Python side)
    float* p;
    const float* g;
    float* m;
    float* v;
    long n;
};

constexpr int ADAM_THREADS = 256;
constexpr int ADAM_PER_BLOCK = ADAM_THREADS * 4 * 4;   // elements per block: 4 x float4 per thread

__global__ void __launch_bounds__(ADAM_THREADS)
adam_multi_kernel(const AdamTensor* __restrict__ tab, const int* __restrict__ blk_first, int n_tensors, float lr_over_bc1,
                  float beta1, float beta2, float eps, float wd, float inv_sqrt_bc2, const int* __restrict__ skip_a,
                  const int* __restrict__ skip_b) {
    // a raised device error flag (out-of-range user / song index in this step's forward: the loss and every gradient are
    // NaN-poisoned) turns the whole update into a no-op, so the parameters stay intact without a host sync per step
    if ((skip_a && *skip_a) || (skip_b && *skip_b)) return;
    // which tensor does this block belong to (<= a few dozen tensors: linear scan of the block prefix)
    int t = 0;
    while (t + 1 < n_tensors && (int)blockIdx.x >= blk_first[t + 1]) ++t;
    const AdamTensor a = tab[t];
    const long base = (long)((int)blockIdx.x - blk_first[t]) * ADAM_PER_BLOCK;
    const bool vec = ((reinterpret_cast<uintptr_t>(a.p) | reinterpret_cast<uintptr_t>(a.g) | reinterpret_cast<uintptr_t>(a.m) |
                       reinterpret_cast<uintptr_t>(a.v)) & 15) == 0;
    const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
    auto upd = [&](float& p, float g, float& m, float& v) {
        g = fmaf(wd, p, g);
        m = fmaf(g - m, omb1, m);                 // exp_avg.lerp_(grad, 1 - beta1)
        v = fmaf(omb2 * g, g, v * beta2);         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const float denom = fmaf(sqrtf(v), inv_sqrt_bc2, eps);
        p -= lr_over_bc1 * (m / denom);
    };
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const long i = base + ((long)it * ADAM_THREADS + threadIdx.x) * 4;
        if (i >= a.n) break;
        if (vec && i + 3 < a.n) {
            float4 p = *reinterpret_cast<float4*>(a.p + i), m = *reinterpret_cast<float4*>(a.m + i),
                   v = *reinterpret_cast<float4*>(a.v + i);
            const float4 g = __ldg(reinterpret_cast<const float4*>(a.g + i));
            upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
            *reinterpret_cast<float4*>(a.p + i) = p;
            *reinterpret_cast<float4*>(a.m + i) = m;
            *reinterpret_cast<float4*>(a.v + i) = v;
        } else {
            for (long j = i; j < a.n && j < i + 4; ++j) {
                float p = a.p[j], m = a.m[j], v = a.v[j];
                upd(p, a.g[j], m, v);
                a.p[j] = p; a.m[j] = m; a.v[j] = v;
            }
        }
    }
}


// Ranger = RAdam + Lookahead (dcrecommend/optim/ranger.py:82-165) for every parameter tensor in one launch.
// Per element, in the reference's order: v = beta2*v + (1-beta2) g^2; m = beta1*m + (1-beta1) g;
// p -= wd*lr*p; p -= step_size*lr * (adaptive ? m / (sqrt(v) + eps) : m); every k-th step slow += alpha*(p - slow), p = slow.
struct RangerTensor {
    float* p;
    const float* g;
    float* m;
    float* v;
    float* slow;
    long n;
};

__global__ void __launch_bounds__(ADAM_THREADS)
ranger_multi_kernel(const RangerTensor* __restrict__ tab, const int* __restrict__ blk_first, int n_tensors, float step_lr,
                    float beta1, float beta2, float eps, float wd_lr, int adaptive, int lookahead, float alpha,
                    const int* __restrict__ skip_a, const int* __restrict__ skip_b) {
    if ((skip_a && *skip_a) || (skip_b && *skip_b)) return;
    int t = 0;
    while (t + 1 < n_tensors && (int)blockIdx.x >= blk_first[t + 1]) ++t;
    const RangerTensor a = tab[t];
    const long base = (long)((int)blockIdx.x - blk_first[t]) * ADAM_PER_BLOCK;
    const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const long i0 = base + ((long)it * ADAM_THREADS + threadIdx.x) * 4;
        for (long j = i0; j < a.n && j < i0 + 4; ++j) {   // scalar accesses: consecutive threads touch consecutive 16-byte groups
            const float g = a.g[j];
            float p = a.p[j], m = a.m[j], v = a.v[j];
            v = __fadd_rn(__fmul_rn(v, beta2), __fmul_rn(__fmul_rn(omb2, g), g));     // mul_(beta2).addcmul_(1-beta2, g, g)
            m = __fadd_rn(__fmul_rn(m, beta1), __fmul_rn(omb1, g));                    // mul_(beta1).add_(1-beta1, g)
            if (wd_lr != 0.f) p = __fadd_rn(p, __fmul_rn(-wd_lr, p));
            if (adaptive) {
                const float denom = __fadd_rn(sqrtf(v), eps);
                p = __fadd_rn(p, __fmul_rn(-step_lr, __fdiv_rn(m, denom)));
            } else {
                p = __fadd_rn(p, __fmul_rn(-step_lr, m));
            }
            if (lookahead) {
                float s = a.slow[j];
                s = __fadd_rn(s, __fmul_rn(alpha, __fsub_rn(p, s)));
                a.slow[j] = s;
                p = s;
            }
            a.p[j] = p; a.m[j] = m; a.v[j] = v;
        }
    }
}

}  // namespace

extern "C" int dcue_ranger_multi_step(const void* table_dev, const int* blk_first_dev, int n_tensors, int total_blocks,
                                      float step_size_times_lr, float beta1, float beta2, float eps, float weight_decay_times_lr,
                                      int adaptive, int lookahead, float alpha, const int* skip_flag_a, const int* skip_flag_b,
                                      void* stream) {
    DCUE_CHECK_ARG(table_dev && blk_first_dev && n_tensors > 0 && total_blocks >= 0);
    if (total_blocks == 0) return 0;
    ranger_multi_kernel<<<total_blocks, ADAM_THREADS, 0, (cudaStream_t)stream>>>((const RangerTensor*)table_dev, blk_first_dev, n_tensors,
                                                                                 step_size_times_lr, beta1, beta2, eps,
                                                                                 weight_decay_times_lr, adaptive, lookahead, alpha,
                                                                                 skip_flag_a, skip_flag_b);
    DCUE_LAUNCH_CHECK();
    return 0;
}

extern "C" int dcue_adam_elems_per_block(void) { return ADAM_PER_BLOCK; }

extern "C" int dcue_adam_multi_step(const void* table_dev, const int* blk_first_dev, int n_tensors, int total_blocks, float lr,
                                    float beta1, float beta2, float eps, float weight_decay, float bias_correction1,
                                    float bias_correction2, const int* skip_flag_a, const int* skip_flag_b, void* stream) {
    DCUE_CHECK_ARG(table_dev && blk_first_dev && n_tensors > 0 && total_blocks >= 0 && bias_correction1 > 0.f && bias_correction2 > 0.f);
    if (total_blocks == 0) return 0;
    adam_multi_kernel<<<total_blocks, ADAM_THREADS, 0, (cudaStream_t)stream>>>((const AdamTensor*)table_dev, blk_first_dev, n_tensors,
                                                                               lr / bias_correction1, beta1, beta2, eps, weight_decay,
                                                                               1.f / sqrtf(bias_correction2), skip_flag_a, skip_flag_b);
    DCUE_LAUNCH_CHECK();
    return 0;
}
