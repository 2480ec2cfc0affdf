// User embedding table: vectorised coalesced gather (+ReLU) and a deterministic
// sort-by-row + segment-sum backward that produces the reference's DENSE [U,E] gradient
// (nn.Embedding(sparse=False), dcrecommend/dcue/embeddings/userembedding.py:27,40-41).
#include "common.cuh"
#include <cub/cub.cuh>

namespace {

// one warp per gathered row; 16 B per lane when E % 4 == 0
__global__ void __launch_bounds__(256)
gather_relu_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx, int B, int U, int E,
                   float* __restrict__ out, float* __restrict__ raw, int* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    const int64_t r = idx[b];
    if (r < 0 || r >= U) {  // nn.Embedding raises: flag it and poison the row so the loss cannot look sane
        if (lane == 0) atomicExch(err, 1);
        for (int i = lane; i < E; i += 32) out[(long)b * E + i] = __int_as_float(0x7fc00000);
        return;
    }
    const float* src = table + r * (long)E;
    float* dst = out + (long)b * E;
    float* rdst = raw ? raw + (long)b * E : nullptr;
    if ((E & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        float4* r4 = reinterpret_cast<float4*>(rdst);
        const int n4 = E / 4;
        for (int i0 = 0; i0 < n4; i0 += 128) {   // 4 independent 16-byte loads per lane in flight before any store
            float4 v[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int i = i0 + t * 32 + lane;
                if (i < n4) v[t] = __ldg(s4 + i);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int i = i0 + t * 32 + lane;
                if (i < n4) {
                    float4 w = v[t];
                    if (r4) r4[i] = w;
                    w.x = fmaxf(w.x, 0.f); w.y = fmaxf(w.y, 0.f); w.z = fmaxf(w.z, 0.f); w.w = fmaxf(w.w, 0.f);
                    d4[i] = w;
                }
            }
        }
    } else {
        for (int i = lane; i < E; i += 32) {
            float v = __ldg(src + i);
            if (rdst) rdst[i] = v;
            dst[i] = fmaxf(v, 0.f);
        }
    }
}

__global__ void iota_kernel(int32_t* p, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// one warp per sorted entry; only segment heads work: they sum their run in position order
__global__ void __launch_bounds__(256)
segment_scatter_kernel(const float* __restrict__ gout, const float* __restrict__ fwd,
                       const int64_t* __restrict__ sidx, const int32_t* __restrict__ spos, int B, int U, int E,
                       float* __restrict__ gtable) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= B) return;
    const int64_t row = sidx[i];
    if (row < 0 || row >= U) return;   // out-of-range index (flagged by the forward gather): never write outside the table
    if (i > 0 && sidx[i - 1] == row) return;
    int end = i + 1;
    while (end < B && sidx[end] == row) ++end;
    float* dst = gtable + row * (long)E;
    if ((E & 3) == 0 && ((reinterpret_cast<uintptr_t>(gout) | reinterpret_cast<uintptr_t>(fwd) | reinterpret_cast<uintptr_t>(gtable)) & 15) == 0) {
        // 16-byte loads/stores: every row starts 16-byte aligned
        for (int e0 = lane * 4; e0 < E; e0 += 128) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int j = i; j < end; ++j) {
                const long p = (long)spos[j] * E + e0;
                const float4 g = __ldg(reinterpret_cast<const float4*>(gout + p));
                const float4 f = __ldg(reinterpret_cast<const float4*>(fwd + p));
                a.x += f.x > 0.f ? g.x : 0.f; a.y += f.y > 0.f ? g.y : 0.f;
                a.z += f.z > 0.f ? g.z : 0.f; a.w += f.w > 0.f ? g.w : 0.f;
            }
            *reinterpret_cast<float4*>(dst + e0) = a;
        }
        return;
    }
    for (int e0 = lane * 4; e0 < E; e0 += 128) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = i; j < end; ++j) {
            const long p = (long)spos[j] * E;
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (e0 + t < E) {
                    float g = __ldg(gout + p + e0 + t);
                    float f = __ldg(fwd + p + e0 + t);
                    a[t] += f > 0.f ? g : 0.f;
                }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (e0 + t < E) dst[e0 + t] = a[t];
    }
}


// Sort-free deterministic segment sum for step-sized batches (B <= DCUE_SCATTER_DIRECT_MAX): one warp per batch row b
// scans the whole index vector (L1/L2 resident, 8 B per entry).  A row with an EARLIER duplicate does nothing; the first
// occurrence of a table row ("head") adds every later duplicate in position order, so the result equals the sorted
// segment sum bit for bit -- without the sort, its workspace, or the two extra launches.
constexpr int SCAT_CHUNK = 512;   // floats of a row handled per scan (4 x float4 per lane)
__global__ void __launch_bounds__(256)
dup_scan_scatter_kernel(const float* __restrict__ gout, const float* __restrict__ fwd, const int64_t* __restrict__ idx,
                        int B, int U, int E, float* __restrict__ gtable) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    const int64_t row = idx[b];
    if (row < 0 || row >= U) return;
    const bool vec = (E & 3) == 0 && ((reinterpret_cast<uintptr_t>(gout) | reinterpret_cast<uintptr_t>(fwd) |
                                      reinterpret_cast<uintptr_t>(gtable)) & 15) == 0;
    float* dst = gtable + row * (long)E;
    for (int c0 = 0; c0 < E; c0 += SCAT_CHUNK) {
        float4 a[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) a[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j0 = 0; j0 < B; j0 += 32) {
            const int j = j0 + lane;
            unsigned m = __ballot_sync(0xffffffffu, j < B && __ldg(idx + j) == row);
            if (j0 + 32 <= b) {          // chunk entirely before b
                if (m) return;           // an earlier duplicate is the head
                continue;
            }
            if (j0 <= b) {               // chunk containing b
                if (m & ((1u << (b - j0)) - 1u)) return;
            }
            while (m) {                  // b itself and later duplicates, ascending position
                const int pj = j0 + __ffs(m) - 1;
                m &= m - 1;
                const long p = (long)pj * E;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int e0 = c0 + t * 128 + lane * 4;
                    if (vec) {
                        if (e0 < E) {
                            const float4 g = __ldg(reinterpret_cast<const float4*>(gout + p + e0));
                            float4 f = make_float4(1.f, 1.f, 1.f, 1.f);
                            if (fwd) f = __ldg(reinterpret_cast<const float4*>(fwd + p + e0));
                            a[t].x += f.x > 0.f ? g.x : 0.f; a[t].y += f.y > 0.f ? g.y : 0.f;
                            a[t].z += f.z > 0.f ? g.z : 0.f; a[t].w += f.w > 0.f ? g.w : 0.f;
                        }
                    } else {
                        float* av = &a[t].x;
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (e0 + q < E) {
                                const float g = __ldg(gout + p + e0 + q);
                                const float f = fwd ? __ldg(fwd + p + e0 + q) : 1.f;
                                av[q] += f > 0.f ? g : 0.f;
                            }
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int e0 = c0 + t * 128 + lane * 4;
            if (vec) {
                if (e0 < E) *reinterpret_cast<float4*>(dst + e0) = a[t];
            } else {
                const float* av = &a[t].x;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (e0 + q < E) dst[e0 + q] = av[q];
            }
        }
    }
}

struct SortWs {
    size_t keys_in_off, vals_in_off, temp_off, temp_bytes, total;
};
SortWs sort_layout(int B) {
    SortWs w;
    size_t temp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp, (const int64_t*)nullptr, (int64_t*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, B);
    w.vals_in_off = 0;
    w.temp_off = round_up_l((long)B * 4, 256);
    w.temp_bytes = temp;
    w.total = w.temp_off + round_up_l((long)temp, 256);
    return w;
}

}  // namespace

extern "C" int dcue_gather_relu_fwd(const float* table, const int64_t* idx, int B, int U, int E, float* out,
                                    float* raw_out, int* err_flag, void* stream) {
    DCUE_CHECK_ARG(table && idx && out && err_flag && B >= 0 && U > 0 && E > 0);
    if (B == 0) return 0;
    gather_relu_kernel<<<ceil_div_i(B, 8), 256, 0, (cudaStream_t)stream>>>(table, idx, B, U, E, out, raw_out,
                                                                            err_flag);
    DCUE_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t dcue_sort_ws_bytes(int B) { return B > 0 ? sort_layout(B).total : 256; }

extern "C" int dcue_sort_indices(const int64_t* idx, int B, int U, int64_t* sorted_idx, int32_t* sorted_pos,
                                 void* ws, size_t ws_bytes, void* stream) {
    DCUE_CHECK_ARG(idx && sorted_idx && sorted_pos && ws && B >= 0 && U > 0);
    if (B == 0) return 0;
    SortWs w = sort_layout(B);
    if (ws_bytes < w.total) DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_sort_indices: workspace %zu < %zu", ws_bytes, w.total);
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* iota = reinterpret_cast<int32_t*>((char*)ws + w.vals_in_off);
    iota_kernel<<<ceil_div_i(B, 256), 256, 0, st>>>(iota, B);
    DCUE_LAUNCH_CHECK();
    int end_bit = 1;
    while (end_bit < 63 && (1LL << end_bit) < (long long)U) ++end_bit;
    size_t temp = w.temp_bytes;
    DCUE_CUDA(cub::DeviceRadixSort::SortPairs((char*)ws + w.temp_off, temp, idx, sorted_idx, iota, sorted_pos, B,
                                              0, end_bit, st));
    return 0;
}

extern "C" int dcue_scatter_add_bwd(const float* grad_out, const float* fwd_out, const int64_t* sorted_idx,
                                    const int32_t* sorted_pos, int B, int U, int E, float* grad_table,
                                    void* stream) {
    DCUE_CHECK_ARG(grad_out && fwd_out && sorted_idx && sorted_pos && grad_table && B >= 0 && U > 0 && E > 0);
    if (B == 0) return 0;
    segment_scatter_kernel<<<ceil_div_i(B, 8), 256, 0, (cudaStream_t)stream>>>(grad_out, fwd_out, sorted_idx,
                                                                                sorted_pos, B, U, E, grad_table);
    DCUE_LAUNCH_CHECK();
    return 0;
}

extern "C" int dcue_scatter_direct_max(void) { return 16384; }

extern "C" int dcue_scatter_add_rows(const float* grad_rows, const float* fwd_mask, const int64_t* idx, int B, int U, int E,
                                     float* grad_table, void* stream) {
    DCUE_CHECK_ARG(grad_rows && idx && grad_table && B >= 0 && U > 0 && E > 0);
    if (B > dcue_scatter_direct_max())
        DCUE_FAIL(DCUE_E_UNSUPPORTED, "dcue_scatter_add_rows: B = %d > %d, use dcue_sort_indices + dcue_scatter_add_bwd", B,
                  dcue_scatter_direct_max());
    if (B == 0) return 0;
    dup_scan_scatter_kernel<<<ceil_div_i(B, 8), 256, 0, (cudaStream_t)stream>>>(grad_rows, fwd_mask, idx, B, U, E, grad_table);
    DCUE_LAUNCH_CHECK();
    return 0;
}
