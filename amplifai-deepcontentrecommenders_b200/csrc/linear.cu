// fp32-accurate GEMMs for the small dense layers of DCUE: the user MLP
// (dcrecommend/dcue/embeddings/userembedding.py:42-44), the k=1 conv and the fc of the song
// tower (truedcuemel1dbn.py:57-59,101).  1.7 % of the step's FLOPs, but as CUDA-core FMAs they took 10 % of the step
// (round 1: 12 launches, 0.30 ms); they now run on the tensor cores as 3xTF32 (mma.sync.m16n8k8: a = a_hi + a_lo with
// both halves TF32, a*b ~ a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, fp32 accumulate), which keeps ~2^-21 relative accuracy --
// the user features still match the reference to ~1e-6.  These GEMMs read 10-20 MB each: warp-level MMA from the same
// shared-memory tiles is enough to make them memory/latency bound; DCUE_LINEAR_IMPL=simt selects the CUDA-core loop (A/B).
//
// One strided kernel:  C[i,j] = sum_r A(i,r) * B(r,j)  (+bias[j]) (relu) (* mask[i,j] > 0)
// 64x64 tile, BK = 16, 256 threads, 4x4 outputs per thread, optional split over r (grid.z)
// into a workspace that a second kernel reduces in fixed order (deterministic).
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int TM = 64, TN = 64, TK = 32;   // k depth per barrier: two 16-deep half tiles (each thread fetches 2 x 16 B per operand)
constexpr int TKH = 16;

struct GemmP {
    const float* A; long sAi, sAr;
    const float* B; long sBr, sBj;
    float* C; long ldc;       // C[i*ldc + j]   (or partial buffer when splits > 1)
    const float* bias;        // [J] or null
    const float* mask; long ldmask;
    int I, J, R, relu, splits;
    long split_stride;        // elements between partial buffers
    float* rowsum;            // nullable [splits][I]: per-split row sums of A (cp.async kernel only)
    int single_tf32;          // 1: one TF32 MMA per product (10-bit mantissa operands, like the 16-bit conv operands); 0: 3xTF32
};

// Global -> register fetch of this thread's 4 elements of the A (TM x TK) and B (TK x TN) tiles at r0.
struct GemmFrag { float a[4], b[4]; };

__device__ __forceinline__ void gemm_fetch(const GemmP& p, int tid, int i0, int j0, int r0, int rend, bool a_r_contig,
                                           bool b_j_contig, bool a_vec, bool b_vec, GemmFrag& f) {
    {
        int li, lr;
        if (a_r_contig) { li = tid >> 2; lr = (tid & 3) * 4; }
        else            { lr = tid >> 4; li = (tid & 15) * 4; }
        const int gi = i0 + li, gr = r0 + lr;
        const float* src = p.A + gi * p.sAi + gr * p.sAr;
#pragma unroll
        for (int q = 0; q < 4; ++q) f.a[q] = 0.f;
        const bool full = a_r_contig ? (gi < p.I && gr + 3 < rend) : (gr < rend && gi + 3 < p.I);
        if (full && a_vec) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(src));
            f.a[0] = t.x; f.a[1] = t.y; f.a[2] = t.z; f.a[3] = t.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int gi2 = a_r_contig ? gi : gi + q, gr2 = a_r_contig ? gr + q : gr;
                if (gi2 < p.I && gr2 < rend) f.a[q] = __ldg(p.A + gi2 * p.sAi + gr2 * p.sAr);
            }
        }
    }
    {
        int lj, lr;
        if (b_j_contig) { lr = tid >> 4; lj = (tid & 15) * 4; }
        else            { lj = tid >> 2; lr = (tid & 3) * 4; }
        const int gj = j0 + lj, gr = r0 + lr;
        const float* src = p.B + gr * p.sBr + gj * p.sBj;
#pragma unroll
        for (int q = 0; q < 4; ++q) f.b[q] = 0.f;
        const bool full = b_j_contig ? (gr < rend && gj + 3 < p.J) : (gj < p.J && gr + 3 < rend);
        if (full && b_vec) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(src));
            f.b[0] = t.x; f.b[1] = t.y; f.b[2] = t.z; f.b[3] = t.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int gj2 = b_j_contig ? gj + q : gj, gr2 = b_j_contig ? gr : gr + q;
                if (gj2 < p.J && gr2 < rend) f.b[q] = __ldg(p.B + gr2 * p.sBr + gj2 * p.sBj);
            }
        }
    }
}

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float r = x - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <bool MMA>
__global__ void __launch_bounds__(256) gemm_kernel(GemmP p) {
    // double-buffered tiles: the next tile's global loads are in flight while the current one is multiplied
    // (the single-buffered version exposed one DRAM/L2 round trip per 16-deep k step: 35 us for 0.55 GFLOP)
    __shared__ __align__(16) float As[2][TK][TM + 8];   // row stride 72 words: the MMA fragment reads hit 32 distinct banks
    __shared__ __align__(16) float Bs[2][TK][TN + 8];
    const int tid = threadIdx.x;
    const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TN;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 4x4
    // split range over r
    const int rchunk = ceil_div_i(ceil_div_i(p.R, TK), p.splits) * TK;
    const int rbeg = blockIdx.z * rchunk;
    const int rend = min(p.R, rbeg + rchunk);

    // SIMT: thread (ty, tx) owns the 4x4 block at rows ty*4, cols tx*4.
    // MMA : warp (wm = warp & 1, wn = warp >> 1) owns rows wm*32..+32, cols wn*16..+16 as 2 x 2 m16n8 tiles; acc[mi*2+ni][0..3]
    //       are the fragment's (row g, col 2t), (g, 2t+1), (g+8, 2t), (g+8, 2t+1) with g = lane / 4, t = lane % 4.
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    const int lane = tid & 31, wid = tid >> 5;
    const int wm = wid & 1, wn = wid >> 1, fg = lane >> 2, ft = lane & 3;

    const bool a_r_contig = (p.sAr == 1);
    const bool b_j_contig = (p.sBj == 1);
    // 16-byte loads need every row start 16-byte aligned: base pointer and the non-unit stride
    const bool a_vec = ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0) && (((a_r_contig ? p.sAi : p.sAr) & 3) == 0) &&
                       (a_r_contig || p.sAi == 1) && ((rbeg & 3) == 0);
    const bool b_vec = ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0) && (((b_j_contig ? p.sBr : p.sBj) & 3) == 0) &&
                       (b_j_contig || p.sBr == 1) && ((rbeg & 3) == 0);

    auto stash = [&](int buf, int koff, const GemmFrag& f) {
        if (a_r_contig) {
            const int li = tid >> 2, lr = koff + (tid & 3) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) As[buf][lr + q][li] = f.a[q];
        } else {
            const int lr = koff + (tid >> 4), li = (tid & 15) * 4;
            *reinterpret_cast<float4*>(&As[buf][lr][li]) = make_float4(f.a[0], f.a[1], f.a[2], f.a[3]);
        }
        if (b_j_contig) {
            const int lr = koff + (tid >> 4), lj = (tid & 15) * 4;
            *reinterpret_cast<float4*>(&Bs[buf][lr][lj]) = make_float4(f.b[0], f.b[1], f.b[2], f.b[3]);
        } else {
            const int lj = tid >> 2, lr = koff + (tid & 3) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) Bs[buf][lr + q][lj] = f.b[q];
        }
    };

    GemmFrag f[2];     // rows beyond rend fetch as zeros
    if (rbeg < rend) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            gemm_fetch(p, tid, i0, j0, rbeg + h * TKH, rend, a_r_contig, b_j_contig, a_vec, b_vec, f[h]);
            stash(0, h * TKH, f[h]);
        }
    }
    __syncthreads();
    int buf = 0;
    for (int r0 = rbeg; r0 < rend; r0 += TK) {
        const bool more = r0 + TK < rend;
        if (more) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
                gemm_fetch(p, tid, i0, j0, r0 + TK + h * TKH, rend, a_r_contig, b_j_contig, a_vec, b_vec, f[h]);
        }
        if (MMA) {
#pragma unroll
            for (int kk = 0; kk < TK; kk += 8) {
                uint32_t ah[2][4], al[2][4], bh[2][2], bl[2][2];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) {
                    const int r = wm * 32 + mi * 16 + fg;
                    split_tf32(As[buf][kk + ft][r], ah[mi][0], al[mi][0]);
                    split_tf32(As[buf][kk + ft][r + 8], ah[mi][1], al[mi][1]);
                    split_tf32(As[buf][kk + ft + 4][r], ah[mi][2], al[mi][2]);
                    split_tf32(As[buf][kk + ft + 4][r + 8], ah[mi][3], al[mi][3]);
                }
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) {
                    const int c = wn * 16 + ni * 8 + fg;
                    split_tf32(Bs[buf][kk + ft][c], bh[ni][0], bl[ni][0]);
                    split_tf32(Bs[buf][kk + ft + 4][c], bh[ni][1], bl[ni][1]);
                }
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 2; ++ni) {      // small terms first
                        mma_tf32(acc[mi * 2 + ni], al[mi], bh[ni]);
                        mma_tf32(acc[mi * 2 + ni], ah[mi], bl[ni]);
                        mma_tf32(acc[mi * 2 + ni], ah[mi], bh[ni]);
                    }
            }
        } else {
#pragma unroll
            for (int k = 0; k < TK; ++k) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
                const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
            }
        }
        if (more) {                    // the other buffer was last read before the previous barrier
#pragma unroll
            for (int h = 0; h < 2; ++h) stash(buf ^ 1, h * TKH, f[h]);
        }
        __syncthreads();
        buf ^= 1;
    }
    float* C = p.C + (long)blockIdx.z * p.split_stride;
    if (MMA) {
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int gi = i0 + wm * 32 + mi * 16 + fg + 8 * h;
                    const int gj = j0 + wn * 16 + ni * 8 + 2 * ft;
                    if (gi >= p.I) continue;
#pragma unroll
                    for (int y = 0; y < 2; ++y) {
                        if (gj + y >= p.J) continue;
                        float v = acc[mi * 2 + ni][2 * h + y];
                        if (p.splits == 1) {
                            if (p.bias) v += p.bias[gj + y];
                            if (p.relu) v = v < 0.f ? 0.f : v;  // NaN-propagating (an out-of-range user row must stay loud)
                            if (p.mask) v = p.mask[gi * p.ldmask + gj + y] > 0.f ? v : 0.f;
                        }
                        C[gi * p.ldc + gj + y] = v;
                    }
                }
        return;
    }
    const int gj0 = j0 + tx * 4;
    const bool c_vec = ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && ((p.ldc & 3) == 0) && gj0 + 3 < p.J &&
                       (!p.mask || (((reinterpret_cast<uintptr_t>(p.mask) & 15) == 0) && ((p.ldmask & 3) == 0)));
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        const int gi = i0 + ty * 4 + x;
        if (gi >= p.I) continue;
        if (c_vec) {   // one 16-byte store per output row segment
            float v[4] = {acc[x][0], acc[x][1], acc[x][2], acc[x][3]};
            if (p.splits == 1) {
                if (p.bias) {
#pragma unroll
                    for (int y = 0; y < 4; ++y) v[y] += p.bias[gj0 + y];
                }
                if (p.relu) {
#pragma unroll
                    for (int y = 0; y < 4; ++y) v[y] = v[y] < 0.f ? 0.f : v[y];
                }
                if (p.mask) {
                    const float4 m = *reinterpret_cast<const float4*>(p.mask + gi * p.ldmask + gj0);
                    v[0] = m.x > 0.f ? v[0] : 0.f; v[1] = m.y > 0.f ? v[1] : 0.f;
                    v[2] = m.z > 0.f ? v[2] : 0.f; v[3] = m.w > 0.f ? v[3] : 0.f;
                }
            }
            *reinterpret_cast<float4*>(C + gi * p.ldc + gj0) = make_float4(v[0], v[1], v[2], v[3]);
            continue;
        }
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int gj = j0 + tx * 4 + y;
            if (gj >= p.J) continue;
            float v = acc[x][y];
            if (p.splits == 1) {
                if (p.bias) v += p.bias[gj];
                if (p.relu) v = v < 0.f ? 0.f : v;  // NaN-propagating (an out-of-range user row must stay loud)
                if (p.mask) v = p.mask[gi * p.ldmask + gj] > 0.f ? v : 0.f;
            }
            C[gi * p.ldc + gj] = v;
        }
    }
}

// ------------------------------------------------------------------ cp.async pipelined 3xTF32 GEMM
// The register-staged kernel above is bound by global-load latency (two blocks per SM, one 16-byte load per thread and
// operand in flight: 34 us for an 11 MB GEMM, the same with CUDA-core FMAs or tensor-core MMAs).  Here 64 x 64 x 32 tiles go
// global -> shared memory with cp.async (16-byte chunks, zero-filled tails) through a 3-stage ring.  Either operand may be
// contiguous along the contraction index r (smem [i][r], row stride TKC+4) or along its own index (smem [r][i], row stride
// 72); both layouts give conflict-free m16n8k8 fragment reads.
// Round 2 (ncu: 2 060 instructions per warp for 64 MMAs, issue slots 55 % busy, 16 warps per SM): the operand layouts are
// template parameters (no per-load selects), 4 warps own 32 x 32 each (5 instead of 7 instructions per MMA), 55 KB of shared
// memory let 4 blocks share an SM, and the weight-gradient form also takes the bias gradient (row sums of its A operand, from
// the unrounded fp32 values already in registers) -- no separate column-sum launches.
#ifndef DCUE_GEMM_STAGES
#define DCUE_GEMM_STAGES 3
#endif
constexpr int TKC = 32, CST = DCUE_GEMM_STAGES;
constexpr int CP_OP_FLOATS = 64 * (TKC + 4);                // one operand tile of one stage (either layout fits: 32 x 72 = 64 x 36)
constexpr size_t CP_SMEM = (size_t)CST * 2 * CP_OP_FLOATS * sizeof(float);
constexpr int CP_THREADS = 128;

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, int src_bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}

// This thread's four 16-byte chunks of a [64 x TKC] operand tile X(i, r) = base[i*s_i + r*s_r]: everything that does not depend
// on the k-chunk is computed ONCE (the first version redid the index arithmetic for every chunk: 25 instructions per cp.async,
// as many as the tile's MMAs).  Rows i >= I and columns r >= rend are zero-filled (cp.async src-size).
template <bool RC>
struct TileLoader {
    const float* g[4];     // global address of the chunk at r0 = 0 (any valid address when the chunk is out of range)
    int lim[4];            // RC: the chunk's offset lr along r, or INT_MAX/2 when its row is out of range
                           // !RC: bytes to copy (0..16, fixed by the i-range), the row offset lr is it * 8 + tid / 16
    uint32_t soff[4];      // float offset inside the operand's shared-memory tile
    int lr0;
    __device__ __forceinline__ void init(const float* base, long s_i, long s_r, int i0, int I, int tid) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int c = tid + it * CP_THREADS;
            if (RC) {
                const int li = c / (TKC / 4), lr = (c % (TKC / 4)) * 4;
                const int gi = i0 + li;
                g[it] = gi < I ? base + gi * s_i + lr : base;
                lim[it] = gi < I ? lr : (1 << 29);
                soff[it] = (uint32_t)(li * (TKC + 4) + lr);
            } else {
                const int lr = c >> 4, li = (c & 15) * 4;
                const int gi = i0 + li;
                const int nb = min(16, (I - gi) * 4);
                g[it] = nb > 0 ? base + lr * s_r + gi : base;
                lim[it] = nb > 0 ? nb : 0;
                soff[it] = (uint32_t)(lr * 72 + li);
            }
        }
        lr0 = tid >> 4;
    }
    // chunk starting at r0 (rem = rend - r0 > 0 columns left) into the stage's tile `sm`
    __device__ __forceinline__ void issue(float* sm, long s_r, int r0, int rem) const {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            if (RC) {
                int bytes = (rem - lim[it]) * 4;
                bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
                cp_async16(sm + soff[it], bytes ? g[it] + r0 : g[it], bytes);
            } else {
                const int bytes = (lr0 + it * (CP_THREADS / 16)) < rem ? lim[it] : 0;
                cp_async16(sm + soff[it], bytes ? g[it] + r0 * s_r : g[it], bytes);
            }
        }
    }
};

template <bool SPLIT, bool ARC, bool BRC>
__global__ void __launch_bounds__(CP_THREADS) gemm_cpasync_kernel(GemmP p) {
    extern __shared__ __align__(16) float cps[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TN;
    const int wm = wid & 1, wn = wid >> 1, fg = lane >> 2, ft = lane & 3;
    const int rchunk = ceil_div_i(ceil_div_i(p.R, TKC), p.splits) * TKC;
    const int rbeg = blockIdx.z * rchunk;
    const int rend = min(p.R, rbeg + rchunk);
    const int nk = rbeg < rend ? ceil_div_i(rend - rbeg, TKC) : 0;
    const bool want_rowsum = p.rowsum != nullptr && blockIdx.x == 0 && wn == 0;

    float acc[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
    float rs[2][2] = {{0.f, 0.f}, {0.f, 0.f}};               // row sums of A: rows (mi, g) and (mi, g + 8)

    TileLoader<ARC> la;
    TileLoader<BRC> lb;
    la.init(p.A, p.sAi, p.sAr, i0, p.I, tid);
    lb.init(p.B, p.sBj, p.sBr, j0, p.J, tid);
    auto issue = [&](int kc) {
        if (kc < nk) {
            float* sa = cps + (size_t)(kc % CST) * 2 * CP_OP_FLOATS;
            const int r0 = rbeg + kc * TKC;
            la.issue(sa, p.sAr, r0, rend - r0);
            lb.issue(sa + CP_OP_FLOATS, p.sBr, r0, rend - r0);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");     // one group per k-chunk, empty ones included
    };
#pragma unroll
    for (int s = 0; s < CST - 1; ++s) issue(s);
    for (int kc = 0; kc < nk; ++kc) {
        asm volatile("cp.async.wait_group %0;" ::"n"(CST - 2) : "memory");   // chunk kc has landed (this thread's copies)
        __syncthreads();                                                      // ... everyone's; and chunk kc-1 is fully consumed
        issue(kc + CST - 1);                                                  // refill the slot of chunk kc-1
        const float* sa = cps + (size_t)(kc % CST) * 2 * CP_OP_FLOATS;
        const float* sb = sa + CP_OP_FLOATS;
#pragma unroll
        for (int kk = 0; kk < TKC; kk += 8) {
            uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                const int r = wm * 32 + mi * 16 + fg;
                float x0, x1, x2, x3;
                if (ARC) {
                    x0 = sa[r * (TKC + 4) + kk + ft]; x1 = sa[(r + 8) * (TKC + 4) + kk + ft];
                    x2 = sa[r * (TKC + 4) + kk + ft + 4]; x3 = sa[(r + 8) * (TKC + 4) + kk + ft + 4];
                } else {
                    x0 = sa[(kk + ft) * 72 + r]; x1 = sa[(kk + ft) * 72 + r + 8];
                    x2 = sa[(kk + ft + 4) * 72 + r]; x3 = sa[(kk + ft + 4) * 72 + r + 8];
                }
                rs[mi][0] += x0 + x2;
                rs[mi][1] += x1 + x3;
                split_tf32(x0, ah[mi][0], al[mi][0]); split_tf32(x1, ah[mi][1], al[mi][1]);
                split_tf32(x2, ah[mi][2], al[mi][2]); split_tf32(x3, ah[mi][3], al[mi][3]);
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int c = wn * 32 + ni * 8 + fg;
                float y0, y1;
                if (BRC) { y0 = sb[c * (TKC + 4) + kk + ft]; y1 = sb[c * (TKC + 4) + kk + ft + 4]; }
                else { y0 = sb[(kk + ft) * 72 + c]; y1 = sb[(kk + ft + 4) * 72 + c]; }
                split_tf32(y0, bh[ni][0], bl[ni][0]); split_tf32(y1, bh[ni][1], bl[ni][1]);
            }
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {      // small terms first
                    if (SPLIT) {
                        mma_tf32(acc[mi][ni], al[mi], bh[ni]);
                        mma_tf32(acc[mi][ni], ah[mi], bl[ni]);
                    }
                    mma_tf32(acc[mi][ni], ah[mi], bh[ni]);
                }
        }
    }
    float* C = p.C + (long)blockIdx.z * p.split_stride;
    // Epilogue (round 2: the per-element bias / relu / mask / bounds tests of the first version were 800 of the kernel's ~1 900
    // executed instructions): one warp-uniform mode switch, the bias pair of every column loaded once, 8-byte stores.
    const int mode = p.splits > 1 ? 0 : (p.mask ? 2 : ((p.bias || p.relu) ? 1 : 0));
    const bool st2 = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 7) == 0);
    float b0[4] = {0.f, 0.f, 0.f, 0.f}, b1[4] = {0.f, 0.f, 0.f, 0.f};
    if (mode == 1 && p.bias) {
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int gj = j0 + wn * 32 + ni * 8 + 2 * ft;
            if (gj < p.J) b0[ni] = p.bias[gj];
            if (gj + 1 < p.J) b1[ni] = p.bias[gj + 1];
        }
    }
    const bool relu = mode == 1 && p.relu;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int gi = i0 + wm * 32 + mi * 16 + fg + 8 * h;
            if (gi >= p.I) continue;
            float* crow = C + gi * p.ldc;
            const float* mrow = mode == 2 ? p.mask + gi * p.ldmask : nullptr;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int gj = j0 + wn * 32 + ni * 8 + 2 * ft;
                if (gj >= p.J) continue;
                float v0 = acc[mi][ni][2 * h] + b0[ni], v1 = acc[mi][ni][2 * h + 1] + b1[ni];
                if (relu) {                              // NaN-propagating (an out-of-range user row must stay loud)
                    v0 = v0 < 0.f ? 0.f : v0;
                    v1 = v1 < 0.f ? 0.f : v1;
                }
                const bool two = gj + 1 < p.J;
                if (mode == 2) {
                    v0 = mrow[gj] > 0.f ? v0 : 0.f;
                    if (two) v1 = mrow[gj + 1] > 0.f ? v1 : 0.f;
                }
                if (two && st2) *reinterpret_cast<float2*>(crow + gj) = make_float2(v0, v1);
                else {
                    crow[gj] = v0;
                    if (two) crow[gj + 1] = v1;
                }
            }
        }
    if (p.rowsum != nullptr) {
        // bias gradient of the weight-gradient form: sum over r of A(i, r), this block's r range -> rowsum[z][i]
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float t = rs[mi][h];
                t += __shfl_xor_sync(0xffffffffu, t, 1);
                t += __shfl_xor_sync(0xffffffffu, t, 2);
                const int gi = i0 + wm * 32 + mi * 16 + fg + 8 * h;
                if (want_rowsum && ft == 0 && gi < p.I) p.rowsum[(long)blockIdx.z * p.I + gi] = t;
            }
    }
}

// cp.async needs 16-byte aligned rows: the base pointer, the non-unit stride (x4 bytes) and the tile origin along the
// contiguous index are multiples of 4 floats (true for every DCUE shape: K, N in {100, 128, 300, 612}, ld multiples of 4)
bool cp_operand_ok(const float* base, long s_i, long s_r) {
    if (reinterpret_cast<uintptr_t>(base) & 15) return false;
    if (s_r == 1) return (s_i & 3) == 0;
    if (s_i == 1) return (s_r & 3) == 0;
    return false;
}

__global__ void split_reduce_kernel(const float* __restrict__ part, int splits, long n, float* __restrict__ out) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += part[(long)k * n + i];
    out[i] = s;
}
// the same for two partial arrays in one launch (weight gradient [splits][n] and bias gradient [splits][n2])
__global__ void split_reduce2_kernel(const float* __restrict__ part, int splits, long n, float* __restrict__ out,
                                     const float* __restrict__ part2, long n2, float* __restrict__ out2) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n + n2) return;
    const bool second = i >= n;
    const float* src = second ? part2 + (i - n) : part + i;
    const long stride = second ? n2 : n;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = 0;
    for (; k + 4 <= splits; k += 4) {
        s0 += src[(long)k * stride]; s1 += src[(long)(k + 1) * stride];
        s2 += src[(long)(k + 2) * stride]; s3 += src[(long)(k + 3) * stride];
    }
    for (; k < splits; ++k) s0 += src[(long)k * stride];
    const float s = (s0 + s1) + (s2 + s3);
    if (second) out2[i - n] = s; else out[i] = s;
}

// column sums of a [M, N] matrix (row stride ld): block (column chunk, row chunk g) writes
// part[g][n]; split_reduce_kernel then sums the row chunks in fixed order (deterministic).
constexpr int COLSUM_G = 64;
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int ld, int M, int N,
                                                     float* __restrict__ part) {
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int rl = threadIdx.x >> 5;
    float s = 0.f;
    if (c < N) {
        const int step = 8 * gridDim.y;
        int m = blockIdx.y * 8 + rl;
        for (; m + 3 * step < M; m += 4 * step) {      // four independent loads in flight (the plain loop was latency bound: 22 us)
            const float v0 = x[(long)m * ld + c], v1 = x[(long)(m + step) * ld + c];
            const float v2 = x[(long)(m + 2 * step) * ld + c], v3 = x[(long)(m + 3 * step) * ld + c];
            s += (v0 + v1) + (v2 + v3);
        }
        for (; m < M; m += step) s += x[(long)m * ld + c];
    }
    red[rl][threadIdx.x & 31] = s;
    __syncthreads();
    if (rl == 0 && c < N) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
        part[(long)blockIdx.y * N + c] = t;
    }
}

bool linear_simt() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DCUE_LINEAR_IMPL");
        v = (e && (e[0] == 's' || e[0] == 'S')) ? 1 : 0;
    }
    return v != 0;
}
int linear_impl() {      // 0 = cp.async pipeline (default), 1 = CUDA cores, 2 = register-staged 3xTF32
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DCUE_LINEAR_IMPL");
        v = !e ? 0 : ((e[0] == 's' || e[0] == 'S') ? 1 : ((e[0] == 'r' || e[0] == 'R') ? 2 : 0));
    }
    return v;
}
template <bool SPLIT, bool ARC, bool BRC>
void launch_cp(const GemmP& p, dim3 grid, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(gemm_cpasync_kernel<SPLIT, ARC, BRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CP_SMEM);
        attr_done = true;
    }
    gemm_cpasync_kernel<SPLIT, ARC, BRC><<<grid, CP_THREADS, CP_SMEM, st>>>(p);
}
template <bool SPLIT>
void launch_cp_layout(const GemmP& p, dim3 grid, cudaStream_t st) {
    const bool arc = p.sAr == 1, brc = p.sBr == 1;
    if (arc && brc) launch_cp<SPLIT, true, true>(p, grid, st);
    else if (arc) launch_cp<SPLIT, true, false>(p, grid, st);
    else if (brc) launch_cp<SPLIT, false, true>(p, grid, st);
    else launch_cp<SPLIT, false, false>(p, grid, st);
}
// -> true when the cp.async kernel ran (it honours p.rowsum); the other kernels ignore p.rowsum
bool launch_gemm(const GemmP& p, dim3 grid, cudaStream_t st) {
    const int impl = linear_impl();
    if (impl == 1) { gemm_kernel<false><<<grid, 256, 0, st>>>(p); return false; }
    if (impl == 0 && cp_operand_ok(p.A, p.sAi, p.sAr) && cp_operand_ok(p.B, p.sBj, p.sBr)) {
        if (p.single_tf32) launch_cp_layout<false>(p, grid, st);
        else launch_cp_layout<true>(p, grid, st);
        return true;
    }
    gemm_kernel<true><<<grid, 256, 0, st>>>(p);
    return false;
}
bool cp_path(const GemmP& p) {
    return linear_impl() == 0 && cp_operand_ok(p.A, p.sAi, p.sAr) && cp_operand_ok(p.B, p.sBj, p.sBr);
}

int wgrad_splits(int M, int K, int N) {
    const int tiles = ceil_div_i(N, TM) * ceil_div_i(K, TN);
    int s = (2 * dcue_num_sms() + tiles - 1) / tiles;
    const int maxs = ceil_div_i(M, 4 * TK);
    if (s > maxs) s = maxs;
    if (s < 1) s = 1;
    if (s > 64) s = 64;
    return s;
}

}  // namespace

static int linear_fwd_impl(const float* X, int ldx, const float* W, const float* b, int M, int K, int N,
                           int relu, float* Y, int ldy, void* stream, int single_tf32) {
    DCUE_CHECK_ARG(X && W && Y && M >= 0 && K > 0 && N > 0 && ldx >= K && ldy >= N);
    if (M == 0) return 0;
    GemmP p{};
    p.A = X; p.sAi = ldx; p.sAr = 1;
    p.B = W; p.sBr = 1; p.sBj = K;  // B(r,j) = W[j,r]
    p.C = Y; p.ldc = ldy; p.bias = b; p.mask = nullptr; p.ldmask = 0;
    p.I = M; p.J = N; p.R = K; p.relu = relu; p.splits = 1; p.split_stride = 0; p.single_tf32 = single_tf32;
    dim3 grid(ceil_div_i(N, TN), ceil_div_i(M, TM), 1);
    launch_gemm(p, grid, (cudaStream_t)stream);
    DCUE_LAUNCH_CHECK();
    return 0;
}

static int linear_dgrad_impl(const float* dY, int lddy, const float* W, int M, int K, int N,
                             const float* mask, int ldmask, float* dX, int lddx, void* stream, int single_tf32) {
    DCUE_CHECK_ARG(dY && W && dX && M >= 0 && K > 0 && N > 0 && lddy >= N && lddx >= K);
    if (M == 0) return 0;
    GemmP p{};
    p.A = dY; p.sAi = lddy; p.sAr = 1;
    p.B = W; p.sBr = K; p.sBj = 1;  // B(r=n, j=k) = W[n,k]
    p.C = dX; p.ldc = lddx; p.bias = nullptr; p.mask = mask; p.ldmask = ldmask;
    p.I = M; p.J = K; p.R = N; p.relu = 0; p.splits = 1; p.split_stride = 0; p.single_tf32 = single_tf32;
    dim3 grid(ceil_div_i(K, TN), ceil_div_i(M, TM), 1);
    launch_gemm(p, grid, (cudaStream_t)stream);
    DCUE_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t dcue_linear_wgrad_ws_bytes(int M, int K, int N) {
    return ((size_t)wgrad_splits(M, K, N) * (size_t)N * (size_t)K + (size_t)COLSUM_G * N) * sizeof(float) + 256;
}

static int linear_wgrad_impl(const float* dY, int lddy, const float* X, int ldx, int M, int K, int N,
                             float* dW, float* db, void* ws, size_t ws_bytes, void* stream, int single_tf32) {
    DCUE_CHECK_ARG(dY && X && dW && M >= 0 && K > 0 && N > 0 && lddy >= N && ldx >= K);
    cudaStream_t st = (cudaStream_t)stream;
    if (M == 0) {
        DCUE_CUDA(cudaMemsetAsync(dW, 0, (size_t)N * K * sizeof(float), st));
        if (db) DCUE_CUDA(cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), st));
        return 0;
    }
    const int splits = wgrad_splits(M, K, N);
    if (!ws || ws_bytes < ((size_t)splits * N * K + (size_t)COLSUM_G * N) * sizeof(float))
        DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_linear_wgrad: workspace too small");
    GemmP p{};
    p.A = dY; p.sAi = 1; p.sAr = lddy;  // A(i=n, r=m) = dY[m,n]
    p.B = X; p.sBr = ldx; p.sBj = 1;    // B(r=m, j=k) = X[m,k]
    p.C = splits > 1 ? (float*)ws : dW; p.ldc = K; p.bias = nullptr; p.mask = nullptr; p.ldmask = 0;
    p.I = N; p.J = K; p.R = M; p.relu = 0; p.splits = splits; p.split_stride = (long)N * K; p.single_tf32 = single_tf32;
    float* part = (float*)ws + (size_t)splits * N * K;      // bias-gradient partials: [splits][N] (fused) or [COLSUM_G][N]
    const bool fuse_db = db != nullptr && cp_path(p);        // splits <= 64 = COLSUM_G rows fit the same region
    p.rowsum = fuse_db ? (splits > 1 ? part : db) : nullptr;
    dim3 grid(ceil_div_i(K, TN), ceil_div_i(N, TM), splits);
    launch_gemm(p, grid, st);
    DCUE_LAUNCH_CHECK();
    const long n = (long)N * K;
    if (splits > 1) {
        if (fuse_db) split_reduce2_kernel<<<ceil_div_i(n + N, 256), 256, 0, st>>>((const float*)ws, splits, n, dW, part, N, db);
        else split_reduce_kernel<<<ceil_div_i(n, 256), 256, 0, st>>>((const float*)ws, splits, n, dW);
        DCUE_LAUNCH_CHECK();
    }
    if (db && !fuse_db) {
        dim3 cg(ceil_div_i(N, 32), COLSUM_G);
        colsum_kernel<<<cg, 256, 0, st>>>(dY, lddy, M, N, part);
        DCUE_LAUNCH_CHECK();
        split_reduce_kernel<<<ceil_div_i(N, 256), 256, 0, st>>>(part, COLSUM_G, N, db);
        DCUE_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int dcue_linear_fwd(const float* X, int ldx, const float* W, const float* b, int M, int K, int N,
                               int relu, float* Y, int ldy, void* stream) {
    return linear_fwd_impl(X, ldx, W, b, M, K, N, relu, Y, ldy, stream, 0);
}
extern "C" int dcue_linear_dgrad(const float* dY, int lddy, const float* W, int M, int K, int N,
                                 const float* mask, int ldmask, float* dX, int lddx, void* stream) {
    return linear_dgrad_impl(dY, lddy, W, M, K, N, mask, ldmask, dX, lddx, stream, 0);
}
extern "C" int dcue_linear_wgrad(const float* dY, int lddy, const float* X, int ldx, int M, int K, int N,
                                 float* dW, float* db, void* ws, size_t ws_bytes, void* stream) {
    return linear_wgrad_impl(dY, lddy, X, ldx, M, K, N, dW, db, ws, ws_bytes, stream, 0);
}
// Single-pass TF32 variants (operands rounded to TF32 = 11 significant bits, the precision of the fp16 conv operands; fp32
// accumulate): the song tower's k = 1 conv and fc, whose inputs already carry 16-bit operand rounding.
extern "C" int dcue_linear_fwd_tf32(const float* X, int ldx, const float* W, const float* b, int M, int K, int N,
                                    int relu, float* Y, int ldy, void* stream) {
    return linear_fwd_impl(X, ldx, W, b, M, K, N, relu, Y, ldy, stream, 1);
}
extern "C" int dcue_linear_dgrad_tf32(const float* dY, int lddy, const float* W, int M, int K, int N,
                                      const float* mask, int ldmask, float* dX, int lddx, void* stream) {
    return linear_dgrad_impl(dY, lddy, W, M, K, N, mask, ldmask, dX, lddx, stream, 1);
}
extern "C" int dcue_linear_wgrad_tf32(const float* dY, int lddy, const float* X, int ldx, int M, int K, int N,
                                      float* dW, float* db, void* ws, size_t ws_bytes, void* stream) {
    return linear_wgrad_impl(dY, lddy, X, ldx, M, K, N, dW, db, ws, ws_bytes, stream, 1);
}
