// tcgen05 / TMEM / TMA-bulk implicit-GEMM kernels for the DCUE song tower (sm_100a only).
//
// Reference ops replaced: nn.Conv1d (+MaxPool1d+ReLU and BatchNorm statistics) forward, its
// data gradient and its weight gradient (truedcuemel1dbn.py:25-54,80-95 and autograd thereof).
//
// Formulation (see dcue_b200.h for the panel layout): channels on the MMA M axis, flat time
// rows on the N axis, so one TMEM lane = one output channel and
//   * max-pooling over time is a per-thread register max over adjacent accumulator columns,
//   * bias / ReLU / BatchNorm partial sums are per-thread scalars (no shuffles, no atomics),
//   * the k conv taps are k K-blocks whose B descriptors differ by +16 bytes (one row): the
//     no-swizzle panel layout makes a tap shift a plain descriptor address offset, so the
//     activation tile is staged ONCE per tile by cp.async.bulk (no im2col, no re-load per tap).
//
//   forward / dgrad:  D[ch, r] = sum_{j<k} sum_{c<128} A[ch][j*128+c] * In[r+j, c]
//       A = packed weights, resident in smem for the whole persistent CTA (K-major, 128 x k*128)
//       B = activation tile of 128+8 rows x 128 channels (K-major), 2-stage mbarrier ring
//       D = 128 lanes x 128 columns fp32 in TMEM, double buffered (epilogue overlaps next MMA)
//   wgrad:            D_j[co, ci] = sum_r dY[r, co] * X[r+j, ci]      (both operands MN-major)
//       k accumulators of 128 columns stay in TMEM for the CTA's whole row range.
//
// Warp roles (192 threads): warp 0 = bulk-copy producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).
#include "common.cuh"
#include "conv_common.cuh"
#include <stdlib.h>

namespace {

constexpr int BN = 128;                       // flat rows per tile (MMA N)
constexpr int HALO = 8;                       // extra rows staged for the taps (>= k-1, keeps 128 B groups)
constexpr int PANELS = 16;                    // 128 channels / 8
constexpr int ROWB = 16;                      // bytes per (row, panel) chunk
constexpr int B_PANEL_BYTES = (BN + HALO) * ROWB;          // 2176
constexpr int B_STAGE_BYTES = PANELS * B_PANEL_BYTES;       // 34816
constexpr int A_PANEL_BYTES = 128 * ROWB;                   // 2048 (128 M rows)
constexpr int NSTAGE = 2;
constexpr int NTHREADS = 192;

// ----------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// A operand from TMEM (TS mode): lane = row of A, each 32-bit column = two consecutive K elements
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// predicated global stores: dead lanes issue nothing.  (They used to write one shared scratch word instead; ncu
// showed every SM serialising on that single L2 sector -- the layer-2 dgrad spent its whole 270 us there.)
__device__ __forceinline__ void st_pred_f32(float* p, float v, bool ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %2, 0;\n\t"
        "@p st.global.f32 [%0], %1;\n\t"
        "}" ::"l"(p), "f"(v), "r"((int)ok) : "memory");
}
__device__ __forceinline__ void st_pred_u8(uint8_t* p, uint32_t v, bool ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %2, 0;\n\t"
        "@p st.global.u8 [%0], %1;\n\t"
        "}" ::"l"(p), "r"(v), "r"((int)ok) : "memory");
}

// shared-memory matrix descriptor, SWIZZLE_NONE (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4 | [16,30) leading byte offset>>4 | [32,46) stride byte offset>>4 | [46,48) version=1
// K-major : LBO = distance between the two 8-element K chunks of one MMA, SBO = between 8-row groups
// MN-major: LBO = distance between 8-row K groups,              SBO = between 8-element MN chunks
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor), dense, fp32 accumulate
__device__ __forceinline__ uint32_t make_idesc(int a_fmt, int b_fmt, int a_mn_major, int b_mn_major, int M, int N) {
    return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)a_mn_major << 15) |
           ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Pipe {
    int stage = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance(int n) {
        if (++stage == n) { stage = 0; phase ^= 1; }
    }
};

// ----------------------------------------------------------------------------- fwd / dgrad
// Folded input-BatchNorm shift: conv outputs whose taps hang over the zero padding lack those taps' constants.  Kept out of
// line so that the (rare, warp-uniform) border windows do not bloat the common epilogue path with predicated code.
__device__ __noinline__ float border_fix(float tb0, float tb1, float tb2, float tb3, int t, int k, int pad, int Lin) {
    const float tb[4] = {tb0, tb1, tb2, tb3};
    return missing_taps(tb, t, k, pad, Lin);
}

// EPI 0: bias + maxpool(POOL) + relu + argmax code + BN partial sums -> z[S*P, Cout]
// EPI 1: store the data rows -> dx[S*Lin, Cout]
// Epilogue work split: 16 warps, warp e owns TMEM lane quarter (warp_id % 4) x column chunk e/4 of
// every tile, i.e. exactly one tcgen05.ld.32x32b.x32 per tile; 4 warps per scheduler hide each other's
// latencies (ncu showed a single epilogue warp per scheduler stalling on fixed-latency dependencies).
constexpr int NEPI = 16;
constexpr int NTHREADS_ROWS = 64 + NEPI * 32;

// EPI 0: bias + maxpool(POOL) + relu + argmax code + BN partial sums -> z[S*P, Cout]
// EPI 1: store the data rows (scaled by 1/s of the gradient operand)   -> dx[S*Lin, Cout]
// ATMEM: the packed weights live in TMEM columns [0, k*64) (TS-mode MMA) instead of shared memory, which
// leaves room for NST = 6 activation stages instead of 2 (the 2-stage ring was load-latency bound) and
// removes the A-operand shared-memory reads.
// EPI 2 = EPI 1 with BnBwdStats: the data gradient dx this kernel writes IS the gradient dy entering the previous
// stage's BatchNorm, so the BatchNorm-backward reductions (sum dy, sum dy*xhat, max|dy|; dcue_bn_bwd_reduce) are taken in
// this epilogue while dx is still in registers -- one extra coalesced read of z instead of a separate sweep over dx and z
// (103 us at layer 1).  Partial rows as in the forward: [(block, chunk)][2][Cout] doubles, then [(block, chunk)] max|dy|.
struct BnBwdStats {
    const float* z;        // [S*Lin, Cout] pre-BatchNorm activations of the previous stage (same row space as dx)
    const float* mean;     // per channel (batch statistics) -- or zeros / ones for the plain sums
    const float* rstd;
    const float* dtp;      // nullable [S, lddtp]: gradient of the time average, added as dtp / Lin to every row
    int lddtp;
    float inv_rows;        // 1 / Lin
};

template <int EPI, int POOL, bool ATMEM>
__global__ void __launch_bounds__(NTHREADS_ROWS, 1)
tc_conv_rows_kernel(const uint4* __restrict__ panel, long panel_rows, int fmt_in, const uint4* __restrict__ wp, int fmt_w,
                    const float* __restrict__ bias, ConvGeom g, float* __restrict__ out, uint8_t* __restrict__ code,
                    double* __restrict__ partial, const float* __restrict__ gscale, const float* __restrict__ tap_bias,
                    float* __restrict__ dummy, BnBwdStats bs) {
    __shared__ float smax[4][4];
    __shared__ double sred[4][2][128];
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NST = ATMEM ? 6 : NSTAGE;
    constexpr int ACC0 = ATMEM ? 256 : 0;              // first accumulator column
    const int a_bytes = ATMEM ? 0 : g.k * PANELS * A_PANEL_BYTES;  // k * 32 KB
    uint8_t* sA = smem;
    uint8_t* sB = smem + a_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + NST * B_STAGE_BYTES);
    // bars: [0..NST) full, [NST..2NST) empty, [2N] wfull, [2N+1,2N+2] tmem_full, [2N+3,2N+4] tmem_empty
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 5);
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (NST + s); };
    const uint32_t WFULL = bar0 + 8u * (2 * NST);
    auto TFULL = [&](int a) { return bar0 + 8u * (2 * NST + 1 + a); };
    auto TEMPTY = [&](int a) { return bar0 + 8u * (2 * NST + 3 + a); };

    const long ntiles = (g.rows_total + BN - 1) / BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        mbar_init(WFULL, ATMEM ? 4 : 1);
        for (int a = 0; a < 2; ++a) { mbar_init(TFULL(a), 1); mbar_init(TEMPTY(a), NEPI); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<ATMEM ? 512 : 256>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== producer: weights once, then one activation tile per stage =====
        if (lane == 0) {
            if (!ATMEM) {
                mbar_expect_tx(WFULL, (uint32_t)a_bytes);
                for (int c = 0; c < g.k * 4; ++c)  // 8 KB pieces
                    bulk_g2s(smem_u32(sA + c * 8192), reinterpret_cast<const uint8_t*>(wp) + (size_t)c * 8192, 8192, WFULL);
            }
            Pipe pp;
            for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                mbar_wait(EMPTY(pp.stage), pp.phase ^ 1);
                mbar_expect_tx(FULL(pp.stage), B_STAGE_BYTES);
                const long r0 = tile * BN;
                uint8_t* dst = sB + pp.stage * B_STAGE_BYTES;
#pragma unroll 4
                for (int q = 0; q < PANELS; ++q)
                    bulk_g2s(smem_u32(dst + q * B_PANEL_BYTES), panel + (long)q * panel_rows + r0, B_PANEL_BYTES,
                             FULL(pp.stage));
                pp.advance(NST);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc(fmt_w, fmt_in, 0, 0, 128, BN);
            mbar_wait(WFULL, 0);
            tc_fence_after();
            Pipe pp;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                mbar_wait(TEMPTY(acc), acc_phase ^ 1);
                mbar_wait(FULL(pp.stage), pp.phase);
                tc_fence_after();
                const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB + pp.stage * B_STAGE_BYTES);
                const uint32_t d = tmem_base + (uint32_t)(ACC0 + acc * BN);
                for (int j = 0; j < g.k; ++j) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {  // 16 channels = 2 panels per MMA
                        const uint64_t bd = make_desc(b0 + (uint32_t)(2 * c * B_PANEL_BYTES + j * ROWB), B_PANEL_BYTES, 128);
                        if (ATMEM) {
                            umma_f16_ts(d, tmem_base + (uint32_t)(j * 64 + c * 8), bd, idesc, (j | c) != 0);
                        } else {
                            const uint64_t ad = make_desc(a0 + (uint32_t)((j * PANELS + 2 * c) * A_PANEL_BYTES), A_PANEL_BYTES, 128);
                            umma_f16(d, ad, bd, idesc, (j | c) != 0);
                        }
                    }
                }
                umma_commit(EMPTY(pp.stage));  // smem stage reusable once these MMAs retire
                umma_commit(TFULL(acc));       // accumulator ready for the epilogue
                pp.advance(NST);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue warps: TMEM -> registers -> (pool, relu, stats) -> global =====
        const int quarter = warp & 3;            // TMEM lane quarter this warp may read
        const int chunk = (warp - 2) >> 2;       // 32-column chunk of every tile
        const int m = quarter * 32 + lane;       // output channel == TMEM lane
        const bool chan_ok = m < g.Cout;
        float bv = EPI == 0 ? ((bias && chan_ok) ? bias[m] : 0.f) : (gscale ? gscale[1] : 1.f);
        float tb[4] = {0.f, 0.f, 0.f, 0.f};
        const bool has_tb = EPI == 0 && tap_bias != nullptr;
        if (has_tb && chan_ok)
            for (int j = 0; j < g.k; ++j) { tb[j] = tap_bias[j * g.Cout + m]; bv += tb[j]; }
        if (ATMEM && chunk == 0) {
            // the four chunk-0 warps cover the four lane quarters: copy this thread's weight row into TMEM.
            // packed weights [K/8][128 rows][8] -> row m, K chunk kk8 is one uint4 = columns 4*kk8 .. 4*kk8+3
            for (int c0 = 0; c0 < g.k * 16; c0 += 8) {
                uint32_t r[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 w = __ldg(wp + (long)(c0 + i) * 128 + m);
                    r[4 * i] = w.x; r[4 * i + 1] = w.y; r[4 * i + 2] = w.z; r[4 * i + 3] = w.w;
                }
                tmem_st32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c0 * 4), r);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(WFULL);
        }
        // BatchNorm partial sums across the CTA's tiles: compensated (Kahan) fp32 instead of fp64 -- the two DADDs per warp and
        // tile took 11 % of the kernel's stall samples (the fp64 pipe is narrow) and the compensated sum is as accurate here
        float st1 = 0.f, st1c = 0.f, st2 = 0.f, st2c = 0.f;
        constexpr bool dstats = EPI == 2;      // own instantiation: the statistics cost registers the plain dgrad keeps free
        float bmean = 0.f, brstd = 1.f, amax = 0.f;
        if (dstats && chan_ok) { bmean = bs.mean[m]; brstd = bs.rstd[m]; }
        int acc = 0;
        uint32_t acc_phase = 0;
        for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int r = (int)(tile * BN) + chunk * 32;
            int s = r / g.Lp;
            int q = r - s * g.Lp;
            // EPI 1 with statistics: the 32 z values this thread needs are requested BEFORE waiting for the accumulator, all at
            // once (32 x 128 B per warp in flight).  Valid rows (pad <= q < pad + Lin, s < S) map to CONSECUTIVE output rows, so one
            // running offset serves every chunk -- with Lin = 33 only 2 of 36 chunk positions lie inside one spectrogram and the
            // first version's per-row fallback (dependent loads behind the stores) made the layer-2 launch 5x slower.
            float zz[EPI == 2 ? 32 : 1];
            long o_first = 0;
            uint32_t vm = 0, wm = 0;      // bit t: row t is a data row / a new spectrogram starts after row t (warp-uniform)
            if (dstats) {
                const int tt0 = min(max(q - g.pad, 0), g.Lin);
                o_first = ((long)s * g.Lin + tt0) * g.Cout + m;
                int s2 = s, q2 = q;
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    vm |= (s2 < g.S && q2 >= g.pad && q2 < g.pad + g.Lin) ? (1u << t) : 0u;
                    ++q2;
                    if (q2 == g.Lp) { q2 = 0; ++s2; wm |= 1u << t; }
                }
                const float* zp = bs.z + o_first;
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    const int nb = __popc(vm & ((1u << t) - 1u));          // data rows before t
                    zz[t] = (((vm >> t) & 1u) && chan_ok) ? __ldg(zp + (long)nb * g.Cout) : bmean;
                }
            }
            mbar_wait(TFULL(acc), acc_phase);
            tc_fence_after();
            const uint32_t tsrc = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ACC0 + acc * BN + chunk * 32);
            float v[32];
            if constexpr (!dstats) {
                tmem_ld32(tsrc, v);
                // accumulator columns are in registers: release the TMEM buffer before the global stores
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(TEMPTY(acc));
            }
            if (EPI == 0) {
                // ncu (round 2): this epilogue, not the MMA, paced the kernel -- 510 instructions per warp and tile, the ALU pipe
                // saturated (math-pipe-throttle the top stall) because the border correction and the (s, p) index arithmetic
                // were if-converted into predicated code that issued for every output.  Rows are warp-uniform, so each
                // pooling window now takes a real branch: 32 of a spectrogram's 34 windows are "interior" (all taps inside the
                // data, same spectrogram, p < P) and run a short straight-line path with an incrementally advanced pointer.
                float ts1 = 0.f, ts2 = 0.f;
                const int q_hi = g.Lin + g.pad - g.k;            // last conv row whose taps all hit data rows
                const int q_end = g.P * POOL;                    // rows beyond are dropped by the floor pooling
                long o = ((long)s * g.P + q / POOL) * g.Cout + m;
                auto pool_store = [&](const float (&w)[POOL], bool ok, long oo) {
                    float best = w[0];
                    int bi = 0;
#pragma unroll
                    for (int i = 1; i < POOL; ++i) {
                        const bool gt = w[i] > best;   // first maximum wins, like ATen
                        best = gt ? w[i] : best;
                        bi = gt ? i : bi;
                    }
                    const float val = ok ? fmaxf(best + bv, 0.f) : 0.f;
                    st_pred_f32(out + oo, val, ok);
                    st_pred_u8(code + oo, (uint32_t)bi, ok);
                    ts1 += val;
                    ts2 = fmaf(val, val, ts2);
                };
                if (s < g.S && q >= g.pad && q + 31 <= q_hi && q + 32 <= q_end) {
                    // the whole 32-row chunk is interior (about 2/3 of the chunks): straight-line, no per-window tests
#pragma unroll
                    for (int t0 = 0; t0 < 32; t0 += POOL) {
                        float w[POOL];
#pragma unroll
                        for (int i = 0; i < POOL; ++i) w[i] = v[t0 + i];
                        pool_store(w, chan_ok, o);
                        o += g.Cout;
                    }
                } else {
                    // chunk touching a spectrogram border (about 1/3 of the chunks): per-window tests, all warp-uniform
#pragma unroll
                    for (int t0 = 0; t0 < 32; t0 += POOL) {
                        float w[POOL];
#pragma unroll
                        for (int i = 0; i < POOL; ++i) w[i] = v[t0 + i];
                        const bool ok = chan_ok && s < g.S && q + POOL <= q_end;
                        if (has_tb && (q < g.pad || q + POOL - 1 > q_hi)) {
#pragma unroll
                            for (int i = 0; i < POOL; ++i) w[i] -= border_fix(tb[0], tb[1], tb[2], tb[3], q + i, g.k, g.pad, g.Lin);
                        }
                        pool_store(w, ok, o);
                        q += POOL;
                        o += g.Cout;
                        if (q >= g.Lp) {                         // next spectrogram (warp-uniform)
                            q -= g.Lp;
                            ++s;
                            o = ((long)s * g.P + q / POOL) * g.Cout + m;
                        }
                    }
                }
                {
                    const float y1 = ts1 - st1c, t1 = st1 + y1;
                    st1c = (t1 - st1) - y1;
                    st1 = t1;
                    const float y2 = ts2 - st2c, t2 = st2 + y2;
                    st2c = (t2 - st2) - y2;
                    st2 = t2;
                }
            } else if (dstats) {
                // two 16-column halves: 32 z values + 16 accumulator columns live instead of 64 registers
                float* dst = out + o_first;
                float ts1 = 0.f, ts2 = 0.f;
                float tpv = 0.f;
                int s_tp = -1;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float vh[16];
                    tmem_ld16(tsrc + (uint32_t)(h * 16), vh);
                    if (h == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(TEMPTY(acc));
                    }
#pragma unroll
                    for (int t = 0; t < 16; ++t) {
                        const int tt = h * 16 + t;
                        const bool okr = (vm >> tt) & 1u;                                 // warp-uniform
                        const bool ok = okr && chan_ok;
                        if (bs.dtp && okr) {                                              // residual towers: once per spectrogram
                            const int st = s + __popc(wm & ((1u << tt) - 1u));
                            if (st != s_tp) {
                                tpv = chan_ok ? __ldg(bs.dtp + (long)st * bs.lddtp + m) * bs.inv_rows : 0.f;
                                s_tp = st;
                            }
                        }
                        const float val = vh[t] * bv;
                        st_pred_f32(dst + (long)__popc(vm & ((1u << tt) - 1u)) * g.Cout, val, ok);
                        const float gg = ok ? val + tpv : 0.f;
                        ts1 += gg;
                        ts2 = fmaf(gg, (zz[tt] - bmean) * brstd, ts2);
                        amax = fmaxf(amax, fabsf(gg));
                    }
                }
                const float y1 = ts1 - st1c, t1 = st1 + y1;
                st1c = (t1 - st1) - y1;
                st1 = t1;
                const float y2 = ts2 - st2c, t2 = st2 + y2;
                st2c = (t2 - st2) - y2;
                st2 = t2;
            } else if (s < g.S && q >= g.pad && q + 31 < g.Lin + g.pad) {
                // all 32 rows are data rows of one spectrogram: one base pointer, no per-row tests
                float* dst = out + ((long)s * g.Lin + (q - g.pad)) * g.Cout + m;
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    st_pred_f32(dst, v[t] * bv, chan_ok);
                    dst += g.Cout;
                }
            } else {
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    const int tt = q - g.pad;
                    const bool ok = chan_ok && s < g.S && tt >= 0 && tt < g.Lin;
                    const long oo = ((long)s * g.Lin + tt) * g.Cout + m;
                    st_pred_f32(out + oo, v[t] * bv, ok);
                    ++q;
                    const bool wrap = q == g.Lp;
                    q = wrap ? 0 : q;
                    s += wrap ? 1 : 0;
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if ((EPI == 0 || dstats) && partial) {
            // the four chunk-warps of a channel combine through shared memory (fixed order): ONE partial row per block, so
            // the finaliser reads 148 rows instead of 592
            sred[chunk][0][m] = (double)st1 - (double)st1c;
            sred[chunk][1][m] = (double)st2 - (double)st2c;
            if (dstats) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
                if (lane == 0) smax[chunk][quarter] = amax;
            }
            asm volatile("bar.sync 3, %0;" ::"n"(NEPI * 32) : "memory");
            if (chunk == 0 && chan_ok) {
                partial[((long)blockIdx.x * 2 + 0) * g.Cout + m] = (sred[0][0][m] + sred[1][0][m]) + (sred[2][0][m] + sred[3][0][m]);
                partial[((long)blockIdx.x * 2 + 1) * g.Cout + m] = (sred[0][1][m] + sred[1][1][m]) + (sred[2][1][m] + sred[3][1][m]);
            }
            if (dstats && warp == 2 && lane == 0) {
                float mx = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) mx = fmaxf(mx, smax[c][qq]);
                double* pmax = partial + (size_t)gridDim.x * 2 * g.Cout;
                pmax[blockIdx.x] = (double)mx;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<ATMEM ? 512 : 256>(tmem_base);
    }
}

// ----------------------------------------------------------------------------- wgrad
constexpr int WG_DY_PANEL = BN * ROWB;                       // 2048
constexpr int WG_STAGE_BYTES = PANELS * WG_DY_PANEL + B_STAGE_BYTES;   // 32768 + 34816
constexpr int WG_NSTAGE = 3;

template <int KTAPS>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_wgrad_kernel(const uint4* __restrict__ dyp, long dy_rows, int fmt_dy, const uint4* __restrict__ xp, long x_rows,
                int fmt_x, long rows_total, float* __restrict__ part /* [grid][128][KTAPS][128] */) {
    constexpr int TCOLS = KTAPS * 128 <= 128 ? 128 : (KTAPS * 128 <= 256 ? 256 : 512);
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_NSTAGE * WG_STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_NSTAGE + 1);
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (WG_NSTAGE + s); };
    const uint32_t DONE = bar0 + 8u * (2 * WG_NSTAGE);

    const long ntiles = (rows_total + BN - 1) / BN;
    // balanced contiguous ranges: grid <= ntiles (tc_grid), so every CTA owns >= 1 tile and its TMEM
    // accumulators are always written before the epilogue reads them
    const long tbeg = ntiles * blockIdx.x / gridDim.x;
    const long tend = ntiles * (blockIdx.x + 1) / gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < WG_NSTAGE; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        mbar_init(DONE, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TCOLS>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            Pipe pp;
            for (long tile = tbeg; tile < tend; ++tile) {
                mbar_wait(EMPTY(pp.stage), pp.phase ^ 1);
                mbar_expect_tx(FULL(pp.stage), WG_STAGE_BYTES);
                const long r0 = tile * BN;
                uint8_t* dy_dst = smem + pp.stage * WG_STAGE_BYTES;
                uint8_t* x_dst = dy_dst + PANELS * WG_DY_PANEL;
#pragma unroll 4
                for (int q = 0; q < PANELS; ++q) {
                    bulk_g2s(smem_u32(dy_dst + q * WG_DY_PANEL), dyp + (long)q * dy_rows + r0, WG_DY_PANEL, FULL(pp.stage));
                    bulk_g2s(smem_u32(x_dst + q * B_PANEL_BYTES), xp + (long)q * x_rows + r0, B_PANEL_BYTES, FULL(pp.stage));
                }
                pp.advance(WG_NSTAGE);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(fmt_dy, fmt_x, 1, 1, 128, 128);
            Pipe pp;
            bool first = true;
            for (long tile = tbeg; tile < tend; ++tile) {
                mbar_wait(FULL(pp.stage), pp.phase);
                tc_fence_after();
                const uint32_t dy0 = smem_u32(smem + pp.stage * WG_STAGE_BYTES);
                const uint32_t x0 = dy0 + PANELS * WG_DY_PANEL;
#pragma unroll
                for (int kk = 0; kk < BN / 16; ++kk) {  // 16 rows (MMA K) per instruction
                    const uint64_t ad = make_desc(dy0 + (uint32_t)(kk * 16 * ROWB), 128, WG_DY_PANEL);
#pragma unroll
                    for (int j = 0; j < KTAPS; ++j) {
                        const uint64_t bd = make_desc(x0 + (uint32_t)((kk * 16 + j) * ROWB), 128, B_PANEL_BYTES);
                        umma_f16(tmem_base + (uint32_t)(j * 128), ad, bd, idesc, !(first && kk == 0));
                    }
                }
                first = false;
                umma_commit(EMPTY(pp.stage));
                pp.advance(WG_NSTAGE);
            }
            umma_commit(DONE);
        }
        __syncwarp();
    } else {
        const int quarter = warp & 3;
        const int co = quarter * 32 + lane;
        mbar_wait(DONE, 0);
        tc_fence_after();
        float* dst = part + ((long)blockIdx.x * 128 + co) * KTAPS * 128;
#pragma unroll 1
        for (int j = 0; j < KTAPS; ++j) {
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(j * 128 + ch * 32), v);
                if (tend <= tbeg) {  // rows_total == 0: nothing was accumulated
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.f;
                }
                float4* d4 = reinterpret_cast<float4*>(dst + j * 128 + ch * 32);
#pragma unroll
                for (int i = 0; i < 8; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TCOLS>(tmem_base);
    }
}

// ----------------------------------------------------------------------------- fused unpool + wgrad
// dW[co,ci,j] = sum_r dY[r,co] X[r+j,ci] where dY is never materialised: the 16-bit gradient operand tile is built in
// shared memory by 8 "expander" warps straight from the pooled fp32 inputs (gradient dy, pre-BatchNorm activation z,
// argmax code): BatchNorm backward (a*dy + b*z + c per channel), ReLU mask, scale, fp16, and the value is placed on
// the row of its pooling window that won the max (the other POOL-1 rows are zero).  Versus unpool -> panel -> wgrad
// this removes one write and one read of the 4x-unpooled panel (2 x 749 MB at layer 1) and a launch.
// lane = pooling window of the 128-row tile (32 windows x POOL 4), warp e = channel panels 2e, 2e+1.  Global loads go
// straight to registers (32-byte sector per lane, as in bn_relu_unpool_rows_kernel) one tile ahead; the smem stores
// rotate the row order per lane pair so that a warp store instruction touches every bank once.
constexpr int WGU_EXP = 8;                          // expander warps
constexpr int WGU_DY_PANEL = WG_DY_PANEL + ROWB;    // dY operand panel stride of the fused kernel (+16 B: bank spread)
constexpr int WGU_STAGE_BYTES = PANELS * WGU_DY_PANEL + B_STAGE_BYTES;
constexpr int WGU_THREADS = 64 + WGU_EXP * 32;

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct UnpoolSrc {
    const float* dy; int lddy;
    const float* dtp; int lddtp;
    const float* z; const uint8_t* code;
    const float* scale; const float* mean; const float* rstd; const double* sums; double count;
    int S, P, Lp;
    const float* gscale;
};

template <int KTAPS>
__global__ void __launch_bounds__(WGU_THREADS, 1)
tc_wgrad_unpool_kernel(UnpoolSrc src, int fmt, const uint4* __restrict__ xp, long x_rows, long rows_total,
                       float* __restrict__ part /* [grid][128][KTAPS][128] */, double* __restrict__ bias_partial /* [grid][128] */) {
    constexpr int TCOLS = KTAPS * 128 <= 128 ? 128 : (KTAPS * 128 <= 256 ? 256 : 512);
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ float bred[WGU_EXP][128];      // per-warp conv-bias partial sums
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_NSTAGE * WGU_STAGE_BYTES);
    // bars: [0..N) full (X tile by bulk copy), [N..2N) empty, [2N..3N) dyfull (expanders), 3N: done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * WG_NSTAGE + 1);
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int st) { return bar0 + 8u * st; };
    auto EMPTY = [&](int st) { return bar0 + 8u * (WG_NSTAGE + st); };
    auto DYFULL = [&](int st) { return bar0 + 8u * (2 * WG_NSTAGE + st); };
    const uint32_t DONE = bar0 + 8u * (3 * WG_NSTAGE);

    const long ntiles = (rows_total + BN - 1) / BN;
    const long tbeg = ntiles * blockIdx.x / gridDim.x;          // balanced: every CTA owns >= 1 tile (grid <= ntiles)
    const long tend = ntiles * (blockIdx.x + 1) / gridDim.x;

    if (threadIdx.x == 0) {
        for (int st = 0; st < WG_NSTAGE; ++st) { mbar_init(FULL(st), 1); mbar_init(EMPTY(st), 1); mbar_init(DYFULL(st), WGU_EXP); }
        mbar_init(DONE, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TCOLS>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            Pipe pp;
            for (long tile = tbeg; tile < tend; ++tile) {
                mbar_wait(EMPTY(pp.stage), pp.phase ^ 1);
                mbar_expect_tx(FULL(pp.stage), B_STAGE_BYTES);
                const long r0 = tile * BN;
                uint8_t* x_dst = smem + pp.stage * WGU_STAGE_BYTES + PANELS * WGU_DY_PANEL;
#pragma unroll 4
                for (int q = 0; q < PANELS; ++q)
                    bulk_g2s(smem_u32(x_dst + q * B_PANEL_BYTES), xp + (long)q * x_rows + r0, B_PANEL_BYTES, FULL(pp.stage));
                pp.advance(WG_NSTAGE);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(fmt, fmt, 1, 1, 128, 128);
            Pipe pp;
            bool first = true;
            for (long tile = tbeg; tile < tend; ++tile) {
                mbar_wait(FULL(pp.stage), pp.phase);
                mbar_wait(DYFULL(pp.stage), pp.phase);
                tc_fence_after();
                const uint32_t dy0 = smem_u32(smem + pp.stage * WGU_STAGE_BYTES);
                const uint32_t x0 = dy0 + PANELS * WGU_DY_PANEL;
#pragma unroll
                for (int kk = 0; kk < BN / 16; ++kk) {  // 16 rows (MMA K) per instruction
                    const uint64_t ad = make_desc(dy0 + (uint32_t)(kk * 16 * ROWB), 128, WGU_DY_PANEL);
#pragma unroll
                    for (int j = 0; j < KTAPS; ++j) {
                        const uint64_t bd = make_desc(x0 + (uint32_t)((kk * 16 + j) * ROWB), 128, B_PANEL_BYTES);
                        umma_f16(tmem_base + (uint32_t)(j * 128), ad, bd, idesc, !(first && kk == 0));
                    }
                }
                first = false;
                umma_commit(EMPTY(pp.stage));
                pp.advance(WG_NSTAGE);
            }
            umma_commit(DONE);
        }
        __syncwarp();
    } else {
        // ===== expanders: pooled fp32 inputs -> the 16-bit dY operand tile of every stage =====
        // lane <-> (pooling window, 8-channel panel): q = lane % 16 is the panel, lanes 0-15 / 16-31 take two consecutive windows,
        // warp e owns windows 4e .. 4e+3 (two chunks per lane and tile).  A warp load instruction therefore covers two
        // contiguous 512-byte pooled rows (8 cache lines; the previous lane = window mapping touched 32 lines per
        // instruction and ncu showed l1tex 90 % busy with the tensor pipe at 43 %), every lane owns a complete 16-byte
        // operand chunk (no lane exchange), and the per-channel BatchNorm constants of its fixed 8 channels live in registers.
        // The dY panels are WGU_DY_PANEL = 2048 + 16 bytes apart so that the 16 panels of one row hit 16 different bank groups.
        const int e = warp - 2;
        const int q = lane & 15, hw = lane >> 4;
        const float gs = src.gscale ? src.gscale[0] : 1.f;
        float ka[8], kb[8], kc[8], acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = q * 8 + i;
            const float sc = src.scale ? src.scale[c] : 1.f;
            float b0 = 0.f, c0 = 0.f;
            if (src.sums) {
                const double inv_n = 1.0 / src.count;
                const double rs2 = (double)src.rstd[c] * src.sums[128 + c] * inv_n;
                b0 = (float)(-(double)sc * rs2);
                c0 = (float)((double)sc * (rs2 * (double)src.mean[c] - src.sums[c] * inv_n));
            }
            ka[i] = sc; kb[i] = b0; kc[i] = c0; acc[i] = 0.f;
        }
        const float invP = 1.f / (float)src.P;
        // raw inputs of this lane's two chunks, TWO tiles ahead (two register sets keep two tiles of loads in flight)
        struct Raw {
            float4 g4[2][2], z4[2][2];
            uint2 cd[2];
            bool valid[2];
            long srow[2];
        };
        auto fetch = [&](Raw& w, long tile) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const unsigned win = (unsigned)(e * 4 + 2 * i + hw);
                const unsigned r = (unsigned)(tile * BN) + 4u * win;
                const unsigned sp = r / (unsigned)src.Lp;
                const unsigned pw = (r - sp * (unsigned)src.Lp) >> 2;
                w.valid[i] = tile < tend && (int)sp < src.S && (int)pw < src.P;
                w.srow[i] = (long)sp;
                if (w.valid[i]) {
                    const long crow = (long)sp * src.P + pw;
                    const float4* gp = reinterpret_cast<const float4*>(src.dy + crow * src.lddy + q * 8);
                    const float4* zp = reinterpret_cast<const float4*>(src.z + crow * 128 + q * 8);
                    w.g4[i][0] = __ldg(gp); w.g4[i][1] = __ldg(gp + 1);
                    w.z4[i][0] = __ldg(zp); w.z4[i][1] = __ldg(zp + 1);
                    w.cd[i] = __ldg(reinterpret_cast<const uint2*>(src.code + crow * 128 + q * 8));
                }
            }
        };
        Pipe pp;
        auto expand = [&](Raw& w, long tile) {
            unsigned hp[2][4], cw[2][2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                hp[i][0] = hp[i][1] = hp[i][2] = hp[i][3] = 0u;
                cw[i][0] = cw[i][1] = 0u;
                if (w.valid[i]) {
                    float g[8] = {w.g4[i][0].x, w.g4[i][0].y, w.g4[i][0].z, w.g4[i][0].w, w.g4[i][1].x, w.g4[i][1].y, w.g4[i][1].z, w.g4[i][1].w};
                    const float zz[8] = {w.z4[i][0].x, w.z4[i][0].y, w.z4[i][0].z, w.z4[i][0].w,
                                         w.z4[i][1].x, w.z4[i][1].y, w.z4[i][1].z, w.z4[i][1].w};
                    if (src.dtp) {
                        const float4* tp = reinterpret_cast<const float4*>(src.dtp + w.srow[i] * src.lddtp + q * 8);
                        const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1);
                        g[0] = fmaf(t0.x, invP, g[0]); g[1] = fmaf(t0.y, invP, g[1]); g[2] = fmaf(t0.z, invP, g[2]); g[3] = fmaf(t0.w, invP, g[3]);
                        g[4] = fmaf(t1.x, invP, g[4]); g[5] = fmaf(t1.y, invP, g[5]); g[6] = fmaf(t1.z, invP, g[6]); g[7] = fmaf(t1.w, invP, g[7]);
                    }
                    unsigned short h[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const float gg = fmaf(ka[t], g[t], fmaf(kb[t], zz[t], kc[t]));
                        const float v = zz[t] > 0.f ? gg : 0.f;
                        acc[t] += v;
                        h[t] = cvt_f32_to16(v * gs, fmt);
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) hp[i][c] = (unsigned)h[2 * c] | ((unsigned)h[2 * c + 1] << 16);
                    cw[i][0] = w.cd[i].x; cw[i][1] = w.cd[i].y;
                }
            }
            fetch(w, tile + 2);                       // refill this register set: two tiles of loads stay in flight
            mbar_wait(EMPTY(pp.stage), pp.phase ^ 1); // the MMAs that read this stage (3 tiles ago) have retired
            uint8_t* dy_dst = smem + pp.stage * WGU_STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                uint8_t* base = dy_dst + q * WGU_DY_PANEL + (e * 4 + 2 * i + hw) * (4 * ROWB);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // the four code bytes of a word become two 16-bit-lane masks with word-wide logic (exact zero-byte test +
                    // byte permutes) instead of a compare and select per channel
                    unsigned out[4];
#pragma unroll
                    for (int wq = 0; wq < 2; ++wq) {
                        const unsigned t = cw[i][wq] ^ ((unsigned)j * 0x01010101u);            // byte == 0 where code == j
                        const unsigned nz = ((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t;             // bit 7 set where the byte is non-zero
                        const unsigned m8 = ((~nz & 0x80808080u) >> 7) * 0xffu;                // 0xff in the matching bytes
                        out[2 * wq] = hp[i][2 * wq] & __byte_perm(m8, 0u, 0x1100);             // channels 0,1 of the word
                        out[2 * wq + 1] = hp[i][2 * wq + 1] & __byte_perm(m8, 0u, 0x3322);     // channels 2,3
                    }
                    *reinterpret_cast<uint4*>(base + j * ROWB) = make_uint4(out[0], out[1], out[2], out[3]);
                }
            }
            fence_proxy_async();                      // generic-proxy stores -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(DYFULL(pp.stage));
            pp.advance(WG_NSTAGE);
        };
        Raw ra, rb;
        fetch(ra, tbeg);
        fetch(rb, tbeg + 1);
        for (long tile = tbeg; tile < tend; tile += 2) {
            expand(ra, tile);
            if (tile + 1 < tend) expand(rb, tile + 1);
        }
        if (bias_partial) {
            // this lane's 8 channels: add the other window half (lane ^ 16), then the 8 warps through shared memory, fixed order
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
            if (hw == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) bred[e][q * 8 + i] = acc[i];
            }
            asm volatile("bar.sync 2, %0;" ::"n"(WGU_EXP * 32) : "memory");
            const int c = threadIdx.x - 64;
            if (c < 128) {
                double t = 0.0;
#pragma unroll
                for (int w8 = 0; w8 < WGU_EXP; ++w8) t += (double)bred[w8][c];
                bias_partial[(long)blockIdx.x * 128 + c] = t;
            }
        }
        if (e < 4) {
            // ===== epilogue: the four warps covering the four TMEM lane quarters write the CTA's partial =====
            const int quarter = warp & 3;
            const int co = quarter * 32 + lane;
            mbar_wait(DONE, 0);
            tc_fence_after();
            float* dst = part + ((long)blockIdx.x * 128 + co) * KTAPS * 128;
#pragma unroll 1
            for (int j = 0; j < KTAPS; ++j) {
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(j * 128 + ch * 32), v);
                    if (tend <= tbeg) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = 0.f;
                    }
                    float4* d4 = reinterpret_cast<float4*>(dst + j * 128 + ch * 32);
#pragma unroll
                    for (int i = 0; i < 8; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TCOLS>(tmem_base);
    }
}

size_t rows_smem_bytes(int k, bool atmem) {
    if (atmem) return (size_t)6 * B_STAGE_BYTES + 8 * (2 * 6 + 5) + 16;
    return (size_t)k * PANELS * A_PANEL_BYTES + NSTAGE * B_STAGE_BYTES + 8 * (2 * NSTAGE + 5) + 16;
}
bool use_atmem() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DCUE_TC_ATMEM");
        v = (e && e[0] == '1') ? 1 : 0;  // default: weights in shared memory (SS); DCUE_TC_ATMEM=1 selects TMEM (TS)
    }
    return v != 0;
}
constexpr size_t WG_SMEM = (size_t)WG_NSTAGE * WG_STAGE_BYTES + 8 * (2 * WG_NSTAGE + 1) + 16;

int tc_grid(long rows_total) {
    const long ntiles = (rows_total + BN - 1) / BN;
    const int sms = dcue_num_sms();
    return (int)(ntiles < sms ? (ntiles > 0 ? ntiles : 1) : sms);
}

template <int EPI, int POOL>
int launch_rows_t(const void* panel, long panel_rows, int fmt_in, const void* w_packed, int fmt_w, const float* bias,
                  const ConvGeom& g, float* out, uint8_t* code, double* partial, const float* gscale,
                  const float* tap_bias, float* dummy, int grid, cudaStream_t st, BnBwdStats bs = BnBwdStats{}) {
    const bool atm = use_atmem();
    const size_t smem = rows_smem_bytes(g.k, atm);
    if (atm) {
        DCUE_CUDA(cudaFuncSetAttribute(tc_conv_rows_kernel<EPI, POOL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_conv_rows_kernel<EPI, POOL, true><<<grid, NTHREADS_ROWS, smem, st>>>((const uint4*)panel, panel_rows, fmt_in,
                                                                                (const uint4*)w_packed, fmt_w, bias, g, out,
                                                                                code, partial, gscale, tap_bias, dummy, bs);
    } else {
        DCUE_CUDA(cudaFuncSetAttribute(tc_conv_rows_kernel<EPI, POOL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_conv_rows_kernel<EPI, POOL, false><<<grid, NTHREADS_ROWS, smem, st>>>((const uint4*)panel, panel_rows, fmt_in,
                                                                                 (const uint4*)w_packed, fmt_w, bias, g, out,
                                                                                 code, partial, gscale, tap_bias, dummy, bs);
    }
    DCUE_LAUNCH_CHECK();
    return 0;
}

}  // namespace

__global__ void dcue_cvt_d2f_kernel(const double* __restrict__ in, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}

size_t dcue_tc_ws_bytes(int k) {
    // [grid*4][2][128] stat partials + one scratch line for dead-lane stores
    const size_t stats = (size_t)dcue_num_sms() * 4 * (2 * 128 + 1) * sizeof(double) + 256;
    const size_t wg = (size_t)dcue_num_sms() * 128 * k * 128 * sizeof(float);
    return stats > wg ? stats : wg;
}

int dcue_tc_conv_fwd(const void* panel, long panel_rows, int fmt, const void* w_packed, const float* bias,
                     const float* tap_bias, const ConvGeom& g, float* z, uint8_t* code, double* sums, void* ws, size_t ws_bytes,
                     cudaStream_t st) {
    if (g.Cin != 128) DCUE_FAIL(DCUE_E_UNSUPPORTED, "tcgen05 conv needs Cin == 128 (got %d)", g.Cin);
    if (!code) DCUE_FAIL(DCUE_E_BADARG, "tcgen05 conv forward needs the argmax code buffer");
    const int grid = tc_grid(g.rows_total);
    const size_t part_bytes = (size_t)grid * 4 * 2 * g.Cout * sizeof(double);
    if (!ws || ws_bytes < part_bytes + 256) DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_conv_pool_fwd(tc): workspace too small");
    double* part = sums ? (double*)ws : nullptr;
    float* dummy = (float*)((char*)ws + part_bytes);
    int e;
    if (g.pool == 4) e = launch_rows_t<0, 4>(panel, panel_rows, fmt, w_packed, fmt, bias, g, z, code, part, nullptr, tap_bias, dummy, grid, st);
    else if (g.pool == 2) e = launch_rows_t<0, 2>(panel, panel_rows, fmt, w_packed, fmt, bias, g, z, code, part, nullptr, tap_bias, dummy, grid, st);
    else e = launch_rows_t<0, 1>(panel, panel_rows, fmt, w_packed, fmt, bias, g, z, code, part, nullptr, tap_bias, dummy, grid, st);
    if (e) return e;
    if (sums && sums != DCUE_STATS_PARTIALS) {
        dcue_reduce_partials_d<<<ceil_div_i(2 * g.Cout, 8), 256, 0, st>>>((const double*)ws, grid, 2 * g.Cout, sums);
        DCUE_LAUNCH_CHECK();
    }
    return 0;
}

int dcue_tc_conv_fwd_nparts(long rows_total) { return tc_grid(rows_total); }

int dcue_tc_conv_dgrad(const void* dy_panel_shifted, long panel_rows, int fmt_dy, const void* w_packed, int fmt_w,
                       const ConvGeom& g, const float* gscale, float* dx, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (g.Cin != 128) DCUE_FAIL(DCUE_E_UNSUPPORTED, "tcgen05 dgrad needs Cout == 128 (got %d)", g.Cin);
    if (fmt_dy != fmt_w) DCUE_FAIL(DCUE_E_UNSUPPORTED, "tcgen05 kind::f16 needs both operands in the same 16-bit format");
    if (!ws || ws_bytes < 256) DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_conv_dgrad(tc): workspace too small");
    return launch_rows_t<1, 1>(dy_panel_shifted, panel_rows, fmt_dy, w_packed, fmt_w, nullptr, g, dx, nullptr, nullptr, gscale,
                               nullptr, (float*)ws, tc_grid(g.rows_total), st);
}

// dgrad + the BatchNorm-backward partial sums of the stage below, left at the start of ws for dcue_bn_bwd_finalize:
// [grid*4][2*Cout] doubles (sum dy, sum dy*xhat) followed by [grid*4] doubles (max|dy|)
int dcue_tc_conv_dgrad_stats(const void* dy_panel_shifted, long panel_rows, int fmt_dy, const void* w_packed, int fmt_w,
                             const ConvGeom& g, const float* gscale, float* dx, const float* z, const float* mean, const float* rstd,
                             const float* dtp, int lddtp, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (g.Cin != 128) DCUE_FAIL(DCUE_E_UNSUPPORTED, "tcgen05 dgrad needs Cout == 128 (got %d)", g.Cin);
    if (fmt_dy != fmt_w) DCUE_FAIL(DCUE_E_UNSUPPORTED, "tcgen05 kind::f16 needs both operands in the same 16-bit format");
    const int grid = tc_grid(g.rows_total);
    const size_t need = (size_t)grid * 4 * (2 * g.Cout + 1) * sizeof(double);
    if (!ws || ws_bytes < need) DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_conv_dgrad_stats(tc): workspace too small");
    BnBwdStats bs{z, mean, rstd, dtp, lddtp, 1.f / (float)g.Lin};
    return launch_rows_t<2, 1>(dy_panel_shifted, panel_rows, fmt_dy, w_packed, fmt_w, nullptr, g, dx, nullptr, (double*)ws, gscale,
                               nullptr, nullptr, grid, st, bs);
}

int dcue_tc_conv_wgrad(const void* dy_panel, long dy_rows, int fmt_dy, const void* x_panel, long x_rows, int fmt_x,
                       long rows_total, int k, int Cin, int Cout, const float* gscale, float* dW, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
    if (fmt_dy != fmt_x) DCUE_FAIL(DCUE_E_UNSUPPORTED, "tcgen05 kind::f16 needs both operands in the same 16-bit format");
    if (Cin != 128 || Cout != 128) DCUE_FAIL(DCUE_E_UNSUPPORTED, "tcgen05 wgrad needs Cin == Cout == 128");
    const int grid = tc_grid(rows_total);
    if (!ws || ws_bytes < (size_t)grid * 128 * k * 128 * sizeof(float))
        DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_conv_wgrad(tc): workspace too small");
#define LAUNCH_WG(KT)                                                                                             \
    do {                                                                                                          \
        DCUE_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM)); \
        tc_wgrad_kernel<KT><<<grid, NTHREADS, WG_SMEM, st>>>((const uint4*)dy_panel, dy_rows, fmt_dy, (const uint4*)x_panel, \
                                                             x_rows, fmt_x, rows_total, (float*)ws);              \
    } while (0)
    switch (k) {
        case 1: LAUNCH_WG(1); break;
        case 2: LAUNCH_WG(2); break;
        case 3: LAUNCH_WG(3); break;
        case 4: LAUNCH_WG(4); break;
        default: DCUE_FAIL(DCUE_E_BADARG, "k must be 1..4");
    }
#undef LAUNCH_WG
    DCUE_LAUNCH_CHECK();
    dcue_wgrad_reduce_kernel<<<ceil_div_i((long)Cout * Cin * k, 256), 256, 0, st>>>((const float*)ws, grid, Cout, Cin, k, gscale, dW);
    DCUE_LAUNCH_CHECK();
    return 0;
}

size_t dcue_tc_wgrad_unpool_ws_bytes(int k) {
    return (size_t)dcue_num_sms() * 128 * k * 128 * sizeof(float) + (size_t)dcue_num_sms() * 128 * sizeof(double) + 256;
}

int dcue_tc_conv_wgrad_unpool(const float* dy, int lddy, const float* dtp, int lddtp, const float* z, const uint8_t* code,
                              const float* scale, const float* mean, const float* rstd, const double* sums, double count, int S,
                              int P, int Lp, const void* x_panel, long x_rows, int fmt, int k, const float* gscale, float* dW,
                              double* bias_sums, float* bias_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (k != 4) DCUE_FAIL(DCUE_E_UNSUPPORTED, "fused unpool+wgrad is built for k == 4 (got %d)", k);
    const long rows_total = (long)S * Lp;
    const int grid = tc_grid(rows_total);
    const size_t part_bytes = (size_t)grid * 128 * k * 128 * sizeof(float);
    if (!ws || ws_bytes < part_bytes + (size_t)grid * 128 * sizeof(double))
        DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_conv_wgrad_unpool: workspace too small");
    double* bpart = bias_sums ? (double*)((char*)ws + part_bytes) : nullptr;
    UnpoolSrc src{dy, lddy, dtp, lddtp, z, code, scale, mean, rstd, sums, count > 0 ? count : 1.0, S, P, Lp, gscale};
    constexpr size_t SMEM = (size_t)WG_NSTAGE * WGU_STAGE_BYTES + 8 * (3 * WG_NSTAGE + 1) + 16;
    DCUE_CUDA(cudaFuncSetAttribute(tc_wgrad_unpool_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    tc_wgrad_unpool_kernel<4><<<grid, WGU_THREADS, SMEM, st>>>(src, fmt, (const uint4*)x_panel, x_rows, rows_total, (float*)ws, bpart);
    DCUE_LAUNCH_CHECK();
    dcue_wgrad_reduce_kernel<<<ceil_div_i((long)128 * 128 * k, 256), 256, 0, st>>>((const float*)ws, grid, 128, 128, k, gscale, dW);
    DCUE_LAUNCH_CHECK();
    if (bias_sums) {
        dcue_reduce_partials_d<<<ceil_div_i(128, 8), 256, 0, st>>>(bpart, grid, 128, bias_sums);
        DCUE_LAUNCH_CHECK();
        if (bias_out) {
            dcue_cvt_d2f_kernel<<<1, 128, 0, st>>>(bias_sums, 128, bias_out);
            DCUE_LAUNCH_CHECK();
        }
    }
    return 0;
}
