// Eval scorer: all-pairs cosine score GEMM (tcgen05, fp32 accumulate in TMEM) with a fused
// streaming top-k kept in shared memory — the score matrix is never materialised.
//
// Generalises DCUE.predict (dcrecommend/nn/dcue.py:495-513: model.sim of one user's factor row
// against every candidate song's factor row) to all users x all songs.
//
//   users on the MMA M axis (one TMEM lane = one user), songs on N.  CTA = 128 users x one
//   contiguous range of songs; the user tile lives in TMEM (TS-mode MMA), song tiles stream through a
//   3-stage cp.async.bulk ring; accumulators are double buffered so the filter overlaps the next MMA.
//   Filter: one TMEM lane = one user, so each epilogue thread owns one user: its threshold (the k-th
//   best score at the last compaction) and its list length live in registers.  A chunk of 32 fresh scores
//   is first reduced with a max tree; only if some lane beats its threshold are the columns scanned,
//   and a passing score is APPENDED to that user's unsorted list in shared memory by its own lane (plain
//   predicated stores, all 32 users in parallel).  When a list reaches CAP entries the warp sorts it
//   cooperatively (bitonic network, 4 entries per lane, 15 shuffle steps), keeps the k best and raises the
//   threshold: one ~400-instruction compaction per CAP-k appends instead of a dependent
//   shuffle/LDS/STS chain per insert (ncu of the first version: 3 % tensor pipe, 780 cycles per insert).
//   Finish: one last sort per user, write the k best (descending, ties by lower song index).
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int TU = 128;        // users per CTA (MMA M)
constexpr int TI = 128;        // songs per tile (MMA N)
constexpr int SLOTS = 128;     // kept candidates per user (k <= SLOTS)
constexpr int ROWB = 16;
constexpr int PANEL_BYTES = 128 * ROWB;  // 2048: one 8-wide K chunk of a 128-row tile
constexpr int NSTAGE = 3;
constexpr int NTHREADS = 192;
constexpr int ACC0 = 64;         // accumulator columns start (the user tile occupies TMEM columns [0, Kp/2 <= 64))
constexpr int NACC = 3;          // accumulator stages: absorbs the jitter of rare compactions / slow scans

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
// asynchronous TMEM load: the registers are valid only after tmem_ld_fence() on the same array
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
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
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// ties the registers to a point after tmem_wait_ld() so no consumer is scheduled before the wait
__device__ __forceinline__ void tmem_ld_fence(uint32_t (&r)[32]) {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                      "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
    asm volatile("" : "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                      "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}
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
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int a_fmt, int b_fmt, int M, int N) {
    return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

constexpr int CAP = SLOTS;      // list capacity per user
constexpr int LD = SLOTS + 1;   // row stride (floats): (user + slot) % 32 banks -> lane-parallel appends do not collide

struct Cand {
    float s;
    int i;
};
// a ranks before b: higher score, ties by lower song index (-1 = empty sorts last among equals)
__device__ __forceinline__ bool before(const Cand& a, const Cand& b) {
    return a.s > b.s || (a.s == b.s && (unsigned)a.i < (unsigned)b.i);
}

// Warp-cooperative bitonic sort (descending) of 128 candidates, element g = lane*4 + e.
__device__ __forceinline__ void bitonic_sort128(Cand (&c)[4], int lane) {
#pragma unroll
    for (int size = 2; size <= 128; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 4) {
                const int lx = stride >> 2;
                const bool lower = (lane & lx) == 0;                    // g < partner
                const bool desc = ((lane * 4) & size) == 0;             // size >= 8: uniform over e
                const bool keep_first = lower == desc;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    Cand o;
                    o.s = __shfl_xor_sync(0xffffffffu, c[e].s, lx);
                    o.i = __shfl_xor_sync(0xffffffffu, c[e].i, lx);
                    const bool mine_first = before(c[e], o);
                    if (mine_first != keep_first) c[e] = o;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int pe = e ^ stride;
                    if (pe > e) {
                        const bool desc = ((lane * 4 + e) & size) == 0;
                        const bool in_order = before(c[e], c[pe]);
                        if (in_order != desc) { const Cand t = c[e]; c[e] = c[pe]; c[pe] = t; }
                    }
                }
            }
        }
    }
}

// Sort user `u`'s list (entries >= n are empty), write it back sorted; returns the k-th best score.
__device__ __forceinline__ float compact_user(float* __restrict__ cs, int* __restrict__ ci, int u, int n, int k, int lane) {
    Cand c[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int g = lane * 4 + e;
        c[e].s = g < n ? cs[u * LD + g] : -INFINITY;
        c[e].i = g < n ? ci[u * LD + g] : -1;
    }
    bitonic_sort128(c, lane);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        cs[u * LD + lane * 4 + e] = c[e].s;
        ci[u * LD + lane * 4 + e] = c[e].i;
    }
    // k-th best = element k-1 = lane (k-1)/4, e = (k-1)%4
    const int ke = (k - 1) & 3;
    const float mine = ke == 0 ? c[0].s : ke == 1 ? c[1].s : ke == 2 ? c[2].s : c[3].s;
    const float kth = __shfl_sync(0xffffffffu, mine, (k - 1) >> 2);
    __syncwarp();
    return kth;
}

struct UserState {   // per-lane (= per-user) filter state, kept in registers
    float thr;       // k-th best score at the last compaction
    int cnt;         // entries in this user's list
};

// Rare path, ONE non-inlined copy (keeps the hot loop small enough for the instruction cache): re-read a
// 32-column chunk of the accumulator from TMEM, append every score that beats its user's threshold (each
// lane appends to its own user's list), and compact any list that fills up.
__device__ __noinline__ UserState scan_chunk(uint32_t taddr, long ib, long iend, UserState st, bool enable,
                                             float* __restrict__ cs, int* __restrict__ ci, int ubase, int k, int lane) {
    float v[32];
    tmem_ld32(taddr, v);
    const int u = ubase + lane;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (enable && v[j] > st.thr && ib + j < iend) {
            cs[u * LD + st.cnt] = v[j];
            ci[u * LD + st.cnt] = (int)(ib + j);
            ++st.cnt;
        }
        unsigned full = __ballot_sync(0xffffffffu, st.cnt == CAP);
        while (full) {
            const int l = __ffs(full) - 1;
            full &= full - 1;
            const float kth = compact_user(cs, ci, ubase + l, CAP, k, lane);
            if (lane == l) { st.thr = kth; st.cnt = k; }
        }
    }
    return st;
}

// compaction of every user whose list is full (state by value: keeps thr/cnt in registers)
__device__ __noinline__ UserState compact_full(unsigned full, UserState st, float* __restrict__ cs, int* __restrict__ ci,
                                               int ubase, int k, int lane) {
    while (full) {
        const int l = __ffs(full) - 1;
        full &= full - 1;
        const float kth = compact_user(cs, ci, ubase + l, CAP, k, lane);
        if (lane == l) { st.thr = kth; st.cnt = k; }
    }
    return st;
}

__global__ void __launch_bounds__(NTHREADS, 1)
topk_kernel(const uint4* __restrict__ users, long u_rows /* panel rows */, long n_users, const uint4* __restrict__ items,
            long i_rows, long n_items, int Kp, int fmt, int k, long item_offset, long items_per_split, int dbg,
            float* __restrict__ out_s, int64_t* __restrict__ out_i /* [splits][n_users][k] */) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npan = Kp / 8;
    const int tile_bytes = npan * PANEL_BYTES;
    uint8_t* sB = smem;
    float* cs = reinterpret_cast<float*>(sB + NSTAGE * tile_bytes);
    int* ci = reinterpret_cast<int*>(cs + TU * LD);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uintptr_t>(ci + TU * LD + 1) & ~(uintptr_t)7);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 1 + 2 * NACC);
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
    const uint32_t AFULL = bar0 + 8u * (2 * NSTAGE);
    auto TFULL = [&](int a) { return bar0 + 8u * (2 * NSTAGE + 1 + a); };
    auto TEMPTY = [&](int a) { return bar0 + 8u * (2 * NSTAGE + 1 + NACC + a); };

    const long u0 = (long)blockIdx.x * TU;
    const long ibeg = (long)blockIdx.y * items_per_split;
    const long iend = ibeg + items_per_split < n_items ? ibeg + items_per_split : n_items;
    const long ntiles = iend > ibeg ? (iend - ibeg + TI - 1) / TI : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        mbar_init(AFULL, 4);
        for (int a = 0; a < NACC; ++a) { mbar_init(TFULL(a), 1); mbar_init(TEMPTY(a), 4); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long t = 0; t < ntiles; ++t) {
                mbar_wait(EMPTY(stage), phase ^ 1);
                mbar_expect_tx(FULL(stage), (uint32_t)tile_bytes);
                const long r0 = ibeg + t * TI;   // multiple of 128: tile index r0 >> 7
                bulk_g2s(smem_u32(sB + stage * tile_bytes), items + (r0 >> 7) * (long)npan * 128, (uint32_t)tile_bytes, FULL(stage));
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(fmt, fmt, TU, TI);
            mbar_wait(AFULL, 0);   // the four epilogue warps have copied the user tile into TMEM
            tc_fence_after();
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (long t = 0; t < ntiles; ++t) {
                mbar_wait(TEMPTY(acc), acc_phase ^ 1);
                mbar_wait(FULL(stage), phase);
                tc_fence_after();
                const uint32_t b0 = smem_u32(sB + stage * tile_bytes);
                for (int c = 0; c < Kp / 16; ++c) {
                    const uint64_t bd = make_desc(b0 + (uint32_t)(2 * c * PANEL_BYTES), PANEL_BYTES, 128);
                    umma_f16_ts(tmem_base + (uint32_t)(ACC0 + acc * TI), tmem_base + (uint32_t)(c * 8), bd, idesc, c != 0);
                }
                umma_commit(EMPTY(stage));
                umma_commit(TFULL(acc));
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        const int quarter = warp & 3;
        const int ubase = quarter * 32;
        const int u = ubase + lane;          // this lane's user within the tile == TMEM lane
        {   // user tile -> TMEM (TS-mode A operand): row m, K chunk kk8 = one uint4 = columns 4*kk8..4*kk8+3
            for (int c0 = 0; c0 < 16; c0 += 8) {
                uint32_t r[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint4 w = make_uint4(0u, 0u, 0u, 0u);
                    if (c0 + i < npan) w = __ldg(users + ((u0 >> 7) * (long)npan + (c0 + i)) * 128 + u);
                    r[4 * i] = w.x; r[4 * i + 1] = w.y; r[4 * i + 2] = w.z; r[4 * i + 3] = w.w;
                }
                tmem_st32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c0 * 4), r);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(AFULL);
        }
        UserState st;
        st.thr = -INFINITY;
        st.cnt = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (long t = 0; t < ntiles; ++t) {
            mbar_wait(TFULL(acc), acc_phase);
            tc_fence_after();
            const long it0 = ibeg + t * TI;
            // hot path: all four 32-column loads in flight, one wait, a max tree per chunk
            uint32_t r0[32], r1[32], r2[32], r3[32];
            const uint32_t tb = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ACC0 + acc * TI);
            if (dbg >= 3) {  // timing experiment: no TMEM reads at all
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(TEMPTY(acc));
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
                continue;
            }
            tmem_ld32_async(tb, r0);
            tmem_ld32_async(tb + 32, r1);
            tmem_ld32_async(tb + 64, r2);
            tmem_ld32_async(tb + 96, r3);
            tmem_wait_ld();
            tmem_ld_fence(r0); tmem_ld_fence(r1); tmem_ld_fence(r2); tmem_ld_fence(r3);
            float mx[4];
#define DCUE_CHUNK_MAX(R, C)                                                                              \
            {                                                                                             \
                float m0 = __uint_as_float(R[0]), m1 = __uint_as_float(R[1]), m2 = __uint_as_float(R[2]),  \
                      m3 = __uint_as_float(R[3]);                                                         \
                _Pragma("unroll") for (int j = 4; j < 32; j += 4) {                                         \
                    m0 = fmaxf(m0, __uint_as_float(R[j])); m1 = fmaxf(m1, __uint_as_float(R[j + 1]));     \
                    m2 = fmaxf(m2, __uint_as_float(R[j + 2])); m3 = fmaxf(m3, __uint_as_float(R[j + 3])); \
                }                                                                                         \
                mx[C] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));                                              \
            }
            DCUE_CHUNK_MAX(r0, 0)
            DCUE_CHUNK_MAX(r1, 1)
            DCUE_CHUNK_MAX(r2, 2)
            DCUE_CHUNK_MAX(r3, 3)
#undef DCUE_CHUNK_MAX
            if (it0 + TI > iend) {
                // last tile: the slow path masks the columns beyond the song range
                for (int c = 0; c < 4; ++c)
                    st = scan_chunk(tb + (uint32_t)(c * 32), it0 + c * 32, iend, st, true, cs, ci, ubase, k, lane);
            } else if (dbg < 1) {
                // A chunk is examined only if some user of this warp beat its threshold there.  Each lane then
                // builds the mask of its passing columns; with exactly one (the common case) the value is the
                // chunk maximum it already holds, so it appends directly -- all 32 users in parallel.  Lanes with
                // several candidates (warm-up) take the slow path that re-reads the chunk from TMEM.
#define DCUE_CHUNK_SCAN(R, C)                                                                                   \
                if (__any_sync(0xffffffffu, mx[C] > st.thr)) {                                                  \
                    unsigned bits = 0;                                                                          \
                    _Pragma("unroll") for (int j = 0; j < 32; ++j)                                                \
                        bits |= (__uint_as_float(R[j]) > st.thr) ? (1u << j) : 0u;                              \
                    const int pc = __popc(bits);                                                                \
                    if (pc == 1) {                                                                              \
                        cs[u * LD + st.cnt] = mx[C];                                                            \
                        ci[u * LD + st.cnt] = (int)(it0 + (C) * 32 + __ffs(bits) - 1);                          \
                        ++st.cnt;                                                                               \
                    }                                                                                           \
                    const unsigned full = __ballot_sync(0xffffffffu, st.cnt == CAP);                            \
                    if (full) st = compact_full(full, st, cs, ci, ubase, k, lane);                              \
                    if (__any_sync(0xffffffffu, pc > 1))                                                        \
                        st = scan_chunk(tb + (uint32_t)((C) * 32), it0 + (C) * 32, iend, st, pc > 1, cs, ci, ubase, k, lane); \
                }
                DCUE_CHUNK_SCAN(r0, 0)
                DCUE_CHUNK_SCAN(r1, 1)
                DCUE_CHUNK_SCAN(r2, 2)
                DCUE_CHUNK_SCAN(r3, 3)
#undef DCUE_CHUNK_SCAN
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(TEMPTY(acc));
            if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
        }
        // ---- finish: final sort of every user's list; lane L writes slots 4L..4L+3
        __syncwarp();
        for (int uu = 0; uu < 32; ++uu) {
            const int usr = ubase + uu;
            const long gu = u0 + usr;
            if (gu >= n_users) break;
            const int n = __shfl_sync(0xffffffffu, st.cnt, uu);
            compact_user(cs, ci, usr, n, k, lane);
            const long ob = ((long)blockIdx.y * n_users + gu) * k;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int slot = lane * 4 + e;
                if (slot < k) {
                    const int it = ci[usr * LD + slot];
                    out_s[ob + slot] = cs[usr * LD + slot];
                    out_i[ob + slot] = it < 0 ? -1 : (int64_t)it + item_offset;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// x[rows,F] fp32 -> x/max(|x|,eps) as 16-bit K-major operand tiles: [rows/128 tiles][Kp/8 panels][128 rows][8]
// (tile-major, so one 128-row operand tile is ONE contiguous cp.async.bulk of Kp*256 bytes)
__global__ void __launch_bounds__(256)
normalize_rows_kernel(const float* __restrict__ x, long rows, int F, float eps, int Kp, int fmt, uint4* __restrict__ out,
                      long panel_rows) {
    const int lane = threadIdx.x & 31;
    const long r = blockIdx.x * 8L + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float* xr = x + r * F;
    float ss = 0.f;
    for (int c = lane; c < F; c += 32) { const float v = xr[c]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), eps);
    for (int q = lane; q < Kp / 8; q += 32) {
        unsigned short h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = q * 8 + j;
            h[j] = cvt_f32_to16(c < F ? xr[c] * inv : 0.f, fmt);
        }
        uint4 o;
        o.x = h[0] | ((unsigned)h[1] << 16);
        o.y = h[2] | ((unsigned)h[3] << 16);
        o.z = h[4] | ((unsigned)h[5] << 16);
        o.w = h[6] | ((unsigned)h[7] << 16);
        out[((r >> 7) * (Kp / 8) + q) * 128 + (r & 127)] = o;
    }
}

// k-way merge of `parts` descending lists per user, one thread per user
__global__ void topk_merge_kernel(const float* __restrict__ s, const int64_t* __restrict__ idx, int parts, long n_users,
                                  int k, float* __restrict__ os, int64_t* __restrict__ oi) {
    const long u = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (u >= n_users) return;
    int head[16];
    for (int p = 0; p < parts; ++p) head[p] = 0;
    for (int r = 0; r < k; ++r) {
        int bp = -1;
        float bs = -INFINITY;
        int64_t bi = -1;
        for (int p = 0; p < parts; ++p) {
            if (head[p] >= k) continue;
            const long o = ((long)p * n_users + u) * k + head[p];
            const float v = s[o];
            const int64_t ii = idx[o];
            if (ii < 0) continue;
            if (bp < 0 || v > bs || (v == bs && ii < bi)) { bp = p; bs = v; bi = ii; }
        }
        os[u * k + r] = bp < 0 ? -INFINITY : bs;
        oi[u * k + r] = bi;
        if (bp >= 0) ++head[bp];
    }
}

int topk_splits(long n_users, long n_items) {
    const long ublocks = (n_users + TU - 1) / TU;
    long want = (2L * dcue_num_sms() + ublocks - 1) / ublocks;
    const long maxs = (n_items + 4 * TI - 1) / (4 * TI);
    if (want > maxs) want = maxs;
    if (want > 16) want = 16;
    if (want < 1) want = 1;
    return (int)want;
}

}  // namespace

extern "C" int dcue_normalize_rows(const float* x, long rows, int F, float eps, int Kp, int fmt, void* out, void* stream) {
    DCUE_CHECK_ARG(x && out && rows >= 0 && F > 0 && Kp >= F && Kp % 16 == 0 && Kp <= 256);
    if (rows == 0) return 0;
    normalize_rows_kernel<<<ceil_div_i(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, rows, F, eps, Kp, fmt, (uint4*)out,
                                                                                 round_up_l(rows, 128));
    DCUE_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t dcue_topk_ws_bytes(int impl, long n_users, long n_items, int k) {
    (void)impl;
    const int splits = topk_splits(n_users, n_items);
    return splits > 1 ? (size_t)splits * n_users * k * (sizeof(float) + sizeof(int64_t)) + 256 : 256;
}

extern "C" int dcue_topk_scores(int impl, const void* users_n, long n_users, const void* items_n, long n_items, int Kp,
                                int fmt, int k, long item_offset, float* top_scores, int64_t* top_idx, void* ws,
                                size_t ws_bytes, void* stream) {
    DCUE_CHECK_ARG(users_n && items_n && top_scores && top_idx && n_users >= 0 && n_items >= 0 && k > 0 &&
                   k <= SLOTS - 8);  // a compaction must free at least 8 slots
    DCUE_CHECK_ARG(Kp % 16 == 0 && Kp >= 16 && Kp <= 128);
    if (impl != DCUE_IMPL_TC) DCUE_FAIL(DCUE_E_UNSUPPORTED, "dcue_topk_scores: only the tcgen05 implementation exists");
    if (n_users == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int splits = topk_splits(n_users, n_items);
    long per = (n_items + splits - 1) / splits;
    per = round_up_l(per > 0 ? per : 1, TI);
    float* os = top_scores;
    int64_t* oi = top_idx;
    if (splits > 1) {
        const size_t need = (size_t)splits * n_users * k * (sizeof(float) + sizeof(int64_t));
        if (!ws || ws_bytes < need) DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_topk_scores: workspace %zu < %zu", ws_bytes, need);
        oi = (int64_t*)ws;
        os = (float*)((char*)ws + (size_t)splits * n_users * k * sizeof(int64_t));
    }
    // slots that are never filled (fewer than k songs in a split) must read as "missing"
    DCUE_CUDA(cudaMemsetAsync(oi, 0xff, (size_t)splits * n_users * k * sizeof(int64_t), st));
    const size_t smem = (size_t)NSTAGE * (Kp / 8) * PANEL_BYTES + (size_t)TU * (SLOTS + 1) * 8 + 16 +
                        8 * (2 * NSTAGE + 1 + 2 * NACC) + 16;
    DCUE_CUDA(cudaFuncSetAttribute(topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((n_users + TU - 1) / TU), splits);
    topk_kernel<<<grid, NTHREADS, smem, st>>>((const uint4*)users_n, round_up_l(n_users, 128), n_users,
                                              (const uint4*)items_n, round_up_l(n_items, 128), n_items, Kp, fmt, k,
                                              item_offset, per, getenv("DCUE_TOPK_DEBUG") ? atoi(getenv("DCUE_TOPK_DEBUG")) : 0, os, oi);
    DCUE_LAUNCH_CHECK();
    if (splits > 1) {
        topk_merge_kernel<<<ceil_div_i(n_users, 128), 128, 0, st>>>(os, oi, splits, n_users, k, top_scores, top_idx);
        DCUE_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int dcue_topk_merge(const float* scores, const int64_t* idx, int parts, long n_users, int k, float* out_scores,
                               int64_t* out_idx, void* stream) {
    DCUE_CHECK_ARG(scores && idx && out_scores && out_idx && parts >= 1 && parts <= 16 && n_users >= 0 && k > 0);
    if (n_users == 0) return 0;
    topk_merge_kernel<<<ceil_div_i(n_users, 128), 128, 0, (cudaStream_t)stream>>>(scores, idx, parts, n_users, k, out_scores,
                                                                                  out_idx);
    DCUE_LAUNCH_CHECK();
    return 0;
}
