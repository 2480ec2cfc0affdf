// Eval scorer: all-pairs cosine score GEMM (tcgen05, fp32 accumulate in TMEM) with a fused
// streaming top-k -- the score matrix is never materialised.
//
// Generalises DCUE.predict (dcrecommend/nn/dcue.py:495-513: model.sim of one user's factor row
// against every candidate song's factor row) to all users x all songs (BASELINE cfg5).
// Design notes are at topk_stream_kernel below.
#include "common.cuh"
#include <stdlib.h>

// phase cycle counters of block 0 / appender warp 0 (diagnostics: dcue_topk_debug_cycles)
__device__ unsigned long long g_topk_dbg[8];

namespace {

constexpr int ROWB = 16;
constexpr int PANEL_BYTES = 128 * ROWB;  // 2048: one 8-wide K chunk of a 128-row operand tile

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
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_fence16(uint32_t (&r)[16]) {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                      "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int a_fmt, int b_fmt, int M, int N) {
    return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// ----------------------------------------------------------------------------- streaming top-k scorer
// Transposed formulation: SONGS on the MMA M axis (one TMEM lane = one song of the 128-song tile), 256 USERS on N
// (one accumulator column = one user).  An epilogue ("appender") thread therefore sees one song against many users
// and the per-user state is addressed by the (compile-time) column: the hot loop is one predicate-accumulating
// compare per score against the user's threshold in shared memory and one warp vote per 4 columns -- no per-lane
// lists, no shuffles, no divergence unless some song beats some user's threshold (a few 1e-4 of the scores).
//   * passing scores are APPENDED to the user's candidate list in an L2-resident global scratch (CAPH slots per
//     user, slot index from a shared-memory counter, warp-aggregated atomics);
//   * a list that has grown beyond LIMIT is FROZEN at a tile boundary ([0, n) no longer changes; appends continue
//     behind it) and queued; dedicated compactor warps pop the queue, radix-select the k best of the frozen part
//     in place (exact k-th largest key, 16 entries per lane, one REDUX per key bit) and publish the new threshold;
//     the owner warp installs it at a later tile boundary (surviving tail entries move down behind the k kept);
//     the appenders never wait for a compaction unless a list is about to overflow, and then they help;
//   * at the end of the song range every list is reduced to <= k entries, sorted (descending, ties by lower song
//     index) and written out.
// History (ncu, 500k songs, k=100): users-on-lanes version 1 600 instructions per warp and tile, 4 % tensor pipe;
// transposed with synchronous bitonic compaction: hot loop 21 % of the instructions, sorts 48 %, 28 % of the samples
// stalled at the epilogue barrier behind one sorting warp -> asynchronous compactors + selection instead of sorting.
constexpr int NU = 256;          // users per CTA (MMA N)
constexpr int TS = 128;          // songs per tile (MMA M)
constexpr int CAPH = 2048;       // candidate slots per user in the global scratch
constexpr int BSTEP = 4;         // tiles between two boundary checks of the appenders
constexpr int LIMIT = 384;       // freeze + compact a list longer than this (the frozen part is capped at 512 entries)
constexpr int HARD = CAPH - BSTEP * TS;  // a list longer than this could overflow before the next check: wait for its compaction
constexpr int NST = 4;           // song-tile stages
constexpr int NEPI = 16;         // appender warps: (TMEM lane quarter) x (column quarter)
constexpr int UPW = NU / NEPI;   // users owned (bookkeeping, final output) per appender warp
constexpr int CPW = 4 * NU / NEPI;   // user columns compared per appender warp
constexpr int NCOMP = 4;         // compactor warps
constexpr int NTHREADS2 = 64 + (NEPI + NCOMP) * 32;
constexpr int UPANEL = NU * ROWB;   // 4096: one 8-wide K chunk of the 256-user operand

struct Cand {
    float s;
    int i;
};
// a ranks before b: higher score, ties by lower song index (-1 = empty sorts last among equals)
__device__ __forceinline__ bool before(const Cand& a, const Cand& b) {
    return a.s > b.s || (a.s == b.s && (unsigned)a.i < (unsigned)b.i);
}

// order-preserving map float bits -> unsigned (larger float = larger key) and back
__device__ __forceinline__ unsigned f2key(int bits) { return (unsigned)bits ^ ((bits < 0) ? 0xffffffffu : 0x80000000u); }
__device__ __forceinline__ int key2f(unsigned key) { return (int)(key ^ ((key & 0x80000000u) ? 0x80000000u : 0xffffffffu)); }

// Keep the k best of lst[0, n) (k <= n <= 512) in lst[0, k), unordered; returns the k-th best score.  Warp
// cooperative, entry g = e*32 + lane (coalesced); exact radix select over the key bits that actually vary.
__device__ __noinline__ float select_topk_inplace(int2* __restrict__ lst, int n, int k, int lane) {
    unsigned key[16];
    int idx[16];
    unsigned vmask = 0, aand = 0xffffffffu, oor = 0u;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const int g = e * 32 + lane;
        key[e] = 0u;
        idx[e] = -1;
        if (g < n) {
            const int2 v = __ldcg(lst + g);      // appended by other warps of this CTA: read through L2
            key[e] = f2key(v.x);
            idx[e] = v.y;
            vmask |= 1u << e;
            aand &= key[e];
            oor |= key[e];
        }
    }
    aand = __reduce_and_sync(0xffffffffu, aand);
    oor = __reduce_or_sync(0xffffffffu, oor);
    const unsigned diff = aand ^ oor;            // key bits on which the entries differ
    unsigned prefix = aand;                      // bits common to all entries (varying bits are 0 here)
    unsigned decided = ~diff;                    // mask of the bits of `prefix` that are final
    int rem = k;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
        const unsigned b = 1u << bit;
        if (!(diff & b)) continue;
        // candidates agree with prefix on every decided bit (the common bits and the varying bits above) and have this bit set
        const unsigned m = decided | b;
        const unsigned want = prefix | b;
        int c = 0;
#pragma unroll
        for (int e = 0; e < 16; ++e) c += (((vmask >> e) & 1u) && ((key[e] & m) == want)) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= rem) prefix |= b; else rem -= c;
        decided |= b;
    }
    const unsigned T = prefix;                   // key of the k-th best; `rem` entries equal to T are still needed
    __syncwarp();                                // every lane has loaded its entries: the front of the list may be overwritten
    const unsigned lt = (1u << lane) - 1u;
    int base = 0;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const bool kp = ((vmask >> e) & 1u) && key[e] > T;
        const unsigned bm = __ballot_sync(0xffffffffu, kp);
        if (kp) __stcg(lst + base + __popc(bm & lt), make_int2(key2f(key[e]), idx[e]));
        base += __popc(bm);
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const bool eq = ((vmask >> e) & 1u) && key[e] == T;
        const unsigned bm = __ballot_sync(0xffffffffu, eq);
        const int pos = base + __popc(bm & lt);
        if (eq && pos < k) __stcg(lst + pos, make_int2(key2f(key[e]), idx[e]));
        base += __popc(bm);
    }
    __syncwarp();
    return __int_as_float(key2f(T));
}

// Warp-cooperative bitonic sort (descending) of 32*EPL candidates, element g = lane*EPL + e (final output only).
template <int EPL>
__device__ __forceinline__ void bitonic_sort_warp(Cand (&c)[EPL], int lane) {
#pragma unroll 1
    for (int size = 2; size <= 32 * EPL; size <<= 1) {
#pragma unroll 1
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= EPL) {
                const int lx = stride / EPL;
                const bool lower = (lane & lx) == 0;                    // g < partner
                const bool desc = ((lane * EPL) & size) == 0;           // size >= 2*EPL: uniform over e
                const bool keep_first = lower == desc;
#pragma unroll
                for (int e = 0; e < EPL; ++e) {
                    Cand o;
                    o.s = __shfl_xor_sync(0xffffffffu, c[e].s, lx);
                    o.i = __shfl_xor_sync(0xffffffffu, c[e].i, lx);
                    const bool mine_first = before(c[e], o);
                    if (mine_first != keep_first) c[e] = o;
                }
            } else {
#define DCUE_INTRA(ST)                                                                           \
    _Pragma("unroll") for (int e = 0; e < EPL; ++e) {                                             \
        const int pe = e ^ (ST);                                                                 \
        if (pe > e) {                                                                            \
            const bool desc = ((lane * EPL + e) & size) == 0;                                    \
            const bool in_order = before(c[e], c[pe]);                                           \
            if (in_order != desc) { const Cand t = c[e]; c[e] = c[pe]; c[pe] = t; }              \
        }                                                                                        \
    }
                if (EPL > 4 && stride == 4) { DCUE_INTRA(4 % EPL) }
                else if (stride == 2) { DCUE_INTRA(2) }
                else { DCUE_INTRA(1) }
#undef DCUE_INTRA
            }
        }
    }
}

// Final output of one user: lst[0, n) with n <= 32*EPL -> the k best, sorted, to out_s / out_i (missing: -inf / -1).
// EPL = 4 (n <= 128: the usual k = 100) sorts half as many elements in fewer steps than EPL = 8 (n <= 256).
template <int EPL>
__device__ __noinline__ void sort_and_write(const int2* __restrict__ lst, int n, int k, long item_offset,
                                            float* __restrict__ os, int64_t* __restrict__ oi, int lane) {
    Cand c[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
        const int g = lane * EPL + e;
        int2 v = make_int2(__float_as_int(-INFINITY), -1);
        if (g < n) v = __ldcg(lst + g);
        c[e].s = __int_as_float(v.x);
        c[e].i = v.y;
    }
    bitonic_sort_warp<EPL>(c, lane);
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
        const int g = lane * EPL + e;
        if (g < k) {
            os[g] = c[e].s;
            oi[g] = c[e].i < 0 ? -1 : (int64_t)c[e].i + item_offset;
        }
    }
}

struct TopkShared {          // per-user state and the compaction queue (shared memory)
    float thr[NU];           // score a candidate must beat
    int cnt[NU];             // entries in the user's list (frozen part included)
    int pend_n[NU];          // > 0: [0, pend_n) is frozen and queued / being compacted
    float newthr[NU];        // published by the compactor
    int done[NU];            // compaction finished, waiting for the owner to install it
    int qbuf[NU];            // ring of queued users (a user is queued at most once at a time)
    int q_res, q_tail, q_head, over, exit_flag, pad[3];
};

// pop one queued user (lane 0 decides), compact it, publish.  Returns false if the queue was empty.
__device__ __forceinline__ bool compact_one(TopkShared* sh, int2* __restrict__ mylists, int k, int lane, bool block,
                                            bool* exiting) {
    int slot = -1;
    if (lane == 0) {
        for (;;) {
            const int h = *(volatile int*)&sh->q_head;
            const int t = *(volatile int*)&sh->q_tail;
            if (h < t) {
                if (atomicCAS(&sh->q_head, h, h + 1) == h) { slot = h; break; }
            } else if (!block) {
                break;
            } else if (*(volatile int*)&sh->exit_flag) {
                slot = -2;
                break;
            } else {
                __nanosleep(1000);
            }
        }
    }
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot == -2 && exiting) *exiting = true;
    if (slot < 0) return false;
    __threadfence_block();
    const int u = *(volatile int*)&sh->qbuf[slot & (NU - 1)];
    const int n = *(volatile int*)&sh->pend_n[u];
    const float t = select_topk_inplace(mylists + (size_t)u * CAPH, n, k, lane);
    if (lane == 0) {
        sh->newthr[u] = t;
        __threadfence_block();
        *(volatile int*)&sh->done[u] = 1;
    }
    __syncwarp();
    return true;
}

__global__ void __launch_bounds__(NTHREADS2, 1)
topk_stream_kernel(const uint4* __restrict__ users, long n_utiles128, long n_users, const uint4* __restrict__ items,
                   long n_items, int Kp, int fmt, int k, long item_offset, long items_per_split, int splits,
                   long n_work, int tile_stride /* score every tile_stride-th 128-song tile (threshold pre-pass) */,
                   const float* __restrict__ init_thr /* nullable: per-user starting threshold, stride thr_ld */, long thr_ld,
                   int* __restrict__ n_failed, int thr_only /* write only the k-th best score per user to out_s[user] */,
                   int allow_short /* a seeded user with fewer than k candidates is NOT a failure: its short list is written */,
                   int2* __restrict__ lists /* [grid][NU][CAPH] */, float* __restrict__ out_s,
                   int64_t* __restrict__ out_i /* [splits][n_users][k] */) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npan = Kp / 8;
    const int tile_bytes = npan * PANEL_BYTES;
    uint8_t* sU = smem;                                   // [npan][256 users][16 B]
    uint8_t* sA = sU + npan * UPANEL;                     // NST song tiles, [npan][128 songs][16 B] each
    TopkShared* sh = reinterpret_cast<TopkShared*>(sA + NST * tile_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sh + 1);
    // bars: [0..NST) full, [NST..2NST) empty, UFULL, UEMPTY, TFULL[2], TEMPTY[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 6);
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (NST + s); };
    const uint32_t UFULL = bar0 + 8u * (2 * NST), UEMPTY = bar0 + 8u * (2 * NST + 1);
    auto TFULL = [&](int a) { return bar0 + 8u * (2 * NST + 2 + a); };
    auto TEMPTY = [&](int a) { return bar0 + 8u * (2 * NST + 4 + a); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        mbar_init(UFULL, 1);
        mbar_init(UEMPTY, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(TFULL(a), 1); mbar_init(TEMPTY(a), NEPI); }
        fence_barrier_init();
        sh->q_res = sh->q_tail = sh->q_head = sh->over = sh->exit_flag = 0;
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    int2* mylists = lists + (size_t)blockIdx.x * NU * CAPH;

    // every role walks the same work list: item w = (256-user tile, song split)
    auto song_range = [&](long w, long& ibeg, long& iend) {
        const int sp = (int)(w % splits);
        ibeg = (long)sp * items_per_split;
        iend = ibeg + items_per_split < n_items ? ibeg + items_per_split : n_items;
        if (iend < ibeg) iend = ibeg;
    };

    if (warp == 0) {
        // ===== producer: the user tile once per work item, then the song tiles of its range =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, uphase = 0;
            for (long w = blockIdx.x; w < n_work; w += gridDim.x) {
                const long ut = w / splits;
                long ibeg, iend;
                song_range(w, ibeg, iend);
                mbar_wait(UEMPTY, uphase ^ 1);       // the previous item's MMAs have finished reading sU
                const int halves = (ut * 2 + 1 < n_utiles128) ? 2 : 1;
                mbar_expect_tx(UFULL, (uint32_t)(halves * tile_bytes));
                for (int h = 0; h < halves; ++h)
                    for (int q = 0; q < npan; ++q)
                        bulk_g2s(smem_u32(sU + q * UPANEL + h * PANEL_BYTES), users + ((ut * 2 + h) * (long)npan + q) * 128,
                                 PANEL_BYTES, UFULL);
                uphase ^= 1;
                const long ntiles = ((iend - ibeg + TS - 1) / TS + tile_stride - 1) / tile_stride;
                for (long t = 0; t < ntiles; ++t) {
                    mbar_wait(EMPTY(stage), phase ^ 1);
                    mbar_expect_tx(FULL(stage), (uint32_t)tile_bytes);
                    const long r0 = ibeg + t * tile_stride * TS;   // multiple of 128: tile index r0 >> 7
                    bulk_g2s(smem_u32(sA + stage * tile_bytes), items + (r0 >> 7) * (long)npan * 128, (uint32_t)tile_bytes,
                             FULL(stage));
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: D[song, user] (128 x 256, fp32) = A (songs, K-major) x B (users, K-major)^T =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc(fmt, fmt, TS, NU);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0, uphase = 0;
            for (long w = blockIdx.x; w < n_work; w += gridDim.x) {
                long ibeg, iend;
                song_range(w, ibeg, iend);
                const long ntiles = ((iend - ibeg + TS - 1) / TS + tile_stride - 1) / tile_stride;
                mbar_wait(UFULL, uphase);
                uphase ^= 1;
                tc_fence_after();
                const uint32_t u0s = smem_u32(sU);
                for (long t = 0; t < ntiles; ++t) {
                    mbar_wait(TEMPTY(acc), acc_phase ^ 1);
                    mbar_wait(FULL(stage), phase);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + stage * tile_bytes);
                    for (int c = 0; c < Kp / 16; ++c) {
                        const uint64_t ad = make_desc(a0 + (uint32_t)(2 * c * PANEL_BYTES), PANEL_BYTES, 128);
                        const uint64_t bd = make_desc(u0s + (uint32_t)(2 * c * UPANEL), UPANEL, 128);
                        umma_f16(tmem_base + (uint32_t)(acc * NU), ad, bd, idesc, c != 0);
                    }
                    umma_commit(EMPTY(stage));
                    umma_commit(TFULL(acc));
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                umma_commit(UEMPTY);   // arrives once every MMA of this item has retired
            }
        }
        __syncwarp();
    } else if (warp >= 2 + NEPI) {
        // ===== compactors: pop frozen lists, select the k best in place, publish the new threshold =====
        bool exiting = false;
        while (!exiting) compact_one(sh, mylists, k, lane, true, &exiting);
    } else {
        // ===== appenders: 8 warps = 4 TMEM lane quarters (songs) x 2 column halves (users) =====
        const int e = warp - 2;
        const int quarter = warp & 3;            // TMEM lane quarter this warp may read
        const int half = e >> 2;                 // users [half*CPW, half*CPW + CPW)
        const int et = threadIdx.x - 64;         // 0..NEPI*32-1
        const bool owner_lane = lane < UPW;      // lanes that own a user in the boundary bookkeeping
        const unsigned lt = (1u << lane) - 1u;
        int acc = 0;
        uint32_t acc_phase = 0;
        auto epi_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(NEPI * 32) : "memory"); };

        // Tile-boundary bookkeeping of the users owned by this warp (u = e*32 + lane): install finished compactions,
        // freeze + queue lists that grew beyond LIMIT.  With finishing == false the appenders only wait (and help)
        // while some list could overflow in the next tile; with finishing == true until nothing is pending.
        auto boundary = [&](bool finishing) {
            for (;;) {
                epi_sync();                                   // all appends of the tile are visible
                const int u = e * UPW + (lane % UPW);
                // (a) install finished compactions: survivors of the tail move down behind the k kept entries
                unsigned inst = __ballot_sync(0xffffffffu, owner_lane && *(volatile int*)&sh->done[u] != 0);
                while (inst) {
                    const int l = __ffs(inst) - 1;
                    inst &= inst - 1;
                    const int uu = e * UPW + l;
                    __threadfence_block();
                    const float nt = *(volatile float*)&sh->newthr[uu];
                    const int n0 = sh->pend_n[uu], c = sh->cnt[uu];
                    int2* lst = mylists + (size_t)uu * CAPH;
                    int w = k;
                    for (int j0 = n0; j0 < c; j0 += 32) {      // destinations are always below the sources
                        const int j = j0 + lane;
                        int2 v = make_int2(0, 0);
                        bool keep = false;
                        if (j < c) { v = __ldcg(lst + j); keep = __int_as_float(v.x) > nt; }
                        const unsigned bm = __ballot_sync(0xffffffffu, keep);
                        if (keep) __stcg(lst + w + __popc(bm & lt), v);
                        w += __popc(bm);
                        __syncwarp();
                    }
                    if (lane == 0) {
                        sh->cnt[uu] = w;
                        sh->thr[uu] = nt;
                        sh->pend_n[uu] = 0;
                        *(volatile int*)&sh->done[uu] = 0;
                    }
                    __syncwarp();
                }
                // (b) freeze + queue
                const int cu = sh->cnt[u];
                const bool pending = sh->pend_n[u] != 0;
                // a seeded stream expects ~4k candidates per user in total: do not compact them mid-stream
                // at the end of the stream only lists beyond the selection capacity (512) need a compaction first
                const bool fr = owner_lane && !pending && cu > (finishing ? 512 : (init_thr ? 1024 : LIMIT));
                const unsigned fm = __ballot_sync(0xffffffffu, fr);
                if (fm) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&sh->q_res, __popc(fm));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (fr) {
                        sh->pend_n[u] = cu < 512 ? cu : 512;
                        sh->qbuf[(base + __popc(fm & lt)) & (NU - 1)] = u;
                    }
                    __syncwarp();
                    if (lane == 0) {
                        __threadfence_block();
                        while (atomicCAS(&sh->q_tail, base, base + __popc(fm)) != base) {}   // publish in order
                    }
                }
                const bool wait_for = owner_lane && (finishing ? (pending || fr) : (sh->cnt[u] > HARD));
                // a long queue means many users are streaming against stale thresholds (the start of an unseeded stream
                // freezes all 256 lists at once): every appender warp helps the 4 compactors until it is short again
                const bool backlog = *(volatile int*)&sh->q_tail - *(volatile int*)&sh->q_head > 16;
                if ((__any_sync(0xffffffffu, wait_for) || backlog) && lane == 0) *(volatile int*)&sh->over = 1;
                epi_sync();
                const bool over = *(volatile int*)&sh->over != 0;
                if (!over) break;
                if (!compact_one(sh, mylists, k, lane, false, nullptr)) __nanosleep(200);   // help instead of idling
                epi_sync();
                if (et == 0) *(volatile int*)&sh->over = 0;
            }
        };

        for (long w = blockIdx.x; w < n_work; w += gridDim.x) {
            const long ut = w / splits;
            const int sp = (int)(w % splits);
            long ibeg, iend;
            song_range(w, ibeg, iend);
            const long ntiles = ((iend - ibeg + TS - 1) / TS + tile_stride - 1) / tile_stride;
            // per-user state: users beyond n_users never accept a candidate; a seeded threshold (two-pass mode) is a guess
            // below which nothing is kept -- a user left with fewer than k candidates is reported as failed
            if (et < NU) {
                const long gu0 = ut * NU + et;
                sh->thr[et] = gu0 < n_users ? (init_thr ? init_thr[gu0 * thr_ld] : -INFINITY) : INFINITY;
                sh->cnt[et] = 0;
                sh->pend_n[et] = 0;
                sh->done[et] = 0;
            }
            epi_sync();
            const bool dbg = blockIdx.x == 0 && et == 0;
            long long c0 = clock64();
            for (long t = 0; t < ntiles; ++t) {
                mbar_wait(TFULL(acc), acc_phase);
                tc_fence_after();
                const long song = ibeg + t * tile_stride * TS + quarter * 32 + lane;     // this lane's song
                const bool song_ok = song < iend;
                const int song_i = (int)song;
                const uint32_t tb = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * NU + half * CPW);
                // 128 user columns in 8 steps of 16 (a run-time loop: the fully unrolled version was 70 KB of SASS and
                // "no instruction" was the top stall in ncu); the next 16 columns are in flight while these are compared
                auto process16 = [&](const uint32_t (&r)[16], int ucol) {
                    // one vote per 16 users: all 16 compares are independent (the per-4-column votes were a chain of
                    // LDS -> FSETP -> VOTE -> BRA latencies, ncu: 0.3 IPC with "wait"/"branch resolving" on top)
                    float tt[16];
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 th = *reinterpret_cast<const float4*>(sh->thr + ucol + j);
                        tt[j] = th.x; tt[j + 1] = th.y; tt[j + 2] = th.z; tt[j + 3] = th.w;
                    }
                    unsigned bits = 0;
#pragma unroll
                    for (int j = 0; j < 16; ++j) bits |= (__uint_as_float(r[j]) > tt[j]) ? (1u << j) : 0u;
                    bits = song_ok ? bits : 0u;
                    const unsigned any = __reduce_or_sync(0xffffffffu, bits);
                    if (any) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (any & (0xfu << (4 * g))) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const int j = 4 * g + q;
                                    if (bits & (1u << j)) {   // lanes rarely collide outside the first few tiles
                                        const int u = ucol + j;
                                        const int pos = atomicAdd(&sh->cnt[u], 1);
                                        if (pos < CAPH)   // cannot fail: cnt <= HARD at the last boundary
                                            __stcg(mylists + (size_t)u * CAPH + pos, make_int2((int)r[j], song_i));
                                    }
                                }
                            }
                        }
                    }
                };
                uint32_t ra[16], rb[16];
                tmem_ld16_async(tb, ra);
#pragma unroll 1
                for (int it = 0; it < CPW / 16; it += 2) {
                    tmem_wait_ld();
                    tmem_ld_fence16(ra);
                    tmem_ld16_async(tb + (uint32_t)((it + 1) * 16), rb);
                    process16(ra, half * CPW + it * 16);
                    tmem_wait_ld();
                    tmem_ld_fence16(rb);
                    if (it + 2 < CPW / 16) tmem_ld16_async(tb + (uint32_t)((it + 2) * 16), ra);
                    process16(rb, half * CPW + it * 16 + 16);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(TEMPTY(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                if ((t & (BSTEP - 1)) == BSTEP - 1) boundary(false);
            }
            // ---- finish: every list down to <= 512 entries with nothing pending, then the k best, sorted
            long long c1 = clock64();
            boundary(true);
            long long c2 = clock64();
            for (int uu = 0; uu < UPW; ++uu) {
                const int u = e * UPW + uu;
                const long gu = ut * NU + u;
                if (gu >= n_users) break;
                int n = sh->cnt[u];
                int2* lst = mylists + (size_t)u * CAPH;
                if (thr_only) {   // threshold pre-pass: the k-th best sampled score seeds the next pass (-inf: too few sampled)
                    const float kth = n >= k ? select_topk_inplace(lst, n, k, lane) : -INFINITY;
                    if (lane == 0) out_s[gu] = kth;
                    continue;
                }
                const long ob = ((long)sp * n_users + gu) * k;
                const long avail = iend - ibeg;
                if (init_thr && !allow_short && n < (avail < k ? (int)avail : k)) {
                    // the seeded threshold was too high for this user: mark the row, the caller re-scores it exactly
                    for (int g = lane; g < k; g += 32) { out_s[ob + g] = -INFINITY; out_i[ob + g] = -2; }
                    if (lane == 0) atomicAdd(n_failed, 1);
                    continue;
                }
                if (n > k) { select_topk_inplace(lst, n, k, lane); n = k; }
                if (k <= 128) sort_and_write<4>(lst, n, k, item_offset, out_s + ob, out_i + ob, lane);
                else sort_and_write<8>(lst, n, k, item_offset, out_s + ob, out_i + ob, lane);
            }
            long long c3 = clock64();
            epi_sync();   // nobody resets the per-user state while another warp still reads it
            if (dbg) {
                atomicAdd(&g_topk_dbg[0], (unsigned long long)(c1 - c0));
                atomicAdd(&g_topk_dbg[1], (unsigned long long)(c2 - c1));
                atomicAdd(&g_topk_dbg[2], (unsigned long long)(c3 - c2));
                atomicAdd(&g_topk_dbg[3], (unsigned long long)(clock64() - c3));
                atomicAdd(&g_topk_dbg[4], 1ull);
                atomicAdd(&g_topk_dbg[5], (unsigned long long)ntiles);
            }
        }
        if (et == 0) {
            __threadfence_block();
            *(volatile int*)&sh->exit_flag = 1;
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
    const long ublocks = (n_users + NU - 1) / NU;
    long want = (2L * dcue_num_sms() + ublocks - 1) / ublocks;
    const long maxs = (n_items + 8 * TS - 1) / (8 * TS);
    if (want > maxs) want = maxs;
    if (want > 16) want = 16;
    if (want < 1) want = 1;
    return (int)want;
}
int topk_grid(long n_users, int splits) {
    const long work = ((n_users + NU - 1) / NU) * splits;
    const long sms = dcue_num_sms();
    return (int)(work < sms ? (work > 0 ? work : 1) : sms);
}
size_t topk_list_bytes(int grid) { return (size_t)grid * NU * CAPH * sizeof(int2); }

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
    const size_t parts = splits > 1 ? (size_t)splits * n_users * k * (sizeof(float) + sizeof(int64_t)) : 0;
    return topk_list_bytes(topk_grid(n_users, splits)) + parts + 512;
}

namespace {
// one launch of the streaming scorer (+ the split merge); ws = [candidate lists][split partials]
int launch_topk(const void* users_n, long n_users, const void* items_n, long n_items, int Kp, int fmt, int k, long item_offset,
                int tile_stride, const float* init_thr, long thr_ld, int* n_failed, int thr_only, float* top_scores,
                int64_t* top_idx, void* ws, size_t ws_bytes, cudaStream_t st, int allow_short = 0) {
    const int splits = init_thr || tile_stride > 1 ? 1 : topk_splits(n_users, n_items);
    const int grid = topk_grid(n_users, splits);
    long per = (n_items + splits - 1) / splits;
    per = round_up_l(per > 0 ? per : 1, TS);
    const size_t list_bytes = topk_list_bytes(grid);
    const size_t part_bytes = splits > 1 ? (size_t)splits * n_users * k * (sizeof(float) + sizeof(int64_t)) : 0;
    if (!ws || ws_bytes < list_bytes + part_bytes)
        DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_topk_scores: workspace %zu < %zu", ws_bytes, list_bytes + part_bytes);
    float* os = top_scores;
    int64_t* oi = top_idx;
    if (splits > 1) {
        oi = (int64_t*)((char*)ws + list_bytes);
        os = (float*)((char*)oi + (size_t)splits * n_users * k * sizeof(int64_t));
    }
    const int npan = Kp / 8;
    const size_t smem = (size_t)npan * UPANEL + (size_t)NST * npan * PANEL_BYTES + sizeof(TopkShared) + 8 * (2 * NST + 6) + 16;
    DCUE_CUDA(cudaFuncSetAttribute(topk_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long n_work = ((n_users + NU - 1) / NU) * splits;
    topk_stream_kernel<<<grid, NTHREADS2, smem, st>>>((const uint4*)users_n, round_up_l(n_users, 128) / 128, n_users,
                                                      (const uint4*)items_n, n_items, Kp, fmt, k, item_offset, per, splits, n_work,
                                                      tile_stride, init_thr, thr_ld, n_failed, thr_only, allow_short, (int2*)ws, os, oi);
    DCUE_LAUNCH_CHECK();
    if (splits > 1) {
        topk_merge_kernel<<<ceil_div_i(n_users, 128), 128, 0, st>>>(os, oi, splits, n_users, k, top_scores, top_idx);
        DCUE_LAUNCH_CHECK();
    }
    return 0;
}

// two-pass plan: sample every s-th song tile, seed each user's threshold with the r-th best sampled score, so that about
// 4k scores of the full stream pass it (k is then reached with probability ~1 - 1e-4 per user; the rest is re-scored)
bool topk_2pass_plan(long n_users, long n_items, int k, int* stride, int* r) {
    if (topk_splits(n_users, n_items) != 1 || n_items < 32768 || k > 128) return false;
    const int C = 4 * k;
    int s = C / 20;
    if (s > 16) s = 16;
    if (s < 2) return false;
    *stride = s;
    *r = (C + s - 1) / s;
    return true;
}
}  // namespace

extern "C" int dcue_topk_scores(int impl, const void* users_n, long n_users, const void* items_n, long n_items, int Kp,
                                int fmt, int k, long item_offset, float* top_scores, int64_t* top_idx, void* ws,
                                size_t ws_bytes, void* stream) {
    DCUE_CHECK_ARG(users_n && items_n && top_scores && top_idx && n_users >= 0 && n_items >= 0 && k > 0 && k <= 256);
    DCUE_CHECK_ARG(Kp % 16 == 0 && Kp >= 16 && Kp <= 128 && n_items < (1L << 31));
    if (impl != DCUE_IMPL_TC) DCUE_FAIL(DCUE_E_UNSUPPORTED, "dcue_topk_scores: only the tcgen05 implementation exists");
    if (n_users == 0) return 0;
    return launch_topk(users_n, n_users, items_n, n_items, Kp, fmt, k, item_offset, 1, nullptr, 0, nullptr, 0, top_scores, top_idx, ws,
                       ws_bytes, (cudaStream_t)stream);
}

extern "C" size_t dcue_topk_2pass_ws_bytes(int impl, long n_users, long n_items, int k) {
    int s = 0, r = 0;
    const size_t base = dcue_topk_ws_bytes(impl, n_users, n_items, k);
    if (!topk_2pass_plan(n_users, n_items, k, &s, &r)) return base;
    return base + 2 * (size_t)n_users * sizeof(float) + 512;
}

extern "C" int dcue_topk_scores_2pass(int impl, const void* users_n, long n_users, const void* items_n, long n_items, int Kp,
                                      int fmt, int k, long item_offset, float* top_scores, int64_t* top_idx, int* n_failed,
                                      void* ws, size_t ws_bytes, void* stream) {
    DCUE_CHECK_ARG(users_n && items_n && top_scores && top_idx && n_failed && n_users >= 0 && n_items >= 0 && k > 0 && k <= 256);
    DCUE_CHECK_ARG(Kp % 16 == 0 && Kp >= 16 && Kp <= 128 && n_items < (1L << 31));
    if (impl != DCUE_IMPL_TC) DCUE_FAIL(DCUE_E_UNSUPPORTED, "dcue_topk_scores: only the tcgen05 implementation exists");
    cudaStream_t st = (cudaStream_t)stream;
    DCUE_CUDA(cudaMemsetAsync(n_failed, 0, sizeof(int), st));
    if (n_users == 0) return 0;
    int s = 0, r = 0;
    if (!topk_2pass_plan(n_users, n_items, k, &s, &r))
        return launch_topk(users_n, n_users, items_n, n_items, Kp, fmt, k, item_offset, 1, nullptr, 0, nullptr, 0, top_scores, top_idx,
                           ws, ws_bytes, st);
    const size_t base = topk_list_bytes(topk_grid(n_users, 1));
    if (!ws || ws_bytes < base + 2 * (size_t)n_users * sizeof(float))
        DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_topk_scores_2pass: workspace too small");
    float* thrA = (float*)((char*)ws + base);
    float* thrB = thrA + n_users;
    // level A (only when the stream is long enough to pay for it): every (16 s)-th tile, unseeded, 20th best -> seeds level B.
    // An unseeded stream is expensive at its START (thresholds at -inf flood the lists), so the unseeded level is kept tiny.
    const long tiles = (n_items + TS - 1) / TS;
    const float* seedB = nullptr;
    if (tiles / (16L * s) >= 12) {
        if (int e = launch_topk(users_n, n_users, items_n, n_items, Kp, fmt, 20, 0, 16 * s, nullptr, 0, nullptr, 1, thrA, nullptr, ws, base, st))
            return e;
        seedB = thrA;
    }
    // level B: every s-th tile from level A's seeds (a user level A left short simply starts at -inf), r-th best -> seeds the
    // full stream; level C: all songs.  A seed that turns out too high only costs that user an exact re-score by the caller.
    if (int e = launch_topk(users_n, n_users, items_n, n_items, Kp, fmt, r, 0, s, seedB, 1, n_failed, 1, thrB, nullptr, ws, base, st))
        return e;
    return launch_topk(users_n, n_users, items_n, n_items, Kp, fmt, k, item_offset, 1, thrB, 1, n_failed, 0, top_scores, top_idx, ws,
                       base, st);
}

// ---- global-threshold protocol of the song-sharded eval (parallel.sharded_topk): every shard returns the r best scores of
// its song SAMPLE (every s-th tile); the caller merges the shards' lists per user, takes the r-th best of the union -- a
// threshold that ~4k songs of ALL shards together exceed -- and streams every shard from it with short lists allowed, so a
// shard keeps ~4k/W candidates per user instead of 4k (8x fewer appends, no selection) and the merge sees ~4k entries.
extern "C" int dcue_topk_sample_r(long n_users, long n_items, int k) {
    int s = 0, r = 0;
    return topk_2pass_plan(n_users, n_items, k, &s, &r) ? r : 0;
}
extern "C" size_t dcue_topk_sample_ws_bytes(long n_users, long n_items, int k) {
    int s = 0, r = 0;
    if (!topk_2pass_plan(n_users, n_items, k, &s, &r)) return 512;
    return topk_list_bytes(topk_grid(n_users, 1)) + (size_t)n_users * (sizeof(float) + (size_t)r * sizeof(int64_t)) + 1024;
}
extern "C" int dcue_topk_sample(int impl, const void* users_n, long n_users, const void* items_n, long n_items, int Kp, int fmt,
                                int k, float* sample_scores /* [n_users][r], descending, -inf padded */, void* ws, size_t ws_bytes,
                                void* stream) {
    DCUE_CHECK_ARG(users_n && items_n && sample_scores && ws && n_users >= 0 && n_items >= 0 && k > 0 && k <= 256);
    DCUE_CHECK_ARG(Kp % 16 == 0 && Kp >= 16 && Kp <= 128 && n_items < (1L << 31));
    if (impl != DCUE_IMPL_TC) DCUE_FAIL(DCUE_E_UNSUPPORTED, "dcue_topk_sample: only the tcgen05 implementation exists");
    int s = 0, r = 0;
    if (!topk_2pass_plan(n_users, n_items, k, &s, &r)) DCUE_FAIL(DCUE_E_UNSUPPORTED, "dcue_topk_sample: stream too short to sample");
    if (n_users == 0) return 0;
    if (ws_bytes < dcue_topk_sample_ws_bytes(n_users, n_items, k)) DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_topk_sample: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t base = topk_list_bytes(topk_grid(n_users, 1));
    float* thrA = (float*)((char*)ws + base);
    int64_t* idx_scratch = (int64_t*)((char*)ws + base + round_up_l((long)n_users * sizeof(float), 256));
    const long tiles = (n_items + TS - 1) / TS;
    const float* seedB = nullptr;
    if (tiles / (16L * s) >= 12) {       // long stream: a tiny unseeded level first (see dcue_topk_scores_2pass)
        if (int e = launch_topk(users_n, n_users, items_n, n_items, Kp, fmt, 20, 0, 16 * s, nullptr, 0, nullptr, 1, thrA, nullptr, ws, base, st))
            return e;
        seedB = thrA;
    }
    return launch_topk(users_n, n_users, items_n, n_items, Kp, fmt, r, 0, s, seedB, 1, nullptr, 0, sample_scores, idx_scratch, ws, base,
                       st, /*allow_short=*/1);
}
extern "C" size_t dcue_topk_seeded_ws_bytes(long n_users) { return topk_list_bytes(topk_grid(n_users, 1)) + 512; }
extern "C" int dcue_topk_scores_seeded(int impl, const void* users_n, long n_users, const void* items_n, long n_items, int Kp,
                                       int fmt, int k, long item_offset, const float* thr /* [n_users] */, float* top_scores,
                                       int64_t* top_idx, void* ws, size_t ws_bytes, void* stream) {
    DCUE_CHECK_ARG(users_n && items_n && thr && top_scores && top_idx && ws && n_users >= 0 && n_items >= 0 && k > 0 && k <= 256);
    DCUE_CHECK_ARG(Kp % 16 == 0 && Kp >= 16 && Kp <= 128 && n_items < (1L << 31));
    if (impl != DCUE_IMPL_TC) DCUE_FAIL(DCUE_E_UNSUPPORTED, "dcue_topk_scores_seeded: only the tcgen05 implementation exists");
    if (n_users == 0) return 0;
    if (ws_bytes < topk_list_bytes(topk_grid(n_users, 1))) DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_topk_scores_seeded: workspace too small");
    return launch_topk(users_n, n_users, items_n, n_items, Kp, fmt, k, item_offset, 1, thr, 1, nullptr, 0, top_scores, top_idx, ws,
                       topk_list_bytes(topk_grid(n_users, 1)), (cudaStream_t)stream, /*allow_short=*/1);
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

extern "C" int dcue_topk_debug_cycles(unsigned long long* host_out8, int reset) {
    DCUE_CHECK_ARG(host_out8);
    DCUE_CUDA(cudaMemcpyFromSymbol(host_out8, g_topk_dbg, sizeof(unsigned long long) * 8));
    if (reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        DCUE_CUDA(cudaMemcpyToSymbol(g_topk_dbg, z, sizeof(z)));
    }
    return 0;
}
