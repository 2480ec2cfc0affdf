// One-shot all-reduce of a tiny fp64 vector over NVLink peer memory (symmetric buffers), one kernel per call.
//
// Data-parallel DCUE needs 13 all-reduces of 2 KB per step (BatchNorm batch statistics forward, the two
// BatchNorm-backward sums per layer: SyncBN semantics, parallel.py) and each sits on the critical path between two
// kernels; an NCCL all-reduce of that size costs ~20-40 us of launch + protocol latency.  Here every rank PUSHES its
// vector into its slot of every rank's symmetric buffer, raises a flag in every peer's signal pad (release, system scope),
// waits for the peers' flags (acquire) and sums the slots of its own buffer in RANK ORDER -- bit-identical on every rank.
// Slots are double buffered by call parity: a rank can only reach call e+2 (which reuses parity e&1) after every rank has
// signalled e+1, i.e. finished call e.  The call counter lives in device memory and is advanced by the kernel
// itself, so the kernel is CUDA-graph capturable (every rank issues the same sequence of calls).
#include "common.cuh"
#include "peer.cuh"

namespace {

__global__ void __launch_bounds__(256)
peer_allreduce_f64_kernel(double* const* __restrict__ bufs, unsigned* const* __restrict__ sigs, unsigned* __restrict__ counter,
                          int rank, int world, double* __restrict__ inout, int n) {
    __shared__ unsigned ep;
    PeerCtx pc{bufs, sigs, counter, rank, world};
    peer_allreduce_block(pc, inout, n, &ep);
}

}  // namespace

extern "C" int dcue_peer_allreduce_slot_doubles(void) { return PEER_SLOT_DOUBLES; }
extern "C" long dcue_peer_allreduce_buffer_doubles(void) { return 2L * PEER_MAX_WORLD * PEER_SLOT_DOUBLES; }

extern "C" int dcue_peer_allreduce_f64(const void* peer_bufs_dev, const void* peer_signals_dev, void* counter, int rank, int world,
                                       double* inout, int n, void* stream) {
    DCUE_CHECK_ARG(peer_bufs_dev && peer_signals_dev && counter && inout && world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world);
    DCUE_CHECK_ARG(n >= 0 && n <= PEER_SLOT_DOUBLES);
    if (n == 0) return 0;
    peer_allreduce_f64_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((double* const*)peer_bufs_dev, (unsigned* const*)peer_signals_dev,
                                                                   (unsigned*)counter, rank, world, inout, n);
    DCUE_LAUNCH_CHECK();
    return 0;
}

// ====================================================================================================================
// Row-sharded / replicated user table over NVLink peer memory (BASELINE cfg4 and the table-gradient exchange of cfg3).
//
// Every rank owns one symmetric "exchange" buffer:  [flags: 2 channels x 64 uint32][pad to 1 KB][idx: cap int64][rows: cap x E f32]
// and (row-sharded table) keeps its shard of the table itself in symmetric memory.  Then
//   forward : rows[b] = shard_of_owner(u[b])[local(u[b])] is a plain gather whose loads go over NVLink -- no index
//             all-to-all, no row all-to-all, no padding traffic (dcue_peer_gather_relu_fwd);
//   backward: the user-MLP data gradient is written straight into the rank's `rows` slot; ONE single-CTA kernel publishes
//             the rank's indices, raises a flag in every peer (release.sys), waits for theirs and copies all index lists
//             (dcue_peer_exchange_i64); the owner then segment-sums the gradient rows it owns, reading them from the peers'
//             slots in (rank, position) order -- deterministic, bit-identical wherever two ranks sum the same rows
//             (dcue_peer_scatter_add_rows).
// Single-slot safety: a rank rewrites its slot only after a barrier that every peer reaches after its previous read
// (the forward barrier of the sharded table / the SyncBN all-reduces of the data-parallel step).
// ====================================================================================================================
namespace {

constexpr int XCH_FLAG_BYTES = 1024;
constexpr int XCH_CHANNELS = 2;

__device__ __forceinline__ bool peer_wait(const unsigned* pad, unsigned e, unsigned* timeout_flag) {
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while ((int)(ld_acquire_sys(pad) - e) < 0) {
        __nanosleep(20);
        if ((++spins & 1023u) == 0) {
            const unsigned long long now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > PEER_TIMEOUT_NS) { atomicExch(timeout_flag, 1u + threadIdx.x); return false; }
        }
    }
    return true;
}

// counter[0..1]: call counters of the two channels, counter[2]: timeout flag
__global__ void __launch_bounds__(256)
peer_exchange_i64_kernel(uint8_t* const* __restrict__ bufs, unsigned* __restrict__ counter, int channel, int rank, int world,
                         const int64_t* __restrict__ mine, int n, int64_t* __restrict__ all_out) {
    __shared__ unsigned ep;
    if (threadIdx.x == 0) ep = ++counter[channel];
    __syncthreads();
    const unsigned e = ep;
    int64_t* my_idx = reinterpret_cast<int64_t*>(bufs[rank] + XCH_FLAG_BYTES);
    for (int i = threadIdx.x; i < n; i += blockDim.x) my_idx[i] = mine[i];
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) {
        unsigned* peer_flags = reinterpret_cast<unsigned*>(bufs[threadIdx.x]) + channel * 64;
        st_release_sys(peer_flags + rank, e);
        const unsigned* my_flags = reinterpret_cast<const unsigned*>(bufs[rank]) + channel * 64;
        peer_wait(my_flags + threadIdx.x, e, counter + 2);
    }
    __syncthreads();
    if (all_out) {
        // eight NVLink reads in flight per thread (one at a time, each followed by its store, made this single-block kernel
        // 57 us on 8 GPUs: 32 serial round trips)
        const int total = world * n;
        for (int b0 = threadIdx.x; b0 < total; b0 += 8 * blockDim.x) {
            int64_t v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int g = b0 + q * blockDim.x;
                v[q] = 0;
                if (g < total) {
                    const int r = g / n, i = g - r * n;
                    const int64_t* src = reinterpret_cast<const int64_t*>(bufs[r] + XCH_FLAG_BYTES) + i;
                    asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(v[q]) : "l"(src) : "memory");
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int g = b0 + q * blockDim.x;
                if (g < total) all_out[g] = v[q];
            }
        }
    }
}

// owner / local row of a global row under the contiguous block partition of parallel.shard_slice
__device__ __forceinline__ void block_owner(long r, long base, long rem, int& owner, long& local) {
    const long cut = (base + 1) * rem;
    if (r < cut) { owner = (int)(r / (base + 1)); local = r - (long)owner * (base + 1); }
    else { const long q = (r - cut) / base; owner = (int)(rem + q); local = r - cut - q * base; }
}

__global__ void __launch_bounds__(256)
peer_gather_relu_kernel(const float* const* __restrict__ shards, long base, long rem, const int64_t* __restrict__ idx, int B, long U,
                        int E, float* __restrict__ out, float* __restrict__ raw, int* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    const int64_t r = idx[b];
    float* dst = out + (long)b * E;
    if (r < 0 || r >= U) {
        if (lane == 0) atomicExch(err, 1);
        for (int i = lane; i < E; i += 32) dst[i] = __int_as_float(0x7fc00000);
        return;
    }
    int owner;
    long local;
    block_owner(r, base, rem, owner, local);
    const float* src = shards[owner] + local * (long)E;
    float* rdst = raw ? raw + (long)b * E : nullptr;
    if ((E & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        const int n4 = E / 4;
        for (int i0 = 0; i0 < n4; i0 += 128) {
            float4 v[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int i = i0 + t * 32 + lane;
                if (i < n4) v[t] = __ldcg(s4 + i);          // peer memory: L2 of the owner, not this SM's L1
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int i = i0 + t * 32 + lane;
                if (i < n4) {
                    float4 w = v[t];
                    if (rdst) reinterpret_cast<float4*>(rdst)[i] = w;
                    w.x = fmaxf(w.x, 0.f); w.y = fmaxf(w.y, 0.f); w.z = fmaxf(w.z, 0.f); w.w = fmaxf(w.w, 0.f);
                    reinterpret_cast<float4*>(dst)[i] = w;
                }
            }
        }
    } else {
        for (int i = lane; i < E; i += 32) {
            const float v = __ldcg(src + i);
            if (rdst) rdst[i] = v;
            dst[i] = fmaxf(v, 0.f);
        }
    }
}

// First pass of the exchanged-row reduction: per table row its FIRST entry (as n - j, so that zero = none) and its number of
// entries.  Integer atomics: the result does not depend on the order.
__global__ void __launch_bounds__(256)
peer_mark_rows_kernel(const int64_t* __restrict__ all_idx, int n, long lo, long hi, unsigned* __restrict__ meta /* [hi-lo][2] */) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int64_t row = all_idx[j];
    if (row < lo || row >= hi) return;
    atomicMax(meta + 2 * (row - lo), (unsigned)(n - j));
    atomicAdd(meta + 2 * (row - lo) + 1, 1u);
}

// warp per entry j of the world*B exchanged (index, gradient row) pairs; only the FIRST entry of a row works: it sums the
// row's entries in increasing j (= (rank, position)) order -- deterministic, no sort.  A row with one entry (most of them) is a
// copy; a duplicated row scans the index list from j on and stops at its last entry.  (Round 2: without the marks every warp
// scanned all n indices -- O(n^2 / 32) ballots, 122 us at n = 8 x 1024.)
__global__ void __launch_bounds__(256)
peer_scatter_add_rows_kernel(uint8_t* const* __restrict__ bufs, long rows_off_bytes, const int64_t* __restrict__ all_idx, int world, int B,
                             long lo, long hi, int E, const unsigned* __restrict__ meta, float* __restrict__ gshard) {
    const int lane = threadIdx.x & 31;
    const int n = world * B;
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= n) return;
    const int64_t row = all_idx[j];
    if (row < lo || row >= hi) return;
    if ((int)(n - meta[2 * (row - lo)]) != j) return;          // an earlier entry owns this row
    const int cnt = (int)meta[2 * (row - lo) + 1];
    float* dst = gshard + (row - lo) * (long)E;
    for (int c0 = 0; c0 < E; c0 += 512) {
        float4 a[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) a[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        int found = 0;
        for (int j0 = j & ~31; j0 < n && found < cnt; j0 += 32) {
            const int jj = j0 + lane;
            unsigned m = cnt == 1 ? (j0 == (j & ~31) ? 1u << (j & 31) : 0u)
                                  : __ballot_sync(0xffffffffu, jj >= j && jj < n && __ldg(all_idx + jj) == row);
            found += __popc(m);
            while (m) {
                const int pj = j0 + __ffs(m) - 1;
                m &= m - 1;
                const int r = pj / B, pos = pj - r * B;
                const float* src = reinterpret_cast<const float*>(bufs[r] + rows_off_bytes) + (long)pos * E;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int e0 = c0 + t * 128 + lane * 4;
                    if ((E & 3) == 0) {
                        if (e0 < E) {
                            const float4 g = __ldcg(reinterpret_cast<const float4*>(src + e0));
                            a[t].x += g.x; a[t].y += g.y; a[t].z += g.z; a[t].w += g.w;
                        }
                    } else {
                        float* av = &a[t].x;
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (e0 + q < E) av[q] += __ldcg(src + e0 + q);
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int e0 = c0 + t * 128 + lane * 4;
            if ((E & 3) == 0) {
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

}  // namespace

extern "C" size_t dcue_peer_exchange_bytes(int capacity_rows, int E) {
    return (size_t)XCH_FLAG_BYTES + (size_t)capacity_rows * 8 + (size_t)capacity_rows * E * 4;
}
extern "C" size_t dcue_peer_exchange_rows_offset(int capacity_rows) { return (size_t)XCH_FLAG_BYTES + (size_t)capacity_rows * 8; }

extern "C" int dcue_peer_exchange_i64(const void* peer_bufs_dev, void* counter, int channel, int rank, int world, const int64_t* mine,
                                      int n, int64_t* all_out, void* stream) {
    DCUE_CHECK_ARG(peer_bufs_dev && counter && channel >= 0 && channel < XCH_CHANNELS && world >= 1 && world <= 64 && rank >= 0 &&
                   rank < world && n >= 0 && (n == 0 || mine));
    peer_exchange_i64_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((uint8_t* const*)peer_bufs_dev, (unsigned*)counter, channel, rank,
                                                                  world, mine, n, all_out);
    DCUE_LAUNCH_CHECK();
    return 0;
}

extern "C" int dcue_peer_gather_relu_fwd(const void* peer_shards_dev, long U, int world, const int64_t* idx, int B, int E, float* out,
                                         float* raw_out, int* err_flag, void* stream) {
    DCUE_CHECK_ARG(peer_shards_dev && idx && out && err_flag && U > 0 && world >= 1 && B >= 0 && E > 0);
    if (B == 0) return 0;
    peer_gather_relu_kernel<<<ceil_div_i(B, 8), 256, 0, (cudaStream_t)stream>>>((const float* const*)peer_shards_dev, U / world,
                                                                                 U % world, idx, B, U, E, out, raw_out, err_flag);
    DCUE_LAUNCH_CHECK();
    return 0;
}

extern "C" int dcue_peer_scatter_add_rows(const void* peer_bufs_dev, int capacity_rows, const int64_t* all_idx, int world, int B, long lo,
                                          long hi, int E, float* grad_shard, void* ws, size_t ws_bytes, void* stream) {
    DCUE_CHECK_ARG(peer_bufs_dev && all_idx && grad_shard && world >= 1 && B >= 0 && B <= capacity_rows && hi >= lo && E > 0);
    if (B == 0 || hi == lo) return 0;
    const size_t need = (size_t)(hi - lo) * 2 * sizeof(unsigned);
    if (!ws || ws_bytes < need) DCUE_FAIL(DCUE_E_WORKSPACE, "dcue_peer_scatter_add_rows: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const long n = (long)world * B;
    DCUE_CUDA(cudaMemsetAsync(ws, 0, need, st));
    peer_mark_rows_kernel<<<ceil_div_i(n, 256), 256, 0, st>>>(all_idx, (int)n, lo, hi, (unsigned*)ws);
    DCUE_LAUNCH_CHECK();
    peer_scatter_add_rows_kernel<<<ceil_div_i(n, 8), 256, 0, st>>>(
        (uint8_t* const*)peer_bufs_dev, (long)dcue_peer_exchange_rows_offset(capacity_rows), all_idx, world, B, lo, hi, E,
        (const unsigned*)ws, grad_shard);
    DCUE_LAUNCH_CHECK();
    return 0;
}

// ====================================================================================================================
// Flat gradient all-reduce over NVLink peer memory (data-parallel step, BASELINE cfg3): ONE multi-CTA kernel replaces
// torch.cat + NCCL all-reduce + _foreach_copy_ of the ~1.5 MB of tower / MLP gradients.  Every CTA owns one 16 KB chunk of
// the flat index space: it gathers the chunk from the gradient tensors (device pointer table) into this rank's symmetric
// slot, raises the chunk's flag in every peer (release.sys), waits for the peers' flags of the SAME chunk and sums the
// peers' copies in RANK ORDER straight back into the gradient tensors -- bit-identical on all ranks, and the exchange of
// chunk c overlaps the gather of chunk c+1 on other SMs.  Two slots by call parity (a rank can only reach call e+2 after
// every peer has signalled e+1, i.e. finished reading call e).  The call counter lives in device memory and is advanced by
// the last CTA to finish, so the kernel is CUDA-graph capturable.
// Symmetric buffer layout: [flags: 2 slots x GR_MAX_CHUNKS x GR_MAX_WORLD uint32][data: 2 slots x GR_MAX_CHUNKS x GR_CHUNK f32]
// ====================================================================================================================
namespace {

constexpr int GR_CHUNK = 4096;            // floats per chunk (16 KB)
constexpr int GR_MAX_CHUNKS = 256;        // 1 Mi floats
constexpr int GR_MAX_WORLD = 16;
constexpr size_t GR_FLAG_BYTES = (size_t)2 * GR_MAX_CHUNKS * GR_MAX_WORLD * sizeof(unsigned);
constexpr size_t GR_DATA_BYTES = (size_t)2 * GR_MAX_CHUNKS * GR_CHUNK * sizeof(float);

struct GradEntry {     // one row of the HOST table (int64 x 2 on the Python side)
    float* g;
    long n;
};
// The tensor list travels as a kernel ARGUMENT (1 KB): no device table, no host-to-device copy, so the call can be captured
// into a CUDA graph together with the backward pass that allocates the gradient tensors.
struct GradList {
    float* g[64];
    long pre[65];
};

__global__ void __launch_bounds__(256)
peer_allreduce_grads_kernel(uint8_t* const* __restrict__ bufs, unsigned* __restrict__ counter /* [0] epoch, [1] ticket, [2] timeout */,
                            int rank, int world, const __grid_constant__ GradList gl, int n_tensors, long n_total) {
    __shared__ long pre[65];
    __shared__ float* gp[64];
    __shared__ unsigned last;
    const unsigned e = counter[0] + 1u;          // every CTA reads the epoch before the last one to finish advances it
    const int c = blockIdx.x;
    const int slot = (int)(e & 1u);
    for (int t = threadIdx.x; t <= n_tensors; t += blockDim.x) {
        pre[t] = gl.pre[t];
        if (t < n_tensors) gp[t] = gl.g[t];
    }
    __syncthreads();
    const long f0 = (long)c * GR_CHUNK;
    const int len = (int)min((long)GR_CHUNK, n_total - f0);
    float* mine = reinterpret_cast<float*>(bufs[rank] + GR_FLAG_BYTES) + ((size_t)slot * GR_MAX_CHUNKS + c) * GR_CHUNK;
    // flat index -> (tensor, offset): the chunk spans few tensors, start from the tensor holding f0
    int t0 = 0;
    {
        int lo = 0, hi = n_tensors - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (pre[mid] <= f0) lo = mid; else hi = mid - 1; }
        t0 = lo;
    }
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
        const long f = f0 + i;
        int t = t0;
        while (pre[t + 1] <= f) ++t;
        mine[i] = gp[t][f - pre[t]];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) {
        unsigned* peer_flags = reinterpret_cast<unsigned*>(bufs[threadIdx.x]) + ((size_t)slot * GR_MAX_CHUNKS + c) * GR_MAX_WORLD;
        st_release_sys(peer_flags + rank, e);
        const unsigned* my_flags = reinterpret_cast<const unsigned*>(bufs[rank]) + ((size_t)slot * GR_MAX_CHUNKS + c) * GR_MAX_WORLD;
        peer_wait(my_flags + threadIdx.x, e, counter + 2);
    }
    __syncthreads();
    // 16-byte NVLink reads, all ranks in flight, added in rank order (one 4-byte read at a time, each followed by its add,
    // serialised world x 16 round trips per thread: that version lost to NCCL's LL ring).  The slot is a whole chunk, so the
    // vector read past `len` stays inside it; those lanes are not written back.
    for (int i4 = threadIdx.x * 4; i4 < len; i4 += blockDim.x * 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r0 = 0; r0 < world; r0 += 8) {
            float4 v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r0 + q < world) {
                    const float* src = reinterpret_cast<const float*>(bufs[r0 + q] + GR_FLAG_BYTES) +
                                       ((size_t)slot * GR_MAX_CHUNKS + c) * GR_CHUNK + i4;
                    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(v[q].x), "=f"(v[q].y), "=f"(v[q].z), "=f"(v[q].w) : "l"(src) : "memory");
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (r0 + q < world) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
        }
        const float av[4] = {acc.x, acc.y, acc.z, acc.w};
        int t = t0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long f = f0 + i4 + k;
            if (i4 + k < len) {
                while (pre[t + 1] <= f) ++t;
                gp[t][f - pre[t]] = av[k];
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = (atomicAdd(&counter[1], 1u) == gridDim.x - 1) ? 1u : 0u;
        if (last) {
            counter[1] = 0u;
            __threadfence();
            counter[0] = e;
        }
    }
}

}  // namespace

extern "C" size_t dcue_peer_grads_bytes(void) { return GR_FLAG_BYTES + GR_DATA_BYTES; }
extern "C" long dcue_peer_grads_max_elems(void) { return (long)GR_MAX_CHUNKS * GR_CHUNK; }

extern "C" int dcue_peer_allreduce_grads(const void* peer_bufs_dev, void* counter, int rank, int world, const void* table_host,
                                         int n_tensors, void* stream) {
    DCUE_CHECK_ARG(peer_bufs_dev && counter && table_host && world >= 1 && world <= GR_MAX_WORLD && rank >= 0 && rank < world &&
                   n_tensors >= 1 && n_tensors <= 63);
    GradList gl{};
    const GradEntry* tab = (const GradEntry*)table_host;
    long n_total = 0;
    for (int t = 0; t < n_tensors; ++t) {
        DCUE_CHECK_ARG(tab[t].g != nullptr && tab[t].n >= 0);
        gl.g[t] = tab[t].g;
        gl.pre[t] = n_total;
        n_total += tab[t].n;
    }
    gl.pre[n_tensors] = n_total;
    DCUE_CHECK_ARG(n_total <= dcue_peer_grads_max_elems());
    if (n_total == 0 || world == 1) return 0;
    const int chunks = (int)((n_total + GR_CHUNK - 1) / GR_CHUNK);
    peer_allreduce_grads_kernel<<<chunks, 256, 0, (cudaStream_t)stream>>>((uint8_t* const*)peer_bufs_dev, (unsigned*)counter, rank, world,
                                                                        gl, n_tensors, n_total);
    DCUE_LAUNCH_CHECK();
    return 0;
}
