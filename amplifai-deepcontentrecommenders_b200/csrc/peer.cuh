// Device-side one-shot all-reduce over NVLink peer memory, shared by peer.cu (stand-alone kernel) and bn.cu (fused into the
// BatchNorm statistic finalisers).  See peer.cu for the protocol.
#pragma once
#include "common.cuh"

constexpr int PEER_SLOT_DOUBLES = 512;
constexpr int PEER_MAX_WORLD = 16;                // source slots per call parity in every rank's buffer
constexpr unsigned long long PEER_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;   // 20 s

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

struct PeerCtx {
    double* const* bufs;       // device array of `world` pointers to the symmetric buffers: [2 parities][PEER_MAX_WORLD][PEER_SLOT_DOUBLES]
    unsigned* const* sigs;     // device array of `world` pointers to the signal pads
    unsigned* counter;         // [0] call counter of this rank, [1] time-out flag
    int rank, world;
};

// All-reduce (SUM, rank order -> bit-identical on every rank) of vec[0, n) held by ONE thread block (shared or global
// memory, visible to all its threads); every thread of the block must call this.  `ep_sh` is a shared scratch word.
// PUSH protocol: every rank stores its vector into slot [parity][rank] of EVERY rank's buffer (fire-and-forget NVLink writes),
// fences, raises its flag in every peer (release.sys) and waits for the peers' flags; the sum then reads only LOCAL memory.
// (The first version pulled: flag, then NVLink reads of the peers' slots -- one more round trip on the critical path of each
// of the 11 BatchNorm exchanges of a step.)  Slots are double buffered by call parity: a rank can only reach call e+2 (which
// rewrites parity e&1 in its peers) after every peer has signalled e+1, i.e. has finished call e.
__device__ __forceinline__ void peer_allreduce_block(const PeerCtx& pc, double* vec, int n, unsigned* ep_sh) {
    __syncthreads();
    if (threadIdx.x == 0) *ep_sh = ++(*pc.counter);
    __syncthreads();
    const unsigned e = *ep_sh;
    const long slot = ((long)(e & 1u) * PEER_MAX_WORLD + pc.rank) * PEER_SLOT_DOUBLES;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = vec[i];
        for (int r = 0; r < pc.world; ++r) pc.bufs[r][slot + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < pc.world) {
        st_release_sys(pc.sigs[threadIdx.x] + pc.rank, e);                 // tell peer `threadIdx.x` that my vector has landed
        const unsigned* my_pad = pc.sigs[pc.rank] + threadIdx.x;
        // peer's call e (or a later one) is published.  The wait is bounded: a rank that died or raised would otherwise hang
        // every other GPU inside this kernel; after PEER_TIMEOUT_NS the flag counter[1] is raised (the result is then invalid)
        unsigned long long t0 = 0;
        unsigned spins = 0;
        while ((int)(ld_acquire_sys(my_pad) - e) < 0) {
            __nanosleep(20);
            if ((++spins & 1023u) == 0) {
                const unsigned long long now = globaltimer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > PEER_TIMEOUT_NS) { atomicExch(reinterpret_cast<int*>(pc.counter) + 1, 1 + (int)threadIdx.x); break; }
            }
        }
    }
    __syncthreads();
    const double* mine = pc.bufs[pc.rank] + (long)(e & 1u) * PEER_MAX_WORLD * PEER_SLOT_DOUBLES;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int r0 = 0; r0 < pc.world; r0 += 8) {
            double v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = r0 + q < pc.world ? ld_relaxed_sys_f64(mine + (long)(r0 + q) * PEER_SLOT_DOUBLES + i) : 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (r0 + q < pc.world) s += v[q];
        }
        vec[i] = s;
    }
    __syncthreads();
}
