// One-shot all-reduce of a tiny fp64 vector over NVLink peer memory (symmetric buffers), one kernel per call.
//
// Data-parallel DCUE needs 13 all-reduces of 2 KB per step (BatchNorm batch statistics forward, the two
// BatchNorm-backward sums per layer: SyncBN semantics, parallel.py) and each sits on the critical path between two
// kernels; an NCCL all-reduce of that size costs ~20-40 us of launch + protocol latency.  Here every rank copies its
// vector into its own symmetric slot, raises a flag in every peer's signal pad (release, system scope), waits for the
// peers' flags (acquire) and sums all slots in RANK ORDER -- so the result is bit-identical on every rank.
// Slots are double buffered by call parity: a rank can only reach call e+2 (which reuses slot e&1) after every rank has
// signalled e+1, i.e. finished reading call e.  The call counter lives in device memory and is advanced by the kernel
// itself, so the kernel is CUDA-graph capturable (every rank issues the same sequence of calls).
#include "common.cuh"

namespace {

constexpr int PEER_SLOT_DOUBLES = 512;
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

__global__ void __launch_bounds__(256)
peer_allreduce_f64_kernel(double* const* __restrict__ bufs, unsigned* const* __restrict__ sigs, unsigned* __restrict__ counter,
                          int rank, int world, double* __restrict__ inout, int n) {
    __shared__ unsigned ep;
    if (threadIdx.x == 0) ep = ++(*counter);
    __syncthreads();
    const unsigned e = ep;
    const int slot = (int)(e & 1u) * PEER_SLOT_DOUBLES;
    double* mine = bufs[rank] + slot;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mine[i] = inout[i];
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) {
        st_release_sys(sigs[threadIdx.x] + rank, e);                       // tell peer `threadIdx.x` that my slot is ready
        const unsigned* my_pad = sigs[rank] + threadIdx.x;
        // peer's call e (or a later one) is published.  The wait is bounded: a rank that died or raised would otherwise hang
        // every other GPU inside this kernel; after PEER_TIMEOUT_NS the flag counter[1] is raised (the result is then invalid)
        unsigned long long t0 = 0;
        unsigned spins = 0;
        while ((int)(ld_acquire_sys(my_pad) - e) < 0) {
            __nanosleep(20);
            if ((++spins & 1023u) == 0) {
                const unsigned long long now = globaltimer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > PEER_TIMEOUT_NS) { atomicExch(reinterpret_cast<int*>(counter) + 1, 1 + (int)threadIdx.x); break; }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += ld_relaxed_sys_f64(bufs[r] + slot + i);
        inout[i] = s;
    }
}

}  // namespace

extern "C" int dcue_peer_allreduce_slot_doubles(void) { return PEER_SLOT_DOUBLES; }

extern "C" int dcue_peer_allreduce_f64(const void* peer_bufs_dev, const void* peer_signals_dev, void* counter, int rank, int world,
                                       double* inout, int n, void* stream) {
    DCUE_CHECK_ARG(peer_bufs_dev && peer_signals_dev && counter && inout && world >= 1 && world <= 64 && rank >= 0 && rank < world);
    DCUE_CHECK_ARG(n >= 0 && n <= PEER_SLOT_DOUBLES);
    if (n == 0) return 0;
    peer_allreduce_f64_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((double* const*)peer_bufs_dev, (unsigned* const*)peer_signals_dev,
                                                                   (unsigned*)counter, rank, world, inout, n);
    DCUE_LAUNCH_CHECK();
    return 0;
}
