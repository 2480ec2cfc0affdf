// Ranking metrics on the device: ROC-AUC and average precision per segment (one user's or one song's candidate list),
// replacing the per-user Python lists + sklearn.metrics.roc_auc_score / average_precision_score round trip of
// DCUE.score / DCUE.score_song (dcrecommend/nn/dcue.py:380-476).
//
// Both metrics are written as exact COUNTS per positive element, so no sort is needed and ties are handled the way
// sklearn handles them (thresholds are the distinct score values):
//   AUC = sum_{i positive} ( #{j negative: s_j < s_i} + 0.5 #{j negative: s_j == s_i} ) / (P * N)      (Mann-Whitney)
//   AP  = (1/P) sum_{i positive} TP(s_i) / (TP(s_i) + FP(s_i)),  TP(t) = #{positives with s >= t}, FP(t) likewise
//         ( = sum over distinct thresholds of (recall step) * precision, sklearn's definition )
// One CTA per segment; the segment streams through shared memory in tiles, every thread owns a strided set of elements.
// Counts are integers and the final sums are reduced in a fixed order in fp64: results are deterministic.
//
// DCUE.score mixes two splits per user (positives of one with negatives of the other, nn/dcue.py:405-418): `group`
// (0/1 per element, nullable) selects the half an element belongs to; the AUC is computed per half, the AP over all
// elements of the segment, exactly as the reference does.
#include "common.cuh"

namespace {

constexpr int MET_THREADS = 256;
constexpr int MET_TILE = 2048;

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
    v = warp_sum_d(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < MET_THREADS / 32; ++w) t += sh[w];     // fixed order
    return t;   // valid on thread 0
}

__global__ void __launch_bounds__(MET_THREADS)
auc_ap_kernel(const float* __restrict__ scores, const uint8_t* __restrict__ targets, const uint8_t* __restrict__ group,
              const int64_t* __restrict__ seg, double* __restrict__ out /* [n_seg][8] */) {
    __shared__ float ts[MET_TILE];
    __shared__ uint8_t tc[MET_TILE];      // bit 0 = target, bit 1 = group
    __shared__ double red[MET_THREADS / 32];
    const long beg = seg[blockIdx.x], end = seg[blockIdx.x + 1];
    const long n = end - beg;
    double auc_part[2] = {0.0, 0.0}, ap_part = 0.0;
    long npos[2] = {0, 0}, ncnt[2] = {0, 0};
    for (long i0 = 0; i0 < n; i0 += MET_THREADS) {
        const long i = i0 + threadIdx.x;
        const bool act = i < n;
        const float si = act ? scores[beg + i] : 0.f;
        const int ti = act ? (targets[beg + i] != 0) : 0;
        const int gi = (act && group) ? (group[beg + i] != 0) : 0;
        if (act) { ncnt[gi]++; npos[gi] += ti; }
        // counts against every element of the segment
        unsigned less_neg = 0, eq_neg = 0, ge_pos = 0, ge_neg = 0;
        for (long j0 = 0; j0 < n; j0 += MET_TILE) {
            __syncthreads();
            for (int t = threadIdx.x; t < MET_TILE && j0 + t < n; t += MET_THREADS) {
                ts[t] = scores[beg + j0 + t];
                tc[t] = (uint8_t)((targets[beg + j0 + t] != 0) | ((group && group[beg + j0 + t]) ? 2 : 0));
            }
            __syncthreads();
            if (act && ti) {
                const int lim = (int)min((long)MET_TILE, n - j0);
                for (int t = 0; t < lim; ++t) {
                    const float sj = ts[t];
                    const int c = tc[t];
                    const int tj = c & 1;
                    const bool same_half = (c >> 1) == gi;
                    ge_pos += (tj && sj >= si);
                    ge_neg += (!tj && sj >= si);
                    less_neg += (!tj && same_half && sj < si);
                    eq_neg += (!tj && same_half && sj == si);
                }
            }
        }
        if (act && ti) {
            auc_part[gi] += (double)less_neg + 0.5 * (double)eq_neg;
            ap_part += (double)ge_pos / (double)(ge_pos + ge_neg);
        }
    }
    double r[7];
    r[0] = block_sum_d(auc_part[0], red);
    r[1] = block_sum_d(auc_part[1], red);
    r[2] = block_sum_d(ap_part, red);
    r[3] = block_sum_d((double)npos[0], red);
    r[4] = block_sum_d((double)npos[1], red);
    r[5] = block_sum_d((double)ncnt[0], red);
    r[6] = block_sum_d((double)ncnt[1], red);
    if (threadIdx.x == 0) {
        double* o = out + (long)blockIdx.x * 8;
        for (int g = 0; g < 2; ++g) {
            const double P = r[3 + g], N = r[5 + g] - r[3 + g];
            // the reference's conventions (nn/dcue.py:433-438, :463-468): all positive -> 1, no positive -> 0
            o[g] = (r[5 + g] > 0 && N == 0) ? 1.0 : (P == 0 ? 0.0 : r[g] / (P * N));
            o[2 + g] = r[5 + g];     // elements in the half
            o[4 + g] = P;            // positives in the half
        }
        const double Pall = r[3] + r[4];
        o[6] = Pall > 0 ? r[2] / Pall : 0.0;
        o[7] = Pall;
    }
}

}  // namespace

extern "C" int dcue_auc_ap_segments(const float* scores, const uint8_t* targets, const uint8_t* group, const int64_t* seg_offsets,
                                    int n_segments, double* out, void* stream) {
    DCUE_CHECK_ARG(scores && targets && seg_offsets && out && n_segments >= 0);
    if (n_segments == 0) return 0;
    auc_ap_kernel<<<n_segments, MET_THREADS, 0, (cudaStream_t)stream>>>(scores, targets, group, seg_offsets, out);
    DCUE_LAUNCH_CHECK();
    return 0;
}
