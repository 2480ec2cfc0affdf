/*
 * dcue_b200.h — C ABI of libdcue_b200.so: hand-written sm_100a kernels for the DCUE
 * (Lee et al. 2018) training + scoring hot path.
 *
 * The reference (estebandito22/Amplifai-DeepContentRecommenders) is pure Python/PyTorch and
 * has no FFI; each entry point below replaces the implicit ATen/cuDNN/cuBLAS kernels that a
 * reference call site issues (file:line given per function, paths relative to the reference
 * root).  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named host_*; buffers are owned by the caller
 *     (PyTorch tensors on the Python side); nothing is allocated or freed here;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on that stream;
 *   - return value: 0 = ok, >0 = cudaError_t, <0 = DCUE_E_* argument error;
 *     dcue_last_error() returns a static message for the calling thread's last failure;
 *   - no global mutable state: re-entrant per stream.
 *
 * 16-bit activation "panel" layout (conv operands, HBM):
 *   element (row r, channel c) of a [rows, C] matrix lives at
 *       base[((c / 8) * panel_rows + r) * 8 + (c % 8)]
 *   i.e. C/8 panels, each a column of 16-byte (8-channel) row chunks.  The same bytes are a
 *   K-major UMMA operand (rows = M/N, channels = K) and an MN-major one (channels = M/N,
 *   rows = K) with no swizzle, so tap shifts are plain 16-byte address offsets.
 *   Rows are "flat": spectrogram s, padded time q -> r = s * Lp + q, with `pad` zero rows in
 *   front of each spectrogram's data and zeros up to Lp.  `base` points at row 0; the buffer
 *   holds DCUE_FRONT_HALO zero rows before it and at least DCUE_BACK_HALO after rows_total.
 */
#ifndef DCUE_B200_H
#define DCUE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCUE_E_BADARG   (-1)
#define DCUE_E_WORKSPACE (-2)
#define DCUE_E_INDEX    (-3)   /* user index out of range (nn.Embedding raises IndexError) */
#define DCUE_E_UNSUPPORTED (-4)

#define DCUE_FMT_F16  0
#define DCUE_FMT_BF16 1

#define DCUE_IMPL_SIMT 0       /* CUDA-core implicit GEMM (bring-up / validator / tiny layers) */
#define DCUE_IMPL_TC   1       /* tcgen05 + TMEM + TMA-bulk implicit GEMM                      */

#define DCUE_FRONT_HALO 8
#define DCUE_BACK_HALO  136

const char* dcue_last_error(void);
int dcue_version(void);
/* number of kernels this library has launched in the calling process (all streams) */
long dcue_launch_count(void);

/* ---------------------------------------------------------------- user tower ------------- */

/* nn.Embedding forward + ReLU   (dcrecommend/dcue/embeddings/userembedding.py:40-41).
 * out[b,:] = relu(table[idx[b],:]); raw (pre-ReLU) rows are optionally copied to raw_out.
 * err_flag (device int, pre-zeroed) is set to 1 if any index is outside [0,U). */
int dcue_gather_relu_fwd(const float* table, const int64_t* idx, int B, int U, int E,
                         float* out, float* raw_out, int* err_flag, void* stream);

/* autograd of the above: dense [U,E] gradient = deterministic sorted segment sum of the
 * ReLU-masked rows (replaces ATen embedding_dense_backward, userembedding.py:27,40).
 * sorted_idx/sorted_pos: idx sorted ascending with the source position of each entry (stable);
 * the sort itself is dcue_sort_indices.  grad_table must be zeroed by the caller. */
int dcue_sort_indices(const int64_t* idx, int B, int U, int64_t* sorted_idx, int32_t* sorted_pos,
                      void* ws, size_t ws_bytes, void* stream);
size_t dcue_sort_ws_bytes(int B);
/* The same dense gradient WITHOUT a sort for step-sized batches (B <= dcue_scatter_direct_max()): one launch, no
 * workspace; the first occurrence of a table row sums its duplicates in position order (bit-identical to the sorted
 * segment sum).  fwd_mask (nullable) = ReLU output of the gather (gradient passes where > 0).  Rows whose index is
 * outside [0,U) are skipped.  grad_table must be zeroed by the caller. */
int dcue_scatter_add_rows(const float* grad_rows, const float* fwd_mask, const int64_t* idx, int B, int U, int E,
                          float* grad_table, void* stream);
int dcue_scatter_direct_max(void);
int dcue_scatter_add_bwd(const float* grad_out, const float* fwd_out /* relu output, mask */,
                         const int64_t* sorted_idx, const int32_t* sorted_pos, int B, int U, int E,
                         float* grad_table, void* stream);

/* nn.Linear (+ optional ReLU) forward / backward, fp32
 * (userembedding.py:42-44; truedcuemel1dbn.py:57-59 (k=1 conv), :101 (fc)).
 * Y[M,N] = act(X[M,K] W[N,K]^T + b[N]); ldx/ldy are row strides in elements. */
int dcue_linear_fwd(const float* X, int ldx, const float* W, const float* b, int M, int K, int N,
                    int relu, float* Y, int ldy, void* stream);
/* dX[M,K] = dY[M,N] W[N,K]; if mask != NULL, dX *= (mask > 0) (ReLU that produced X). */
int dcue_linear_dgrad(const float* dY, int lddy, const float* W, int M, int K, int N,
                      const float* mask, int ldmask, float* dX, int lddx, void* stream);
/* dW[N,K] = dY^T X ; db[N] = colsum(dY).  ws: dcue_linear_wgrad_ws_bytes(M,K,N). */
int dcue_linear_wgrad(const float* dY, int lddy, const float* X, int ldx, int M, int K, int N,
                      float* dW, float* db, void* ws, size_t ws_bytes, void* stream);
size_t dcue_linear_wgrad_ws_bytes(int M, int K, int N);
/* The same three GEMMs with ONE TF32 tensor-core pass per product instead of the fp32-accurate 3xTF32 split: operands are
 * rounded to TF32 (11 significant bits -- the precision of the fp16 conv operands), accumulation is fp32.  Used for the song
 * tower's k = 1 conv and fc (truedcuemel1dbn.py:57-59, :101), whose inputs already carry 16-bit operand rounding; the user
 * MLP keeps the fp32-accurate form.  When the cp.async path is not applicable (unaligned operands, DCUE_LINEAR_IMPL) these
 * fall back to the fp32-accurate kernels. */
int dcue_linear_fwd_tf32(const float* X, int ldx, const float* W, const float* b, int M, int K, int N,
                         int relu, float* Y, int ldy, void* stream);
int dcue_linear_dgrad_tf32(const float* dY, int lddy, const float* W, int M, int K, int N,
                           const float* mask, int ldmask, float* dX, int lddx, void* stream);
int dcue_linear_wgrad_tf32(const float* dY, int lddy, const float* X, int ldx, int M, int K, int N,
                           float* dW, float* db, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- song tower ------------- */

/* Per-channel sum / sum of squares of the NCL fp32 input batch, without torch.cat
 * (dcrecommend/dcue/dcue.py:90 + bn0, truedcuemel1dbn.py:79).  sums = double[2*C]. */
int dcue_ncl_stats(const float* pos, int S_pos, const float* neg, int S_neg, int C, int L,
                   double* sums, void* ws, size_t ws_bytes, void* stream);
size_t dcue_ncl_stats_ws_bytes(int C);

/* Index-based feed (replaces the per-sample torch.load + default_collate of
 * dcrecommend/datasets/dcuedataset.py:235-256 and the 1.4 GB/step H2D copy, nn/dcue.py:195-199):
 * spectrogram s = pool[idx[s], :, off[s] : off[s]+L] of a resident fp32 pool [n_songs, C, T] (off NULL = 0).
 * err_flag (device int) is set if an index / offset is out of range. */
int dcue_ncl_stats_indexed(const float* pool, long n_songs, long T, const int64_t* idx, const int32_t* off, int S,
                           int C, int L, int* err_flag, double* sums, void* ws, size_t ws_bytes, void* stream);
int dcue_ncl_pack_indexed(const float* pool, long n_songs, long T, const int64_t* idx, const int32_t* off, int S,
                          int C, int L, int* err_flag, const float* scale, const float* shift, void* panel,
                          long panel_rows, int Lp, int pad, int fmt, void* stream);

/* BatchNorm finalize (truedcuemel1dbn.py:24,30,38,46,54,61): from sums/count produce
 * scale = gamma*rstd, shift = beta - mean*scale, mean, rstd; training!=0 uses batch statistics
 * and updates running_mean/var (momentum, unbiased var) and num_batches_tracked; otherwise the
 * running statistics are used and nothing is updated.  gamma/beta NULL = identity affine. */
int dcue_bn_finalize(const double* sums, double count, int C, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, int64_t* num_batches_tracked,
                     float momentum, float eps, int training,
                     const float* center /* nullable: sums are statistics of x - center (may alias running_mean) */,
                     float* scale, float* shift, float* mean, float* rstd, void* stream);

/* Fused statistic finalisers: ONE launch replaces "reduce the producers' per-block partials" (+ the data-parallel all-reduce
 * of the sums over NVLink peer memory) + dcue_bn_finalize / dcue_grad_scale.  The producers below leave their partial rows at
 * the START of their workspace when called in "partials" mode (dcue_conv_pool_fwd_parts; sums == NULL for
 * dcue_ncl_center_pack_stats[_indexed] and dcue_bn_bwd_reduce); the row count is the matching *_nparts query.
 * peer_*: as dcue_peer_allreduce_f64 (world == 1: pointers may be NULL).  count = GLOBAL element count per channel.
 * dcue_bn_stats_finalize (training mode): sums_out = global double[2C] (required: the blocks hand their column sums to the
 *   last block through it); other outputs as dcue_bn_finalize.
 * dcue_bn_bwd_finalize: partial = dcue_bn_bwd_reduce's rows ([nparts][2C] sums, then [nparts] max|dy|); sums_out = global
 *   double[2C]; dbeta / dgamma (nullable) fp32 copies; absmax_out (nullable) = local max|dy|; gscale_out (nullable) = {s, 1/s}
 *   as dcue_grad_scale(absmax, scale, C, count). */
int dcue_bn_stats_finalize(const double* partial, int nparts, double count, int C, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                           float eps, const float* center, const void* peer_bufs_dev, const void* peer_signals_dev,
                           void* peer_counter, int rank, int world,
                           void* ticket /* one zero-initialised device uint32 (self-resetting block ticket) */,
                           double* sums_out, float* scale, float* shift, float* mean, float* rstd, void* stream);
int dcue_bn_bwd_finalize(const double* partial, int nparts, int C, const float* scale, double count,
                         const void* peer_bufs_dev, const void* peer_signals_dev, void* peer_counter, int rank, int world,
                         void* ticket, double* sums_out, float* dbeta, float* dgamma, float* absmax_out, float* gscale_out,
                         void* stream);
size_t dcue_ncl_center_pack_stats_nparts(int S, int C);
size_t dcue_bn_bwd_reduce_nparts(int S, int P);
size_t dcue_conv_pool_fwd_nparts(int impl, int S, int Lp);
/* dcue_conv_pool_fwd with the BatchNorm statistic partials ([nparts][2*Cout] doubles) left at the start of ws */
int dcue_conv_pool_fwd_parts(int impl, const void* panel, long panel_rows, int fmt, const void* w_packed,
                             const float* bias, const float* tap_bias, int S, int Lp, int Lin, int pad, int P, int pool,
                             int k, int Cin, int Cout, float* z, uint8_t* code, void* ws, size_t ws_bytes, void* stream);

/* Single pass over the fp32 input (replaces dcue_ncl_stats + dcue_ncl_pack for the BatchNorm towers):
 * u = x - center[c] is written as the 16-bit layer-1 operand panel and its per-channel sum / sum of squares
 * are accumulated in the same sweep (sums = double[2*C]).  center = bn0.running_mean keeps u small whatever
 * the data offset; the exact batch normalisation x-hat = rstd*u + shift is folded into layer 1. */
int dcue_ncl_center_pack_stats(const float* pos, int S_pos, const float* neg, int S_neg, int C, int L,
                               const float* center, void* panel, long panel_rows, int Lp, int pad, int fmt,
                               double* sums, void* ws, size_t ws_bytes, void* stream);
int dcue_ncl_center_pack_stats_indexed(const float* pool, long n_songs, long T, const int64_t* idx,
                                       const int32_t* off, int S, int C, int L, int* err_flag, const float* center,
                                       void* panel, long panel_rows, int Lp, int pad, int fmt, double* sums,
                                       void* ws, size_t ws_bytes, void* stream);

/* x[S,C,L] fp32 (pos rows then neg rows) -> scale*x+shift -> 16-bit panel rows
 * r = s*Lp + pad + t  (fused transpose + convert; scale/shift NULL = identity). */
int dcue_ncl_pack(const float* pos, int S_pos, const float* neg, int S_neg, int C, int L,
                  const float* scale, const float* shift, void* panel, long panel_rows,
                  int Lp, int pad, int fmt, void* stream);

/* Conv1d weight [Cout,Cin,k] fp32 -> 16-bit UMMA A operand [128 x k*128], K-major panels:
 * mode 0 (forward): A[co][j*Cin+ci] = W[co][ci][j];
 * mode 1 (dgrad):   A[ci][jj*Cout+co] = W[co][ci][k-1-jj].   Rows/cols beyond C are zero. */
int dcue_pack_conv_weight(const float* W, int Cout, int Cin, int k, int mode, int fmt,
                          const float* col_scale /* mode 0: W[co][ci][j] *= col_scale[ci]; nullable */,
                          const float* col_scale2 /* second per-ci factor; nullable */, void* out, void* stream);

/* Folding an input-side BatchNorm (gamma, beta) into the conv that consumes it, so that the conv
 * operand is the plain normalised input xhat (truedcuemel1dbn.py:79-80):
 *   conv(gamma*xhat + beta) = conv_{W*gamma}(xhat) + sum over taps j that hit a data row of tapB[j],
 * tapB[j][co] = sum_ci W[co][ci][j]*beta[ci].  tap_bias out = float[k][Cout]. */
int dcue_conv_tap_bias(const float* W, int Cout, int Cin, int k, const float* beta,
                       const float* gamma /* nullable */, const float* shift /* nullable: constant = beta + gamma*shift */,
                       float* tap_bias, void* stream);
/* out[z][i][c] = gscale[1] * (sum over the z-th slice of the spectrograms) panel[s*Lp + r_i][c], i < 4 (r_i < 0 = unused):
 * border row sums of dY as dcue_panel_row_sums_parts() partial sums, out = float[parts][4][128]; dcue_bn_fold_grads adds
 * the parts in fixed order (deterministic). */
int dcue_panel_row_sums_parts(void);
int dcue_panel_row_sums(const void* panel, long panel_rows, int fmt, int S, int Lp, int r0, int r1, int r2, int r3,
                        const float* gscale, float* out, void* stream);
/* Backward of the folded pair from G = wgrad(dY, xhat):  dW = gamma*G + beta*T, dgamma = sum W*G,
 * dbeta = sum W*T, with T[co][j] = Tall[co] - sum of the border row sums E whose tap j is padding.
 * This replaces the layer-1 data gradient and the BatchNorm-backward pass over the input batch. */
int dcue_bn_fold_grads(const float* G, const float* W, const float* gamma, const float* beta,
                       const float* xs_scale, const float* xs_shift /* operand u with x-hat = xs_scale*u + xs_shift; nullable */,
                       const float* Tall, const float* E, int r0, int r1, int r2, int r3, int Cout, int Cin, int k, int pad, int Lin,
                       float* dW, float* dgamma, float* dbeta, void* stream);

/* Conv1d + bias + MaxPool1d(pool) + ReLU with BatchNorm partial sums, implicit GEMM over
 * shifted row views of the panel (truedcuemel1dbn.py:80-83 etc.; all four tower variants).
 * z[S*P, Cout] fp32 = relu(max_{i<pool} conv[s, p*pool+i, :] + bias); code[S*P,Cout] u8 = argmax i;
 * sums (nullable) = double[2*Cout] sum z, sum z^2. */
int dcue_conv_pool_fwd(int impl, const void* panel, long panel_rows, int fmt, const void* w_packed,
                       const float* bias, const float* tap_bias /* dcue_conv_tap_bias output, nullable */,
                       int S, int Lp, int Lin, int pad, int P, int pool, int k, int Cin, int Cout,
                       float* z, uint8_t* code, double* sums, void* ws, size_t ws_bytes, void* stream);

/* Conv1d data gradient: dX[s, t, ci] = sum_{j,co} W[co,ci,j] dY[s, t + pad - j, co] for the
 * Lin data rows of every spectrogram; dY is a bf16/f16 panel in the forward's flat row space
 * (conv output t at row s*Lp + t). dx[S*Lin, Cin] fp32. */
int dcue_conv_dgrad(int impl, const void* dy_panel, long panel_rows, int fmt_dy, const void* w_packed_dgrad,
                    int fmt_w, int S, int Lp, int Lin, int pad, int k, int Cin, int Cout,
                    const float* gscale /* {s, 1/s} of the scaled dY operand, nullable */, float* dx,
                    void* ws, size_t ws_bytes, void* stream);

/* dcue_conv_dgrad (tcgen05 only, Cin == 128) that ALSO takes the BatchNorm-backward reductions of the stage below in its
 * epilogue: the dx it writes is the gradient dy entering that stage's BatchNorm, z / mean / rstd are that stage's
 * pre-BatchNorm activations [S*Lin, Cin] and statistics (zeros / ones for plain sums), dtp (nullable, [S, lddtp]) the gradient
 * of the time average (added as dtp / Lin).  Leaves dcue_bn_bwd_reduce's partial rows at the start of ws
 * (dcue_conv_pool_fwd_nparts(DCUE_IMPL_TC, S, Lp) rows) for dcue_bn_bwd_finalize: replaces a separate sweep over dx and z. */
int dcue_conv_dgrad_stats(const void* dy_panel, long panel_rows, int fmt_dy, const void* w_packed_dgrad, int fmt_w, int S, int Lp,
                          int Lin, int pad, int k, int Cin, int Cout, const float* gscale, float* dx, const float* z,
                          const float* mean, const float* rstd, const float* dtp, int lddtp, void* ws, size_t ws_bytes,
                          void* stream);

/* Conv1d weight gradient dW[co,ci,j] = sum_r dY[r,co] X[r+j,ci] over all flat rows
 * (reference layout [Cout,Cin,k] fp32 out). */
int dcue_conv_wgrad(int impl, const void* dy_panel, long dy_panel_rows, int fmt_dy, const void* x_panel,
                    long x_panel_rows, int fmt_x, long rows_total, int k, int Cin, int Cout,
                    const float* gscale /* {s, 1/s}, nullable */, float* dW, void* ws, size_t ws_bytes,
                    void* stream);
size_t dcue_conv_ws_bytes(int impl, int S, int Lp, int k, int Cin, int Cout);

/* Fused BatchNorm-backward + ReLU mask + MaxPool unpooling + conv weight gradient (tcgen05 only): the 16-bit dY operand
 * is built in shared memory from the pooled inputs (arguments as dcue_bn_relu_unpool_bwd) and never written to HBM.
 * Used where no data gradient is needed afterwards (layer 1).  dW as dcue_conv_wgrad; bias_sums[128] (fp64) /
 * bias_out[128] (fp32, nullable) get the conv bias gradient.  pool == 4, k == 4, 128 channels. */
int dcue_conv_wgrad_unpool(const float* dy, int lddy, const float* dtp /* nullable */, int lddtp, const float* z,
                           const uint8_t* code, const float* scale, const float* mean, const float* rstd,
                           const double* sums /* nullable: no batch-statistics terms */, double count, int S, int P, int pool,
                           int Lp, const void* x_panel, long x_panel_rows, int fmt, int k, int Cin, int Cout,
                           const float* gscale, float* dW, double* bias_sums, float* bias_out, void* ws, size_t ws_bytes,
                           void* stream);
size_t dcue_conv_wgrad_unpool_ws_bytes(int k);
/* Border row sums of the (never materialised) dY taken from the pooled inputs; out as dcue_panel_row_sums. */
int dcue_border_row_sums(const float* dy, int lddy, const float* dtp, int lddtp, const float* z, const uint8_t* code,
                         const float* scale, const float* mean, const float* rstd, const double* sums, double count,
                         int S, int P, int C, int pool, int r0, int r1, int r2, int r3, float* out, void* stream);

/* y = scale*z+shift written (a) as the next layer's 16-bit panel rows s*Lp+pad+p (panel nullable)
 * and/or (b) as fp32 [S*P, C] (y nullable); tp (nullable) [S, ldtp] gets the time average of y
 * (AvgPool1d over the stage's whole extent, truedcuemel1dresbn.py:92,98,104,110). */
int dcue_affine_pack(const float* z, int S, int P, int C, const float* scale, const float* shift,
                     void* panel, long panel_rows, int Lp, int pad, int fmt, float* y, float* tp,
                     int ldtp, void* stream);

/* BatchNorm backward reductions: sums = double[2*C]: sum dy, sum dy*xhat, where xhat=(z-mean)*rstd.
 * dy has row stride lddy (elements); dtp (nullable, [S, lddtp]) is the gradient of the time
 * average, added as dtp/P to every row. */
int dcue_bn_bwd_reduce(const float* dy, int lddy, const float* dtp, int lddtp, const float* z, const float* mean,
                       const float* rstd, int S, int P, int C, double* sums,
                       float* absmax /* nullable: max |dy (+dtp/P)| */,
                       float* dbeta /* nullable: fp32 copy of sums[0:C] */, float* dgamma /* nullable: sums[C:2C] */,
                       void* ws, size_t ws_bytes, void* stream);
/* Power-of-two scale for the 16-bit conv-backward operand (tcgen05 kind::f16 needs both operands in
 * one format, and unscaled fp16 gradients would underflow): out = {s, 1/s}, s the largest power of two
 * with s*bound <= 2^14, bound = max_c|scale_c| * absmax * (count > 0 ? 2 + sqrt(count) : 1) >= max|dz|. */
int dcue_grad_scale(const float* absmax, const float* scale, int C, double count, float* out, void* stream);
size_t dcue_bn_bwd_ws_bytes(int C);

/* BatchNorm backward apply + ReLU mask + MaxPool unpooling in one pass:
 *   dz = scale*(dy - s1/n - xhat*s2/n)   (training)   or   scale*dy (eval / no BN: sums NULL)
 *   dz = 0 where z <= 0
 * and writes (a) dY panel rows s*Lp + p*pool + code (the other pool-1 rows of the group zero)
 * when dy_panel != NULL, or (b) dense dz[S*P,C] fp32 when dz_out != NULL.
 * bias_sums double[C] (nullable) receives sum dz (the conv bias gradient). */
int dcue_bn_relu_unpool_bwd(const float* dy, int lddy, const float* dtp, int lddtp, const float* z,
                            const uint8_t* code, const float* scale, const float* mean, const float* rstd,
                            const double* sums, double count, int S, int P, int C, int pool, int Lp,
                            void* dy_panel, long panel_rows, int fmt,
                            const float* gscale /* panel values are multiplied by gscale[0]; nullable */,
                            float* dz_out, double* bias_sums, float* bias_out /* nullable: fp32 copy of bias_sums */,
                            void* ws, size_t ws_bytes, void* stream);

/* bn0 backward reductions (truedcuemel1dbn.py:79): dx = channels-last [S*L, C] gradient of the
 * normalised input (layer1 dgrad), pos/neg = the NCL fp32 input; sums = double[2*C] as above.
 * ws: dcue_bn_bwd_ws_bytes(C). */
int dcue_ncl_bn_bwd_reduce(const float* dx, const float* pos, int S_pos, const float* neg, int S_neg, int C,
                           int L, const float* mean, const float* rstd, double* sums, void* ws,
                           size_t ws_bytes, void* stream);

/* double[n] -> float[n] with a scale (BN weight/bias grads, conv bias grads). */
int dcue_cvt_f64_f32(const double* in, int n, double mul, float* out, void* stream);

/* ---------------------------------------------------------------- scoring + loss --------- */

/* nn.CosineSimilarity(dim=1) scores (dcrecommend/dcue/dcue.py:93-106):
 * feats = [B positive rows ; B*N negative rows] x F.  scores[b,n] = cos(u_b,pos_b)-cos(u_b,neg_bn). */
int dcue_score_fwd(const float* u, const float* feats, int B, int N, int F, float eps, float* scores,
                   void* stream);
int dcue_score_bwd(const float* u, const float* feats, const float* gscores, int B, int N, int F,
                   float eps, float* du, float* dfeats, void* stream);
/* Fused cosine scores + hinge loss (dcrecommend/nn/dcue.py:167-170) + analytic backward:
 * loss_rows[b] = sum_n max(0, margin - scores[b,n]); du/dfeats = d(mean_b loss_rows)/d(.) with
 * batch_total = global batch size (DP divides by the global B). */
int dcue_score_hinge_fwdbwd(const float* u, const float* feats, int B, int N, int F, float eps,
                            float margin, int batch_total, float* scores, float* loss_rows, float* du,
                            float* dfeats, void* stream);

/* The scalar loss of DCUE._loss_func from the kernel's per-row sums: loss_out[0] = sum_b loss_rows[b] / batch_total (one block,
 * fixed order), and the backward's rescale by the incoming gradient g (a DEVICE scalar): a_out = a * g, b_out = b * g in one
 * launch -- together they replace the ~8 ATen kernels of `loss_rows.sum() / B` and its autograd (dcrecommend/nn/dcue.py:167-170,208). */
int dcue_loss_mean(const float* loss_rows, int B, int batch_total, float* loss_out, void* stream);
int dcue_scale_pair(const float* a, long na, const float* b, long nb, const float* g_dev, float* a_out, float* b_out,
                    void* stream);

/* ---------------------------------------------------------------- optimizer ---------------- */

/* torch.optim.Adam.step() for ALL parameter tensors in one launch (dcrecommend/nn/dcue.py:143-147, :209).
 * table_dev: device array of n_tensors rows {float* p, const float* g, float* m, float* v, int64 n};
 * blk_first_dev[i] = first block of tensor i when tensor j takes ceil(n_j / dcue_adam_elems_per_block()) blocks;
 * total_blocks = their sum.  bias_correction{1,2} = 1 - beta{1,2}^step.  amsgrad = False, L2 weight decay. */
int dcue_adam_multi_step(const void* table_dev, const int* blk_first_dev, int n_tensors, int total_blocks, float lr,
                         float beta1, float beta2, float eps, float weight_decay, float bias_correction1,
                         float bias_correction2,
                         const int* skip_flag_a /* nullable device flags: a non-zero flag makes the step a no-op */,
                         const int* skip_flag_b, void* stream);
int dcue_adam_elems_per_block(void);
/* Ranger.step() = RAdam + Lookahead for ALL parameter tensors in one launch (dcrecommend/optim/ranger.py:82-165).
 * table rows {float* p, const float* g, float* exp_avg, float* exp_avg_sq, float* slow_buffer, int64 n}; blocks as above.
 * step_size_times_lr = RAdam step size (rectified when adaptive != 0) * lr; lookahead != 0 on every k-th step:
 * slow += alpha*(p - slow); p = slow. */
int dcue_ranger_multi_step(const void* table_dev, const int* blk_first_dev, int n_tensors, int total_blocks,
                           float step_size_times_lr, float beta1, float beta2, float eps, float weight_decay_times_lr,
                           int adaptive, int lookahead, float alpha, const int* skip_flag_a, const int* skip_flag_b,
                           void* stream);

/* ---------------------------------------------------------------- multi-GPU ---------------- */

/* One-shot SUM all-reduce of inout[n] (fp64, n <= dcue_peer_allreduce_slot_doubles()) over NVLink peer memory: replaces
 * the 13 latency-bound NCCL all-reduces per data-parallel step (SyncBN statistics; the reference has no multi-GPU path,
 * BASELINE cfg3).  peer_bufs_dev / peer_signals_dev: DEVICE arrays of `world` pointers to every rank's symmetric buffer
 * (2 slots of slot_doubles fp64) and zero-initialised signal pad (>= world uint32); counter: TWO zero-initialised device
 * uint32 of the calling rank (call counter; [1] is raised when a peer did not answer within 20 s -- the result is then
 * invalid and the caller must abort).  Every rank must issue the same sequence of calls.  Result bit-identical on all ranks. */
int dcue_peer_allreduce_f64(const void* peer_bufs_dev, const void* peer_signals_dev, void* counter, int rank, int world,
                            double* inout, int n, void* stream);
int dcue_peer_allreduce_slot_doubles(void);
long dcue_peer_allreduce_buffer_doubles(void);   /* doubles every rank's symmetric buffer must hold (zero-initialised) */

/* User table over NVLink peer memory (BASELINE cfg4: table row-sharded over the GPUs; cfg3: exchange of the table-gradient
 * rows).  Every rank owns a zero-initialised symmetric "exchange" buffer of dcue_peer_exchange_bytes(capacity, E) bytes
 * ([flags][idx: capacity int64][rows: capacity x E f32], the rows start at dcue_peer_exchange_rows_offset(capacity));
 * peer_bufs_dev = DEVICE array of `world` pointers to them.  counter = 3 zero-initialised device uint32 of the calling rank
 * (call counters of the two flag channels, [2] = time-out flag as in dcue_peer_allreduce_f64).
 * dcue_peer_exchange_i64: publish mine[n] in my idx slot, barrier over all ranks (channel 0 or 1), copy every rank's list
 *   to all_out[world][n] (nullable; n == 0 = pure barrier).  Every rank issues the same sequence of calls per channel.
 * dcue_peer_gather_relu_fwd: as dcue_gather_relu_fwd for a table whose rows are block-sharded (rank r owns
 *   [r*U/W + min(r, U%W), ...)) in symmetric memory: peer_shards_dev = `world` pointers to the shards; rows come over NVLink.
 * dcue_peer_scatter_add_rows: grad_shard[row - lo] = sum of the gradient rows (read from the peers' `rows` slots) whose index
 *   (all_idx from dcue_peer_exchange_i64) is `row`, for lo <= row < hi, summed in (rank, position) order.  grad_shard zeroed
 *   by the caller; ws = 8 * (hi - lo) bytes of scratch (first entry and entry count of every row). */
size_t dcue_peer_exchange_bytes(int capacity_rows, int E);
size_t dcue_peer_exchange_rows_offset(int capacity_rows);
int dcue_peer_exchange_i64(const void* peer_bufs_dev, void* counter, int channel, int rank, int world, const int64_t* mine,
                           int n, int64_t* all_out, void* stream);
int dcue_peer_gather_relu_fwd(const void* peer_shards_dev, long U, int world, const int64_t* idx, int B, int E, float* out,
                              float* raw_out, int* err_flag, void* stream);
int dcue_peer_scatter_add_rows(const void* peer_bufs_dev, int capacity_rows, const int64_t* all_idx, int world, int B, long lo,
                               long hi, int E, float* grad_shard, void* ws /* 8 * (hi - lo) bytes */, size_t ws_bytes,
                               void* stream);

/* Flat SUM all-reduce of a list of fp32 gradient tensors over NVLink peer memory in ONE multi-CTA kernel (replaces
 * torch.cat + NCCL all-reduce + scatter-back of the data-parallel step's tower / MLP gradients, BASELINE cfg3).
 * peer_bufs_dev: `world` pointers to zero-initialised symmetric buffers of dcue_peer_grads_bytes() bytes; counter: 3
 * zero-initialised device uint32 of the calling rank (epoch, CTA ticket, time-out flag); table_host: n_tensors <= 63 rows
 * {float* grad, int64 n} in HOST memory (they travel as a kernel argument: graph-capturable, no copy); sum of n <=
 * dcue_peer_grads_max_elems().
 * The sums are written back into the gradient tensors, bit-identical on every rank.  Every rank issues the same calls. */
size_t dcue_peer_grads_bytes(void);
long dcue_peer_grads_max_elems(void);
int dcue_peer_allreduce_grads(const void* peer_bufs_dev, void* counter, int rank, int world, const void* table_host,
                              int n_tensors, void* stream);

/* ---------------------------------------------------------------- eval scorer ------------ */

/* row-normalise factors (x / max(||x||,eps)) into 16-bit K-major rows padded to Kp (mult of 16). */
int dcue_normalize_rows(const float* x, long rows, int F, float eps, int Kp, int fmt, void* out,
                        void* stream);
/* All-pairs cosine score GEMM + fused top-k (generalises DCUE.predict, nn/dcue.py:495-513):
 * for each user row the k best (score, item index + item_offset) among `items` rows, sorted
 * descending (k <= 256; missing slots: score -inf, index -1).  users_n/items_n: outputs of
 * dcue_normalize_rows.  ws: dcue_topk_ws_bytes() bytes (per-CTA candidate lists + split partials). */
int dcue_topk_scores(int impl, const void* users_n, long n_users, const void* items_n, long n_items,
                     int Kp, int fmt, int k, long item_offset, float* top_scores, int64_t* top_idx,
                     void* ws, size_t ws_bytes, void* stream);
size_t dcue_topk_ws_bytes(int impl, long n_users, long n_items, int k);
/* Same result in two passes: pass 1 scores every s-th 128-song tile and seeds each user's threshold with the r-th best
 * sampled score (s*r ~ 4k), pass 2 streams all songs from those thresholds (4x fewer candidates, no list compaction).
 * A user whose seed was too high (fewer than k candidates, probability ~1e-4) gets index -2 in every slot and is counted
 * in *n_failed (device int): the caller re-scores those rows with dcue_topk_scores.  Falls back to the single pass for
 * small inputs. */
int dcue_topk_scores_2pass(int impl, const void* users_n, long n_users, const void* items_n, long n_items,
                           int Kp, int fmt, int k, long item_offset, float* top_scores, int64_t* top_idx,
                           int* n_failed, void* ws, size_t ws_bytes, void* stream);
size_t dcue_topk_2pass_ws_bytes(int impl, long n_users, long n_items, int k);
/* Diagnostics: cycle counters of the scorer's phases accumulated by block 0's first appender warp since the last reset:
 * host_out8 (HOST pointer) = {stream loop, final boundary, select+sort+write, trailing barrier, work items, song tiles, 0, 0}. */
int dcue_topk_debug_cycles(unsigned long long* host_out8, int reset);
/* Global-threshold protocol of the song-sharded eval: dcue_topk_sample returns, per user, the r = dcue_topk_sample_r()
 * best scores of THIS shard's song sample (every s-th 128-song tile; r*s ~ 4k; r == 0: stream too short, do not sample);
 * the caller merges the shards' lists, takes the r-th best of the union as the user's threshold and calls
 * dcue_topk_scores_seeded on every shard: all songs above the threshold (at most the k best), descending; a user with fewer
 * than k of them gets a short list padded with (-inf, -1) -- the merged lists then hold ~4k entries per user, and a user
 * whose merged list is shorter than k must be re-scored with dcue_topk_scores (probability ~1e-4). */
int dcue_topk_sample_r(long n_users, long n_items, int k);
size_t dcue_topk_sample_ws_bytes(long n_users, long n_items, int k);
int dcue_topk_sample(int impl, const void* users_n, long n_users, const void* items_n, long n_items, int Kp, int fmt,
                     int k, float* sample_scores, void* ws, size_t ws_bytes, void* stream);
size_t dcue_topk_seeded_ws_bytes(long n_users);
int dcue_topk_scores_seeded(int impl, const void* users_n, long n_users, const void* items_n, long n_items, int Kp,
                            int fmt, int k, long item_offset, const float* thr, float* top_scores,
                            int64_t* top_idx, void* ws, size_t ws_bytes, void* stream);
/* merge `parts` per-shard top-k lists [parts][n_users][k] into one (song-sharded eval). */
int dcue_topk_merge(const float* scores, const int64_t* idx, int parts, long n_users, int k,
                    float* out_scores, int64_t* out_idx, void* stream);

/* ---------------------------------------------------------------- ranking metrics -------- */

/* ROC-AUC and average precision per segment on the device (replaces sklearn's roc_auc_score / average_precision_score
 * over Python lists in DCUE.score / DCUE.score_song, dcrecommend/nn/dcue.py:380-476).  Segment i = elements
 * [seg_offsets[i], seg_offsets[i+1]) of scores / targets (0/1) / group (0/1, nullable = all 0).  Ties are handled like
 * sklearn (thresholds = distinct scores).  out[i] = double[8]: {AUC of half 0, AUC of half 1, elements in half 0 / 1,
 * positives in half 0 / 1, AP over the whole segment, positives in the segment}; a half with only positives has AUC 1,
 * one with no positive AUC 0 (the reference's conventions). */
int dcue_auc_ap_segments(const float* scores, const uint8_t* targets, const uint8_t* group, const int64_t* seg_offsets,
                         int n_segments, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DCUE_B200_H */
