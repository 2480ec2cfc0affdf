// extern "C" entry points of the conv stage kernels: argument validation + dispatch to the
// tcgen05 (DCUE_IMPL_TC) or CUDA-core (DCUE_IMPL_SIMT) implementation.
#include "common.cuh"
#include "conv_common.cuh"

thread_local char g_dcue_err[256] = "";
long g_dcue_launches = 0;

extern "C" const char* dcue_last_error(void) { return g_dcue_err; }
extern "C" int dcue_version(void) { return 100; }
extern "C" long dcue_launch_count(void) { return g_dcue_launches; }

static int check_geom(const ConvGeom& g, long panel_rows) {
    DCUE_CHECK_ARG(g.S >= 0 && g.Lp > 0 && g.k >= 1 && g.k <= 4 && g.Cin > 0 && g.Cin <= 128 && g.Cin % 8 == 0);
    DCUE_CHECK_ARG(g.Cout > 0 && g.Cout <= 128);
    DCUE_CHECK_ARG(g.rows_total < (1L << 31) - 4096);  // flat rows are indexed with 32-bit ints in the epilogues
    // tiles may read k-1 rows past a 128-row boundary: the back halo covers it
    DCUE_CHECK_ARG(panel_rows >= round_up_l(g.rows_total, 128) + 16);
    return 0;
}

extern "C" size_t dcue_conv_ws_bytes(int impl, int S, int Lp, int k, int Cin, int Cout) {
    (void)S; (void)Lp; (void)Cin; (void)Cout;
    size_t stats = (size_t)dcue_num_sms() * 4 * (2 * 128 + 1) * sizeof(double) + 256;
    size_t wg = (size_t)32 * 128 * k * 128 * sizeof(float);
    size_t tc = impl == DCUE_IMPL_TC ? dcue_tc_ws_bytes(k) : 0;
    size_t m = stats > wg ? stats : wg;
    return (m > tc ? m : tc) + 256;
}

extern "C" int dcue_conv_pool_fwd(int impl, const void* panel, long panel_rows, int fmt, const void* w_packed,
                                  const float* bias, const float* tap_bias, int S, int Lp, int Lin, int pad, int P, int pool,
                                  int k, int Cin, int Cout, float* z, uint8_t* code, double* sums, void* ws, size_t ws_bytes,
                                  void* stream) {
    DCUE_CHECK_ARG(panel && w_packed && z && (pool == 1 || pool == 2 || pool == 4) && P > 0 && P * pool <= Lp &&
                   Lp % pool == 0);
    DCUE_CHECK_ARG(Lin > 0 && pad >= 0 && Lin + pad <= Lp);
    ConvGeom g{S, Lp, Lin, pad, k, pool, P, Cin, Cout, (long)S * Lp};
    if (int e = check_geom(g, panel_rows)) return e;
    if (S == 0) return 0;
    if (impl == DCUE_IMPL_TC)
        return dcue_tc_conv_fwd(panel, panel_rows, fmt, w_packed, bias, tap_bias, g, z, code, sums, ws, ws_bytes, (cudaStream_t)stream);
    return dcue_simt_conv_fwd(panel, panel_rows, fmt, w_packed, bias, tap_bias, g, z, code, sums, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int dcue_conv_pool_fwd_parts(int impl, const void* panel, long panel_rows, int fmt, const void* w_packed,
                                        const float* bias, const float* tap_bias, int S, int Lp, int Lin, int pad, int P, int pool,
                                        int k, int Cin, int Cout, float* z, uint8_t* code, void* ws, size_t ws_bytes, void* stream) {
    return dcue_conv_pool_fwd(impl, panel, panel_rows, fmt, w_packed, bias, tap_bias, S, Lp, Lin, pad, P, pool, k, Cin, Cout, z, code,
                              DCUE_STATS_PARTIALS, ws, ws_bytes, stream);
}
extern "C" size_t dcue_conv_pool_fwd_nparts(int impl, int S, int Lp) {
    const long rows = (long)S * Lp;
    return (size_t)(impl == DCUE_IMPL_TC ? dcue_tc_conv_fwd_nparts(rows) : dcue_simt_conv_fwd_nparts(rows));
}

extern "C" int dcue_conv_dgrad(int impl, const void* dy_panel, long panel_rows, int fmt_dy, const void* w_packed_dgrad,
                               int fmt_w, int S, int Lp, int Lin, int pad, int k, int Cin, int Cout, const float* gscale,
                               float* dx, void* ws, size_t ws_bytes, void* stream) {
    DCUE_CHECK_ARG(dy_panel && w_packed_dgrad && dx && Lin > 0 && pad >= 0 && Lin + pad <= Lp && k - 1 <= DCUE_FRONT_HALO);
    // GEMM view: contraction over the conv's Cout, output channels = the conv's Cin
    ConvGeom g{S, Lp, Lin, pad, k, 1, 0, Cout, Cin, (long)S * Lp};
    DCUE_CHECK_ARG(Cout % 8 == 0);
    if (int e = check_geom(g, panel_rows)) return e;
    if (S == 0) return 0;
    // In'[r] = dY[r - (k-1)]: shift the base pointer back by k-1 rows (front halo rows are zero)
    const char* shifted = (const char*)dy_panel - (size_t)(k - 1) * 16;
    if (impl == DCUE_IMPL_TC)
        return dcue_tc_conv_dgrad(shifted, panel_rows, fmt_dy, w_packed_dgrad, fmt_w, g, gscale, dx, ws, ws_bytes, (cudaStream_t)stream);
    return dcue_simt_conv_dgrad(shifted, panel_rows, fmt_dy, w_packed_dgrad, fmt_w, g, gscale, dx, (cudaStream_t)stream);
}

extern "C" int dcue_conv_dgrad_stats(const void* dy_panel, long panel_rows, int fmt_dy, const void* w_packed_dgrad, int fmt_w, int S, int Lp,
                                     int Lin, int pad, int k, int Cin, int Cout, const float* gscale, float* dx, const float* z,
                                     const float* mean, const float* rstd, const float* dtp, int lddtp, void* ws, size_t ws_bytes,
                                     void* stream) {
    DCUE_CHECK_ARG(dy_panel && w_packed_dgrad && dx && z && mean && rstd && Lin > 0 && pad >= 0 && Lin + pad <= Lp &&
                   k - 1 <= DCUE_FRONT_HALO);
    DCUE_CHECK_ARG(Cout % 8 == 0 && Cin == 128 && (!dtp || lddtp >= Cin));
    ConvGeom g{S, Lp, Lin, pad, k, 1, 0, Cout, Cin, (long)S * Lp};
    if (int e = check_geom(g, panel_rows)) return e;
    if (S == 0) return 0;
    const char* shifted = (const char*)dy_panel - (size_t)(k - 1) * 16;
    return dcue_tc_conv_dgrad_stats(shifted, panel_rows, fmt_dy, w_packed_dgrad, fmt_w, g, gscale, dx, z, mean, rstd, dtp, lddtp, ws,
                                    ws_bytes, (cudaStream_t)stream);
}

extern "C" int dcue_conv_wgrad(int impl, const void* dy_panel, long dy_panel_rows, int fmt_dy, const void* x_panel,
                               long x_panel_rows, int fmt_x, long rows_total, int k, int Cin, int Cout,
                               const float* gscale, float* dW, void* ws, size_t ws_bytes, void* stream) {
    DCUE_CHECK_ARG(dy_panel && x_panel && dW && rows_total >= 0 && k >= 1 && k <= 4 && Cin > 0 && Cin <= 128 &&
                   Cin % 8 == 0 && Cout > 0 && Cout <= 128 && Cout % 8 == 0);
    DCUE_CHECK_ARG(dy_panel_rows >= round_up_l(rows_total, 128) + 16 && x_panel_rows >= round_up_l(rows_total, 128) + 16);
    if (impl == DCUE_IMPL_TC)
        return dcue_tc_conv_wgrad(dy_panel, dy_panel_rows, fmt_dy, x_panel, x_panel_rows, fmt_x, rows_total, k, Cin, Cout,
                                  gscale, dW, ws, ws_bytes, (cudaStream_t)stream);
    return dcue_simt_conv_wgrad(dy_panel, dy_panel_rows, fmt_dy, x_panel, x_panel_rows, fmt_x, rows_total, k, Cin, Cout,
                                gscale, dW, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" size_t dcue_conv_wgrad_unpool_ws_bytes(int k) { return dcue_tc_wgrad_unpool_ws_bytes(k); }

extern "C" int dcue_conv_wgrad_unpool(const float* dy, int lddy, const float* dtp, int lddtp, const float* z, const uint8_t* code,
                                      const float* scale, const float* mean, const float* rstd, const double* sums, double count,
                                      int S, int P, int pool, int Lp, const void* x_panel, long x_panel_rows, int fmt, int k,
                                      int Cin, int Cout, const float* gscale, float* dW, double* bias_sums, float* bias_out,
                                      void* ws, size_t ws_bytes, void* stream) {
    DCUE_CHECK_ARG(dy && z && code && x_panel && dW && S >= 0 && P > 0 && lddy >= 128 && lddy % 4 == 0);
    DCUE_CHECK_ARG(((uintptr_t)dy & 15) == 0 && ((uintptr_t)z & 15) == 0 && ((uintptr_t)code & 15) == 0);
    DCUE_CHECK_ARG(!dtp || (lddtp % 4 == 0 && ((uintptr_t)dtp & 15) == 0));
    DCUE_CHECK_ARG(!sums || (mean && rstd && count > 0));
    if (pool != 4 || k != 4 || Cin != 128 || Cout != 128)
        DCUE_FAIL(DCUE_E_UNSUPPORTED, "dcue_conv_wgrad_unpool: built for pool 4, k 4, 128 channels");
    DCUE_CHECK_ARG(Lp % 4 == 0 && P * pool <= Lp && (long)S * Lp < (1L << 31) - 4096);
    DCUE_CHECK_ARG(x_panel_rows >= round_up_l((long)S * Lp, 128) + 16);
    if (S == 0) {
        DCUE_CUDA(cudaMemsetAsync(dW, 0, (size_t)Cout * Cin * k * sizeof(float), (cudaStream_t)stream));
        if (bias_sums) DCUE_CUDA(cudaMemsetAsync(bias_sums, 0, 128 * sizeof(double), (cudaStream_t)stream));
        if (bias_out) DCUE_CUDA(cudaMemsetAsync(bias_out, 0, 128 * sizeof(float), (cudaStream_t)stream));
        return 0;
    }
    return dcue_tc_conv_wgrad_unpool(dy, lddy, dtp, lddtp, z, code, scale, mean, rstd, sums, count, S, P, Lp, x_panel,
                                     x_panel_rows, fmt, k, gscale, dW, bias_sums, bias_out, ws, ws_bytes, (cudaStream_t)stream);
}
