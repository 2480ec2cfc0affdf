"""Achieved HBM bandwidth of the memory-bound DCUE kernels, each timed alone with CUDA events on its
launch stream at a size that is not launch-latency bound (at the cfg2 batch of 1024 the embedding and loss
kernels move 2-18 MB and finish in 5-25 us, which says nothing about the kernel).  Bytes are the ALGORITHMIC
bytes of SURVEY.md section 8(d); the denominator is MEASURED_PEAKS.json's copy bandwidth.
  python tools/hbm_kernel_bench.py [rows]          # rows = triplets for the embedding / loss kernels"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps=10, flush=None):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()                      # > L2: the next launch starts from a cold cache
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def run(B=262144, tower=True, S=21504):
    """-> {"hbm_peak_GB/s", "rows", "kernels": {name: {ms, GB/s, frac_of_peak, algorithmic_MB}}}; B = triplets (rows) for the
    embedding / loss kernels; tower=True adds the song-tower glue sweeps at S spectrograms."""
    pkg = importlib.import_module("amplifai-deepcontentrecommenders_b200")
    L, ops = pkg._lib, pkg.ops
    dev = torch.device("cuda")
    peak = 6546.2
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = L.stream()
    g = torch.Generator(device=dev).manual_seed(0)
    out = {}

    def report(name, nbytes, ms):
        out[name] = {"ms": round(ms, 4), "GB/s": round(nbytes / ms / 1e6, 1), "frac_of_peak": round(nbytes / ms / 1e6 / peak, 3),
                     "algorithmic_MB": round(nbytes / 1e6, 1)}

    # ---- embedding gather (+ReLU) and the sorted segment scatter-add, U = 1M x 300 (cfg4 table)
    U, E = 1000000, 300
    table = torch.randn(U, E, device=dev, generator=g)
    idx = torch.randint(0, U, (B,), device=dev, generator=g)
    h0 = torch.empty(B, E, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    report("gather_relu_fwd", B * (8 + 2 * 4 * E),
           timed(lambda: L.call("dcue_gather_relu_fwd", table.data_ptr(), idx.data_ptr(), B, U, E, h0.data_ptr(), None, err.data_ptr(), st),
                 flush=flush))
    dh0 = torch.randn(B, E, device=dev, generator=g)
    gtab = torch.zeros(U, E, device=dev)
    if B <= L.lib().dcue_scatter_direct_max():
        # step-sized batch: the sort-free single-launch scatter the training step uses
        report("scatter_add_rows", B * (8 + 3 * 4 * E),
               timed(lambda: L.call("dcue_scatter_add_rows", dh0.data_ptr(), h0.data_ptr(), idx.data_ptr(), B, U, E, gtab.data_ptr(), st),
                     flush=flush))
    else:
        sidx = torch.empty(B, dtype=torch.int64, device=dev)
        spos = torch.empty(B, dtype=torch.int32, device=dev)
        nscr = L.query("dcue_sort_ws_bytes", B)
        scr = torch.empty(nscr, dtype=torch.uint8, device=dev)
        L.call("dcue_sort_indices", idx.data_ptr(), B, U, sidx.data_ptr(), spos.data_ptr(), scr.data_ptr(), nscr, st)
        report("segment_scatter_add_bwd", B * (8 + 4 + 3 * 4 * E),   # grad row + relu mask row read, table row written
               timed(lambda: L.call("dcue_scatter_add_bwd", dh0.data_ptr(), h0.data_ptr(), sidx.data_ptr(), spos.data_ptr(), B, U, E,
                                    gtab.data_ptr(), st), flush=flush))
    del table, gtab, dh0, h0

    # ---- fused cosine score + hinge loss + backward, N = 20 negatives, F = 100
    N, F = 20, 100
    Bs = min(B, 131072)
    uf = torch.randn(Bs, F, device=dev, generator=g)
    feats = torch.randn(Bs * (1 + N), F, device=dev, generator=g)
    scores = torch.empty(Bs, N, device=dev)
    lrows = torch.empty(Bs, device=dev)
    du, df = torch.empty_like(uf), torch.empty_like(feats)
    report("score_hinge_fwdbwd", Bs * (4 * F * (2 + N) * 2 + 4 * N),
           timed(lambda: L.call("dcue_score_hinge_fwdbwd", uf.data_ptr(), feats.data_ptr(), Bs, N, F, 1e-8, 0.2, Bs, scores.data_ptr(),
                                lrows.data_ptr(), du.data_ptr(), df.data_ptr(), st), flush=flush))
    del feats, df

    if not tower:
        return {"hbm_peak_GB/s": peak, "rows": B, "kernels": out}
    # ---- song-tower glue at cfg2 size (S = 21 504 spectrograms)
    geo = ops.tower_geometry(131)[0]
    pos = torch.randn(S, 128, 131, device=dev, generator=g)
    X = ops.Panel(S, geo["Lp"], dev)
    sums = torch.zeros(256, dtype=torch.float64, device=dev)
    rm = torch.zeros(128, device=dev)
    nws = max(L.query("dcue_ncl_stats_ws_bytes", 128), L.query("dcue_bn_bwd_ws_bytes", 128))
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    report("ncl_center_pack_stats", S * (128 * 131 * 4 + geo["Lp"] * 256),
           timed(lambda: L.call("dcue_ncl_center_pack_stats", pos.data_ptr(), S, None, 0, 128, 131, rm.data_ptr(), X.base, X.panel_rows,
                                geo["Lp"], geo["pad"], L.FMT_F16, sums.data_ptr(), ws.data_ptr(), nws, st)))
    del pos
    rows = S * geo["P"]
    z = torch.rand(rows, 128, device=dev, generator=g)
    dy = torch.randn(rows, 128, device=dev, generator=g)
    code = torch.randint(0, 4, (rows, 128), dtype=torch.uint8, device=dev, generator=g)
    bn = torch.ones(4, 128, device=dev)
    gsc = torch.ones(2, device=dev)
    bsum = torch.zeros(128, dtype=torch.float64, device=dev)
    dY = ops.Panel(S, geo["Lp"], dev)
    report("bn_relu_unpool_bwd(layer1)", rows * (512 + 512 + 128) + S * geo["P"] * 4 * 256,
           timed(lambda: L.call("dcue_bn_relu_unpool_bwd", dy.data_ptr(), 128, None, 0, z.data_ptr(), code.data_ptr(), bn[0].data_ptr(),
                                bn[2].data_ptr(), bn[3].data_ptr(), sums.data_ptr(), float(rows), S, geo["P"], 128, 4, geo["Lp"],
                                dY.base, dY.panel_rows, L.FMT_F16, gsc.data_ptr(), None, bsum.data_ptr(), None, ws.data_ptr(), nws, st)))
    g2 = ops.tower_geometry(131)[1]
    X2 = ops.Panel(S, g2["Lp"], dev)
    report("affine_pack(layer1->2)", rows * (512 + 256),
           timed(lambda: L.call("dcue_affine_pack", z.data_ptr(), S, geo["P"], 128, bn[0].data_ptr(), bn[1].data_ptr(), X2.base,
                                X2.panel_rows, g2["Lp"], g2["pad"], L.FMT_F16, None, None, 0, st)))
    amax = torch.zeros(1, device=dev)
    report("bn_bwd_reduce(layer1)", rows * 1024,
           timed(lambda: L.call("dcue_bn_bwd_reduce", dy.data_ptr(), 128, None, 0, z.data_ptr(), bn[2].data_ptr(), bn[3].data_ptr(), S,
                                geo["P"], 128, sums.data_ptr(), amax.data_ptr(), None, None, ws.data_ptr(), nws, st), flush=flush))
    return {"hbm_peak_GB/s": peak, "rows": B, "kernels": out}


if __name__ == "__main__":
    print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 262144), indent=1))
