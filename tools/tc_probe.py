"""GPU probe: run one tcgen05 backward configuration per subprocess (a trapped kernel poisons the
CUDA context) and report max relative error vs the CUDA-core validator.
usage: python tools/tc_probe.py            # runs the whole matrix
       python tools/tc_probe.py wgrad 1 0  # one case: op fmt_dy fmt_other"""
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(op, fmt_dy, fmt_o, gi=0, S=5):
    import torch
    pkg = importlib.import_module("amplifai-deepcontentrecommenders_b200")
    L, ops = pkg._lib, pkg.ops
    dev = "cuda"
    geo = ops.tower_geometry(131)[gi]
    g = torch.Generator().manual_seed(1)
    st = L.stream()
    x = torch.randn(S, 128, geo["Lin"], generator=g).to(dev)
    w = (torch.randn(128, 128, geo["k"], generator=g) * 0.06).to(dev)
    X = ops.Panel(S, geo["Lp"], dev)
    L.call("dcue_ncl_pack", x.data_ptr(), S, None, 0, 128, geo["Lin"], None, None, X.base, X.panel_rows, geo["Lp"], geo["pad"], fmt_o, st)
    # a dense random dY (valid rows only) written through the unpool kernel with pool = 1 semantics is not
    # available, so pack a random [S, Lout] gradient with the NCL packer (pad = 0) instead
    dy = torch.randn(S, 128, geo["P"] * geo["pool"], generator=g).to(dev)
    dY = ops.Panel(S, geo["Lp"], dev)
    L.call("dcue_ncl_pack", dy.data_ptr(), S, None, 0, 128, geo["P"] * geo["pool"], None, None, dY.base, dY.panel_rows, geo["Lp"], 0, fmt_dy, st)
    nws = L.query("dcue_conv_ws_bytes", L.IMPL_TC, S, geo["Lp"], geo["k"], 128, 128)
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    outs = []
    for impl in (L.IMPL_SIMT, L.IMPL_TC):
        if op == "wgrad":
            o = torch.zeros(128, 128, geo["k"], device=dev)
            L.call("dcue_conv_wgrad", impl, dY.base, dY.panel_rows, fmt_dy, X.base, X.panel_rows, fmt_o, S * geo["Lp"], geo["k"], 128, 128,
                   None, o.data_ptr(), ws.data_ptr(), nws, st)
        else:
            wpd = torch.empty(128 * geo["k"] * 128, dtype=torch.int16, device=dev)
            L.call("dcue_pack_conv_weight", w.data_ptr(), 128, 128, geo["k"], 1, fmt_o, None, None, wpd.data_ptr(), st)
            o = torch.zeros(S * geo["Lin"], 128, device=dev)
            L.call("dcue_conv_dgrad", impl, dY.base, dY.panel_rows, fmt_dy, wpd.data_ptr(), fmt_o, S, geo["Lp"], geo["Lin"], geo["pad"], geo["k"],
                   128, 128, None, o.data_ptr(), ws.data_ptr(), nws, st)
        torch.cuda.synchronize()
        outs.append(o.double().cpu())
    err = ((outs[0] - outs[1]).abs().max() / outs[0].abs().max()).item()
    print("RESULT %s fmt_dy=%d fmt_other=%d gi=%d: max rel err tc vs simt = %.3e (|ref|max %.3e)" % (op, fmt_dy, fmt_o, gi, err, outs[0].abs().max()))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), *(int(a) for a in sys.argv[4:]))
        sys.exit(0)
    for op in ("dgrad", "wgrad"):
        for fd, fo in ((0, 0), (1, 1), (1, 0), (0, 1)):
            for gi in ((0, 3) if (fd, fo) == (0, 0) else (0,)):
                r = subprocess.run([sys.executable, __file__, op, str(fd), str(fo), str(gi)], capture_output=True, text=True, timeout=300)
                lines = [l for l in (r.stdout + r.stderr).splitlines() if "RESULT" in l or "Error" in l or "error" in l]
                print(op, fd, fo, gi, "rc=%d" % r.returncode, "|", (lines[-1] if lines else "no output")[:200], flush=True)
