"""DCUE benchmark (BASELINE.json metric: DCUE train triplets/sec at 1/2/4/8 B200; eval scored users/sec).

  python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the box's host cores

One "step" = zero_grad + DCUENet forward(u, pos, neg) + hinge loss + backward + optimizer.step +
scheduler.batch_step over one batch (the loop body of the reference's _train_epoch,
dcrecommend/nn/dcue.py:202-210).  Workload = BASELINE configs[1] (cfg2): truedcuemel1dbn, 20k users,
emb 300, batch 1024 per GPU, 20 sampled negatives, margin 0.2, Adam(1e-5, (0.9, 0.99), 1e-8);
at N > 1 this is cfg3 (data parallel, global batch 1024 N).  Prints ONE JSON line on rank 0.  Besides the
contract's keys the line carries
  e2e / e2e_indexed      dense feed through forward(u,pos,neg)-compatible host tensors / index feed
  operand_bf16           the same step with bf16 conv operands (DCUE_OPERAND=bf16), BASELINE's "bf16 tower"
  eval                   cfg5: ALL 1M users x 500k songs, fused top-100, song-sharded + all-to-all by user block
  cfg4                   1M-user table row-sharded over the ranks (NVLink peer memory), negatives 20 -> 200
  roofline               dominant kernel (CUDA events, live) + whole-step tensor fraction + HBM kernels
  dp_parity / eval_parity / table_parity   (N > 1) multi-rank results against rank 0 recomputing them on one GPU
  cpu_baseline, eval.cpu_baseline          the oracle port of the reference on the host cores (N = 1)
"""
import argparse
import glob
import importlib
import json
import os
import gc
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC, UNIT = "DCUE train triplets/sec", "triplets/s"
CFG = dict(model_type="truedcuemel1dbn", users=20000, emb=300, feat=100, hidden=128, frames=131, margin=0.2,
           lr=1e-5, betas=(0.9, 0.99), eps=1e-8)
L1_FLOP_PER_SPEC = 2 * 128 * 128 * 4 * 132      # live MACs*2 of layer1 per spectrogram
STEP_FLOP_PER_SPEC_LIVE = 68.16e6               # forward + dgrad + wgrad, live output columns only
USER_MLP_FLOP_PER_TRIPLET = 0.72e6


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(burst=d.get("bf16_tflops", 1590.0), sustained=d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)),
                    hbm=d.get("hbm_gbs", 6650.0), source="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def host_threads():
    """All host cores this process may use -- torchrun exports OMP_NUM_THREADS=1, which made the round-1 CPU arm run on one
    thread at N > 1 and voided those ratios."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


_SAMPLER_SRC = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
print(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM), flush=True)
get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
import select
while True:
    if select.select([sys.stdin], [], [], 0.004)[0]:
        break
    print(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), get(h), flush=True)
"""


class ClockSampler:
    """SM clock and throttle reasons sampled every ~4 ms DURING the timed region by a SEPARATE PROCESS (NVML).  A sampling
    thread inside the benchmark process competed for the GIL with the thread that launches the step; with the data-parallel
    step's bounded launch queue that cost 0.5 ms per step at N >= 2 (rounds 1-2: 635 k vs 758 k triplets/s on 2 GPUs)."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        import subprocess
        self.maxclk, self.err, self.proc = None, None, None
        try:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(phys)], stdin=subprocess.PIPE,
                                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            first = self.proc.stdout.readline()            # NVML is initialised: sampling has started
            self.maxclk = int(first.strip())
        except Exception as e:  # noqa: BLE001
            self.err = "nvml sampler unavailable: %s" % e
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": self.maxclk, "reasons": [self.err or "no samples"]}
        try:
            out, _ = self.proc.communicate(input="stop\n", timeout=5.0)
        except Exception as e:  # noqa: BLE001
            self.proc.kill()
            return {"sm_mhz": None, "sm_max_mhz": self.maxclk, "reasons": ["sampler: %s" % e]}
        samples = []
        for line in out.splitlines():
            parts = line.split()
            if len(parts) == 2:
                samples.append((int(parts[0]), int(parts[1])))
        if not samples:
            return {"sm_mhz": None, "sm_max_mhz": self.maxclk, "reasons": ["no samples"]}
        clk = sorted(c for c, _ in samples)
        reasons = [n for n, bit in self.BAD.items() if any(r & bit for _, r in samples)]
        return {"sm_mhz": clk[len(clk) // 2], "sm_max_mhz": self.maxclk, "reasons": reasons, "samples": len(samples),
                "sampler": "separate process, NVML every 4 ms"}


def workload_config(args):
    return {"workload": "cfg2: DCUE %s, %d users x %d emb, feature %d, batch %d per GPU, 1 pos + %d sampled negs, hinge margin %.1f, Adam lr 1e-5"
                        % (CFG["model_type"], CFG["users"], CFG["emb"], CFG["feat"], args.batch, args.negs, CFG["margin"]),
            "global_batch": args.batch * args.gpus, "negatives": args.negs, "parallelism": "dp%d" % args.gpus,
            "l2": "inputs (%.2f GB/step/GPU fp32 spectrograms) exceed the 126 MB L2; no flush needed" %
                  (args.batch * (1 + args.negs) * 128 * CFG["frames"] * 4 / 1e9)}


# ------------------------------------------------------------------------------------- CPU arm (oracle port)
class OracleTrainer:
    """The reference algorithm (oracle/dcue_oracle.py, pinned to the reference's own outputs) + torch Adam on host cores."""

    def __init__(self, users):
        from oracle import fixtures
        self.p = fixtures.make_params(CFG["model_type"], seed=0, user_count=users)
        self.names = [k for k, v in self.p.items() if v.is_floating_point() and "running_" not in k]
        self.leaves = [torch.nn.Parameter(self.p[k].clone()) for k in self.names]
        self.opt = torch.optim.Adam(self.leaves, CFG["lr"], CFG["betas"], CFG["eps"], 0)

    def step(self, u, pos, neg):
        from oracle import dcue_oracle as O
        cur = dict(self.p)
        cur.update({k: l.detach() for k, l in zip(self.names, self.leaves)})
        r = O.train_step_grads(cur, u, pos, neg, CFG["model_type"], CFG["margin"])
        for l, k in zip(self.leaves, self.names):
            l.grad = r["grads"].get(k)
        self.opt.step()
        self.p.update(r["new_stats"])
        return r["loss"].item()


def cpu_arm(batch, negs, steps, warmup, budget_s=None):
    """-> (triplets/s, seconds per step, steps run, threads)."""
    from oracle import fixtures
    threads = host_threads()
    tr = OracleTrainer(CFG["users"])
    u, pos, neg = fixtures.make_inputs(batch, negs, CFG["users"], seed=1)
    for _ in range(warmup):
        tr.step(u, pos, neg)
    t0, n = time.perf_counter(), 0
    while n < steps:
        tr.step(u, pos, neg)
        n += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s and n >= 2:
            break
    dt = time.perf_counter() - t0
    return batch * n / dt, dt / n, n, threads


def cpu_eval_arm(n_users=1024, n_items=500000, k=100):
    """reference-style model.sim on all pairs for a 1 024-user slice x 500k songs (chunked) + top-k,
    on the host cores -> users/s (oracle.topk_scores restates dcrecommend/nn/dcue.py:513 for all pairs)."""
    from oracle import dcue_oracle as O
    threads = host_threads()
    g = torch.Generator().manual_seed(3)
    uf = torch.randn(n_users, CFG["feat"], generator=g)
    itf = torch.randn(n_items, CFG["feat"], generator=g)
    O.topk_scores(uf[:64], itf, k, chunk=64)
    t0 = time.perf_counter()
    O.topk_scores(uf, itf, k, chunk=256)
    dt = time.perf_counter() - t0
    return {"value": n_users / dt, "unit": "users/s", "cores": threads, "kind": "port",
            "sample": "%d-user slice x %d songs, fp32 normalise + matmul + torch.topk(k=%d), chunks of 256 users, %.2f s" %
                      (n_users, n_items, k, dt)}


def run_reference(args):
    """Reference arm: the reference's algorithm (oracle port: the reference itself is pure Python on torch and cannot travel to
    the GPU box; the port is pinned to its outputs by tests/golden) on ALL host cores, same config / metric / unit as ours.
    Each step is one full cfg2 batch when that fits the time budget, else the largest power-of-two slice that does."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    # probe at batch 64 to size the sample: K + W steps must end within ~4 minutes
    v64, s64, _, _ = cpu_arm(64, args.negs, 1, 1)
    per_triplet = s64 / 64.0
    total_steps = args.steps + min(args.warmup, 2)
    sample_b = args.batch
    while sample_b > 64 and per_triplet * sample_b * total_steps > 240.0:
        sample_b //= 2
    while True:
        try:
            value, sps, n, _ = cpu_arm(sample_b, args.negs, args.steps, min(args.warmup, 2))
            break
        except (RuntimeError, MemoryError) as exc:          # host RAM too small for the full-batch autograd graph
            if sample_b <= 64 or "alloc" not in str(exc).lower():
                raise
            sample_b //= 2
    sample = ("oracle port of the reference train step (dcrecommend/nn/dcue.py:202-210) on %d host threads; each step = %d of the "
              "%d triplets of a cfg2 batch x %d negs (BatchNorm statistics over that slice), %d steps, %.2f s/step"
              % (threads, sample_b, args.batch, args.negs, n, sps))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
           "warmup": min(args.warmup, 2), "ms_per_step": sps * 1e3 * args.batch / sample_b, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------- GPU arm helpers
def load_ncu_traffic():
    """DRAM bytes per launch of the big kernels, read from the committed ncu summaries (profiles/*traffic*.json: the newest
    file wins).  Values are at S = 21 504 spectrograms and scale linearly with S."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*traffic*.json")))
    if not files:
        return {}, None
    d = json.load(open(files[-1]))
    return d.get("kernels", {}), os.path.basename(files[-1])


def kernel_roofline(pkg, S, dev):
    """CUDA-event timing, on this stream, of the step's heaviest kernels at layer-1 size: the tcgen05 GEMMs (tensor-bound)
    and the BatchNorm-backward/unpool sweep (HBM-bound).  achieved = algorithmic FLOPs (bytes) per launch / average launch time."""
    L, ops = pkg._lib, pkg.ops
    geo = ops.tower_geometry(CFG["frames"])[0]
    st = L.stream()
    X, dY = ops.Panel(S, geo["Lp"], dev), ops.Panel(S, geo["Lp"], dev)
    X.buf.random_(0, 15000)   # arbitrary finite positive fp16 bit patterns (values < 1)
    dY.buf.random_(0, 15000)
    wp = torch.randint(0, 15000, (128 * 4 * 128,), dtype=torch.int16, device=dev)
    bias = torch.zeros(128, device=dev)
    rows = S * geo["P"]
    z = torch.rand(rows, 128, device=dev)
    dyn = torch.randn(rows, 128, device=dev)
    code = torch.randint(0, 4, (rows, 128), dtype=torch.uint8, device=dev)
    sums = torch.zeros(256, dtype=torch.float64, device=dev)
    bsum = torch.zeros(128, dtype=torch.float64, device=dev)
    bn = torch.ones(4, 128, device=dev)
    gsc = torch.ones(2, device=dev)
    dW = torch.empty(128, 128, 4, device=dev)
    nws = max(L.query("dcue_conv_ws_bytes", L.IMPL_TC, S, geo["Lp"], 4, 128, 128), L.query("dcue_bn_bwd_ws_bytes", 128),
              L.query("dcue_conv_wgrad_unpool_ws_bytes", 4))
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    calls = {
        "conv1_fwd_pool": (lambda: L.call("dcue_conv_pool_fwd", L.IMPL_TC, X.base, X.panel_rows, 0, wp.data_ptr(), bias.data_ptr(), None, S,
                                          geo["Lp"], geo["Lin"], 2, geo["P"], 4, 4, 128, 128, z.data_ptr(), code.data_ptr(),
                                          sums.data_ptr(), ws.data_ptr(), nws, st), "tensor", L1_FLOP_PER_SPEC * S),
        "conv1_wgrad": (lambda: L.call("dcue_conv_wgrad", L.IMPL_TC, dY.base, dY.panel_rows, 0, X.base, X.panel_rows, 0, S * geo["Lp"], 4,
                                       128, 128, None, dW.data_ptr(), ws.data_ptr(), nws, st), "tensor", L1_FLOP_PER_SPEC * S),
        # fused layer-1 backward (the step's longest kernel): BatchNorm-backward + unpool built in smem + weight gradient
        "conv1_wgrad_unpool": (lambda: L.call("dcue_conv_wgrad_unpool", dyn.data_ptr(), 128, None, 0, z.data_ptr(), code.data_ptr(),
                                              bn[0].data_ptr(), bn[2].data_ptr(), bn[3].data_ptr(), sums.data_ptr(), float(rows), S,
                                              geo["P"], 4, geo["Lp"], X.base, X.panel_rows, 0, 4, 128, 128, gsc.data_ptr(),
                                              dW.data_ptr(), bsum.data_ptr(), None, ws.data_ptr(), nws, st), "tensor", L1_FLOP_PER_SPEC * S),
        # reads dy, z (fp32) and the argmax code, writes the 4x unpooled fp16 panel: 1152 + 1024 B per pooled row
        "bn_relu_unpool_bwd1": (lambda: L.call("dcue_bn_relu_unpool_bwd", dyn.data_ptr(), 128, None, 0, z.data_ptr(), code.data_ptr(),
                                               bn[0].data_ptr(), bn[2].data_ptr(), bn[3].data_ptr(), sums.data_ptr(), float(rows), S,
                                               geo["P"], 128, 4, geo["Lp"], dY.base, dY.panel_rows, 0, gsc.data_ptr(), None,
                                               bsum.data_ptr(), None, ws.data_ptr(), nws, st), "hbm", (1152 + 1024) * rows),
    }
    res = {}
    for name, (fn, bound, work) in calls.items():
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        reps = 10
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[name] = {"ms": ms, "bound": bound, "achieved": work / (ms * 1e-3) / (1e12 if bound == "tensor" else 1e9),
                     "unit": "TFLOP/s" if bound == "tensor" else "GB/s"}
    return res


def build_model(pkg, users, dev, seed=0):
    torch.manual_seed(seed)
    return pkg.DCUENet({"feature_dim": CFG["feat"], "conv_hidden": CFG["hidden"], "user_embdim": CFG["emb"], "user_count": users,
                        "model_type": CFG["model_type"]}).to(dev).train()


def l2rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def timed_steps(fn, n, barrier, dev, world):
    """n calls of fn bracketed by barrier + synchronize, CUDA events, MAX over ranks -> total ms."""
    import torch.distributed as dist
    gc.collect()
    gc.disable()          # no collector pause of the launching process inside the timed region
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    barrier()
    gc.enable()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


# ------------------------------------------------------------------------------------- multi-rank parity legs
def timeline_leg(step, u, pos, neg, path, rank, steps=4):
    """Kernel timeline of the (graph-replayed) step from CUPTI activity records (torch.profiler): per kernel its stream, start
    offset, duration and the idle gap before it on its stream -- the in-graph, warm-cache view that ncu's serialised launch
    list cannot give.  One text file per rank: `path`.rank<r>.txt (diagnostic; never part of a timed region)."""
    import json as _json
    import tempfile
    from torch.profiler import ProfilerActivity, profile
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            step(u, pos, neg)
        torch.cuda.synchronize()
    with tempfile.NamedTemporaryFile(suffix=".json") as f:
        prof.export_chrome_trace(f.name)
        tr = _json.load(open(f.name))
    ks = sorted((e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e),
                key=lambda e: e["ts"])
    # a step starts at the single-pass input kernel (the first launch of the forward); older layouts: the user-tower gather
    starts = [i for i, e in enumerate(ks) if "center_pack_stats" in e["name"]]
    if len(starts) < 3:
        starts = [i for i, e in enumerate(ks) if "gather_relu" in e["name"]]
    if len(starts) < 3:
        return None
    a, b = starts[-2], starts[-1]
    evs = ks[a:b]
    t0 = evs[0]["ts"]
    last_end = {}
    lines, busy = [], {}
    for e in evs:
        st = e.get("args", {}).get("stream", 0)
        gap = e["ts"] - last_end[st] if st in last_end else 0.0
        last_end[st] = e["ts"] + e["dur"]
        busy[st] = busy.get(st, 0.0) + e["dur"]
        name = e["name"].replace("(anonymous namespace)::", "").replace("void ", "")
        lines.append("%9.1f %8.1f %7.1f  s%-3s %s" % (e["ts"] - t0, e["dur"], gap, st, name[:100]))
    span = ks[b]["ts"] - t0
    agg = {}
    for e in evs:
        n = e["name"].replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:70]
        c = agg.setdefault(n, [0, 0.0])
        c[0] += 1
        c[1] += e["dur"]
    out = "%s.rank%d.txt" % (path, rank)
    with open(out, "w") as f:
        f.write("step span %.1f us, %d kernels; busy per stream: %s\n" % (span, len(evs), {k: round(v, 1) for k, v in busy.items()}))
        f.write("--- per kernel (count, total us, share of span)\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-72s n=%3d %9.1f us %5.1f%%\n" % (n, c, t, 100 * t / span))
        f.write("--- timeline: start_us dur_us gap_before_us stream kernel\n")
        f.write("\n".join(lines) + "\n")
    return {"span_us": span, "kernels": len(evs), "file": out}


def dp_parity_leg(pkg, par, dev, rank, world, per_rank=32, negs=4, users=500):
    """Global batch split over the ranks (SyncBN statistics, peer-memory table-row exchange, flat gradient all-reduce) against
    rank 0 recomputing the FULL batch on one GPU with the same kernels: loss, every gradient, BatchNorm buffers; then the
    CUDA-graph replay of the data-parallel step against the eager one."""
    import torch.distributed as dist
    Bg = per_rank * world
    g = torch.Generator().manual_seed(77)
    u = torch.randint(0, users, (Bg,), generator=g)
    u[1] = u[0]
    u[Bg - 1] = u[0]                                   # the same user on the first and the last rank
    pos = torch.randn(Bg, 128, CFG["frames"], generator=g)
    neg = torch.randn(Bg, negs, 128, CFG["frames"], generator=g)
    lo, hi = par.shard_slice(Bg, rank, world)
    model = build_model(pkg, users, dev, seed=5)
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    dp = par.DataParallelDCUE(model)
    ul, pl, nl = u[lo:hi].to(dev), pos[lo:hi].to(dev), neg[lo:hi].to(dev)
    model.zero_grad(set_to_none=True)
    loss = dp.loss_step(ul, pl, nl, CFG["margin"])
    loss.backward()
    dp.reduce_gradients()
    loss_g = dp.reduce_loss(loss).item()
    del loss       # no autograd graph (and no AccumulateGrad node bound to this stream) may survive into the graph capture below
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    bufs = {k: v.detach().clone() for k, v in model.named_buffers()}
    # every rank must hold the same reduced gradients: compare with rank 0's
    worst_rank_diff = 0.0
    for k in sorted(grads):
        ref = grads[k].clone()
        dist.broadcast(ref, 0)
        worst_rank_diff = max(worst_rank_diff, (grads[k] - ref).abs().max().item())
    wr = torch.tensor([worst_rank_diff], device=dev)
    dist.all_reduce(wr, op=dist.ReduceOp.MAX)
    # graph replay of the DP step
    model.load_state_dict(sd0)
    gstep = pkg.GraphedTrainStep(model, CFG["margin"], ul, pl, nl, warmup=2, dp=dp)
    gl = gstep(ul, pl, nl)
    graph_loss = dp.reduce_loss(gl).item()
    graph_grad_diff = max((p.grad - grads[k]).abs().max().item() / grads[k].abs().max().clamp_min(1e-30).item()
                          for k, p in model.named_parameters())
    gd = torch.tensor([graph_grad_diff], device=dev)
    dist.all_reduce(gd, op=dist.ReduceOp.MAX)
    gstep.release()
    dp.check_peers()
    out = None
    if rank == 0:
        single = build_model(pkg, users, dev, seed=5)
        single.load_state_dict(sd0)
        single.zero_grad(set_to_none=True)
        ls = single.hinge_loss_step(u.to(dev), pos.to(dev), neg.to(dev), CFG["margin"])
        ls.backward()
        worst, worst_name = 0.0, ""
        for k, p in single.named_parameters():
            e = l2rel(grads[k], p.grad)
            if e > worst:
                worst, worst_name = e, k
        sb = dict(single.named_buffers())
        buf_err = max(((bufs[k].double() - sb[k].double()).abs().max() / sb[k].double().abs().max().clamp_min(1e-30)).item()
                      for k in bufs)
        out = {"global_batch": Bg, "negatives": negs, "loss_rel": abs(loss_g - ls.item()) / abs(ls.item()),
               "worst_grad_l2": worst, "worst_grad": worst_name, "buffers": buf_err, "max_abs_diff_between_ranks": wr.item(),
               "graph_vs_eager_loss_rel": abs(graph_loss - loss_g) / abs(loss_g), "graph_vs_eager_grad_relmax": gd.item(),
               "what": "ranks' slices vs rank 0 recomputing the full batch on one GPU (same kernels, fp16 operands)"}
    model.conv._dp = None
    model.user_embd._dp = None
    dist.barrier()
    return out


def eval_parity_leg(par, ev, dev, rank, world, n_users=3000, n_items=None, k=100):
    """sharded_topk (songs over the ranks, all-to-all by user block, per-block merge) against rank 0 scoring all songs alone."""
    import torch.distributed as dist
    n_items = n_items or 40000 * world      # >= 32768 songs per shard: the global-threshold protocol is exercised
    g = torch.Generator(device=dev).manual_seed(11)
    uf = torch.randn(n_users, CFG["feat"], generator=g, device=dev)
    itf = torch.randn(n_items, CFG["feat"], generator=g, device=dev)
    lo, hi = par.shard_slice(n_items, rank, world)
    s, i = par.sharded_topk(uf, itf[lo:hi].contiguous(), k, lo, gather=True, user_tile=n_users // 3)
    out = None
    if rank == 0:
        s1, i1 = ev.topk_scores(uf, itf, k)
        same = (i == i1)
        # a differing index is only legitimate between tied scores
        bad = int(((~same) & ((s - s1).abs() > 0)).sum())
        out = {"users": n_users, "songs": n_items, "k": k, "scores_max_abs_diff": (s - s1).abs().max().item(),
               "index_mismatch_beyond_ties": bad, "index_equal_frac": same.float().mean().item(),
               "what": "song-sharded top-k over %d ranks (3 user tiles) vs one GPU scoring all songs" % world}
    dist.barrier()
    return out


def table_parity_leg(pkg, par, optim, dev, rank, world, users=1003, per_rank=48, negs=3):
    """Row-sharded user table over NVLink peer memory against the replicated table (both data parallel, same batch):
    loss, user-MLP gradients, the owner's shard gradient vs the same rows of the dense gradient, and the table after one Adam step."""
    import torch.distributed as dist
    Bg = per_rank * world
    g = torch.Generator().manual_seed(99)
    u = torch.randint(0, users, (Bg,), generator=g)
    u[1] = u[0]
    u[Bg - 1] = u[0]
    pos = torch.randn(Bg, 128, CFG["frames"], generator=g)
    neg = torch.randn(Bg, negs, 128, CFG["frames"], generator=g)
    lo_b, hi_b = par.shard_slice(Bg, rank, world)
    ul, pl, nl = u[lo_b:hi_b].to(dev), pos[lo_b:hi_b].to(dev), neg[lo_b:hi_b].to(dev)

    def one(sharded):
        model = build_model(pkg, users, dev, seed=9)
        if sharded:
            par.shard_user_table(model)
        dp = par.DataParallelDCUE(model)
        opt = optim.FusedAdam(model.parameters(), 1e-3, CFG["betas"], CFG["eps"], 0)
        model.zero_grad(set_to_none=True)
        loss = dp.loss_step(ul, pl, nl, CFG["margin"])
        loss.backward()
        dp.reduce_gradients()
        lg = dp.reduce_loss(loss).item()
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        opt.step()
        table = model.user_embd.gather_full_table() if sharded else model.user_embd.embeddings.weight.detach().clone()
        span = (model.user_embd.lo, model.user_embd.hi) if sharded else None
        model.raise_if_index_error()
        model.conv._dp = None
        return lg, grads, table, span

    l_rep, g_rep, t_rep, _ = one(False)
    l_sh, g_sh, t_sh, (lo, hi) = one(True)
    errs = torch.tensor([abs(l_sh - l_rep) / abs(l_rep),
                         max(l2rel(g_sh[k], g_rep[k]) for k in g_sh if k in g_rep),
                         (g_sh["user_embd.shard"] - g_rep["user_embd.embeddings.weight"][lo:hi]).abs().max().item(),
                         (t_sh - t_rep).abs().max().item()], device=dev, dtype=torch.float64)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank != 0:
        return None
    return {"users": users, "global_batch": Bg, "loss_rel": errs[0].item(), "worst_shared_grad_l2": errs[1].item(),
            "shard_grad_max_abs_diff": errs[2].item(), "table_after_adam_max_abs_diff": errs[3].item(),
            "what": "ShardedUserTable(transport=peer) vs replicated table, both data parallel over %d ranks" % world}


# ------------------------------------------------------------------------------------- cfg4 leg
def cfg4_leg(pkg, par, optim, dev, rank, world, args, barrier):
    """BASELINE configs[3]: 1M users x 300 row-sharded over the ranks, index feed, negatives swept 20 -> 200.  Per N: whole-job
    triplets/s of the full step (graph: forward + loss + backward [+ collectives]; then the fused Adam), the time of the dense
    Adam pass alone and of the user-table path alone (gather over NVLink + MLP forward/backward + row exchange + scatter)."""
    U4, B = 1000000, args.batch
    model = build_model(pkg, U4, dev, seed=4)
    if world > 1:
        par.shard_user_table(model, capacity=B)
    dp = par.DataParallelDCUE(model)
    opt = optim.FusedAdam(model.parameters(), CFG["lr"], CFG["betas"], CFG["eps"], 0)
    opt.set_skip_flags(model.error_flags())
    pool_songs = 2048
    g = torch.Generator(device=dev).manual_seed(40 + rank)
    pool = torch.randn(pool_songs, 128, CFG["frames"], generator=g, device=dev)
    res = {"users": U4, "emb": CFG["emb"], "batch_per_gpu": B, "table": ("row-sharded over %d ranks, NVLink peer-memory gather / "
           "gradient-row exchange (csrc/peer.cu)" % world) if world > 1 else "one GPU holds the whole table",
           "table_rows_per_rank": (U4 + world - 1) // world, "sweep": []}
    for N in args.cfg4_negs:
        u = torch.randint(0, U4, (B,), generator=g, device=dev)
        pi = torch.randint(0, pool_songs, (B,), generator=g, device=dev)
        ni = torch.randint(0, pool_songs, (B, N), generator=g, device=dev)
        gstep = pkg.GraphedTrainStep(model, CFG["margin"], u, pi, ni, pool=pool, warmup=2, dp=dp if world > 1 else None)

        def step():
            gstep()
            opt.step()

        for _ in range(2):
            step()
        n = args.cfg4_steps
        ms = timed_steps(step, n, barrier, dev, world) / n
        adam_ms = timed_steps(opt.step, n, barrier, dev, world) / n

        def table_path():
            out = model.user_embd(u)
            out.backward(torch.ones_like(out))

        model.zero_grad(set_to_none=True)
        for _ in range(2):
            table_path()
        table_ms = timed_steps(table_path, n, barrier, dev, world) / n
        model.raise_if_index_error()
        gstep.release()
        del gstep
        pkg.ops.clear_workspaces()
        torch.cuda.empty_cache()
        # restore persistent gradient storage for the next capture
        model.zero_grad(set_to_none=True)
        res["sweep"].append({"negatives": N, "value": B * world / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                             "adam_ms": adam_ms, "adam_share": adam_ms / ms, "table_path_ms": table_ms,
                             "table_path_share": table_ms / ms,
                             "tflops_live_per_gpu": (B * (1 + N) * STEP_FLOP_PER_SPEC_LIVE) / (ms * 1e-3) / 1e12})
    model.conv._dp = None
    del model, opt, pool
    pkg.ops.clear_workspaces()
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module("amplifai-deepcontentrecommenders_b200")
    par = importlib.import_module("amplifai-deepcontentrecommenders_b200.parallel")
    optim = importlib.import_module("amplifai-deepcontentrecommenders_b200.optim")
    ev = importlib.import_module("amplifai-deepcontentrecommenders_b200.eval")
    L = pkg._lib
    B, N, U = args.batch, args.negs, CFG["users"]
    graphs = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_trainer(seed=0):
        model = build_model(pkg, U, dev, seed)
        dp = par.DataParallelDCUE(model)
        opt = optim.FusedAdam(model.parameters(), CFG["lr"], CFG["betas"], CFG["eps"], 0)
        opt.set_skip_flags(model.error_flags())
        sched = optim.CyclicLRWithRestarts(opt, B * world, epoch_size=B * world * 100000, restart_period=30, t_mult=2, policy="cosine")
        sched.step()
        return model, dp, opt, sched

    model, dp, opt, sched = make_trainer()
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    u = torch.randint(0, U, (B,), generator=g, device=dev)
    pos = torch.randn(B, 128, CFG["frames"], generator=g, device=dev)
    neg = torch.randn(B, N, 128, CFG["frames"], generator=g, device=dev)
    loss_acc = torch.zeros((), device=dev)

    def step_eager(u_, pos_, neg_):
        opt.zero_grad(set_to_none=False)
        loss = dp.loss_step(u_, pos_, neg_, CFG["margin"])
        loss.backward()
        dp.reduce_gradients()
        opt.step()
        sched.batch_step()
        return loss.detach()

    use_graph = not args.no_graph      # under DP the collectives are captured into the graph as well
    graph_launches = 0
    step = step_eager
    if use_graph:
        c0 = L.lib().dcue_launch_count()
        gstep = pkg.GraphedTrainStep(model, CFG["margin"], u, pos, neg, warmup=3, dp=dp if world > 1 else None)
        graphs.append(gstep)
        graph_launches = (L.lib().dcue_launch_count() - c0) // gstep.passes   # warm-up passes (+ priming pass) + the captured one

        def step(u_, pos_, neg_):  # noqa: F811  (same step: forward+loss+backward replayed as one CUDA graph)
            if u_ is not u:
                return step_eager(u_, pos_, neg_)      # other buffers (dense e2e double buffering): eager launches
            loss = gstep()
            opt.step()
            sched.batch_step()
            return loss.detach().clone()

    # ---------------- device-resident timing ("value")
    # the last two warm-up steps run AFTER the sampler process has been forked: the first step after the fork measured 7-15 ms
    # (copy-on-write faults of the launching process), which must not land in the timed region
    late = min(2, args.warmup)
    # warm-up runs the timed loop's exact statement: with graph-owned gradients no ATen add had been launched before the first
    # `loss_acc +=`, and the lazy load of that kernel's module cost the first timed step 10-17 ms
    for _ in range(args.warmup - late + int(os.environ.get("DCUE_BENCH_SETTLE", "0"))):
        loss_acc += step(u, pos, neg)
    # the sampler process takes ~0.1 s to start (python + NVML init): start it BEFORE the barrier, or the other ranks enter
    # the timed loop first and spend that time spinning on rank 0's peer flags inside their own timed region
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(late):
        loss_acc += step(u, pos, neg)
    loss_acc.zero_()
    # no collector pause of the launching process inside the timed region
    gc.collect()
    gc.disable()
    barrier()
    launches0 = L.lib().dcue_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per_step = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)] if os.environ.get("DCUE_BENCH_PER_STEP") else None
    e0.record()
    for i in range(args.steps):
        loss_acc += step(u, pos, neg)
        if per_step:
            per_step[i].record()
    e1.record()
    barrier()
    gc.enable()
    if per_step and rank == 0:      # diagnostic: device time of every step of the timed region
        ts = [e0.elapsed_time(e) for e in per_step]
        print("per-step ms:", " ".join("%.3f" % (b - a) for a, b in zip([0.0] + ts[:-1], ts)), file=sys.stderr, flush=True)
    launches = L.lib().dcue_launch_count() - launches0
    # under graph replay the library's launch counter only ticks for the optimizer: add the per-step count seen at capture
    launches_per_step = launches / args.steps + (graph_launches if use_graph else 0)
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = ms.item()
    value = B * world * args.steps / (ms_total * 1e-3)
    final_loss = dp.reduce_loss(loss_acc / args.steps).item()
    dp.check_peers()

    if args.timeline:
        timeline_leg(step, u, pos, neg, args.timeline, rank)
        barrier()

    # ---------------- end to end: pinned host inputs -> H2D -> step -> loss D2H, every step
    hu, hpos, hneg = (t.cpu().pin_memory() for t in (u, pos, neg))
    bufs = [(torch.empty_like(u), torch.empty_like(pos), torch.empty_like(neg)) for _ in range(2)]
    copy_stream = torch.cuda.Stream(dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def enqueue_copy(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[slot])
            for d, h in zip(bufs[slot], (hu, hpos, hneg)):
                d.copy_(h, non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_loop(n):
        for s in range(2):
            freed[s].record(torch.cuda.current_stream())
        enqueue_copy(0)
        for i in range(n):
            slot = i & 1
            if i + 1 < n:
                enqueue_copy(slot ^ 1)          # overlaps with this step's kernels
            torch.cuda.current_stream().wait_event(ready[slot])
            lossv = step(*bufs[slot])
            freed[slot].record(torch.cuda.current_stream())
            lossv.item()                        # device -> host read of the step's result

    e2e_steps = max(2, min(args.steps, 6))
    e2e_loop(2)
    barrier()
    t0 = time.perf_counter()
    e2e_loop(e2e_steps)
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = B * world * e2e_steps / dt.item()
    h2d = u.numel() * 8 + (pos.numel() + neg.numel()) * 4
    del bufs, hpos, hneg

    # ---------------- end to end on the INDEX feed: the song pool is resident on the
    # device like the user table; a step's host inputs are u, positive / negative song indices only
    pool_songs = 4096
    pool = torch.randn(pool_songs, 128, CFG["frames"], generator=g, device=dev)
    gi = torch.Generator().manual_seed(2 + rank)
    n_idx_batches = 4
    hidx = [(torch.randint(0, U, (B,), generator=gi).pin_memory(), torch.randint(0, pool_songs, (B,), generator=gi).pin_memory(),
             torch.randint(0, pool_songs, (B, N), generator=gi).pin_memory()) for _ in range(n_idx_batches)]
    gidx = None
    if use_graph:
        u0_, p0_, n0_ = (t.to(dev) for t in hidx[0])
        gidx = pkg.GraphedTrainStep(model, CFG["margin"], u0_, p0_, n0_, pool=pool, dp=dp if world > 1 else None)
        graphs.append(gidx)

    def idx_step(i):
        hu_, hp_, hn_ = hidx[i % n_idx_batches]
        if gidx is not None:
            loss = gidx(hu_, hp_, hn_)           # pinned host -> static device index buffers, then one graph launch
        else:
            u_, p_, n_ = hu_.to(dev, non_blocking=True), hp_.to(dev, non_blocking=True), hn_.to(dev, non_blocking=True)
            opt.zero_grad(set_to_none=False)
            loss = dp.loss_step_indexed(u_, pool, p_, n_, CFG["margin"])
            loss.backward()
            dp.reduce_gradients()
        opt.step()
        sched.batch_step()
        return loss.detach().item()      # device -> host read of the step's result

    for i in range(3):
        idx_step(i)
    barrier()
    idx_steps = args.steps
    t0 = time.perf_counter()
    for i in range(idx_steps):
        idx_step(i)
    barrier()
    dti = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dti, op=dist.ReduceOp.MAX)
    e2e_idx_value = B * world * idx_steps / dti.item()
    model.raise_if_index_error()
    h2d_idx = B * 8 * 2 + B * N * 8

    # ---------------- the same step with bf16 conv operands (BASELINE cfg2 says "bf16 tower"; fp16 is this repo's default
    # because it is 8x closer to fp32 at the same tensor-core rate -- both are measured)
    for gs_ in graphs:
        gs_.release()
    graphs.clear()
    model.conv._dp = None
    del model, dp, opt, sched, pool, gidx
    if use_graph:
        del gstep
    pkg.ops.clear_workspaces()
    torch.cuda.empty_cache()
    bf16 = None
    if not args.no_bf16:
        os.environ["DCUE_OPERAND"] = "bf16"
        m2, dp2, opt2, sched2 = make_trainer()
        g2 = pkg.GraphedTrainStep(m2, CFG["margin"], u, pos, neg, warmup=3, dp=dp2 if world > 1 else None)

        def step_bf():
            g2()
            opt2.step()
            sched2.batch_step()

        for _ in range(3):
            step_bf()
        nb = max(5, args.steps // 2)
        msb = timed_steps(step_bf, nb, barrier, dev, world)
        bf16 = {"value": B * world * nb / (msb * 1e-3), "unit": UNIT, "ms_per_step": msb / nb, "steps": nb, "dtype": "bf16",
                "final_loss": dp2.reduce_loss(g2.loss.detach()).item(),
                "note": "DCUE_OPERAND=bf16: same kernels, bf16 conv operands (parity table: profiles/r02_cfg1_error_table.md)"}
        g2.release()
        m2.conv._dp = None
        del m2, dp2, opt2, sched2, g2
        os.environ.pop("DCUE_OPERAND", None)
        if os.environ.get("DCUE_BENCH_RETIME") == "1":       # diagnostic: the fp16 step timed again at this point of the run
            pkg.ops.clear_workspaces()
            torch.cuda.empty_cache()
            m3, dp3, opt3, sched3 = make_trainer()
            g3 = pkg.GraphedTrainStep(m3, CFG["margin"], u, pos, neg, warmup=3, dp=dp3 if world > 1 else None)

            def step_re():
                g3()
                opt3.step()
                sched3.batch_step()

            for _ in range(3):
                step_re()
            msr = timed_steps(step_re, nb, barrier, dev, world)
            bf16["fp16_retimed"] = {"value": B * world * nb / (msr * 1e-3), "ms_per_step": msr / nb}
            g3.release()
            m3.conv._dp = None
            del m3, dp3, opt3, sched3, g3
        pkg.ops.clear_workspaces()
        torch.cuda.empty_cache()
    del pos, neg

    # ---------------- eval scorer (second half of BASELINE.json's metric): cfg5 = all-pairs cosine scores of ALL 1M user
    # factors against 500k song factors with the fused top-100; songs sharded over the ranks, the per-rank lists exchanged
    # all-to-all by user block, every rank merges its own users
    ev_users, ev_items, ev_k = args.eval_users, 500000, 100
    ge = torch.Generator(device=dev).manual_seed(3)
    ufac = torch.randn(ev_users, CFG["feat"], generator=ge, device=dev)
    lo_i, hi_i = par.shard_slice(ev_items, rank, world)
    ifac = torch.randn(ev_items, CFG["feat"], generator=ge, device=dev)[lo_i:hi_i].contiguous()   # same factors on every rank
    par.sharded_topk(ufac[:8192], ifac, ev_k, lo_i)
    barrier()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    ev_out = par.sharded_topk(ufac, ifac, ev_k, lo_i, user_tile=args.eval_user_tile if world > 1 else None)
    ee1.record()
    barrier()
    ems = torch.tensor([ee0.elapsed_time(ee1)], device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    eval_out = {"metric": "eval scored users/sec (top-%d)" % ev_k, "value": ev_users / (ems.item() * 1e-3), "unit": "users/s",
                "ms": ems.item(), "users": ev_users, "songs": ev_items, "songs_per_gpu": hi_i - lo_i, "k": ev_k,
                "tflops_algorithmic": 2.0 * CFG["feat"] * ev_users * ev_items / (ems.item() * 1e-3) / 1e12,
                "merge": "all-to-all by user block, each rank merges U/%d users (tiles of %s users overlap exchange and scoring)"
                         % (world, args.eval_user_tile) if world > 1 else "single GPU: no merge",
                "workload": "cfg5: %d users x %d songs, fp16 factors, fp32 accumulate, songs sharded over %d GPU(s), merged top-k"
                            % (ev_users, ev_items, world)}
    # The scorer reads every fp32 score from tensor memory exactly once; TMEM reads run at 64 B/clk/SM (tcgen05.ld, measured in
    # B300_MICROARCH.md), which for a 128 x 256 tile is 2 048 clocks against 900 for its MMAs: that, not the tensor pipe, is the
    # roofline of this formulation.  All-pairs bytes / (time x SMs of all ranks x SM clock).
    sm_clk = ((clocks or {}).get("sm_mhz") or 1965) * 1e6
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    tmem_bpc = 4.0 * ev_users * ev_items / (ems.item() * 1e-3) / (n_sm * world) / sm_clk
    eval_out["roofline"] = {"bound": "tmem_read", "achieved": tmem_bpc, "peak": 64.0, "unit": "B/clk/SM", "frac": tmem_bpc / 64.0,
                            "peak_source": "B300_MICROARCH.md (LDTM throughput 64 B/cyc/SM)", "algorithmic": "4 B per (user, song) score"}
    # 2-D variant at N >= 4: 2 song shards x N/2 user groups (longer song streams per work item; every rank holds half the songs)
    if world >= 4 and not args.no_eval_hybrid:
        lay = par.eval_layout(2)
        ulo, uhi = par.shard_slice(ev_users, lay["user_group"], lay["n_user_groups"])
        slo, shi = par.shard_slice(ev_items, lay["song_shard"], 2)
        ge2 = torch.Generator(device=dev).manual_seed(3)
        torch.randn(ev_users, CFG["feat"], generator=ge2, device=dev)            # advance the generator past the user factors
        ifac2 = torch.randn(ev_items, CFG["feat"], generator=ge2, device=dev)[slo:shi].contiguous()
        ug = ufac[ulo:uhi].contiguous()
        par.sharded_topk(ug[:8192], ifac2, ev_k, slo, group=lay["group"])
        hms = timed_steps(lambda: par.sharded_topk(ug, ifac2, ev_k, slo, group=lay["group"]), 1, barrier, dev, world)
        eval_out["hybrid_2_song_shards"] = {"value": ev_users / (hms * 1e-3), "unit": "users/s", "ms": hms,
                                            "layout": "2 song shards x %d user groups (each rank: %d users x %d songs)"
                                                      % (lay["n_user_groups"], uhi - ulo, shi - slo)}
        del ifac2, ug
    del ufac, ifac, ev_out
    torch.cuda.empty_cache()

    # ---------------- multi-rank parity legs + cfg4
    parity = {}
    if world > 1:
        parity["dp_parity"] = dp_parity_leg(pkg, par, dev, rank, world)
        parity["eval_parity"] = eval_parity_leg(par, ev, dev, rank, world)
        parity["table_parity"] = table_parity_leg(pkg, par, optim, dev, rank, world)
        pkg.ops.clear_workspaces()
        torch.cuda.empty_cache()
    cfg4 = None if args.no_cfg4 else cfg4_leg(pkg, par, optim, dev, rank, world, args, barrier)

    out = None
    if rank == 0:
        pk = peaks()
        S = B * (1 + N)
        kern = kernel_roofline(pkg, S, dev)
        top = max((k for k in kern if kern[k]["bound"] == "tensor"), key=lambda k: kern[k]["ms"])
        traffic_tab, traffic_src = load_ncu_traffic()
        traffic = traffic_tab.get(top, {}).get("dram_bytes_per_launch_S21504")
        traffic = None if traffic is None else traffic * S / 21504.0
        step_ms = ms_total / args.steps
        step_flop = S * STEP_FLOP_PER_SPEC_LIVE + B * USER_MLP_FLOP_PER_TRIPLET
        step_tf = step_flop / (step_ms * 1e-3) / 1e12
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        hbm = None
        try:
            import hbm_kernel_bench
            hbm = {"at_cfg2_batch": hbm_kernel_bench.run(B, tower=True, S=S)["kernels"],
                   "at_262144_rows": hbm_kernel_bench.run(262144, tower=False)["kernels"],
                   "note": "embedding / loss kernels move 2-18 MB at the cfg2 batch (launch-latency bound: 4-25 us); the second "
                           "table times them at a size where bandwidth is the limit.  Peak = measured copy bandwidth."}
        except Exception as e:  # noqa: BLE001
            hbm = {"error": str(e)}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f16", "data": "synthetic", "config": workload_config(args), "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "steps": e2e_steps,
                       "api": "DCUENet.forward-compatible dense feed: fp32 [B,128,131] + [B,N,128,131] from pinned host memory (PCIe-bound)"},
               "e2e_indexed": {"value": e2e_idx_value, "unit": UNIT, "h2d_bytes_per_step": h2d_idx, "d2h_bytes_per_step": 4,
                               "steps": idx_steps,
                               "api": "hinge_loss_step_indexed: resident pool of %d songs on the device, host sends u + song indices" % pool_songs},
               "gpu_launches": int(launches_per_step * args.steps), "launches_per_step": launches_per_step,
               "final_loss": final_loss, "cuda_graph": bool(use_graph), "operand_bf16": bf16,
               "eval": eval_out, "cfg4": cfg4,
               "roofline": {"bound": "tensor", "kernel": top, "achieved": kern[top]["achieved"], "peak": pk["burst"],
                            "unit": "TFLOP/s", "frac": kern[top]["achieved"] / pk["burst"], "traffic": traffic,
                            "traffic_source": traffic_src,
                            "peak_source": pk["source"] + " (burst bf16 cuBLAS: the kernel is timed alone)",
                            "step": {"tflops_live": step_tf, "frac_of_sustained": step_tf / pk["sustained"],
                                     "flop_per_step_live": step_flop, "peak_sustained": pk["sustained"],
                                     "note": "whole step incl. optimizer vs sustained bf16 cuBLAS (kernels timed inside a long step)"},
                            "kernels": {k: dict(v, frac=v["achieved"] / (pk["burst"] if v["bound"] == "tensor" else pk["hbm"]))
                                        for k, v in kern.items()},
                            "hbm_kernels": hbm}}
        out.update({k: v for k, v in parity.items() if v is not None})
    if world > 1:
        dist.barrier()
    if rank == 0:
        if world == 1 and not args.no_cpu:
            v, sps, n, threads = cpu_arm(64, N, 1000, 1, budget_s=args.cpu_seconds)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                   "sample": "oracle port of the reference train step (the reference is pure Python on torch and cannot "
                                             "travel to this box; the port is pinned to its outputs, tests/golden): cfg1 shape, batch 64 x "
                                             "%d negs, %d steps, %.2f s/step" % (N, n, sps)}
            out["eval"]["cpu_baseline"] = cpu_eval_arm()
        print(json.dumps(out), flush=True)
    # clean shutdown: captured graphs (they hold NCCL work) are gone, so the process group can be destroyed normally
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="triplets per GPU")
    ap.add_argument("--negs", type=int, default=20)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--eval-users", type=int, default=1000000, help="users scored in the eval leg (cfg5: 1M)")
    ap.add_argument("--eval-user-tile", type=int, default=250000, help="users per exchange tile of the song-sharded eval (N > 1)")
    ap.add_argument("--cfg4-negs", type=int, nargs="*", default=[20, 50, 100, 200])
    ap.add_argument("--cfg4-steps", type=int, default=5)
    ap.add_argument("--no-cfg4", action="store_true")
    ap.add_argument("--no-eval-hybrid", action="store_true")
    ap.add_argument("--no-bf16", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--timeline", default=None, help="write a CUPTI kernel timeline of one step to PATH.rank<r>.txt (diagnostic)")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the CUDA-graph step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
