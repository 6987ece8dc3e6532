#!/usr/bin/env python
"""Benchmark of the MMLRec training hot path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload ae_ple_t4] [--batch 4096] [--precision fp32|bf16] [--no-extras]

One "step" = one full training step (gather -> experts/gates/towers -> BCE -> backward -> fused
embedding + dense optimizer update) over one batch of synthetic input of the named shape.
Prints ONE JSON line on rank 0.  The headline workload is BASELINE.json config 2 (PLE on the AliExpress shape,
T = 4, B = 4096, bf16 tensor-core mode); the same line carries, under ``extra_workloads``, the other BASELINE
configs (census MMoE, KuaiRec SharedBottom / ESMM, MovieLens STAR / PEPNet, the 26 x 10M-row MMoE), the fp32-mode
value of the headline and the reference's step body run as torch eager on the same GPU.
"""
import argparse
import copy
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from mmlrec_b200 import synthetic  # noqa: E402

METRIC = "train samples/sec"
UNIT = "samples/s"
HEADLINE = "ae_ple_t4"
# the other BASELINE.json configs: (workload, batch, label)
EXTRA_WORKLOADS = [("census_mmoe", 1024, "config 1"), ("kuairec_sharedbottom", 4096, "config 3"), ("kuairec_esmm", 4096, "config 3"),
                   ("movielens_star", 4096, "config 4"), ("movielens_pepnet", 4096, "config 4"), ("synth26_mmoe", 4096, "config 5")]


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d["hbm_gbs"], d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def workload_config(args, name=None):
    name = name or args.workload
    kw = {}
    if getattr(args, "vocab", 0) and name == "synth26_mmoe":
        kw["vocab"] = args.vocab
    cfg, fields = synthetic.workload(name, **kw)
    return cfg, fields


def k1_bytes_per_sample(cfg, fields):
    fs = sum(1 for f in fields if f[1] == "sparse")
    fd = len(fields) - fs
    d = cfg["model_config"]["emb"]
    return 4 * (fs + fd) + 4 * d * fs + 4 * (fs * d + fd)


def gemm_flops_per_sample(model):
    """2 * sum d_in*d_out over every Linear that receives a gradient + mixture flops; x3 for fwd+bwd."""
    fwd = 0
    for name, p in model.named_parameters():
        if p.dim() == 2 and "embedding_dict" not in name:
            fwd += 2 * p.shape[0] * p.shape[1]
    return fwd, 3 * fwd


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0, period_ms=20):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", str(period_ms), "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        rows = [r for r in rows if len(r) >= 9]
        sm = [float(r[1]) for r in rows]
        mx = [float(r[2]) for r in rows]
        pw = [float(r[3]) for r in rows if r[3].replace(".", "", 1).isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        # "under load" = the samples drawing more than idle power + 20 % of the observed range
        load = [s for s, p in zip(sm, pw)] if len(pw) != len(sm) or not pw else \
            [s for s, p in zip(sm, pw) if p >= min(pw) + 0.2 * (max(pw) - min(pw))]
        return {"sm_mhz": float(np.median(load or sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(load),
                "power_w_max": max(pw) if pw else None}


def build_ours(cfg, fields, device, precision, shard=None):
    from mmlrec_b200.model import get_model_class
    from mmlrec_b200.model.utils import DenseFeat, SparseFeat
    cfg = copy.deepcopy(cfg)
    cfg["b200_config"] = {"precision": precision, "cuda_graph": True}
    if shard:
        cfg["b200_config"]["shard_tables"] = {"rank": shard[0], "world": shard[1]}
    emb = cfg["model_config"]["emb"]
    cols = [SparseFeat(n, v, emb) if k == "sparse" else DenseFeat(n, 1) for n, k, v in fields]
    torch.manual_seed(0)
    model = get_model_class(cfg["model_config"]["model_name"])(cols, device=device, config=cfg)
    model.compile(cfg["optim_config"]["optimizer"], cfg["optim_config"]["loss"], [])
    model.train()
    return model


def build_oracle(cfg, fields, seed=0, device="cpu"):
    """The reference's implementation of the path as restated in oracle/ (kind = "port"): parameters drawn with the
    reference's init distributions under the same seed as build_ours (identical initial state).  device='cuda' runs the
    same torch-eager step on the GPU -- the "library kernels" line of SURVEY section 8d."""
    from oracle.mmlrec_oracle import Column, OracleTrainer
    from mmlrec_b200.model import get_model_class
    from mmlrec_b200.model.utils import DenseFeat, SparseFeat
    emb = cfg["model_config"]["emb"]
    cols = [SparseFeat(n, v, emb) if k == "sparse" else DenseFeat(n, 1) for n, k, v in fields]
    torch.manual_seed(seed)
    shell = get_model_class(cfg["model_config"]["model_name"])(cols, device="cpu", config=copy.deepcopy(cfg))
    names = [n for n, _ in shell.named_parameters()]
    sd = shell.state_dict()
    params = {k: v.to(device) for k, v in sd.items() if k in names}
    if hasattr(shell, "linears") and len(shell.linears) and hasattr(shell.linears[0], "spec_weights"):  # STAR's frozen tensors
        for prefix, mods in (("linears", shell.linears), ("final_layers", shell.final_layers)):
            for j, m in enumerate(mods):
                for i in range(m.num_domains - 1):
                    params[f"{prefix}.{j}.specific_weights.{i}"] = getattr(m, f"frozen_weight_{i}").detach().clone().to(device)
                    params[f"{prefix}.{j}.specific_biases.{i}"] = getattr(m, f"frozen_bias_{i}").detach().clone().to(device)
    ocols = [Column(n, k, v, emb if k == "sparse" else 1) for n, k, v in fields]
    return OracleTrainer(cfg, ocols, params, {k: v.to(device) for k, v in sd.items() if k not in names}, names)


def table_elements(cfg, fields):
    return sum(v for _, k, v in fields if k == "sparse") * cfg["model_config"]["emb"]


def time_oracle(cfg, fields, batch, budget_s=20.0, max_steps=12, warmup=2, device="cpu"):
    tr = build_oracle(cfg, fields, device=device)
    batches = [synthetic.make_batch(cfg, fields, batch, seed=s) for s in range(4)]
    batches = [(torch.from_numpy(x).to(device), torch.from_numpy(y).to(device)) for x, y in batches]
    sync = torch.cuda.synchronize if device != "cpu" else (lambda: None)
    for s in range(warmup):
        tr.step(*batches[s % 4])
    sync()
    t0, n = time.perf_counter(), 0
    while n < max_steps and (time.perf_counter() - t0) < budget_s:
        tr.step(*batches[n % 4])
        n += 1
    sync()
    dt = time.perf_counter() - t0
    return batch * n / dt, n, dt


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (oracle port: /root/reference does not exist on the GPU box),
    all host threads, each step a full training step of one per-GPU batch of the workload."""
    if rank != 0:
        return
    cfg, fields = workload_config(args)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    tr = build_oracle(cfg, fields)
    batches = [synthetic.make_batch(cfg, fields, args.batch, seed=s) for s in range(4)]
    batches = [(torch.from_numpy(x), torch.from_numpy(y)) for x, y in batches]
    steps = min(args.steps, 600)   # bounded sample: ~0.1 s per step of the headline workload on 16+ cores
    for s in range(min(args.warmup, 5)):
        tr.step(*batches[s % 4])
    t0 = time.perf_counter()
    for s in range(steps):
        tr.step(*batches[s % 4])
    dt = time.perf_counter() - t0
    val = args.batch * steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "batch_per_gpu": args.batch, "global_batch": args.batch * max(args.gpus, 1),
                       "model": cfg["model_config"]["model_name"], "optimizer": cfg["optim_config"]["optimizer"],
                       "precision": "fp32", "parallelism": "host cpu (each step = one per-GPU batch of the workload)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{steps} full training steps of batch {args.batch} "
                                       "(oracle/mmlrec_oracle.py: the reference's step body on torch CPU)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def time_graph(fn, reps=20, iters=5):
    """Device time of fn(), with fn's launches captured `reps` times back to back in a CUDA graph so that host launch
    overhead (about 10 us per ctypes launch, more than several of these kernels take) is out of the measurement."""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                fn()
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = float("inf")
    for _ in range(iters):
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / reps)
    return best


def gather_roofline(model, cfg, fields, hbm_peak, which, rows=1 << 20):
    """K1 alone on 2^20 resident rows: achieved = algorithmic bytes / CUDA-event time."""
    from mmlrec_b200 import lib as L
    n = rows
    X, _ = synthetic.make_batch(cfg, fields, n, seed=99)
    Xd = torch.from_numpy(X).cuda()
    fs = [t for t in model.embedding_layout]
    D = model.emb_dim
    meta = torch.tensor([v for t in fs for v in (t[0]._mm_off, t[1], t[2], t[3])], dtype=torch.int64, device="cuda")
    dense_cols = model.dense_x_cols
    in_dim = model.input_dim_total
    ld = (in_dim + 3) // 4 * 4
    out = torch.empty(n, ld, device="cuda")
    lib = L.load()
    dc = torch.tensor(dense_cols or [0], dtype=torch.int32, device="cuda")
    oob = torch.zeros(1, dtype=torch.int32, device="cuda")

    def run():
        L.check(lib.mmlrec_gather_concat(Xd.data_ptr(), Xd.stride(0), n, model.store.emb.data_ptr(), meta.data_ptr(),
                                         len(fs), D, dc.data_ptr(), len(dense_cols), len(fs) * D, out.data_ptr(), ld,
                                         None, 0, oob.data_ptr(), torch.cuda.current_stream().cuda_stream))
    ms = time_graph(run, reps=5)
    bytes_alg = n * k1_bytes_per_sample(cfg, fields)
    ach = bytes_alg / (ms * 1e-3) / 1e9
    traffic, src = None, None
    tpath = os.path.join(ROOT, "profiles", "gather_traffic_r02.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath))
        if t.get("workload") == cfg.get("_workload_name") and t.get("rows") == n:
            traffic, src = t["dram_bytes_per_launch"], "profiles/gather_traffic_r02.json (ncu --set full of this launch)"
    del out, Xd
    return {"kernel": "gather_concat_kernel", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
            "frac": ach / hbm_peak, "traffic": traffic, "traffic_source": src, "peak_source": which, "rows": n, "ms": ms,
            "algorithmic_bytes_per_row": k1_bytes_per_sample(cfg, fields),
            "l2": "X (1.3e8+ B) and the output (8e8+ B) exceed the 126 MB L2; the tables are what the workload makes them"}


def gemm_roofline(model, plan, tf_burst, which):
    """The step's dominant kernel: the tcgen05 grouped GEMM (bf16) / gemm_grouped_f32_kernel (fp32).  Every GEMM launch
    of the step is timed alone with CUDA events around a graph of 20 back-to-back replays of that launch (operands
    L2-warm, as they are inside the step where the previous kernel has just produced them; burst peak as denominator);
    achieved = algorithmic flops of all launches / their summed duration."""
    from mmlrec_b200.engine.core import LinearStage
    st = torch.cuda.current_stream
    B = plan.B
    total_ms, total_flops, per = 0.0, 0.0, []
    for s in plan.stages:
        if not isinstance(s, LinearStage):
            continue
        fwd_flops = sum(2.0 * B * g.N * g.K for g in s.groups)
        bwd_flops = sum((4.0 if g.x.group.need_grad else 2.0) * B * g.N * g.K for g in s.live_groups)
        for label, tables, flops in (("fwd", s.fwd, fwd_flops), ("bwd", s.bwd, bwd_flops)):
            if not tables:
                continue
            ms = time_graph(lambda: [s._launch(t, st().cuda_stream, "bench") for t in tables])
            total_ms += ms
            total_flops += flops
            per.append({"launch": f"{label}:{s.label}", "ms": ms, "tflops": flops / (ms * 1e-3) / 1e12})
    ach = total_flops / (total_ms * 1e-3) / 1e12
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "tc_gemm_traffic_r02.json")
    if plan.b.tc and os.path.exists(tpath) and per:
        t = json.load(open(tpath))
        if len(t["dram_bytes_per_launch"]) == len(per):
            traffic = t["dram_bytes_per_step"] / len(per)
            traffic_src = "profiles/tc_gemm_traffic_r02.json (ncu --set full of this workload, mean over the step's launches)"
    kname = ("gemm_grouped_tc2_kernel" if plan.b.tc_kernel == 2 else "gemm_grouped_tc_kernel") if plan.b.tc else "gemm_grouped_f32_kernel"
    return {"kernel": kname, "bound": "tensor", "achieved": ach, "peak": tf_burst, "unit": "TFLOP/s", "frac": ach / tf_burst,
            "traffic": traffic, "traffic_unit": "bytes/launch", "traffic_source": traffic_src,
            "peak_source": which + " (bf16 cuBLAS burst)", "launches_per_step": len(per), "flops_per_step": total_flops,
            "ms_per_step_in_gemm": total_ms, "timing": "CUDA events around a graph of 20 back-to-back replays of each launch on the same operands (L2-warm; "
                                                        "consecutive launches overlap through programmatic dependent launch), best of 5",
            "per_launch": per}


def stage_breakdown(model, plan, iters=10):
    """Eager per-stage device time (CUDA events on the launch stream) of one training step.  Host launch overhead
    inflates stages of a few microseconds; the graph-timed roofline entries are the kernel numbers."""
    from mmlrec_b200 import lib as L
    st = torch.cuda.current_stream().cuda_stream
    acc = {}

    def timed(label, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        return label, a, b

    for it in range(iters + 2):
        evs = []
        evs.append(timed("clock+catch-up+sort", lambda: (plan.advance_clock(st), plan.gather.catch_up(st), plan.gather.sort(st))))
        for i, s in enumerate(plan.stages):
            evs.append(timed(f"fwd:{i}:{s.name}:{getattr(s, 'label', '')}", lambda s=s: s.forward(st, True)))
        for i, s in reversed(list(enumerate(plan.stages))):
            evs.append(timed(f"bwd:{i}:{s.name}:{getattr(s, 'label', '')}", lambda s=s: s.backward(st)))
        store = model.store
        p = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        evs.append(timed("dense_optimizer", lambda: L.check(plan.b.lib.mmlrec_dense_optimizer_step_sliced(
            store.dense.data_ptr(), store.dense_grad.data_ptr(), p(store.dense_s1), p(store.dense_s2), store.n_dense,
            model.hyper_dev.data_ptr(), p(store.dense_bf16), plan.grad_slices, store.slice_stride, st))))
        torch.cuda.synchronize()
        if it >= 2:
            for label, a, b in evs:
                acc[label] = acc.get(label, 0.0) + a.elapsed_time(b)
    return {k: v / iters for k, v in acc.items()}


def self_check(model, cfg, fields, B, precision):
    """First training step of the timed configuration against the oracle on the same seeded state and batch
    (loss and predictions; 2e-2 in bf16 mode, 1e-4 / 2e-5 in fp32 mode).  Skipped when the CPU copy of the tables
    would not fit comfortably (the 26 x 10M-row workload: 8.3 GB + dense gradients of the same size)."""
    if table_elements(cfg, fields) > 3e8:
        return {"checked": False, "why": "tables too large for a CPU oracle copy; this shape is covered with 1e5-row "
                                         "tables by tests/test_bench_shapes_gpu.py"}
    tr = build_oracle(cfg, fields)
    X, y = synthetic.make_batch(cfg, fields, B, seed=4242)
    pred_o, loss_o = tr.step(torch.from_numpy(X), torch.from_numpy(y))
    loss = model.train_on_batch(X, y)
    torch.cuda.synchronize()
    got, want = float(loss[-1].item()), float(loss_o)
    pred = model.plan(B).pred.cpu()
    perr = float((pred - pred_o).norm() / pred_o.norm())
    ltol, ptol = (2e-2, 2e-2) if precision == "bf16" else (1e-4, 2e-5)
    ok = abs(got - want) <= ltol * abs(want) and perr <= ptol
    if not ok:
        raise AssertionError(f"bench self-check failed: loss {got} vs oracle {want}, prediction rel err {perr:.3e}")
    return {"checked": True, "loss": got, "oracle_loss": want, "pred_rel_err": perr, "tolerance": ptol}


def measure(args, workload, B, precision, rank, world, device, sharded, steps, warmup, sampler_hook=None, check=True):
    """Build the model of `workload`, verify its first step, time `steps` steps resident + end to end."""
    from mmlrec_b200 import lib as L
    cfg, fields = workload_config(args, workload)
    cfg["_workload_name"] = workload
    model = build_ours(cfg, fields, device, precision, shard=(rank, world) if sharded else None)
    if world > 1:
        from mmlrec_b200 import parallel
        if sharded:
            parallel.attach_sharded(model)
        else:
            parallel.attach(model, rank, world)
    out = {"model": model, "cfg": cfg, "fields": fields}
    if check and world == 1:
        out["self_check"] = self_check(model, cfg, fields, B, precision)
    pool = 8
    host = [synthetic.make_batch(cfg, fields, B, seed=1000 * rank + s) for s in range(pool)]
    Xh = [torch.from_numpy(x).pin_memory() for x, _ in host]
    yh = [torch.from_numpy(y).pin_memory() for _, y in host]
    Xd = [x.to(device) for x in Xh]
    yd = [y.to(device) for y in yh]
    plan = model.plan(B)
    out["plan"] = plan

    def step_resident(i):
        plan.X.copy_(Xd[i % pool], non_blocking=True)
        plan.y.copy_(yd[i % pool], non_blocking=True)
        model._run_train(plan)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    l0 = L.launch_count()
    step_resident(0)          # eager (also surfaces launch errors); the next one captures the CUDA graph
    out["launches_per_step"] = L.launch_count() - l0
    for i in range(1, max(warmup, 3)):
        step_resident(i)
    barrier()
    if sampler_hook:
        sampler_hook("start")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    a.record()
    for i in range(steps):
        step_resident(i)
    b.record()
    barrier()
    ms = a.elapsed_time(b)
    # repeat blocks of the same length (stability of the number; the clock sampler keeps running)
    reps = []
    for _ in range(args.repeats):
        a.record()
        for i in range(steps):
            step_resident(i)
        b.record()
        torch.cuda.synchronize()
        reps.append(a.elapsed_time(b) / steps)
    # end to end through the public API with HOST buffers (H2D of X,y + D2H of the loss every step)
    for i in range(3):
        model.train_on_batch(Xh[i % pool], yh[i % pool])
    barrier()
    t0 = time.perf_counter()
    a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a2.record()
    last = 0.0
    # every step: H2D of that step's pinned inputs (inside train_on_batch) and a D2H read of a step's loss into pinned
    # memory.  The read is issued behind the step on the stream and waited for one step later, so the host prepares and
    # launches step i+1 while step i runs (what any input pipeline does; fit() works the same way)
    loss_host = [torch.empty(plan.loss.numel(), dtype=torch.float32).pin_memory() for _ in range(2)]
    read = [torch.cuda.Event(), torch.cuda.Event()]
    for i in range(steps):
        loss = model.train_on_batch(Xh[i % pool], yh[i % pool])
        loss_host[i & 1].copy_(loss, non_blocking=True)   # device -> host read of the step's result
        read[i & 1].record()
        if i > 0:
            read[(i - 1) & 1].synchronize()
            last = float(loss_host[(i - 1) & 1][-1])
    read[(steps - 1) & 1].synchronize()
    last = float(loss_host[(steps - 1) & 1][-1])
    b2.record()
    barrier()
    ms_e2e = max(a2.elapsed_time(b2), 1e3 * (time.perf_counter() - t0))
    if sampler_hook:
        sampler_hook("stop")
    if sharded:
        model.shard.check()   # raises if a peer-memory barrier ever gave up waiting (a rank died)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    out.update(ms=ms, ms_e2e=ms_e2e, last_loss=last, repeat_ms_per_step=reps,
               h2d=int(Xh[0].numel() * 4 + yh[0].numel() * 4), d2h=int(plan.loss.numel() * 4))
    return out


def summarize(m, workload, B, world, steps, precision, sharded):
    cfg = m["cfg"]
    fwd_fl, tot_fl = gemm_flops_per_sample(m["model"])
    return {"workload": workload, "value": world * B * steps / (m["ms"] * 1e-3), "unit": UNIT, "ms_per_step": m["ms"] / steps,
            "e2e": {"value": world * B * steps / (m["ms_e2e"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": m["h2d"],
                    "d2h_bytes_per_step": m["d2h"],
                    "ms_per_step": m["ms_e2e"] / steps,
                    "how": "train_on_batch(pinned host X, y) every step; the step's loss vector is copied to pinned host "
                           "memory behind the step and read one step later"},
            "batch_per_gpu": B, "model": cfg["model_config"]["model_name"], "optimizer": cfg["optim_config"]["optimizer"],
            "precision": precision, "dtype": "f32" if precision == "fp32" else "bf16",
            "tables": "row-sharded" if sharded else ("replicated" if world > 1 else "local"),
            "launches_per_step": int(m["launches_per_step"]), "last_loss": m["last_loss"],
            "self_check": m.get("self_check"), "gemm_flops_per_sample": {"fwd": fwd_fl, "fwd_bwd": tot_fl}}


def run_ours(args, rank, world):
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU path)"
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(device))
    # multi-GPU default: row-sharded tables (every rank sorts / updates / sweeps only the rows it owns)
    sharded = world > 1 and args.tables in ("sharded", "auto")
    B = args.batch
    sampler = {"s": None}

    def hook(what):
        if rank != 0:
            return
        if what == "start":
            sampler["s"] = ClockSampler(local)
        else:
            sampler["clocks"] = sampler["s"].stop()

    m = measure(args, args.workload, B, args.precision, rank, world, device, sharded, args.steps, args.warmup, hook)
    model, plan, cfg, fields = m["model"], m["plan"], m["cfg"], m["fields"]
    hbm, tf_burst, tf_sus, which = measured_peaks()
    head = summarize(m, args.workload, B, world, args.steps, args.precision, sharded)
    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": head["dtype"], "data": "synthetic",
                "config": {"workload": args.workload, "batch_per_gpu": B, "global_batch": B * world,
                           "model": head["model"], "optimizer": head["optimizer"], "precision": args.precision,
                           "parallelism": f"dp{world}" + (" + row-sharded tables (owner = id mod R, NVLink peer memory)"
                                                          if sharded else ""),
                           "vocab_override": args.vocab or None,
                           "l2": "tables + Adam state (>0.25 GB) exceed the 126 MB L2 and are streamed every step; "
                                 "batches rotate through a pool of 8"},
                "e2e": head["e2e"], "gpu_launches": int(m["launches_per_step"] * args.steps),
                "launches_per_step": int(m["launches_per_step"]), "last_loss": m["last_loss"], "self_check": head["self_check"],
                "repeat_ms_per_step": m["repeat_ms_per_step"], "clocks": sampler.get("clocks"),
                "gemm_flops_per_sample": head["gemm_flops_per_sample"],
                "achieved_gemm_tflops_whole_step": world * B * head["gemm_flops_per_sample"]["fwd_bwd"] /
                (head["ms_per_step"] * 1e-3) / 1e12}
    if world == 1 and not args.no_extras:
        try:
            line["breakdown_ms"] = stage_breakdown(model, plan)
        except Exception as e:  # noqa: BLE001
            line["breakdown_ms"] = {"error": repr(e)}
        try:
            line["roofline"] = gemm_roofline(model, plan, tf_burst, which)
        except Exception as e:  # noqa: BLE001
            line["roofline"] = {"error": repr(e)}
        line["roofline_gather"] = gather_roofline(model, cfg, fields, hbm, which)
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        v, n, dt = time_oracle(cfg, fields, B)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{n} full training steps of batch {B} in {dt:.1f} s "
                                          "(oracle/mmlrec_oracle.py, torch CPU)"}
        try:   # the reference's step body as torch eager on this GPU: the "library kernels" number
            v, n, dt = time_oracle(cfg, fields, B, budget_s=10.0, max_steps=30, warmup=3, device=device)
            line["reference_eager_gpu"] = {"value": v, "unit": UNIT, "kind": "port on cuda (oracle/mmlrec_oracle.py: the same "
                                           "ATen ops + torch.optim as the reference's basemodel.py:261-313, fp32, eager)",
                                           "steps": n, "ms_per_step": 1e3 * dt / n}
        except Exception as e:  # noqa: BLE001
            line["reference_eager_gpu"] = {"error": repr(e)}
    del m, model, plan
    torch.cuda.empty_cache()
    # ---- the other configurations (same measurement, fewer steps)
    extras = []
    if not args.no_extras and args.workload == HEADLINE:
        todo = ([(HEADLINE, B, "config 2, fp32 mode", "fp32")] if world == 1 else []) + \
               [(w, b, lab, "bf16") for w, b, lab in EXTRA_WORKLOADS if world == 1 or w == "synth26_mmoe"]
        for wl, b_, label, prec in todo:
            try:
                steps = min(args.steps, 50)
                mm = measure(args, wl, b_, prec, rank, world, device, sharded, steps, 3)
                e = summarize(mm, wl, b_, world, steps, prec, sharded)
                e["baseline_config"] = label
                if rank == 0 and world == 1:
                    if prec == "bf16":
                        e["roofline_gather"] = gather_roofline(mm["model"], mm["cfg"], mm["fields"], hbm, which, rows=1 << 19)
                        try:
                            r = gemm_roofline(mm["model"], mm["plan"], tf_burst, which)
                            e["roofline"] = {k: r[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac",
                                                               "launches_per_step", "ms_per_step_in_gemm")}
                        except Exception as ex:  # noqa: BLE001
                            e["roofline"] = {"error": repr(ex)}
                    if table_elements(mm["cfg"], mm["fields"]) <= 3e8 and wl != HEADLINE:
                        torch.set_num_threads(os.cpu_count() or 1)
                        v, n, dt = time_oracle(mm["cfg"], mm["fields"], b_, budget_s=6.0, max_steps=8, warmup=1)
                        e["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                             "sample": f"{n} training steps of batch {b_} in {dt:.1f} s"}
                extras.append(e)
                del mm
                torch.cuda.empty_cache()
            except Exception as ex:  # noqa: BLE001
                extras.append({"workload": wl, "error": repr(ex)})
    if rank == 0:
        line["extra_workloads"] = extras
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        # no teardown: destroying the process group while captured graphs still hold its collectives on two streams can
        # block at interpreter exit; every rank has passed the barrier, the measurement is printed -> leave
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--repeats", type=int, default=4, help="extra timed blocks of --steps steps (reported, not the headline)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=HEADLINE)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--vocab", type=int, default=0, help="rows per table for synth26_mmoe (default 10M)")
    ap.add_argument("--tables", default="auto", choices=["auto", "replicated", "sharded"],
                    help="multi-GPU table placement; auto = row-sharded")
    ap.add_argument("--no-extras", action="store_true", help="skip breakdown / roofline / cpu baseline / other workloads")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
