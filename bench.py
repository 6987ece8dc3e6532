#!/usr/bin/env python
"""Benchmark of the MMLRec training hot path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload ae_ple_t4] [--batch 4096] [--precision fp32|bf16]

One "step" = one full training step (gather -> experts/gates/towers -> BCE -> backward -> fused
embedding + dense optimizer update) over one batch of synthetic input of the named shape.
Prints ONE JSON line on rank 0.
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


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d["hbm_gbs"], d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def workload_config(args):
    kw = {}
    if args.vocab and args.workload == "synth26_mmoe":
        kw["vocab"] = args.vocab
    cfg, fields = synthetic.workload(args.workload, **kw)
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

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm = [float(r[1]) for r in rows if len(r) >= 9]
        mx = [float(r[2]) for r in rows if len(r) >= 9]
        reasons = set()
        for r in rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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


def build_oracle(cfg, fields, seed=0):
    """The reference's CPU implementation of the path, as restated in oracle/ (kind = "port"):
    parameters drawn with the reference's init distributions."""
    from oracle.mmlrec_oracle import Column, OracleTrainer
    from mmlrec_b200.model import get_model_class
    from mmlrec_b200.model.utils import DenseFeat, SparseFeat
    emb = cfg["model_config"]["emb"]
    cols = [SparseFeat(n, v, emb) if k == "sparse" else DenseFeat(n, 1) for n, k, v in fields]
    torch.manual_seed(seed)
    shell = get_model_class(cfg["model_config"]["model_name"])(cols, device="cpu", config=copy.deepcopy(cfg))
    names = {n for n, _ in shell.named_parameters()}
    sd = shell.state_dict()
    ocols = [Column(n, k, v, emb if k == "sparse" else 1) for n, k, v in fields]
    return OracleTrainer(cfg, ocols, {k: v for k, v in sd.items() if k in names},
                         {k: v for k, v in sd.items() if k not in names}, sorted(names))


def time_oracle(cfg, fields, batch, budget_s=20.0, max_steps=12, warmup=2):
    tr = build_oracle(cfg, fields)
    batches = [synthetic.make_batch(cfg, fields, batch, seed=s) for s in range(4)]
    batches = [(torch.from_numpy(x), torch.from_numpy(y)) for x, y in batches]
    for s in range(warmup):
        tr.step(*batches[s % 4])
    t0, n = time.perf_counter(), 0
    while n < max_steps and (time.perf_counter() - t0) < budget_s:
        tr.step(*batches[n % 4])
        n += 1
    dt = time.perf_counter() - t0
    return batch * n / dt, n, dt


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    cfg, fields = workload_config(args)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    tr = build_oracle(cfg, fields)
    batches = [synthetic.make_batch(cfg, fields, args.batch, seed=s) for s in range(4)]
    batches = [(torch.from_numpy(x), torch.from_numpy(y)) for x, y in batches]
    for s in range(args.warmup):
        tr.step(*batches[s % 4])
    t0 = time.perf_counter()
    for s in range(args.steps):
        tr.step(*batches[s % 4])
    dt = time.perf_counter() - t0
    val = args.batch * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "batch_per_gpu": args.batch, "global_batch": args.batch * max(args.gpus, 1),
                       "model": cfg["model_config"]["model_name"], "optimizer": cfg["optim_config"]["optimizer"],
                       "precision": "fp32", "parallelism": "host cpu (each step = one per-GPU batch of the workload)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{args.steps} full training steps of batch {args.batch} "
                                       "(oracle/mmlrec_oracle.py: the reference's step body on torch CPU)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def time_kernel_eager(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def gather_roofline(model, cfg, fields, hbm_peak, which):
    """K1 alone on 2^20 resident rows: achieved = algorithmic bytes / CUDA-event time."""
    from mmlrec_b200 import ops
    n = 1 << 20
    X, _ = synthetic.make_batch(cfg, fields, n, seed=99)
    Xd = torch.from_numpy(X).cuda()
    fs = [t for t in model.embedding_layout]
    D = model.emb_dim
    meta = torch.tensor([v for t in fs for v in (t[0]._mm_off, t[1], t[2], t[3])], dtype=torch.int64, device="cuda")
    dense_cols = model.dense_x_cols
    in_dim = model.input_dim_total
    ld = (in_dim + 3) // 4 * 4
    out = torch.empty(n, ld, device="cuda")
    from mmlrec_b200 import lib as L
    lib = L.load()
    dc = torch.tensor(dense_cols or [0], dtype=torch.int32, device="cuda")
    oob = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def run():
        L.check(lib.mmlrec_gather_concat(Xd.data_ptr(), Xd.stride(0), n, model.store.emb.data_ptr(), meta.data_ptr(),
                                         len(fs), D, dc.data_ptr(), len(dense_cols), len(fs) * D, out.data_ptr(), ld,
                                         None, 0, oob.data_ptr(), st))
    ms = time_kernel_eager(run)
    bytes_alg = n * k1_bytes_per_sample(cfg, fields)
    ach = bytes_alg / (ms * 1e-3) / 1e9
    del out, Xd
    return {"kernel": "gather_concat_kernel", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
            "frac": ach / hbm_peak, "traffic": None, "peak_source": which, "rows": n, "ms": ms,
            "algorithmic_bytes_per_row": k1_bytes_per_sample(cfg, fields)}


def gemm_roofline(model, plan, tf_burst, which, iters=20):
    """The step's dominant kernel: gemm_grouped_tc_kernel (bf16) / gemm_grouped_f32_kernel (fp32).  Every GEMM
    launch of the step is timed on its own with CUDA events on the launch stream (kernel timed alone -> burst
    peak); achieved = algorithmic flops of all launches / their summed duration."""
    from mmlrec_b200.engine.core import LinearStage
    st = torch.cuda.current_stream().cuda_stream
    B = plan.B
    total_ms, total_flops, per = 0.0, 0.0, []
    for s in plan.stages:
        if not isinstance(s, LinearStage):
            continue
        fwd_flops = sum(2.0 * B * g.N * g.K for g in s.groups)
        bwd_flops = sum((4.0 if g.x.group.need_grad else 2.0) * B * g.N * g.K for g in s.live_groups)
        for label, tables, flops in (("fwd", s.fwd, fwd_flops), ("bwd", s.bwd, bwd_flops)):
            ms = time_kernel_eager(lambda: [s._launch(t, st, "bench") for t in tables], iters=iters)
            total_ms += ms
            total_flops += flops
            per.append({"launch": f"{label}:{s.label}", "ms": ms, "tflops": flops / (ms * 1e-3) / 1e12})
    ach = total_flops / (total_ms * 1e-3) / 1e12
    # DRAM bytes per launch (read + write) from the committed `ncu --set full` capture of this workload
    traffic, traffic_src = None, None
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "tc_gemm_traffic_r01.json")
    if plan.b.tc and os.path.exists(tpath) and per:
        t = json.load(open(tpath))
        if len(t["dram_bytes_per_launch"]) == len(per):
            traffic = t["dram_bytes_per_step"] / len(per)
            traffic_src = "profiles/tc_gemm_traffic_r01.json (ncu --set full, mean over the step's launches)"
    return {"kernel": "gemm_grouped_tc_kernel" if plan.b.tc else "gemm_grouped_f32_kernel", "bound": "tensor",
            "achieved": ach, "peak": tf_burst, "unit": "TFLOP/s", "frac": ach / tf_burst, "traffic": traffic,
            "traffic_unit": "bytes/launch", "traffic_source": traffic_src,
            "peak_source": which + " (bf16 cuBLAS burst)", "launches_per_step": len(per),
            "flops_per_step": total_flops, "ms_per_step_in_gemm": total_ms, "per_launch": per}


def stage_breakdown(model, plan, iters=10):
    """Eager per-stage device time (CUDA events on the launch stream) of one training step."""
    from mmlrec_b200 import lib as L
    st = torch.cuda.current_stream().cuda_stream
    names, acc = [], {}

    def timed(label, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        return label, a, b

    for it in range(iters + 2):
        evs = []
        evs.append(timed("clock+sort", lambda: (L.check(plan.b.lib.mmlrec_hyper_advance(model.hyper_dev.data_ptr(), st)),
                                                plan.gather.sort(st))))
        for i, s in enumerate(plan.stages):
            evs.append(timed(f"fwd:{i}:{s.name}:{getattr(s, 'label', '')}", lambda s=s: s.forward(st, True)))
        for i, s in reversed(list(enumerate(plan.stages))):
            evs.append(timed(f"bwd:{i}:{s.name}:{getattr(s, 'label', '')}", lambda s=s: s.backward(st)))
        store = model.store
        p = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        evs.append(timed("dense_optimizer", lambda: L.check(plan.b.lib.mmlrec_dense_optimizer_step(
            store.dense.data_ptr(), store.dense_grad.data_ptr(), p(store.dense_s1), p(store.dense_s2), store.n_dense,
            model.hyper_dev.data_ptr(), p(store.dense_bf16), st))))
        torch.cuda.synchronize()
        if it >= 2:
            for label, a, b in evs:
                acc[label] = acc.get(label, 0.0) + a.elapsed_time(b)
    return {k: v / iters for k, v in acc.items()}


def run_ours(args, rank, world):
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU path)"
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(device))
    from mmlrec_b200 import lib as L
    cfg, fields = workload_config(args)
    # multi-GPU default: row-sharded tables (every rank sorts / updates / sweeps only the rows it owns; measured
    # faster than replicated tables already at 2 GPUs, profiles/bench_ple_dp2_*_r01.json)
    sharded = world > 1 and args.tables in ("sharded", "auto")
    model = build_ours(cfg, fields, device, args.precision, shard=(rank, world) if sharded else None)
    if world > 1:
        from mmlrec_b200 import parallel
        if sharded:
            parallel.attach_sharded(model)
        else:
            parallel.attach(model, rank, world)
    B = args.batch
    pool = 8
    host = [synthetic.make_batch(cfg, fields, B, seed=1000 * rank + s) for s in range(pool)]
    Xh = [torch.from_numpy(x).pin_memory() for x, _ in host]
    yh = [torch.from_numpy(y).pin_memory() for _, y in host]
    Xd = [x.to(device) for x in Xh]
    yd = [y.to(device) for y in yh]
    plan = model.plan(B)

    def step_resident(i):
        plan.X.copy_(Xd[i % pool], non_blocking=True)
        plan.y.copy_(yd[i % pool], non_blocking=True)
        model._run_train(plan)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (first step eager, second captures the CUDA graph)
    l0 = L.launch_count()
    step_resident(0)
    l1 = L.launch_count()
    launches_per_step = l1 - l0
    for i in range(1, max(args.warmup, 3)):
        step_resident(i)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    # ---- timed: K steps, inputs resident in HBM
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    a.record()
    for i in range(args.steps):
        step_resident(i)
    b.record()
    barrier()
    ms = a.elapsed_time(b)
    # ---- timed: end to end through the public API with HOST buffers (H2D of X,y + D2H of the loss)
    for i in range(3):
        model.train_on_batch(Xh[i % pool], yh[i % pool])
    barrier()
    t0 = time.perf_counter()
    a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a2.record()
    last = 0.0
    for i in range(args.steps):
        loss = model.train_on_batch(Xh[i % pool], yh[i % pool])
        last = float(loss[-1].item())  # device -> host read of the step's result
    b2.record()
    barrier()
    ms_e2e = max(a2.elapsed_time(b2), 1e3 * (time.perf_counter() - t0))
    clocks = sampler.stop() if sampler else None
    if sharded:
        model.shard.check()   # raises if a peer-memory barrier ever gave up waiting (a rank died)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        return
    hbm, tf_burst, tf_sus, which = measured_peaks()
    value = world * B * args.steps / (ms * 1e-3)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)
    fwd_fl, tot_fl = gemm_flops_per_sample(model)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": args.workload, "batch_per_gpu": B, "global_batch": B * world,
                       "model": cfg["model_config"]["model_name"], "optimizer": cfg["optim_config"]["optimizer"],
                       "precision": args.precision,
                       "parallelism": f"dp{world}" + (" + row-sharded tables (owner = id mod R, NVLink peer memory)"
                                                      if sharded else ""),
                       "vocab_override": args.vocab or None,
                       "l2": "tables + Adam state (>0.25 GB) exceed the 126 MB L2 and are streamed by the dense-Adam "
                             "sweep every step; batches rotate through a pool of 8"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(Xh[0].numel() * 4 + yh[0].numel() * 4),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches_per_step * args.steps), "launches_per_step": int(launches_per_step),
            "last_loss": last, "clocks": clocks,
            "gemm_flops_per_sample": {"fwd": fwd_fl, "fwd_bwd": tot_fl},
            "achieved_gemm_tflops_whole_step": world * B * tot_fl / (ms / args.steps * 1e-3) / 1e12}
    if world == 1 and not args.no_extras:
        try:
            line["breakdown_ms"] = stage_breakdown(model, plan)
            dom = max(line["breakdown_ms"].items(), key=lambda kv: kv[1])
            line["dominant_stage"] = {"stage": dom[0], "ms": dom[1]}
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
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ae_ple_t4")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--vocab", type=int, default=0, help="rows per table for synth26_mmoe (default 10M)")
    ap.add_argument("--tables", default="auto", choices=["auto", "replicated", "sharded"],
                    help="multi-GPU table placement; auto = row-sharded")
    ap.add_argument("--no-extras", action="store_true", help="skip breakdown / roofline / cpu baseline")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
