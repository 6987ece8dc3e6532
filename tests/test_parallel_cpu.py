"""CPU, world_size 2, gloo: the host-side collective logic of the data-parallel step
(mmlrec_b200/parallel.py): rank-major global batch assembly and SUM (not mean) gradient reduction."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mmlrec_b200.parallel import DataParallelContext
    ctx = DataParallelContext(rank, world)
    b = 5
    local = torch.arange(b * 3, dtype=torch.float32).reshape(b, 3) + 100 * rank
    out = torch.zeros(world * b, 3)
    ctx.gather_rows(local, out)
    grad = torch.full((7,), float(rank + 1))
    ctx.sum_gradients(grad)
    lo, hi = ctx.shard(world * b)
    results[rank] = (out.clone(), grad.clone(), (lo, hi), ctx.global_batch(b))
    dist.destroy_process_group()


def test_gather_is_rank_major_and_reduction_is_sum():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
        res = dict(results)
    b = 5
    base = torch.arange(b * 3, dtype=torch.float32).reshape(b, 3)
    want = torch.cat([base, base + 100])
    for r in range(world):
        out, grad, (lo, hi), gb = res[r]
        assert torch.equal(out, want), "global batch = ranks' batches in rank order"
        assert torch.equal(grad, torch.full((7,), 3.0)), "dense gradients are SUMMED (loss reduction='sum')"
        assert (lo, hi) == (r * b, (r + 1) * b) and gb == world * b
        assert torch.equal(out[lo:hi], base + 100 * r)


# ---------------------------------------------------------------------------------------------
# row-sharded tables: owner(id) = id mod R, local row id // R (host-side layout; the exchange kernels are
# exercised on GPUs by tools/sharded_equivalence.py)
# ---------------------------------------------------------------------------------------------
def _shard_worker(rank, world, port, results):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import copy
    from mmlrec_b200 import parallel, synthetic
    from mmlrec_b200.model import get_model_class
    from mmlrec_b200.model.utils import SparseFeat
    cfg, _ = synthetic.workload("synth26_mmoe", vocab=11)
    fields = [("a", "sparse", 11), ("b", "sparse", 4), ("c", "sparse", 1)]   # ragged: 11 and 1 do not divide by 2
    cfg = copy.deepcopy(cfg)
    cfg["b200_config"] = {"shard_tables": {"rank": rank, "world": world}}
    cols = [SparseFeat(n, v, 8) for n, _, v in fields]
    model = get_model_class("mmoe")(cols, device="cpu", config=cfg)
    sh = model.shard
    g = torch.Generator().manual_seed(5)
    tables = {n: torch.randn(v, 8, generator=g) for n, _, v in fields}
    parallel.load_full_tables(model, tables)
    shapes = {n: tuple(model.embedding_dict[n].weight.shape) for n, _, _ in fields}
    local_ok = all(torch.equal(model.embedding_dict[n].weight[:len(sh.owned_ids(v))], tables[n][sh.owned_ids(v)])
                   for n, _, v in fields)
    back = {n: parallel.full_table(model, n) for n, _, _ in fields}
    results[rank] = (shapes, local_ok, all(torch.equal(back[n], tables[n]) for n in tables),
                     [sh.local_rows(v) for _, _, v in fields])
    dist.destroy_process_group()


def test_sharded_table_layout_round_trips():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_shard_worker, args=(world, port, results), nprocs=world, join=True)
        res = dict(results)
    for r in range(world):
        shapes, local_ok, round_trip, rows = res[r]
        assert shapes == {"a": (6, 8), "b": (2, 8), "c": (1, 8)}, "every rank allocates ceil(V / R) rows per table"
        assert rows == [6, 2, 1]
        assert local_ok, "local row k of rank r holds global id r + k * R"
        assert round_trip, "all-gather + interleave rebuilds the full table bit for bit"


def test_shard_context_rejects_bad_rank_and_mode():
    import copy
    import pytest
    from mmlrec_b200 import synthetic
    from mmlrec_b200.model import get_model_class
    from mmlrec_b200.model.utils import SparseFeat
    from mmlrec_b200.parallel import ShardContext
    with pytest.raises(ValueError):
        ShardContext(2, 2)
    cfg, _ = synthetic.workload("synth26_mmoe", vocab=8)
    cfg = copy.deepcopy(cfg)
    cfg["b200_config"] = {"shard_tables": {"rank": 0, "world": 2, "gather": "bogus"}}
    with pytest.raises(ValueError):
        get_model_class("mmoe")([SparseFeat("a", 8, 8), SparseFeat("b", 8, 8)], device="cpu", config=cfg)


# ---------------------------------------------------------------------------------------------
# the data-parallel step itself on 2 gloo ranks: the planned program of each rank runs on the CPU (tests/plan_emulator.py)
# with the REAL DataParallelContext collectives -- Sync-BatchNorm's all-gather of per-rank moments and all-reduce of the
# backward sums, the SUM all-reduce of the dense gradient -- and must reproduce what the REFERENCE computed in one
# process on the whole batch (tests/golden/*.npz)
# ---------------------------------------------------------------------------------------------
def _dp_step_worker(rank, world, port, case, results):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import load_golden
    from plan_emulator import EmulatedPlan
    from test_plan_emulation_cpu import _load, _model
    from mmlrec_b200.parallel import DataParallelContext
    torch.set_num_threads(2)
    z, cfg, fields = load_golden(case)
    model = _model(cfg, fields, "fp32")
    ctx = DataParallelContext(rank, world)
    X, y = z["step0/X"], z["step0/y"]
    lo, hi = ctx.shard(X.shape[0])
    plan = EmulatedPlan(model, hi - lo, "fp32", dp=ctx)
    _load(model, z)
    plan.build()
    pred, loss = plan.forward_backward(X[lo:hi], y[lo:hi])
    grads = {n: plan.grad(p) for n, p in model.named_parameters() if getattr(p, "_mm_kind", "") == "dense"}
    bufs = {n: b.clone() for n, b in model.named_buffers()}
    results[rank] = (pred, loss, grads, bufs, [c for c in plan.fake.calls if c.startswith("bn_")])
    dist.destroy_process_group()


def _run_dp_case(case):
    from helpers import load_golden, rel_err
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_dp_step_worker, args=(world, port, case, results), nprocs=world, join=True)
        res = dict(results)
    z, cfg, _ = load_golden(case)
    pred = torch.cat([res[r][0] for r in range(world)])
    assert rel_err(pred, z["step0/pred"]) < 1e-5, "rank-major predictions == the single-process reference"
    total = sum(float(res[r][1][-1]) for r in range(world))
    assert abs(total - float(z["step0/loss"])) <= 1e-5 * abs(float(z["step0/loss"])), "sum-reduced loss"
    use_bn = cfg["model_config"].get("dnn_use_bn", False)
    for name, g in res[0][2].items():
        assert torch.equal(g, res[1][2][name]), f"{name}: the all-reduced gradient is the same on both ranks"
        key = "grad0/" + name
        if key not in z.files or (use_bn and ".linears." in name and name.endswith(".bias")):
            continue
        w = torch.from_numpy(z[key])
        assert float((g - w).abs().max()) <= 1e-5 * float(w.abs().max()) + 1e-9, name
    return res, z


def test_two_rank_step_with_sync_batchnorm_equals_the_reference():
    """census MMoE with BatchNorm: 2 ranks x 24 rows normalise over the GLOBAL 48 rows, like the reference's one process."""
    res, z = _run_dp_case("mmoe_census_bn_adam")
    calls = res[0][4]
    assert "bn_stats" in calls and "bn_combine" in calls and "bn_backward_sums" in calls and "bn_backward_synced" in calls
    # running statistics after the step: identical on both ranks and equal to the single-process program's at 48 rows
    # (momentum 0.1 towards the GLOBAL-batch moments, unbiased variance with the global count)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import load_golden
    from plan_emulator import EmulatedPlan
    from test_plan_emulation_cpu import _load, _model
    z, cfg, fields = load_golden("mmoe_census_bn_adam")
    model = _model(cfg, fields, "fp32")
    plan = EmulatedPlan(model, int(z["step0/X"].shape[0]), "fp32")
    _load(model, z)
    plan.build()
    plan.forward_backward(z["step0/X"], z["step0/y"])
    single = dict(model.named_buffers())
    moved = 0
    for name, buf in res[0][3].items():
        assert torch.equal(buf, res[1][3][name]), f"{name}: identical BatchNorm buffers on both ranks"
        if buf.dtype == torch.float32:
            assert float((buf - single[name]).abs().max()) <= 1e-6 * float(single[name].abs().max()) + 1e-9, name
            moved += int(not torch.equal(buf, torch.from_numpy(z["init/" + name])))
        else:
            assert int(buf) == int(single[name]) == 1, name
    assert moved > 0


def test_two_rank_step_of_ple_equals_the_reference():
    _run_dp_case("ple_ae_t4_adam")
