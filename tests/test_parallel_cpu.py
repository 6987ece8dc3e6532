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
