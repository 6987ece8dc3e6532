"""Data-parallel execution of the step on one multi-GPU box (one process per GPU, NCCL over NVLink/NVSwitch).

The reference has no functional multi-GPU path, so the contract is: R ranks with per-rank batch
b produce exactly the single-process step at the global batch R*b drawn in rank-major order.

  * dense parameters: local forward/backward, then ONE ``all_reduce(SUM)`` over the flat ``dense_grad`` bucket
    (SUM, not mean: the loss is ``reduction='sum'``, basemodel.py:294-296);
  * embedding tables (replicated): the id columns are all-gathered at the start of the step so every rank
    sorts the GLOBAL batch; after the local backward the ``d(dnn_input)`` rows are all-gathered and every rank
    runs the same K2 segmented reduce + fused row update (+ dense-Adam sweep) -> tables stay bit-identical;
  * BatchNorm statistics are per rank (no Sync-BN yet; only the census shape uses BatchNorm).

All collectives are issued on the step's stream through ``torch.distributed`` so they are captured in the
step's CUDA graph.  Row-sharded tables with id / row / row-gradient all-to-all (BASELINE config 5) are the next
item of SURVEY section 8(e) and are not implemented here.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


class DataParallelContext:
    def __init__(self, rank: int, world: int, group: Optional[dist.ProcessGroup] = None):
        self.rank, self.world, self.group = rank, world, group

    # ---- collectives (thin wrappers so the gloo CPU tests exercise exactly what the step calls)
    def gather_rows(self, local: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """out[r*b:(r+1)*b] = rank r's ``local`` rows (rank-major global batch order)."""
        assert out.shape[0] == local.shape[0] * self.world and out.shape[1:] == local.shape[1:]
        if dist.get_backend(self.group) == "gloo":  # CPU tests
            parts = list(out.chunk(self.world, dim=0))
            dist.all_gather(parts, local.contiguous(), group=self.group)
        else:
            dist.all_gather_into_tensor(out, local, group=self.group)
        return out

    def sum_gradients(self, flat_grad: torch.Tensor) -> torch.Tensor:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        return flat_grad

    def global_batch(self, b: int) -> int:
        return b * self.world

    def shard(self, n: int):
        """Row range of the global batch owned by this rank (used by input pipelines and tests)."""
        b = n // self.world
        return self.rank * b, (self.rank + 1) * b


def attach(model, rank: int, world: int, group: Optional[dist.ProcessGroup] = None) -> DataParallelContext:
    """Turn ``model`` into one replica of a data-parallel job.  Parameters must already be identical on every
    rank (same seed or ``broadcast_parameters``); plans are rebuilt with the collective-aware step."""
    ctx = DataParallelContext(rank, world, group)
    model.dp = ctx
    model._plans.clear()
    return ctx


def broadcast_parameters(model, src: int = 0, group: Optional[dist.ProcessGroup] = None) -> None:
    st = model.store
    for t in (st.dense, st.emb, st.stats, st.counts):
        dist.broadcast(t, src=src, group=group)
    st.refresh_bf16()
