"""Data-parallel execution of the step on one multi-GPU box (one process per GPU, NCCL over NVLink/NVSwitch).

The reference has no functional multi-GPU path, so the contract is: R ranks with per-rank batch
b produce exactly the single-process step at the global batch R*b drawn in rank-major order.

  * dense parameters: local forward/backward, then a SUM all-reduce of the flat dense gradient (SUM, not mean: the loss
    is ``reduction='sum'``, basemodel.py:294-296) -- with row-sharded tables a reduce-scatter + all-gather kernel over
    NVLink peer memory (csrc/peer.cu), otherwise NCCL on stage-boundary buckets overlapped with the backward pass;
  * embedding tables (replicated): the id columns are all-gathered at the start of the step so every rank
    sorts the GLOBAL batch; after the local backward the ``d(dnn_input)`` rows are all-gathered and every rank
    runs the same K2 segmented reduce + fused row update (+ dense-Adam sweep) -> tables stay bit-identical;
  * BatchNorm normalises over the GLOBAL batch (synchronised statistics: per-rank moments are all-gathered and
    combined, the backward sums all-reduced; engine/core.py LinearStage, csrc/fused_ops.cu bn_stats / bn_combine).

All collectives are issued on the step's stream through ``torch.distributed`` so they are captured in the
step's CUDA graph.

Row-sharded tables (BASELINE config 5): ``b200_config["shard_tables"] = {"rank": r, "world": R}``
builds every table as the shard ``rows r::R`` in IPC-exported device memory; ``attach_sharded`` opens the peers'
shards.  The id / row / row-gradient all-to-all exchanges are one-sided stores over NVLink peer memory
(csrc/peer.cu): ids are pushed to their owner, the owner serves its local rows into the requester's staging rows,
the backward pushes gradient rows to the owner, which sorted the request keys on a side stream meanwhile; barriers
are a flag exchange through peer memory, and the dense-gradient all-reduce doubles as the barrier before the
owner's K2.  ``shard_tables["gather"] = "peer_read"`` selects the barrier-free forward in which K1 reads rows
straight from the owners' shards (faster while the tables fit the peer TLB reach).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

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


# ----------------------------------------------------------------------------------------------
# row-sharded tables over NVLink peer memory
# ----------------------------------------------------------------------------------------------
class _RawCuda:
    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class PeerBuffer:
    """A cudaMalloc allocation exported with CUDA IPC (``mmlrec_peer_*``).  ``tensor()`` is a zero-copy torch view of
    the local memory; after ``exchange()`` ``peer_table`` is a device int64 vector with every rank's base pointer
    (this rank's own pointer at index ``rank``) that the kernels index by owner."""

    def __init__(self, nbytes: int, device: torch.device):
        from . import lib as L
        self._L, self.lib = L, L.load()
        self.nbytes, self.device = int(nbytes), device
        ptr = C.c_void_p()
        with torch.cuda.device(device):
            L.check(self.lib.mmlrec_peer_alloc(C.byref(ptr), self.nbytes), "peer_alloc")
        self.ptr = int(ptr.value)
        self.peer_ptrs: Optional[List[int]] = None
        self.peer_table: Optional[torch.Tensor] = None

    def tensor(self, dtype: torch.dtype) -> torch.Tensor:
        typestr, size = {torch.float32: ("<f4", 4), torch.int64: ("<i8", 8), torch.int32: ("<i4", 4)}[dtype]
        return torch.as_tensor(_RawCuda(self.ptr, self.nbytes // size, typestr), device=self.device)

    def exchange(self, rank: int, world: int, group=None) -> None:
        handle = C.create_string_buffer(64)
        with torch.cuda.device(self.device):
            self._L.check(self.lib.mmlrec_peer_export(self.ptr, handle), "peer_export")
            torch.cuda.synchronize()
            mine = torch.tensor(list(handle.raw), dtype=torch.uint8, device=self.device)
            every = torch.empty(world * 64, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(every, mine, group=group)
            raw = bytes(every.cpu().tolist())
            ptrs = []
            for r in range(world):
                if r == rank:
                    ptrs.append(self.ptr)
                    continue
                out = C.c_void_p()
                self._L.check(self.lib.mmlrec_peer_import(raw[r * 64:(r + 1) * 64], C.byref(out)), f"peer_import rank {r}")
                ptrs.append(int(out.value))
            self.peer_ptrs = ptrs
            self.peer_table = torch.tensor(ptrs, dtype=torch.int64, device=self.device)


class ShardContext:
    """owner(id) = id mod world; the owner keeps row ``id`` at local row ``id // world``."""

    def __init__(self, rank: int, world: int, group=None):
        if not (0 <= rank < world):
            raise ValueError(f"shard_tables: rank {rank} outside world {world}")
        self.rank, self.world, self.group = rank, world, group
        self.emb: Optional[PeerBuffer] = None
        self._token: Optional[torch.Tensor] = None
        self.flags: Optional[PeerBuffer] = None      # int32 [world]: barrier epochs written by the peers
        self.epoch: Optional[torch.Tensor] = None
        self.err: Optional[torch.Tensor] = None
        self.grad_in: Optional[PeerBuffer] = None    # dense-gradient all-reduce over peer memory (csrc/peer.cu)
        self.grad_out: Optional[PeerBuffer] = None

    def local_rows(self, vocabulary: int) -> int:
        return (vocabulary + self.world - 1) // self.world

    def owned_ids(self, vocabulary: int) -> torch.Tensor:
        """Global ids stored by this rank, in local-row order."""
        return torch.arange(self.rank, vocabulary, self.world)

    def alloc_emb(self, n_floats: int, device: torch.device) -> torch.Tensor:
        self.emb = PeerBuffer(4 * n_floats, device)
        return self.emb.tensor(torch.float32)

    def alloc_exchanged(self, nbytes: int, device: torch.device) -> PeerBuffer:
        buf = PeerBuffer(nbytes, device)
        buf.exchange(self.rank, self.world, self.group)
        return buf

    def connect(self) -> None:
        if self.emb is not None and self.emb.peer_table is None:
            dev = self.emb.device
            self.emb.exchange(self.rank, self.world, self.group)
            self._token = torch.zeros(1, device=dev)
            self.flags = self.alloc_exchanged(4 * max(self.world, 4), dev)
            self.epoch = torch.zeros(1, dtype=torch.int32, device=dev)
            self.err = torch.zeros(1, dtype=torch.int32, device=dev)

    def barrier(self) -> None:
        """Stream-ordered cross-rank barrier as a collective (4-byte all-reduce); set-up paths."""
        dist.all_reduce(self._token, group=self.group)

    def flag_barrier(self, stream: int) -> None:
        """Stream-ordered cross-rank barrier through peer memory (one tiny kernel, graph-capturable)."""
        self.emb._L.check(self.emb.lib.mmlrec_peer_barrier(self.flags.peer_table.data_ptr(), self.epoch.data_ptr(),
                                                           self.rank, self.world, self.err.data_ptr(), stream),
                          "peer_barrier")

    def connect_gradients(self, n_floats: int, device: torch.device) -> None:
        """Exchange buffers of the peer-memory dense-gradient all-reduce (collective; called by attach_sharded)."""
        if self.grad_in is None:
            self.grad_in = self.alloc_exchanged(4 * n_floats, device)
            self.grad_out = self.alloc_exchanged(4 * n_floats, device)
            self.grad_floats = n_floats

    def allreduce_gradients(self, stream: int) -> None:
        """grad_out (every rank) = sum over ranks of grad_in, in rank order; flag barriers on both sides."""
        self.flag_barrier(stream)
        self.emb._L.check(self.emb.lib.mmlrec_peer_allreduce_f32(self.grad_in.peer_table.data_ptr(),
                                                                 self.grad_out.peer_table.data_ptr(), self.grad_floats,
                                                                 self.rank, self.world, stream), "peer_allreduce")
        self.flag_barrier(stream)

    def check(self) -> None:
        """Raise if a barrier gave up waiting for a peer (host-side, outside the captured step)."""
        if int(self.err.item()) != 0:
            raise RuntimeError("row-sharded tables: a peer never reached the barrier (err flag set by mmlrec_peer_barrier)")


def attach_sharded(model, group: Optional[dist.ProcessGroup] = None) -> ShardContext:
    """Connect a model built with ``b200_config["shard_tables"]`` to its peers (collective: every rank calls it).
    Dense parameters are data-parallel exactly like ``attach``; the tables are row-sharded."""
    sh = model.shard
    if sh is None:
        raise ValueError('the model was not built with b200_config["shard_tables"]')
    sh.group = group
    model.dp = DataParallelContext(sh.rank, sh.world, group)
    sh.connect()
    import os
    if model.store is not None and os.environ.get("MMLREC_NCCL_ALLREDUCE") is None \
            and model.b200_config.get("peer_allreduce", True):
        sh.connect_gradients(model.store.slice_stride, model.device_obj)
    model._plans.clear()
    return sh


def load_full_tables(model, tables) -> None:
    """Fill the local shards from full ``[V, D]`` tables keyed by embedding name (tests / checkpoints)."""
    sh = model.shard
    with torch.no_grad():
        for name, full in tables.items():
            w = model.embedding_dict[name].weight
            rows = full[sh.rank::sh.world]
            w[:rows.shape[0]].copy_(rows.to(w.device))


def full_table(model, name: str) -> torch.Tensor:
    """All-gather one sharded table back into its ``[V, D]`` form (collective; tests / checkpoints)."""
    sh = model.shard
    w = model.embedding_dict[name].weight.detach()
    parts = [torch.empty_like(w) for _ in range(sh.world)]
    dist.all_gather(parts, w.contiguous(), group=sh.group)
    vocab = next(fc.vocabulary_size for fc in model.dnn_feature_columns if getattr(fc, "embedding_name", None) == name)
    out = torch.empty(sh.local_rows(vocab) * sh.world, w.shape[1], device=w.device)
    for r, part in enumerate(parts):
        out[r::sh.world] = part
    return out[:vocab]
