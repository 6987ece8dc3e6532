"""Static step program.

A model describes its step ONCE per batch size as a list of *stages* (gather, grouped linear
layers, gate-mix, heads) wired by ``Act`` handles (column slices of wide row-major device
buffers).  Building the program allocates every workspace buffer and uploads every kernel's
problem table; running it is a fixed sequence of C-ABI launches with no allocation, no host
synchronisation and no Python-side tensor math, so the whole step is captured in one CUDA graph.

Two arithmetic modes share the program:
  * ``fp32``  every GEMM on the CUDA cores (gemm_simt.cu) -- the parity mode (1e-5 vs the reference);
  * ``bf16``  every GEMM on the tcgen05 tensor cores (gemm_tc.cu): bf16 operands, fp32 accumulation
              in TMEM, fp32 master weights + bf16 shadow; activations that feed a GEMM are kept in
              bf16, activations that feed the gate / head / BatchNorm kernels in fp32 (a buffer gets
              exactly the copies its consumers asked for).

Gradient convention: a gradient buffer holds dL/d(pre-activation) of the stage that PRODUCED the
activation -- every kernel that writes an input gradient applies the producer's ReLU mask itself
(``Act.relu``), which fuses ``threshold_backward`` into the GEMM / gate / head epilogues.  The first
writer of a gradient assigns, later writers accumulate (``Act.grad_written``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .. import lib as L
from .store import FlatStore, _align


class ActGroup:
    """One wide buffer family [B, sum(widths)]: forward values and gradients, each optionally in
    fp32 and/or bf16 (decided by the consumers before ``materialize``)."""

    def __init__(self, widths: Sequence[int], relu: bool, name: str, grad_dtype: str, act: Optional[str] = None,
                 align: int = 1):
        # align: every member starts on a multiple of `align` columns (members whose widths are not multiples of 8 still
        # start on the 16-byte boundaries the tensor-core kernel's tensor maps need; APG's k, k*k, k wide outputs)
        self.widths, self.relu, self.name, self.grad_dtype = list(widths), relu, name, grad_dtype
        self.act = act if act is not None else ("relu" if relu else None)   # activation that produced the values
        self.need_f32 = self.need_bf16 = False
        self.need_grad = True
        self.want_bits = False      # a tensor-core dgrad wants this (ReLU) activation as a 1-bit mask
        self.bits = None            # int32 [ceil(B/32), chunks, 32]: layout of MmlrecGemmTcDesc.relu_bits_out
        self.bits_written = False   # set by the producing stage when its GEMM epilogue really emits the bits
        self.buf = self.buf16 = self.gbuf = self.gbuf16 = None
        self.acts: List["Act"] = []
        at = 0
        for i, w in enumerate(widths):
            self.acts.append(Act(self, at, w, f"{name}[{i}]"))
            at += w if i == len(widths) - 1 else _align(w, align)
        self.total = at

    def span(self) -> "Act":
        """All columns of the group as ONE activation (torch.cat of its members, cross_stitch.py:17: free here,
        the members are adjacent columns of one buffer)."""
        if any(x.col != sum(self.widths[:i]) for i, x in enumerate(self.acts)):
            raise NotImplementedError(f"{self.name}: members of widths {self.widths} are padded to 16-byte boundaries in "
                                      "bf16 mode; their concatenation is not contiguous (use widths that are multiples of 8 "
                                      "or the fp32 mode)")
        a = Act(self, 0, self.total, f"{self.name}[*]")
        a.members = list(self.acts)
        return a

    def materialize(self, b: "Builder") -> None:
        if not (self.need_f32 or self.need_bf16):  # produced but never consumed: keep one copy
            self.need_f32, self.need_bf16 = (not b.tc), b.tc
        if self.need_f32:
            self.buf = b.zeros(b.B, _align(self.total, 4))
        if self.need_bf16:
            self.buf16 = b.zeros(b.B, _align(self.total, 8), dtype=torch.bfloat16)
        if self.want_bits and self.need_grad and all(a.col % 32 == 0 for a in self.acts):
            self.bits_chunks = (self.total + 31) // 32
            self.bits = b.zeros((b.B + 31) // 32 * self.bits_chunks * 32, dtype=torch.int32)
        if self.need_grad:
            if self.grad_dtype == "f32":
                self.gbuf = b.zeros(b.B, _align(self.total, 4))
            else:
                self.gbuf16 = b.zeros(b.B, _align(self.total, 8), dtype=torch.bfloat16)


class Act:
    """[B, width] activation = columns [col, col+width) of its group's buffers."""

    def __init__(self, group: ActGroup, col: int, width: int, name: str):
        self.group, self.col, self.width, self.name = group, col, width, name
        self._grad_written = False
        self.need_f32 = self.need_bf16 = False   # which copies THIS activation's consumers read
        self.parent: Optional["Act"] = None      # set by sub(): this activation is a column range of `parent`
        self.children: List["Act"] = []          # column ranges handed out by sub()
        self.members: List["Act"] = []           # set by ActGroup.span(): this activation covers these whole ones

    # "some writer has produced (part of) this gradient": a range written through its sub-ranges counts, and writing
    # a span marks every activation it covers
    @property
    def grad_written(self) -> bool:
        return self._grad_written or any(c.grad_written for c in self.children)

    @grad_written.setter
    def grad_written(self, v: bool) -> None:
        self._grad_written = v
        for m in self.members:
            m.grad_written = v

    relu = property(lambda self: self.group.relu)
    # derivative folded into gradient writes: 0 none, 1 ReLU, 2 "2*sigmoid" (GateNN output)
    dkind = property(lambda self: {"relu": 1, "sigmoid2": 2}.get(self.group.act, 0))

    def sub(self, col: int, width: int) -> "Act":
        """A column sub-range of this activation (e.g. one field's embedding inside dnn_input)."""
        a = Act(self.group, self.col + col, width, f"{self.name}[{col}:{col + width}]")
        a.parent = self
        self.children.append(a)
        return a
    # fp32 view
    ld = property(lambda self: self.group.buf.stride(0))
    ptr = property(lambda self: self.group.buf.data_ptr() + 4 * self.col)
    has_f32 = property(lambda self: self.group.buf is not None)
    # bf16 view
    ld16 = property(lambda self: self.group.buf16.stride(0))
    ptr16 = property(lambda self: self.group.buf16.data_ptr() + 2 * self.col)
    has_bf16 = property(lambda self: self.group.buf16 is not None)
    # gradients
    grad_is_f32 = property(lambda self: self.group.grad_dtype == "f32")
    gld = property(lambda self: (self.group.gbuf if self.grad_is_f32 else self.group.gbuf16).stride(0))
    gptr = property(lambda self: (self.group.gbuf.data_ptr() + 4 * self.col) if self.grad_is_f32
                    else (self.group.gbuf16.data_ptr() + 2 * self.col))

    def same_as(self, o: "Act") -> bool:
        return self.group is o.group and self.col == o.col and self.width == o.width

    def tensor(self) -> torch.Tensor:
        return self.group.buf[:, self.col:self.col + self.width]

    def grad_tensor(self) -> torch.Tensor:
        g = self.group.gbuf if self.grad_is_f32 else self.group.gbuf16
        return g[:, self.col:self.col + self.width]

    def want(self, f32: bool = False, bf16: bool = False) -> "Act":
        self.need_f32 |= f32
        self.need_bf16 |= bf16
        self.group.need_f32 |= f32
        self.group.need_bf16 |= bf16
        if self.parent is not None:
            self.parent.want(f32, bf16)
        for m in self.members:
            m.want(f32, bf16)
        return self


class LinearSpec:
    """One nn.Linear (+ optional BatchNorm1d) as the program sees it."""

    def __init__(self, x: Act, linear: nn.Module, bn: Optional[nn.Module] = None, transposed: bool = False):
        self.x, self.linear, self.bn = x, linear, bn
        self.W: nn.Parameter = linear.weight
        self.b: Optional[nn.Parameter] = getattr(linear, "bias", None)
        # transposed: the parameter is stored [K, N] and applied as x @ W (cross_stitch.py:18) or x @ W + b (apg.py:92, :99)
        # instead of x @ W^T; its bias gradient comes from a column-sum kernel (the wgrad problem has dZ as the B operand)
        self.transposed = transposed
        self.N, self.K = (self.W.shape[1], self.W.shape[0]) if transposed else self.W.shape
        assert not (transposed and bn is not None), "transposed weights: x @ W [+ b] only"


class Builder:
    """Collects stages.  In ``dry`` mode nothing is allocated: only the order in which parameters
    and buffers are consumed is recorded (it becomes the flat-store layout)."""

    def __init__(self, B: int, device, store: Optional[FlatStore], dry: bool, precision: str = "fp32"):
        self.B, self.device, self.store, self.dry = B, device, store, dry
        self.tc = precision == "bf16"
        # tensor-core kernel: 2 = CTA pairs (tcgen05 cta_group::2, 256-wide tiles), 1 = one CTA per 128x128 tile
        import os
        self.tc_kernel = int(os.environ.get("MMLREC_TC_KERNEL", "2"))
        self.stages: List["Stage"] = []
        self.groups: List[ActGroup] = []
        self.param_order: List[nn.Parameter] = []
        self._noted = 0
        self.buffer_order: List[torch.Tensor] = []
        self.keep: List[object] = []       # tensors whose addresses are baked into tables
        self.aux_floats = 0                # dry run: size of the derived-weight region
        self.lib = None if dry else L.load()
        if store is not None:
            store.reset_aux()

    # ---- allocation helpers
    def zeros(self, *shape, dtype=torch.float32) -> torch.Tensor:
        t = torch.zeros(*shape, dtype=dtype, device=self.device)
        self.keep.append(t)
        return t

    def table(self, structs: Sequence[C.Structure]) -> torch.Tensor:
        t = torch.frombuffer(bytearray(L.struct_bytes(structs)), dtype=torch.uint8).to(self.device)
        self.keep.append(t)
        return t

    def ints(self, values: Sequence[int], dtype=torch.int32) -> torch.Tensor:
        t = torch.tensor(list(values), dtype=dtype, device=self.device)
        self.keep.append(t)
        return t

    def tc_table(self, descs: Sequence[L.GemmTcDesc]):
        """Encode tensor-core problems (tensor maps built on the host) -> launch tuple
        (records, prefix, n, tiles, tile order, per-unit starts, units); a unit is a CTA or a CTA pair."""
        pairs = self.tc_kernel == 2
        if pairs:
            descs = L.split_two_output_problems(descs)
        rb = int(self.lib.mmlrec_tc2_record_bytes() if pairs else self.lib.mmlrec_tc_record_bytes())
        host = (C.c_uint8 * (rb * len(descs)))()
        pre, at = [0], 0
        for i, d in enumerate(descs):
            if pairs:
                L.check(self.lib.mmlrec_tc2_encode_problem(C.byref(d), C.addressof(host) + i * rb), "tc2_encode_problem")
                at += int(self.lib.mmlrec_tc2_num_tiles(C.byref(d)))
            else:
                L.check(self.lib.mmlrec_tc_encode_problem(C.byref(d), C.addressof(host) + i * rb), "tc_encode_problem")
                at += int(self.lib.mmlrec_tc_num_tiles(d.M, d.N))
            pre.append(at)
        rec = torch.frombuffer(bytearray(bytes(host)), dtype=torch.uint8).to(self.device)
        assert rec.data_ptr() % 128 == 0
        self.keep.append(rec)
        order, starts, n_units = self._tc_schedule(descs, pre)
        return rec, self.ints(pre), len(descs), at, self.ints(order), self.ints(starts), n_units

    def tc_launch(self, tbl, stream, stamps=None):
        if self.tc_kernel == 2:
            return self.lib.mmlrec_gemm_grouped_tc2(tbl[0].data_ptr(), tbl[1].data_ptr(), tbl[2], tbl[3], tbl[4].data_ptr(),
                                                    tbl[5].data_ptr(), tbl[6], stamps, stream)
        if stamps is not None:
            return self.lib.mmlrec_gemm_grouped_tc_debug(tbl[0].data_ptr(), tbl[1].data_ptr(), tbl[2], tbl[3],
                                                         tbl[4].data_ptr(), tbl[5].data_ptr(), tbl[6], stamps, stream)
        return self.lib.mmlrec_gemm_grouped_tc_scheduled(tbl[0].data_ptr(), tbl[1].data_ptr(), tbl[2], tbl[3],
                                                         tbl[4].data_ptr(), tbl[5].data_ptr(), tbl[6], stream)

    def _tc_schedule(self, descs, pre):
        """Static longest-processing-time schedule of the launch's tiles over the SMs (CTA pairs for kernel 2).  Tile
        cost model in cycles, from the per-tile timelines under profiles/: a 64-wide k-block costs what its operand
        bytes cost at the L2 -> SM rate (~37 B/cycle/SM with every SM pulling), plus the epilogue of the tile."""
        import heapq
        n_sm = max(int(self.lib.mmlrec_tc_sm_count()), 1)
        pairs = self.tc_kernel == 2
        n_units = max(n_sm // 2, 1) if pairs else n_sm
        tiles = []
        for i, d in enumerate(descs):
            kb = (d.K + 63) // 64
            if pairs:
                # bytes one CTA of the pair moves for the tile (operand k-blocks in, its 128 x BN block out) at the
                # ~30 B/cycle/SM the L2 sustains with every SM pulling and pushing (profiles/tc_timeline_r02.txt:
                # forward K = 256 tiles 5.7-6.5 k cycles, K = 128 dgrad 4 k, 16-k-block wgrad 18-20 k)
                bn = 128 if (d.N <= 128 or (d.colsum and d.N > 240)) else 256
                out_b = 128 * min(bn, d.N) * (4 if d.C_f32 else 2)
                cost = (kb * (16384 + bn * 64) + out_b) // 30 + 600
            else:
                cost = 900 * kb + 2600
            tiles.extend((cost, t) for t in range(pre[i], pre[i + 1]))
        n_units = min(n_units, len(tiles))
        tiles.sort(key=lambda ct: (-ct[0], ct[1]))
        heap = [(0, c) for c in range(n_units)]
        lists = [[] for _ in range(n_units)]
        for cost, t in tiles:
            load, c = heapq.heappop(heap)
            lists[c].append(t)
            heapq.heappush(heap, (load + cost, c))
        order, starts = [], [0]
        for lst in lists:
            order.extend(sorted(lst))   # ascending tile index: consecutive tiles of a problem share operand rows in L2
            starts.append(len(order))
        return order, starts, n_units

    def aux_matrix(self, rows: int, cols: int):
        if self.dry:
            self.aux_floats = _align(self.aux_floats, 8) + rows * _align(cols, 8)
            return None
        return self.store.aux_matrix(rows, cols)

    def aux_vector(self, n: int):
        if self.dry:
            self.aux_floats += n
            return None
        return self.store.aux_vector(n)

    def new_group(self, widths: Sequence[int], relu: bool, name: str, grad_dtype: str = "f32",
                  act: Optional[str] = None, align: int = 1) -> List[Act]:
        # tensor-core mode: every member starts on a 16-byte boundary (8 bf16 / 8 fp32 columns) -- what the tensor maps of
        # the GEMMs reading or writing it need.  A no-op for widths that are multiples of 8 (every BASELINE shape); odd
        # widths (PEPNet on the 199-wide AliExpress input, APG's k-wide outputs) get pad columns nobody reads
        g = ActGroup(widths, relu, name, grad_dtype if self.tc else "f32", act, max(align, 8) if self.tc else align)
        self.groups.append(g)
        return g.acts

    def note_params(self, params: Sequence[Optional[nn.Parameter]]) -> None:
        for p in params:   # a parameter shared by several consumers (MLP's one final layer) is laid out once
            if isinstance(p, nn.Parameter) and not any(p is q for q in self.param_order):
                self.param_order.append(p)

    def note_buffers(self, bufs: Sequence[Optional[torch.Tensor]]) -> None:
        self.buffer_order.extend(t for t in bufs if t is not None)

    def add(self, stage: "Stage") -> "Stage":
        # the parameters a stage noted in its constructor (= a contiguous range of the flat store, in stage order)
        stage._params = self.param_order[self._noted:]
        self._noted = len(self.param_order)
        self.stages.append(stage)
        return stage

    def materialize(self) -> None:
        for g in self.groups:
            g.materialize(self)
        for s in self.stages:
            s.finalize()


class Stage:
    name = "stage"

    def finalize(self) -> None:
        """Called once after every buffer exists: build the forward tables."""

    def forward(self, stream: int, training: bool) -> None:
        raise NotImplementedError

    def plan_backward(self) -> None:
        """Called once, in reverse stage order, to build the backward tables."""

    def backward(self, stream: int) -> None:
        pass


# ----------------------------------------------------------------------------------------------
# K1 / K2
# ----------------------------------------------------------------------------------------------
class GatherStage(Stage):
    """Multi-field gather + concat (K1) and, in backward, sort + segmented reduce + fused row
    update (K2).  model/basemodel.py:461-487 + model/utils.py:434-446."""
    name = "gather"

    def __init__(self, b: Builder, model):
        self.b, self.model = b, model
        self.in_dim = model.input_dim_total
        (self.out,) = b.new_group([self.in_dim], relu=False, name="dnn_input", grad_dtype="f32")

    def finalize(self):
        b, model = self.b, self.model
        meta = []
        for f in model.embedding_layout:  # (param, vocab, x_col, out_col)
            meta += [f[0]._mm_off, f[1], f[2], f[3]]
        self.F_s, self.D = len(model.embedding_layout), model.emb_dim
        self.meta = b.ints(meta if meta else [0, 0, 0, 0], dtype=torch.int64)
        self.dense_cols = b.ints(model.dense_x_cols if model.dense_x_cols else [0])
        self.F_d = len(model.dense_x_cols)
        self.dense_out_col = self.F_s * self.D
        self.X = b.zeros(b.B, model.num_x_cols)
        self.oob = b.zeros(1, dtype=torch.int32)
        # data parallel: ids and d(dnn_input) of ALL ranks (the embedding update runs on the global batch)
        self.dp = getattr(model, "dp", None)
        self.sh = getattr(model, "shard", None)
        if self.sh is not None and (self.dp is None or self.sh.emb is None or self.sh.emb.peer_table is None):
            raise RuntimeError("row-sharded tables: call mmlrec_b200.parallel.attach_sharded(model, group) on every rank")
        self.B_all = self.dp.global_batch(b.B) if self.dp else b.B
        if self.sh is not None and self.F_s:
            # exchange buffers of the one-sided all-to-alls (csrc/peer.cu); collective: every rank builds its plan
            self.X_all = self.X
            row_w = self.F_s * self.D
            self.peer_read = self.sh.gather_mode == "peer_read"
            self.rq_keys = self.sh.alloc_exchanged(2 * self.F_s * self.B_all * 8, b.device)   # [2][F_s][B_all] u64
            self.rx_grad = self.sh.alloc_exchanged(self.B_all * row_w * 4, b.device)          # [B_all][F_s*D]
            self.rows_in = self.sh.alloc_exchanged(b.B * row_w * 4, b.device)                 # [b][F_s*D]
            L.check(b.lib.mmlrec_peer_fill_u64(self.rq_keys.ptr, 2 * self.F_s * self.B_all, (1 << 64) - 1, None), "rq fill")
            torch.cuda.synchronize()
            self.sh.barrier()   # every rank's sentinel fill is complete before anyone pushes
            torch.cuda.synchronize()
            self.ev_served = torch.cuda.Event()
        elif self.dp:
            self.X_all = b.zeros(self.B_all, model.num_x_cols)
            self.dgrad_all = b.zeros(self.B_all, self.out.group.gbuf.shape[1])
        else:
            self.X_all = self.X
        if self.F_s:
            n_pad = 32
            while n_pad < self.B_all:
                n_pad <<= 1
            self.sorted_ids = b.zeros(self.F_s, self.B_all, dtype=torch.int32)
            self.sorted_pos = b.zeros(self.F_s, self.B_all, dtype=torch.int32)
            self.keys_ws = b.zeros(self.F_s * n_pad, dtype=torch.int64)

    def forward(self, stream, training):
        b, st, o = self.b, self.b.store, self.out
        if self.sh is not None and self.F_s:
            sh, hy = self.sh, self.model.hyper_dev
            outs = (o.ptr if o.has_f32 else None, o.ld if o.has_f32 else 0,
                    o.ptr16 if o.has_bf16 else None, o.ld16 if o.has_bf16 else 0)
            # [ids all-to-all] every (sample, field) key goes to the row's owner (also feeds the owner's sort)
            L.check(b.lib.mmlrec_emb_push_ids(self.X.data_ptr(), self.X.stride(0), b.B, self.meta.data_ptr(), self.F_s,
                                              sh.rank, sh.world, self.B_all, self.rq_keys.peer_table.data_ptr(),
                                              hy.data_ptr(), 0, self.oob.data_ptr(), stream), "emb_push_ids")
            sh.flag_barrier(stream)
            if training and self.model.lazy_adam:   # lazy dense-Adam: the rows about to be served catch up first
                m = self.model
                L.check(b.lib.mmlrec_emb_adam_catch_up_keys(
                    self.rq_keys.ptr, self.B_all, self.meta.data_ptr(), self.F_s, self.D, st.emb.data_ptr(),
                    st.emb_s1.data_ptr(), st.emb_s2.data_ptr(), st.row_touch.data_ptr(), hy.data_ptr(), 0,
                    m.adam_hist.data_ptr(), m.adam_hist_cap, stream), "emb_adam_catch_up_keys")
            if self.peer_read:   # K1 reads the rows straight from the owners' shards
                L.check(b.lib.mmlrec_gather_concat_sharded(
                    self.X.data_ptr(), self.X.stride(0), b.B, sh.emb.peer_table.data_ptr(), sh.world,
                    self.meta.data_ptr(), self.F_s, self.D, self.dense_cols.data_ptr(), self.F_d, self.dense_out_col,
                    *outs, self.oob.data_ptr(), stream), "gather_concat_sharded")
            else:                # [rows all-to-all] the owner serves its local rows into the requesters' staging rows
                L.check(b.lib.mmlrec_emb_serve_rows(self.rq_keys.ptr, st.emb.data_ptr(), self.meta.data_ptr(), self.F_s,
                                                    self.D, b.B, self.B_all, self.rows_in.peer_table.data_ptr(),
                                                    hy.data_ptr(), 0, stream), "emb_serve_rows")
            self.ev_served.record(torch.cuda.current_stream())
            if not training:     # nobody sorts (= consumes) the keys outside a training step: reset this half now
                half = self.F_s * self.B_all
                par = int(self.model.hyper_host_step()) & 1
                L.check(b.lib.mmlrec_peer_fill_u64(self.rq_keys.ptr + par * half * 8, half, (1 << 64) - 1, stream), "rq reset")
            if not self.peer_read:
                sh.flag_barrier(stream)
                L.check(b.lib.mmlrec_gather_concat_staged(
                    self.X.data_ptr(), self.X.stride(0), b.B, self.rows_in.ptr, self.meta.data_ptr(), self.F_s, self.D,
                    self.dense_cols.data_ptr(), self.F_d, self.dense_out_col, *outs, stream), "gather_concat_staged")
            elif not training:
                sh.flag_barrier(stream)   # the reset above must not race with the peers' next push
            return
        L.check(b.lib.mmlrec_gather_concat(
            self.X.data_ptr(), self.X.stride(0), b.B, st.emb.data_ptr(), self.meta.data_ptr(), self.F_s, self.D,
            self.dense_cols.data_ptr(), self.F_d, self.dense_out_col,
            o.ptr if o.has_f32 else None, o.ld if o.has_f32 else 0,
            o.ptr16 if o.has_bf16 else None, o.ld16 if o.has_bf16 else 0,
            self.oob.data_ptr(), stream), "gather_concat")

    def sort(self, stream):
        """Side branch of the step: sort the batch ids; with Adam also stamp the touched rows and run
        the dense-Adam sweep of all UNtouched rows (zero-gradient update: needs no gradient, and the
        gather only reads touched rows, which the sweep skips)."""
        if not self.F_s:
            return
        b, st, hy = self.b, self.b.store, self.model.hyper_dev
        if self.sh is not None:   # the owner sorts the request keys it received in the forward exchange
            L.check(b.lib.mmlrec_sort_field_keys(self.rq_keys.ptr, self.B_all, self.F_s, hy.data_ptr(),
                                                 self.sorted_ids.data_ptr(), self.sorted_pos.data_ptr(),
                                                 self.keys_ws.data_ptr(), stream), "sort_field_keys")
        else:
            L.check(b.lib.mmlrec_sort_field_ids(self.X_all.data_ptr(), self.X_all.stride(0), self.B_all,
                                                self.meta.data_ptr(), self.F_s, self.sorted_ids.data_ptr(),
                                                self.sorted_pos.data_ptr(), self.keys_ws.data_ptr(), stream),
                    "sort_field_ids")
        if self.model.optimizer_name in ("adam", "rmsprop") and not self.model.lazy_adam:   # zero-gradient update is not a no-op
            L.check(b.lib.mmlrec_emb_stamp_rows(self.sorted_ids.data_ptr(), self.meta.data_ptr(), self.F_s, self.B_all, self.D,
                                                st.row_touch.data_ptr(), hy.data_ptr(), stream), "emb_stamp_rows")
            L.check(b.lib.mmlrec_emb_adam_dense_sweep(st.emb.data_ptr(), st.emb_s1.data_ptr(),
                                                      st.emb_s2.data_ptr() if st.emb_s2 is not None else None,
                                                      st.row_touch.data_ptr(), st.n_emb // self.D, self.D,
                                                      hy.data_ptr(), stream), "emb_adam_dense_sweep")

    def catch_up(self, stream):
        """Lazy dense-Adam: the rows this step reads are brought up to the previous optimizer step (start of the step)."""
        if not self.F_s or not self.model.lazy_adam or self.sh is not None:
            return   # (row-sharded tables: the owner catches up on the request keys, see forward)
        b, st, m = self.b, self.b.store, self.model
        L.check(b.lib.mmlrec_emb_adam_catch_up(self.X_all.data_ptr(), self.X_all.stride(0), self.B_all, self.meta.data_ptr(),
                                               self.F_s, self.D, st.emb.data_ptr(), st.emb_s1.data_ptr(), st.emb_s2.data_ptr(),
                                               st.row_touch.data_ptr(), m.hyper_dev.data_ptr(), m.adam_hist.data_ptr(),
                                               m.adam_hist_cap, stream), "emb_adam_catch_up")

    def backward(self, stream):
        if not self.F_s or not self.out.grad_written:
            return
        b, st, hy = self.b, self.b.store, self.model.hyper_dev
        p = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        d_ptr, d_ld = self.out.gptr, self.out.gld
        if self.sh is not None:  # [row-grad all-to-all] every local (sample, field) gradient row goes to its owner
            L.check(b.lib.mmlrec_emb_push_grads(
                self.X.data_ptr(), self.X.stride(0), b.B, d_ptr, d_ld, self.meta.data_ptr(), self.F_s, self.D,
                self.sh.rank, self.sh.world, self.B_all, self.rx_grad.peer_table.data_ptr(), stream), "emb_push_grads")
            return
        if self.dp:  # every rank reduces the gradient rows of the GLOBAL batch
            self.dp.gather_rows(self.out.group.gbuf, self.dgrad_all)
            d_ptr, d_ld = self.dgrad_all.data_ptr(), self.dgrad_all.stride(0)
        L.check(b.lib.mmlrec_emb_backward_update(
            d_ptr, d_ld, self.B_all, self.sorted_ids.data_ptr(), self.sorted_pos.data_ptr(),
            self.meta.data_ptr(), self.F_s, self.D, st.emb.data_ptr(), p(st.emb_s1), p(st.emb_s2), p(st.row_touch),
            hy.data_ptr(), None, stream), "emb_backward_update")


    def post_reduce(self, stream):
        """Sharded tables, after the dense-gradient all-reduce (= every gradient row has landed): segmented
        reduce + fused row update of the rows this rank owns, over the ids it sorted during forward/backward."""
        if self.sh is None or not self.F_s or not self.out.grad_written:
            return
        b, st, hy = self.b, self.b.store, self.model.hyper_dev
        p = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        L.check(b.lib.mmlrec_emb_backward_update_sharded(
            self.rx_grad.ptr, self.F_s * self.D, self.B_all, self.sorted_ids.data_ptr(), self.sorted_pos.data_ptr(),
            self.meta.data_ptr(), self.F_s, self.D, st.emb.data_ptr(), p(st.emb_s1), p(st.emb_s2), p(st.row_touch),
            hy.data_ptr(), stream), "emb_backward_update_sharded")


# ----------------------------------------------------------------------------------------------
# K3: grouped linear layers
# ----------------------------------------------------------------------------------------------
class _Group:
    """Adjacent members of a stage that read the same input and whose parameters are contiguous:
    executed as one wide problem."""

    def __init__(self, members: List[LinearSpec], outs: List[Act]):
        self.members, self.outs = members, outs
        self.y_col = outs[0].col
        self.x = members[0].x
        self.K = members[0].K
        self.N = sum(m.N for m in members)
        self.W, self.b = members[0].W, members[0].b
        self.transposed = members[0].transposed
        assert not self.transposed or len(members) == 1


def _tile_prefix_f32(problems: Sequence[L.GemmF32]) -> Tuple[List[int], int]:
    pre, at = [0], 0
    for p in problems:
        at += ((p.M + 63) // 64) * ((p.N + 63) // 64)
        pre.append(at)
    return pre, at


class LinearStage(Stage):
    """A set of independent ``y = act([bn](x W^T + b))`` layers run by ONE grouped-GEMM launch
    (model/utils.py:146-161), and in backward by one launch holding every wgrad + dgrad problem."""
    name = "linear"

    def __init__(self, b: Builder, specs: List[LinearSpec], act: Optional[str], label: str = "", align_outs: int = 1):
        self.b, self.specs, self.act, self.label = b, specs, act, label
        self.use_bn = specs[0].bn is not None
        assert all((s.bn is not None) == self.use_bn for s in specs)
        # parameter order == flat-store layout: weights, biases, bn gammas, bn betas
        b.note_params([s.W for s in specs])
        b.note_params([s.b for s in specs])
        if self.use_bn:
            b.note_params([s.bn.weight for s in specs])
            b.note_params([s.bn.bias for s in specs])
            b.note_buffers([s.bn.running_mean for s in specs])
            b.note_buffers([s.bn.running_var for s in specs])
            b.note_buffers([s.bn.num_batches_tracked for s in specs])
        widths = [s.N for s in specs]
        # without BatchNorm the backward GEMMs read dZ = d(out) directly (bf16 in tensor-core mode);
        # with BatchNorm d(out) feeds the BN backward kernel (fp32) which emits dZ
        self.outs = b.new_group(widths, relu=(act == "relu"), name=f"{label}.y",
                                grad_dtype="f32" if self.use_bn else "bf16", act=act, align=align_outs)
        self.zs = b.new_group(widths, relu=False, name=f"{label}.z", grad_dtype="bf16", align=align_outs) if self.use_bn else None
        if self.use_bn:
            self.zs[0].want(f32=True)
        for s in specs:
            s.x.want(f32=not b.tc, bf16=b.tc)
            if b.tc and b.tc_kernel == 2 and s.x.relu:
                s.x.group.want_bits = True   # dgrad through the producer's ReLU reads 1 bit per element, not the bf16 value

    # ---- forward tables
    def finalize(self):
        b, st, specs = self.b, self.b.store, self.specs
        self.groups: List[_Group] = []
        cur: List[LinearSpec] = []

        def flush():
            if cur:
                self.groups.append(_Group(list(cur), [self.outs[specs.index(m)] for m in cur]))

        for s in specs:
            ok = bool(cur) and not s.transposed and not cur[-1].transposed \
                and self.outs[specs.index(s)].col == self.outs[specs.index(cur[-1])].col + cur[-1].N \
                and s.x.same_as(cur[-1].x) and s.K == cur[-1].K and st.contiguous_after(cur[-1].W, s.W) \
                and ((s.b is None) == (cur[-1].b is None)) and (s.b is None or st.contiguous_after(cur[-1].b, s.b))
            if ok and self.use_bn:
                p, q = cur[-1].bn, s.bn
                ok = st.contiguous_after(p.weight, q.weight) and st.contiguous_after(p.bias, q.bias) \
                    and p.running_mean._mm_off + cur[-1].N == q.running_mean._mm_off \
                    and p.running_var._mm_off + cur[-1].N == q.running_var._mm_off \
                    and p.num_batches_tracked._mm_off + 1 == q.num_batches_tracked._mm_off
            if not ok:
                flush()
                cur.clear()
            cur.append(s)
        flush()
        yg = self.outs[0].group
        tgt = self.zs[0].group if self.use_bn else yg
        act_code = L.ACT_NONE if self.use_bn else L.ACT_CODES[self.act]
        if not b.tc:
            probs = []
            for g in self.groups:
                p = L.GemmF32()
                p.A, p.a_rs, p.a_cs = g.x.ptr, g.x.ld, 1
                p.B, p.b_rs, p.b_cs = (g.W.data_ptr(), 1, g.W._mm_ld) if g.transposed else (g.W.data_ptr(), g.W._mm_ld, 1)
                p.C, p.ldc = tgt.buf.data_ptr() + 4 * g.y_col, tgt.buf.stride(0)
                p.bias = g.b.data_ptr() if g.b is not None else None
                p.M, p.N, p.K, p.act = b.B, g.N, g.K, act_code
                probs.append(p)
            pre, tiles = _tile_prefix_f32(probs)
            self.fwd = [(b.table(probs), b.ints(pre), len(probs), tiles)]
        else:
            # forward problems may be narrower than the merged group: a run of members whose consumers
            # read the same copies (e.g. expert hidden -> bf16 only, gate hidden -> fp32 only) becomes one
            # problem, so no column is written in a precision nobody reads
            descs = []
            for g in self.groups:
                runs, at = [], 0
                for m, o in zip(g.members, g.outs):
                    if self.use_bn:
                        key = (True, False)  # the pre-BatchNorm buffer is fp32
                    else:
                        key = (o.need_f32, o.need_bf16)
                        if not (key[0] or key[1]):  # produced but never read: write whichever copy exists
                            key = (tgt.buf16 is None, tgt.buf16 is not None)
                    if runs and runs[-1][0] == key:
                        runs[-1][2] += m.N
                    else:
                        runs.append([key, at, m.N])
                    at += m.N
                for (f32, bf16), off, n in runs:
                    d = L.GemmTcDesc()
                    d.A, d.lda, d.a_mn_major = g.x.ptr16, g.x.ld16, 0
                    if g.transposed:   # the array is [K, N]: read MN-major
                        d.B, d.ldb, d.b_mn_major = st.bf16_ptr(g.W), g.W._mm_ld, 1
                    else:
                        d.B, d.ldb, d.b_mn_major = st.bf16_ptr(g.W) + 2 * off * g.W._mm_ld, g.W._mm_ld, 0
                    d.M, d.N, d.K = b.B, n, g.K
                    col = g.y_col + off
                    if f32:
                        d.C_f32, d.ldc_f32 = tgt.buf.data_ptr() + 4 * col, tgt.buf.stride(0)
                    if bf16:
                        d.C_bf16, d.ldc_bf16 = tgt.buf16.data_ptr() + 2 * col, tgt.buf16.stride(0)
                    d.bias = (g.b.data_ptr() + 4 * off) if g.b is not None else None
                    d.act = act_code
                    if tgt.bits is not None and self.act == "relu" and not self.use_bn and col % 32 == 0:
                        d.relu_bits_out, d.bits_out_chunks, d.bits_out_chunk0 = tgt.bits.data_ptr(), tgt.bits_chunks, col // 32
                        tgt.bits_written = True
                    descs.append(d)
            self.fwd = self._tc_tables(descs)
        if self.use_bn:
            n_total = _align(yg.total, 4)
            self.save_mean, self.save_invstd = b.zeros(n_total), b.zeros(n_total)
            # data parallel: BatchNorm statistics of the GLOBAL batch (what the single-process reference normalises over)
            self.dp = getattr(b, "dp", None)
            if self.dp is not None and self.dp.world > 1:
                self.bn_stats = [(b.zeros(1, 2 * g.N), b.zeros(self.dp.world, 2 * g.N)) for g in self.groups]
                self.bn_sums = [b.zeros(2 * g.N) for g in self.groups]
            else:
                self.dp = None

    MAX_TC_PROBLEMS = 96  # the tensor-core kernel caches its tile table in shared memory

    def _tc_tables(self, descs):
        return [self.b.tc_table(descs[i:i + self.MAX_TC_PROBLEMS]) for i in range(0, len(descs), self.MAX_TC_PROBLEMS)]

    def _launch(self, tbl, stream, what):
        lib = self.b.lib
        if self.b.tc:
            rc = self.b.tc_launch(tbl, stream)
        else:
            rc = lib.mmlrec_gemm_grouped_f32(tbl[0].data_ptr(), tbl[1].data_ptr(), tbl[2], tbl[3], stream)
        L.check(rc, f"{what} {self.label}")

    def forward(self, stream, training):
        b = self.b
        for tbl in self.fwd:
            self._launch(tbl, stream, "linear fwd")
        if self.use_bn:
            zg, yg = self.zs[0].group, self.outs[0].group
            for gi, g in enumerate(self.groups):
                bn0, c = g.members[0].bn, g.y_col
                mode = 1 if training else 0
                if training and self.dp is not None:
                    local, every = self.bn_stats[gi]
                    L.check(b.lib.mmlrec_bn_stats(zg.buf.data_ptr() + 4 * c, zg.buf.stride(0), b.B, g.N, local.data_ptr(),
                                                  stream), f"bn stats {self.label}")
                    self.dp.gather_rows(local, every)
                    L.check(b.lib.mmlrec_bn_combine(every.data_ptr(), self.dp.world, b.B, g.N, bn0.running_mean.data_ptr(),
                                                    bn0.running_var.data_ptr(), bn0.num_batches_tracked.data_ptr(),
                                                    len(g.members), self.save_mean.data_ptr() + 4 * c,
                                                    self.save_invstd.data_ptr() + 4 * c, stream), f"bn combine {self.label}")
                    mode = 2
                L.check(b.lib.mmlrec_bn_forward(
                    zg.buf.data_ptr() + 4 * c, zg.buf.stride(0), b.B, g.N, bn0.weight.data_ptr(), bn0.bias.data_ptr(),
                    bn0.running_mean.data_ptr(), bn0.running_var.data_ptr(), bn0.num_batches_tracked.data_ptr(),
                    len(g.members), self.save_mean.data_ptr() + 4 * c, self.save_invstd.data_ptr() + 4 * c,
                    (yg.buf.data_ptr() + 4 * c) if yg.buf is not None else None,
                    yg.buf.stride(0) if yg.buf is not None else 0,
                    (yg.buf16.data_ptr() + 2 * c) if yg.buf16 is not None else None,
                    yg.buf16.stride(0) if yg.buf16 is not None else 0,
                    L.ACT_CODES[self.act], mode, stream), f"bn fwd {self.label}")

    # ---- backward
    def plan_backward(self):
        b, st = self.b, self.b.store
        self.live_groups: List[_Group] = []
        waves: List[list] = [[]]
        # deterministic split-K of the wgrad problems (tensor-core mode): the contraction runs over the batch, so an
        # unsplit tile is B/64 k-blocks long while the launch has only a handful of them -- far fewer than SMs.  The batch
        # is cut into S slices; every slice is its own problem that writes its partial tile into gradient slice k of the
        # flat store (one bulk tensor store per box: a partial tile costs little) and the dense optimizer adds the
        # slices up in slice order -- no reduction pass, no atomics.  Kernel 1 (one CTA per tile) keeps the old rule:
        # only launches with a handful of tiles are split.  Derived weights (STAR) are consumed by a fold kernel that
        # reads slice 0 only, so their stages are not split.
        self.split_k = 1
        derived = any(g.W._mm_off >= st.aux_base for g in self.groups)
        if b.tc and b.B % 1024 == 0 and b.B >= 2048 and not derived:
            if b.tc_kernel == 2:
                self.split_k = min(b.B // 1024, st.max_grad_slices)
            else:
                wg_tiles = sum(((g.N + 127) // 128) * ((g.K + 127) // 128)
                               for g in self.groups if any(o.grad_written for o in g.outs))
                for cand in (4, 2):
                    if cand <= min(b.B // 1024, st.max_grad_slices) and 0 < wg_tiles * cand <= 32:
                        self.split_k = cand
                        break
        self.zero_before_bwd = []   # gradient buffers a K-split dgrad accumulates into from zero
        self.colsums = []           # (dZ fp32 ptr, dZ bf16 ptr, ld, N, bias-gradient ptr) of [K, N]-stored layers with a bias
        dzg = self.zs[0].group if self.use_bn else self.outs[0].group   # where dZ lives
        # one Linear applied to several inputs in this stage (AITM's h1/h2/h3 on both tokens, aitm.py:86-88): every
        # application writes its weight / bias gradient into its OWN gradient slices (the optimizer adds slices up in
        # slice order anyway), so the launch holds no two problems with the same destination
        live = [g for g in self.groups if any(o.grad_written for o in g.outs)]
        uses: Dict[int, int] = {}
        for g in live:
            uses[g.W.data_ptr()] = uses.get(g.W.data_ptr(), 0) + 1
        dup = max(uses.values(), default=1)
        if dup > 1:
            assert dup <= st.max_grad_slices and not derived, "too many applications of one Linear in a stage"
            self.split_k = max(min(self.split_k, st.max_grad_slices // dup), 1)
        S = self.split_k
        self.split_k = S * dup      # gradient slices this stage writes (StepPlan.grad_slices)
        seen: Dict[int, int] = {}
        # tensor-core mode, dZ kept in fp32 because several stages add into it (an activation with consumers in
        # different stages): the GEMMs read a bf16 copy made at the start of this stage's backward
        self.dz_stage = None
        if b.tc and not self.use_bn and dzg.gbuf16 is None and live:
            self.dz_stage = b.zeros(b.B, _align(dzg.total, 8), dtype=torch.bfloat16)
        for g in self.groups:
            if not any(o.grad_written for o in g.outs):
                continue  # nothing flows into this group: its parameters keep a zero gradient
            slice0 = seen.get(g.W.data_ptr(), 0) * S     # first gradient slice of this application
            seen[g.W.data_ptr()] = seen.get(g.W.data_ptr(), 0) + 1
            g_off = 4 * slice0 * st.slice_stride
            for o in g.outs:
                for part in ([o] if not o.children else o.children):
                    if not part.grad_written:
                        part.grad_tensor().zero_()  # stays zero: the buffer is never written afterwards
            self.live_groups.append(g)
            x = g.x
            want_dx = x.group.need_grad
            wave = sum(1 for gg in self.live_groups[:-1] if gg.x.same_as(x)) if want_dx else 0
            while len(waves) <= wave:
                waves.append([])
            accumulate = 1 if (want_dx and (x.grad_written or wave > 0)) else 0
            if g.transposed and g.b is not None:
                if b.tc:
                    zb = self.dz_stage if self.dz_stage is not None else dzg.gbuf16
                    self.colsums.append((None, zb.data_ptr() + 2 * g.y_col, zb.stride(0), g.N, st.grad_ptr(g.b) + g_off))
                else:
                    self.colsums.append((dzg.gbuf.data_ptr() + 4 * g.y_col, None, dzg.gbuf.stride(0), g.N,
                                         st.grad_ptr(g.b) + g_off))
            if not b.tc:
                dz_ptr, dz_ld = dzg.gbuf.data_ptr() + 4 * g.y_col, dzg.gbuf.stride(0)
                p = L.GemmF32()   # wgrad: dW[n,k] = sum_b dZ[b,n] X[b,k]; rowsum_a = bias gradient
                if g.transposed:  # the parameter is [K, N]: dW[k,n] = sum_b X[b,k] dZ[b,n]
                    p.A, p.a_rs, p.a_cs = x.ptr, 1, x.ld
                    p.B, p.b_rs, p.b_cs = dz_ptr, 1, dz_ld
                    p.M, p.N, p.K = g.K, g.N, b.B
                else:
                    p.A, p.a_rs, p.a_cs = dz_ptr, 1, dz_ld
                    p.B, p.b_rs, p.b_cs = x.ptr, 1, x.ld
                    p.rowsum_a = (st.grad_ptr(g.b) + g_off) if g.b is not None else None
                    p.M, p.N, p.K = g.N, g.K, b.B
                p.C, p.ldc = st.grad_ptr(g.W) + g_off, g.W._mm_ld
                waves[0].append(p)
                if want_dx:
                    q = L.GemmF32()   # dgrad: dX[b,k] = sum_n dZ[b,n] W[n,k], masked by the producer's ReLU
                    q.A, q.a_rs, q.a_cs = dz_ptr, dz_ld, 1
                    q.B, q.b_rs, q.b_cs = (g.W.data_ptr(), g.W._mm_ld, 1) if g.transposed else (g.W.data_ptr(), 1, g.W._mm_ld)
                    q.C, q.ldc = x.gptr, x.gld
                    q.M, q.N, q.K = b.B, g.K, g.N
                    assert x.dkind in (0, 1), "a GEMM cannot back-propagate into a sigmoid output directly"
                    if x.relu:
                        q.mask, q.ldmask = x.ptr, x.ld
                    q.accumulate = accumulate
                    waves[wave].append(q)
            else:
                dz_buf = self.dz_stage if self.dz_stage is not None else dzg.gbuf16
                dz16, dz_ld = dz_buf.data_ptr() + 2 * g.y_col, dz_buf.stride(0)
                # wgrad: both operands MN-major (no transposed copies); batch slice k -> gradient slice slice0 + k
                # a layer with few outputs: dW^T = X^T dZ (M = K_in fills the 256-row pair tiles; with M = N_out <= 128
                # half of every tile is empty and the peer CTA's operand path idles), stored transposed from registers;
                # the bias gradient = column sums of dZ from an all-ones A tile (MmlrecGemmTcDesc.colsum_b)
                swapped = b.tc_kernel == 2 and not g.transposed and g.N <= 128 and g.K >= g.N \
                    and os.environ.get("MMLREC_NO_WGRAD_SWAP") is None
                for k in range(S):
                    rows = b.B // S
                    d = L.GemmTcDesc()
                    if swapped:
                        d.A, d.lda, d.a_mn_major = x.ptr16 + 2 * k * rows * x.ld16, x.ld16, 1
                        d.B, d.ldb, d.b_mn_major = dz16 + 2 * k * rows * dz_ld, dz_ld, 1
                        d.M, d.N, d.K = g.K, g.N, rows
                        d.C_f32, d.ldc_f32 = st.grad_ptr(g.W) + g_off + 4 * k * st.slice_stride, g.W._mm_ld
                        d.c_transposed = 1
                        d.colsum_b = (st.grad_ptr(g.b) + g_off + 4 * k * st.slice_stride) if g.b is not None else None
                        waves[0].append(d)
                        continue
                    if g.transposed:   # dW[k,n] for a [K, N] parameter: the operands swap roles
                        d.A, d.lda, d.a_mn_major = x.ptr16 + 2 * k * rows * x.ld16, x.ld16, 1
                        d.B, d.ldb, d.b_mn_major = dz16 + 2 * k * rows * dz_ld, dz_ld, 1
                        d.M, d.N, d.K = g.K, g.N, rows
                    else:
                        d.A, d.lda, d.a_mn_major = dz16 + 2 * k * rows * dz_ld, dz_ld, 1
                        d.B, d.ldb, d.b_mn_major = x.ptr16 + 2 * k * rows * x.ld16, x.ld16, 1
                        d.M, d.N, d.K = g.N, g.K, rows
                        d.colsum = (st.grad_ptr(g.b) + g_off + 4 * k * st.slice_stride) if g.b is not None else None
                    d.C_f32, d.ldc_f32 = st.grad_ptr(g.W) + g_off + 4 * k * st.slice_stride, g.W._mm_ld
                    waves[0].append(d)
                if want_dx:
                    e = L.GemmTcDesc()   # dgrad: A = dZ (K-major), B = W read MN-major
                    e.A, e.lda, e.a_mn_major = dz16, dz_ld, 0
                    e.B, e.ldb, e.b_mn_major = st.bf16_ptr(g.W), g.W._mm_ld, 0 if g.transposed else 1
                    e.M, e.N, e.K = b.B, g.K, g.N
                    if x.grad_is_f32:
                        e.C_f32, e.ldc_f32 = x.gptr, x.gld
                        e.accumulate = accumulate
                    else:
                        assert not accumulate, "a bf16 gradient buffer cannot be accumulated into"
                        e.C_bf16, e.ldc_bf16 = x.gptr, x.gld
                    if x.relu:
                        e.mask, e.ldmask = x.ptr16, x.ld16
                        xg = x.group
                        if xg.bits is not None and xg.bits_written and x.col % 32 == 0:
                            e.mask_bits, e.mask_bits_chunks, e.mask_bits_chunk0 = xg.bits.data_ptr(), xg.bits_chunks, x.col // 32
                    if (b.tc_kernel == 2 and g.N >= 2048 and x.grad_is_f32 and not accumulate and not g.transposed
                            and x.width == x.group.total):
                        # a very long contraction (K = the merged layer width) on a few tiles: two K halves as two
                        # problems that ADD into the zeroed gradient buffer (0 + a + b is order-independent: deterministic)
                        self.zero_before_bwd.append(x.group.gbuf)
                        h0 = (g.N // 2 + 63) // 64 * 64
                        for k0, kn in ((0, h0), (h0, g.N - h0)):
                            e2 = L.GemmTcDesc()
                            e2.A, e2.lda, e2.a_mn_major = dz16 + 2 * k0, dz_ld, 0
                            e2.B, e2.ldb, e2.b_mn_major = st.bf16_ptr(g.W) + 2 * k0 * g.W._mm_ld, g.W._mm_ld, 1
                            e2.M, e2.N, e2.K = b.B, g.K, kn
                            e2.C_f32, e2.ldc_f32, e2.accumulate = x.gptr, x.gld, 1
                            e2.mask, e2.ldmask = e.mask, e.ldmask
                            e2.mask_bits, e2.mask_bits_chunks, e2.mask_bits_chunk0 = e.mask_bits, e.mask_bits_chunks, e.mask_bits_chunk0
                            waves[wave].append(e2)
                    else:
                        waves[wave].append(e)
            if want_dx:
                x.grad_written = True
        self.bwd = []
        for wv in waves:
            if not wv:
                continue
            if b.tc:
                self.bwd.extend(self._tc_tables(wv))
            else:
                pre, tiles = _tile_prefix_f32(wv)
                self.bwd.append((b.table(wv), b.ints(pre), len(wv), tiles))

    def backward(self, stream):
        b = self.b
        if self.use_bn:
            zg, yg = self.zs[0].group, self.outs[0].group
            for g in self.live_groups:
                bn0, c = g.members[0].bn, g.y_col
                if self.dp is not None:
                    sums = self.bn_sums[self.groups.index(g)]
                    dz32 = ((zg.gbuf.data_ptr() + 4 * c) if zg.gbuf is not None else None,
                            zg.gbuf.stride(0) if zg.gbuf is not None else 0)
                    dz16 = ((zg.gbuf16.data_ptr() + 2 * c) if zg.gbuf16 is not None else None,
                            zg.gbuf16.stride(0) if zg.gbuf16 is not None else 0)
                    L.check(b.lib.mmlrec_bn_backward_sums(
                        yg.gbuf.data_ptr() + 4 * c, yg.gbuf.stride(0), zg.buf.data_ptr() + 4 * c, zg.buf.stride(0), b.B, g.N,
                        self.save_mean.data_ptr() + 4 * c, self.save_invstd.data_ptr() + 4 * c, sums.data_ptr(),
                        b.store.grad_ptr(bn0.weight), b.store.grad_ptr(bn0.bias), stream), f"bn bwd sums {self.label}")
                    self.dp.sum_gradients(sums)
                    L.check(b.lib.mmlrec_bn_backward_synced(
                        yg.gbuf.data_ptr() + 4 * c, yg.gbuf.stride(0), zg.buf.data_ptr() + 4 * c, zg.buf.stride(0), b.B, g.N,
                        bn0.weight.data_ptr(), self.save_mean.data_ptr() + 4 * c, self.save_invstd.data_ptr() + 4 * c,
                        dz32[0], dz32[1], dz16[0], dz16[1], sums.data_ptr(), b.B * self.dp.world, stream),
                        f"bn bwd {self.label}")
                    continue
                L.check(b.lib.mmlrec_bn_backward(
                    yg.gbuf.data_ptr() + 4 * c, yg.gbuf.stride(0), zg.buf.data_ptr() + 4 * c, zg.buf.stride(0), b.B, g.N,
                    bn0.weight.data_ptr(), self.save_mean.data_ptr() + 4 * c, self.save_invstd.data_ptr() + 4 * c,
                    (zg.gbuf.data_ptr() + 4 * c) if zg.gbuf is not None else None,
                    zg.gbuf.stride(0) if zg.gbuf is not None else 0,
                    (zg.gbuf16.data_ptr() + 2 * c) if zg.gbuf16 is not None else None,
                    zg.gbuf16.stride(0) if zg.gbuf16 is not None else 0,
                    b.store.grad_ptr(bn0.weight), b.store.grad_ptr(bn0.bias), stream), f"bn bwd {self.label}")
        if self.dz_stage is not None:
            dzg = self.outs[0].group
            L.check(b.lib.mmlrec_copy_cols(dzg.gbuf.data_ptr(), dzg.gbuf.stride(0), None, 0, self.dz_stage.data_ptr(),
                                           self.dz_stage.stride(0), b.B, dzg.total, stream), f"dZ -> bf16 {self.label}")
        for buf in self.zero_before_bwd:
            L.check(b.lib.mmlrec_fill_f32(buf.data_ptr(), buf.numel(), 0.0, stream), f"zero d(input) {self.label}")
        for tbl in self.bwd:
            self._launch(tbl, stream, "linear bwd")
        for z32, z16, ld, n, out in self.colsums:
            L.check(b.lib.mmlrec_colsum(z32, z16, ld, b.B, n, out, stream), f"bias gradient {self.label}")


def mlp_stages(b: Builder, items: Sequence[Tuple[Act, nn.Module]], label: str) -> List[Act]:
    """Run several DNN blocks (``model/utils.py`` DNN: ``linears`` [+ ``bn``]) side by side: layer l of
    every block goes into ONE LinearStage, so blocks that read the same input become one wide GEMM
    and the rest one grouped launch.  Blocks may have different depths."""
    cur = [x for x, _ in items]
    depth = max(len(d.linears) for _, d in items)
    for l in range(depth):
        idx = [i for i, (_, d) in enumerate(items) if len(d.linears) > l]
        specs = [LinearSpec(cur[i], *items[i][1].layer(l)) for i in idx]
        stage = b.add(LinearStage(b, specs, items[idx[0]][1].activation, label=f"{label}.l{l}"))
        for i, o in zip(idx, stage.outs):
            cur[i] = o
    return cur


# ----------------------------------------------------------------------------------------------
# element-wise stages (PEPNet) and derived weights (STAR)
# ----------------------------------------------------------------------------------------------
class ConcatStage(Stage):
    """cat(sources, -1) of DETACHED inputs into a fresh buffer without a gradient
    (pepnet.py:73, :139: ``torch.cat([x.detach(), emb], -1)``)."""
    name = "concat"

    def __init__(self, b: Builder, sources: List[Act], label: str = ""):
        self.b, self.sources, self.label = b, sources, label
        for a in sources:
            a.want(f32=True)
        (self.out,) = b.new_group([sum(a.width for a in sources)], relu=False, name=f"{label}.cat")
        self.out.group.need_grad = False

    def forward(self, stream, training):
        b, o, at = self.b, self.out, 0
        for a in self.sources:
            L.check(b.lib.mmlrec_copy_cols(a.ptr, a.ld, (o.ptr + 4 * at) if o.has_f32 else None, o.ld if o.has_f32 else 0,
                                           (o.ptr16 + 2 * at) if o.has_bf16 else None, o.ld16 if o.has_bf16 else 0,
                                           b.B, a.width, stream), f"concat {self.label}")
            at += a.width


class MulStage(Stage):
    """out_i = a_i * b_i for a list of pairs (pepnet.py:78 ``hidden * gw``, :140 ``feature_gate * dnn_input``);
    backward writes d(a), d(b) with the producers' activation derivatives folded in."""
    name = "mul"

    def __init__(self, b: Builder, pairs: List[Tuple[Act, Act]], label: str = ""):
        self.b, self.pairs, self.label = b, pairs, label
        for a, c in pairs:
            assert a.width == c.width
            a.want(f32=True)
            c.want(f32=True)
        self.outs = b.new_group([a.width for a, _ in pairs], relu=False, name=f"{label}.mul", grad_dtype="f32")

    def forward(self, stream, training):
        b = self.b
        for (a, c), o in zip(self.pairs, self.outs):
            L.check(b.lib.mmlrec_mul_forward(a.ptr, a.ld, c.ptr, c.ld, o.ptr if o.has_f32 else None,
                                             o.ld if o.has_f32 else 0, o.ptr16 if o.has_bf16 else None,
                                             o.ld16 if o.has_bf16 else 0, b.B, a.width, stream), f"mul fwd {self.label}")

    def plan_backward(self):
        self.bwd_args = []
        for (a, c), o in zip(self.pairs, self.outs):
            if not o.grad_written:
                continue
            side = []
            for t in (a, c):
                if not t.group.need_grad:
                    side.append((None, None, 0, 0, 0))
                    continue
                f32 = t.grad_is_f32
                acc = 1 if t.grad_written else 0
                assert f32 or not acc, "a bf16 gradient buffer cannot be accumulated into"
                side.append((t.gptr if f32 else None, None if f32 else t.gptr, t.gld, t.dkind, acc))
                t.grad_written = True
            self.bwd_args.append((o, a, c, side))

    def backward(self, stream):
        b = self.b
        for o, a, c, (sa, sc) in self.bwd_args:
            L.check(b.lib.mmlrec_mul_backward(o.gptr, o.gld, a.ptr, a.ld, c.ptr, c.ld, sa[0], sa[1], sa[2], sa[3], sa[4],
                                              sc[0], sc[1], sc[2], sc[3], sc[4], b.B, a.width, stream),
                    f"mul bwd {self.label}")


class PairAttentionStage(Stage):
    """AITM's information transfer between two consecutive tasks (aitm.py:82-91): the tokens p (transferred from the
    previous task) and q (this task's own feature) each carry value / key / query projections; ``out = sum_j
    softmax_j(<K_j, Q_j> / sqrt(H)) V_j``.  ``vkq`` = the six [B, H] projections V_p K_p Q_p V_q K_q Q_q as ADJACENT
    columns of one buffer (the outputs of one LinearStage, in that order)."""
    name = "aitm_attention"

    def __init__(self, b: Builder, vkq: List[Act], label: str = ""):
        self.b, self.vkq, self.label = b, vkq, label
        H = vkq[0].width
        assert len(vkq) == 6 and all(a.group is vkq[0].group and a.width == H and a.col == vkq[0].col + i * H
                                     for i, a in enumerate(vkq)), "the six projections must be adjacent columns"
        self.H = H
        for a in vkq:
            a.want(f32=True)
        (self.out,) = b.new_group([H], relu=False, name=f"{label}.att", grad_dtype="f32")

    def finalize(self):
        self.attn = self.b.zeros(self.b.B, 2)

    def forward(self, stream, training):
        b, o, x = self.b, self.out, self.vkq[0]
        L.check(b.lib.mmlrec_aitm_attention_forward(x.ptr, x.ld, b.B, self.H, o.ptr if o.has_f32 else None,
                                                    o.ld if o.has_f32 else 0, o.ptr16 if o.has_bf16 else None,
                                                    o.ld16 if o.has_bf16 else 0, self.attn.data_ptr(), stream),
                f"attention fwd {self.label}")

    def plan_backward(self):
        self.live = self.out.grad_written and self.vkq[0].group.need_grad
        if self.live:
            for a in self.vkq:
                a.grad_written = True

    def backward(self, stream):
        if not self.live:
            return
        b, o, x = self.b, self.out, self.vkq[0]
        f32 = x.grad_is_f32
        L.check(b.lib.mmlrec_aitm_attention_backward(o.gptr, o.gld, x.ptr, x.ld, self.attn.data_ptr(), b.B, self.H,
                                                     x.gptr if f32 else None, None if f32 else x.gptr, x.gld, stream),
                f"attention bwd {self.label}")


class ApgMixStage(Stage):
    """APG's per-sample product (apg.py:96-99): ``kk[b] = nk[b] @ wkk[b].view(k, k) + bkk[b]`` with the sample's matrix
    and bias generated from its scene embedding.  ``nk`` [B, k], ``wkk`` [B, k*k], ``bkk`` [B, k] are outputs of the
    preceding LinearStage; this stage is their only consumer, so backward assigns all three gradients."""
    name = "apg_mix"

    def __init__(self, b: Builder, nk: Act, wkk: Act, bkk: Act, label: str = ""):
        self.b, self.nk, self.wkk, self.bkk, self.label = b, nk, wkk, bkk, label
        self.k = nk.width
        assert wkk.width == self.k * self.k and bkk.width == self.k
        for a in (nk, wkk, bkk):
            assert a.dkind == 0, "the inputs of the APG product carry no activation"
            a.want(f32=True)
        (self.out,) = b.new_group([self.k], relu=False, name=f"{label}.kk", grad_dtype="f32")

    def forward(self, stream, training):
        b, o = self.b, self.out
        L.check(b.lib.mmlrec_apg_mix_forward(self.nk.ptr, self.nk.ld, self.wkk.ptr, self.wkk.ld, self.bkk.ptr, self.bkk.ld,
                                             b.B, self.k, o.ptr if o.has_f32 else None, o.ld if o.has_f32 else 0,
                                             o.ptr16 if o.has_bf16 else None, o.ld16 if o.has_bf16 else 0, stream),
                f"apg mix fwd {self.label}")

    def plan_backward(self):
        self.live = self.out.grad_written
        if self.live:
            for a in (self.nk, self.wkk, self.bkk):
                assert a.group.need_grad and not a.grad_written, "the APG product is the only consumer of its inputs"
                a.grad_written = True

    def backward(self, stream):
        if not self.live:
            return
        b, o = self.b, self.out

        def dst(a):
            return (a.gptr if a.grad_is_f32 else None, None if a.grad_is_f32 else a.gptr, a.gld)

        L.check(b.lib.mmlrec_apg_mix_backward(o.gptr, o.gld, self.nk.ptr, self.nk.ld, self.wkk.ptr, self.wkk.ld, b.B, self.k,
                                              *dst(self.nk), *dst(self.wkk), *dst(self.bkk), stream),
                f"apg mix bwd {self.label}")


class DerivedLinear:
    """Quacks like nn.Linear for LinearSpec / HeadSpec: weight [N,K] (+ bias [N]) living in the aux region."""

    def __init__(self, weight, bias):
        self.weight, self.bias = weight, bias


class StarWeightStage(Stage):
    """SharedSpecificLinear (model/utils.py:163-223): effective weights W_eff[t] = W_spec[t] * W_shared and
    biases b_spec[t] + b_shared of ALL domains, materialised once per step in nn.Linear layout
    ([T*out, in]) so the ordinary grouped-GEMM stages run the layers; backward folds d(W_eff) back into
    d(W_shared), d(b_shared) and the one registered specific weight (index T-1).  Placed right
    after the gather so that its backward runs after every wgrad."""
    name = "star_weights"

    def __init__(self, b: Builder, layers: List[nn.Module], label: str = ""):
        self.b, self.layers, self.label = b, layers, label
        self.live_flags = {}
        self.derived: List[Optional[DerivedLinear]] = []
        for m in layers:
            b.note_params([m.shared_weight, m.shared_bias, m.specific_weight, m.specific_bias])
            K, N = m.in_features, m.out_features
            T = len(m.spec_weights())
            w = b.aux_matrix(T * N, K)
            bias = b.aux_vector(T * N)
            self.derived.append(None if b.dry else DerivedLinear(w, bias))

    def linear(self, layer: int, t: int) -> Optional[DerivedLinear]:
        """nn.Linear-like view of domain t of layer `layer` (rows [t*N, (t+1)*N) of the derived matrix)."""
        if self.b.dry:
            m = self.layers[layer]
            return _DryLinear(m.out_features, m.in_features)
        d, m = self.derived[layer], self.layers[layer]
        N = m.out_features
        w = d.weight[t * N:(t + 1) * N]
        w._mm_off, w._mm_ld, w._mm_span, w._mm_kind = d.weight._mm_off + t * N * d.weight._mm_ld, d.weight._mm_ld, \
            N * d.weight._mm_ld, "dense"
        bb = d.bias[t * N:(t + 1) * N]
        bb._mm_off, bb._mm_ld, bb._mm_span, bb._mm_kind = d.bias._mm_off + t * N, 0, N, "dense"
        return DerivedLinear(w, bb)

    def finalize(self):
        b = self.b
        self.ptrs = []
        for m in self.layers:
            sw = b.ints([w.data_ptr() for w in m.spec_weights()], dtype=torch.int64)
            sb = b.ints([w.data_ptr() for w in m.spec_biases()], dtype=torch.int64)
            self.ptrs.append((sw, sb))
        self.live = [b.ints(self.live_flags.get(i, [1] * len(m.spec_weights()))) for i, m in enumerate(self.layers)]

    def set_live(self, layer: int, flags: List[int]):
        """Domains of `layer` whose effective weight receives a gradient (a per-domain head only uses its own)."""
        self.live_flags[layer] = list(flags)

    def forward(self, stream, training):
        b, st = self.b, self.b.store
        for m, d, (sw, sb) in zip(self.layers, self.derived, self.ptrs):
            T = len(m.spec_weights())
            L.check(b.lib.mmlrec_star_weights(sw.data_ptr(), sb.data_ptr(), m.shared_weight.data_ptr(),
                                              m.shared_bias.data_ptr(), T, m.in_features, m.out_features,
                                              d.weight.data_ptr(), d.weight._mm_ld,
                                              st.bf16_ptr(d.weight) if st.dense_bf16 is not None else None,
                                              d.bias.data_ptr(), stream), f"star weights {self.label}")

    def backward(self, stream):
        b, st = self.b, self.b.store
        for i, (m, d, (sw, sb)) in enumerate(zip(self.layers, self.derived, self.ptrs)):
            T = len(m.spec_weights())
            live = self.live[i]
            L.check(b.lib.mmlrec_star_fold(st.grad_ptr(d.weight), d.weight._mm_ld, st.grad_ptr(d.bias), sw.data_ptr(),
                                           m.shared_weight.data_ptr(), live.data_ptr(), T, m.in_features, m.out_features,
                                           st.grad_ptr(m.shared_weight), st.grad_ptr(m.shared_bias),
                                           st.grad_ptr(m.specific_weight), st.grad_ptr(m.specific_bias), stream),
                    f"star fold {self.label}")


class SnrGateStage(Stage):
    """SNR-trans gate (snr_trans.py:9-50) as a derived weight: ``out_i = sum_j z_ij (x_j @ M_ij)`` for all i is ONE Linear
    over the concatenated inputs with ``W_eff[i*U+v, j*U+u] = z_ij M_ij[u, v]``, rebuilt every step from the trained
    hard-concrete parameters (``u``, ``alpha``) and the constant trans matrices; an ordinary LinearStage applies it, and
    this stage's backward folds d(W_eff) into d(u), d(alpha)."""
    name = "snr_gate"

    def __init__(self, b: Builder, gate: nn.Module, label: str = ""):
        self.b, self.gate, self.label = b, gate, label
        # SNR-trans trains u [n_out, n_in]; MSSM's u [n_out, n_in, U] is an unregistered constant (mssm.py:27-29)
        self.u_trained = isinstance(gate.u, nn.Parameter)
        b.note_params([gate.alpha, gate.u] if self.u_trained else [gate.alpha])
        self.n_out, self.n_in, self.U = gate.output_dim, gate.input_dim, gate.units
        self.zdim = self.U if gate.u.dim() == 3 else 1
        w = b.aux_matrix(self.n_out * self.U, self.n_in * self.U)
        self.derived = _DryLinear(self.n_out * self.U, self.n_in * self.U, bias=False) if b.dry else DerivedLinear(w, None)

    def finalize(self):
        self.dz = self.b.zeros(self.n_out * self.n_in * self.zdim)

    def forward(self, stream, training):
        b, st, g, w = self.b, self.b.store, self.gate, self.derived.weight
        L.check(b.lib.mmlrec_snr_gate_weights(g.u.data_ptr(), g.alpha.data_ptr(), g.trans_matrix.data_ptr(), self.n_out,
                                              self.n_in, self.U, self.zdim, w.data_ptr(), w._mm_ld,
                                              st.bf16_ptr(w) if st.dense_bf16 is not None else None, stream),
                f"snr gate weights {self.label}")

    def backward(self, stream):
        b, st, g, w = self.b, self.b.store, self.gate, self.derived.weight
        L.check(b.lib.mmlrec_snr_gate_fold(st.grad_ptr(w), w._mm_ld, g.trans_matrix.data_ptr(), g.u.data_ptr(),
                                           g.alpha.data_ptr(), self.n_out, self.n_in, self.U, self.zdim,
                                           self.dz.data_ptr(), st.grad_ptr(g.u) if self.u_trained else None,
                                           st.grad_ptr(g.alpha), stream), f"snr gate fold {self.label}")


class _DryLinear:
    """Shape-only stand-in used while recording the parameter order."""

    def __init__(self, n, k, bias=True):
        self.weight = _DryTensor((n, k))
        self.bias = _DryTensor((n,)) if bias else None


class _DryTensor:
    def __init__(self, shape):
        self.shape = shape


# ----------------------------------------------------------------------------------------------
# gate head + softmax + mixture
# ----------------------------------------------------------------------------------------------
class GateSpec:
    """``detach[e]``: the gate mixes expert e as a constant (``x.detach()``, hmoe.py:130): value used, no gradient into it."""

    def __init__(self, gate_in: Act, head: nn.Module, experts: List[Act], detach: Optional[List[bool]] = None):
        self.gate_in, self.head, self.experts = gate_in, head, experts
        self.detach = list(detach) if detach is not None else [False] * len(experts)
        assert len(self.detach) == len(experts)


class GateMixStage(Stage):
    """softmax(gate_in Wg^T) @ stack(experts) for every gate of a level in one launch
    (model/mmoe.py:80-88, model/ple.py:127-152)."""
    name = "gate_mix"

    def __init__(self, b: Builder, gates: List[GateSpec], label: str = ""):
        self.b, self.gates, self.label = b, gates, label
        b.note_params([g.head.weight for g in gates])
        H = gates[0].experts[0].width
        assert all(e.width == H for g in gates for e in g.experts)
        self.H = H
        self.any_detach = any(any(g.detach) for g in gates)
        self.outs = b.new_group([H] * len(gates), relu=False, name=f"{label}.mix", grad_dtype="f32")
        self.outs[0].want(f32=True)
        for g in gates:
            g.gate_in.want(f32=True)
            for e in g.experts:
                e.want(f32=True)
        self.any_live = False

    def finalize(self):
        self.probs = [self.b.zeros(self.b.B, len(g.experts)) for g in self.gates]
        # distinct expert activations of the level, in first-use order
        self.uniq: List[Act] = []
        for g in self.gates:
            for a in g.experts:
                if not any(a.same_as(u) for u in self.uniq):
                    self.uniq.append(a)
        self.total_wg = sum(len(g.experts) * g.gate_in.width for g in self.gates)
        self.fused = (len(self.gates) <= L.LEVEL_MAX_GATES and len(self.uniq) <= L.LEVEL_MAX_EXPERTS
                      and self.H % 4 == 0 and self.total_wg <= L.LEVEL_MAX_WG
                      and all(a.col % 4 == 0 for a in self.uniq) and all(o.col % 4 == 0 for o in self.outs))
        if self.fused:
            rec = self._level_record(False)
            self.level_table = self.b.table([rec])
            self.total_ne = sum(len(g.experts) for g in self.gates)
            self.total_hg = sum(g.gate_in.width for g in self.gates)
            # the tiled kernels move whole rows with 16-byte cp.async / vector stores
            self.tiled_fwd = (all(rec.Hg[i] % 4 == 0 and (rec.gate_in[i] or 0) % 16 == 0 and rec.ld_gate_in[i] % 4 == 0
                                  and (rec.Wg[i] or 0) % 16 == 0 and rec.ld_Wg[i] % 4 == 0
                                  for i in range(len(self.gates)))
                              and not self.b.dry and self.total_ne <= 256
                              and self.b.lib.mmlrec_gate_level_forward_tiled_smem(
                                  len(self.uniq), self.H, self.total_wg, self.total_ne, self.total_hg) <= 110 * 1024)
        else:
            self.gate_table = self.b.table([self._gate_record(i, False) for i in range(len(self.gates))])

    def _level_record(self, backward: bool) -> L.GateLevel:
        """Record of the row-fused kernels.  In backward mode also claims the gradient buffers
        (assign / accumulate bookkeeping identical to the per-gate path)."""
        st = self.b.store
        r = L.GateLevel()
        r.n_gates, r.n_experts, r.H = len(self.gates), len(self.uniq), self.H
        r.expert_relu = 1 if self.uniq[0].relu else 0
        assert all(a.relu == self.uniq[0].relu and a.ld == self.uniq[0].ld for a in self.uniq)
        r.ld_expert = self.uniq[0].ld
        for u, a in enumerate(self.uniq):
            r.expert[u] = a.ptr
            for g in range(L.LEVEL_MAX_GATES):
                r.slot[u][g] = -1
        live = [backward and self.outs[i].grad_written for i in range(len(self.gates))]
        for i, g in enumerate(self.gates):
            o, gi = self.outs[i], g.gate_in
            r.gate_in[i], r.ld_gate_in[i], r.Hg[i], r.n_e[i] = gi.ptr, gi.ld, gi.width, len(g.experts)
            r.Wg[i], r.ld_Wg[i] = g.head.weight.data_ptr(), g.head.weight._mm_ld
            r.probs[i] = self.probs[i].data_ptr()
            r.mix[i], r.ld_mix[i] = o.ptr, o.ld
            if o.has_bf16:
                r.mix_bf16[i], r.ld_mix_bf16[i] = o.ptr16, o.ld16
            for e, a in enumerate(g.experts):
                u = next(k for k, x in enumerate(self.uniq) if x.same_as(a))
                r.slot[u][i] = e
                if g.detach[e]:
                    r.detach_mask[u] |= 1 << i
            if live[i]:
                r.d_mix[i], r.ld_d_mix[i] = o.gptr, o.gld
                if gi.group.need_grad:
                    if gi.grad_is_f32:
                        r.d_gate_in[i], r.ld_d_gate_in[i] = gi.gptr, gi.gld
                        r.accumulate_d_gate_in[i] = 1 if gi.grad_written else 0
                    else:
                        assert not gi.grad_written, "a bf16 gradient buffer cannot be accumulated into"
                        r.d_gate_in_bf16[i], r.ld_d_gate_in_bf16[i] = gi.gptr, gi.gld
                    r.relu_mask_gate_in[i] = 1 if gi.relu else 0
                    gi.grad_written = True
                r.dWg[i] = st.grad_ptr(g.head.weight)
        if backward:
            for u, a in enumerate(self.uniq):
                if not any(live[i] and any(a.same_as(x) and not g.detach[e] for e, x in enumerate(g.experts))
                           for i, g in enumerate(self.gates)):
                    continue   # no live gate sends a gradient into this expert
                assert not a.grad_written, "an expert output consumed elsewhere must be accumulated"
                if a.grad_is_f32:
                    r.d_expert[u], r.ld_d_expert = a.gptr, a.gld
                else:
                    r.d_expert_bf16[u], r.ld_d_expert_bf16 = a.gptr, a.gld
                a.grad_written = True
        return r

    def _gate_record(self, i: int, backward: bool) -> L.Gate:
        g, st, o = self.gates[i], self.b.store, self.outs[i]
        r = L.Gate()
        r.gate_in, r.ld_gate_in, r.Hg, r.n_e = g.gate_in.ptr, g.gate_in.ld, g.gate_in.width, len(g.experts)
        assert r.n_e <= L.MAX_GATE_EXPERTS
        r.Wg, r.ld_Wg = g.head.weight.data_ptr(), g.head.weight._mm_ld
        for e, a in enumerate(g.experts):
            r.expert[e] = a.ptr
        assert all(a.ld == g.experts[0].ld for a in g.experts)
        r.ld_expert, r.H = g.experts[0].ld, self.H
        r.probs = self.probs[i].data_ptr()
        r.mix, r.ld_mix = o.ptr, o.ld
        if o.has_bf16:
            r.mix_bf16, r.ld_mix_bf16 = o.ptr16, o.ld16
        if backward and o.grad_written:
            r.d_mix, r.ld_d_mix = o.gptr, o.gld
            gi = g.gate_in
            if gi.group.need_grad:
                if gi.grad_is_f32:
                    r.d_gate_in, r.ld_d_gate_in = gi.gptr, gi.gld
                    r.accumulate_d_gate_in = 1 if gi.grad_written else 0
                else:
                    assert not gi.grad_written, "a bf16 gradient buffer cannot be accumulated into"
                    r.d_gate_in_bf16, r.ld_d_gate_in_bf16 = gi.gptr, gi.gld
                r.relu_mask_gate_in = 1 if gi.relu else 0
                gi.grad_written = True
            r.dWg = st.grad_ptr(g.head.weight)
        return r

    def forward(self, stream, training):
        if self.fused and self.tiled_fwd:
            L.check(self.b.lib.mmlrec_gate_level_forward_tiled(self.level_table.data_ptr(), self.b.B, len(self.uniq), self.H,
                                                               self.total_wg, self.total_ne, self.total_hg, stream),
                    f"gate_level fwd (tiled) {self.label}")
        elif self.fused:
            L.check(self.b.lib.mmlrec_gate_level_forward(self.level_table.data_ptr(), self.b.B, stream),
                    f"gate_level fwd {self.label}")
        else:
            L.check(self.b.lib.mmlrec_gate_mix_forward(self.gate_table.data_ptr(), len(self.gates), self.b.B, stream),
                    f"gate_mix fwd {self.label}")

    def plan_backward(self):
        b = self.b
        live = [i for i in range(len(self.gates)) if self.outs[i].grad_written]
        self.any_live = bool(live)
        if not live:
            return
        if self.fused:
            rec = self._level_record(True)
            self.level_table = b.table([rec])
            G, E = len(self.gates), len(self.uniq)
            self.total_ne = sum(len(g.experts) for g in self.gates)
            self.total_hg = sum(g.gate_in.width for g in self.gates)

            def ok16(ptr, ld):
                return (ptr or 0) % 16 == 0 and ld % 4 == 0
            # the tiled kernel moves whole rows with 16-byte cp.async / vector stores
            self.tiled = (all(rec.Hg[i] % 4 == 0 and ok16(rec.gate_in[i], rec.ld_gate_in[i])
                              and ok16(rec.Wg[i], rec.ld_Wg[i]) and ok16(rec.d_gate_in[i], rec.ld_d_gate_in[i])
                              and (rec.d_gate_in_bf16[i] or 0) % 8 == 0 and rec.ld_d_gate_in_bf16[i] % 4 == 0
                              for i in range(G))
                          and b.lib.mmlrec_gate_level_backward_tiled_smem(G, E, self.H, self.total_wg, self.total_ne,
                                                                          self.total_hg) <= 110 * 1024)
            if self.any_detach and not self.tiled:
                raise NotImplementedError("detached experts (HMoE task weights) need the tiled gate backward")
            if self.tiled:
                n = b.lib.mmlrec_gate_level_backward_tiled_scratch(self.total_wg, b.B)
            else:
                n = b.lib.mmlrec_gate_level_backward_scratch(self.total_wg, b.B)
            self.scratch = b.zeros(int(n))
            self.counters = b.zeros(1, dtype=torch.int32)
            return
        if self.any_detach:
            raise NotImplementedError("detached experts (HMoE task weights) need the level-fused gate kernels")
        # gates that share an input (no gate DNN: every head reads the level input) must not race on
        # d(gate_in): the kernel then walks the gates in order inside each CTA
        seen, self.serialize = [], 0
        for i in live:
            gi = self.gates[i].gate_in
            if any(gi.same_as(o) for o in seen):
                self.serialize = 1
            seen.append(gi)
        self.gate_table = b.table([self._gate_record(i, True) for i in range(len(self.gates))])
        # experts: every distinct expert activation gets d = sum over its user gates of p * d_mix
        recs, uniq = [], []
        for i in live:
            for a in self.gates[i].experts:
                if not any(a.same_as(u) for u in uniq):
                    uniq.append(a)
        for a in uniq:
            r = L.ExpertGrad()
            assert not a.grad_written, "an expert output consumed elsewhere must be accumulated"
            r.expert, r.ld_expert, r.H = a.ptr, a.ld, self.H
            if a.grad_is_f32:
                r.d_expert, r.ld_d_expert = a.gptr, a.gld
            else:
                r.d_expert_bf16, r.ld_d_expert_bf16 = a.gptr, a.gld
            r.relu_mask = 1 if a.relu else 0
            n = 0
            for i in live:
                for e, ea in enumerate(self.gates[i].experts):
                    if ea.same_as(a):
                        r.user_probs[n] = self.probs[i].data_ptr()
                        r.user_prob_ld[n], r.user_prob_col[n] = self.probs[i].stride(0), e
                        r.user_d_mix[n], r.user_d_mix_ld[n] = self.outs[i].gptr, self.outs[i].gld
                        n += 1
            assert n <= L.MAX_TASKS + 1
            r.n_users = n
            a.grad_written = True
            recs.append(r)
        self.expert_table, self.n_expert_recs = b.table(recs), len(recs)
        self.max_ne = max(len(g.experts) for g in self.gates)
        self.max_hg = max(g.gate_in.width for g in self.gates)
        n = b.lib.mmlrec_gate_mix_backward_scratch(len(self.gates), self.max_ne, self.max_hg, b.B)
        self.scratch = b.zeros(int(n))
        self.counters = b.zeros(len(self.gates), dtype=torch.int32)

    def backward(self, stream):
        if not self.any_live:
            return
        b = self.b
        if self.fused and self.tiled:
            L.check(b.lib.mmlrec_gate_level_backward_tiled(self.level_table.data_ptr(), b.B, len(self.gates),
                                                           len(self.uniq), self.H, self.total_wg, self.total_ne,
                                                           self.total_hg, self.scratch.data_ptr(), stream),
                    f"gate_level bwd (tiled) {self.label}")
            return
        if self.fused:
            L.check(b.lib.mmlrec_gate_level_backward(self.level_table.data_ptr(), b.B, self.total_wg, self.total_ne,
                                                     self.total_hg, self.scratch.data_ptr(), self.counters.data_ptr(),
                                                     stream),
                    f"gate_level bwd {self.label}")
            return
        L.check(b.lib.mmlrec_gate_mix_backward(self.gate_table.data_ptr(), len(self.gates), self.expert_table.data_ptr(),
                                               self.n_expert_recs, b.B, self.max_ne, self.max_hg, self.serialize,
                                               self.scratch.data_ptr(), self.counters.data_ptr(), stream),
                f"gate_mix bwd {self.label}")


# ----------------------------------------------------------------------------------------------
# heads + loss
# ----------------------------------------------------------------------------------------------
class HeadSpec:
    def __init__(self, h: Act, final: nn.Module, bias: Optional[nn.Parameter], task: str, bias2=None):
        self.h, self.final, self.bias, self.task, self.bias2 = h, final, bias, task, bias2


class HeadStage(Stage):
    """Bias-free 1-unit head + PredictionLayer + sum-BCE, forward and backward in one kernel
    (model/mmoe.py:97-100, model/utils.py:242-248, model/basemodel.py:294-296)."""
    name = "heads"

    def __init__(self, b: Builder, heads: List[HeadSpec], esmm: bool = False, cumulative_bias: bool = False,
                 escm: Optional[Tuple[float, float]] = None):
        # cumulative_bias: task t's logit carries the biases of tasks 0..t (mlp.py:47 hands ONE logit tensor to every
        # PredictionLayer, whose ``output += self.bias`` works in place, model/utils.py:243-245)
        self.b, self.heads, self.esmm = b, heads, esmm
        self.flags = (1 if esmm else 0) | (2 if cumulative_bias else 0)
        # ESCM (escm.py:74-96 + the loss of basemodel.py:284-292): the kernel gives the two probabilities (shared bias),
        # the three-column prediction, the IPW-weighted loss and its gradient with respect to the probabilities are a
        # dozen element-wise torch ops on [B]-sized tensors, then the kernel's external-gradient backward takes over
        self.escm = escm   # (counterfactual_w, global_w) or None
        if escm is not None:
            assert len(heads) == 2 and not esmm
            self.flags = 4
        b.note_params([h.final.weight for h in heads])
        b.note_params([h.bias for h in heads])
        b.note_params([h.bias2 for h in heads])
        self.T = len(heads)
        for h in heads:
            h.h.want(f32=True)
        # heads sharing one final layer (MLP): only the one-launch head kernel sums their weight gradients
        shared = len({id(h.final.weight) for h in heads}) < len(heads)
        wmax = max(h.h.width for h in heads)
        if shared and not ((self.T <= 8 and wmax <= 128) or (self.T <= 4 and wmax <= 256 and cumulative_bias)):
            raise NotImplementedError("a final layer shared by several heads needs T <= 8 tasks of width <= 128 "
                                      "(or T <= 4 of width <= 256 with cumulative biases: MLP)")

    def finalize(self):
        b, st = self.b, self.b.store
        self.y = b.zeros(b.B, self.T)
        # scenario mask (b200_config["domain_mask"]): [B, num_domains] of 0 / 1, head t reads column t (msl) or
        # t % num_domains (mtmsl) -- mmoe.py:101-106
        self.mask = None
        D = getattr(self.b, "mask_domains", 0)
        if D:
            assert not self.esmm and all(h.task == "binary" for h in self.heads), "scenario masks: binary heads only"
            self.mask = b.zeros(b.B, D)
            self.mask.fill_(1.0)
        self.d_pred = b.zeros(b.B, self.T)
        self.pred = b.zeros(b.B, self.T)
        if self.escm is not None:
            self.pred2, self.pred = self.pred, b.zeros(b.B, 3)   # kernel output [p_ctr, p_cvr]; model output adds ctcvr
        self.loss = b.zeros(self.T + 1)
        recs = []
        for h in self.heads:
            r = L.Head()
            r.h, r.ld_h, r.H = h.h.ptr, h.h.ld, h.h.width
            r.kind = L.HEAD_SIGMOID_BCE if h.task == "binary" else L.HEAD_IDENTITY_MSE
            r.w = h.final.weight.data_ptr()
            r.bias = h.bias.data_ptr() if h.bias is not None else None
            if h.h.group.need_grad:
                # an activation read by several heads of this stage: the head kernel sums their contributions
                again = any(o.h.same_as(h.h) for o in self.heads[:self.heads.index(h)])
                assert again or not h.h.grad_written
                assert not again or (self.T <= 8 and h.h.width <= 128) or \
                    (self.T <= 4 and h.h.width <= 256 and (self.flags & 6)), "heads sharing an input need the one-launch kernel"
                if h.h.grad_is_f32:
                    r.d_h, r.ld_d_h = h.h.gptr, h.h.gld
                else:
                    r.d_h_bf16, r.ld_d_h_bf16 = h.h.gptr, h.h.gld
                r.relu_mask = 1 if h.h.relu else 0
                h.h.grad_written = True
            r.dw = st.grad_ptr(h.final.weight)
            r.dbias = st.grad_ptr(h.bias) if h.bias is not None else None
            if h.bias2 is not None:
                r.bias2, r.dbias2 = h.bias2.data_ptr(), st.grad_ptr(h.bias2)
            r.mask_col = (len(recs) % D) if D else 0
            recs.append(r)
        self.table = b.table(recs)
        max_h = max(h.h.width for h in self.heads)
        self.scratch = b.zeros(int(b.lib.mmlrec_heads_scratch(self.T, max_h, b.B)))
        self.counter = b.zeros(1, dtype=torch.int32)

    def forward(self, stream, training):
        b = self.b
        if self.escm is not None:
            return self._escm_forward(stream, training)
        if self.mask is not None:
            L.check(b.lib.mmlrec_heads_forward_backward_masked(
                self.table.data_ptr(), self.T, b.B, self.y.data_ptr() if training else None, self.T, self.mask.data_ptr(),
                self.mask.stride(0), self.pred.data_ptr(), self.T, self.loss.data_ptr(), self.flags, 1 if training else 0,
                self.scratch.data_ptr(), self.scratch.numel(), self.counter.data_ptr(), stream), "heads (masked)")
            return
        L.check(b.lib.mmlrec_heads_forward_backward(
            self.table.data_ptr(), self.T, b.B, self.y.data_ptr() if training else None, self.T, self.pred.data_ptr(),
            self.T, self.loss.data_ptr(), self.flags, 1 if training else 0, self.scratch.data_ptr(),
            self.scratch.numel(), self.counter.data_ptr(), stream), "heads")

    def _escm_forward(self, stream, training):
        b, B = self.b, self.b.B
        L.check(b.lib.mmlrec_heads_forward_backward(
            self.table.data_ptr(), self.T, B, None, self.T, self.pred2.data_ptr(), self.T, self.loss.data_ptr(), self.flags,
            0, self.scratch.data_ptr(), self.scratch.numel(), self.counter.data_ptr(), stream), "heads (escm forward)")
        p0, p1 = self.pred2[:, 0], self.pred2[:, 1]
        p2 = p0 * p1                                             # escm.py:88  ctcvr = ctr * cvr
        torch.stack([p0, p1, p2], dim=1, out=self.pred)
        if not training:
            return
        w_cf, w_g = self.escm
        y0, y1 = self.y[:, 0], self.y[:, 1]

        def bce(p, y):        # F.binary_cross_entropy per element (log terms clamped at -100) and its derivative in p
            val = (y - 1.0) * torch.log1p(-p).clamp_min(-100.0) - y * torch.log(p).clamp_min(-100.0)
            return val, (p - y) / ((1.0 - p) * p).clamp_min(1e-12)

        (l0, g_l0), (l1, g_l1), (l2, g_l2) = bce(p0, y0), bce(p1, y1), bce(p2, y1)
        L0, L1, L2 = l0.sum(), l1.sum(), l2.sum()
        # counterfact_ipw (escm.py:98-111; `ips.stop_gradient = True` is a no-op in torch: the gradient flows through ips)
        n = y0.sum()
        ps_raw = p0 * n
        ps = ps_raw.clamp_min(1e-6)
        inv = 1.0 / ps
        ips = inv.clamp(-15.0, 15.0) * float(B)
        s = (ips * y0).mean()
        total = L0 + w_cf * (L1 * s) + w_g * L2                  # basemodel.py:284-292
        d_ips = torch.where((inv >= -15.0) & (inv <= 15.0) & (ps_raw > 1e-6), -n / (ps * ps), torch.zeros_like(ps)) * float(B)
        g0 = g_l0 + w_cf * L1 * (y0 * d_ips) / float(B) + w_g * g_l2 * p1
        g1 = w_cf * s * g_l1 + w_g * g_l2 * p0
        self.backward_external(stream, torch.stack([g0, g1], dim=1))
        self.loss.copy_(torch.stack([L0, total - L0, total]))

    def backward_external(self, stream, d_pred: torch.Tensor):
        """Backward for an upstream gradient dL/d(pred) [B, T] handed in by autograd (differentiable forward())."""
        b = self.b
        self.d_pred.copy_(d_pred)
        pred = self.pred2 if self.escm is not None else self.pred   # (the kernel's own [B, T] output buffer)
        L.check(b.lib.mmlrec_heads_backward_external(
            self.table.data_ptr(), self.T, b.B, self.d_pred.data_ptr(), self.T, pred.data_ptr(), self.T,
            self.loss.data_ptr(), self.flags, self.scratch.data_ptr(), self.scratch.numel(),
            self.counter.data_ptr(), stream), "heads (external gradient)")


# ----------------------------------------------------------------------------------------------
# the program
# ----------------------------------------------------------------------------------------------
class StepPlan:
    """The compiled step of one model at one batch size."""

    def __init__(self, model, B: int):
        self.model, self.B = model, B
        self.b = Builder(B, model.device_obj, model.store, dry=False, precision=model.precision)
        self.b.dp = getattr(model, "dp", None)
        self.b.mask_domains = model.num_domains if getattr(model, "use_domain_mask", False) else 0
        model.build_graph(self.b)
        self.b.materialize()
        self.stages = self.b.stages
        self.gather: GatherStage = next(s for s in self.stages if isinstance(s, GatherStage))
        self.heads: HeadStage = next(s for s in self.stages if isinstance(s, HeadStage))
        for s in reversed(self.stages):
            s.plan_backward()
        # gradient slices this program writes (split-K wgrad, see LinearStage.plan_backward)
        self.grad_slices = max([getattr(s, "split_k", 1) for s in self.stages] + [1])
        self.fold_seg = self.b.ints([0, 0, model.store.slice_stride], dtype=torch.int64)
        self._plan_buckets()
        self.reg_loss = self.b.zeros(1)     # value of the L2 term of the last step (0 without regularisation)
        self.l2_scratch = self.b.zeros(int(self.b.lib.mmlrec_l2_scratch())) if model.store.l2_coef is not None else None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.side = torch.cuda.Stream(device=model.device_obj)
        self.ev_fork, self.ev_join = torch.cuda.Event(), torch.cuda.Event()
        self.ev_bwd, self.ev_join2 = torch.cuda.Event(), torch.cuda.Event()

    # ---- data parallel: gradient buckets
    def _plan_buckets(self) -> None:
        """Cut the flat dense-gradient store into buckets along stage boundaries.  The store is laid out in stage order
        and backward runs the stages in reverse, so once stage k's backward has been issued the range [first parameter
        of stage k, end) is final: its slices are folded and all-reduced on a communication stream while the earlier
        stages' backward is still running; only the last bucket (the first stages' parameters) is reduced on the main
        stream, after the last backward kernel.  ``b200_config["grad_bucket_floats"]`` = minimum bucket size (0: one
        bucket = the old behaviour)."""
        m, st = self.model, self.model.store
        self.buckets, self.bucket_after, self.first_bucket_hi = [], {}, st.slice_stride
        import os
        min_floats = int(os.environ.get("MMLREC_GRAD_BUCKET_FLOATS", m.b200_config.get("grad_bucket_floats", 1 << 18)))
        self.fold_all = self.b.ints([0, 0, st.slice_stride], dtype=torch.int64)
        sh = getattr(m, "shard", None)
        if sh is not None and sh.grad_in is not None:
            return   # the peer-memory all-reduce takes the whole store after the last backward kernel
        if getattr(m, "dp", None) is None or min_floats <= 0 or st.aux_floats > 64:
            return   # (derived weights -- STAR -- are folded by a stage that runs last: one bucket)
        firsts = []
        for s in self.stages:
            offs = [p._mm_off for p in getattr(s, "_params", []) if getattr(p, "_mm_kind", "") == "dense"]
            if offs:
                firsts.append((s, min(offs)))
        if len(firsts) < 2 or any(a[1] >= b[1] for a, b in zip(firsts, firsts[1:])) or firsts[0][1] != 0:
            return
        hi = st.slice_stride
        for s, lo in reversed(firsts[1:]):
            if hi - lo >= min_floats and lo >= min_floats:
                seg = self.b.ints([lo, lo, hi - lo], dtype=torch.int64)
                self.bucket_after[id(s)] = len(self.buckets)
                self.buckets.append((lo, hi, seg))
                hi = lo
        self.first_bucket_hi = hi
        if self.buckets:
            self.comm = torch.cuda.Stream(device=m.device_obj)
            self.ev_bucket = [torch.cuda.Event() for _ in self.buckets]
            self.ev_comm = torch.cuda.Event()
            self.fold_seg = self.b.ints([0, 0, hi], dtype=torch.int64)

    def _reduce_bucket(self, i: int, main) -> None:
        """Fold + all-reduce bucket i on the communication stream, behind everything issued on `main` so far."""
        lo, hi, seg = self.buckets[i]
        st, lib = self.model.store, self.b.lib
        self.ev_bucket[i].record(main)
        self.comm.wait_event(self.ev_bucket[i])
        with torch.cuda.stream(self.comm):
            if self.grad_slices > 1:
                L.check(lib.mmlrec_sum_slices(seg.data_ptr(), 1, hi - lo, st.dense_grad.data_ptr(), st.grad_slices.data_ptr(),
                                              self.grad_slices, st.slice_stride, self.comm.cuda_stream), "fold bucket")
            self.model.dp.sum_gradients(st.dense_grad[lo:hi])
            self.ev_comm.record(self.comm)

    # ---- static inputs / outputs
    X = property(lambda self: self.gather.X)
    y = property(lambda self: self.heads.y)
    pred = property(lambda self: self.heads.pred)
    loss = property(lambda self: self.heads.loss)

    def forward(self, training: bool, stream: Optional[int] = None) -> None:
        stream = torch.cuda.current_stream().cuda_stream if stream is None else stream
        for s in self.stages:
            s.forward(stream, training)

    # ---- differentiable forward (autograd): model(x) -> probabilities with a grad_fn
    def autograd_forward(self, training: bool) -> None:
        """Forward of every stage; the heads only predict (their backward waits for autograd's upstream gradient)."""
        stream = torch.cuda.current_stream().cuda_stream
        for s in self.stages:
            s.forward(stream, False if s is self.heads else training)

    def autograd_backward(self, d_pred: torch.Tensor):
        """Backward of every stage for dL/d(pred); leaves dense gradients in the flat store (gradient slices) and
        returns d(dnn_input) [B, in_dim] for the embedding tables (None when nothing flows there)."""
        stream = torch.cuda.current_stream().cuda_stream
        self.model.store.grad_slices.zero_()
        self.heads.backward_external(stream, d_pred)
        for s in reversed(self.stages):
            if s is not self.gather and s is not self.heads:
                s.backward(stream)
        self.model.store.live_slices = self.grad_slices
        return self.gather.out.grad_tensor() if (self.gather.F_s and self.gather.out.grad_written) else None

    def advance_clock(self, stream: int) -> None:
        """step += 1 and the step's Adam bias corrections (recorded in the history ring when the tables are lazy)."""
        m, lib = self.model, self.b.lib
        if m.lazy_adam:
            L.check(lib.mmlrec_hyper_advance_hist(m.hyper_dev.data_ptr(), m.adam_hist.data_ptr(), m.adam_hist_cap, stream),
                    "hyper_advance")
        else:
            L.check(lib.mmlrec_hyper_advance(m.hyper_dev.data_ptr(), stream), "hyper_advance")

    def train_step(self, stream: Optional[int] = None) -> None:
        """advance clock -> sort ids -> forward (+ fused head backward) -> backward -> optimizer."""
        m = self.model
        stream = torch.cuda.current_stream().cuda_stream if stream is None else stream
        lib = self.b.lib
        self.advance_clock(stream)
        # the id sort only feeds the last backward kernel (K2): fork it onto a side stream so it overlaps
        # the whole forward / backward (a parallel branch of the captured graph)
        main = torch.cuda.current_stream()
        dp = getattr(m, "dp", None)
        sh = getattr(m, "shard", None)
        if sh is not None and self.gather.F_s:
            if self.gather.peer_read:
                sh.flag_barrier(stream)   # every owner finished the previous step's row updates before rows are read
        elif dp is not None:
            dp.gather_rows(self.gather.X, self.gather.X_all)
        self.gather.catch_up(stream)
        if sh is None or not self.gather.F_s:
            self.ev_fork.record(main)
            self.side.wait_event(self.ev_fork)
            self.gather.sort(self.side.cuda_stream)
            self.ev_join.record(self.side)
        probe_l2 = m.store.l2_coef is not None and not m.store.l2_probed and not torch.cuda.is_current_stream_capturing()
        if probe_l2:
            # first (eager) step with L2 regularisation: find the gradient entries that NO backward kernel writes (PLE's
            # allocated-but-unused shared experts, a dead gate): the reg kernel must assign those, not accumulate
            m.store.dense_grad[:m.store.n_dense].fill_(float("nan"))
        for s in self.stages:
            s.forward(stream, True)
            if s is self.gather and sh is not None and self.gather.F_s:
                # sharded tables: the keys to sort arrived in the forward exchange; the sort (+ stamp + dense-Adam
                # sweep) runs beside the rest of forward / backward once the owner has served them
                self.side.wait_event(self.gather.ev_served)
                self.gather.sort(self.side.cuda_stream)
                self.ev_join.record(self.side)
        for s in reversed(self.stages):
            if s is not self.gather:
                s.backward(stream)
                if dp is not None and id(s) in self.bucket_after:
                    self._reduce_bucket(self.bucket_after[id(s)], main)
        st = m.store
        p = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        if probe_l2:
            g = st.dense_grad[:st.n_dense]
            unwritten = torch.isnan(g)
            g[unwritten] = 0.0
            st.l2_coef[unwritten] = -st.l2_coef[unwritten].abs()
            st.l2_probed = True

        def dense_step(n_slices, grad_ptr=None):
            if st.l2_coef is not None:   # reg gradient into the buffer the optimizer reads (slice 0), reg value beside the loss
                L.check(lib.mmlrec_l2_regularize(st.dense.data_ptr(), grad_ptr or st.dense_grad.data_ptr(),
                                                 st.l2_coef.data_ptr(), st.n_dense, self.reg_loss.data_ptr(),
                                                 self.l2_scratch.data_ptr(), stream), "l2 regularisation")
            L.check(lib.mmlrec_dense_optimizer_step_sliced(
                st.dense.data_ptr(), grad_ptr or st.dense_grad.data_ptr(), p(st.dense_s1), p(st.dense_s2), st.n_dense,
                m.hyper_dev.data_ptr(), p(st.dense_bf16), n_slices, st.slice_stride, stream), "dense_optimizer_step")

        def fold_slices():
            """data parallel: the all-reduce wants ONE gradient buffer -> slice 0 += slices 1..S-1 (fixed order)"""
            if self.grad_slices > 1:
                L.check(lib.mmlrec_sum_slices(self.fold_seg.data_ptr(), 1, self.first_bucket_hi, st.dense_grad.data_ptr(),
                                              st.grad_slices.data_ptr(), self.grad_slices, st.slice_stride, stream),
                        "fold gradient slices")

        def reduce_rest():
            """the bucket that is final only now (the first stages' parameters), on the main stream; then join the
            buckets that went out on the communication stream during the backward pass"""
            dp.sum_gradients(st.dense_grad if not self.buckets else st.dense_grad[:self.first_bucket_hi])
            if self.buckets:
                main.wait_event(self.ev_comm)

        if dp is not None and sh is None:
            # replicated tables: collectives stay on the main stream in program order
            main.wait_event(self.ev_join)
            self.gather.backward(stream)
            fold_slices()
            reduce_rest()
            dense_step(1)
            return
        # the table update (K2) and the dense optimizer touch disjoint buffers: K2 runs on the side stream, behind the
        # sort / sweep it depends on, while the main stream runs the dense optimizer
        peer_ar = sh is not None and sh.grad_in is not None
        if peer_ar:
            # dense-gradient all-reduce over peer memory: the slices are folded straight into the exchange buffer, the
            # reduced gradient lands in grad_out on every rank (bit-identical), which the optimizer then reads; its
            # first flag barrier is also the barrier between the gradient-row pushes and the owners' K2
            self.gather.backward(stream)
            L.check(lib.mmlrec_sum_slices(self.fold_all.data_ptr(), 1, st.slice_stride, sh.grad_in.ptr,
                                          st.grad_slices.data_ptr(), self.grad_slices, st.slice_stride, stream),
                    "fold gradient slices -> exchange buffer")
            sh.allreduce_gradients(stream)
        elif sh is not None:
            self.gather.backward(stream)        # push gradient rows to their owners
            fold_slices()
            reduce_rest()                       # also the barrier between the pushes and the owners' K2
        self.ev_bwd.record(main)
        self.side.wait_event(self.ev_bwd)
        if sh is not None:
            self.gather.post_reduce(self.side.cuda_stream)
        else:
            self.gather.backward(self.side.cuda_stream)
        self.ev_join2.record(self.side)
        if peer_ar:
            dense_step(1, sh.grad_out.ptr)
        else:
            dense_step(1 if sh is not None else self.grad_slices)
        main.wait_event(self.ev_join2)

    def capture(self) -> None:
        """Capture train_step into a CUDA graph (after one eager warm-up run has happened)."""
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                self.train_step()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = g

    def replay(self) -> None:
        self.graph.replay()
