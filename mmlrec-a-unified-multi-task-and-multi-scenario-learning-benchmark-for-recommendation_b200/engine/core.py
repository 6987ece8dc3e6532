"""Static step program.

A model describes its step ONCE per batch size as a list of *stages* (gather, grouped linear
layers, gate-mix, heads) wired by ``Act`` handles (column slices of wide row-major device
buffers).  Building the program allocates every workspace buffer and uploads every kernel's
problem table; running it is a fixed sequence of C-ABI launches with no allocation, no host
synchronisation and no Python-side tensor math, so the whole step is captured in one CUDA graph.

Gradient convention: ``Act.gbuf`` holds dL/d(pre-activation) of the stage that PRODUCED the
activation -- every kernel that writes an input gradient applies the producer's ReLU mask itself
(``Act.relu``), which fuses ``threshold_backward`` into the GEMM / gate / head epilogues.  The first
writer of a gradient assigns, later writers accumulate (``Act.grad_written``).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .. import lib as L
from .store import FlatStore, _align


class Act:
    """[B, width] activation = columns [col, col+width) of ``buf`` (fp32 [B, ld]); ``gbuf`` is its
    gradient twin with the same geometry."""

    def __init__(self, buf: Optional[torch.Tensor], gbuf: Optional[torch.Tensor], col: int, width: int,
                 relu: bool = False, needs_grad: bool = True, name: str = ""):
        self.buf, self.gbuf, self.col, self.width = buf, gbuf, col, width
        self.relu, self.needs_grad, self.name = relu, needs_grad, name
        self.grad_written = False

    @property
    def ld(self) -> int:
        return self.buf.stride(0)

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr() + 4 * self.col

    @property
    def gptr(self) -> int:
        return self.gbuf.data_ptr() + 4 * self.col

    def same_as(self, o: "Act") -> bool:
        return self.buf is o.buf and self.col == o.col and self.width == o.width

    def tensor(self) -> torch.Tensor:
        return self.buf[:, self.col:self.col + self.width]

    def grad_tensor(self) -> torch.Tensor:
        return self.gbuf[:, self.col:self.col + self.width]


class LinearSpec:
    """One nn.Linear (+ optional BatchNorm1d) as the program sees it."""

    def __init__(self, x: Act, linear: nn.Module, bn: Optional[nn.Module] = None):
        self.x, self.linear, self.bn = x, linear, bn
        self.W: nn.Parameter = linear.weight
        self.b: Optional[nn.Parameter] = getattr(linear, "bias", None)
        self.N, self.K = self.W.shape

    def params(self) -> List[nn.Parameter]:
        return [self.W] + ([self.b] if self.b is not None else [])


class Builder:
    """Collects stages.  In ``dry`` mode nothing is allocated: only the order in which parameters
    are consumed is recorded (it becomes the flat-store layout)."""

    def __init__(self, B: int, device, store: Optional[FlatStore], dry: bool):
        self.B, self.device, self.store, self.dry = B, device, store, dry
        self.stages: List["Stage"] = []
        self.param_order: List[nn.Parameter] = []
        self.buffer_order: List[torch.Tensor] = []
        self.keep: List[object] = []       # tensors whose addresses are baked into tables
        self.lib = None if dry else L.load()

    # ---- allocation helpers
    def zeros(self, *shape, dtype=torch.float32) -> Optional[torch.Tensor]:
        if self.dry:
            return None
        t = torch.zeros(*shape, dtype=dtype, device=self.device)
        self.keep.append(t)
        return t

    def table(self, structs: Sequence[C.Structure]) -> Optional[torch.Tensor]:
        if self.dry:
            return None
        raw = L.struct_bytes(structs)
        t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device)
        self.keep.append(t)
        return t

    def ints(self, values: Sequence[int], dtype=torch.int32) -> Optional[torch.Tensor]:
        if self.dry:
            return None
        t = torch.tensor(list(values), dtype=dtype, device=self.device)
        self.keep.append(t)
        return t

    def new_act_buffer(self, widths: Sequence[int], relu: bool, name: str, needs_grad: bool = True) -> List[Act]:
        """One wide buffer (and its gradient twin) holding several activations side by side."""
        total = _align(sum(widths), 4)
        buf = self.zeros(self.B, total)
        gbuf = self.zeros(self.B, total) if needs_grad else None
        acts, at = [], 0
        for i, w in enumerate(widths):
            acts.append(Act(buf, gbuf, at, w, relu=relu, needs_grad=needs_grad, name=f"{name}[{i}]"))
            at += w
        return acts

    def note_params(self, params: Sequence[Optional[nn.Parameter]]) -> None:
        self.param_order.extend(p for p in params if p is not None)

    def note_buffers(self, bufs: Sequence[Optional[torch.Tensor]]) -> None:
        self.buffer_order.extend(t for t in bufs if t is not None)

    def add(self, stage: "Stage") -> "Stage":
        self.stages.append(stage)
        return stage


class Stage:
    name = "stage"

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
        (self.out,) = b.new_act_buffer([self.in_dim], relu=False, name="dnn_input")
        if b.dry:
            return
        st = b.store
        emb, D = model.embedding_layout, model.emb_dim
        meta = []
        for f in emb:  # (param, vocab, x_col, out_col)
            meta += [f[0]._mm_off, f[1], f[2], f[3]]
        self.F_s, self.D = len(emb), D
        self.meta = b.ints(meta if meta else [0, 0, 0, 0], dtype=torch.int64)
        self.dense_cols = b.ints(model.dense_x_cols if model.dense_x_cols else [0])
        self.F_d = len(model.dense_x_cols)
        self.dense_out_col = self.F_s * D
        self.X = b.zeros(b.B, model.num_x_cols)
        self.oob = b.zeros(1, dtype=torch.int32)
        if self.F_s:
            n_pad = 32
            while n_pad < b.B:
                n_pad <<= 1
            self.sorted_ids = b.zeros(self.F_s, b.B, dtype=torch.int32)
            self.sorted_pos = b.zeros(self.F_s, b.B, dtype=torch.int32)
            self.keys_ws = b.zeros(self.F_s * n_pad, dtype=torch.int64)

    def forward(self, stream, training):
        b, st = self.b, self.b.store
        L.check(b.lib.mmlrec_gather_concat(
            self.X.data_ptr(), self.X.stride(0), b.B, st.emb.data_ptr(), self.meta.data_ptr(), self.F_s, self.D,
            self.dense_cols.data_ptr(), self.F_d, self.dense_out_col, self.out.ptr, self.out.ld, None, 0,
            self.oob.data_ptr(), stream), "gather_concat")

    def sort(self, stream):
        if self.F_s:
            b = self.b
            L.check(b.lib.mmlrec_sort_field_ids(self.X.data_ptr(), self.X.stride(0), b.B, self.meta.data_ptr(),
                                                self.F_s, self.sorted_ids.data_ptr(), self.sorted_pos.data_ptr(),
                                                self.keys_ws.data_ptr(), stream), "sort_field_ids")

    def backward(self, stream):
        if not self.F_s or not self.out.grad_written:
            return
        b, st, hy = self.b, self.b.store, self.model.hyper_dev
        p = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        L.check(b.lib.mmlrec_emb_backward_update(
            self.out.gptr, self.out.gbuf.stride(0), b.B, self.sorted_ids.data_ptr(), self.sorted_pos.data_ptr(),
            self.meta.data_ptr(), self.F_s, self.D, st.emb.data_ptr(), p(st.emb_s1), p(st.emb_s2), p(st.row_touch),
            hy.data_ptr(), None, stream), "emb_backward_update")
        if self.model.optimizer_name == "adam":
            L.check(b.lib.mmlrec_emb_adam_dense_sweep(st.emb.data_ptr(), st.emb_s1.data_ptr(), st.emb_s2.data_ptr(),
                                                      st.row_touch.data_ptr(), st.n_emb // self.D, self.D,
                                                      hy.data_ptr(), stream), "emb_adam_dense_sweep")


# ----------------------------------------------------------------------------------------------
# K3: grouped linear layers
# ----------------------------------------------------------------------------------------------
class _Group:
    """Adjacent members of a stage that read the same input and whose parameters are contiguous:
    executed as one wide problem."""

    def __init__(self, members: List[LinearSpec], y_col: int):
        self.members, self.y_col = members, y_col
        self.x = members[0].x
        self.K = members[0].K
        self.N = sum(m.N for m in members)
        self.W, self.b = members[0].W, members[0].b
        self.has_bn = members[0].bn is not None


def _tile_prefix(problems: Sequence[L.GemmF32]) -> Tuple[List[int], int]:
    pre, at = [0], 0
    for p in problems:
        at += ((p.M + 63) // 64) * ((p.N + 63) // 64)
        pre.append(at)
    return pre, at


class LinearStage(Stage):
    """A set of independent ``y = act([bn](x W^T + b))`` layers run by ONE grouped-GEMM launch
    (model/utils.py:146-161), and in backward by one launch holding every wgrad + dgrad problem."""
    name = "linear"

    def __init__(self, b: Builder, specs: List[LinearSpec], act: Optional[str], label: str = ""):
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
        relu = act == "relu"
        self.outs = b.new_act_buffer([s.N for s in specs], relu=relu, name=f"{label}.y")
        self.zs = b.new_act_buffer([s.N for s in specs], relu=False, name=f"{label}.z") if self.use_bn else None
        if b.dry:
            return
        st = b.store
        # ---- merge adjacent members sharing x with contiguous parameters
        self.groups: List[_Group] = []
        cur: List[LinearSpec] = []
        for s in specs:
            ok = bool(cur) and s.x.same_as(cur[-1].x) and s.K == cur[-1].K and st.contiguous_after(cur[-1].W, s.W) \
                and ((s.b is None) == (cur[-1].b is None)) and (s.b is None or st.contiguous_after(cur[-1].b, s.b))
            if ok and self.use_bn:
                ok = st.contiguous_after(cur[-1].bn.weight, s.bn.weight) and st.contiguous_after(cur[-1].bn.bias, s.bn.bias) \
                    and cur[-1].bn.running_mean._mm_off + cur[-1].N == s.bn.running_mean._mm_off \
                    and cur[-1].bn.running_var._mm_off + cur[-1].N == s.bn.running_var._mm_off
            if ok:
                cur.append(s)
            else:
                if cur:
                    self.groups.append(_Group(cur, self.outs[specs.index(cur[0])].col))
                cur = [s]
        self.groups.append(_Group(cur, self.outs[specs.index(cur[0])].col))
        ybuf = self.outs[0].buf
        fwd_target = self.zs[0].buf if self.use_bn else ybuf
        probs = []
        for g in self.groups:
            p = L.GemmF32()
            p.A, p.a_rs, p.a_cs = g.x.ptr, g.x.ld, 1
            p.B, p.b_rs, p.b_cs = g.W.data_ptr(), g.W._mm_ld, 1
            p.C, p.ldc = fwd_target.data_ptr() + 4 * g.y_col, fwd_target.stride(0)
            p.bias = g.b.data_ptr() if g.b is not None else None
            p.M, p.N, p.K = b.B, g.N, g.K
            p.act = L.ACT_NONE if self.use_bn else L.ACT_CODES[act]
            probs.append(p)
        self.fwd_table = b.table(probs)
        pre, self.fwd_tiles = _tile_prefix(probs)
        self.fwd_prefix = b.ints(pre)
        self.n_fwd = len(probs)
        if self.use_bn:
            n_total = sum(g.N for g in self.groups)
            self.save_mean = b.zeros(_align(n_total, 4))
            self.save_invstd = b.zeros(_align(n_total, 4))

    # ---- forward
    def forward(self, stream, training):
        b = self.b
        L.check(b.lib.mmlrec_gemm_grouped_f32(self.fwd_table.data_ptr(), self.fwd_prefix.data_ptr(), self.n_fwd,
                                              self.fwd_tiles, stream), f"linear fwd {self.label}")
        if self.use_bn:
            zbuf, ybuf = self.zs[0].buf, self.outs[0].buf
            for g in self.groups:
                bn0 = g.members[0].bn
                L.check(b.lib.mmlrec_bn_forward(
                    zbuf.data_ptr() + 4 * g.y_col, zbuf.stride(0), b.B, g.N, bn0.weight.data_ptr(), bn0.bias.data_ptr(),
                    bn0.running_mean.data_ptr(), bn0.running_var.data_ptr(), bn0.num_batches_tracked.data_ptr(),
                    len(g.members), self.save_mean.data_ptr() + 4 * g.y_col, self.save_invstd.data_ptr() + 4 * g.y_col,
                    ybuf.data_ptr() + 4 * g.y_col, ybuf.stride(0), None, 0, L.ACT_CODES[self.act], 1 if training else 0,
                    stream), f"bn fwd {self.label}")

    # ---- backward
    def plan_backward(self):
        b, st = self.b, self.b.store
        self.live_groups: List[_Group] = []
        waves: List[List[L.GemmF32]] = [[]]
        gy, gz = self.outs[0].gbuf, (self.zs[0].gbuf if self.use_bn else None)
        for g in self.groups:
            outs = [self.outs[self.specs.index(m)] for m in g.members]
            if not any(o.grad_written for o in outs):
                continue  # nothing flows into this group: parameters keep a zero gradient
            for o in outs:
                if not o.grad_written:
                    o.grad_tensor().zero_()  # stays zero: the buffer is never written afterwards
            self.live_groups.append(g)
            dz = gz if self.use_bn else gy
            dz_ptr, dz_ld = dz.data_ptr() + 4 * g.y_col, dz.stride(0)
            # wgrad: dW[n,k] = sum_b dZ[b,n] X[b,k]; rowsum_a = bias gradient
            p = L.GemmF32()
            p.A, p.a_rs, p.a_cs = dz_ptr, 1, dz_ld
            p.B, p.b_rs, p.b_cs = g.x.ptr, 1, g.x.ld
            p.C, p.ldc = st.grad_ptr(g.W), g.W._mm_ld
            p.rowsum_a = st.grad_ptr(g.b) if g.b is not None else None
            p.M, p.N, p.K = g.N, g.K, b.B
            waves[0].append(p)
            # dgrad: dX[b,k] = sum_n dZ[b,n] W[n,k]  (masked by the producer's ReLU)
            if g.x.needs_grad and g.x.gbuf is not None:
                q = L.GemmF32()
                q.A, q.a_rs, q.a_cs = dz_ptr, dz_ld, 1
                q.B, q.b_rs, q.b_cs = g.W.data_ptr(), 1, g.W._mm_ld
                q.C, q.ldc = g.x.gptr, g.x.gbuf.stride(0)
                q.M, q.N, q.K = b.B, g.K, g.N
                if g.x.relu:
                    q.mask, q.ldmask = g.x.ptr, g.x.ld
                q.accumulate = 1 if g.x.grad_written else 0
                # two problems of one launch must not write the same gradient: later ones wait a wave
                w = sum(1 for gg in self.live_groups[:-1] if gg.x.same_as(g.x) and gg.x.needs_grad)
                while len(waves) <= w:
                    waves.append([])
                if w > 0:
                    q.accumulate = 1
                waves[w].append(q)
                g.x.grad_written = True
        self.bwd = []
        for wv in waves:
            if wv:
                pre, tiles = _tile_prefix(wv)
                self.bwd.append((b.table(wv), b.ints(pre), len(wv), tiles))

    def backward(self, stream):
        b = self.b
        if self.use_bn:
            zbuf, gy, gz = self.zs[0].buf, self.outs[0].gbuf, self.zs[0].gbuf
            for g in self.live_groups:
                bn0 = g.members[0].bn
                L.check(b.lib.mmlrec_bn_backward(
                    gy.data_ptr() + 4 * g.y_col, gy.stride(0), zbuf.data_ptr() + 4 * g.y_col, zbuf.stride(0), b.B, g.N,
                    bn0.weight.data_ptr(), self.save_mean.data_ptr() + 4 * g.y_col,
                    self.save_invstd.data_ptr() + 4 * g.y_col, gz.data_ptr() + 4 * g.y_col, gz.stride(0), None, 0,
                    b.store.grad_ptr(bn0.weight), b.store.grad_ptr(bn0.bias), stream), f"bn bwd {self.label}")
        for table, prefix, n, tiles in self.bwd:
            L.check(b.lib.mmlrec_gemm_grouped_f32(table.data_ptr(), prefix.data_ptr(), n, tiles, stream),
                    f"linear bwd {self.label}")


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
# gate head + softmax + mixture
# ----------------------------------------------------------------------------------------------
class GateSpec:
    def __init__(self, gate_in: Act, head: nn.Module, experts: List[Act]):
        self.gate_in, self.head, self.experts = gate_in, head, experts


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
        self.outs = b.new_act_buffer([H] * len(gates), relu=False, name=f"{label}.mix")
        if b.dry:
            return
        self.probs = [b.zeros(b.B, len(g.experts)) for g in gates]
        self._fill_tables(backward=False)

    def _gate_record(self, i: int, backward: bool) -> L.Gate:
        g, st = self.gates[i], self.b.store
        r = L.Gate()
        r.gate_in, r.ld_gate_in, r.Hg, r.n_e = g.gate_in.ptr, g.gate_in.ld, g.gate_in.width, len(g.experts)
        assert r.n_e <= L.MAX_GATE_EXPERTS
        r.Wg, r.ld_Wg = g.head.weight.data_ptr(), g.head.weight._mm_ld
        for e, a in enumerate(g.experts):
            r.expert[e] = a.ptr
        assert all(a.ld == g.experts[0].ld for a in g.experts)
        r.ld_expert, r.H = g.experts[0].ld, self.H
        r.probs = self.probs[i].data_ptr()
        r.mix, r.ld_mix = self.outs[i].ptr, self.outs[i].ld
        if backward and self.outs[i].grad_written:
            r.d_mix, r.ld_d_mix = self.outs[i].gptr, self.outs[i].gbuf.stride(0)
            gi = g.gate_in
            if gi.needs_grad and gi.gbuf is not None:
                r.d_gate_in, r.ld_d_gate_in = gi.gptr, gi.gbuf.stride(0)
                r.relu_mask_gate_in = 1 if gi.relu else 0
                r.accumulate_d_gate_in = 1 if gi.grad_written else 0
                gi.grad_written = True
            r.dWg = st.grad_ptr(g.head.weight)
        return r

    def _fill_tables(self, backward: bool):
        self.gate_table = self.b.table([self._gate_record(i, backward) for i in range(len(self.gates))])

    def forward(self, stream, training):
        L.check(self.b.lib.mmlrec_gate_mix_forward(self.gate_table.data_ptr(), len(self.gates), self.b.B, stream),
                f"gate_mix fwd {self.label}")

    def plan_backward(self):
        b = self.b
        live = [i for i in range(len(self.gates)) if self.outs[i].grad_written]
        self.any_live = bool(live)
        if not live:
            return
        # gates that share an input (no gate DNN: every head reads the level input) must not race on
        # d(gate_in): the kernel then walks the gates in order inside each CTA
        seen, self.serialize = [], 0
        for i in live:
            gi = self.gates[i].gate_in
            if any(gi.same_as(o) for o in seen):
                self.serialize = 1
            seen.append(gi)
        self._fill_tables(backward=True)
        # experts: every distinct expert activation gets d = sum over its user gates of p * d_mix
        recs, uniq = [], []
        for i in live:
            for a in self.gates[i].experts:
                if not any(a.same_as(u) for u in uniq):
                    uniq.append(a)
        for a in uniq:
            r = L.ExpertGrad()
            assert not a.grad_written, "an expert output consumed elsewhere must be accumulated"
            r.expert, r.ld_expert = a.ptr, a.ld
            r.d_expert, r.ld_d_expert, r.H = a.gptr, a.gbuf.stride(0), self.H
            r.relu_mask = 1 if a.relu else 0
            n = 0
            for i in live:
                for e, ea in enumerate(self.gates[i].experts):
                    if ea.same_as(a):
                        r.user_probs[n] = self.probs[i].data_ptr()
                        r.user_prob_ld[n], r.user_prob_col[n] = self.probs[i].stride(0), e
                        r.user_d_mix[n], r.user_d_mix_ld[n] = self.outs[i].gptr, self.outs[i].gbuf.stride(0)
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
        L.check(b.lib.mmlrec_gate_mix_backward(self.gate_table.data_ptr(), len(self.gates), self.expert_table.data_ptr(),
                                               self.n_expert_recs, b.B, self.max_ne, self.max_hg, self.serialize,
                                               self.scratch.data_ptr(), self.counters.data_ptr(), stream),
                f"gate_mix bwd {self.label}")


# ----------------------------------------------------------------------------------------------
# heads + loss
# ----------------------------------------------------------------------------------------------
class HeadSpec:
    def __init__(self, h: Act, final: nn.Module, bias: Optional[nn.Parameter], task: str):
        self.h, self.final, self.bias, self.task = h, final, bias, task


class HeadStage(Stage):
    """Bias-free 1-unit head + PredictionLayer + sum-BCE, forward and backward in one kernel
    (model/mmoe.py:97-100, model/utils.py:242-248, model/basemodel.py:294-296)."""
    name = "heads"

    def __init__(self, b: Builder, heads: List[HeadSpec], esmm: bool = False):
        self.b, self.heads, self.esmm = b, heads, esmm
        b.note_params([h.final.weight for h in heads])
        b.note_params([h.bias for h in heads])
        self.T = len(heads)
        if b.dry:
            return
        st = b.store
        self.y = b.zeros(b.B, self.T)
        self.pred = b.zeros(b.B, self.T)
        self.loss = b.zeros(self.T + 1)
        recs = []
        for h in heads:
            r = L.Head()
            r.h, r.ld_h, r.H = h.h.ptr, h.h.ld, h.h.width
            r.kind = L.HEAD_SIGMOID_BCE if h.task == "binary" else L.HEAD_IDENTITY_MSE
            r.w = h.final.weight.data_ptr()
            r.bias = h.bias.data_ptr() if h.bias is not None else None
            if h.h.needs_grad and h.h.gbuf is not None:
                assert not h.h.grad_written
                r.d_h, r.ld_d_h, r.relu_mask = h.h.gptr, h.h.gbuf.stride(0), 1 if h.h.relu else 0
                h.h.grad_written = True
            r.dw = st.grad_ptr(h.final.weight)
            r.dbias = st.grad_ptr(h.bias) if h.bias is not None else None
            recs.append(r)
        self.table = b.table(recs)
        max_h = max(h.h.width for h in heads)
        self.scratch = b.zeros(int(b.lib.mmlrec_heads_scratch(self.T, max_h, b.B)))
        self.counter = b.zeros(1, dtype=torch.int32)

    def forward(self, stream, training):
        b = self.b
        L.check(b.lib.mmlrec_heads_forward_backward(
            self.table.data_ptr(), self.T, b.B, self.y.data_ptr() if training else None, self.T, self.pred.data_ptr(),
            self.T, self.loss.data_ptr(), 1 if self.esmm else 0, 1 if training else 0, self.scratch.data_ptr(),
            self.scratch.numel(), self.counter.data_ptr(), stream), "heads")


# ----------------------------------------------------------------------------------------------
# the program
# ----------------------------------------------------------------------------------------------
class StepPlan:
    """The compiled step of one model at one batch size."""

    def __init__(self, model, B: int):
        self.model, self.B = model, B
        self.b = Builder(B, model.device_obj, model.store, dry=False)
        model.build_graph(self.b)
        self.stages = self.b.stages
        self.gather: GatherStage = next(s for s in self.stages if isinstance(s, GatherStage))
        self.heads: HeadStage = next(s for s in self.stages if isinstance(s, HeadStage))
        for s in reversed(self.stages):
            s.plan_backward()
        self.graph: Optional[torch.cuda.CUDAGraph] = None

    # ---- static inputs / outputs
    @property
    def X(self) -> torch.Tensor:
        return self.gather.X

    @property
    def y(self) -> torch.Tensor:
        return self.heads.y

    @property
    def pred(self) -> torch.Tensor:
        return self.heads.pred

    @property
    def loss(self) -> torch.Tensor:
        return self.heads.loss

    def forward(self, training: bool, stream: Optional[int] = None) -> None:
        stream = torch.cuda.current_stream().cuda_stream if stream is None else stream
        for s in self.stages:
            s.forward(stream, training)

    def train_step(self, stream: Optional[int] = None) -> None:
        """advance clock -> sort ids -> forward (+ fused head backward) -> backward -> optimizer."""
        m = self.model
        stream = torch.cuda.current_stream().cuda_stream if stream is None else stream
        lib = self.b.lib
        L.check(lib.mmlrec_hyper_advance(m.hyper_dev.data_ptr(), stream), "hyper_advance")
        self.gather.sort(stream)
        for s in self.stages:
            s.forward(stream, True)
        for s in reversed(self.stages):
            s.backward(stream)
        st = m.store
        p = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        L.check(lib.mmlrec_dense_optimizer_step(st.dense.data_ptr(), st.dense_grad.data_ptr(), p(st.dense_s1),
                                                p(st.dense_s2), st.n_dense, m.hyper_dev.data_ptr(), p(st.dense_bf16),
                                                stream), "dense_optimizer_step")

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
