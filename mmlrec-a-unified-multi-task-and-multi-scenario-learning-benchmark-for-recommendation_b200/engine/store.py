"""Flat parameter store: every tensor the optimizer touches lives in a few big device buffers.

Layout in HBM (all fp32 unless noted):
  * ``dense``      every non-embedding parameter, in the order the step program consumes them, so
                   that the Linear layers of one stage that read the same input are adjacent and
                   run as ONE wide GEMM (their weights form a single [sum N, K] matrix);
  * ``dense_grad`` same geometry; written by the wgrad kernels, read by the optimizer / all-reduce;
  * ``dense_s1/s2`` optimizer state (Adagrad sum / Adam exp_avg, exp_avg_sq);
  * ``dense_bf16`` bf16 shadow of ``dense`` for the tensor-core mode (refreshed by the optimizer);
  * ``emb``        all embedding tables back to back, rows of D floats, plus ``emb_s1/s2`` and the
                   per-row ``row_touch`` stamp used by the dense-Adam sweep;
  * ``stats`` / ``counts``  BatchNorm running statistics (fp32) and num_batches_tracked (int64).

The ``nn.Parameter`` / buffer objects of the module tree are re-pointed at views of these buffers,
so ``state_dict()`` / ``load_state_dict()`` keep the reference's key names and shapes.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch
import torch.nn as nn


def _align(n: int, a: int) -> int:
    return (n + a - 1) // a * a


class FlatStore:
    MAX_GRAD_SLICES = 4

    def __init__(self, model: nn.Module, ordered: Iterable[nn.Parameter], emb_params: List[nn.Parameter],
                 device: torch.device, want_bf16: bool, ordered_buffers: Iterable[torch.Tensor] = (),
                 aux_floats: int = 0, emb_alloc=None):
        self.device = device
        emb_ids = {id(p) for p in emb_params}
        seen, order = set(), []
        for p in ordered:
            if id(p) not in seen and id(p) not in emb_ids:
                seen.add(id(p))
                order.append(p)
        for p in model.parameters():
            if id(p) not in seen and id(p) not in emb_ids:
                seen.add(id(p))
                order.append(p)
        # ---- dense parameters.  Matrices are stored with their row stride padded to 8 elements and
        # start on an 8-element boundary (16 B in the bf16 shadow: a legal TMA operand); vectors are
        # packed back to back so the biases / BatchNorm vectors of fused layers form one vector.
        offs, lds, at = {}, {}, 0
        for p in order:
            if p.dim() == 2:
                at = _align(at, 8)
                # `_mm_nopad`: matrices only touched by element-wise kernels (STAR's [in,out] tensors) stay dense
                lds[id(p)] = p.shape[1] if getattr(p, "_mm_nopad", False) else _align(p.shape[1], 8)
                offs[id(p)] = at
                at += p.shape[0] * lds[id(p)]
            else:
                lds[id(p)] = 0
                offs[id(p)] = at
                at += p.numel()
        self.n_dense = _align(max(at, 8), 8)          # what the optimizer owns
        # aux region behind the parameters: derived weights recomputed every step (STAR's spec*shared) and
        # their gradients; same offsets in dense / dense_grad / dense_bf16, never touched by the optimizer
        self.aux_base, self.aux_floats, self._aux_at = self.n_dense, _align(aux_floats, 8), 0
        self.dense = torch.zeros(self.n_dense + self.aux_floats, dtype=torch.float32, device=device)
        # gradient slices: slice 0 is THE gradient buffer; in tensor-core mode the split-K wgrad problems write the
        # partial tiles of batch slices 1..S-1 into the further slices (same offsets) and the optimizer adds them up
        # (fp32 mode: no split-K; the second slice serves a Linear applied twice in one stage, LinearStage.plan_backward)
        self.max_grad_slices = self.MAX_GRAD_SLICES if want_bf16 else 2
        self.slice_stride = self.n_dense + self.aux_floats
        self.grad_slices = torch.zeros(self.max_grad_slices, self.slice_stride, dtype=torch.float32, device=device)
        self.dense_grad = self.grad_slices[0]
        self.live_slices = 1          # slices the last executed step program wrote (engine/core.py StepPlan)
        self.l2_coef = None           # per-element l2 of the registered weights (set_l2), None while nothing is regularised
        self.l2_probed = False        # the sign of l2_coef marks entries no backward kernel writes (StepPlan.train_step)
        self.dense_s1: Optional[torch.Tensor] = None
        self.dense_s2: Optional[torch.Tensor] = None
        self.dense_bf16 = (torch.zeros(self.n_dense + self.aux_floats, dtype=torch.bfloat16, device=device)
                           if want_bf16 else None)
        # ---- embedding tables
        eoffs, at = {}, 0
        for p in emb_params:
            eoffs[id(p)] = at
            at += p.numel()
        self.n_emb = at
        # `emb_alloc`: row-sharded tables live in IPC-exported memory the peers can address (parallel.ShardContext)
        self.emb = (emb_alloc(max(at, 4), device) if emb_alloc is not None
                    else torch.zeros(max(at, 4), dtype=torch.float32, device=device))
        self.emb_s1: Optional[torch.Tensor] = None
        self.emb_s2: Optional[torch.Tensor] = None
        self.row_touch: Optional[torch.Tensor] = None
        self.emb_dim = emb_params[0].shape[1] if emb_params else 4
        self.emb_offset: Dict[int, int] = {}
        # ---- buffers (BatchNorm statistics)
        # statistics of layers fused into one stage are adjacent: [means...][vars...]
        rank = {id(t): i for i, t in enumerate(ordered_buffers)}
        allb = [(m, n, b) for m in model.modules() for n, b in m._buffers.items() if b is not None]
        allb.sort(key=lambda e: rank.get(id(e[2]), len(rank)))
        fbufs = [e for e in allb if e[2].dtype == torch.float32]
        ibufs = [e for e in allb if e[2].dtype == torch.int64]
        self.stats = torch.zeros(max(sum(b.numel() for _, _, b in fbufs), 4), dtype=torch.float32, device=device)
        self.counts = torch.zeros(max(sum(b.numel() for _, _, b in ibufs), 1), dtype=torch.int64, device=device)

        # ---- re-point the module tree
        owners = {}
        for m in model.modules():
            for n, p in m._parameters.items():
                if p is not None:
                    owners.setdefault(id(p), []).append((m, n))
        self.params: List[nn.Parameter] = []
        with torch.no_grad():
            for p in order:
                o, ld = offs[id(p)], lds[id(p)]
                if p.dim() == 2:
                    view = self.dense[o:o + p.shape[0] * ld].view(p.shape[0], ld)[:, :p.shape[1]]
                else:
                    view = self.dense[o:o + p.numel()].view(p.shape)
                view.copy_(p.detach().to(device))
                newp = nn.Parameter(view, requires_grad=p.requires_grad)
                newp._mm_off = o
                newp._mm_ld = ld
                newp._mm_span = p.shape[0] * ld if p.dim() == 2 else p.numel()
                newp._mm_kind = "dense"
                for m, n in owners[id(p)]:
                    m._parameters[n] = newp
                self.params.append(newp)
            self.emb_params: List[nn.Parameter] = []
            for p in emb_params:
                o = eoffs[id(p)]
                view = self.emb[o:o + p.numel()].view(p.shape)
                view.copy_(p.detach().to(device))
                newp = nn.Parameter(view, requires_grad=p.requires_grad)
                newp._mm_off = o
                newp._mm_ld = p.shape[1]
                newp._mm_span = p.numel()
                newp._mm_kind = "emb"
                for m, n in owners[id(p)]:
                    m._parameters[n] = newp
                self.emb_params.append(newp)
            at = 0
            for m, n, b in fbufs:
                view = self.stats[at:at + b.numel()].view(b.shape)
                view.copy_(b.to(device))
                view._mm_off = at
                m._buffers[n] = view
                at += b.numel()
            at = 0
            for m, n, b in ibufs:
                view = self.counts[at:at + b.numel()].view(b.shape)
                view.copy_(b.to(device))
                view._mm_off = at
                m._buffers[n] = view
                at += b.numel()
        if self.dense_bf16 is not None:
            self.refresh_bf16()

    # ------------------------------------------------------------------ derived weights
    def reset_aux(self) -> None:
        self._aux_at = 0

    def aux_matrix(self, rows: int, cols: int) -> torch.Tensor:
        """[rows, cols] fp32 view (row stride padded to 8) in the aux region, tagged like a parameter so the
        GEMM stages can use it as a weight; every step program asks for its derived tensors in the same
        order, so all programs of a model share them."""
        ld = _align(cols, 8) if rows > 1 or cols > 1 else 1
        self._aux_at = _align(self._aux_at, 8)
        off = self.aux_base + self._aux_at
        assert self._aux_at + rows * ld <= self.aux_floats, "aux region too small"
        self._aux_at += rows * ld
        t = self.dense[off:off + rows * ld].view(rows, ld)[:, :cols]
        t._mm_off, t._mm_ld, t._mm_span, t._mm_kind = off, ld, rows * ld, "dense"
        return t

    def aux_vector(self, n: int) -> torch.Tensor:
        off = self.aux_base + self._aux_at
        assert self._aux_at + n <= self.aux_floats, "aux region too small"
        self._aux_at += n
        t = self.dense[off:off + n]
        t._mm_off, t._mm_ld, t._mm_span, t._mm_kind = off, 0, n, "dense"
        return t

    # ------------------------------------------------------------------ views / pointers
    def ptr(self, p: torch.Tensor) -> int:
        return p.data_ptr()

    def grad_ptr(self, p: nn.Parameter) -> int:
        assert p._mm_kind == "dense"
        return self.dense_grad.data_ptr() + 4 * p._mm_off

    def grad_view(self, p: nn.Parameter) -> torch.Tensor:
        """The gradient of `p` left by the last step (the sum of the live gradient slices)."""
        assert p._mm_kind == "dense"
        g = self.dense_grad if self.live_slices == 1 else self.grad_slices[:self.live_slices].sum(0)
        if p.dim() == 2:
            return g[p._mm_off:p._mm_off + p._mm_span].view(p.shape[0], p._mm_ld)[:, :p.shape[1]]
        return g[p._mm_off:p._mm_off + p.numel()].view(p.shape)

    def bf16_ptr(self, p: nn.Parameter) -> int:
        assert self.dense_bf16 is not None and p._mm_kind == "dense"
        return self.dense_bf16.data_ptr() + 2 * p._mm_off

    def contiguous_after(self, a: nn.Parameter, b: nn.Parameter) -> bool:
        """True when ``b`` starts exactly where ``a`` ends (so [a;b] is one matrix)."""
        return (a._mm_kind == b._mm_kind == "dense" and a._mm_off + a._mm_span == b._mm_off
                and a._mm_ld == b._mm_ld)

    def set_l2(self, p: nn.Parameter, l2: float) -> None:
        """Register dense parameter `p` for L2 regularisation (coefficients add up when it is registered twice, as the
        reference's list of weight lists would count it twice)."""
        if self.l2_coef is None:
            self.l2_coef = torch.zeros(self.n_dense, dtype=torch.float32, device=self.device)
        self.l2_coef[p._mm_off:p._mm_off + p._mm_span] += float(l2)

    def refresh_bf16(self) -> None:
        if self.dense_bf16 is not None:
            self.dense_bf16[:self.n_dense].copy_(self.dense[:self.n_dense])

    # ------------------------------------------------------------------ optimizer state
    def ensure_optimizer_state(self, optimizer: str) -> None:
        need1 = optimizer in ("adagrad", "adam", "rmsprop")
        need2 = optimizer == "adam"
        if need1 and self.dense_s1 is None:
            self.dense_s1 = torch.zeros(self.n_dense, dtype=torch.float32, device=self.device)
            self.emb_s1 = torch.zeros_like(self.emb)
        if need2 and self.dense_s2 is None:
            self.dense_s2 = torch.zeros(self.n_dense, dtype=torch.float32, device=self.device)
            self.emb_s2 = torch.zeros_like(self.emb)

        if optimizer in ("adam", "rmsprop") and self.row_touch is None:   # dense zero-gradient sweep (K2 stamps rows)
            self.row_touch = torch.full((max(self.n_emb // max(self.emb_dim, 1), 1),), -1, dtype=torch.int32,
                                        device=self.device)

    def reset_optimizer_state(self) -> None:
        for t in (self.dense_s1, self.dense_s2, self.emb_s1, self.emb_s2):
            if t is not None:
                t.zero_()
        if self.row_touch is not None:
            self.row_touch.fill_(-1)
