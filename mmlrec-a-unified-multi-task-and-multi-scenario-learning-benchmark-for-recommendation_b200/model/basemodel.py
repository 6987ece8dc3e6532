"""Training runtime with the reference's ``BaseModel`` surface (``/root/reference/model/basemodel.py``):
constructor signature (:69-71), ``compile`` (:557-567), ``fit`` (:135-371), ``evaluate`` (:373-393),
``predict`` (:395-457), ``input_from_feature_columns`` / ``compute_input_dim`` (:461-507).

What differs underneath: the step body (``basemodel.py:262-313``) is not a chain of ATen calls but a
static program of hand-written sm_100a kernels (``engine/core.py``) replayed as one CUDA graph;
parameters live in flat device buffers (``engine/store.py``); the embedding tables are updated by
the fused sort + segmented-reduce kernel instead of dense [V,D] gradients; per-batch metrics are
computed after the epoch from predictions kept on the device instead of forcing a host sync every
step.  There is no CPU execution path: any compute call without CUDA + the library raises.
"""
from __future__ import annotations

import copy
import time
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from .. import lib as L
from ..engine.core import Builder, StepPlan
from ..engine.store import FlatStore
from .utils import (DenseFeat, PredictionLayer, SparseFeat, VarLenSparseFeat, build_input_features,
                    create_embedding_matrix, get_mask)


class _FusedOptimizerHandle:
    """What ``model.optim`` holds: the update itself is fused into the CUDA step, so this only names it."""

    def __init__(self, name: str, lr: float):
        self.name, self.lr = name, lr

    def __repr__(self):
        return f"FusedOptimizer({self.name}, lr={self.lr})"


class _FusedForward(torch.autograd.Function):
    """``model(X)`` as one autograd node: forward = the step program's forward kernels, backward = its backward kernels
    fed with autograd's dL/d(pred).  Parameters enter as inputs so that their ``.grad`` is filled like the reference's
    (dense ``[V, D]`` gradients for the embedding tables, as ``nn.Embedding(sparse=False)`` produces).  The program's
    buffers are reused: call ``backward()`` before the next forward of the same batch size."""

    @staticmethod
    def forward(ctx, model, plan, X, *params):
        plan.X.copy_(X, non_blocking=True)
        plan.autograd_forward(model.training)
        ctx.model, ctx.plan, ctx.params = model, plan, params
        return plan.pred.clone()

    @staticmethod
    def backward(ctx, d_pred):
        model, plan = ctx.model, ctx.plan
        d_in = plan.autograd_backward(d_pred.contiguous())
        st = model.store
        emb_of = {id(t[0]): t for t in model.embedding_layout}
        grads = []
        for p in ctx.params:
            if id(p) in emb_of:
                if d_in is None:
                    grads.append(None)
                    continue
                _, vocab, xcol, ocol = emb_of[id(p)]
                ids = plan.X[:, xcol].long().clamp_(0, vocab - 1)
                g = torch.zeros_like(p)
                g.index_add_(0, ids, d_in[:, ocol:ocol + model.emb_dim].float())
                grads.append(g)
            elif getattr(p, "_mm_kind", "") == "dense":
                grads.append(st.grad_view(p).clone())
            else:
                grads.append(None)
        return (None, None, None, *grads)


class BaseModel(nn.Module):
    def __init__(self, linear_feature_columns, dnn_feature_columns, init_std=0.0001, device="cpu", gpus=None,
                 config=None):
        super().__init__()
        if config is None:
            raise ValueError("config is required")
        self.dnn_feature_columns = dnn_feature_columns
        self.config = config
        self.data_config = config["data_config"]
        self.model_config = config["model_config"]
        self.optim_config = config["optim_config"]
        self.training_config = config["training_config"]
        self.b200_config = config.get("b200_config", {})
        self.save_layer_output = False
        self.use_cka_loss = self.model_config.get("use_cka_loss", False)
        if self.use_cka_loss:
            raise NotImplementedError("use_cka_loss: utils.CKA does not exist in the reference either (SURVEY Q14)")
        self.device = device
        self.device_obj = torch.device(device)
        self.gpus = gpus
        if gpus and str(gpus[0]) not in str(device):
            raise ValueError("`gpus[0]` should be the same gpu with `device`")

        self.task_name = self.model_config.get("task_name", "mtl")
        self.task_names = self.model_config.get("task_names", ["ctr", "ctcvr"])
        self.task_types = self.model_config.get("task_types", ["binary", "binary"])
        self.num_domains = self.data_config.get("num_domains", 1)
        if self.task_name == "msl":
            self.num_tasks = self.num_domains
        elif self.task_name == "mtmsl":
            self.num_tasks = len(self.data_config["label_columns"])
        else:
            self.num_tasks = len(self.task_names)
        if self.num_tasks <= 1:
            raise ValueError("num_tasks must be greater than 1!")
        if len(dnn_feature_columns) == 0:
            raise ValueError("dnn_feature_columns is null!")
        if len(self.task_types) != self.num_tasks:
            raise ValueError("num_tasks must be equal to the length of task_types")
        for task_type in self.task_types:
            if task_type not in ["binary", "regression"]:
                raise ValueError("task must be binary or regression, {} is illegal".format(task_type))
        # L2 regularisation (basemodel.py:115-130, :514-540).  l2_reg_dnn is fused into the step (the model classes
        # register the same weight lists as the reference's); l2_reg_linear has nothing to act on (no linear columns);
        # l2_reg_embedding > 0 would put a dense gradient on every table row every step and is not implemented.  An
        # absent key counts as 0 here (the reference defaults the linear / embedding terms to 1e-5; every shipped
        # config sets all three to 0).
        if self.model_config.get("l2_reg_embedding", 0) not in (0, 0.0):
            raise NotImplementedError("l2_reg_embedding > 0 is not implemented in the fused step")
        self.regularization_weight = []
        self._regularized_modules = []

        self.feature_index = build_input_features(list(linear_feature_columns) + list(dnn_feature_columns))
        # row-sharded tables (BASELINE config 5): this process holds rows rank::world of every table
        self.shard = None
        sh = self.b200_config.get("shard_tables")
        if sh:
            from ..parallel import ShardContext
            self.shard = ShardContext(int(sh["rank"]), int(sh["world"]))
            # "owner_serve" (default): ids -> owner -> rows; works with the lazy dense-Adam.  "peer_read": requesters read
            # the shard in place (no row round trip, but the owner must sweep its shard every step).  "auto": peer_read
            # for shards within the peer TLB reach (measured slower than owner_serve + lazy Adam at 2 GPUs on PLE-AE)
            import os
            self.shard.gather_mode = sh.get("gather", os.environ.get("MMLREC_SHARD_GATHER", "owner_serve"))
            if self.shard.gather_mode not in ("auto", "owner_serve", "peer_read"):
                raise ValueError('shard_tables["gather"] must be "auto", "owner_serve" or "peer_read"')
        self.embedding_dict = create_embedding_matrix(dnn_feature_columns, init_std, sparse=False, device="cpu",
                                                      shard_world=self.shard.world if self.shard else 1)
        self.out = PredictionLayer(self.model_config.get("task", "binary"))
        self.init_std = init_std
        self.store: Optional[FlatStore] = None
        self._plans: Dict[int, StepPlan] = {}
        self._steps_on_plan: Dict[int, int] = {}
        self.use_cuda_graph = bool(self.b200_config.get("cuda_graph", True))
        self.precision = self.b200_config.get("precision", "fp32")
        self.dp = None  # set by mmlrec_b200.parallel.attach()
        # The scenario mask the reference's classes and loop are written for but never receive (its loop sets
        # domain_mask = None unconditionally, basemodel.py:265-266).  Opt-in: predictions are multiplied by the
        # sample's domain-mask entry (mmoe.py:101-106) and the per-task BCE is weighted by it (basemodel.py:273-282).
        self.use_domain_mask = bool(self.b200_config.get("domain_mask", False))
        if self.use_domain_mask:
            if self.task_name not in ("msl", "mtmsl"):
                raise ValueError('b200_config["domain_mask"] needs task_name "msl" or "mtmsl"')
            mname = self.model_config.get("model_name", "").lower()
            if mname == "star" and self.model_config.get("dnn_use_bn", False):
                # star.py:50-51: only with dnn_use_bn does the mask route the first layer through DomainBatchNorm
                raise NotImplementedError("STAR with a domain mask AND BatchNorm goes through DomainBatchNorm (not built)")
            if mname in ("esmm", "escm"):
                # esmm.py:46-62 / escm.py:74-96 take the argument and never use it: there is no masked ESMM to reproduce
                raise NotImplementedError(f"{mname}: the reference's forward ignores the domain mask")
        self.optimizer_name: Optional[str] = None
        self.lazy_adam, self.adam_hist, self._steps_since_flush = False, None, 0
        self._user_optimizer = None
        self.hyper_dev: Optional[torch.Tensor] = None
        self.metrics, self.metrics_names = {}, ["loss"]

    # ------------------------------------------------------------------ feature bookkeeping
    def compute_input_dim(self, feature_columns, include_sparse=True, include_dense=True, feature_group=False):
        sparse = [fc for fc in feature_columns if isinstance(fc, (SparseFeat, VarLenSparseFeat))]
        dense = [fc for fc in feature_columns if isinstance(fc, DenseFeat)]
        total = 0
        if include_sparse:
            total += len(sparse) if feature_group else sum(fc.embedding_dim for fc in sparse)
        if include_dense:
            total += sum(fc.dimension for fc in dense)
        return total

    @property
    def embedding_size(self):
        dims = {fc.embedding_dim for fc in self.dnn_feature_columns if isinstance(fc, (SparseFeat, VarLenSparseFeat))}
        if len(dims) > 1:
            raise ValueError("embedding_dim of SparseFeat and VarlenSparseFeat must be same in this model!")
        return list(dims)[0]

    def _index_features(self):
        """Column bookkeeping for K1: (table param, vocab, X column, output column) per sparse field
        in list order, then the X columns of the dense features (model/utils.py:434-446 layout)."""
        sparse = [fc for fc in self.dnn_feature_columns if isinstance(fc, SparseFeat)]
        dense = [fc for fc in self.dnn_feature_columns if isinstance(fc, DenseFeat)]
        if any(isinstance(fc, VarLenSparseFeat) for fc in self.dnn_feature_columns):
            raise NotImplementedError("VarLenSparseFeat is outside the B200 hot path")
        dims = {fc.embedding_dim for fc in sparse}
        if len(dims) > 1:
            raise NotImplementedError("the gather kernel needs one embedding_dim for all sparse fields")
        self.emb_dim = dims.pop() if dims else 4
        if self.emb_dim % 4 != 0:
            raise NotImplementedError("embedding_dim must be a multiple of 4 (128-bit row chunks)")
        self.embedding_layout = []
        for j, fc in enumerate(sparse):
            self.embedding_layout.append((self.embedding_dict[fc.embedding_name].weight, fc.vocabulary_size,
                                          self.feature_index[fc.name][0], j * self.emb_dim))
        self.dense_x_cols = []
        for fc in dense:
            a, e = self.feature_index[fc.name]
            self.dense_x_cols.extend(range(a, e))
        self.input_dim_total = len(sparse) * self.emb_dim + len(self.dense_x_cols)
        self.num_x_cols = max(e for _, e in self.feature_index.values())

    # ------------------------------------------------------------------ regularisation
    def add_regularization_weight(self, weight_list, l1=0.0, l2=0.0):
        """basemodel.py:514-523: a parameter, or an iterable of parameters / (name, parameter) pairs."""
        weight_list = [weight_list] if isinstance(weight_list, nn.Parameter) else list(weight_list)
        if l1:
            raise NotImplementedError("l1 regularisation is not fused (the reference never passes l1 > 0)")
        self.regularization_weight.append((weight_list, l1, l2))

    def regularize(self, modules, l2) -> None:
        """The reference's idiom (e.g. mmoe.py:60-63): every '*weight*' that is not a BatchNorm weight of each module.
        The MODULES are remembered, not the tensors: _finalize re-creates the parameters as views of the flat store."""
        self._regularized_modules.extend((m, l2) for m in modules)

    def _regularized(self):
        """(parameter, l2) pairs in registration order (a parameter registered twice appears twice)."""
        for m, l2 in self._regularized_modules:
            for name, prm in m.named_parameters():
                if "weight" in name and "bn" not in name:
                    yield prm, l2
        for weight_list, _, l2 in self.regularization_weight:
            for w in weight_list:
                yield (w[1] if isinstance(w, tuple) else w), l2

    def get_regularization_loss(self) -> torch.Tensor:
        """basemodel.py:524-540, as torch ops on the parameters (the differentiable-forward / external-optimizer path;
        the fused step computes the same value and its gradient in mmlrec_l2_regularize)."""
        total = torch.zeros((1,), device=self.device_obj)
        for prm, l2 in self._regularized():
            if l2 > 0:
                total = total + torch.sum(l2 * torch.square(prm))
        return total

    # ------------------------------------------------------------------ program construction
    def build_graph(self, b: Builder) -> None:  # pragma: no cover - abstract
        raise NotImplementedError

    def _finalize(self) -> None:
        """End of construction (the reference's ``self.to(device)``): lay the parameters out in the
        order the step program consumes them and move everything into the flat device store."""
        self._index_features()
        if self.device_obj.type != "cuda":
            return  # parameters stay ordinary CPU tensors; any compute call raises (no CPU path)
        dry = Builder(2, self.device_obj, None, dry=True, precision=self.precision)
        self.build_graph(dry)
        emb_params = [t[0] for t in self.embedding_layout]
        self.store = FlatStore(self, dry.param_order, emb_params, self.device_obj, want_bf16=self.precision == "bf16",
                               ordered_buffers=dry.buffer_order, aux_floats=dry.aux_floats + 64,
                               emb_alloc=self.shard.alloc_emb if self.shard else None)
        self._index_features()  # re-read the re-pointed table parameters
        for prm, l2 in self._regularized():
            if l2:
                if getattr(prm, "_mm_kind", "") != "dense":
                    raise NotImplementedError("L2 regularisation of an embedding table is not implemented in the fused step")
                self.store.set_l2(prm, l2)
        if self.shard is not None and self.shard.gather_mode == "auto":
            limit = int(self.b200_config.get("peer_read_max_shard_bytes", 64 << 20))
            self.shard.gather_mode = "peer_read" if 4 * self.store.n_emb <= limit else "owner_serve"

    def hyper_host_step(self) -> int:
        """Optimizer step count read back from the device clock (synchronises; evaluation paths only)."""
        if self.hyper_dev is None:
            return 0
        return int(self.hyper_dev[:4].view(torch.int32).item())

    def _require_cuda(self):
        if self.store is None:
            raise RuntimeError("mmlrec_b200 models compute only on a CUDA device through libmmlrec_b200.so; "
                               f"this model was built for device={self.device!r} (there is no CPU fallback)")
        L.load()

    def plan(self, B: int) -> StepPlan:
        self._require_cuda()
        p = self._plans.get(B)
        if p is None:
            with torch.cuda.device(self.device_obj):
                p = StepPlan(self, B)
            self._plans[B] = p
            self._steps_on_plan[B] = 0
        return p

    # ------------------------------------------------------------------ compile
    def compile(self, optimizer, loss=None, metrics=None):
        self.metrics_names = ["loss"]
        self._user_optimizer = None
        if isinstance(optimizer, torch.optim.Optimizer):
            # an optimizer INSTANCE (basemodel.py:557-567 accepts one): its update cannot be fused, so training goes
            # through the differentiable forward + autograd + optimizer.step() -- the reference's own step body
            self._user_optimizer = optimizer
            optimizer = "sgd"   # (fused-state bookkeeping below; the fused update is not used)
        elif not isinstance(optimizer, str):
            raise NotImplementedError("optimizer must be a name or a torch.optim.Optimizer over model.parameters()")
        if optimizer not in ("sgd", "adam", "adagrad", "rmsprop"):
            raise NotImplementedError
        losses = [loss] * self.num_tasks if isinstance(loss, str) else list(loss or [])
        for t, (lname, ttype) in enumerate(zip(losses, self.task_types)):
            ok = (lname == "binary_crossentropy" and ttype == "binary") or (lname == "mse" and ttype == "regression")
            if not ok:
                raise NotImplementedError(f"loss {lname!r} with task type {ttype!r} is not fused")
        self.loss_names = losses
        self.optimizer_name = optimizer
        self.metrics = self._get_metrics(metrics)
        if self.store is not None:
            lr = self.optim_config.get("lr", 1e-3)
            self.store.ensure_optimizer_state(optimizer)
            self.store.reset_optimizer_state()
            h = L.make_hyper(optimizer, lr)
            self.hyper_dev = torch.frombuffer(bytearray(bytes(h)), dtype=torch.uint8).to(self.device_obj)
            self._plans.clear()
            # exact lazy dense-Adam on the tables (csrc/emb_update.cu): rows catch up on the zero-gradient steps they
            # missed when they are next read, instead of a sweep over every table every step.  Row-sharded tables: the
            # owner catches up the rows named by the request keys before it serves them (the peer-read forward, where
            # peers read the shard directly, keeps the per-shard sweep).
            self.lazy_adam = (optimizer == "adam" and bool(self.b200_config.get("lazy_adam", True)) and self.store.n_emb > 0
                              and (self.shard is None or self.shard.gather_mode == "owner_serve"))
            self.adam_hist_cap = 1 << 14
            self.adam_hist = (torch.zeros(2 * self.adam_hist_cap, dtype=torch.float32, device=self.device_obj)
                              if self.lazy_adam else None)
            self._steps_since_flush = 0
        self.optim = self._user_optimizer or _FusedOptimizerHandle(optimizer, self.optim_config.get("lr", 1e-3))

    def _get_metrics(self, metrics):
        """Metric table with the reference's names (basemodel.py:616-635).  The functions are the device-side
        equivalents of the sklearn calls (mmlrec_b200/metrics.py): they take CUDA tensors and return python floats."""
        from .. import metrics as M
        table = {}
        for m in metrics or []:
            if m in ("binary_crossentropy", "logloss"):
                table[m] = M.log_loss
            if m == "auc":
                table[m] = M.roc_auc_score
            if m == "mse":
                table[m] = M.mean_squared_error
            if m in ("accuracy", "acc"):
                table[m] = M.accuracy_score
            self.metrics_names.append(m)
        return table

    # ------------------------------------------------------------------ one step
    def train_on_batch(self, X, y, domain_mask=None) -> torch.Tensor:
        """One optimizer step on a batch (host or device tensors / arrays).  Returns the device
        tensor ``[T+1]`` of per-task BCE sums and their total (no host synchronisation).  ``domain_mask``
        ``[B, num_domains]`` of 0 / 1 is required (and only used) with ``b200_config["domain_mask"]``."""
        if self.hyper_dev is None:
            raise RuntimeError("call compile() before training")
        X = torch.as_tensor(X)
        y = torch.as_tensor(y)
        p = self.plan(X.shape[0])
        if self.use_domain_mask:
            if domain_mask is None:
                raise ValueError('this model was built with b200_config["domain_mask"]: pass domain_mask')
            p.heads.mask.copy_(torch.as_tensor(domain_mask).reshape(X.shape[0], -1), non_blocking=True)
        if X.device.type == "cpu" and X.is_pinned() and y.device.type == "cpu" and y.is_pinned() \
                and X.dtype == torch.float32 and y.dtype == torch.float32:
            self._stage_pinned(p, X, y.reshape(X.shape[0], -1))
        else:
            p.X.copy_(X, non_blocking=True)
            p.y.copy_(y.reshape(X.shape[0], -1), non_blocking=True)
        self._run_train(p)
        return p.loss

    def _stage_pinned(self, p: StepPlan, X: torch.Tensor, y: torch.Tensor) -> None:
        """Pinned host batch -> the step program's inputs through a two-deep device staging ring: the host-to-device
        copy runs on a copy stream, i.e. beside the PREVIOUS step when the caller does not synchronise in between
        (same scheme as fit()'s host-resident path); the main stream only does a device-to-device copy."""
        ring = getattr(p, "_pin_ring", None)
        if ring is None:
            dev = self.device_obj
            ring = p._pin_ring = {"k": 0, "stream": torch.cuda.Stream(device=dev),
                                  "buf": [(torch.empty_like(p.X), torch.empty_like(p.y)) for _ in range(2)],
                                  "copied": [torch.cuda.Event(), torch.cuda.Event()],
                                  "consumed": [torch.cuda.Event(), torch.cuda.Event()]}
        k = ring["k"] & 1
        ring["k"] += 1
        dx, dy = ring["buf"][k]
        main = torch.cuda.current_stream()
        with torch.cuda.stream(ring["stream"]):
            ring["stream"].wait_event(ring["consumed"][k])   # the step that last read this pair has moved it on
            dx.copy_(X, non_blocking=True)
            dy.copy_(y, non_blocking=True)
            ring["copied"][k].record()
        main.wait_event(ring["copied"][k])
        p.X.copy_(dx, non_blocking=True)
        p.y.copy_(dy, non_blocking=True)
        ring["consumed"][k].record(main)

    def check_ids(self) -> None:
        """Raise IndexError if any batch so far carried an id outside its table (the reference's nn.Embedding raises at
        the offending batch; the kernels clamp the id, set a device flag and go on -- read here, at host sync points)."""
        for p in self._plans.values():
            if p.gather.F_s and int(p.gather.oob.item()) != 0:
                p.gather.oob.zero_()
                raise IndexError("index out of range in self: a sparse feature id lies outside [0, vocabulary_size) "
                                 "(ids are clamped on the device; the batch that carried it has already been processed)")
        if self.shard is not None:
            self.shard.check()   # a peer-memory barrier that gave up waiting means a rank is gone: stop, do not train on

    def sync_parameters(self) -> None:
        """After parameters were changed from outside the fused step (an external optimizer, manual edits): refresh
        the bf16 shadow of the dense parameters the tensor-core GEMMs read."""
        if self.store is not None:
            self.store.refresh_bf16()

    def flush_tables(self) -> None:
        """Lazy dense-Adam: bring every table row up to the current optimizer step (no-op otherwise).  Runs before the
        tables are read outside a training step (eval forward, predict, state_dict, deepcopy)."""
        if not self.lazy_adam or self._steps_since_flush == 0:
            return
        st = self.store
        with torch.cuda.device(self.device_obj):
            L.check(L.load().mmlrec_emb_adam_flush(st.emb.data_ptr(), st.emb_s1.data_ptr(), st.emb_s2.data_ptr(),
                                                   st.row_touch.data_ptr(), st.n_emb // self.emb_dim, self.emb_dim,
                                                   self.hyper_dev.data_ptr(), self.adam_hist.data_ptr(), self.adam_hist_cap,
                                                   torch.cuda.current_stream().cuda_stream), "emb_adam_flush")
        self._steps_since_flush = 0

    def state_dict(self, *args, **kwargs):
        self.flush_tables()
        return super().state_dict(*args, **kwargs)

    def _autograd_step(self, p: StepPlan) -> None:
        """The reference's step body (basemodel.py:262-313) with a user-supplied optimizer: forward -> sum of the
        per-task losses -> backward -> optimizer.step(), on the fused forward / backward kernels."""
        import torch.nn.functional as F
        fns = {"binary_crossentropy": F.binary_cross_entropy, "mse": F.mse_loss, "mae": F.l1_loss}
        with torch.enable_grad():
            pred = _FusedForward.apply(self, p, p.X.clone(), *[q for q in self.parameters()])
            self._user_optimizer.zero_grad()
            per_task = [fns[self.loss_names[t]](pred[:, t], p.y[:, t], reduction="sum") for t in range(self.num_tasks)]
            total = sum(per_task)
            total.backward()
        self._user_optimizer.step()
        self.sync_parameters()
        p.loss.copy_(torch.stack([v.detach() for v in per_task] + [total.detach()]))

    def _run_train(self, p: StepPlan) -> None:
        if self._user_optimizer is not None:
            self._autograd_step(p)
            return
        n = self._steps_on_plan[p.B]
        if p.B == 1 and self.model_config.get("dnn_use_bn", False):
            raise ValueError("Expected more than 1 value per channel when training")   # what nn.BatchNorm1d raises
        if self.lazy_adam:
            if self._steps_since_flush >= self.adam_hist_cap - 2:   # the history ring is about to wrap
                self.flush_tables()
            self._steps_since_flush += 1
        self.store.live_slices = 1 if self.dp is not None else p.grad_slices
        if not self.use_cuda_graph:
            p.train_step()
        elif p.graph is not None:
            p.replay()
        elif n == 0:
            p.train_step()          # eager warm-up (also the step that surfaces launch errors)
        else:
            p.capture()
            p.replay()
        self._steps_on_plan[p.B] = n + 1

    def forward(self, X, domain_mask=None):
        """Probabilities ``[B, T]`` (sigmoid applied for 'binary' heads), like the reference's ``forward``
        (model/mmoe.py:65).  With autograd enabled the result is differentiable with respect to every parameter
        (``loss.backward()`` fills ``.grad``; any ``torch.optim`` optimizer over ``model.parameters()`` then works --
        call ``model.sync_parameters()`` after ``optimizer.step()`` in bf16 mode).  The fused training step of
        ``fit`` / ``train_on_batch`` does not come through here."""
        X = torch.as_tensor(X)
        self.flush_tables()
        p = self.plan(X.shape[0])
        if p.heads.mask is not None:
            p.heads.mask.fill_(1.0)   # the mask is applied to the returned tensor below
        if torch.is_grad_enabled() and any(q.requires_grad for q in self.parameters()) and p.heads.escm is None:
            out = _FusedForward.apply(self, p, X.to(self.device_obj), *[q for q in self.parameters()])
        else:
            p.X.copy_(X, non_blocking=True)
            p.forward(training=self.training)
            out = p.pred.clone()
        if domain_mask is not None:
            dm = torch.as_tensor(domain_mask).to(out)
            if self.task_name == "msl":
                out = out * dm
            elif self.task_name == "mtmsl":
                out = out * dm[:, [i % self.num_domains for i in range(self.num_tasks)]]
        return out

    # ------------------------------------------------------------------ data plumbing
    def _stack_inputs(self, x) -> np.ndarray:
        if isinstance(x, dict):
            x = [x[name] for name in self.feature_index]
        cols = []
        for a in x:
            a = np.asarray(a.values if hasattr(a, "values") else a)
            cols.append(a.reshape(len(a), -1))
        return np.concatenate(cols, axis=-1).astype(np.float32)

    # ------------------------------------------------------------------ fit / evaluate / predict
    def fit(self, x=None, y=None, batch_size=None, epochs=1, initial_epoch=0, validation_split=0.,
            validation_data=None, shuffle=True):
        self._require_cuda()
        from torch.utils.data import DataLoader, TensorDataset
        dm, dm_val = None, None
        if self.use_domain_mask:
            col, vals = self.data_config.get("mask_column", ""), self.data_config.get("mask_values", [])
            if col == "" or not isinstance(x, dict):
                raise ValueError('b200_config["domain_mask"]: fit() needs dict inputs and data_config["mask_column"]')
            dm = get_mask(list(np.asarray(x[col])), vals, self.num_domains).float()
            if validation_data:
                dm_val = get_mask(list(np.asarray(validation_data[0][col])), vals, self.num_domains).float()
        X = self._stack_inputs(x)
        y = np.asarray(y, dtype=np.float32).reshape(len(X), self.num_tasks)
        val = None
        if validation_data:
            if len(validation_data) not in (2, 3):
                raise ValueError("validation_data must be (x_val, y_val) or (x_val, y_val, sample_weights)")
            val = (self._stack_inputs(validation_data[0]),
                   np.asarray(validation_data[1], dtype=np.float32).reshape(-1, self.num_tasks))
        elif validation_split and 0. < validation_split < 1.:
            cut = int(len(X) * (1. - validation_split))
            X, y, val = X[:cut], y[:cut], (X[cut:], y[cut:])
            if dm is not None:
                dm, dm_val = dm[:cut], dm[cut:]
        batch_size = 256 if batch_size is None else batch_size
        dev = self.device_obj
        # input pipeline.  A dataset that fits is kept in HBM and batches are index_select'ed on the
        # device; a larger one stays in PINNED host memory and every batch is gathered on the host into one of two pinned
        # staging buffers and copied on a side stream while the previous step is still running (double buffering).
        resident = (X.nbytes + y.nbytes) <= int(self.b200_config.get("max_resident_bytes", 16 << 30))
        if resident:
            Xd, yd = torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev)
        else:
            Xh, yh = torch.from_numpy(X).pin_memory(), torch.from_numpy(y).pin_memory()
            stage = [(torch.empty(batch_size, X.shape[1]).pin_memory(), torch.empty(batch_size, y.shape[1]).pin_memory())
                     for _ in range(2)]
            dstage = [(torch.empty(batch_size, X.shape[1], device=dev), torch.empty(batch_size, y.shape[1], device=dev))
                      for _ in range(2)]
            copy_stream = torch.cuda.Stream(device=dev)
            copied, consumed = [torch.cuda.Event(), torch.cuda.Event()], [torch.cuda.Event(), torch.cuda.Event()]
        if dm is not None:
            dm = dm.to(dev)
        n = len(X)
        steps = (n - 1) // batch_size + 1
        # the reference's DataLoader(shuffle=True) consumes the global torch RNG; iterating a DataLoader
        # over the row indices reproduces its batch order exactly
        loader = DataLoader(TensorDataset(torch.arange(n)), shuffle=shuffle, batch_size=batch_size)
        print("Train on {0} samples, validate on {1} samples, {2} steps per epoch".format(
            n, 0 if val is None else len(val[1]), steps))
        self.train()
        best_auc, stall, best_model = 0, 0, None
        for epoch in range(initial_epoch, epochs):
            t0 = time.time()
            preds, idxs, losses, ys = [], [], [], []
            for k, (idx,) in enumerate(loader):
                p = self.plan(len(idx))
                if resident:
                    idx_d = idx.to(dev, non_blocking=True)
                    torch.index_select(Xd, 0, idx_d, out=p.X)
                    torch.index_select(yd, 0, idx_d, out=p.y)
                else:
                    (sx, sy), (dx, dy), nb = stage[k & 1], dstage[k & 1], len(idx)
                    copied[k & 1].synchronize()   # host: the H2D copy that last read this pinned pair (batch k-2) is done
                    torch.index_select(Xh, 0, idx, out=sx[:nb])   # host gather -- the GPU is still busy with step k-1
                    torch.index_select(yh, 0, idx, out=sy[:nb])
                    with torch.cuda.stream(copy_stream):          # H2D beside the running step
                        copy_stream.wait_event(consumed[k & 1])   # (batch k-2 has been moved out of this device pair)
                        dx[:nb].copy_(sx[:nb], non_blocking=True)
                        dy[:nb].copy_(sy[:nb], non_blocking=True)
                        copied[k & 1].record()
                    main = torch.cuda.current_stream()
                    main.wait_event(copied[k & 1])
                    p.X.copy_(dx[:nb], non_blocking=True)         # device -> device into the step program's input
                    p.y.copy_(dy[:nb], non_blocking=True)
                    consumed[k & 1].record(main)
                if dm is not None:
                    torch.index_select(dm, 0, idx.to(dev, non_blocking=True), out=p.heads.mask)
                self._run_train(p)
                preds.append(p.pred.clone())
                ys.append(p.y.clone())
                # (the epoch log is the TOTAL loss: task losses + the L2 term, basemodel.py:303-308, :333)
                losses.append(p.loss[self.num_tasks] + p.reg_loss[0])
                idxs.append(idx)
            total = float(torch.stack(losses).sum().item())
            self.check_ids()
            logs = {"loss": total / n, "cka_loss": 0.0}
            if self.metrics:
                sums = {k: 0.0 for k in self.metrics}
                for yt, yp in zip(ys, preds):
                    for name, fn in self.metrics.items():
                        try:
                            sums[name] += self._train_metric(fn, yt, yp)
                        except ValueError:
                            pass  # a batch with a single class has no AUC
                for k, v in sums.items():
                    logs[k] = v / steps
            if val is not None:
                res = self.evaluate(val[0], val[1], batch_size, dm_val)
                print(res)
                if res.get("auc", 0) > best_auc:
                    best_auc, best_model, stall = res["auc"], copy.deepcopy(self), 0
                else:
                    stall += 1
                for k, v in res.items():
                    logs["val_" + k] = v
                self.train()
            print("Epoch {0}/{1}".format(epoch + 1, epochs))
            line = "{0}s - loss: {1: .4f} - cka_loss: {2: .4f}".format(int(time.time() - t0), logs["loss"], 0.0)
            for name in self.metrics:
                line += " - {0}: {1: .4f}".format(name, logs.get(name, float("nan")))
            if val is not None:
                for name in self.metrics:
                    line += " - val_{0}: {1: .4f}".format(name, logs.get("val_" + name, float("nan")))
            print(line)
            if stall >= self.optim_config.get("early_stop", 3):
                break
        return best_model if best_model is not None else self

    def _train_metric(self, fn, y_true, y_pred):
        """The reference's metric inputs per task type (basemodel.py:316-331, :383-392); tensors on the device."""
        if y_pred.shape[1] == 3 and self.model_config.get("model_name", "").lower() == "escm":
            y_pred = y_pred[:, [0, 2]]   # ctr and ctcvr (basemodel.py:328-329, :438-441)
        if self.task_name == "msl":
            return fn(y_true[:, 0], y_pred.sum(dim=-1))
        if self.task_name == "mtmsl":
            d = self.num_domains
            return fn(y_true[:, [0, d]], torch.stack([y_pred[:, :d].sum(-1), y_pred[:, d:].sum(-1)], dim=-1))
        return fn(y_true, y_pred)

    def evaluate(self, x, y, batch_size=256, domain_mask=None):
        pred = self.predict_device(x, batch_size, domain_mask)
        yt = torch.as_tensor(np.asarray(y, dtype=np.float32)).reshape(pred.shape[0], -1).to(self.device_obj)
        return {name: self._train_metric(fn, yt, pred) for name, fn in self.metrics.items()}

    def predict_device(self, x, batch_size=256, domain_mask=None) -> torch.Tensor:
        """``predict`` without the trip to the host: float32 ``[N, T]`` on the device."""
        self._require_cuda()
        self.flush_tables()
        was_training = self.training
        self.eval()
        arr = x.astype(np.float32) if isinstance(x, np.ndarray) and x.ndim == 2 else self._stack_inputs(x)
        X = torch.from_numpy(np.ascontiguousarray(arr)).to(self.device_obj)
        # the reference iterates a DataLoader here (basemodel.py:425-426); creating its iterator draws one int64 from
        # the global torch RNG (the loader's base seed) even without shuffling.  Drawing it too keeps the RNG stream
        # -- hence the shuffle order of the NEXT training epoch -- identical to the reference's
        torch.empty((), dtype=torch.int64).random_()
        out = []
        for a in range(0, len(X), batch_size):
            xb = X[a:a + batch_size]
            p = self.plan(len(xb))
            if p.heads.mask is not None:
                p.heads.mask.fill_(1.0)
            p.X.copy_(xb)
            p.forward(training=False)
            out.append(p.pred.clone())
        self.train(was_training)
        res = torch.cat(out)
        if res.shape[1] == 3 and self.model_config.get("model_name", "").lower() == "escm":
            res = res[:, [0, 2]].contiguous()   # basemodel.py:438-441: predict() returns ctr and ctcvr
        if domain_mask is not None:
            dm = torch.as_tensor(domain_mask).to(res)
            if self.task_name == "msl":
                res = res * dm
            elif self.task_name == "mtmsl":
                res = res * dm[:, [i % self.num_domains for i in range(self.num_tasks)]]
        self.check_ids()
        return res

    def predict(self, x, batch_size=256, domain_mask=None):
        return self.predict_device(x, batch_size, domain_mask).cpu().numpy().astype("float64")

    def update_save(self, value=True):
        self.save_layer_output = value

    # ------------------------------------------------------------------ copies / checkpoints
    def __deepcopy__(self, memo):
        """``fit`` keeps the best epoch as ``deepcopy(model)`` (basemodel.py:344).  Parameters are
        views of flat buffers, so the copy is rebuilt through the constructor and the flat buffers
        are cloned wholesale."""
        self.flush_tables()
        with torch.random.fork_rng(devices=[]):   # the constructor draws (and discards) initial weights: keep the caller's RNG stream
            clone = type(self)(self.dnn_feature_columns, init_std=self.init_std, device=self.device, gpus=self.gpus,
                               config=copy.deepcopy(self.config))
        if self.store is not None:
            for name in ("dense", "emb", "stats", "counts"):
                getattr(clone.store, name).copy_(getattr(self.store, name))
            clone.store.refresh_bf16()
        else:
            clone.load_state_dict(self.state_dict())
        if self.optimizer_name is not None:
            clone.compile(self.optimizer_name, self.loss_names, [m for m in self.metrics_names if m != "loss"])
        clone.train(self.training)
        return clone

    _CKPT_BUFFERS = ("dense", "emb", "stats", "counts", "dense_s1", "dense_s2", "emb_s1", "emb_s2", "row_touch")

    def save_checkpoint(self, path: str) -> str:
        """Everything a resumed run needs to continue bit for bit: the flat parameter / BatchNorm buffers, the optimizer
        state of the dense parameters and of the tables (incl. the lazy dense-Adam row stamps and history ring) and the
        device-side optimizer clock.  The reference keeps no checkpoint of its own (``deepcopy`` of the best
        epoch only); the tensors are stored under their flat-store names, ``state_dict()`` rides along for interchange
        with a reference model.  Row-sharded tables: every rank calls this and writes ITS shard to ``<path>.rank<r>``."""
        self._require_cuda()
        if self.hyper_dev is None:
            raise RuntimeError("call compile() before save_checkpoint()")
        torch.cuda.synchronize(self.device_obj)
        st = self.store
        blob = {"format": 1, "model": type(self).__name__, "optimizer": self.optimizer_name,
                "precision": self.precision, "hyper": self.hyper_dev.cpu(),
                "steps_since_flush": self._steps_since_flush,
                "adam_hist": None if self.adam_hist is None else self.adam_hist.cpu(),
                "store": {k: getattr(st, k).cpu() for k in self._CKPT_BUFFERS if getattr(st, k, None) is not None},
                "shard": None if self.shard is None else (self.shard.rank, self.shard.world),
                "state_dict": {k: v.cpu() for k, v in super().state_dict().items()}}
        if self.shard is not None:
            path = f"{path}.rank{self.shard.rank}"
        torch.save(blob, path)
        return path

    def load_checkpoint(self, path: str) -> None:
        """Restore ``save_checkpoint``'s state into a model built with the same columns / config and compiled with the
        same optimizer; training then continues exactly where the saved run stopped."""
        self._require_cuda()
        if self.shard is not None:
            path = f"{path}.rank{self.shard.rank}"
        blob = torch.load(path, map_location="cpu", weights_only=False)
        if blob.get("format") != 1 or blob["model"] != type(self).__name__:
            raise ValueError(f"{path}: not a checkpoint of a {type(self).__name__}")
        if self.hyper_dev is None or blob["optimizer"] != self.optimizer_name:
            raise RuntimeError(f"compile() the model with optimizer {blob['optimizer']!r} before load_checkpoint()")
        shard = None if self.shard is None else (self.shard.rank, self.shard.world)
        if blob["shard"] != shard:
            raise ValueError(f"{path}: saved with table shard {blob['shard']}, this model has {shard}")
        st = self.store
        for k, v in blob["store"].items():
            dst = getattr(st, k, None)
            if dst is None or dst.shape != v.shape:
                raise ValueError(f"{path}: buffer {k!r} does not match this model")
            dst.copy_(v)
        self.hyper_dev.copy_(blob["hyper"])
        if self.adam_hist is not None and blob["adam_hist"] is not None:
            self.adam_hist.copy_(blob["adam_hist"])
        self._steps_since_flush = int(blob["steps_since_flush"])
        st.refresh_bf16()
        torch.cuda.synchronize(self.device_obj)

    def load_state_dict(self, state_dict, strict=True, assign=False):
        out = super().load_state_dict(state_dict, strict=strict, assign=False)
        if self.store is not None:
            self.store.refresh_bf16()
        return out
