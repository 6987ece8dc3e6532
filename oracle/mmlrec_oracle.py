"""CPU oracle for the MMLRec training hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product package never does (it fails loudly when the CUDA
library is missing instead of falling back to anything here).

What it is: a *functional* restatement, in plain fp32 PyTorch on the CPU, of the reference's
step body (``/root/reference/model/basemodel.py:262-313``): multi-field gather + concat, the
expert / gate / tower networks of twelve model families, sigmoid + BCE(sum), backward and the
``torch.optim`` step.  All arithmetic of the reference lives in the third-party module ``torch``
(reference prose pins "PyTorch 1.11.0", ``README.md:52``; no lock file; operative version in
this image: torch 2.11.0+cu128), so the restatement calls the same ATen ops in the same order.
Parameters are a flat ``dict`` keyed by the reference's own ``state_dict`` names, which is what
lets a reference model's weights be dropped in unchanged.

Parity pin: the reference ships no tests or golden vectors, so the oracle
is pinned against outputs of the reference itself, run in the build container by
``tests/golden/make_golden.py`` (imports ``/root/reference`` read-only) and committed as
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` replays them on every CPU test run.

Reference citations are to ``/root/reference``.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# schema (model/utils.py:328-346 SparseFeat, :388-395 DenseFeat, :407-431 build_input_features)
# --------------------------------------------------------------------------------------
class Column:
    """One input column: sparse (id -> embedding row) or dense (scalar passthrough)."""

    __slots__ = ("name", "kind", "vocab", "dim")

    def __init__(self, name: str, kind: str, vocab: int = 0, dim: int = 1):
        assert kind in ("sparse", "dense")
        self.name, self.kind, self.vocab, self.dim = name, kind, int(vocab), int(dim)


def feature_layout(columns: Sequence[Column]) -> "OrderedDict[str, Tuple[int, int]]":
    """name -> (start, end) column span in X.  model/utils.py:407-431: every SparseFeat takes one
    column, every DenseFeat ``dimension`` columns, first occurrence of a name wins."""
    out: "OrderedDict[str, Tuple[int, int]]" = OrderedDict()
    at = 0
    for c in columns:
        if c.name in out:
            continue
        w = 1 if c.kind == "sparse" else c.dim
        out[c.name] = (at, at + w)
        at += w
    return out


def input_dim(columns: Sequence[Column]) -> int:
    """model/basemodel.py:489-507."""
    return sum(c.dim for c in columns)


# --------------------------------------------------------------------------------------
# gather + concat (model/basemodel.py:461-487, model/utils.py:434-446)
# --------------------------------------------------------------------------------------
def gather_fields(X: Tensor, params: Params, columns: Sequence[Column]) -> Tuple[List[Tensor], List[Tensor]]:
    lay = feature_layout(columns)
    rows, dense = [], []
    for c in columns:
        a, b = lay[c.name]
        if c.kind == "sparse":
            ids = X[:, a:b].long()  # fp32 carrier truncated toward zero, basemodel.py:476
            rows.append(F.embedding(ids, params[f"embedding_dict.{c.name}.weight"]))  # [B,1,D]
        else:
            dense.append(X[:, a:b])
    return rows, dense


def concat_fields(rows: List[Tensor], dense: List[Tensor]) -> Tensor:
    """model/utils.py:434-446: sparse block (field-major) then dense block."""
    parts = []
    if rows:
        parts.append(torch.flatten(torch.cat(rows, dim=-1), start_dim=1))
    if dense:
        parts.append(torch.flatten(torch.cat(dense, dim=-1), start_dim=1))
    if not parts:
        raise NotImplementedError
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=-1)


def gather_concat(X: Tensor, params: Params, columns: Sequence[Column]) -> Tensor:
    return concat_fields(*gather_fields(X, params, columns))


# --------------------------------------------------------------------------------------
# MLP block (model/utils.py:92-161), prediction layer (:225-248), loss (basemodel.py:294-296)
# --------------------------------------------------------------------------------------
def _count_layers(params: Params, prefix: str) -> int:
    n = 0
    while f"{prefix}.linears.{n}.weight" in params:
        n += 1
    return n


def mlp(params: Params, buffers: Params, prefix: str, x: Tensor, use_bn: bool, training: bool,
        activation: str = "relu") -> Tensor:
    """Linear -> [BatchNorm1d] -> activation per layer; dropout p=0 in every shipped config."""
    for i in range(_count_layers(params, prefix)):
        x = F.linear(x, params[f"{prefix}.linears.{i}.weight"], params[f"{prefix}.linears.{i}.bias"])
        if use_bn:
            bn = f"{prefix}.bn.{i}"
            x = F.batch_norm(x, buffers[f"{bn}.running_mean"], buffers[f"{bn}.running_var"],
                             params[f"{bn}.weight"], params[f"{bn}.bias"], training, 0.1, 1e-5)
            if training:
                buffers[f"{bn}.num_batches_tracked"] += 1
        if activation == "relu":
            x = torch.relu(x)
        elif activation == "sigmoid":
            x = torch.sigmoid(x)
        elif activation is not None:
            raise NotImplementedError(activation)
    return x


def predict_head(logit: Tensor, bias: Tensor, task: str) -> Tensor:
    """model/utils.py:242-248 (bias add, sigmoid when binary)."""
    z = logit + bias
    return torch.sigmoid(z) if task == "binary" else z


def bce_sum(pred: Tensor, y: Tensor) -> Tensor:
    """basemodel.py:294-296: sum over tasks of binary_cross_entropy on PROBABILITIES, reduction='sum'
    (ATen clamps each log term at -100)."""
    return sum(F.binary_cross_entropy(pred[:, t], y[:, t], reduction="sum") for t in range(pred.shape[1]))


def loss_sum(pred: Tensor, y: Tensor, loss_names: Optional[Sequence[str]] = None) -> Tensor:
    """basemodel.py:294-296 with the per-task loss functions of ``_get_loss_func_single`` (:595-604):
    'binary_crossentropy' -> F.binary_cross_entropy, 'mse' -> F.mse_loss, 'mae' -> F.l1_loss, all reduction='sum'."""
    if loss_names is None:
        return bce_sum(pred, y)
    fns = {"binary_crossentropy": F.binary_cross_entropy, "mse": F.mse_loss, "mae": F.l1_loss}
    return sum(fns[loss_names[t]](pred[:, t], y[:, t], reduction="sum") for t in range(pred.shape[1]))


def _gate_mix(gate_logits: Tensor, experts: List[Tensor]) -> Tensor:
    """softmax(gate) @ stack(experts)  (mmoe.py:86, ple.py:139)."""
    stack = torch.stack(experts, 1)  # [B,E,H]
    return torch.matmul(gate_logits.softmax(1).unsqueeze(1), stack).squeeze(1)


# --------------------------------------------------------------------------------------
# model families
# --------------------------------------------------------------------------------------
class Spec:
    """The few config values the forward functions need (all read with the reference's defaults)."""

    def __init__(self, config: dict, columns: Sequence[Column]):
        mc, dc = config["model_config"], config["data_config"]
        self.columns = list(columns)
        self.model = mc["model_name"].lower()
        self.task_name = mc.get("task_name", "mtl")
        self.task_types = mc.get("task_types", ["binary", "binary"])
        if self.task_name == "msl":  # basemodel.py:97-102
            self.num_tasks = dc.get("num_domains", 1)
        elif self.task_name == "mtmsl":
            self.num_tasks = len(dc["label_columns"])
        else:
            self.num_tasks = len(mc.get("task_names", ["ctr", "ctcvr"]))
        self.task = mc.get("task", "binary")
        self.use_bn = mc.get("dnn_use_bn", False)
        self.act = mc.get("dnn_activation", "relu")
        self.num_experts = mc.get("num_experts", 4)
        self.shared_n = mc.get("shared_expert_num", 1)
        self.specific_n = mc.get("specific_expert_num", 3)
        self.levels = mc.get("num_levels", 1)
        self.gate_units = mc.get("gate_dnn_hidden_units", [64])
        self.expert_units = mc.get("expert_dnn_hidden_units", [256, 128])
        self.tower_units = mc.get("tower_dnn_hidden_units", [64])
        self.dnn_units = mc.get("dnn_hidden_units", [256, 128])
        self.use_shared = mc.get("use_shared", True)
        self.task_weight_units = mc.get("task_weight_hidden_units", [64])
        self.scene_feature = dc.get("scene_feature", "")
        self.user_sf = dc.get("user_sf", "")
        self.item_sf = dc.get("item_sf", "")


def _towers(p: Params, b: Params, s: Spec, tower_in: List[Tensor], training: bool) -> Tensor:
    """tower DNN -> bias-free 1-unit head -> PredictionLayer, per task (mmoe.py:93-108)."""
    outs = []
    for t in range(s.num_tasks):
        h = tower_in[t]
        if len(s.tower_units) > 0:
            h = mlp(p, b, f"tower_dnn.{t}", h, s.use_bn, training, s.act)
        logit = F.linear(h, p[f"tower_dnn_final_layer.{t}.weight"])
        outs.append(predict_head(logit, p[f"out.{t}.bias"], s.task_types[t]))
    return torch.cat(outs, -1)


def forward_mmoe(p: Params, b: Params, s: Spec, X: Tensor, training: bool) -> Tensor:
    """model/mmoe.py:65-108."""
    x = gather_concat(X, p, s.columns)
    experts = [mlp(p, b, f"expert_dnn.{e}", x, s.use_bn, training, s.act) for e in range(s.num_experts)]
    mixed = []
    for t in range(s.num_tasks):
        g = mlp(p, b, f"gate_dnn.{t}", x, s.use_bn, training, s.act) if len(s.gate_units) > 0 else x
        mixed.append(_gate_mix(F.linear(g, p[f"gate_dnn_final_layer.{t}.weight"]), experts))
    return _towers(p, b, s, mixed, training)


def forward_ple(p: Params, b: Params, s: Spec, X: Tensor, training: bool) -> Tensor:
    """model/ple.py:107-198.  Only ``shared_expert_num`` of the allocated shared experts are used
    (ple.py:120); the top level's shared-gate output never reaches a tower."""
    x = gather_concat(X, p, s.columns)
    inputs = [x] * (s.num_tasks + 1)
    T, S, Sh = s.num_tasks, s.specific_n, s.shared_n
    for lv in range(s.levels):
        spec = [mlp(p, b, f"specific_experts.{lv}.{t}.{j}", inputs[t], s.use_bn, training, s.act)
                for t in range(T) for j in range(S)]
        shared = [mlp(p, b, f"shared_experts.{lv}.0.{k}", inputs[-1], s.use_bn, training, s.act) for k in range(Sh)]
        outs = []
        for t in range(T):
            g = inputs[t]
            if len(s.gate_units) > 0:
                g = mlp(p, b, f"specific_gate_dnn.{lv}.{t}.0", g, s.use_bn, training, s.act)
            g = F.linear(g, p[f"specific_gate_dnn_final_layer.{lv}.{t}.weight"])
            outs.append(_gate_mix(g, spec[t * S:(t + 1) * S] + shared))
        g = inputs[-1]
        if len(s.gate_units) > 0:
            g = mlp(p, b, f"shared_gate_dnn.{lv}", g, s.use_bn, training, s.act)
        g = F.linear(g, p[f"shared_gate_dnn_final_layer.{lv}.weight"])
        outs.append(_gate_mix(g, spec + shared))
        inputs = outs
    return _towers(p, b, s, inputs[:T], training)


def forward_sharedbottom(p: Params, b: Params, s: Spec, X: Tensor, training: bool) -> Tensor:
    """model/sharedbottom.py:52-86."""
    x = gather_concat(X, p, s.columns)
    h = mlp(p, b, "bottom_dnn", x, s.use_bn, training, s.act)
    return _towers(p, b, s, [h] * s.num_tasks, training)


def forward_esmm(p: Params, b: Params, s: Spec, X: Tensor, training: bool) -> Tensor:
    """model/esmm.py:45-62: ONE shared PredictionLayer bias (``out.bias``) serves both heads."""
    x = gather_concat(X, p, s.columns)
    ctr = F.linear(mlp(p, b, "ctr_dnn", x, s.use_bn, training, s.act), p["ctr_dnn_final_layer.weight"])
    cvr = F.linear(mlp(p, b, "cvr_dnn", x, s.use_bn, training, s.act), p["cvr_dnn_final_layer.weight"])
    p_ctr = predict_head(ctr, p["out.bias"], s.task)
    p_cvr = predict_head(cvr, p["out.bias"], s.task)
    return torch.cat([p_ctr, p_ctr * p_cvr], -1)


def forward_escm(p: Params, b: Params, s: Spec, X: Tensor, training: bool) -> Tensor:
    """model/escm.py:74-96 (the 'escm' variant): ESMM's towers, output [p_ctr, p_cvr, p_ctr * p_cvr]."""
    x = gather_concat(X, p, s.columns)
    ctr = F.linear(mlp(p, b, "ctr_dnn", x, s.use_bn, training, s.act), p["ctr_dnn_final_layer.weight"])
    cvr = F.linear(mlp(p, b, "cvr_dnn", x, s.use_bn, training, s.act), p["cvr_dnn_final_layer.weight"])
    p_ctr = predict_head(ctr, p["out.bias"], s.task)
    p_cvr = predict_head(cvr, p["out.bias"], s.task)
    return torch.cat([p_ctr, p_cvr, p_ctr * p_cvr], -1)


def escm_loss(pred: Tensor, y: Tensor, counterfactual_w: float = 0.1, global_w: float = 1.0) -> Tensor:
    """model/basemodel.py:284-292 with model/escm.py:98-111 (counterfact_ipw; `ips.stop_gradient = True` is a no-op
    attribute assignment in torch, so the gradient flows through the weights)."""
    loss_0 = F.binary_cross_entropy(pred[:, 0], y[:, 0], reduction="sum")
    loss_1 = F.binary_cross_entropy(pred[:, 1], y[:, 1], reduction="sum")
    loss_2 = F.binary_cross_entropy(pred[:, 2], y[:, 1], reduction="sum")
    ctr_num, o = torch.sum(y[:, 0]), y[:, 0].float()
    ps = torch.maximum(pred[:, 0] * ctr_num.float(), torch.full_like(pred[:, 0], 0.000001))
    ips = torch.clip(torch.reciprocal(ps), min=-15, max=15) * torch.sum(torch.full_like(o, 1).float(), 0)
    loss_1 = torch.mean(loss_1 * ips * o)
    return loss_0 + loss_1 * counterfactual_w + loss_2 * global_w


def star_weight(p: Params, prefix: str, i: int, last: int, which: str) -> Tensor:
    """SharedSpecificLinear keeps per-domain tensors in plain lists (model/utils.py:181-191); only
    the last one is a registered parameter (``<prefix>.specific_<which>``).  The oracle stores the
    unregistered ones under ``<prefix>.specific_weights.<i>`` / ``.specific_biases.<i>`` (frozen constants)."""
    if i == last:
        return p[f"{prefix}.specific_{which}"]
    plural = "weights" if which == "weight" else "biases"
    return p[f"{prefix}.specific_{plural}.{i}"]


def _star_linear(p: Params, prefix: str, x: Tensor, i: int, last: int, use_shared: bool) -> Tensor:
    """model/utils.py:214-218."""
    w, bias = star_weight(p, prefix, i, last, "weight"), star_weight(p, prefix, i, last, "bias")
    if use_shared:
        return torch.matmul(x, w * p[f"{prefix}.shared_weight"]) + bias + p[f"{prefix}.shared_bias"]
    return torch.matmul(x, w) + bias


def forward_star(p: Params, b: Params, s: Spec, X: Tensor, training: bool) -> Tensor:
    """model/star.py:41-64 with domain_mask=None (always, basemodel.py:265-266), so DomainBatchNorm
    is never entered."""
    x = gather_concat(X, p, s.columns)
    T, outs = s.num_tasks, []
    for i in range(T):
        h = x
        for j in range(len(s.dnn_units)):
            h = _star_linear(p, f"linears.{j}", h, i, T - 1, s.use_shared)
            h = torch.relu(h) if s.act == "relu" else torch.sigmoid(h)
        z = _star_linear(p, f"final_layers.{i}", h, i, T - 1, s.use_shared)
        outs.append(predict_head(z, p[f"out.{i}.bias"], s.task_types[i]))
    return torch.cat(outs, -1)


def _gate_nn(p: Params, prefix: str, x: Tensor) -> Tensor:
    """model/pepnet.py:8-32: Linear -> ReLU -> Linear -> Sigmoid, times 2."""
    h = torch.relu(F.linear(x, p[f"{prefix}.gate.0.weight"], p[f"{prefix}.gate.0.bias"]))
    return torch.sigmoid(F.linear(h, p[f"{prefix}.gate.2.weight"], p[f"{prefix}.gate.2.bias"])) * 2


def forward_pepnet(p: Params, b: Params, s: Spec, X: Tensor, training: bool) -> Tensor:
    """model/pepnet.py:119-156.  ``sparse_embedding_list[scene_index[0]]`` indexes the list of
    sparse fields by the scene feature's X column (identical while sparse columns come first)."""
    rows, dense = gather_fields(X, p, s.columns)
    x = concat_fields(rows, dense)
    lay = feature_layout(s.columns)
    side = [rows[lay[s.scene_feature][0]].squeeze(1).detach()]
    if s.user_sf != "" and s.item_sf != "":
        side += [rows[lay[s.user_sf][0]].squeeze(1).detach(), rows[lay[s.item_sf][0]].squeeze(1).detach()]
    scene, task_side = side[0], torch.cat(side, -1) if len(side) > 1 else side[0]
    x = _gate_nn(p, "feature_gate", torch.cat([x.detach(), scene], -1)) * x
    L, outs = len(s.dnn_units), []
    for t in range(s.num_tasks):
        gate_in = torch.cat([x.detach(), task_side], -1)
        h = x
        for i in range(L + 1):
            h = h * _gate_nn(p, f"ppn.{t}.gate_layers.{i}", gate_in)
            if i < L:
                h = torch.relu(F.linear(h, p[f"ppn.{t}.mlp_layers.{i}.0.weight"], p[f"ppn.{t}.mlp_layers.{i}.0.bias"]))
            else:
                h = F.linear(h, p[f"ppn.{t}.mlp_layers.{i}.weight"], p[f"ppn.{t}.mlp_layers.{i}.bias"])
        outs.append(predict_head(h, p[f"out.{t}.bias"], s.task_types[t]))
    return torch.cat(outs, -1)


def forward_mlp(p: Params, b: Params, s: Spec, X: Tensor, training: bool) -> Tensor:
    """model/mlp.py:36-52: a stack of single-layer DNN blocks (no BatchNorm: the blocks are built with the DNN
    defaults), ONE bias-free final layer; every task's PredictionLayer is applied to the same logit."""
    h = gather_concat(X, p, s.columns)
    for i in range(len(s.dnn_units)):
        h = mlp(p, b, f"mlp_layers.{i}", h, False, training, "relu")
    z = F.linear(h, p["final_layer.weight"])
    outs = []
    for t in range(s.num_tasks):  # PredictionLayer.forward does ``output += self.bias`` IN PLACE (utils.py:243-245) on the
        z = z + p[f"out.{t}.bias"]   # one logit tensor all tasks share: task t sees biases 0..t
        outs.append(torch.sigmoid(z) if s.task_types[t] == "binary" else z)
    return torch.cat(outs, -1)


def forward_cross_stitch(p: Params, b: Params, s: Spec, X: Tensor, training: bool) -> Tensor:
    """model/cross_stitch.py:82-112: shared layer, then per level the T task layers and the cross-stitch unit
    ``cat(outputs) @ W`` (:17-18; W is used as stored, [in, out]), split back per task."""
    x = gather_concat(X, p, s.columns)
    shared = mlp(p, b, "shared_layer", x, s.use_bn, training, s.act)
    cur = [shared] * s.num_tasks
    for i in range(len(s.dnn_units)):
        cur = [mlp(p, b, f"cross_stitch.task_layer_{i}.{t}", cur[t], s.use_bn, training, s.act)
               for t in range(s.num_tasks)]
        stitched = torch.matmul(torch.cat(cur, -1), p[f"cross_stitch.gate_{i}.cross_stitch_weight"])
        d = s.dnn_units[i]
        cur = [stitched[:, t * d:(t + 1) * d] for t in range(s.num_tasks)]
    return _towers(p, b, s, cur, training)


def forward_hmoe(p: Params, b: Params, s: Spec, X: Tensor, training: bool) -> Tensor:
    """model/hmoe.py:82-139: MMoE up to the towers; task i's head reads sum_j softmax(task_weight_i)[j] * tower_j
    with tower_j detached for j != i (:126-131)."""
    x = gather_concat(X, p, s.columns)
    T = s.num_tasks
    experts = [mlp(p, b, f"expert_dnn.{e}", x, s.use_bn, training, s.act) for e in range(s.num_experts)]
    mixed, weights, towers = [], [], []
    for t in range(T):
        g = mlp(p, b, f"gate_dnn.{t}", x, s.use_bn, training, s.act) if len(s.gate_units) > 0 else x
        mixed.append(_gate_mix(F.linear(g, p[f"gate_dnn_final_layer.{t}.weight"]), experts))
    for t in range(T):
        w = mlp(p, b, f"task_weight.{t}", x, s.use_bn, training, s.act) if len(s.task_weight_units) > 0 else x
        weights.append(F.linear(w, p[f"task_weight_final_layer.{t}.weight"]).softmax(1))
    for t in range(T):
        towers.append(mlp(p, b, f"tower_dnn.{t}", mixed[t], s.use_bn, training, s.act) if len(s.tower_units) > 0
                      else mixed[t])
    outs = []
    for i in range(T):
        h = weights[i][:, i].view(-1, 1) * towers[i]
        for j in range(T):
            if j != i:
                h = h + weights[i][:, j].view(-1, 1) * towers[j].detach()
        logit = F.linear(h, p[f"tower_dnn_final_layer.{i}.weight"])
        outs.append(predict_head(logit, p[f"out.{i}.bias"], s.task_types[i]))
    return torch.cat(outs, -1)


def forward_aitm(p: Params, b: Params, s: Spec, X: Tensor, training: bool) -> Tensor:
    """model/aitm.py:77-108: per-task bottoms (built from ``expert_dnn_hidden_units``, aitm.py:20); task i >= 1 replaces
    its feature by the attention over [g(feat_{i-1}), feat_i] with values / keys / queries from h1 / h2 / h3."""
    x = gather_concat(X, p, s.columns)
    feat = [mlp(p, b, f"bottom.{t}", x, s.use_bn, training, s.act) for t in range(s.num_tasks)]
    H = feat[0].shape[1]
    for i in range(1, s.num_tasks):
        tok = torch.cat([F.linear(feat[i - 1], p[f"g.{i - 1}.weight"], p[f"g.{i - 1}.bias"]).unsqueeze(1),
                         feat[i].unsqueeze(1)], dim=1)
        V = F.linear(tok, p["h1.weight"], p["h1.bias"])
        K = F.linear(tok, p["h2.weight"], p["h2.bias"])
        Q = F.linear(tok, p["h3.weight"], p["h3.bias"])
        a = torch.softmax(torch.sum(K * Q, 2, True) / math.sqrt(H), dim=1)
        feat[i] = torch.sum(a * V, 1)
    return _towers(p, b, s, feat, training)


def _snr_gate(p: Params, prefix: str, xs: List[Tensor], n_out: int) -> List[Tensor]:
    """snr_trans.py:36-50 / mssm.py:38-60: hard-concrete gates z from (u, alpha) -- a scalar per connection [n_out, n_in]
    (SNR-trans) or a vector over the output units [n_out, n_in, U] (MSSM); out_i = sum_j (x_j @ M_ij) * z_ij.
    The transformation matrices ``<prefix>.trans_matrix.<i>.<j>`` are unregistered constants in the reference."""
    alpha = p[f"{prefix}.alpha"]

    def hard_concrete(u):
        s = torch.sigmoid(torch.log(u) - torch.log(1 - u) + torch.log(alpha) / 0.9)
        s_ = s * (1.1 - -0.1) + -0.1
        z = (s_ > 0).float() * s_
        return (z > 1).float() + (z <= 1).float() * z

    if f"{prefix}.u" in p:
        z = hard_concrete(p[f"{prefix}.u"])
    else:   # mssm.py:26-29, :40-50: one [U] vector of gate logits per connection (unregistered constants like the
        # matrices), each pushed through the gate separately -- d(alpha) accumulates connection by connection
        z = [[hard_concrete(p[f"{prefix}.u.{i}.{j}"]) for j in range(len(xs))] for i in range(n_out)]
    outs = []
    for i in range(n_out):
        o = torch.stack([torch.matmul(xs[j], p[f"{prefix}.trans_matrix.{i}.{j}"]) * z[i][j] for j in range(len(xs))], 1)
        outs.append(o.sum(1))
    return outs


def forward_snr_trans(p: Params, b: Params, s: Spec, X: Tensor, training: bool) -> Tensor:
    """model/snr_trans.py:120-156: per level E single-layer experts (level 0 on dnn_input, level l on mixed feature j),
    then the gate; the last gate has one output per task."""
    x = gather_concat(X, p, s.columns)
    feats = [x] * s.num_experts
    levels = len(s.expert_units)
    for l in range(levels):
        outs = [mlp(p, b, f"trans.trans{l + 1}.{j}", feats[j], s.use_bn, training, s.act) for j in range(s.num_experts)]
        feats = _snr_gate(p, f"trans.gate{l + 1}", outs, s.num_tasks if l == levels - 1 else s.num_experts)
    return _towers(p, b, s, feats, training)


def forward_mssm(p: Params, b: Params, s: Spec, X: Tensor, training: bool) -> Tensor:
    """model/mssm.py:140-179: SNR-trans's level structure under the names ``mssm.expert<l>.<j>`` / ``mssm.gate<l>``."""
    x = gather_concat(X, p, s.columns)
    feats = [x] * s.num_experts
    levels = len(s.expert_units)
    for l in range(levels):
        outs = [mlp(p, b, f"mssm.expert{l + 1}.{j}", feats[j], s.use_bn, training, s.act) for j in range(s.num_experts)]
        feats = _snr_gate(p, f"mssm.gate{l + 1}", outs, s.num_tasks if l == levels - 1 else s.num_experts)
    return _towers(p, b, s, feats, training)


def forward_apg(p: Params, b: Params, s: Spec, X: Tensor, training: bool) -> Tensor:
    """model/apg.py:152-176 with the layer of :76-125 in the variant the model builds (use_uv_shared, no P module, no inner /
    generating activation): nk = x @ W_nk + b_nk; kk = nk @ W_kk(scene) + b_kk(scene) per sample; out = act(kk @ W_km + b_km).
    The scene embedding is sparse_embedding_list[X column of the scene feature], detached (:158-159)."""
    rows, dense = gather_fields(X, p, s.columns)
    h = concat_fields(rows, dense)
    scene_col = feature_layout(s.columns)[s.scene_feature][0]
    scene = rows[scene_col].squeeze().detach()
    for l in range(len(s.dnn_units)):
        pre = f"apg_layers.{l}"
        k = p[f"{pre}.shared_weight_nk"].shape[1]
        w_kk = mlp(p, b, f"{pre}.specific_weight_kk", scene, False, training, None).view(-1, k, k)
        b_kk = mlp(p, b, f"{pre}.specific_bias_kk", scene, False, training, None)
        nk = torch.matmul(h, p[f"{pre}.shared_weight_nk"]) + p[f"{pre}.shared_bias_nk"]
        kk = torch.matmul(nk.unsqueeze(1), w_kk).squeeze() + b_kk
        h = torch.matmul(kk, p[f"{pre}.shared_weight_km"]) + p[f"{pre}.shared_bias_km"]
        if s.act is not None:
            h = torch.relu(h)
    outs = [predict_head(F.linear(h, p[f"final_layer.{t}.weight"]), p[f"out.{t}.bias"], s.task_types[t])
            for t in range(s.num_tasks)]
    return torch.cat(outs, -1)


FORWARDS = {
    "mmoe": forward_mmoe, "pcg": forward_mmoe, "ple": forward_ple, "sharedbottom": forward_sharedbottom,
    "esmm": forward_esmm, "star": forward_star, "pepnet": forward_pepnet, "mlp": forward_mlp,
    "cross_stitch": forward_cross_stitch, "hmoe": forward_hmoe, "escm": forward_escm, "aitm": forward_aitm,
    "snr_trans": forward_snr_trans, "mssm": forward_mssm, "apg": forward_apg,
}


# --------------------------------------------------------------------------------------
# L2 regularisation (basemodel.py:514-540; the per-model registrations, e.g. mmoe.py:36-38, :49-51, :60-63)
# --------------------------------------------------------------------------------------
# module prefixes each model registers with l2_reg_dnn, in registration order (a module listed twice counts twice)
REG_MODULES = {
    "mmoe": ["gate_dnn", "tower_dnn", "expert_dnn", "gate_dnn_final_layer", "tower_dnn_final_layer"],
    "ple": ["specific_gate_dnn", "shared_gate_dnn", "tower_dnn", "specific_experts", "shared_experts",
            "specific_gate_dnn_final_layer", "shared_gate_dnn_final_layer", "tower_dnn_final_layer"],
    "sharedbottom": ["tower_dnn", "bottom_dnn", "tower_dnn_final_layer"],
    "esmm": ["ctr_dnn", "cvr_dnn", "ctr_dnn_final_layer", "cvr_dnn_final_layer"],
    "hmoe": ["gate_dnn", "tower_dnn", "task_weight", "expert_dnn", "gate_dnn_final_layer", "task_weight_final_layer",
             "tower_dnn_final_layer"],
    "mlp": ["mlp_layers"],
    "cross_stitch": ["tower_dnn"],
    "star": [], "pepnet": [],
    "escm": ["ctr_dnn", "cvr_dnn", "ctr_dnn_final_layer", "cvr_dnn_final_layer"],
    "aitm": ["tower_dnn", "bottom", "tower_dnn_final_layer"],
    "snr_trans": ["tower_dnn"],
    "mssm": ["tower_dnn"],
    "apg": [],
}
REG_MODULES["pcg"] = REG_MODULES["mmoe"]


def regularization_loss(params: Params, model: str, l2: float) -> Tensor:
    """sum over the registered modules of l2 * sum(w^2) for every parameter whose name inside the module contains
    'weight' and not 'bn' (the reference's filter)."""
    total = torch.zeros((1,))
    if not l2:
        return total
    for mod in REG_MODULES[model]:
        for k, v in params.items():
            if k.startswith(mod + "."):
                local = k[len(mod) + 1:]
                if "weight" in local and "bn" not in local:
                    total = total + torch.sum(l2 * torch.square(v))
    return total


# --------------------------------------------------------------------------------------
# the step body (basemodel.py:262-313)
# --------------------------------------------------------------------------------------
class OracleTrainer:
    """Holds leaf parameter tensors + buffers and reproduces forward -> sum BCE -> backward ->
    torch.optim step.  ``trainable`` is the set of keys that are registered parameters in the
    reference (everything in ``named_parameters()``); the rest are frozen constants."""

    def __init__(self, config: dict, columns: Sequence[Column], params: Params, buffers: Optional[Params] = None,
                 trainable: Optional[Sequence[str]] = None):
        self.spec = Spec(config, columns)
        self.fwd = FORWARDS[self.spec.model]
        keys = list(params.keys()) if trainable is None else list(trainable)
        self.params = {k: v.detach().clone().float() for k, v in params.items()}
        for k in keys:
            self.params[k].requires_grad_(True)
        self.trainable = keys
        self.buffers = {k: v.detach().clone() for k, v in (buffers or {}).items()}
        oc = config["optim_config"]
        self.l2_dnn = config["model_config"].get("l2_reg_dnn", 0)
        self.last_reg = 0.0
        lr, name = oc.get("lr", 1e-3), oc.get("optimizer", "adagrad")
        loss = oc.get("loss", "binary_crossentropy")
        self.loss_names = [loss] * self.spec.num_tasks if isinstance(loss, str) else list(loss)
        leaves = [self.params[k] for k in keys]
        if name == "adam":  # basemodel.py:569-584
            self.optim = torch.optim.Adam(leaves, lr=lr)
        elif name == "adagrad":
            self.optim = torch.optim.Adagrad(leaves, lr=lr)
        elif name == "sgd":
            self.optim = torch.optim.SGD(leaves, lr=lr)
        elif name == "rmsprop":
            self.optim = torch.optim.RMSprop(leaves, lr=lr)
        else:
            raise NotImplementedError(name)

    def _mask_columns(self, domain_mask: Tensor) -> Tensor:
        """[B, T]: the mask entry head t is multiplied by -- column t for 'msl', t % num_domains for 'mtmsl'
        (mmoe.py:101-106)."""
        D = domain_mask.shape[1]
        return domain_mask.float()[:, [t % D for t in range(self.spec.num_tasks)]]

    def forward(self, X: Tensor, training: bool = False, domain_mask: Optional[Tensor] = None) -> Tensor:
        pred = self.fwd(self.params, self.buffers, self.spec, X.float(), training)
        if domain_mask is not None:   # every model multiplies its task outputs by the mask at the very end of forward
            pred = pred * self._mask_columns(domain_mask).to(pred.device)
        return pred

    def loss_and_grads(self, X: Tensor, y: Tensor, domain_mask: Optional[Tensor] = None
                       ) -> Tuple[Tensor, Tensor, Dict[str, Optional[Tensor]]]:
        pred = self.forward(X, training=True, domain_mask=domain_mask)
        self.optim.zero_grad()
        if domain_mask is not None:   # basemodel.py:273-282: BCE weighted by the task's mask column
            w = self._mask_columns(domain_mask).to(pred.device)
            loss = sum(F.binary_cross_entropy(pred[:, t], y.float()[:, t], weight=w[:, t], reduction="sum")
                       for t in range(pred.shape[1]))
        elif self.spec.model == "escm":
            loss = escm_loss(pred, y.float())
        else:
            loss = loss_sum(pred, y.float(), self.loss_names)
        reg = regularization_loss(self.params, self.spec.model, self.l2_dnn).to(loss.device)
        self.last_reg = float(reg.detach())
        (loss + reg).backward()   # basemodel.py:303 total_loss = loss + reg_loss (+ aux / cka terms that are zero)
        grads = {k: (None if self.params[k].grad is None else self.params[k].grad.detach().clone())
                 for k in self.trainable}
        return pred.detach(), loss.detach(), grads

    def step(self, X: Tensor, y: Tensor, domain_mask: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        pred, loss, _ = self.loss_and_grads(X, y, domain_mask)
        self.optim.step()
        return pred, loss

    def state(self) -> Params:
        out = {k: v.detach().clone() for k, v in self.params.items()}
        out.update({k: v.clone() for k, v in self.buffers.items()})
        return out


# --------------------------------------------------------------------------------------
# stand-alone references for single kernels (used by kernel-level parity tests)
# --------------------------------------------------------------------------------------
def embedding_row_grads(X: Tensor, d_input: Tensor, columns: Sequence[Column]) -> Dict[str, Tuple[Tensor, Tensor]]:
    """Per sparse field: (sorted unique ids, summed gradient rows) -- what ATen's
    embedding_dense_backward leaves in the non-zero rows of the dense [V,D] gradient."""
    lay = feature_layout(columns)
    out, at = {}, 0
    for c in columns:
        if c.kind != "sparse":
            continue
        ids = X[:, lay[c.name][0]].long()
        g = d_input[:, at:at + c.dim].double()
        uniq, inv = torch.unique(ids, return_inverse=True)
        acc = torch.zeros(len(uniq), c.dim, dtype=torch.float64).index_add_(0, inv, g)
        out[c.name] = (uniq, acc.float())
        at += c.dim
    return out


def adam_reference(p: Tensor, g: Tensor, m: Tensor, v: Tensor, t: int, lr: float,
                   b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> None:
    """torch/optim/adam.py _single_tensor_adam op order (in place)."""
    m.lerp_(g, 1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


def adagrad_reference(p: Tensor, g: Tensor, s: Tensor, lr: float, eps: float = 1e-10) -> None:
    """torch/optim/adagrad.py _single_tensor_adagrad op order (lr_decay=0, in place)."""
    s.addcmul_(g, g, value=1)
    p.addcdiv_(g, s.sqrt().add_(eps), value=-lr)
