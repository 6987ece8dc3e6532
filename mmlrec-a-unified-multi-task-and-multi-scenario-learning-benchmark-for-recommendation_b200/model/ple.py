"""PLE / CGC (reference: ``/root/reference/model/ple.py``): ``num_levels`` CGC layers, each with
T x ``specific_expert_num`` task experts, ``shared_expert_num`` shared experts, T task gates over
(own + shared) experts and one shared gate over all experts; then towers and heads.

Reference quirks kept on purpose: the constructor allocates ``specific_expert_num``
shared experts per level but the forward uses only the first ``shared_expert_num``; the last
level's shared gate feeds nothing.  Those parameters exist (same ``state_dict``) and never receive
a gradient; the dead last-level shared gate is not computed at all.

Step program per level: ONE grouped launch per MLP layer covering every expert and gate MLP (at
level 0 all of them read dnn_input, so it is a single wide GEMM), one gate-mix launch for all
gates, and in backward one launch with every wgrad + dgrad problem of the layer.
"""
import torch.nn as nn

from ..engine.core import (Builder, GateMixStage, GateSpec, GatherStage, HeadSpec, HeadStage, mlp_stages)
from .basemodel import BaseModel
from .utils import DNN, PredictionLayer


class PLE(BaseModel):
    def __init__(self, dnn_feature_columns, init_std=0.0001, device="cpu", gpus=None, config=None):
        super().__init__(linear_feature_columns=[], dnn_feature_columns=dnn_feature_columns, init_std=init_std,
                         device=device, gpus=gpus, config=config)
        mc = self.model_config
        self.num_experts = mc.get("num_experts", 4)
        self.input_dim = self.compute_input_dim(dnn_feature_columns)
        self.shared_expert_num = mc.get("shared_expert_num", 1)
        self.specific_expert_num = mc.get("specific_expert_num", 3)
        self.num_levels = mc.get("num_levels", 1)
        self.expert_dnn_hidden_units = mc.get("expert_dnn_hidden_units", [256, 128])
        self.gate_dnn_hidden_units = mc.get("gate_dnn_hidden_units", [64])
        self.tower_dnn_hidden_units = mc.get("tower_dnn_hidden_units", [64])
        kw = dict(activation=mc.get("dnn_activation", "relu"), l2_reg=mc.get("l2_reg_dnn", 0),
                  dropout_rate=mc.get("dnn_dropout", 0), use_bn=mc.get("dnn_use_bn", False), init_std=init_std)
        self.dnn_use_bn = kw["use_bn"]
        T, S, Sh, Lv = self.num_tasks, self.specific_expert_num, self.shared_expert_num, self.num_levels
        H = self.expert_dnn_hidden_units[-1]

        def level_in(lv):
            return self.input_dim if lv == 0 else H

        def grid(n_groups, n_each, hidden):
            return nn.ModuleList(nn.ModuleList(nn.ModuleList(DNN(level_in(lv), hidden, **kw) for _ in range(n_each))
                                               for _ in range(n_groups)) for lv in range(Lv))

        self.specific_experts = grid(T, S, self.expert_dnn_hidden_units)
        self.shared_experts = grid(1, S, self.expert_dnn_hidden_units)  # S allocated, Sh used (reference behaviour)
        has_gate_dnn = len(self.gate_dnn_hidden_units) > 0
        if has_gate_dnn:
            self.specific_gate_dnn = grid(T, 1, self.gate_dnn_hidden_units)

        def gate_in(lv):
            return self.gate_dnn_hidden_units[-1] if has_gate_dnn else level_in(lv)

        self.specific_gate_dnn_final_layer = nn.ModuleList(
            nn.ModuleList(nn.Linear(gate_in(lv), S + Sh, bias=False) for _ in range(T)) for lv in range(Lv))
        if has_gate_dnn:
            self.shared_gate_dnn = nn.ModuleList(DNN(level_in(lv), self.gate_dnn_hidden_units, **kw) for lv in range(Lv))
        self.shared_gate_dnn_final_layer = nn.ModuleList(
            nn.Linear(gate_in(lv), T * S + Sh, bias=False) for lv in range(Lv))
        tower_in = H
        if len(self.tower_dnn_hidden_units) > 0:
            self.tower_dnn = nn.ModuleList(DNN(H, self.tower_dnn_hidden_units, **kw) for _ in range(T))
            tower_in = self.tower_dnn_hidden_units[-1]
        self.tower_dnn_final_layer = nn.ModuleList(nn.Linear(tower_in, 1, bias=False) for _ in range(T))
        self.out = nn.ModuleList(PredictionLayer(task) for task in self.task_types)
        # ple.py:57-59, :74-76, :89-91, :99-103 (the unused shared experts are registered too: under L2 they do move)
        self.regularize(([self.specific_gate_dnn, self.shared_gate_dnn] if has_gate_dnn else [])
                        + ([self.tower_dnn] if len(self.tower_dnn_hidden_units) > 0 else [])
                        + [self.specific_experts, self.shared_experts, self.specific_gate_dnn_final_layer,
                           self.shared_gate_dnn_final_layer, self.tower_dnn_final_layer], mc.get("l2_reg_dnn", 0))
        self._finalize()

    def build_graph(self, b: Builder) -> None:
        T, S, Sh, Lv = self.num_tasks, self.specific_expert_num, self.shared_expert_num, self.num_levels
        has_gate_dnn = len(self.gate_dnn_hidden_units) > 0
        x = b.add(GatherStage(b, self)).out
        inputs = [x] * (T + 1)
        for lv in range(Lv):
            last = lv == Lv - 1
            # blocks ordered so that everything reading the same input is adjacent (=> one wide GEMM)
            blocks, where = [], {}
            for t in range(T):
                for j in range(S):
                    where[("spec", t, j)] = len(blocks)
                    blocks.append((inputs[t], self.specific_experts[lv][t][j]))
                if has_gate_dnn:
                    where[("gate", t)] = len(blocks)
                    blocks.append((inputs[t], self.specific_gate_dnn[lv][t][0]))
            for k in range(Sh):
                where[("shared", k)] = len(blocks)
                blocks.append((inputs[T], self.shared_experts[lv][0][k]))
            # the last level's shared gate feeds nothing (Q10) and is skipped -- unless its DNN carries BatchNorm: the
            # reference still runs it (ple.py:141-148), so its running statistics move every training step; here it then
            # rides along in the level's wide GEMM, forward only (no gradient reaches it: its parameters stay put)
            if has_gate_dnn and (not last or self.dnn_use_bn):
                where[("sgate",)] = len(blocks)
                blocks.append((inputs[T], self.shared_gate_dnn[lv]))
            outs = mlp_stages(b, blocks, f"cgc{lv}")
            spec = [outs[where[("spec", t, j)]] for t in range(T) for j in range(S)]
            shared = [outs[where[("shared", k)]] for k in range(Sh)]
            gates = []
            for t in range(T):
                gin = outs[where[("gate", t)]] if has_gate_dnn else inputs[t]
                gates.append(GateSpec(gin, self.specific_gate_dnn_final_layer[lv][t], spec[t * S:(t + 1) * S] + shared))
            if not last:
                gin = outs[where[("sgate",)]] if has_gate_dnn else inputs[T]
                gates.append(GateSpec(gin, self.shared_gate_dnn_final_layer[lv], spec + shared))
            inputs = b.add(GateMixStage(b, gates, label=f"cgc{lv}")).outs
        tower = inputs[:T]
        if len(self.tower_dnn_hidden_units) > 0:
            tower = mlp_stages(b, [(tower[t], self.tower_dnn[t]) for t in range(T)], "towers")
        b.add(HeadStage(b, [HeadSpec(tower[t], self.tower_dnn_final_layer[t], self.out[t].bias, self.task_types[t])
                            for t in range(T)]))
