"""MMoE (reference: ``/root/reference/model/mmoe.py``): E expert MLPs and T gate MLPs over the same
gathered input, bias-free gate heads -> softmax -> expert mixture per task, T tower MLPs, bias-free
1-unit heads, PredictionLayers.  Same constructor, attribute names and ``state_dict`` keys.

Step program: gather -> [experts + gates, layer by layer: one wide GEMM for everything that reads
dnn_input] -> gate-mix (all tasks, one launch) -> towers (grouped) -> fused heads + BCE.
"""
import torch.nn as nn

from ..engine.core import (Builder, GateMixStage, GateSpec, GatherStage, HeadSpec, HeadStage, mlp_stages)
from .basemodel import BaseModel
from .utils import DNN, PredictionLayer


class MMOE(BaseModel):
    def __init__(self, dnn_feature_columns, init_std=0.0001, device="cpu", gpus=None, config=None):
        super().__init__(linear_feature_columns=[], dnn_feature_columns=dnn_feature_columns, init_std=init_std,
                         device=device, gpus=gpus, config=config)
        mc = self.model_config
        self.num_experts = mc.get("num_experts", 4)
        self.input_dim = self.compute_input_dim(dnn_feature_columns)
        self.expert_dnn_hidden_units = mc.get("expert_dnn_hidden_units", [256, 128])
        self.gate_dnn_hidden_units = mc.get("gate_dnn_hidden_units", [64])
        self.tower_dnn_hidden_units = mc.get("tower_dnn_hidden_units", [64])
        kw = dict(activation=mc.get("dnn_activation", "relu"), l2_reg=mc.get("l2_reg_dnn", 0),
                  dropout_rate=mc.get("dnn_dropout", 0), use_bn=mc.get("dnn_use_bn", False), init_std=init_std)
        T, E = self.num_tasks, self.num_experts
        self.expert_dnn = nn.ModuleList(DNN(self.input_dim, self.expert_dnn_hidden_units, **kw) for _ in range(E))
        gate_in = self.input_dim
        if len(self.gate_dnn_hidden_units) > 0:
            self.gate_dnn = nn.ModuleList(DNN(self.input_dim, self.gate_dnn_hidden_units, **kw) for _ in range(T))
            gate_in = self.gate_dnn_hidden_units[-1]
        self.gate_dnn_final_layer = nn.ModuleList(nn.Linear(gate_in, E, bias=False) for _ in range(T))
        tower_in = self.expert_dnn_hidden_units[-1]
        if len(self.tower_dnn_hidden_units) > 0:
            self.tower_dnn = nn.ModuleList(DNN(tower_in, self.tower_dnn_hidden_units, **kw) for _ in range(T))
            tower_in = self.tower_dnn_hidden_units[-1]
        self.tower_dnn_final_layer = nn.ModuleList(nn.Linear(tower_in, 1, bias=False) for _ in range(T))
        self.out = nn.ModuleList(PredictionLayer(task) for task in self.task_types)
        # mmoe.py:36-38, :49-51, :60-63
        self.regularize(([self.gate_dnn] if len(self.gate_dnn_hidden_units) > 0 else [])
                        + ([self.tower_dnn] if len(self.tower_dnn_hidden_units) > 0 else [])
                        + [self.expert_dnn, self.gate_dnn_final_layer, self.tower_dnn_final_layer], mc.get("l2_reg_dnn", 0))
        self._finalize()

    def build_graph(self, b: Builder) -> None:
        T, E = self.num_tasks, self.num_experts
        x = b.add(GatherStage(b, self)).out
        blocks = [(x, self.expert_dnn[e]) for e in range(E)]
        if len(self.gate_dnn_hidden_units) > 0:
            blocks += [(x, self.gate_dnn[t]) for t in range(T)]
        outs = mlp_stages(b, blocks, "experts_gates")
        experts = outs[:E]
        gate_inputs = outs[E:] if len(self.gate_dnn_hidden_units) > 0 else [x] * T
        mix = b.add(GateMixStage(b, [GateSpec(gate_inputs[t], self.gate_dnn_final_layer[t], experts)
                                     for t in range(T)], label="mmoe")).outs
        tower = mix
        if len(self.tower_dnn_hidden_units) > 0:
            tower = mlp_stages(b, [(mix[t], self.tower_dnn[t]) for t in range(T)], "towers")
        b.add(HeadStage(b, [HeadSpec(tower[t], self.tower_dnn_final_layer[t], self.out[t].bias, self.task_types[t])
                            for t in range(T)]))
