"""HMoE (reference: ``/root/reference/model/hmoe.py``): MMoE whose tower outputs are mixed once more -- task i's head
reads ``sum_j softmax(task_weight_i)[j] * tower_j`` with every ``tower_j``, j != i, DETACHED (hmoe.py:126-131): the
other tasks' towers contribute their values and steer task i's weights, but receive no gradient from task i.
Same constructor, attribute names and ``state_dict`` keys.

Step program: MMoE's stages, with the task-weight DNNs riding in the same wide GEMM as the experts and gate DNNs,
then a second gate-mix stage over the tower outputs whose (gate i, expert j != i) pairs are marked detached."""
import torch.nn as nn

from ..engine.core import (Builder, GateMixStage, GateSpec, GatherStage, HeadSpec, HeadStage, mlp_stages)
from .basemodel import BaseModel
from .utils import DNN, PredictionLayer


class HMOE(BaseModel):
    def __init__(self, dnn_feature_columns, init_std=0.0001, device="cpu", gpus=None, config=None):
        super().__init__(linear_feature_columns=[], dnn_feature_columns=dnn_feature_columns, init_std=init_std,
                         device=device, gpus=gpus, config=config)
        mc = self.model_config
        self.num_experts = mc.get("num_experts", 4)
        self.input_dim = self.compute_input_dim(dnn_feature_columns)
        self.expert_dnn_hidden_units = mc.get("expert_dnn_hidden_units", [256, 128])
        self.gate_dnn_hidden_units = mc.get("gate_dnn_hidden_units", [64])
        self.tower_dnn_hidden_units = mc.get("tower_dnn_hidden_units", [64])
        self.task_weight_hidden_units = mc.get("task_weight_hidden_units", [64])
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
        tw_in = self.input_dim
        if len(self.task_weight_hidden_units) > 0:
            self.task_weight = nn.ModuleList(DNN(self.input_dim, self.task_weight_hidden_units, **kw) for _ in range(T))
            tw_in = self.task_weight_hidden_units[-1]
        self.task_weight_final_layer = nn.ModuleList(nn.Linear(tw_in, T, bias=False) for _ in range(T))
        self.tower_dnn_final_layer = nn.ModuleList(nn.Linear(tower_in, 1, bias=False) for _ in range(T))
        self.out = nn.ModuleList(PredictionLayer(task) for task in self.task_types)
        # hmoe.py:39-41, :51-53, :61-63, :75-80
        self.regularize(([self.gate_dnn] if len(self.gate_dnn_hidden_units) > 0 else [])
                        + ([self.tower_dnn] if len(self.tower_dnn_hidden_units) > 0 else [])
                        + ([self.task_weight] if len(self.task_weight_hidden_units) > 0 else [])
                        + [self.expert_dnn, self.gate_dnn_final_layer, self.task_weight_final_layer,
                           self.tower_dnn_final_layer], mc.get("l2_reg_dnn", 0))
        self._finalize()

    def build_graph(self, b: Builder) -> None:
        T, E = self.num_tasks, self.num_experts
        x = b.add(GatherStage(b, self)).out
        blocks = [(x, self.expert_dnn[e]) for e in range(E)]
        has_gate, has_tw = len(self.gate_dnn_hidden_units) > 0, len(self.task_weight_hidden_units) > 0
        if has_gate:
            blocks += [(x, self.gate_dnn[t]) for t in range(T)]
        if has_tw:
            blocks += [(x, self.task_weight[t]) for t in range(T)]
        outs = mlp_stages(b, blocks, "experts_gates")
        experts = outs[:E]
        at = E
        gate_inputs = outs[at:at + T] if has_gate else [x] * T
        at += T if has_gate else 0
        tw_inputs = outs[at:at + T] if has_tw else [x] * T
        mix = b.add(GateMixStage(b, [GateSpec(gate_inputs[t], self.gate_dnn_final_layer[t], experts)
                                     for t in range(T)], label="mmoe")).outs
        tower = mix
        if len(self.tower_dnn_hidden_units) > 0:
            tower = mlp_stages(b, [(mix[t], self.tower_dnn[t]) for t in range(T)], "towers")
        blend = b.add(GateMixStage(b, [GateSpec(tw_inputs[t], self.task_weight_final_layer[t], list(tower),
                                                detach=[j != t for j in range(T)])
                                       for t in range(T)], label="task_weights")).outs
        b.add(HeadStage(b, [HeadSpec(blend[t], self.tower_dnn_final_layer[t], self.out[t].bias, self.task_types[t])
                            for t in range(T)]))
