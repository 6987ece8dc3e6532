"""Shared-Bottom (reference: ``/root/reference/model/sharedbottom.py``): one bottom MLP, T tower
MLPs on its output (one wide GEMM: they all read the same activation), heads."""
import torch.nn as nn

from ..engine.core import Builder, GatherStage, HeadSpec, HeadStage, mlp_stages
from .basemodel import BaseModel
from .utils import DNN, PredictionLayer


class SharedBottom(BaseModel):
    def __init__(self, dnn_feature_columns, init_std=0.0001, device="cpu", gpus=None, config=None):
        super().__init__(linear_feature_columns=[], dnn_feature_columns=dnn_feature_columns, init_std=init_std,
                         device=device, gpus=gpus, config=config)
        mc = self.model_config
        self.num_experts = mc.get("num_experts", 4)
        self.input_dim = self.compute_input_dim(dnn_feature_columns)
        self.bottom_dnn_hidden_units = mc.get("bottom_dnn_hidden_units", [256, 128])
        self.gate_dnn_hidden_units = mc.get("gate_dnn_hidden_units", [64])
        self.tower_dnn_hidden_units = mc.get("tower_dnn_hidden_units", [64])
        kw = dict(activation=mc.get("dnn_activation", "relu"), dropout_rate=mc.get("dnn_dropout", 0),
                  use_bn=mc.get("dnn_use_bn", False), init_std=init_std)
        T = self.num_tasks
        self.bottom_dnn = DNN(self.input_dim, self.bottom_dnn_hidden_units, **kw)
        tower_in = self.bottom_dnn_hidden_units[-1]
        if len(self.tower_dnn_hidden_units) > 0:
            self.tower_dnn = nn.ModuleList(DNN(tower_in, self.tower_dnn_hidden_units, **kw) for _ in range(T))
            tower_in = self.tower_dnn_hidden_units[-1]
        self.tower_dnn_final_layer = nn.ModuleList(nn.Linear(tower_in, 1, bias=False) for _ in range(T))
        self.out = nn.ModuleList(PredictionLayer(task) for task in self.task_types)
        # sharedbottom.py:36-38, :45-49
        self.regularize(([self.tower_dnn] if len(self.tower_dnn_hidden_units) > 0 else [])
                        + [self.bottom_dnn, self.tower_dnn_final_layer], mc.get("l2_reg_dnn", 0))
        self._finalize()

    def build_graph(self, b: Builder) -> None:
        T = self.num_tasks
        x = b.add(GatherStage(b, self)).out
        (bottom,) = mlp_stages(b, [(x, self.bottom_dnn)], "bottom")
        tower = [bottom] * T
        if len(self.tower_dnn_hidden_units) > 0:
            tower = mlp_stages(b, [(bottom, self.tower_dnn[t]) for t in range(T)], "towers")
        b.add(HeadStage(b, [HeadSpec(tower[t], self.tower_dnn_final_layer[t], self.out[t].bias, self.task_types[t])
                            for t in range(T)]))
