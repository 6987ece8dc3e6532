"""AITM (reference: ``/root/reference/model/aitm.py``): one bottom MLP per task; task 1's feature is replaced by an
attention over two tokens -- ``g(feat_0)`` (the information transferred from task 0) and its own ``feat_1`` -- whose
values / keys / queries come from the SAME three Linear layers ``h1`` / ``h2`` / ``h3`` (aitm.py:82-91); then the towers
and heads.  Same constructor, attribute names and ``state_dict`` keys (``g.0``, ``h1``, ``h2``, ``h3``, ``bottom.i``,
``tower_dnn.i``, ``tower_dnn_final_layer.i``, ``out.i``).

Step program: both bottoms in one grouped GEMM per layer -> ``g`` -> ONE launch with the six projections (h1 | h2 | h3
are adjacent in the flat store, so each token is one [B, H] x [H, 3H] problem; the two applications of the shared
weights write their weight gradients into separate gradient slices) -> the pair-attention kernel -> towers -> heads.
``feat_0`` is read by two stages (``g`` and tower 0), so its gradient is accumulated in fp32."""
import torch.nn as nn

from ..engine.core import (Builder, GatherStage, HeadSpec, HeadStage, LinearSpec, LinearStage, PairAttentionStage,
                           mlp_stages)
from .basemodel import BaseModel
from .utils import DNN, PredictionLayer


class AITM(BaseModel):
    def __init__(self, dnn_feature_columns, init_std=0.0001, device="cpu", gpus=None, config=None):
        super().__init__(linear_feature_columns=[], dnn_feature_columns=dnn_feature_columns, init_std=init_std,
                         device=device, gpus=gpus, config=config)
        mc = self.model_config
        self.num_experts = mc.get("num_experts", 4)
        self.input_dim = self.compute_input_dim(dnn_feature_columns)
        self.bottom_dnn_hidden_units = mc.get("expert_dnn_hidden_units", [256, 128])   # aitm.py:20: the experts' key
        self.gate_dnn_hidden_units = mc.get("gate_dnn_hidden_units", [64])
        self.tower_dnn_hidden_units = mc.get("tower_dnn_hidden_units", [64])
        if self.num_tasks != 2:   # aitm.py:31-32
            raise ValueError("the length of task_names must be equal to 2")
        for task_type in self.task_types:   # aitm.py:38-40
            if task_type != "binary":
                raise ValueError("task must be binary in ESMM, {} is illegal".format(task_type))
        kw = dict(activation=mc.get("dnn_activation", "relu"), dropout_rate=mc.get("dnn_dropout", 0),
                  use_bn=mc.get("dnn_use_bn", False), init_std=init_std)
        T, H = self.num_tasks, self.bottom_dnn_hidden_units[-1]
        self.g = nn.ModuleList(nn.Linear(H, H) for _ in range(T - 1))
        self.h1, self.h2, self.h3 = nn.Linear(H, H), nn.Linear(H, H), nn.Linear(H, H)
        self.bottom = nn.ModuleList(DNN(self.input_dim, self.bottom_dnn_hidden_units, l2_reg=mc.get("l2_reg_dnn", 0), **kw)
                                    for _ in range(T))
        tower_in = H
        if len(self.tower_dnn_hidden_units) > 0:
            self.tower_dnn = nn.ModuleList(DNN(H, self.tower_dnn_hidden_units, **kw) for _ in range(T))
            tower_in = self.tower_dnn_hidden_units[-1]
        self.tower_dnn_final_layer = nn.ModuleList(nn.Linear(tower_in, 1, bias=False) for _ in range(T))
        self.out = nn.ModuleList(PredictionLayer(task) for task in self.task_types)
        # aitm.py:59-61, :70-75 (g and h1..h3 are not regularised)
        self.regularize(([self.tower_dnn] if len(self.tower_dnn_hidden_units) > 0 else [])
                        + [self.bottom, self.tower_dnn_final_layer], mc.get("l2_reg_dnn", 0))
        self._finalize()

    def build_graph(self, b: Builder) -> None:
        T = self.num_tasks
        x = b.add(GatherStage(b, self)).out
        feat = mlp_stages(b, [(x, self.bottom[t]) for t in range(T)], "bottom")
        # feat[0] feeds g AND tower 0 (two stages write its gradient): accumulate in fp32
        feat[0].group.grad_dtype = "f32"
        (p,) = b.add(LinearStage(b, [LinearSpec(feat[0], self.g[0])], None, label="g")).outs
        proj = b.add(LinearStage(b, [LinearSpec(tok, h) for tok in (p, feat[1]) for h in (self.h1, self.h2, self.h3)],
                                 None, label="h")).outs          # V_p K_p Q_p V_q K_q Q_q (aitm.py:86-88)
        att = b.add(PairAttentionStage(b, list(proj), label="ait")).out
        tower = [feat[0], att]
        if len(self.tower_dnn_hidden_units) > 0:
            tower = mlp_stages(b, [(tower[t], self.tower_dnn[t]) for t in range(T)], "towers")
        b.add(HeadStage(b, [HeadSpec(tower[t], self.tower_dnn_final_layer[t], self.out[t].bias, self.task_types[t])
                            for t in range(T)]))
