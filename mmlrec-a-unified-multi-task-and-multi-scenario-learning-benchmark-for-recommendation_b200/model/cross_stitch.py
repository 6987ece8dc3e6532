"""Cross-Stitch network (reference: ``/root/reference/model/cross_stitch.py``): a shared first layer, then per level
T task layers followed by a cross-stitch unit -- ``cat(task outputs) @ W`` with W [T*d, T*d] initialised to the
identity (cross_stitch.py:7-27) -- then T towers and heads.  Same constructor, attribute names and ``state_dict``
keys (``cross_stitch.task_layer_<i>.<t>``, ``cross_stitch.gate_<i>.cross_stitch_weight``).

Step program: the T task layers of a level are one grouped launch whose outputs are adjacent columns of one
buffer, so the concatenation is free and the unit is ONE [B, T*d] x [T*d, T*d] GEMM reading the weight in its
stored (K-by-N) orientation; the per-task slices of its output feed the next level."""
import torch
import torch.nn as nn

from ..engine.core import Builder, GatherStage, HeadSpec, HeadStage, LinearSpec, LinearStage, mlp_stages
from .basemodel import BaseModel
from .utils import DNN, PredictionLayer


class CrossStitchLayer(nn.Module):
    def __init__(self, input_dims, device="cpu"):
        super().__init__()
        self.last_dims = list(input_dims)
        self.total_last_dim = sum(self.last_dims)
        self.cross_stitch_weight = nn.Parameter(nn.init.eye_(torch.empty(self.total_last_dim, self.total_last_dim)))

    weight = property(lambda self: self.cross_stitch_weight)   # what LinearSpec reads
    bias = None


class CrossStitch(BaseModel):
    def __init__(self, dnn_feature_columns, init_std=0.0001, device="cpu", gpus=None, config=None):
        super().__init__(linear_feature_columns=[], dnn_feature_columns=dnn_feature_columns, init_std=init_std,
                         device=device, gpus=gpus, config=config)
        mc = self.model_config
        self.shared_hidden_unit = mc.get("shared_hidden_unit", 256)
        self.dnn_hidden_units = mc.get("dnn_hidden_units", [256, 128])
        self.tower_dnn_hidden_units = mc.get("tower_dnn_hidden_units", [64])
        kw = dict(activation=mc.get("dnn_activation", "relu"), l2_reg=mc.get("l2_reg_dnn", 0),
                  dropout_rate=mc.get("dnn_dropout", 0), use_bn=mc.get("dnn_use_bn", False), init_std=init_std)
        T, units = self.num_tasks, self.dnn_hidden_units
        self.input_dim = self.compute_input_dim(dnn_feature_columns)
        self.shared_layer = DNN(self.input_dim, [self.shared_hidden_unit], **kw)
        self.cross_stitch = nn.ModuleDict()
        for i in range(len(units)):
            d_in = self.shared_hidden_unit if i == 0 else units[i - 1]
            self.cross_stitch[f"task_layer_{i}"] = nn.ModuleList(DNN(d_in, [units[i]], **kw) for _ in range(T))
            self.cross_stitch[f"gate_{i}"] = CrossStitchLayer([units[i]] * T)
        tower_in = units[-1]
        if len(self.tower_dnn_hidden_units) > 0:
            self.tower_dnn = nn.ModuleList(DNN(tower_in, self.tower_dnn_hidden_units, **kw) for _ in range(T))
            tower_in = self.tower_dnn_hidden_units[-1]
        self.tower_dnn_final_layer = nn.ModuleList(nn.Linear(tower_in, 1, bias=False) for _ in range(T))
        self.out = nn.ModuleList(PredictionLayer(task) for task in self.task_types)
        if len(self.tower_dnn_hidden_units) > 0:
            self.regularize([self.tower_dnn], mc.get("l2_reg_dnn", 0))   # cross_stitch.py:70-72 (only the towers are registered)
        self._finalize()

    def build_graph(self, b: Builder) -> None:
        T = self.num_tasks
        x = b.add(GatherStage(b, self)).out
        (shared,) = mlp_stages(b, [(x, self.shared_layer)], "shared")
        cur = [shared] * T
        for i, d in enumerate(self.dnn_hidden_units):
            outs = mlp_stages(b, [(cur[t], self.cross_stitch[f"task_layer_{i}"][t]) for t in range(T)], f"task{i}")
            unit = self.cross_stitch[f"gate_{i}"]
            stage = b.add(LinearStage(b, [LinearSpec(outs[0].group.span(), unit, None, transposed=True)], None,
                                      label=f"stitch{i}"))
            cur = [stage.outs[0].sub(t * d, d) for t in range(T)]
        tower = cur
        if len(self.tower_dnn_hidden_units) > 0:
            tower = mlp_stages(b, [(cur[t], self.tower_dnn[t]) for t in range(T)], "towers")
        b.add(HeadStage(b, [HeadSpec(tower[t], self.tower_dnn_final_layer[t], self.out[t].bias, self.task_types[t])
                            for t in range(T)]))
