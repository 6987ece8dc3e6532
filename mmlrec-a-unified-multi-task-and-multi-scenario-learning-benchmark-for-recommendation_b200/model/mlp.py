"""MLP baseline (reference: ``/root/reference/model/mlp.py``): one shared stack of single-layer DNN blocks and ONE
bias-free 1-unit final layer whose logit every task shares; the tasks differ only by their PredictionLayer biases,
which ACCUMULATE along the task list (task t's logit carries biases 0..t: the PredictionLayer adds its bias in place to
the one shared logit tensor, mlp.py:45-52, model/utils.py:243-245).  Same constructor, attribute names and ``state_dict`` keys."""
import torch.nn as nn

from ..engine.core import Builder, GatherStage, HeadSpec, HeadStage, mlp_stages
from .basemodel import BaseModel
from .utils import DNN, PredictionLayer


class MLP(BaseModel):
    def __init__(self, dnn_feature_columns, init_std=0.0001, device="cpu", gpus=None, config=None):
        super().__init__(linear_feature_columns=[], dnn_feature_columns=dnn_feature_columns, init_std=init_std,
                         device=device, gpus=gpus, config=config)
        mc = self.model_config
        self.dnn_use_bn = mc.get("dnn_use_bn", False)
        self.dnn_hidden_units = mc.get("dnn_hidden_units", [256, 128])
        dims = [self.compute_input_dim(dnn_feature_columns)] + list(self.dnn_hidden_units)
        # mlp.py:24-26: the blocks take neither init_std nor use_bn from the config (DNN defaults: 1e-4, no BatchNorm)
        self.mlp_layers = nn.ModuleList(DNN(dims[i], [dims[i + 1]], activation="relu", l2_reg=mc.get("l2_reg_dnn", 0))
                                        for i in range(len(self.dnn_hidden_units)))
        self.final_layer = nn.Linear(self.dnn_hidden_units[-1], 1, bias=False)
        self.out = nn.ModuleList(PredictionLayer(task) for task in self.task_types)
        if any(t != "binary" for t in self.task_types):
            # the in-place bias add of the PredictionLayers aliases ONE logit tensor (see HeadStage.cumulative_bias); with
            # a regression task the reference's own outputs alias each other -- not reproduced
            raise NotImplementedError("MLP supports binary tasks only")
        self.regularize([self.mlp_layers], mc.get("l2_reg_dnn", 0))   # mlp.py:31-33
        self._finalize()

    def build_graph(self, b: Builder) -> None:
        h = b.add(GatherStage(b, self)).out
        for i, block in enumerate(self.mlp_layers):
            (h,) = mlp_stages(b, [(h, block)], f"mlp{i}")
        # every head reads the same activation through the same weight: the head kernel adds their dw contributions
        b.add(HeadStage(b, [HeadSpec(h, self.final_layer, self.out[t].bias, self.task_types[t])
                            for t in range(self.num_tasks)], cumulative_bias=True))
