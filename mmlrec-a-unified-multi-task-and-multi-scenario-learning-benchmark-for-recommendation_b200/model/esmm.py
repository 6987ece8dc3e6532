"""ESMM (reference: ``/root/reference/model/esmm.py``): CTR and CVR MLPs over the same input (one
wide GEMM per layer), bias-free heads, ONE shared PredictionLayer bias (``out.bias``) used by both
heads, output ``[pCTR, pCTR * pCVR]`` (the product and both BCE terms live in the fused head kernel)."""
import torch.nn as nn

from ..engine.core import Builder, GatherStage, HeadSpec, HeadStage, mlp_stages
from .basemodel import BaseModel
from .utils import DNN


class ESMM(BaseModel):
    def __init__(self, dnn_feature_columns, init_std=0.0001, device="cpu", gpus=None, config=None):
        super().__init__(linear_feature_columns=[], dnn_feature_columns=dnn_feature_columns, init_std=init_std,
                         device=device, gpus=gpus, config=config)
        mc = self.model_config
        if self.num_tasks != 2:
            raise ValueError("ESMM has exactly two outputs (ctr, ctcvr)")
        self.num_experts = mc.get("num_experts", 4)
        self.input_dim = self.compute_input_dim(dnn_feature_columns)
        self.expert_dnn_hidden_units = mc.get("expert_dnn_hidden_units", [256, 128])
        self.gate_dnn_hidden_units = mc.get("gate_dnn_hidden_units", [64])
        self.tower_dnn_hidden_units = mc.get("tower_dnn_hidden_units", [64])
        kw = dict(activation=mc.get("dnn_activation", "relu"), dropout_rate=mc.get("dnn_dropout", 0),
                  use_bn=mc.get("dnn_use_bn", False), init_std=init_std)
        self.ctr_dnn = DNN(self.input_dim, self.expert_dnn_hidden_units, **kw)
        self.cvr_dnn = DNN(self.input_dim, self.expert_dnn_hidden_units, **kw)
        self.ctr_dnn_final_layer = nn.Linear(self.expert_dnn_hidden_units[-1], 1, bias=False)
        self.cvr_dnn_final_layer = nn.Linear(self.expert_dnn_hidden_units[-1], 1, bias=False)
        # esmm.py:38-43
        self.regularize([self.ctr_dnn, self.cvr_dnn, self.ctr_dnn_final_layer, self.cvr_dnn_final_layer],
                        mc.get("l2_reg_dnn", 0))
        self._finalize()

    def build_graph(self, b: Builder) -> None:
        x = b.add(GatherStage(b, self)).out
        ctr, cvr = mlp_stages(b, [(x, self.ctr_dnn), (x, self.cvr_dnn)], "ctr_cvr")
        task = self.model_config.get("task", "binary")
        b.add(HeadStage(b, [HeadSpec(ctr, self.ctr_dnn_final_layer, self.out.bias, task),
                            HeadSpec(cvr, self.cvr_dnn_final_layer, self.out.bias, task)], esmm=True))
