"""ESCM-IPW (reference: ``/root/reference/model/escm.py`` with the loss of ``model/basemodel.py:284-292``): ESMM's two
towers over the gathered input with ONE shared PredictionLayer bias; the model returns ``[p_ctr, p_cvr, p_ctr * p_cvr]``
and trains on ``BCE(ctr) + 0.1 * IPW(BCE(cvr)) + 1 * BCE(ctcvr)`` where the inverse-propensity weights come from the
predicted click probability (``counterfact_ipw``, escm.py:98-111).  ``predict`` / ``evaluate`` use the ctr and ctcvr
columns (basemodel.py:438-441).  Same constructor, attribute names and ``state_dict`` keys; the doubly-robust variant
(``model_name = "escm_dr"``, an extra imputation tower that the reference's loss never uses) is not built."""
import torch.nn as nn

from ..engine.core import Builder, GatherStage, HeadSpec, HeadStage, mlp_stages
from .basemodel import BaseModel
from .utils import DNN


class ESCM(BaseModel):
    def __init__(self, dnn_feature_columns, init_std=0.0001, device="cpu", gpus=None, config=None):
        super().__init__(linear_feature_columns=[], dnn_feature_columns=dnn_feature_columns, init_std=init_std,
                         device=device, gpus=gpus, config=config)
        mc = self.model_config
        self.model_name = mc.get("model_name", "escm")
        if self.model_name == "escm_dr":
            raise NotImplementedError("escm_dr (imputation tower) is not built")
        if self.num_tasks != 2:
            raise ValueError("the length of task_names must be equal to 2")
        if any(t != "binary" for t in self.task_types):
            raise ValueError("task must be binary in ESCM")
        self.counterfactual_w, self.global_w = 0.1, 1     # escm.py:33-34
        self.input_dim = self.compute_input_dim(dnn_feature_columns)
        self.expert_dnn_hidden_units = mc.get("expert_dnn_hidden_units", [256, 128])
        kw = dict(activation=mc.get("dnn_activation", "relu"), dropout_rate=mc.get("dnn_dropout", 0),
                  use_bn=mc.get("dnn_use_bn", False), init_std=init_std)
        self.ctr_dnn = DNN(self.input_dim, self.expert_dnn_hidden_units, **kw)
        self.cvr_dnn = DNN(self.input_dim, self.expert_dnn_hidden_units, **kw)
        self.ctr_dnn_final_layer = nn.Linear(self.expert_dnn_hidden_units[-1], 1, bias=False)
        self.cvr_dnn_final_layer = nn.Linear(self.expert_dnn_hidden_units[-1], 1, bias=False)
        # escm.py:66-71
        self.regularize([self.ctr_dnn, self.cvr_dnn, self.ctr_dnn_final_layer, self.cvr_dnn_final_layer],
                        mc.get("l2_reg_dnn", 0))
        self._finalize()

    def build_graph(self, b: Builder) -> None:
        x = b.add(GatherStage(b, self)).out
        ctr, cvr = mlp_stages(b, [(x, self.ctr_dnn), (x, self.cvr_dnn)], "ctr_cvr")
        task = self.model_config.get("task", "binary")
        b.add(HeadStage(b, [HeadSpec(ctr, self.ctr_dnn_final_layer, self.out.bias, task),
                            HeadSpec(cvr, self.cvr_dnn_final_layer, self.out.bias, task)],
                        escm=(self.counterfactual_w, self.global_w)))
