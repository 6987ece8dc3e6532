"""SNR-trans (reference: ``/root/reference/model/snr_trans.py``): levels of E single-layer experts; after each level a
``gate`` mixes the E expert outputs into E (inner levels) or T (last level) new features,
``out_i = sum_j z_ij * (x_j @ M_ij)``, with one hard-concrete scalar ``z_ij`` per connection (trained through ``u`` and
``alpha``) and a [U, U] transformation ``M_ij`` per connection; level l+1's expert j reads mixed feature j; towers and
heads on the last gate's outputs.  Same constructor, attribute names and ``state_dict`` keys (``trans.trans<l>.<j>``,
``trans.gate<l>.alpha`` / ``.u``, ``tower_dnn.i``, ``tower_dnn_final_layer.i``, ``out.i``).

Reference quirk kept on purpose: the transformation matrices live in a plain Python list of lists
(snr_trans.py:31-34), so they are never registered, never moved by the optimizer and absent from ``state_dict`` -- they
stay at their Xavier-normal initial values.  Here they are one non-persistent buffer per gate (same RNG draw order).

Step program: every level is one grouped GEMM for the E experts, a ``SnrGateStage`` that rebuilds the gate's derived
weight ``W_eff[i*U+v, j*U+u] = z_ij M_ij[u, v]`` (one tiny kernel), and ONE GEMM [B, E*U] x [E*U, n_out*U] applying it to
the concatenated expert outputs (adjacent columns of one buffer: no copy); backward folds d(W_eff) into d(u), d(alpha)."""
import torch
import torch.nn as nn

from ..engine.core import (Builder, GatherStage, HeadSpec, HeadStage, LinearSpec, LinearStage, SnrGateStage, mlp_stages)
from .basemodel import BaseModel
from .utils import DNN, PredictionLayer


class gate(nn.Module):   # (the reference's class name)
    def __init__(self, input_dim, output_dim, units, device="cpu", **unused):
        super().__init__()
        self.input_dim, self.output_dim, self.units = input_dim, output_dim, units
        self.beta, self.gamma, self.eplison, self.e = 0.9, -0.1, 1.1, 1e-8
        self.alpha = nn.Parameter(torch.rand((1,)))
        self.u = nn.Parameter(nn.init.uniform_(torch.empty(output_dim, input_dim), self.e, 1 - self.e))
        self.u._mm_nopad = True       # dense [n_out, n_in]: read by the gate kernels only
        m = torch.empty(output_dim, input_dim, units, units)
        for i in range(output_dim):   # snr_trans.py:31-34: same draw order
            for j in range(input_dim):
                nn.init.xavier_normal_(m[i, j])
        self.register_buffer("trans_matrix", m, persistent=False)


class SNR_trans(BaseModel):
    def __init__(self, dnn_feature_columns, init_std=0.0001, device="cpu", gpus=None, config=None):
        super().__init__(linear_feature_columns=[], dnn_feature_columns=dnn_feature_columns, init_std=init_std,
                         device=device, gpus=gpus, config=config)
        mc = self.model_config
        self.num_experts = mc.get("num_experts", 4)
        self.input_dim = self.compute_input_dim(dnn_feature_columns)
        self.expert_dnn_hidden_units = mc.get("expert_dnn_hidden_units", [256, 128])
        self.gate_dnn_hidden_units = mc.get("gate_dnn_hidden_units", [64])
        self.tower_dnn_hidden_units = mc.get("tower_dnn_hidden_units", [64])
        if self.num_tasks <= 1:
            raise ValueError("num_tasks must be greater than 1")
        if self.num_experts <= 1:
            raise ValueError("num_experts must be greater than 1")
        kw = dict(activation=mc.get("dnn_activation", "relu"), l2_reg=mc.get("l2_reg_dnn", 0),
                  dropout_rate=mc.get("dnn_dropout", 0), use_bn=mc.get("dnn_use_bn", False), init_std=init_std)
        E, T, units = self.num_experts, self.num_tasks, list(self.expert_dnn_hidden_units)
        self.hidden_units = [self.input_dim] + units
        self.trans = nn.ModuleDict()
        for l, width in enumerate(units):   # snr_trans.py:73-98: trans<l+1> then gate<l+1>
            self.trans[f"trans{l + 1}"] = nn.ModuleList(DNN(self.hidden_units[l], [width], **kw) for _ in range(E))
            self.trans[f"gate{l + 1}"] = gate(E, T if l == len(units) - 1 else E, width)
        tower_in = units[-1]
        if len(self.tower_dnn_hidden_units) > 0:
            self.tower_dnn = nn.ModuleList(DNN(tower_in, self.tower_dnn_hidden_units, **kw) for _ in range(T))
            tower_in = self.tower_dnn_hidden_units[-1]
        self.tower_dnn_final_layer = nn.ModuleList(nn.Linear(tower_in, 1, bias=False) for _ in range(T))
        self.out = nn.ModuleList(PredictionLayer(task) for task in self.task_types)
        # snr_trans.py:106-108: only the towers are registered for L2
        self.regularize([self.tower_dnn] if len(self.tower_dnn_hidden_units) > 0 else [], mc.get("l2_reg_dnn", 0))
        self._finalize()

    def load_unregistered(self, tensors):
        """Set the constant transformation matrices from ``trans.gate<l>.trans_matrix.<i>.<j>`` (the names
        tests/golden/make_golden.py exports the reference's unregistered tensors under)."""
        with torch.no_grad():
            for key, value in tensors.items():
                prefix, i, j = key.rsplit(".", 2)
                mod_name, attr = prefix.rsplit(".", 1)
                assert attr == "trans_matrix", key
                self.get_submodule(mod_name).trans_matrix[int(i), int(j)].copy_(torch.as_tensor(value))

    def build_graph(self, b: Builder) -> None:
        T, E = self.num_tasks, self.num_experts
        x = b.add(GatherStage(b, self)).out
        feats = [x] * E
        for l, width in enumerate(self.expert_dnn_hidden_units):
            outs = mlp_stages(b, [(feats[j], self.trans[f"trans{l + 1}"][j]) for j in range(E)], f"trans{l + 1}")
            g = self.trans[f"gate{l + 1}"]
            derive = b.add(SnrGateStage(b, g, label=f"gate{l + 1}"))
            # the E expert outputs are adjacent columns of one buffer: their concatenation is free
            (mixed,) = b.add(LinearStage(b, [LinearSpec(outs[0].group.span(), derive.derived)], None,
                                         label=f"gate{l + 1}.apply")).outs
            feats = [mixed.sub(i * width, width) for i in range(g.output_dim)]
        tower = feats
        if len(self.tower_dnn_hidden_units) > 0:
            tower = mlp_stages(b, [(feats[t], self.tower_dnn[t]) for t in range(T)], "towers")
        b.add(HeadStage(b, [HeadSpec(tower[t], self.tower_dnn_final_layer[t], self.out[t].bias, self.task_types[t])
                            for t in range(T)]))
