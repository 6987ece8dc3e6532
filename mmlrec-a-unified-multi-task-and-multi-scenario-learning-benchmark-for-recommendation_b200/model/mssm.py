"""MSSM (reference: ``/root/reference/model/mssm.py``): SNR-trans's layout -- levels of E single-layer experts, after each
level a ``gate`` mixing the E expert outputs into E (inner levels) or T (last level) features through one [U, U]
transformation ``M_ij`` per connection -- with the hard-concrete gate applied PER OUTPUT UNIT:
``out_i = sum_j (x_j @ M_ij) * z_ij``, ``z_ij`` a vector of U gates.  Same constructor, attribute names and
``state_dict`` keys (``mssm.expert<l>.<j>``, ``mssm.gate<l>.alpha``, ``tower_dnn.i``, ``tower_dnn_final_layer.i``,
``out.i``).

Reference quirk kept on purpose: BOTH the gate logits ``u`` and the transformation matrices live in plain Python lists
of lists (mssm.py:26-36), so neither is registered, moved by the optimizer or present in ``state_dict`` -- the only
trained quantity of a gate is the scalar ``alpha``.  Here they are two non-persistent buffers per gate, filled by the
same RNG calls in the same order.

Step program: the derived-weight stage of SNR-trans with ``zdim = U``
(``W_eff[i*U+v, j*U+u] = z_ij[v] M_ij[u, v]``, ``mmlrec_snr_gate_weights``), ONE GEMM [B, E*U] x [E*U, n_out*U] over the
adjacent expert outputs, and the fold of d(W_eff) into d(alpha) in backward."""
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
        u = torch.empty(output_dim, input_dim, units)
        for i in range(output_dim):   # mssm.py:26-29: one uniform_ call per connection, i-major
            for j in range(input_dim):
                nn.init.uniform_(u[i, j], self.e, 1 - self.e)
        m = torch.empty(output_dim, input_dim, units, units)
        for i in range(output_dim):   # mssm.py:31-36
            for j in range(input_dim):
                nn.init.xavier_normal_(m[i, j])
        self.register_buffer("u", u, persistent=False)
        self.register_buffer("trans_matrix", m, persistent=False)


class MSSM(BaseModel):
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
        self.mssm = nn.ModuleDict()
        for l, width in enumerate(units):   # mssm.py:85-119: expert<l+1> then gate<l+1>
            self.mssm[f"expert{l + 1}"] = nn.ModuleList(DNN(self.hidden_units[l], [width], **kw) for _ in range(E))
            self.mssm[f"gate{l + 1}"] = gate(E, T if l == len(units) - 1 else E, width)
        tower_in = units[-1]
        if len(self.tower_dnn_hidden_units) > 0:
            self.tower_dnn = nn.ModuleList(DNN(tower_in, self.tower_dnn_hidden_units, **kw) for _ in range(T))
            tower_in = self.tower_dnn_hidden_units[-1]
        self.tower_dnn_final_layer = nn.ModuleList(nn.Linear(tower_in, 1, bias=False) for _ in range(T))
        self.out = nn.ModuleList(PredictionLayer(task) for task in self.task_types)
        # mssm.py:127-129: only the towers are registered for L2
        self.regularize([self.tower_dnn] if len(self.tower_dnn_hidden_units) > 0 else [], mc.get("l2_reg_dnn", 0))
        self._finalize()

    def load_unregistered(self, tensors):
        """Set the constant gate tensors from ``mssm.gate<l>.trans_matrix.<i>.<j>`` / ``mssm.gate<l>.u.<i>.<j>`` (the
        names tests/golden/make_golden.py exports the reference's unregistered tensors under)."""
        with torch.no_grad():
            for key, value in tensors.items():
                prefix, i, j = key.rsplit(".", 2)
                mod_name, attr = prefix.rsplit(".", 1)
                assert attr in ("trans_matrix", "u"), key
                getattr(self.get_submodule(mod_name), attr)[int(i), int(j)].copy_(torch.as_tensor(value))

    def build_graph(self, b: Builder) -> None:
        T, E = self.num_tasks, self.num_experts
        x = b.add(GatherStage(b, self)).out
        feats = [x] * E
        for l, width in enumerate(self.expert_dnn_hidden_units):
            outs = mlp_stages(b, [(feats[j], self.mssm[f"expert{l + 1}"][j]) for j in range(E)], f"expert{l + 1}")
            g = self.mssm[f"gate{l + 1}"]
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
