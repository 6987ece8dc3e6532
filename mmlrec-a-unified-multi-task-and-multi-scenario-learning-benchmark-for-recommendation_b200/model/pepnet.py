"""PEPNet (reference: ``/root/reference/model/pepnet.py``): EPNet feature gate
``x <- GateNN(cat[x.detach(), scene_emb]) * x`` followed by one PPNet block per task in which every layer's
input is scaled by its own ``GateNN(cat[x.detach(), task_emb])`` (``GateNN = Linear -> ReLU -> Linear ->
Sigmoid, times 2``).  Linear layers keep PyTorch's default init, like the reference.

Step program: the first Linear of EVERY GateNN of every task reads the same detached concat -> one wide GEMM;
the second Linears are one grouped launch (2*sigmoid epilogue); per layer one element-wise product launch per
task and one grouped GEMM; the last Linear(H,1) + its bias + the PredictionLayer bias are the fused head.
"""
import torch.nn as nn

from ..engine.core import (Builder, ConcatStage, GatherStage, HeadSpec, HeadStage, LinearSpec, LinearStage, MulStage)
from .basemodel import BaseModel
from .utils import PredictionLayer


class GateNN(nn.Module):
    def __init__(self, input_dim, hidden_dim=None, output_dim=None, hidden_activation="relu", dropout_rate=0.0,
                 batch_norm=False, device="cpu"):
        super().__init__()
        if batch_norm or dropout_rate > 0 or hidden_activation != "relu":
            raise NotImplementedError("GateNN: only the reference defaults are fused")
        hidden_dim = output_dim if hidden_dim is None else hidden_dim
        self.gate = nn.Sequential(nn.Linear(input_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, output_dim),
                                  nn.Sigmoid())


class PPNetBlock(nn.Module):
    def __init__(self, input_dim, output_dim=1, gate_input_dim=32, gate_hidden_dim=None, hidden_units=(),
                 hidden_activations="relu", dropout_rates=0.0, batch_norm=False, use_bias=True, device="cpu"):
        super().__init__()
        if batch_norm or dropout_rates or hidden_activations != "relu" or not use_bias or output_dim != 1:
            raise NotImplementedError("PPNetBlock: only the reference defaults are fused")
        dims = [input_dim] + list(hidden_units)
        self.gate_layers, self.mlp_layers = nn.ModuleList(), nn.ModuleList()
        for i in range(len(dims) - 1):  # creation order = RNG order of the reference
            dense = nn.Linear(dims[i], dims[i + 1], bias=True)
            self.gate_layers.append(GateNN(gate_input_dim, gate_hidden_dim, output_dim=dims[i]))
            self.mlp_layers.append(nn.Sequential(dense, nn.ReLU()))
        self.gate_layers.append(GateNN(gate_input_dim, gate_hidden_dim, output_dim=dims[-1]))
        self.mlp_layers.append(nn.Linear(dims[-1], output_dim, bias=True))


class PepNet(BaseModel):
    def __init__(self, dnn_feature_columns, init_std=0.0001, device="cpu", gpus=None, config=None):
        super().__init__(linear_feature_columns=[], dnn_feature_columns=dnn_feature_columns, init_std=init_std,
                         device=device, gpus=gpus, config=config)
        mc, dc = self.model_config, self.data_config
        self.dnn_hidden_units = mc.get("dnn_hidden_units", [256, 128])
        emb = mc.get("emb", 8)
        self.scene_feature = dc.get("scene_feature", "")
        self.user_sf, self.item_sf = dc.get("user_sf", ""), dc.get("item_sf", "")
        if self.scene_feature == "":
            raise ValueError("PepNet needs data_config.scene_feature")
        self.side_features = [self.scene_feature]
        if self.user_sf != "" and self.item_sf != "":
            self.side_features += [self.user_sf, self.item_sf]
        task_dim = emb * (1 + (self.user_sf != "") + (self.item_sf != ""))
        input_dim = self.compute_input_dim(dnn_feature_columns)
        self.feature_gate = GateNN(input_dim=input_dim + emb, hidden_dim=128, output_dim=input_dim)
        self.ppn = nn.ModuleList(PPNetBlock(input_dim=input_dim, output_dim=1, gate_input_dim=input_dim + task_dim,
                                            gate_hidden_dim=None, hidden_units=list(self.dnn_hidden_units))
                                 for _ in range(self.num_tasks))
        self.out = nn.ModuleList(PredictionLayer(task) for task in self.task_types)
        self._finalize()

    def _field_act(self, x, name):
        # the reference indexes the list of sparse embeddings by the feature's X column (pepnet.py:126); with the
        # sparse columns first that is the field's position, i.e. columns [j*D, (j+1)*D) of dnn_input
        j = self.feature_index[name][0]
        return x.sub(j * self.emb_dim, self.emb_dim)

    def build_graph(self, b: Builder) -> None:
        T, Lh = self.num_tasks, len(self.dnn_hidden_units)
        x = b.add(GatherStage(b, self)).out
        scene = self._field_act(x, self.scene_feature)
        side = [self._field_act(x, n) for n in self.side_features]
        # EPNet: feature gate on the detached input
        g0 = b.add(ConcatStage(b, [x, scene], label="epnet_in")).out
        fg_h = b.add(LinearStage(b, [LinearSpec(g0, self.feature_gate.gate[0])], "relu", label="epnet.l0")).outs[0]
        fg = b.add(LinearStage(b, [LinearSpec(fg_h, self.feature_gate.gate[2])], "sigmoid2", label="epnet.l1")).outs[0]
        x1 = b.add(MulStage(b, [(fg, x)], label="epnet")).outs[0]
        # PPNet: all GateNNs read cat[x1.detach(), task embedding]
        gin = b.add(ConcatStage(b, [x1] + side, label="ppnet_in")).out
        gates = [(t, i) for t in range(T) for i in range(Lh + 1)]
        gh = b.add(LinearStage(b, [LinearSpec(gin, self.ppn[t].gate_layers[i].gate[0]) for t, i in gates], "relu",
                               label="ppnet.gate0")).outs
        gw = b.add(LinearStage(b, [LinearSpec(gh[k], self.ppn[t].gate_layers[i].gate[2]) for k, (t, i) in enumerate(gates)],
                               "sigmoid2", label="ppnet.gate1")).outs
        gw = {ti: a for ti, a in zip(gates, gw)}
        hidden = [x1] * T
        for i in range(Lh):
            scaled = b.add(MulStage(b, [(hidden[t], gw[(t, i)]) for t in range(T)], label=f"ppnet.mul{i}")).outs
            hidden = b.add(LinearStage(b, [LinearSpec(scaled[t], self.ppn[t].mlp_layers[i][0]) for t in range(T)], "relu",
                                       label=f"ppnet.l{i}")).outs
        scaled = b.add(MulStage(b, [(hidden[t], gw[(t, Lh)]) for t in range(T)], label=f"ppnet.mul{Lh}")).outs
        b.add(HeadStage(b, [HeadSpec(scaled[t], self.ppn[t].mlp_layers[Lh], self.ppn[t].mlp_layers[Lh].bias,
                                     self.task_types[t], bias2=self.out[t].bias) for t in range(T)]))
