"""APG -- adaptive parameter generation (reference: ``/root/reference/model/apg.py``): every layer is the low-rank chain
``x -> x @ W_nk + b_nk  (shared, [in, k])  ->  . @ W_kk(scene) + b_kk(scene)  (per-sample [k, k], generated from the
sample's scene embedding)  ->  . @ W_km + b_km  (shared, [k, out])  -> relu`` with ``k = ceil(min(in, out) / mf_k)``; one
bias-free 1-unit head per task on the last layer's output.  Same constructor, attribute names and ``state_dict`` keys
(``apg_layers.<l>.specific_weight_kk.linears.0.*``, ``.specific_bias_kk.linears.0.*``, ``.shared_weight_nk``,
``.shared_bias_nk``, ``.shared_weight_km``, ``.shared_bias_km``, ``final_layer.<t>``, ``out.<t>``).

Reference behaviour kept on purpose: the scene embedding is looked up in the list of sparse embeddings at the scene
feature's X COLUMN (apg.py:136, :158) and detached (:159), so no gradient reaches the scene table through the generated
matrices; the generating DNNs take the DNN defaults (weights N(0, 1e-4), no activation), not the model's ``init_std``.
Only the variant the model ships is built (``use_uv_shared=True, use_mf_p=False, mf_k=4``, apg.py:143-147).

Step program per layer: ONE grouped launch for the three Linear layers that read either the layer input or the (copied,
gradient-less) scene embedding -- the shared [in, k] weight is read in its stored orientation -- then the per-sample
product (``mmlrec_apg_mix_forward`` / ``_backward``: each sample streams its own k*k matrix once), then one GEMM
``relu(. @ W_km + b_km)``; bias gradients of the [K, N]-stored weights come from ``mmlrec_colsum``."""
import math

import torch
import torch.nn as nn

from ..engine.core import (ApgMixStage, Builder, ConcatStage, GatherStage, HeadSpec, HeadStage, LinearSpec, LinearStage)
from .basemodel import BaseModel
from .utils import DNN, PredictionLayer


class _StoredKN:
    """What LinearSpec reads for a weight stored [K, N] and applied as ``x @ W + b``."""

    def __init__(self, weight, bias):
        self.weight, self.bias = weight, bias


class APGLayer(nn.Module):
    def __init__(self, input_dim, output_dim, scene_emb_dim, activation="relu", generate_activation=None,
                 inner_activation=None, use_uv_shared=True, mf_k=16, use_mf_p=True, mf_p=4, device="cpu"):
        super().__init__()
        if generate_activation is not None or inner_activation is not None or not use_uv_shared or use_mf_p:
            raise NotImplementedError("only the APG variant the reference model builds is on the fused step "
                                      "(use_uv_shared=True, use_mf_p=False, no inner / generating activations)")
        if activation not in ("relu", None):
            raise NotImplementedError(f"activation {activation!r} is not supported by the fused step")
        self.activation = activation
        self.input_dim, self.output_dim = input_dim, output_dim
        min_dim = min(int(input_dim), int(output_dim))
        self.p_dim = math.ceil(float(min_dim) / float(mf_p))
        self.k_dim = k = math.ceil(float(min_dim) / float(mf_k))
        # apg.py:34-37, :54-60: same construction (= RNG) order
        self.specific_weight_kk = DNN(scene_emb_dim, [k * k], activation=None)
        self.specific_bias_kk = DNN(scene_emb_dim, [k], activation=None)
        self.shared_weight_nk = nn.Parameter(nn.init.xavier_uniform_(torch.empty(input_dim, k)))
        self.shared_bias_nk = nn.Parameter(torch.zeros(k))
        self.shared_weight_km = nn.Parameter(nn.init.xavier_uniform_(torch.empty(k, output_dim)))
        self.shared_bias_km = nn.Parameter(torch.zeros(output_dim))


class APG(BaseModel):
    def __init__(self, dnn_feature_columns, init_std=0.0001, device="cpu", gpus=None, config=None):
        super().__init__(linear_feature_columns=[], dnn_feature_columns=dnn_feature_columns, init_std=init_std,
                         device=device, gpus=gpus, config=config)
        mc = self.model_config
        self.dnn_use_bn = mc.get("dnn_use_bn", False)
        self.dnn_hidden_units = mc.get("dnn_hidden_units", [256, 128])
        scene_emb_dim = mc.get("emb", 8)
        scene_feature = self.data_config.get("scene_feature", "")
        if scene_feature == "":
            raise ValueError("APG needs data_config['scene_feature'] (the reference's forward fails without it, apg.py:158)")
        if len(dnn_feature_columns) == 0:
            raise ValueError("dnn_feature_columns is null!")
        self.scene_index = self.feature_index[scene_feature]
        dims = [self.compute_input_dim(dnn_feature_columns)] + list(self.dnn_hidden_units)
        self.apg_layers = nn.ModuleList(
            APGLayer(input_dim=dims[i], output_dim=dims[i + 1], scene_emb_dim=scene_emb_dim,
                     activation=mc.get("dnn_activation", "relu"), use_uv_shared=True, use_mf_p=False, mf_k=4, mf_p=4)
            for i in range(len(self.dnn_hidden_units)))
        self.final_layer = nn.ModuleList(nn.Linear(self.dnn_hidden_units[-1], 1, bias=False) for _ in range(self.num_tasks))
        self.out = nn.ModuleList(PredictionLayer(task) for task in self.task_types)
        self._finalize()

    def build_graph(self, b: Builder) -> None:
        x = b.add(GatherStage(b, self)).out
        # apg.py:158-159: sparse_embedding_list[scene X column], detached -> a gradient-less copy of those D columns
        n_sparse = len(self.embedding_layout)
        if not 0 <= self.scene_index[0] < n_sparse:
            raise IndexError("APG: the scene feature's X column does not index the sparse embedding list")
        scene = b.add(ConcatStage(b, [x.sub(self.scene_index[0] * self.emb_dim, self.emb_dim)], label="scene")).out
        h = x
        for l, layer in enumerate(self.apg_layers):
            wkk, bkk = layer.specific_weight_kk.linears[0], layer.specific_bias_kk.linears[0]
            nk, w, bias = b.add(LinearStage(b, [
                LinearSpec(h, _StoredKN(layer.shared_weight_nk, layer.shared_bias_nk), None, transposed=True),
                LinearSpec(scene, wkk), LinearSpec(scene, bkk)], None, label=f"apg{l}.nk", align_outs=8)).outs
            kk = b.add(ApgMixStage(b, nk, w, bias, label=f"apg{l}")).out
            (h,) = b.add(LinearStage(b, [LinearSpec(kk, _StoredKN(layer.shared_weight_km, layer.shared_bias_km), None,
                                                    transposed=True)], layer.activation, label=f"apg{l}.km")).outs
        b.add(HeadStage(b, [HeadSpec(h, self.final_layer[t], self.out[t].bias, self.task_types[t])
                            for t in range(self.num_tasks)]))
