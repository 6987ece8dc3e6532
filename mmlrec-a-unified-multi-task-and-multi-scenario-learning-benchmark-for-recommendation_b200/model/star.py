"""STAR (reference: ``/root/reference/model/star.py`` + ``SharedSpecificLinear``, ``model/utils.py:163-223``):
per domain i and layer j, ``out = x @ (W_spec[j][i] * W_shared[j]) + b_spec[j][i] + b_shared[j]`` -> ReLU; one
``SharedSpecificLinear(hidden, 1, T)`` head per domain, of which domain i uses slot i.

Reference quirk kept on purpose: the per-domain tensors live in plain Python lists and only the
LAST one (``specific_weight`` / ``specific_bias``) is a registered parameter, so the others stay at their
initial values forever and are absent from ``state_dict``.  Here they are non-persistent buffers (same
``state_dict`` keys as the reference, same RNG consumption order at construction).

Step program: a ``StarWeightStage`` materialises the effective weights of all domains once per step in
``nn.Linear`` layout, after which the ordinary grouped-GEMM stages run the layers (layer 0 of all domains is
ONE wide GEMM because every domain reads dnn_input) and the fused head kernel applies the per-domain heads;
its backward folds d(W_eff) into d(W_shared) and the registered specific tensors.
"""
import math

import torch
import torch.nn as nn

from ..engine.core import Builder, GatherStage, HeadSpec, HeadStage, LinearSpec, LinearStage, StarWeightStage
from .basemodel import BaseModel
from .utils import PredictionLayer


class SharedSpecificLinear(nn.Module):
    def __init__(self, in_features, out_features, num_domains, use_shared=True, use_bias=True, device="cpu"):
        super().__init__()
        if not (use_shared and use_bias):
            raise NotImplementedError("the fused STAR step implements use_shared=True with biases (reference default)")
        self.in_features, self.out_features, self.num_domains = in_features, out_features, num_domains
        self.shared_weight = nn.Parameter(torch.empty((in_features, out_features)))
        self.shared_bias = nn.Parameter(torch.empty(out_features))
        self.shared_weight._mm_nopad = True
        nn.init.kaiming_uniform_(self.shared_weight, a=math.sqrt(5))
        fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.shared_weight)
        bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
        nn.init.uniform_(self.shared_bias, -bound, bound)
        for i in range(num_domains):  # same draw order as the reference loop (weight, bias per domain)
            w = torch.empty((in_features, out_features))
            bvec = torch.empty(out_features)
            nn.init.kaiming_uniform_(w, a=math.sqrt(5))
            nn.init.uniform_(bvec, -bound, bound)
            if i == num_domains - 1:
                self.specific_weight, self.specific_bias = nn.Parameter(w), nn.Parameter(bvec)
                self.specific_weight._mm_nopad = True   # dense [in,out]: read by the STAR weight kernels only
            else:
                self.register_buffer(f"frozen_weight_{i}", w, persistent=False)
                self.register_buffer(f"frozen_bias_{i}", bvec, persistent=False)

    def spec_weights(self):
        return [getattr(self, f"frozen_weight_{i}") for i in range(self.num_domains - 1)] + [self.specific_weight]

    def spec_biases(self):
        return [getattr(self, f"frozen_bias_{i}") for i in range(self.num_domains - 1)] + [self.specific_bias]


class STAR(BaseModel):
    def __init__(self, dnn_feature_columns, init_std=0.0001, device="cpu", gpus=None, config=None):
        super().__init__(linear_feature_columns=[], dnn_feature_columns=dnn_feature_columns, init_std=init_std,
                         device=device, gpus=gpus, config=config)
        mc = self.model_config
        self.dnn_use_bn = mc.get("dnn_use_bn", False)  # DomainBatchNorm only runs with a domain mask: never (Q4)
        self.dnn_hidden_units = mc.get("dnn_hidden_units", [256, 128])
        if mc.get("dnn_activation", "relu") != "relu":
            raise NotImplementedError("STAR: only relu is fused")
        use_shared = mc.get("use_shared", True)
        dims = [self.compute_input_dim(dnn_feature_columns)] + list(self.dnn_hidden_units)
        T = self.num_tasks
        self.linears = nn.ModuleList(SharedSpecificLinear(dims[i], dims[i + 1], T, use_shared=use_shared)
                                     for i in range(len(dims) - 1))
        self.final_layers = nn.ModuleList(SharedSpecificLinear(dims[-1], 1, T, use_shared=use_shared) for _ in range(T))
        self.out = nn.ModuleList(PredictionLayer(task) for task in self.task_types)
        self._finalize()

    def load_unregistered(self, tensors):
        """Set the frozen per-domain tensors from ``<prefix>.specific_weights.<i>`` / ``.specific_biases.<i>``
        (the names tests/golden/make_golden.py exports them under)."""
        with torch.no_grad():
            for key, value in tensors.items():
                prefix, kind, idx = key.rsplit(".", 2)
                mod = self.get_submodule(prefix)
                name = ("frozen_weight_" if kind == "specific_weights" else "frozen_bias_") + idx
                getattr(mod, name).copy_(torch.as_tensor(value))

    def build_graph(self, b: Builder) -> None:
        T, Lh = self.num_tasks, len(self.dnn_hidden_units)
        x = b.add(GatherStage(b, self)).out
        star = b.add(StarWeightStage(b, list(self.linears) + list(self.final_layers), label="star"))
        cur = [x] * T
        for j in range(Lh):
            stage = b.add(LinearStage(b, [LinearSpec(cur[t], star.linear(j, t)) for t in range(T)], "relu",
                                      label=f"star.l{j}"))
            cur = stage.outs
        heads = []
        for t in range(T):
            fin = star.linear(Lh + t, t)  # final_layers[t] evaluated for domain t: one row of its derived matrix
            heads.append(HeadSpec(cur[t], fin, fin.bias, self.task_types[t], bias2=self.out[t].bias))
            if not b.dry:
                star.set_live(Lh + t, [1 if i == t else 0 for i in range(T)])
        b.add(HeadStage(b, heads))
