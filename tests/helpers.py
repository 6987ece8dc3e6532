"""Shared test helpers: golden fixture loading and oracle construction (test infrastructure)."""
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith("fit_"))
# cases verified on the CPU only (oracle + plan emulation): the -m gpu step tests enumerate GOLDEN_CASES, not these
GOLDEN_CPU_ONLY = sorted(f[:-4] for f in os.listdir(os.path.join(GOLDEN, "cpu_only")) if f.endswith(".npz")) \
    if os.path.isdir(os.path.join(GOLDEN, "cpu_only")) else []


def load_golden(case):
    path = os.path.join(GOLDEN, case + ".npz")
    if not os.path.exists(path):
        path = os.path.join(GOLDEN, "cpu_only", case + ".npz")
    z = np.load(path, allow_pickle=False)
    cfg = json.loads(str(z["meta/config"]))
    fields = [tuple(f) for f in json.loads(str(z["meta/fields"]))]
    return z, cfg, fields


def oracle_columns(cfg, fields):
    from oracle.mmlrec_oracle import Column
    emb = cfg["model_config"]["emb"]
    return [Column(n, k, v, emb if k == "sparse" else 1) for n, k, v in fields]


def golden_init(z):
    buffers = set(str(s) for s in z["meta/buffers"])
    trainable = [str(s) for s in z["meta/trainable"]]
    params, bufs = {}, {}
    for k in z.files:
        if k.startswith("init/"):
            name = k[5:]
            t = torch.from_numpy(z[k].copy())
            (bufs if name in buffers else params)[name] = t
    return params, bufs, trainable


def make_oracle(case):
    from oracle.mmlrec_oracle import OracleTrainer
    z, cfg, fields = load_golden(case)
    params, bufs, trainable = golden_init(z)
    return OracleTrainer(cfg, oracle_columns(cfg, fields), params, bufs, trainable), z, cfg, fields


def rel_err(a, b):
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def build_pair(workload, kw=None, precision="fp32", init_std=1e-4, seed=0, mc_over=None, oc_over=None, cuda_graph=True):
    """The CUDA model (public API) and the CPU oracle holding the SAME seeded initial state, for a synthetic
    workload at its full (unshrunk) BASELINE shape unless ``mc_over`` shrinks it."""
    import copy
    from mmlrec_b200 import synthetic
    from mmlrec_b200.model import get_model_class
    from mmlrec_b200.model.utils import DenseFeat, SparseFeat
    from oracle.mmlrec_oracle import OracleTrainer
    cfg, fields = synthetic.workload(workload, **(kw or {}))
    cfg["model_config"].update(mc_over or {})
    cfg["optim_config"].update(oc_over or {})
    cfg["b200_config"] = {"precision": precision, "cuda_graph": cuda_graph}
    emb = cfg["model_config"]["emb"]
    cols = [SparseFeat(n, v, emb) if k == "sparse" else DenseFeat(n, 1) for n, k, v in fields]
    torch.manual_seed(seed)
    model = get_model_class(cfg["model_config"]["model_name"])(cols, init_std=init_std, device="cuda:0",
                                                               config=copy.deepcopy(cfg))
    model.compile(cfg["optim_config"]["optimizer"], cfg["optim_config"]["loss"], [])
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    names = [n for n, _ in model.named_parameters()]
    params = {k: v for k, v in sd.items() if k in names}
    if hasattr(model, "linears") and hasattr(model.linears[0], "spec_weights"):  # STAR's frozen per-domain tensors
        for prefix, mods in (("linears", model.linears), ("final_layers", model.final_layers)):
            for j, m in enumerate(mods):
                for i in range(m.num_domains - 1):
                    params[f"{prefix}.{j}.specific_weights.{i}"] = getattr(m, f"frozen_weight_{i}").detach().cpu().clone()
                    params[f"{prefix}.{j}.specific_biases.{i}"] = getattr(m, f"frozen_bias_{i}").detach().cpu().clone()
    oracle = OracleTrainer(cfg, oracle_columns(cfg, fields), params, {k: v for k, v in sd.items() if k not in names}, names)
    model.train()
    return model, oracle, cfg, fields, sd
