"""Shared test helpers: golden fixture loading and oracle construction (test infrastructure)."""
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))


def load_golden(case):
    z = np.load(os.path.join(GOLDEN, case + ".npz"), allow_pickle=False)
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
