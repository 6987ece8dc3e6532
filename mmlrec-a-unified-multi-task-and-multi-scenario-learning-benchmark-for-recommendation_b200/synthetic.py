"""Synthetic workloads with the shapes of the BASELINE.json configs.

The reference derives every shape from a CSV (``utils/data_utils.py:14-94``): vocabulary sizes are
``max id + 1``, sparse columns come first, dense columns are MinMax-scaled to [0,1].  There is no
dataset in this environment, so each workload here is a JSON-compatible ``config`` dict (same
sections and keys the reference reads) plus the list of feature columns the data pipeline would
have produced, plus a seeded generator for ``X`` / ``y`` of that shape.
"""
from __future__ import annotations

import copy
from typing import Dict, List, Optional, Tuple

import numpy as np

# (name, kind, vocabulary) ; kind in {"sparse", "dense"}
FieldSpec = Tuple[str, str, int]

_CENSUS_SPARSE = [
    ("class_worker", 9), ("det_ind_code", 52), ("det_occ_code", 47), ("education", 17), ("hs_college", 3),
    ("major_ind_code", 24), ("major_occ_code", 15), ("race", 5), ("hisp_origin", 10), ("sex", 2),
    ("union_member", 3), ("unemp_reason", 6), ("full_or_part_emp", 8), ("tax_filer_stat", 6),
    ("region_prev_res", 6), ("state_prev_res", 51), ("det_hh_fam_stat", 38), ("det_hh_summ", 8),
    ("mig_chg_msa", 10), ("mig_chg_reg", 9), ("mig_move_reg", 10), ("mig_same", 3), ("mig_prev_sunbelt", 4),
    ("fam_under_18", 5), ("country_father", 43), ("country_mother", 43), ("country_self", 43),
    ("citizenship", 5), ("vet_question", 3),
]
_CENSUS_DENSE = ["age", "wage_per_hour", "capital_gains", "capital_losses", "stock_dividends", "instance_weight",
                 "num_emp", "own_or_self", "vet_benefits", "weeks_worked", "year"]


def _base_config(task_name: str, model_name: str, emb: int, labels: List[str], task_names: List[str]) -> dict:
    t = len(labels)
    return {
        "data_config": {"data_name": "synthetic", "feature_columns": [], "dense_columns": [], "ignore_columns": [],
                        "label_columns": labels},
        "model_config": {"task_name": task_name, "model_name": model_name, "task": "binary", "emb": emb,
                         "task_names": task_names, "task_types": ["binary"] * t,
                         "num_experts": 4, "shared_expert_num": 2, "specific_expert_num": 3, "num_levels": 2,
                         "l2_reg_linear": 0.0, "l2_reg_embedding": 0.0, "l2_reg_dnn": 0.0,
                         "dnn_use_bn": False, "dnn_dropout": 0.0, "dnn_activation": "relu", "use_cka_loss": False},
        "optim_config": {"lr": 1e-3, "optimizer": "adam", "loss": ["binary_crossentropy"] * t,
                         "metrics": ["auc", "acc"], "early_stop": 3},
        "training_config": {"train_batch_size": 4096, "test_batch_size": 4096, "epochs": 1},
        "save_config": {"save_layer_output": False},
    }


def _finish(cfg: dict, fields: List[FieldSpec]) -> Tuple[dict, List[FieldSpec]]:
    cfg["data_config"]["feature_columns"] = [n for n, k, _ in fields if k == "sparse"]
    cfg["data_config"]["dense_columns"] = [n for n, k, _ in fields if k == "dense"]
    return cfg, fields


def census(model_name: str = "mmoe", vocab_scale: float = 1.0) -> Tuple[dict, List[FieldSpec]]:
    """C1: Census-Income MTL shape (configs_mtl/config_census.json): 29 sparse + 11 dense, emb 4,
    experts/gates/towers [128], BatchNorm on, Adam 3e-4, 2 tasks."""
    cfg = _base_config("mtl", model_name, 4, ["label_income", "label_marital"], ["income", "marital"])
    cfg["model_config"].update(expert_dnn_hidden_units=[128], gate_dnn_hidden_units=[128],
                               tower_dnn_hidden_units=[128], bottom_dnn_hidden_units=[128],
                               dnn_hidden_units=[128], dnn_use_bn=True)
    cfg["optim_config"].update(lr=3e-4, early_stop=5)
    fields = [(n, "sparse", max(2, int(v * vocab_scale))) for n, v in _CENSUS_SPARSE]
    fields += [(n, "dense", 0) for n in _CENSUS_DENSE]
    return _finish(cfg, fields)


def aliexpress(model_name: str = "ple", num_domains: int = 4, max_vocab: int = 1_000_000) -> Tuple[dict, List[FieldSpec]]:
    """C2: AliExpress MSL shape (configs_msl/config_AE.json): 16 sparse + appended scene column,
    63 dense, emb 8, PLE 2 levels (3 specific + 2 shared), experts [256,128], gates/towers [64],
    Adam 5e-3.  ``num_domains=4`` is the BASELINE.json override, 2 is the shipped file."""
    cfg = _base_config("msl", model_name, 8, ["label"] * num_domains, ["ctr"] * num_domains)
    cfg["data_config"].update(num_domains=num_domains, mask_values=list(range(num_domains)),
                              mask_column="scene", scene_feature="scene")
    cfg["model_config"].update(expert_dnn_hidden_units=[256, 128], dnn_hidden_units=[256, 128, 64],
                               bottom_dnn_hidden_units=[256, 128], gate_dnn_hidden_units=[64],
                               tower_dnn_hidden_units=[64])
    cfg["optim_config"].update(lr=5e-3)
    names = [f"c{i}" for i in range(1, 10)] + ["iin18", "iin33", "iin34", "iin35", "iin39", "iin40", "iin41"]
    lo = 100.0
    fields: List[FieldSpec] = []
    for j, n in enumerate(names):  # log-spaced 1e2 .. max_vocab
        v = int(round(lo * (max(max_vocab, lo) / lo) ** (j / (len(names) - 1))))
        fields.append((n, "sparse", max(2, v)))
    fields.append(("scene", "sparse", num_domains))
    dense = [f"n{i}" for i in range(1, 24)] + [f"in{i}" for i in
                                               list(range(1, 18)) + list(range(19, 33)) + [36, 37, 38] + list(range(42, 48))]
    assert len(dense) == 63
    fields += [(n, "dense", 0) for n in dense]
    return _finish(cfg, fields)


def kuairec(model_name: str = "sharedbottom", max_vocab: int = 10_728) -> Tuple[dict, List[FieldSpec]]:
    """C3: KuaiRec MTL shape (configs_mtl/config_kuairec.json): 32 sparse, no dense, emb 8,
    bottom/expert [512,256], gate/tower [128], Adam 1e-3, 2 tasks."""
    cfg = _base_config("mtl", model_name, 8, ["label", "label2"], ["ctr", "ctr"])
    cfg["model_config"].update(expert_dnn_hidden_units=[512, 256], dnn_hidden_units=[512, 256],
                               bottom_dnn_hidden_units=[512, 256], gate_dnn_hidden_units=[128],
                               tower_dnn_hidden_units=[128])
    cycle = [max_vocab, 7176, 2, 8, 31, 2, 128, 9, 512, 2, 2048, 50, 7, 2, 340, 1000]
    fields = [(f"f{j}", "sparse", max(2, min(cycle[j % len(cycle)], max_vocab))) for j in range(32)]
    return _finish(cfg, fields)


def movielens(model_name: str = "star", vocab_scale: float = 1.0) -> Tuple[dict, List[FieldSpec]]:
    """C4: MovieLens MTMSL shape (configs_mtmsl/config_movielens.json): 7 sparse, emb 8, hidden
    [128,128], no gate/tower DNN, 2 domains x 2 labels = 4 heads, scene = gender_tag, Adam 1e-2."""
    labels = ["label2", "label2", "label3", "label3"]
    cfg = _base_config("mtmsl", model_name, 8, labels, ["ctr", "ctr", "ctcvr", "ctcvr"])
    cfg["data_config"].update(num_domains=2, mask_values=[0, 1], mask_column="gender_tag", scene_feature="gender_tag")
    cfg["model_config"].update(expert_dnn_hidden_units=[128, 128], dnn_hidden_units=[128, 128],
                               bottom_dnn_hidden_units=[128, 128], gate_dnn_hidden_units=[],
                               tower_dnn_hidden_units=[])
    cfg["optim_config"].update(lr=1e-2, early_stop=5)
    fields = [("user_tag", "sparse", 6041), ("movie_tag", "sparse", 3884), ("gender_tag", "sparse", 2),
              ("age_tag", "sparse", 7), ("occupation_tag", "sparse", 21), ("zip_tag", "sparse", 3440),
              ("genres_tag", "sparse", 302)]
    fields = [(n, k, v if n == "gender_tag" else max(2, int(v * vocab_scale))) for n, k, v in fields]
    return _finish(cfg, fields)


def synth26(model_name: str = "mmoe", vocab: int = 10_000_000, emb: int = 8, optimizer: str = "adagrad"
            ) -> Tuple[dict, List[FieldSpec]]:
    """C5: 26 sparse fields x ``vocab`` rows, KuaiRec MLP dims, Adagrad headline (exactly sparse)."""
    cfg = _base_config("mtl", model_name, emb, ["label", "label2"], ["ctr", "cvr"])
    cfg["model_config"].update(expert_dnn_hidden_units=[512, 256], gate_dnn_hidden_units=[128],
                               tower_dnn_hidden_units=[128], bottom_dnn_hidden_units=[512, 256])
    cfg["optim_config"].update(lr=1e-2, optimizer=optimizer)
    fields = [(f"s{j}", "sparse", vocab) for j in range(26)]
    return _finish(cfg, fields)


WORKLOADS = {
    "census_mmoe": lambda **k: census("mmoe", **k),
    "ae_ple_t4": lambda **k: aliexpress("ple", 4, **k),
    "ae_ple_t2": lambda **k: aliexpress("ple", 2, **k),
    "kuairec_sharedbottom": lambda **k: kuairec("sharedbottom", **k),
    "kuairec_esmm": lambda **k: kuairec("esmm", **k),
    "movielens_star": lambda **k: movielens("star", **k),
    "movielens_pepnet": lambda **k: movielens("pepnet", **k),
    "synth26_mmoe": lambda **k: synth26("mmoe", **k),
}


_DATASETS = {"census": census, "ae_t4": lambda m, **k: aliexpress(m, 4, **k), "ae_t2": lambda m, **k: aliexpress(m, 2, **k),
             "kuairec": kuairec, "movielens": movielens, "synth26": synth26}


def workload(name: str, **kw) -> Tuple[dict, List[FieldSpec]]:
    """A named BASELINE workload, or ``"<dataset>:<model>"`` -- any model of the zoo on a dataset shape, e.g.
    ``"movielens:apg"``, ``"kuairec:aitm"``, ``"kuairec:mssm"`` (datasets: census, ae_t4, ae_t2, kuairec, movielens, synth26)."""
    if name not in WORKLOADS and ":" in name:
        dataset, model_name = name.split(":", 1)
        if dataset not in _DATASETS:
            raise KeyError(f"unknown dataset shape {dataset!r} (have {sorted(_DATASETS)})")
        cfg, fields = _DATASETS[dataset](model_name, **kw)
    else:
        cfg, fields = WORKLOADS[name](**kw)
    return copy.deepcopy(cfg), list(fields)


def num_tasks(cfg: dict) -> int:
    """model/basemodel.py:97-102."""
    mc, dc = cfg["model_config"], cfg["data_config"]
    if mc.get("task_name", "mtl") == "msl":
        return dc.get("num_domains", 1)
    if mc.get("task_name", "mtl") == "mtmsl":
        return len(dc["label_columns"])
    return len(mc.get("task_names", ["ctr", "ctcvr"]))


def make_batch(cfg: dict, fields: List[FieldSpec], n: int, seed: int = 0, zipf: Optional[float] = None,
               label_rate: float = 0.3) -> Tuple[np.ndarray, np.ndarray]:
    """``X [n, F] float32`` (ids carried as fp32 exactly like ``basemodel.py:262``; dense ~ U[0,1))
    and ``y [n, T] float32`` (Bernoulli; columns that share a label name share the draw, which is
    what ``df.reindex`` with duplicated label columns produces)."""
    rng = np.random.default_rng(seed)
    cols = []
    for name, kind, vocab in fields:
        if kind == "sparse":
            if zipf is None:
                ids = rng.integers(0, vocab, size=n)
            else:
                ids = np.minimum(rng.zipf(zipf, size=n) - 1, vocab - 1)
            cols.append(ids.astype(np.float32))
        else:
            cols.append(rng.random(n, dtype=np.float32))
    X = np.stack(cols, axis=1)
    labels = cfg["data_config"]["label_columns"]
    draws: Dict[str, np.ndarray] = {}
    ys = []
    for lab in labels:
        if lab not in draws:
            draws[lab] = (rng.random(n) < label_rate).astype(np.float32)
        ys.append(draws[lab])
    T = num_tasks(cfg)
    while len(ys) < T:  # msl: num_domains heads on the single label
        ys.append(ys[-1])
    return X, np.stack(ys[:T], axis=1)
