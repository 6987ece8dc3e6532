"""GPU: the Python boundary end to end: ``fit`` (epochs, validation, best-epoch copy), ``evaluate``,
``predict`` and the repo's ``main.py`` flow, against what the REFERENCE's own ``fit()`` produced on the same seeded
synthetic mtl dataset (tests/golden/fit_sharedbottom_kuairec.npz, made by tests/golden/make_fit_golden.py)."""
import contextlib
import io
import json
import os
import re
import sys

import numpy as np
import pandas as pd
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import GOLDEN  # noqa: E402


def _load():
    z = np.load(os.path.join(GOLDEN, "fit_sharedbottom_kuairec.npz"), allow_pickle=False)
    cfg, fields = json.loads(str(z["meta/config"])), [tuple(f) for f in json.loads(str(z["meta/fields"]))]
    args = json.loads(str(z["meta/args"]))
    return z, cfg, fields, args


def _parse(text):
    epochs, evals = [], []
    for line in text.splitlines():
        line = line.strip()
        if line.startswith("{") and "auc" in line:
            evals.append({k: float(v) for k, v in re.findall(r"'(\w+)': (?:np\.float64\()?([-0-9.e]+)", line)})
        m = re.match(r"\d+s - (.*)", line)
        if m:
            epochs.append({k.strip(): float(v) for k, v in (kv.split(":") for kv in m.group(1).split(" - "))})
    return epochs, evals


def _build(cfg, fields, args, extra=None):
    import copy
    from mmlrec_b200.model import get_model_class
    from mmlrec_b200.model.utils import DenseFeat, SparseFeat
    cfg = copy.deepcopy(cfg)
    cfg["b200_config"] = dict({"precision": "fp32"}, **(extra or {}))
    emb = cfg["model_config"]["emb"]
    cols = [SparseFeat(n, v, emb) if k == "sparse" else DenseFeat(n, 1) for n, k, v in fields]
    torch.manual_seed(args["seed"])
    np.random.seed(args["seed"])
    model = get_model_class(cfg["model_config"]["model_name"])(cols, init_std=args["init_std"], device="cuda:0", config=cfg)
    model.compile(cfg["optim_config"]["optimizer"], cfg["optim_config"]["loss"], cfg["optim_config"]["metrics"])
    return model


@pytest.mark.parametrize("streaming", [False, True])
def test_fit_evaluate_predict_track_the_reference(streaming):
    """Same seed -> same initial weights and the same DataLoader permutation as the reference, so the per-epoch log
    line (loss, batch-averaged train AUC / accuracy, validation AUC / accuracy) and the best model's predictions must
    agree.  ``streaming``: the dataset is kept in pinned host memory and fed through the double-buffered H2D path."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    z, cfg, fields, args = _load()
    names = [f[0] for f in fields]
    train_df, val_df = pd.DataFrame(z["train/X"], columns=names), pd.DataFrame(z["val/X"], columns=names)
    model = _build(cfg, fields, args, {"max_resident_bytes": 0} if streaming else None)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        best = model.fit({n: train_df[n] for n in names}, z["train/y"], batch_size=args["batch"], epochs=args["epochs"],
                         validation_data=({n: val_df[n] for n in names}, z["val/y"]))
    epochs, evals = _parse(buf.getvalue())
    want_epochs, want_evals = json.loads(str(z["log/epochs"])), json.loads(str(z["log/evals"]))
    assert len(epochs) == len(want_epochs) == args["epochs"], buf.getvalue()
    for e, (got, want) in enumerate(zip(epochs, want_epochs)):
        assert set(got) == set(want), (got, want)
        for k in want:
            assert abs(got[k] - want[k]) <= 2e-4 + 1e-3 * abs(want[k]), f"epoch {e + 1} {k}: {got[k]} vs reference {want[k]}"
    for e, (got, want) in enumerate(zip(evals, want_evals)):
        assert abs(got["auc"] - want["auc"]) < 1e-3 and abs(got["acc"] - want["acc"]) < 5e-3, (e, got, want)
    assert best is not model, "fit returns the copy taken at the best validation AUC"
    final = json.loads(str(z["final/eval"]))
    res = best.evaluate([val_df[n].values for n in names], z["val/y"], args["batch"])
    assert abs(res["auc"] - final["auc"]) < 1e-3 and abs(res["acc"] - final["acc"]) < 5e-3
    pred = best.predict([val_df[n].values for n in names], args["batch"])
    assert pred.dtype == np.float64 and pred.shape == z["final/pred"].shape
    assert float(np.abs(pred - z["final/pred"]).max()) < 2e-3


def test_main_py_runs_end_to_end_on_a_csv(tmp_path):
    """The repo's main.py (the reference's CLI, with its three NameErrors fixed) on a synthetic CSV: dataset build,
    model factory, compile, fit, predict, per-task LogLoss / AUC row appended to the results file."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    z, cfg, fields, args = _load()
    names = [f[0] for f in fields]
    for part in ("train", "val"):
        df = pd.DataFrame(z[f"{part}/X"], columns=names)
        df["label"], df["label2"] = z[f"{part}/y"][:, 0].astype(int), z[f"{part}/y"][:, 1].astype(int)
        df.to_csv(tmp_path / f"{part}.csv", index=False)
    cfg["data_config"].update(all_columns=names + ["label", "label2"], train_dataset_path=str(tmp_path / "train.csv"),
                              test_dataset_path=str(tmp_path / "val.csv"), test_result_path=str(tmp_path / "res.csv"))
    cfg["training_config"].update(train_batch_size=256, test_batch_size=256, epochs=2)
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(root, "main.py"), "--config", str(tmp_path / "config.json"), "--seed", "2",
                          "--device", "cuda:0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "test AUC" in out.stdout and "Epoch 2/2" in out.stdout
    res = pd.read_csv(tmp_path / "res.csv")
    assert {"auc_0", "auc_1", "log_loss_0", "log_loss_1"} <= set(res.columns) and len(res) == 1
    assert 0.5 < float(res["auc_0"][0]) <= 1.0


def test_fit_with_domain_mask_follows_the_masked_oracle():
    """``b200_config["domain_mask"]``: fit() builds the mask from ``data_config["mask_column"]`` like the reference's fit()
    (basemodel.py:152-161) and -- unlike the reference, whose loop then drops it (:265-266) -- applies it.  One epoch
    without shuffling must give the epoch loss of the oracle stepping the same batches with the same masks, and
    predict(domain_mask) must zero the other scenarios' heads."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import copy
    from helpers import oracle_columns
    from mmlrec_b200 import synthetic
    from mmlrec_b200.model import get_model_class
    from mmlrec_b200.model.utils import DenseFeat, SparseFeat, get_mask
    from oracle.mmlrec_oracle import OracleTrainer
    cfg, fields = synthetic.workload("ae_ple_t4", max_vocab=200)
    cfg["model_config"].update(expert_dnn_hidden_units=[16, 8], gate_dnn_hidden_units=[8], tower_dnn_hidden_units=[8])
    cfg["b200_config"] = {"precision": "fp32", "domain_mask": True}
    emb = cfg["model_config"]["emb"]
    cols = [SparseFeat(n, v, emb) if k == "sparse" else DenseFeat(n, 1) for n, k, v in fields]
    torch.manual_seed(3)
    model = get_model_class("ple")(cols, init_std=0.05, device="cuda:0", config=copy.deepcopy(cfg))
    model.compile(cfg["optim_config"]["optimizer"], cfg["optim_config"]["loss"], ["auc"])
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    names = [n for n, _ in model.named_parameters()]
    oracle = OracleTrainer(cfg, oracle_columns(cfg, fields), {k: v for k, v in sd.items() if k in names},
                           {k: v for k, v in sd.items() if k not in names}, names)
    N, B = 512, 128
    X, y = synthetic.make_batch(cfg, fields, N, seed=77)
    cols_names = [f[0] for f in fields]
    dc = cfg["data_config"]
    dm = get_mask(list(X[:, cols_names.index(dc["mask_column"])]), dc["mask_values"], dc["num_domains"]).float()
    total = 0.0
    for a in range(0, N, B):
        _, loss = oracle.step(torch.from_numpy(X[a:a + B]), torch.from_numpy(y[a:a + B]), dm[a:a + B])
        total += float(loss)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        model.fit({n: X[:, i] for i, n in enumerate(cols_names)}, y, batch_size=B, epochs=1, shuffle=False)
    epochs, _ = _parse(buf.getvalue())
    assert abs(epochs[0]["loss"] - total / N) <= 1e-4 * abs(total / N), (epochs, total / N)
    pred = model.predict([X[:, i] for i in range(X.shape[1])], B, domain_mask=dm.numpy())
    assert np.all(pred[dm.numpy() == 0] == 0) and np.all(pred[dm.numpy() == 1] > 0)
