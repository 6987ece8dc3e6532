"""Golden record of the REFERENCE's ``fit()`` / ``evaluate()`` / ``predict()`` (``/root/reference/model/basemodel.py:
135-457``) on a small synthetic mtl dataset: per-epoch training loss / metrics and validation metrics exactly as the
reference prints them, the evaluate dict of every epoch, and the predictions of the returned best model.  Runs ONLY in
the build container (imports /root/reference read-only).  ``tests/test_fit_gpu.py`` replays the same call on the CUDA
path.       python tests/golden/make_fit_golden.py
"""
import contextlib
import io
import json
import os
import re
import sys

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from mmlrec_b200 import synthetic  # noqa: E402
from make_golden import SMALL, build_reference  # noqa: E402

N_TRAIN, N_VAL, BATCH, EPOCHS, SEED = 1500, 500, 128, 3, 2


def dataset(cfg, fields, n, seed):
    """ids uniform per field; labels depend on a few id parities so that AUC moves away from 0.5"""
    X, y = synthetic.make_batch(cfg, fields, n, seed=seed)
    rng = np.random.default_rng(1000 + seed)
    s = (X[:, 0].astype(np.int64) % 2) * 1.5 + (X[:, 1].astype(np.int64) % 3 == 0) * 1.0 - 1.2
    for t in range(y.shape[1]):
        p = 1 / (1 + np.exp(-(s + 0.6 * ((X[:, 2 + t].astype(np.int64) % 2) - 0.5))))
        y[:, t] = (rng.random(n) < p).astype(np.float32)
    names = [f[0] for f in fields]
    return pd.DataFrame(X.astype(np.int64), columns=names), y


def parse_log(text):
    epochs, evals = [], []
    for line in text.splitlines():
        line = line.strip()
        if line.startswith("{") and "auc" in line:
            evals.append({k: float(v) for k, v in re.findall(r"'(\w+)': (?:np\.float64\()?([-0-9.e]+)", line)})
        m = re.match(r"\d+s - (.*)", line)
        if m:
            epochs.append({k.strip(): float(v) for k, v in (kv.split(":") for kv in m.group(1).split(" - "))})
    return epochs, evals


def main():
    cfg, fields = synthetic.workload("kuairec_sharedbottom", max_vocab=200)
    cfg["model_config"].update(SMALL)
    cfg["optim_config"].update(lr=5e-3, metrics=["auc", "acc"], early_stop=5)
    torch.manual_seed(SEED)
    np.random.seed(SEED)
    train_df, y_train = dataset(cfg, fields, N_TRAIN, 11)
    val_df, y_val = dataset(cfg, fields, N_VAL, 12)
    model = build_reference(cfg, fields, init_std=0.05)
    names = [f[0] for f in fields]
    train_in = {n: train_df[n] for n in names}
    val_in = {n: val_df[n] for n in names}
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        best = model.fit(train_in, y_train, batch_size=BATCH, epochs=EPOCHS, validation_data=(val_in, y_val))
        final_eval = best.evaluate([val_df[n].values for n in names], y_val, BATCH)
        pred = best.predict([val_df[n].values for n in names], BATCH)
    epochs, evals = parse_log(buf.getvalue())
    assert len(epochs) == EPOCHS and len(evals) >= EPOCHS, buf.getvalue()
    out = {"meta/config": np.array(json.dumps(cfg)), "meta/fields": np.array(json.dumps(fields)),
           "meta/args": np.array(json.dumps(dict(n_train=N_TRAIN, n_val=N_VAL, batch=BATCH, epochs=EPOCHS, seed=SEED,
                                                 init_std=0.05))),
           "train/X": train_df.values.astype(np.int64), "train/y": y_train, "val/X": val_df.values.astype(np.int64), "val/y": y_val,
           "log/epochs": np.array(json.dumps(epochs)), "log/evals": np.array(json.dumps(evals[:EPOCHS])),
           "final/eval": np.array(json.dumps({k: float(v) for k, v in final_eval.items()})), "final/pred": pred}
    path = os.path.join(HERE, "fit_sharedbottom_kuairec.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")
    print(epochs)
    print(evals[:EPOCHS], final_eval)


if __name__ == "__main__":
    main()
