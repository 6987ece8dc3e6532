"""Experiment driver with the reference's CLI (``/root/reference/main.py``): for each seed build the dataset,
the model named by the config (or ``--run True --model_name X``), ``compile``, ``fit``, ``predict`` and append
per-task LogLoss / AUC to the results CSV.  Differences from the reference, all bug fixes: metric functions
are imported, label columns are de-duplicated in order (Q3), ``--seed`` is honoured when given.

    python main.py --config configs/ae.json --run True --model_name ple --device cuda
"""
import argparse
import os
import random

import numpy as np
import pandas as pd
import torch
from sklearn.metrics import log_loss, roc_auc_score

from mmlrec_b200.model import get_model
from mmlrec_b200.utils.data_utils import ctrdataset, unserialize


def set_seed(seed):
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


def masked(values, mask):
    return np.asarray(values)[np.asarray(mask, dtype=bool)].reshape(-1, 1)


def run(args):
    seeds = [args.seed] if args.seed is not None else [0, 2, 4, 8]
    for seed in seeds:
        print("seed:", seed)
        set_seed(seed)
        config = unserialize(args.config)
        dc, mc, oc, tc = (config[k] for k in ("data_config", "model_config", "optim_config", "training_config"))
        if args.b200_precision:
            config.setdefault("b200_config", {})["precision"] = args.b200_precision
        if args.run:
            mc["model_name"] = args.model_name
        target = list(dict.fromkeys(dc.get("label_columns", ["label"])))
        train, test, test_mask, train_in, test_in, _, df_columns = ctrdataset(config)
        model = get_model(mc.get("model_name", "sharedbottom"), df_columns, config, device=args.device)
        model.compile(optimizer=oc.get("optimizer", "adagrad"),
                      loss=oc.get("loss", ["binary_crossentropy", "binary_crossentropy"]),
                      metrics=oc.get("metrics", ["auc", "acc"]))
        y_train = train[dc["label_columns"]].values if len(target) < len(dc["label_columns"]) else train[target].values
        y_test = test[dc["label_columns"]].values if len(target) < len(dc["label_columns"]) else test[target].values
        best = model.fit(train_in, y_train, batch_size=tc.get("train_batch_size", 4096), epochs=tc.get("epochs", 10),
                         validation_data=(test_in, y_test))
        pred = best.predict(test_in, tc.get("test_batch_size", 4096))
        row = {"type": "_".join([dc.get("data_name", ""), mc.get("task_name", ""), mc.get("model_name", ""), str(seed)])}
        nd = dc.get("num_domains", 0)
        for i, name in enumerate(model.task_types):
            if model.task_name in ("msl", "mtmsl"):
                m = test_mask[:, i if model.task_name == "msl" else i % nd]
                yt, yp = masked(y_test[:, i], m), masked(pred[:, i], m)
            else:
                yt, yp = y_test[:, i], pred[:, i]
            row[f"log_loss_{i}"], row[f"auc_{i}"] = round(log_loss(yt, yp), 4), round(roc_auc_score(yt, yp), 4)
            print("%s test LogLoss" % name, row[f"log_loss_{i}"])
            print("%s test AUC" % name, row[f"auc_{i}"])
        if model.task_name == "msl":
            row["total_auc"] = round(roc_auc_score(y_test[:, 0], pred.sum(-1)), 4)
        elif model.task_name == "mtmsl":
            agg = np.stack([pred[:, :nd].sum(-1), pred[:, nd:].sum(-1)], -1)
            row["total_auc"] = round(roc_auc_score(y_test[:, [0, nd]], agg), 4)
        print(row)
        out = dc.get("test_result_path", "")
        if out:
            pd.DataFrame([row]).to_csv(out, mode="a", index=False, header=not os.path.exists(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--run", type=lambda s: str(s).lower() not in ("", "0", "false"), default=False)
    ap.add_argument("--model_name", type=str, default="")
    ap.add_argument("--config", type=str, default="")
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--b200_precision", default="", choices=["", "fp32", "bf16"])
    run(ap.parse_args())
