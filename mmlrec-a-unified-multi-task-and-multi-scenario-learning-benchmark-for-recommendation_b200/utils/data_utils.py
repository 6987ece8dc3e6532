"""Host-side data surface with the reference's interface (``/root/reference/utils/data_utils.py``):
``ctrdataset(config)`` returns ``(train, test, test_mask, train_model_input, test_model_input,
linear_feature_columns, dnn_feature_columns)``, ``unserialize(path)`` loads a JSON / npy / pickle config,
``get_test_mask`` builds the per-domain 0/1 mask.  It runs once per seed, outside the training step; it is
plain pandas and is not part of the accelerated path."""
import json
import os
import pickle

import numpy as np
import pandas as pd

from ..model.utils import DenseFeat, SparseFeat, get_feature_names


def unserialize(path):
    ext = os.path.basename(path).rsplit(".", 1)[-1]
    if ext == "json":
        with open(path, "r") as f:
            return json.load(f)
    if ext == "ny":  # the reference's spelling of ".npy"
        return np.load(path)
    with open(path, "rb") as f:
        return pickle.load(f)


def get_test_mask(domain_values, mask_values, num_domains):
    dv = np.asarray(domain_values).reshape(-1, 1)
    mv = np.asarray(mask_values).reshape(1, -1)
    assert mv.shape[1] == num_domains
    return (dv == mv).astype(np.float32)


def _encode(df, all_columns, dense_columns, skip, as_str):
    """LabelEncoder semantics (sorted unique -> 0..n-1) for categorical columns, MinMax scaling for dense."""
    for col in all_columns:
        if col in skip:
            continue
        if as_str:
            df[col] = df[col].astype(str)
        if col in dense_columns:
            lo, hi = df[col].min(), df[col].max()
            df[col] = (df[col] - lo) / (hi - lo) if hi > lo else 0.0
        else:
            df[col] = pd.Categorical(df[col], categories=np.sort(df[col].unique())).codes.astype(np.int64)
    return df


def ctrdataset(config):
    dc, mc = config["data_config"], config["model_config"]
    all_columns = dc.get("all_columns", [])
    sparse = list(dc.get("feature_columns", []))
    dense = list(dc.get("dense_columns", []))
    labels = list(dc.get("label_columns", ["label"]))
    train_path, test_path = dc.get("train_dataset_path", ""), dc.get("test_dataset_path", "")
    train_df = pd.read_csv(train_path, usecols=all_columns)
    test_df = pd.read_csv(test_path, usecols=all_columns)
    if "kuairec" in train_path:
        for col in all_columns:
            if "onehot" in col:
                train_df[col], test_df[col] = train_df[col].astype(str), test_df[col].astype(str)
        train_df = train_df[train_df["user_active_degree"] != "0"]
    if "iaac" in train_path:
        for d in (train_df, test_df):
            d["predict_category_property"] = d["predict_category_property"].astype(str)
        test_df = test_df[:-2]
    n_train = len(train_df)
    df = pd.concat([train_df, test_df])
    scene = dc.get("scene_feature", "")
    if scene and scene not in sparse:
        sparse.append(scene)
    df = _encode(df, all_columns, dense, set(labels) | set(dc.get("ignore_columns", [])), "amazon_new" in train_path)
    task_name, mask_col = mc.get("task_name", "mtl"), dc.get("mask_column", "")
    columns = sparse + dense + labels
    with_mask = task_name in ("msl", "mtmsl") and mask_col != ""
    if with_mask and mask_col not in sparse:
        columns = columns + [mask_col]
    df = df.reindex(columns=columns)
    emb = mc.get("emb", 4)
    feature_columns = [SparseFeat(c, vocabulary_size=int(df[c].max()) + 1, embedding_dim=emb) for c in sparse]
    feature_columns += [DenseFeat(c, 1) for c in dense]
    names = get_feature_names(feature_columns)
    train, test = df[:n_train], df[n_train:]
    train_in = {n: train[n] for n in names}
    test_in = {n: test[n] for n in names}
    test_mask = None
    if with_mask:
        if mask_col not in sparse:
            train_in[mask_col], test_in[mask_col] = train[mask_col], test[mask_col]
        test_mask = get_test_mask(test[mask_col], dc.get("mask_values", []), dc.get("num_domains", 1))
    return train, test, test_mask, train_in, test_in, feature_columns, feature_columns
