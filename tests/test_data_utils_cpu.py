"""CPU: the data-pipeline surface (utils/data_utils.py interface of the reference) on a synthetic CSV."""
import json

import numpy as np
import pandas as pd


def test_ctrdataset_schema_and_encoding(tmp_path):
    from mmlrec_b200.utils.data_utils import ctrdataset, get_test_mask, unserialize
    rng = np.random.default_rng(0)
    n = 40
    df = pd.DataFrame({"c1": rng.choice(["x", "y", "z"], n), "c2": rng.integers(10, 15, n), "scene": rng.integers(0, 2, n),
                       "n1": rng.normal(size=n) * 3 + 7, "label": rng.integers(0, 2, n)})
    df.iloc[:30].to_csv(tmp_path / "tr.csv", index=False)
    df.iloc[30:].to_csv(tmp_path / "te.csv", index=False)
    cfg = {"data_config": {"train_dataset_path": str(tmp_path / "tr.csv"), "test_dataset_path": str(tmp_path / "te.csv"),
                           "all_columns": ["c1", "c2", "scene", "n1", "label"], "feature_columns": ["c1", "c2"],
                           "dense_columns": ["n1"], "ignore_columns": [], "label_columns": ["label", "label"],
                           "num_domains": 2, "mask_values": [0, 1], "mask_column": "scene", "scene_feature": "scene"},
           "model_config": {"task_name": "msl", "emb": 8}}
    (tmp_path / "c.json").write_text(json.dumps(cfg))
    cfg = unserialize(str(tmp_path / "c.json"))
    train, test, test_mask, tr_in, te_in, lin_cols, dnn_cols = ctrdataset(cfg)
    assert [c.name for c in dnn_cols] == ["c1", "c2", "scene", "n1"], "scene feature appended to the sparse list"
    assert [getattr(c, "vocabulary_size", None) for c in dnn_cols] == [3, 5, 2, None]
    assert all(getattr(c, "embedding_dim", 8) == 8 for c in dnn_cols[:3])
    assert len(train) == 30 and len(test) == 10 and list(tr_in) == ["c1", "c2", "scene", "n1"]
    full = pd.concat([train, test])
    assert full["n1"].min() == 0.0 and full["n1"].max() == 1.0            # MinMax over train+test
    assert sorted(full["c1"].unique()) == [0, 1, 2]                        # LabelEncoder: sorted uniques -> codes
    assert np.array_equal(full["c1"].values, pd.Categorical(df["c1"], categories=["x", "y", "z"]).codes)
    assert train[["label"]].shape[1] == 2, "duplicated label names become duplicated columns (SURVEY Q3)"
    assert np.array_equal(test_mask, get_test_mask(test["scene"], [0, 1], 2)) and test_mask.shape == (10, 2)
