"""Digest of the reference's 12 shipped JSON configs (``/root/reference/configs_*/*.json``) for the CPU sweep in
tests/test_shipped_configs_cpu.py: per config the SHAPE the models see -- how many sparse / dense columns, where the
scene / side features sit in the sparse list, the number of labels and domains -- plus the model / optimizer
hyper-parameters.  Column names, paths and the rest of the data section are not carried over (columns are renamed
s0.., d0..).  Runs only in the build container:

    python tests/golden/make_config_shapes.py        # rewrites tests/golden/shipped_configs.json
"""
import glob
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    out = []
    for path in sorted(glob.glob("/root/reference/configs_*/*.json")):
        cfg = json.load(open(path))
        dc = cfg["data_config"]
        sparse = list(dc.get("feature_columns", []))
        scene = dc.get("scene_feature", "")
        if scene and scene not in sparse:   # utils/data_utils.py:49-50 appends the scene column to the sparse features
            sparse.append(scene)

        def pos(name):
            return sparse.index(name) if name and name in sparse else -1
        out.append({
            "config": "/".join(path.split("/")[-2:]),
            "n_sparse": len(sparse), "n_dense": len(dc.get("dense_columns", [])),
            "scene_pos": pos(scene), "mask_pos": pos(dc.get("mask_column", "")),
            "user_sf_pos": pos(dc.get("user_sf", "")), "item_sf_pos": pos(dc.get("item_sf", "")),
            "n_labels": len(dc["label_columns"]), "num_domains": dc.get("num_domains", 1),
            "mask_values": dc.get("mask_values", []),
            "model_config": cfg["model_config"], "optim_config": cfg["optim_config"]})
    with open(os.path.join(HERE, "shipped_configs.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(f"{len(out)} configs")


if __name__ == "__main__":
    main()
