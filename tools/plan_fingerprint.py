"""SHA-1 fingerprints of the planned programs (activation groups with their column layout, gradient precisions, parameter
order = flat-store layout, stage list) of the seven BASELINE workloads and of every golden case, in both arithmetic modes.
No GPU needed.  Two commits with equal fingerprints ask the GPU for the same programs, so a GPU test run of one holds for
the other as far as the planner is concerned (kernel sources aside):

    python tools/plan_fingerprint.py [repo root] > profiles/plan_fingerprints_r02.json
"""
import hashlib
import json
import os
import sys

root = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
sys.path.insert(0, os.path.join(root, "tests"))

import torch  # noqa: E402

from helpers import GOLDEN_CASES, load_golden  # noqa: E402
from mmlrec_b200 import synthetic  # noqa: E402
from mmlrec_b200.engine.core import Builder  # noqa: E402
from mmlrec_b200.model import get_model_class  # noqa: E402
from mmlrec_b200.model.utils import DenseFeat, SparseFeat  # noqa: E402


def fingerprint(cfg, fields, precision, B):
    cfg["b200_config"] = {"precision": precision}
    emb = cfg["model_config"]["emb"]
    cols = [SparseFeat(n, v, emb) if k == "sparse" else DenseFeat(n, 1) for n, k, v in fields]
    torch.manual_seed(0)
    model = get_model_class(cfg["model_config"]["model_name"])(cols, device="cpu", config=cfg)
    b = Builder(B, torch.device("cpu"), None, dry=True, precision=precision)
    model.build_graph(b)
    names = {id(p): n for n, p in model.named_parameters()}
    fp = {"groups": [(g.name, g.widths, [a.col for a in g.acts], g.total, g.grad_dtype, g.relu) for g in b.groups],
          "params": [names[id(p)] for p in b.param_order], "stages": [(s.name, getattr(s, "label", "")) for s in b.stages]}
    return hashlib.sha1(json.dumps(fp, sort_keys=True).encode()).hexdigest()


def main():
    out = {}
    small = {"ae_ple_t4": dict(max_vocab=100), "ae_ple_t2": dict(max_vocab=100), "census_mmoe": {},
             "kuairec_sharedbottom": dict(max_vocab=100), "kuairec_esmm": dict(max_vocab=100),
             "movielens_star": dict(vocab_scale=0.01), "movielens_pepnet": dict(vocab_scale=0.01), "synth26_mmoe": dict(vocab=100)}
    for wl, kw in small.items():
        for precision in ("bf16", "fp32"):
            out[f"workload/{wl}/{precision}"] = fingerprint(*synthetic.workload(wl, **kw), precision, 4096)
    for case in GOLDEN_CASES:
        for precision in ("bf16", "fp32"):
            _, cfg, fields = load_golden(case)
            out[f"golden/{case}/{precision}"] = fingerprint(cfg, fields, precision, 48)
    print(json.dumps(out, indent=0, sort_keys=True))


if __name__ == "__main__":
    main()
