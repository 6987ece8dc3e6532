"""K1 alone on 2^20 resident rows of a workload (for timing and `ncu --set full -k regex:gather_concat`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import json  # noqa: E402

import torch  # noqa: E402

import bench  # noqa: E402


class A:
    workload = sys.argv[1] if len(sys.argv) > 1 else "ae_ple_t4"
    vocab = 0


cfg, fields = bench.workload_config(A)
cfg["_workload_name"] = A.workload
model = bench.build_ours(cfg, fields, "cuda:0", "bf16")
hbm, _, _, which = bench.measured_peaks()
print(json.dumps(bench.gather_roofline(model, cfg, fields, hbm, which)))
