"""Run the grouped-GEMM launches of ONE LinearStage a few times (for `ncu -k regex:gemm_grouped_tc`):
    python tools/ncu_one_stage.py <workload> <stage-label> <fwd|bwd> [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from mmlrec_b200 import lib as L, synthetic  # noqa: E402


class A:
    workload = sys.argv[1]
    vocab = 0


label, which = sys.argv[2], sys.argv[3]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
cfg, fields = bench.workload_config(A)
model = bench.build_ours(cfg, fields, "cuda:0", "bf16")
X, y = synthetic.make_batch(cfg, fields, 4096, seed=0)
for _ in range(2):
    model.train_on_batch(X, y)
torch.cuda.synchronize()
plan = model.plan(4096)
st = torch.cuda.current_stream().cuda_stream
for s in plan.stages:
    if s.name == "linear" and s.label == label:
        for _ in range(reps):
            for t in (s.fwd if which == "fwd" else s.bwd):
                L.check(plan.b.tc_launch(t, st))
torch.cuda.synchronize()
