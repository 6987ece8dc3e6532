"""Long-run check: the loss of many training steps over a fixed pool of batches, fp32 mode vs bf16 mode vs the
oracle's step body run as torch eager on the same GPU (same seeded initial state, same batches).

    python tools/loss_trajectory.py [workload] [steps] [every]

Three steps of parity say nothing about whether the rounding of the tensor-core mode drifts over a real run; this does.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from mmlrec_b200 import synthetic  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "ae_ple_t4"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
every = int(sys.argv[3]) if len(sys.argv) > 3 else 100
B, pool = 4096, 8


class A:
    vocab = 0


cfg, fields = bench.workload_config(A(), wl)
host = [synthetic.make_batch(cfg, fields, B, seed=s) for s in range(pool)]
rows = {}
for prec in ("fp32", "bf16"):
    model = bench.build_ours(cfg, fields, "cuda:0", prec)
    tr = []
    for i in range(steps):
        loss = model.train_on_batch(host[i % pool][0], host[i % pool][1])
        if i % every == 0 or i == steps - 1:
            tr.append((i, float(loss[-1].item())))
    rows[prec] = tr
    del model
    torch.cuda.empty_cache()
if os.environ.get("ORACLE", "1") == "1":
    o = bench.build_oracle(cfg, fields, device="cuda:0")
    tr = []
    for i in range(steps):
        X, y = host[i % pool]
        _, loss = o.step(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda())
        if i % every == 0 or i == steps - 1:
            tr.append((i, float(loss)))
    rows["oracle (torch eager, cuda)"] = tr
print(f"# {wl}, B={B}, pool of {pool} batches, loss of the step (sum over tasks)")
keys = list(rows)
print("step".rjust(6) + "".join(k.rjust(30) for k in keys))
for j in range(len(rows[keys[0]])):
    print(str(rows[keys[0]][j][0]).rjust(6) + "".join(f"{rows[k][j][1]:30.4f}" for k in keys))
