"""Timeline of the tcgen05 grouped-GEMM launches inside a training step (profiling aid, 1 GPU):
per-tile clock64 stamps of CTA 0 (producer / MMA issuer / epilogue roles) and, for every CTA, global-timer stamps of
setup / producer / MMA / epilogue completion.   python tools/tc_timeline.py [workload] [stage-label ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from mmlrec_b200 import lib as L, synthetic  # noqa: E402


class A:
    workload = sys.argv[1] if len(sys.argv) > 1 else "ae_ple_t4"
    vocab = 0


labels = set(sys.argv[2:])
cfg, fields = bench.workload_config(A)
model = bench.build_ours(cfg, fields, "cuda:0", "bf16")
B = 4096
X, y = synthetic.make_batch(cfg, fields, B, seed=0)
for _ in range(2):
    model.train_on_batch(X, y)
torch.cuda.synchronize()
plan = model.plan(B)
lib = L.load()
st = torch.cuda.current_stream().cuda_stream
names = ["P.start", "P.table", "P.slot0", "P.issued", "M.start", "M.accfree", "M.data0", "M.commit", "E.start", "E.bias",
         "E.accrdy", "E.done", "E.release"]


def run(tag, tbl, tiles_shown=10):
    n_ctas = tbl[6] * (2 if plan.b.tc_kernel == 2 else 1)
    stamps = torch.zeros(1024 + 8 * max(n_ctas, 148), dtype=torch.int64, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for i in range(3):
        if i == 2:
            ev[0].record()
        L.check(plan.b.tc_launch(tbl, st, stamps.data_ptr()))
    ev[1].record()
    torch.cuda.synchronize()
    s = stamps.cpu().numpy()
    per_tile, cta = s[:1024].reshape(64, 16), s[1024:1024 + 8 * n_ctas].reshape(n_ctas, 8)
    t0 = int(per_tile[0, 0])
    print(f"=== {tag}: {tbl[2]} problems, {tbl[3]} tiles on {n_ctas} CTAs, launch {ev[0].elapsed_time(ev[1]) * 1e3:.1f} us")
    print("tile " + " ".join(f"{n:>9s}" for n in names) + "   (CTA 0, cycles from its first tile start)")
    for i in range(min(64, tiles_shown)):
        if int(per_tile[i, 0]) == 0:
            break
        cells = [f"{int(per_tile[i, k]) - t0:9d}" for k in range(13)]
        if plan.b.tc_kernel == 2:   # slot 1 carries (problem << 32 | k-blocks)
            cells[1] = f"p{int(per_tile[i, 1]) >> 32}k{int(per_tile[i, 1]) & 0xffff}".rjust(9)
        print(f"{i:4d} " + " ".join(cells))
    g0 = cta[:, 0].min()
    rel = (cta[:, :4] - g0) / 1e3
    print(f"per CTA (us from the earliest setup-done): setup-done max {rel[:, 0].max():.2f}; producer done "
          f"median {np.median(rel[:, 1]):.2f} max {rel[:, 1].max():.2f}; MMA done median {np.median(rel[:, 2]):.2f} "
          f"max {rel[:, 2].max():.2f}; epilogue done median {np.median(rel[:, 3]):.2f} max {rel[:, 3].max():.2f}; "
          f"tiles/CTA min {cta[:, 4].min()} max {cta[:, 4].max()}")
    print(f"   kernel entry -> setup done: median {np.median(cta[:, 0] - cta[:, 5]) / 1e3:.2f} us max {(cta[:, 0] - cta[:, 5]).max() / 1e3:.2f}; "
          f"entry spread {(cta[:, 5].max() - cta[:, 5].min()) / 1e3:.2f} us; first entry -> last exit {(cta[:, 6].max() - cta[:, 5].min()) / 1e3:.2f} us; "
          f"epilogue done -> exit median {np.median(cta[:, 6] - cta[:, 3]) / 1e3:.2f} us")
    dur = (cta[:, 3] - cta[:, 0]) / 1e3
    per = dur / np.maximum(cta[:, 4], 1)
    print(f"   busy time per CTA (setup-done -> epilogue done) median {np.median(dur):.2f} us, per tile median {np.median(per):.2f} us")


for s in plan.stages:
    if s.name != "linear" or (labels and s.label not in labels):
        continue
    for i, t in enumerate(s.fwd):
        run(f"fwd {s.label}[{i}]", t)
    for i, t in enumerate(s.bwd):
        run(f"bwd {s.label}[{i}]", t)
