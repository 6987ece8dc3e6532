"""Per-tile timeline (clock64 stamps of CTA 0) of tensor-core GEMM launches inside the PLE-AE step."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mmlrec_b200 import synthetic, lib as L

class A: workload = "ae_ple_t4"
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
names = ["P.start", "P.table", "P.slot0", "P.issued", "M.start", "M.accfree", "M.data0", "M.commit", "E.start", "E.bias", "E.accrdy", "E.done", "E.ld0", "E.ph1", "E.ph2"]
def run(tag, tbl):
    stamps = torch.zeros(64, 16, dtype=torch.int64, device="cuda")
    for _ in range(2):
        L.check(lib.mmlrec_gemm_grouped_tc_debug(tbl[0].data_ptr(), tbl[1].data_ptr(), tbl[2], tbl[3], stamps.data_ptr(), st))
    torch.cuda.synchronize()
    s = stamps.cpu()
    t0 = int(s[0, 0])
    print(f"=== {tag}: {tbl[2]} problems, {tbl[3]} tiles  (cycles relative to CTA0's first tile start)")
    print("tile " + " ".join(f"{n:>9s}" for n in names))
    for i in range(64):
        if int(s[i, 0]) == 0:
            break
        print(f"{i:4d} " + " ".join(f"{int(s[i, k]) - t0:9d}" for k in range(15)))
for idx, s in enumerate(plan.stages):
    if s.name != "linear":
        continue
    if s.label in ("cgc0.l0",):
        run(f"fwd {s.label}", s.fwd[0])
        run(f"bwd {s.label}", s.bwd[0])
