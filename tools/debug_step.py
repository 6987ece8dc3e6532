"""Debug aid: run one training step of a golden case stage by stage with a sync after each launch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import load_golden
from test_step_gpu import build_model, load_init
from mmlrec_b200 import lib as L

case = sys.argv[1] if len(sys.argv) > 1 else "mmoe_census_bn_adagrad"
z, cfg, fields = load_golden(case)
model, cfg = build_model(cfg, fields, cuda_graph=False)
load_init(model, z)
model.compile(cfg["optim_config"]["optimizer"], cfg["optim_config"]["loss"], [])
model.train()
X, y = torch.from_numpy(z["step0/X"]), torch.from_numpy(z["step0/y"])
p = model.plan(X.shape[0])
p.X.copy_(X); p.y.copy_(y)
st = torch.cuda.current_stream().cuda_stream
def sync(tag):
    torch.cuda.synchronize(); print("ok", tag, flush=True)
L.check(p.b.lib.mmlrec_hyper_advance(model.hyper_dev.data_ptr(), st)); sync("hyper")
p.gather.sort(st); sync("sort")
for i, s in enumerate(p.stages):
    print("fwd", i, s.name, getattr(s, "label", ""), flush=True)
    s.forward(st, True); sync(f"fwd {i}")
for i, s in reversed(list(enumerate(p.stages))):
    print("bwd", i, s.name, getattr(s, "label", ""), flush=True)
    s.backward(st); sync(f"bwd {i}")
print("pred err", float((p.pred.cpu() - torch.from_numpy(z["step0/pred"])).abs().max()))
print("loss", p.loss.cpu().tolist(), float(z["step0/loss"]))
import numpy as np
for name, prm in model.named_parameters():
    if getattr(prm, "_mm_kind", "") != "dense" or ("grad0/" + name) not in z.files:
        continue
    g = model.store.grad_view(prm).cpu().numpy()
    want = z["grad0/" + name]
    err = np.abs(g - want)
    bad = err > 1e-5 * np.abs(want).max() + 1e-9
    print(f"{name:45s} scale {np.abs(want).max():.3e} maxerr {err.max():.3e} bad {int(bad.sum())}/{bad.size}",
          ("rows " + str(sorted(set(np.argwhere(bad)[:, 0].tolist()))[:8])) if bad.any() and bad.ndim == 2 else "", flush=True)
