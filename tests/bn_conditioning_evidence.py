"""Evidence for the BatchNorm conditioning note: the SAME oracle code (plain torch) run on the CPU and on
CUDA disagrees in first-step gradients by ~1e-2 when init_std=1e-4, and by ~1e-6 when init_std=0.05."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import load_golden, golden_init, oracle_columns
from oracle.mmlrec_oracle import OracleTrainer

for case in ("mmoe_census_bn_default_init_adam", "mmoe_census_bn_adam"):
    z, cfg, fields = load_golden(case)
    params, bufs, trainable = golden_init(z)
    X, y = torch.from_numpy(z["step0/X"]), torch.from_numpy(z["step0/y"])
    cpu = OracleTrainer(cfg, oracle_columns(cfg, fields), params, bufs, trainable)
    _, _, g_cpu = cpu.loss_and_grads(X, y)
    gpu = OracleTrainer(cfg, oracle_columns(cfg, fields), {k: v.cuda() for k, v in params.items()},
                        {k: v.cuda() for k, v in bufs.items()}, trainable)
    _, _, g_gpu = gpu.loss_and_grads(X.cuda(), y.cuda())
    worst = []
    for k in trainable:
        if g_cpu[k] is None or "embedding" in k:
            continue
        scale = float(g_cpu[k].abs().max())
        err = float((g_cpu[k] - g_gpu[k].cpu()).abs().max())
        worst.append((err / max(scale, 1e-30), k, scale, err))
    worst.sort(reverse=True)
    print(f"{case} (init_std={float(z['meta/init_std'])}): torch-CPU vs torch-CUDA first-step gradients, worst 5 by relative error")
    for rel, k, scale, err in worst[:5]:
        print(f"   {k:40s} scale {scale:.3e} max|diff| {err:.3e} rel {rel:.2e}")
