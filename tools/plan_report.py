"""Print the planned step program of a workload WITHOUT a GPU: the stage list, every grouped-GEMM problem of the forward
and backward launches (M, N, K, operand majors, outputs, split-K slice) with its flops, and the sequence of C-ABI calls
one forward + backward issues.  The program is built by the product planner (engine/core.py) on CPU memory and its
argument tables are read back through tests/plan_emulator.py, so this is what a B200 would be asked to run.

    python tools/plan_report.py [--workload ae_ple_t4] [--batch 4096] [--precision bf16] [--model-name NAME] [--run]
"""
import argparse
import collections
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from mmlrec_b200 import synthetic  # noqa: E402
from mmlrec_b200.engine import core  # noqa: E402
from mmlrec_b200.model import get_model_class  # noqa: E402
from mmlrec_b200.model.utils import DenseFeat, SparseFeat  # noqa: E402
import plan_emulator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="ae_ple_t4")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--model-name", default=None, help="override model_config.model_name (e.g. apg on movielens_star)")
    ap.add_argument("--max-vocab", type=int, default=2000, help="rows per table (the program does not depend on it)")
    ap.add_argument("--run", action="store_true", help="also run one forward + backward on the CPU and list the calls")
    args = ap.parse_args()
    kw = {"max_vocab": args.max_vocab} if args.workload.startswith(("ae_", "kuairec")) else (
        {"vocab": args.max_vocab} if args.workload.startswith("synth26") else {"vocab_scale": 0.05}
        if args.workload.startswith("movielens") else {})   # ("<dataset>:<model>" names start with the dataset too)
    cfg, fields = synthetic.workload(args.workload, **kw)
    if args.model_name:
        cfg["model_config"]["model_name"] = args.model_name
    cfg = copy.deepcopy(cfg)
    cfg["b200_config"] = {"precision": args.precision, "cuda_graph": False}
    emb = cfg["model_config"]["emb"]
    cols = [SparseFeat(n, v, emb) if k == "sparse" else DenseFeat(n, 1) for n, k, v in fields]
    torch.manual_seed(0)
    model = get_model_class(cfg["model_config"]["model_name"])(cols, device="cpu", config=cfg)
    B = args.batch
    plan = plan_emulator.EmulatedPlan(model, B, args.precision)
    plan.build()
    print(f"# {args.workload} / {cfg['model_config']['model_name']}  B={B}  {args.precision}  "
          f"dense parameters {model.store.n_dense}  gradient slices {plan.grad_slices}")
    total = collections.Counter()
    launches = collections.Counter()

    def describe(tbls, what, label):
        for tbl in tbls:
            if tbl[0] == "captured":
                descs = plan_launches(plan)[tbl[1]]
                rows = [(d.M, d.N, d.K, "mn" if d.a_mn_major else "k", "mn" if d.b_mn_major else "k",
                         "f32" if d.C_f32 else "bf16", bool(d.accumulate), bool(d.c_transposed)) for d in descs]
            else:
                import ctypes as C
                from mmlrec_b200 import lib as L
                raw = bytes(plan_emulator.view(tbl[0].data_ptr(), plan_emulator.np.uint8, tbl[2] * C.sizeof(L.GemmF32)))
                ps = [L.GemmF32.from_buffer_copy(raw, i * C.sizeof(L.GemmF32)) for i in range(tbl[2])]
                rows = [(p.M, p.N, p.K, "-", "-", "f32", bool(p.accumulate), False) for p in ps]
            fl = sum(2 * m * n * k for m, n, k, *_ in rows)
            total[what] += fl
            launches[what] += 1
            print(f"  {what} launch {label}: {len(rows)} problems, {fl / 1e9:.3f} GFLOP")
            for m, n, k, am, bm, out, acc, ct in rows:
                print(f"      M={m:<6} N={n:<6} K={k:<6} A:{am:<2} B:{bm:<2} -> {out}{' +=' if acc else ''}{' (stored transposed)' if ct else ''}")

    for i, st in enumerate(plan.stages):
        print(f"[{i}] {st.name} {getattr(st, 'label', '')}")
        if isinstance(st, core.LinearStage):
            describe(st.fwd, "fwd", st.label)
            describe(st.bwd, "bwd", st.label)
            if st.colsums:
                print(f"  bwd: {len(st.colsums)} column-sum bias gradient(s)")
    print(f"# GEMM launches per step: {launches['fwd']} forward + {launches['bwd']} backward; "
          f"{(total['fwd'] + total['bwd']) / 1e9:.2f} GFLOP per step "
          f"({(total['fwd'] + total['bwd']) / B / 1e6:.3f} MFLOP per sample, forward {total['fwd'] / B / 1e6:.3f})")
    if args.run:
        X, y = synthetic.make_batch(cfg, fields, B, seed=0)
        plan.forward_backward(X, y)
        print("# C-ABI calls of one forward + backward (the table update, optimizer and clock kernels come on top):")
        print("  " + " | ".join(plan.fake.calls))


def plan_launches(plan):
    return plan.launch_tables


if __name__ == "__main__":
    main()
