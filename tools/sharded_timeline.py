"""Run under torchrun: eager CUDA-event timing of the pieces of the row-sharded step on rank 0
(barrier, peer-read gather, push, dense all-reduce, owner sort, owner K2) next to the whole captured step."""
import os, sys, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from mmlrec_b200 import synthetic, parallel, lib as L
from mmlrec_b200.model import get_model_class
from mmlrec_b200.model.utils import DenseFeat, SparseFeat

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
vocab = int(os.environ.get("VOCAB", 10_000_000))
B = int(os.environ.get("BATCH", 4096))
cfg, fields = synthetic.workload("synth26_mmoe", vocab=vocab)
cfg["b200_config"] = {"precision": "bf16", "cuda_graph": True,
                      "shard_tables": {"rank": rank, "world": world, "gather": os.environ.get("GATHER", "owner_serve")}}
emb = cfg["model_config"]["emb"]
cols = [SparseFeat(n, v, emb) if k == "sparse" else DenseFeat(n, 1) for n, k, v in fields]
torch.manual_seed(0)
model = get_model_class("mmoe")(cols, device=f"cuda:{local}", config=cfg)
model.compile(cfg["optim_config"]["optimizer"], cfg["optim_config"]["loss"], [])
parallel.attach_sharded(model)
model.train()
X, y = synthetic.make_batch(cfg, fields, B, seed=rank)
for _ in range(4):
    model.train_on_batch(X, y)
plan = model.plan(B)
g, st = plan.gather, torch.cuda.current_stream().cuda_stream
lib, hy = plan.b.lib, model.hyper_dev


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


rows = []
rows.append(("barrier: 4-byte NCCL all-reduce", timed(model.shard.barrier)))
rows.append(("barrier: flag exchange through peer memory", timed(lambda: model.shard.flag_barrier(st))))
rows.append((f"forward exchange + K1 ({model.shard.gather_mode})", timed(lambda: g.forward(st, False))))
rows.append(("owner: sort of the received keys", timed(lambda: (g.forward(st, True), g.sort(st)))))
rows.append(("push gradient rows to the owners", timed(lambda: g.backward(st))))
rows.append(("dense-gradient all-reduce", timed(lambda: model.dp.sum_gradients(model.store.dense_grad))))
rows.append(("owner: K2 on the received rows", timed(lambda: g.post_reduce(st))))
L.check(lib.mmlrec_hyper_advance(hy.data_ptr(), st))
rows.append(("whole step (graph replay)", timed(lambda: model._run_train(plan))))
if rank == 0:
    print(f"row-sharded synth26_mmoe, {world} GPUs, per-GPU batch {B}, vocab {vocab}, gather={model.shard.gather_mode} "
          "(us, eager, rank 0; 'sort' row includes a training-mode forward exchange)")
    for k, v in rows:
        print(f"  {k:48s} {v:9.1f}")
dist.barrier()
os._exit(0)
