"""Row-sharded embedding tables (csrc/peer.cu) on ONE GPU: with world = 1 every id is owned locally, so the whole
one-sided exchange (push ids -> flag barrier -> serve rows -> staged K1 -> side-stream key sort -> push gradient
rows -> K2 over the receive rows) must reproduce the ordinary single-GPU step BIT FOR BIT, which in turn is pinned
to the reference's golden vectors by test_step_gpu.py.  The multi-GPU equivalence is tools/sharded_equivalence.py
(profiles/sharded_equivalence_2gpu_r01.txt)."""
import copy
import os
import socket

import pytest
import torch

from helpers import golden_init, load_golden
from test_step_gpu import build_model, load_init

pytestmark = pytest.mark.gpu


def _group():
    import torch.distributed as dist
    if not dist.is_initialized():
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda:0"))
    return dist


@pytest.mark.parametrize("gather", ["owner_serve", "peer_read"])
@pytest.mark.parametrize("case,precision", [("mmoe_synth26_adagrad", "fp32"), ("ple_ae_t4_adam", "fp32"),
                                            ("ple_ae_t4_adam", "bf16")])
def test_world1_sharded_step_is_bit_identical_to_plain_step(case, precision, gather):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from mmlrec_b200 import parallel, synthetic
    from mmlrec_b200.model import get_model_class
    from mmlrec_b200.model.utils import DenseFeat, SparseFeat
    _group()
    z, cfg, fields = load_golden(case)
    plain, _ = build_model(cfg, fields, precision=precision, cuda_graph=True)
    load_init(plain, z)
    cfg_s = copy.deepcopy(cfg)
    cfg_s["b200_config"] = {"precision": precision, "cuda_graph": True,
                            "shard_tables": {"rank": 0, "world": 1, "gather": gather}}
    emb = cfg["model_config"]["emb"]
    cols = [SparseFeat(n, v, emb) if k == "sparse" else DenseFeat(n, 1) for n, k, v in fields]
    sharded = get_model_class(cfg["model_config"]["model_name"])(cols, device="cuda:0", config=cfg_s)
    load_init(sharded, z)   # world = 1: the shard of every table is the whole table
    for m in (plain, sharded):
        m.compile(cfg["optim_config"]["optimizer"], cfg["optim_config"]["loss"], [])
        m.train()
    parallel.attach_sharded(sharded)
    for s in range(4):   # eager step, graph capture, replays
        X, y = synthetic.make_batch(cfg, fields, 160, seed=300 + s)
        la = plain.train_on_batch(X, y)
        lb = sharded.train_on_batch(X, y)
        assert float(la[-1].item()) == float(lb[-1].item()), f"step {s} loss"
    Xp, _ = synthetic.make_batch(cfg, fields, 160, seed=999)
    assert (plain.predict(Xp, 160) == sharded.predict(Xp, 160)).all(), "forward-only path"
    X, y = synthetic.make_batch(cfg, fields, 160, seed=555)   # a training step after the evaluation pass
    plain.train_on_batch(X, y)
    sharded.train_on_batch(X, y)
    torch.cuda.synchronize()
    sharded.shard.check()
    sa, sb = plain.state_dict(), sharded.state_dict()
    for k, v in sa.items():
        assert torch.equal(v, sb[k]), k
    assert int(sharded.plan(160).gather.oob.item()) == 0


@pytest.mark.parametrize("R,n", [(1, 1000), (2, 4096), (8, 2397 * 4), (3, 44)])
def test_peer_allreduce_kernel_simulated_ranks(R, n):
    """mmlrec_peer_allreduce_f32 with all R "ranks" living on one GPU: rank r reduces slice r of every input buffer in
    rank order and stores it into slice r of every output buffer, so after the R launches every output equals the
    rank-ordered sum -- identical on all ranks, bit for bit."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from mmlrec_b200 import lib as L
    lib = L.load()
    n = (n + 3) // 4 * 4
    g = torch.Generator().manual_seed(R * 1000 + n)
    ins = [torch.randn(n, generator=g).cuda() for _ in range(R)]
    outs = [torch.full((n,), float("nan"), device="cuda") for _ in range(R)]
    tin = torch.tensor([t.data_ptr() for t in ins], dtype=torch.int64, device="cuda")
    tout = torch.tensor([t.data_ptr() for t in outs], dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for r in range(R):
        L.check(lib.mmlrec_peer_allreduce_f32(tin.data_ptr(), tout.data_ptr(), n, r, R, st), "peer_allreduce")
    torch.cuda.synchronize()
    want = ins[0].clone()
    for t in ins[1:]:
        want = want + t          # the kernel's order: ((in0 + in1) + in2) + ...
    for o in outs:
        assert torch.equal(o, want)
