"""Checkpoint / resume: a run that is saved after two steps and resumed in a NEW model object continues bit for bit
(parameters, tables, optimizer state incl. the lazy dense-Adam row stamps, the device-side step clock)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import build_pair  # noqa: E402


@pytest.mark.parametrize("workload,precision,opt", [("ae_ple_t2", "fp32", "adam"), ("ae_ple_t2", "bf16", "adam"),
                                                    ("census_mmoe", "fp32", "adagrad"), ("kuairec_esmm", "bf16", "rmsprop")])
def test_resume_is_bit_identical(tmp_path, workload, precision, opt):
    from mmlrec_b200 import synthetic
    small = dict(expert_dnn_hidden_units=[32, 16], gate_dnn_hidden_units=[8], tower_dnn_hidden_units=[8],
                 bottom_dnn_hidden_units=[32, 16], dnn_hidden_units=[32, 16])
    kw = dict(max_vocab=300) if workload != "census_mmoe" else {}
    oc = dict(optimizer=opt, lr=1e-2)
    a, _, cfg, fields, _ = build_pair(workload, kw, precision, init_std=0.05, mc_over=small, oc_over=oc)
    B = 256
    batches = [synthetic.make_batch(cfg, fields, B, seed=s) for s in range(4)]
    for X, y in batches[:2]:
        a.train_on_batch(X, y)
    path = a.save_checkpoint(str(tmp_path / "ckpt.pt"))
    for X, y in batches[2:]:
        loss_a = a.train_on_batch(X, y)
    torch.cuda.synchronize()
    want = {k: v.detach().cpu().clone() for k, v in a.state_dict().items()}
    loss_a = loss_a.cpu().clone()
    b, _, _, _, _ = build_pair(workload, kw, precision, init_std=0.05, seed=123, mc_over=small, oc_over=oc)  # other weights
    b.load_checkpoint(str(tmp_path / "ckpt.pt"))
    for X, y in batches[2:]:
        loss_b = b.train_on_batch(X, y)
    torch.cuda.synchronize()
    assert torch.equal(loss_b.cpu(), loss_a)
    got = b.state_dict()
    for k, v in want.items():
        assert torch.equal(got[k].cpu(), v), k
    for name in ("dense_s1", "dense_s2", "emb_s1", "emb_s2"):
        sa, sb = getattr(a.store, name, None), getattr(b.store, name, None)
        if sa is not None:
            assert torch.equal(sa, sb), name
