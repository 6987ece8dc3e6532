"""GPU: training dynamics parity.  Train the same model from the same weights on the same batches with the
CUDA path (public API: compile / train_on_batch / predict) and with the oracle, on synthetic data whose labels
depend on the features (so AUC moves away from 0.5), and compare per-task AUC on held-out rows.
BASELINE.json: per-task AUC after N steps within 1e-3."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import oracle_columns  # noqa: E402


def _learnable_batch(cfg, fields, n, seed):
    from mmlrec_b200 import synthetic
    X, y = synthetic.make_batch(cfg, fields, n, seed=seed)
    rng = np.random.default_rng(1000 + seed)
    # labels driven by the parity of two id columns and one more column (learnable through the embeddings)
    s = ((X[:, 0].astype(np.int64) % 2) * 1.5 + (X[:, 1].astype(np.int64) % 3 == 0) * 1.0 - 1.2)
    for t in range(y.shape[1]):
        shift = 0.6 * ((X[:, 2 + t % 3].astype(np.int64) % 2) - 0.5)
        p = 1 / (1 + np.exp(-(s + shift)))
        y[:, t] = (rng.random(n) < p).astype(np.float32)
    return X, y


@pytest.mark.parametrize("wl,kw,precision,tol", [
    ("ae_ple_t4", dict(max_vocab=2000), "fp32", 1e-3),
    ("ae_ple_t4", dict(max_vocab=2000), "bf16", 5e-3),
    ("kuairec_esmm", dict(max_vocab=500), "fp32", 1e-3),
    ("movielens_star", dict(vocab_scale=0.05), "fp32", 1e-3),
])
def test_auc_after_training_matches_oracle(wl, kw, precision, tol):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from sklearn.metrics import roc_auc_score
    from mmlrec_b200 import synthetic
    from mmlrec_b200.model import get_model_class
    from mmlrec_b200.model.utils import DenseFeat, SparseFeat
    from oracle.mmlrec_oracle import OracleTrainer
    cfg, fields = synthetic.workload(wl, **kw)
    cfg["model_config"].update(expert_dnn_hidden_units=[32, 16], gate_dnn_hidden_units=[16], tower_dnn_hidden_units=[16],
                               dnn_hidden_units=[32, 16])
    cfg["optim_config"]["lr"] = 5e-3
    cfg["b200_config"] = {"precision": precision}
    emb = cfg["model_config"]["emb"]
    cols = [SparseFeat(n, v, emb) if k == "sparse" else DenseFeat(n, 1) for n, k, v in fields]
    torch.manual_seed(7)
    model = get_model_class(cfg["model_config"]["model_name"])(cols, init_std=0.05, device="cuda:0", config=copy.deepcopy(cfg))
    model.compile(cfg["optim_config"]["optimizer"], cfg["optim_config"]["loss"], ["auc"])
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    names = [n for n, _ in model.named_parameters()]
    extra = {}
    if hasattr(model, "linears") and hasattr(model.linears[0], "spec_weights"):  # STAR's frozen per-domain tensors
        for prefix, mods in (("linears", model.linears), ("final_layers", model.final_layers)):
            for j, m in enumerate(mods):
                for i in range(m.num_domains - 1):
                    extra[f"{prefix}.{j}.specific_weights.{i}"] = getattr(m, f"frozen_weight_{i}").detach().cpu().clone()
                    extra[f"{prefix}.{j}.specific_biases.{i}"] = getattr(m, f"frozen_bias_{i}").detach().cpu().clone()
    params = {k: v for k, v in sd.items() if k in names}
    params.update(extra)
    oracle = OracleTrainer(cfg, oracle_columns(cfg, fields), params, {k: v for k, v in sd.items() if k not in names}, names)
    model.train()
    B, steps = 512, 40
    for s in range(steps):
        X, y = _learnable_batch(cfg, fields, B, seed=s)
        model.train_on_batch(X, y)
        oracle.step(torch.from_numpy(X), torch.from_numpy(y))
    Xt, yt = _learnable_batch(cfg, fields, 4096, seed=10_000)
    ours = model.predict(Xt, 1024)
    with torch.no_grad():
        ref = oracle.forward(torch.from_numpy(Xt), training=False).numpy()
    for t in range(yt.shape[1]):
        a, b = roc_auc_score(yt[:, t], ours[:, t]), roc_auc_score(yt[:, t], ref[:, t])
        assert b > 0.6, f"task {t}: the oracle itself did not learn (AUC {b:.3f})"
        assert abs(a - b) < tol, f"task {t}: AUC {a:.5f} vs oracle {b:.5f}"
