"""GPU: device-side metrics (mmlrec_b200/metrics.py) against the sklearn functions the reference calls."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,ties", [(50, False), (4096, False), (4096, True), (100_000, True)])
def test_device_metrics_match_sklearn(n, ties):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from sklearn.metrics import accuracy_score, log_loss, mean_squared_error, roc_auc_score
    from mmlrec_b200 import metrics as M
    rng = np.random.default_rng(n + ties)
    y = (rng.random((n, 3)) < 0.3).astype(np.float32)
    p = rng.random((n, 3)).astype(np.float32)
    if ties:
        p = np.round(p, 2)     # many equal scores: tie groups
    p = np.clip(p * 0.6 + 0.3 * y, 0, 1).astype(np.float32)
    yd, pd_ = torch.from_numpy(y).cuda(), torch.from_numpy(p).cuda()
    p64 = p.astype("float64")
    assert abs(M.roc_auc_score(yd, pd_) - roc_auc_score(y, p64)) < 1e-9                     # multilabel, macro average
    assert abs(M.roc_auc_score(yd[:, 0], pd_[:, 0]) - roc_auc_score(y[:, 0], p64[:, 0])) < 1e-9
    assert abs(M.accuracy_score(yd, pd_) - accuracy_score(y, np.where(p64 > 0.5, 1, 0))) < 1e-12
    assert abs(M.log_loss(yd[:, 1], pd_[:, 1]) - log_loss(y[:, 1], p64[:, 1])) < 1e-9
    assert abs(M.mean_squared_error(yd, pd_) - mean_squared_error(y, p64)) < 1e-9


def test_auc_of_a_single_class_raises_like_sklearn():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from mmlrec_b200 import metrics as M
    with pytest.raises(ValueError):
        M.roc_auc_score(torch.ones(10, device="cuda"), torch.rand(10, device="cuda"))
