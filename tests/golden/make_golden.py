"""Generate the golden vectors that pin ``oracle/mmlrec_oracle.py`` to the reference.

Runs ONLY in the build container (needs ``/root/reference``, which is imported read-only and
never copied).  For each case it builds the reference model class from a synthetic config of the
BASELINE shape (widths shrunk so fixtures stay small), then executes the reference's own step
body -- ``model/basemodel.py:262-313``: ``model(x, None).squeeze()`` -> ``optim.zero_grad()`` ->
sum of ``F.binary_cross_entropy(.., reduction='sum')`` -> ``+ reg + aux + zeros`` -> ``backward``
-> ``optim.step()`` -- for a few steps on seeded inputs and stores inputs, initial parameters,
per-step predictions / losses, first-step gradients and the final parameters/buffers.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz
"""
import io
import contextlib
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from mmlrec_b200 import synthetic  # noqa: E402

STEPS = 3
BATCH = 48

# case -> (workload name, workload kwargs, model_config overrides, optim overrides)
SMALL = dict(expert_dnn_hidden_units=[16, 8], gate_dnn_hidden_units=[8], tower_dnn_hidden_units=[8],
             bottom_dnn_hidden_units=[16, 8], dnn_hidden_units=[16, 8])
# BatchNorm cases are built with init_std=0.05 (the constructor argument every reference model takes):
# at the default 1e-4 the Linear output is "bias + 1e-5-scale signal", `z - mean` keeps 3-4 digits and
# any 1-ulp difference in GEMM summation order moves gradients by ~1e-2 (true for torch-CPU vs
# torch-CUDA as well), so such a case can only pin predictions / losses; see *_default_init below.
INIT_STD = {"mmoe_census_bn_adam": 0.05, "mmoe_census_bn_adagrad": 0.05}
CASES = {
    "mmoe_census_bn_adam": ("census_mmoe", {}, dict(expert_dnn_hidden_units=[16], gate_dnn_hidden_units=[16],
                                                    tower_dnn_hidden_units=[16]), {}),
    "mmoe_census_bn_adagrad": ("census_mmoe", {}, dict(expert_dnn_hidden_units=[16], gate_dnn_hidden_units=[16],
                                                       tower_dnn_hidden_units=[16]), dict(optimizer="adagrad", lr=1e-2)),
    "mmoe_census_bn_default_init_adam": ("census_mmoe", {}, dict(expert_dnn_hidden_units=[16], gate_dnn_hidden_units=[16],
                                                                 tower_dnn_hidden_units=[16]), {}),
    "ple_ae_t4_adam": ("ae_ple_t4", dict(max_vocab=300), SMALL, {}),
    "ple_ae_t2_adam": ("ae_ple_t2", dict(max_vocab=300), SMALL, {}),
    "sharedbottom_kuairec_adam": ("kuairec_sharedbottom", dict(max_vocab=200), SMALL, {}),
    "esmm_kuairec_adam": ("kuairec_esmm", dict(max_vocab=200), SMALL, {}),
    "star_movielens_adam": ("movielens_star", dict(vocab_scale=0.02), dict(dnn_hidden_units=[16, 16]), {}),
    "pepnet_movielens_adam": ("movielens_pepnet", dict(vocab_scale=0.02), dict(dnn_hidden_units=[16, 16]), {}),
    "mmoe_synth26_adagrad": ("synth26_mmoe", dict(vocab=97), SMALL, {}),
    "ple_ae_t2_l3_adam": ("ae_ple_t2", dict(max_vocab=300), dict(SMALL, num_levels=3), {}),
    "sharedbottom_kuairec_adagrad": ("kuairec_sharedbottom", dict(max_vocab=200), SMALL, dict(optimizer="adagrad", lr=1e-2)),
    "mmoe_synth26_adam": ("synth26_mmoe", dict(vocab=97), SMALL, dict(optimizer="adam", lr=1e-3)),
    "sharedbottom_kuairec_sgd": ("kuairec_sharedbottom", dict(max_vocab=200), SMALL, dict(optimizer="sgd", lr=1e-2)),
    "esmm_kuairec_rmsprop": ("kuairec_esmm", dict(max_vocab=200), SMALL, dict(optimizer="rmsprop", lr=1e-3)),
    "mmoe_nogate_notower_adam": ("movielens_star", dict(vocab_scale=0.02),
                                 dict(model_name="mmoe", expert_dnn_hidden_units=[16, 16]), {}),
    # regression head: PredictionLayer('regression') is the identity and the task's loss is F.mse_loss(sum)
    # (model/utils.py:246-247, model/basemodel.py:595-604); labels of task 1 are continuous
    "sharedbottom_kuairec_mse_adam": ("kuairec_sharedbottom", dict(max_vocab=200),
                                      dict(SMALL, task_types=["binary", "regression"]),
                                      dict(loss=["binary_crossentropy", "mse"])),
    # saturated logits: head bias 30 -> sigmoid == 1.0f: BCE-on-probabilities gives loss 100 per
    # negative sample (log clamp at -100) and a ZERO gradient, unlike bce_with_logits
    "sharedbottom_kuairec_saturated_sgd": ("kuairec_sharedbottom", dict(max_vocab=200), SMALL,
                                           dict(optimizer="sgd", lr=1e-2)),
    # model zoo on the same stages: one shared final layer (MLP), identity-initialised cross-stitch
    # units applied as x @ W (CrossStitch), detached cross-task tower mixing (HMoE)
    "mlp_kuairec_adam": ("kuairec_sharedbottom", dict(max_vocab=200), dict(SMALL, model_name="mlp"), {}),
    "cross_stitch_kuairec_adam": ("kuairec_sharedbottom", dict(max_vocab=200),
                                  dict(SMALL, model_name="cross_stitch", shared_hidden_unit=16), {}),
    "hmoe_kuairec_adam": ("kuairec_sharedbottom", dict(max_vocab=200),
                          dict(SMALL, model_name="hmoe", task_weight_hidden_units=[8]), {}),
    # 'pcg' = MMOE whose optimizer is wrapped in PCGrad (main.py:53-54, basemodel.py:564-565); the step goes through
    # optim.pc_backward(total_loss) (basemodel.py:309-310) -- one objective, so no projection ever happens
    "pcg_kuairec_adam": ("kuairec_sharedbottom", dict(max_vocab=200), dict(SMALL, model_name="pcg"), {}),
    # l2_reg_dnn > 0 (basemodel.py:514-540): reg gradient 2*l2*w on the weights each model registers, also on PLE's
    # allocated-but-unused shared experts
    "mmoe_kuairec_l2_adam": ("kuairec_sharedbottom", dict(max_vocab=200), dict(SMALL, model_name="mmoe", l2_reg_dnn=1e-2), {}),
    "ple_ae_t2_l2_sgd": ("ae_ple_t2", dict(max_vocab=300), dict(SMALL, l2_reg_dnn=1e-2), dict(optimizer="sgd", lr=1e-2)),
    "esmm_kuairec_l2_adagrad": ("kuairec_esmm", dict(max_vocab=200), dict(SMALL, l2_reg_dnn=1e-2),
                                dict(optimizer="adagrad", lr=1e-2)),
    # the scenario mask the reference is written for but never passes (basemodel.py:265-266 sets it to None): here the
    # reference's own modules are called WITH the mask (model(x, domain_mask), mmoe.py:101-106) and the loss is the
    # masked branch of its loop (basemodel.py:273-282), computed by this script
    "ple_ae_t4_masked_adam": ("ae_ple_t4", dict(max_vocab=300), SMALL, {}),
    "mmoe_movielens_masked_adam": ("movielens_star", dict(vocab_scale=0.02),
                                   dict(model_name="mmoe", expert_dnn_hidden_units=[16, 16]), {}),
}
MASKED = {"ple_ae_t4_masked_adam", "mmoe_movielens_masked_adam"}
# ESCM-IPW: three-column output, the loop's special loss (basemodel.py:284-292)
CASES["escm_kuairec_adam"] = ("kuairec_esmm", dict(max_vocab=200), dict(SMALL, model_name="escm"), {})
CASES["escm_kuairec_sgd"] = ("kuairec_esmm", dict(max_vocab=200), dict(SMALL, model_name="escm"), dict(optimizer="sgd", lr=1e-3))
INIT_STD.update({"escm_kuairec_adam": 0.05, "escm_kuairec_sgd": 0.05})
# AITM (aitm.py): h1 / h2 / h3 applied to both tokens (shared weights), feat_0 read by g and by tower 0
CASES["aitm_kuairec_adam"] = ("kuairec_sharedbottom", dict(max_vocab=200), dict(SMALL, model_name="aitm"), {})
CASES["aitm_kuairec_notower_l2_sgd"] = ("kuairec_sharedbottom", dict(max_vocab=200),
                                        dict(SMALL, model_name="aitm", expert_dnn_hidden_units=[32, 40],
                                             tower_dnn_hidden_units=[], l2_reg_dnn=1e-2), dict(optimizer="sgd", lr=1e-2))
INIT_STD.update({"aitm_kuairec_adam": 0.05, "aitm_kuairec_notower_l2_sgd": 0.05})
# SNR-trans (snr_trans.py): hard-concrete gates over unregistered (constant) transformation matrices
CASES["snr_trans_kuairec_adam"] = ("kuairec_sharedbottom", dict(max_vocab=200), dict(SMALL, model_name="snr_trans"), {})
CASES["snr_trans_kuairec_1level_sgd"] = ("kuairec_sharedbottom", dict(max_vocab=200),
                                         dict(SMALL, model_name="snr_trans", expert_dnn_hidden_units=[24], num_experts=3,
                                              tower_dnn_hidden_units=[]), dict(optimizer="sgd", lr=1e-2))
INIT_STD.update({"snr_trans_kuairec_adam": 0.05, "snr_trans_kuairec_1level_sgd": 0.05})
# MSSM (mssm.py): SNR-trans's structure with a hard-concrete gate per output unit; u AND the matrices are unregistered
CASES["mssm_kuairec_adam"] = ("kuairec_sharedbottom", dict(max_vocab=200), dict(SMALL, model_name="mssm"), {})
CASES["mssm_kuairec_1level_l2_sgd"] = ("kuairec_sharedbottom", dict(max_vocab=200),
                                       dict(SMALL, model_name="mssm", expert_dnn_hidden_units=[24], num_experts=3,
                                            tower_dnn_hidden_units=[16], l2_reg_dnn=1e-2), dict(optimizer="sgd", lr=1e-2))
INIT_STD.update({"mssm_kuairec_adam": 0.05, "mssm_kuairec_1level_l2_sgd": 0.05})
# APG (apg.py): per-sample k x k matrices generated from the detached scene embedding between two shared low-rank maps;
# the generating DNNs are built with the DNN default init (1e-4), so the seeded state is perturbed after construction
CASES["apg_movielens_adam"] = ("movielens_star", dict(vocab_scale=0.02), dict(model_name="apg", dnn_hidden_units=[16, 16]), {})
CASES["apg_movielens_odd_sgd"] = ("movielens_star", dict(vocab_scale=0.02), dict(model_name="apg", dnn_hidden_units=[24, 10]),
                                  dict(optimizer="sgd", lr=1e-2))
INIT_STD.update({"apg_movielens_adam": 0.05, "apg_movielens_odd_sgd": 0.05})
# BatchNorm variants added after the round's GPU budget was spent: checked on the CPU only (oracle + plan emulation), kept in
# tests/golden/cpu_only/ so that the -m gpu step tests do not enumerate them.  PLE with BatchNorm: the reference keeps running
# the dead last-level shared-gate DNN, whose running statistics therefore move; CrossStitch with BatchNorm (census shape).
CASES["ple_census_bn_adam"] = ("census_mmoe", {}, dict(model_name="ple", expert_dnn_hidden_units=[16, 8],
                                                       gate_dnn_hidden_units=[8], tower_dnn_hidden_units=[8],
                                                       shared_expert_num=2, specific_expert_num=2, num_levels=2), {})
CASES["cross_stitch_census_bn_adam"] = ("census_mmoe", {}, dict(model_name="cross_stitch", shared_hidden_unit=16,
                                                                dnn_hidden_units=[16, 8], tower_dnn_hidden_units=[8]), {})
CPU_ONLY = {"ple_census_bn_adam", "cross_stitch_census_bn_adam"}
INIT_STD.update({"ple_census_bn_adam": 0.05, "cross_stitch_census_bn_adam": 0.05})
# cases whose identity / 1e-4 initial state would leave parts of the model untested: perturbed after construction
INIT_STD.update({"cross_stitch_kuairec_adam": 0.05, "hmoe_kuairec_adam": 0.05, "mlp_kuairec_adam": 0.05,
                 "pcg_kuairec_adam": 0.05, "mmoe_kuairec_l2_adam": 0.05, "ple_ae_t2_l2_sgd": 0.05,
                 "esmm_kuairec_l2_adagrad": 0.05, "ple_ae_t4_masked_adam": 0.05, "mmoe_movielens_masked_adam": 0.05})


def post_build(case, model):
    """Edits of the seeded initial state; returns the names touched (stored as meta/init_overridden)."""
    if case == "sharedbottom_kuairec_saturated_sgd":
        with torch.no_grad():
            model.out[0].bias.fill_(30.0)
        return ["out.0.bias"]
    if case.startswith("apg_"):   # generated matrices / biases that really depend on the scene embedding
        touched = []
        with torch.no_grad():
            for name, prm in model.named_parameters():
                if ".specific_" in name and name.endswith(".weight"):
                    prm.add_(0.5 * torch.randn(prm.shape, generator=torch.Generator().manual_seed(11)))
                    touched.append(name)
                if name.endswith("shared_bias_nk") or name.endswith("shared_bias_km"):
                    prm.add_(0.05 * torch.randn(prm.shape, generator=torch.Generator().manual_seed(12)))
                    touched.append(name)
        return touched
    if case in ("cross_stitch_kuairec_adam", "cross_stitch_census_bn_adam"):   # identity units exercise no off-diagonal weight: add seeded noise
        touched = []
        with torch.no_grad():
            for name, prm in model.named_parameters():
                if name.endswith("cross_stitch_weight"):
                    prm.add_(0.1 * torch.randn(prm.shape, generator=torch.Generator().manual_seed(7)))
                    touched.append(name)
        return touched
    return []


def make_labels(case, y, seed):
    if case == "sharedbottom_kuairec_mse_adam":
        y = y.copy()
        y[:, 1] = np.random.default_rng(5000 + seed).normal(0.3, 1.0, size=len(y)).astype(np.float32)
    return y


def build_reference(cfg, fields, init_std=0.0001):
    from model.utils import SparseFeat, DenseFeat
    from model.mmoe import MMOE
    from model.ple import PLE
    from model.sharedbottom import SharedBottom
    from model.esmm import ESMM
    from model.star import STAR
    from model.pepnet import PepNet
    from model.mlp import MLP
    from model.cross_stitch import CrossStitch
    from model.hmoe import HMOE
    from model.escm import ESCM
    from model.aitm import AITM
    from model.snr_trans import SNR_trans
    from model.mssm import MSSM
    from model.apg import APG
    emb = cfg["model_config"]["emb"]
    cols = [SparseFeat(n, vocabulary_size=v, embedding_dim=emb) if k == "sparse" else DenseFeat(n, 1)
            for n, k, v in fields]
    cls = {"mmoe": MMOE, "ple": PLE, "sharedbottom": SharedBottom, "esmm": ESMM, "star": STAR,
           "pepnet": PepNet, "mlp": MLP, "cross_stitch": CrossStitch, "hmoe": HMOE, "pcg": MMOE, "escm": ESCM, "aitm": AITM, "snr_trans": SNR_trans,
           "mssm": MSSM, "apg": APG}[cfg["model_config"]["model_name"].lower()]
    with contextlib.redirect_stdout(io.StringIO()):
        model = cls(cols, init_std=init_std, device="cpu", config=cfg)
        model.compile(optimizer=cfg["optim_config"]["optimizer"], loss=cfg["optim_config"]["loss"],
                      metrics=cfg["optim_config"]["metrics"])
    return model


def unregistered_star_tensors(model):
    """SharedSpecificLinear keeps all but the last per-domain weight outside ``named_parameters``
    (model/utils.py:181-191); export them so the oracle can treat them as constants."""
    extra = {}
    if type(model).__name__ == "SNR_trans":   # gate.trans_matrix: a plain list of lists of Parameters (snr_trans.py:31-34)
        for name, mod in model.trans.items():
            if name.startswith("gate"):
                for i, row in enumerate(mod.trans_matrix):
                    for j, m in enumerate(row):
                        extra[f"trans.{name}.trans_matrix.{i}.{j}"] = m.detach().clone()
        return extra
    if type(model).__name__ == "MSSM":   # gate.u AND gate.trans_matrix: plain lists of lists (mssm.py:26-36)
        for name, mod in model.mssm.items():
            if name.startswith("gate"):
                for attr in ("trans_matrix", "u"):
                    for i, row in enumerate(getattr(mod, attr)):
                        for j, m in enumerate(row):
                            extra[f"mssm.{name}.{attr}.{i}.{j}"] = m.detach().clone()
        return extra
    if type(model).__name__ != "STAR":
        return extra
    for prefix, mods in (("linears", model.linears), ("final_layers", model.final_layers)):
        for j, lin in enumerate(mods):
            n = len(lin.specific_weights)
            for i in range(n - 1):
                extra[f"{prefix}.{j}.specific_weights.{i}"] = lin.specific_weights[i].detach().clone()
                extra[f"{prefix}.{j}.specific_biases.{i}"] = lin.specific_biases[i].detach().clone()
    return extra


def domain_mask_of(cfg, fields, X):
    """basemodel.py:152-161: get_mask(values of data_config['mask_column'], mask_values, num_domains)."""
    from model.utils import get_mask
    dc = cfg["data_config"]
    col = [n for n, _, _ in fields].index(dc["mask_column"])
    return get_mask(list(X[:, col]), dc["mask_values"], dc["num_domains"]).float()


def reference_step(model, X, y, domain_mask=None):
    """model/basemodel.py:262-313 verbatim in effect (domain_mask is always None, :265-266); with a mask: the branch of
    the same loop that the unconditional None makes unreachable (:273-282)."""
    x = X.float()
    y = y.float()
    y_pred = model(x, domain_mask).squeeze()
    model.optim.zero_grad()
    # model.loss_func is the list compile() built through _get_loss_func_single (basemodel.py:595-604)
    if domain_mask is not None:
        D = model.num_domains
        loss = sum(model.loss_func[i](y_pred[:, i], y[:, i], weight=domain_mask[:, i % D], reduction="sum")
                   for i in range(model.num_tasks))
    elif model.model_config["model_name"] == "escm":   # basemodel.py:284-292, line for line
        loss_func = model.loss_func
        loss_0 = loss_func[0](y_pred[:, 0], y[:, 0], reduction="sum")
        loss_1 = loss_func[1](y_pred[:, 1], y[:, 1], reduction="sum")
        loss_2 = loss_func[1](y_pred[:, 2], y[:, 1], reduction="sum")
        ctr_num = torch.sum(y[:, 0])
        o = y[:, 0].float()
        loss_1 = model.counterfact_ipw(loss_1, ctr_num, o, y_pred[:, 0])
        loss = loss_0 + loss_1 * model.counterfactual_w + loss_2 * model.global_w
    else:
        loss = sum(model.loss_func[i](y_pred[:, i], y[:, i], reduction="sum") for i in range(model.num_tasks))
    total = loss + model.get_regularization_loss() + model.aux_loss + torch.zeros((1,))
    if model.model_config["model_name"] == "pcg":
        model.optim.pc_backward(total)
    else:
        total.backward()
    grads = {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in model.named_parameters()}
    model.optim.step()
    return y_pred.detach(), loss.detach(), grads


def main():
    only = set(sys.argv[1:])   # optional: regenerate only the named cases
    for case, (wl, kw, mc_over, oc_over) in CASES.items():
        if only and case not in only:
            continue
        cfg, fields = synthetic.workload(wl, **kw)
        cfg["model_config"].update(mc_over)
        cfg["optim_config"].update(oc_over)
        torch.manual_seed(1234)
        np.random.seed(1234)
        model = build_reference(cfg, fields, INIT_STD.get(case, 0.0001))
        overridden = post_build(case, model)
        model.train()
        blob = {"meta/init_overridden": np.array(overridden, dtype=str)}
        for k, v in model.state_dict().items():
            blob["init/" + k] = v.detach().numpy().copy()
        for k, v in unregistered_star_tensors(model).items():
            blob["init/" + k] = v.numpy().copy()
        blob["meta/trainable"] = np.array([n for n, _ in model.named_parameters()])
        blob["meta/buffers"] = np.array([n for n, _ in model.named_buffers()])
        for s in range(STEPS):
            X, y = synthetic.make_batch(cfg, fields, BATCH, seed=100 + s)
            y = make_labels(case, y, s)
            blob[f"step{s}/X"], blob[f"step{s}/y"] = X, y
            dm = domain_mask_of(cfg, fields, X) if case in MASKED else None
            if dm is not None:
                blob[f"step{s}/mask"] = dm.numpy()
            pred, loss, grads = reference_step(model, torch.from_numpy(X), torch.from_numpy(y), dm)
            blob[f"step{s}/pred"] = pred.numpy()
            blob[f"step{s}/loss"] = loss.numpy()
            if s == 0:
                for n, g in grads.items():
                    if g is not None:
                        blob["grad0/" + n] = g.numpy()
                blob["meta/gradless"] = np.array([n for n, g in grads.items() if g is None])
        model.eval()
        with torch.no_grad():
            Xe, _ = synthetic.make_batch(cfg, fields, BATCH, seed=999)
            blob["eval/X"] = Xe
            dme = domain_mask_of(cfg, fields, Xe) if case in MASKED else None
            if dme is not None:
                blob["eval/mask"] = dme.numpy()
            blob["eval/pred"] = model(torch.from_numpy(Xe).float(), dme).numpy()
        for k, v in model.state_dict().items():
            blob["final/" + k] = v.detach().numpy().copy()
        import json
        blob["meta/init_std"] = np.array(INIT_STD.get(case, 0.0001))
        blob["meta/config"] = np.array(json.dumps(cfg))
        blob["meta/fields"] = np.array(json.dumps(fields))
        path = os.path.join(HERE, "cpu_only", case + ".npz") if case in CPU_ONLY else os.path.join(HERE, case + ".npz")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.savez_compressed(path, **blob)
        print(f"{case:32s} {os.path.getsize(path) / 1024:8.1f} KiB  loss0={float(blob['step0/loss']):.6f}")


if __name__ == "__main__":
    main()
