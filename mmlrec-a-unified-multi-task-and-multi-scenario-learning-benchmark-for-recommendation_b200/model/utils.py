"""Feature schema and building blocks with the reference's names and state_dict keys.

Mirrors the public surface of ``/root/reference/model/utils.py`` that the hot path touches:
``SparseFeat`` / ``DenseFeat`` / ``VarLenSparseFeat`` (:328-395), ``build_input_features``
(:407-431), ``get_feature_names``, ``create_embedding_matrix`` (:466-488), ``DNN`` (:92-161),
``PredictionLayer`` (:225-248), ``get_mask`` (:639-645).  The modules here only OWN parameters
(so checkpoints interchange with the reference); the arithmetic is done by the step program in
``engine/`` through the CUDA library, never by ``nn.Module.forward`` of these classes.

Construction order and init calls follow the reference exactly, so a model built under the same
``torch.manual_seed`` starts from bit-identical weights.
"""
from __future__ import annotations

from collections import OrderedDict, namedtuple
from typing import List, Sequence

import torch
import torch.nn as nn

DEFAULT_GROUP_NAME = "default_group"

_SparseBase = namedtuple("SparseFeat", ["name", "vocabulary_size", "embedding_dim", "use_hash", "dtype",
                                        "embedding_name", "group_name"])
_VarLenBase = namedtuple("VarLenSparseFeat", ["sparsefeat", "maxlen", "combiner", "length_name"])
_DenseBase = namedtuple("DenseFeat", ["name", "dimension", "dtype"])


class SparseFeat(_SparseBase):
    """A categorical column: one id per sample, looked up in a [vocabulary_size, embedding_dim] table."""
    __slots__ = ()

    def __new__(cls, name, vocabulary_size, embedding_dim=4, use_hash=False, dtype="int32", embedding_name=None,
                group_name=DEFAULT_GROUP_NAME):
        if embedding_dim == "auto":
            embedding_dim = 6 * int(pow(vocabulary_size, 0.25))
        return super().__new__(cls, name, vocabulary_size, embedding_dim, use_hash, dtype,
                               name if embedding_name is None else embedding_name, group_name)

    def __hash__(self):
        return hash(self.name)


class VarLenSparseFeat(_VarLenBase):
    """Kept for API compatibility; the data pipeline never produces one (utils/data_utils.py:73-75)
    and the step program rejects it."""
    __slots__ = ()

    def __new__(cls, sparsefeat, maxlen, combiner="mean", length_name=None):
        return super().__new__(cls, sparsefeat, maxlen, combiner, length_name)

    name = property(lambda self: self.sparsefeat.name)
    vocabulary_size = property(lambda self: self.sparsefeat.vocabulary_size)
    embedding_dim = property(lambda self: self.sparsefeat.embedding_dim)
    embedding_name = property(lambda self: self.sparsefeat.embedding_name)

    def __hash__(self):
        return hash(self.name)


class DenseFeat(_DenseBase):
    """A numeric column passed through unchanged."""
    __slots__ = ()

    def __new__(cls, name, dimension=1, dtype="float32"):
        return super().__new__(cls, name, dimension, dtype)

    def __hash__(self):
        return hash(self.name)


def build_input_features(feature_columns) -> "OrderedDict[str, tuple]":
    """name -> (start, end) span in X; first occurrence of a name wins."""
    spans, cursor = OrderedDict(), 0
    for fc in feature_columns:
        if fc.name in spans:
            continue
        if isinstance(fc, SparseFeat):
            width = 1
        elif isinstance(fc, DenseFeat):
            width = fc.dimension
        elif isinstance(fc, VarLenSparseFeat):
            raise NotImplementedError("variable-length sparse features are outside the B200 hot path")
        else:
            raise TypeError("Invalid feature column type,got", type(fc))
        spans[fc.name] = (cursor, cursor + width)
        cursor += width
    return spans


def get_feature_names(feature_columns) -> List[str]:
    return list(build_input_features(feature_columns).keys())


def create_embedding_matrix(feature_columns, init_std=0.0001, linear=False, sparse=False, device="cpu",
                            shard_world: int = 1) -> nn.ModuleDict:
    """One nn.Embedding per sparse column keyed by embedding_name; all tables are constructed first
    (default N(0,1) draw) and then re-drawn N(0, init_std) in a second pass, like the reference.
    ``shard_world > 1``: only this rank's shard (ceil(V / world) rows) is built; the draw then differs from
    the reference's RNG stream (same distribution) -- parity tests load full tables explicitly."""
    sparse_cols = [fc for fc in feature_columns if isinstance(fc, SparseFeat)]
    tables = nn.ModuleDict()
    for fc in sparse_cols:
        if shard_world > 1:
            emb = nn.Embedding((fc.vocabulary_size + shard_world - 1) // shard_world, 1 if linear else fc.embedding_dim,
                               sparse=sparse, _weight=torch.empty((fc.vocabulary_size + shard_world - 1) // shard_world,
                                                                  1 if linear else fc.embedding_dim))
            tables[fc.embedding_name] = emb
            continue
        tables[fc.embedding_name] = nn.Embedding(fc.vocabulary_size, 1 if linear else fc.embedding_dim, sparse=sparse)
    for emb in tables.values():
        nn.init.normal_(emb.weight, mean=0, std=init_std)
    return tables


class DNN(nn.Module):
    """Parameter container of an MLP block: ``linears`` (+ ``bn``) with the reference's init
    (weights N(0, init_std); biases keep nn.Linear's default draw)."""

    def __init__(self, inputs_dim, hidden_units, activation="relu", l2_reg=0, dropout_rate=0, use_bn=False,
                 init_std=0.0001, dice_dim=3, device="cpu"):
        super().__init__()
        if len(hidden_units) == 0:
            raise ValueError("hidden_units is empty!!")
        if dropout_rate and dropout_rate > 0:
            raise NotImplementedError("dnn_dropout > 0 is not supported by the fused step (all shipped configs use 0)")
        if activation not in ("relu", None):
            # 'sigmoid' / 'prelu' hidden layers: the fused gradient writers fold only the ReLU mask (and GateNN's
            # 2*sigmoid) into their epilogues; no shipped config asks for anything else (dnn_activation = relu)
            raise NotImplementedError(f"activation {activation!r} is not supported by the fused step")
        self.activation, self.use_bn, self.l2_reg, self.dropout_rate = activation, use_bn, l2_reg, dropout_rate
        dims = [inputs_dim] + list(hidden_units)
        self.linears = nn.ModuleList(nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1))
        if use_bn:
            self.bn = nn.ModuleList(nn.BatchNorm1d(dims[i + 1]) for i in range(len(dims) - 1))
        for name, tensor in self.linears.named_parameters():
            if "weight" in name:
                nn.init.normal_(tensor, mean=0, std=init_std)

    def layer(self, i: int):
        return self.linears[i], (self.bn[i] if self.use_bn else None)

    def forward(self, inputs):  # pragma: no cover - the step program owns the arithmetic
        raise RuntimeError("DNN.forward is not used: the fused CUDA step program runs the layers")


class PredictionLayer(nn.Module):
    """Scalar bias (+ sigmoid for 'binary'); applied inside the fused head kernel."""

    def __init__(self, task="binary", use_bias=True, **kwargs):
        if task not in ["binary", "multiclass", "regression"]:
            raise ValueError("task must be binary,multiclass or regression")
        super().__init__()
        self.task, self.use_bias = task, use_bias
        if use_bias:
            self.bias = nn.Parameter(torch.zeros((1,)))


def get_mask(domain_values, mask_values, num_domains) -> torch.Tensor:
    """[N, num_domains] int mask: sample n belongs to domain d iff its scene value == mask_values[d]."""
    dv = torch.as_tensor(list(domain_values), dtype=torch.float32).reshape(-1, 1)
    mv = torch.as_tensor(list(mask_values), dtype=torch.float32).reshape(1, -1)
    assert mv.shape[1] == num_domains
    return (dv == mv).int()
