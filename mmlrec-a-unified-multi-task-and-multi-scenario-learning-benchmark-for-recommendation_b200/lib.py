"""ctypes binding of ``libmmlrec_b200.so`` (the C ABI declared in ``include/mmlrec_b200.h``).

There is no fallback: if the library cannot be loaded every compute entry point raises.
Structures mirror the C declarations field for field; tables of them are serialised to bytes and
shipped to the device once, at plan time.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmmlrec_b200.so")

MAX_GATE_EXPERTS = 32
MAX_TASKS = 16

ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SIGMOID2 = 0, 1, 2, 3
OPT_SGD, OPT_ADAGRAD, OPT_ADAM, OPT_RMSPROP = 0, 1, 2, 3
HEAD_SIGMOID_BCE, HEAD_IDENTITY_MSE = 0, 1
ACT_CODES = {None: ACT_NONE, "none": ACT_NONE, "relu": ACT_RELU, "sigmoid": ACT_SIGMOID, "sigmoid2": ACT_SIGMOID2}
OPT_CODES = {"sgd": OPT_SGD, "adagrad": OPT_ADAGRAD, "adam": OPT_ADAM, "rmsprop": OPT_RMSPROP}

vp, i32, i64, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double


class Hyper(C.Structure):
    _fields_ = [("step", i32), ("optimizer", i32), ("lr", f32), ("beta1", f32), ("beta2", f32), ("eps", f32),
                ("step_size", f32), ("bc2_sqrt", f32), ("alpha", f32), ("one_minus_beta1", f32),
                ("one_minus_beta2", f32), ("one_minus_alpha", f32), ("lr_d", f64), ("beta1_d", f64), ("beta2_d", f64)]


class GemmF32(C.Structure):
    _fields_ = [("A", vp), ("B", vp), ("C", vp), ("a_rs", i64), ("a_cs", i64), ("b_rs", i64), ("b_cs", i64),
                ("ldc", i64), ("bias", vp), ("mask", vp), ("ldmask", i64), ("rowsum_a", vp),
                ("M", i32), ("N", i32), ("K", i32), ("act", i32), ("accumulate", i32), ("reserved", i32)]


class GemmTcDesc(C.Structure):
    _fields_ = [("A", vp), ("B", vp), ("lda", i64), ("ldb", i64), ("a_mn_major", i32), ("b_mn_major", i32),
                ("M", i32), ("N", i32), ("K", i32), ("C_f32", vp), ("ldc_f32", i64), ("C_bf16", vp),
                ("ldc_bf16", i64), ("bias", vp), ("mask", vp), ("ldmask", i64), ("colsum", vp),
                ("act", i32), ("accumulate", i32),
                ("relu_bits_out", vp), ("bits_out_chunks", i32), ("bits_out_chunk0", i32),
                ("mask_bits", vp), ("mask_bits_chunks", i32), ("mask_bits_chunk0", i32),
                ("c_transposed", i32), ("pad1", i32), ("colsum_b", vp)]


def split_two_output_problems(descs):
    """The CTA-pair kernel writes one output precision per problem: a problem that wants fp32 AND bf16 copies of its
    result is listed twice (the bf16 copy keeps the ReLU bit output)."""
    out = []
    for d in descs:
        if d.C_f32 and d.C_bf16:
            a, b = GemmTcDesc.from_buffer_copy(bytes(d)), GemmTcDesc.from_buffer_copy(bytes(d))
            a.C_bf16, a.ldc_bf16, a.relu_bits_out = None, 0, None
            b.C_f32, b.ldc_f32, b.colsum = None, 0, None
            out += [a, b]
        else:
            out.append(d)
    return out


class Gate(C.Structure):
    _fields_ = [("gate_in", vp), ("ld_gate_in", i64), ("Hg", i32), ("n_e", i32), ("Wg", vp), ("ld_Wg", i64),
                ("expert", vp * MAX_GATE_EXPERTS), ("ld_expert", i64), ("H", i32), ("pad0", i32),
                ("probs", vp), ("mix", vp), ("ld_mix", i64), ("mix_bf16", vp), ("ld_mix_bf16", i64),
                ("d_mix", vp), ("ld_d_mix", i64), ("d_gate_in", vp), ("ld_d_gate_in", i64),
                ("relu_mask_gate_in", i32), ("accumulate_d_gate_in", i32),
                ("d_gate_in_bf16", vp), ("ld_d_gate_in_bf16", i64), ("dWg", vp)]


class ExpertGrad(C.Structure):
    _fields_ = [("expert", vp), ("ld_expert", i64), ("d_expert", vp), ("ld_d_expert", i64), ("H", i32),
                ("n_users", i32), ("user_probs", vp * (MAX_TASKS + 1)), ("user_prob_ld", i32 * (MAX_TASKS + 1)),
                ("user_prob_col", i32 * (MAX_TASKS + 1)), ("user_d_mix", vp * (MAX_TASKS + 1)),
                ("user_d_mix_ld", i64 * (MAX_TASKS + 1)), ("relu_mask", i32), ("pad0", i32),
                ("d_expert_bf16", vp), ("ld_d_expert_bf16", i64)]


LEVEL_MAX_GATES, LEVEL_MAX_EXPERTS, LEVEL_MAX_WG = 8, 32, 3072


class GateLevel(C.Structure):
    _G, _E = LEVEL_MAX_GATES, LEVEL_MAX_EXPERTS
    _fields_ = [("n_gates", i32), ("n_experts", i32), ("H", i32), ("expert_relu", i32),
                ("expert", vp * _E), ("ld_expert", i64), ("d_expert", vp * _E), ("ld_d_expert", i64),
                ("d_expert_bf16", vp * _E), ("ld_d_expert_bf16", i64), ("slot", (C.c_int8 * _G) * _E),
                ("gate_in", vp * _G), ("ld_gate_in", i64 * _G), ("Wg", vp * _G), ("ld_Wg", i64 * _G),
                ("Hg", i32 * _G), ("n_e", i32 * _G), ("probs", vp * _G), ("mix", vp * _G), ("ld_mix", i64 * _G),
                ("mix_bf16", vp * _G), ("ld_mix_bf16", i64 * _G), ("d_mix", vp * _G), ("ld_d_mix", i64 * _G),
                ("d_gate_in", vp * _G), ("ld_d_gate_in", i64 * _G), ("d_gate_in_bf16", vp * _G),
                ("ld_d_gate_in_bf16", i64 * _G), ("relu_mask_gate_in", i32 * _G), ("accumulate_d_gate_in", i32 * _G),
                ("dWg", vp * _G), ("detach_mask", C.c_uint32 * _E)]


class Head(C.Structure):
    _fields_ = [("h", vp), ("ld_h", i64), ("H", i32), ("kind", i32), ("w", vp), ("bias", vp),
                ("d_h", vp), ("ld_d_h", i64), ("relu_mask", i32), ("mask_col", i32), ("dw", vp), ("dbias", vp),
                ("d_h_bf16", vp), ("ld_d_h_bf16", i64), ("bias2", vp), ("dbias2", vp)]


_SIGNATURES = {
    "mmlrec_abi_version": (C.c_int, []),
    "mmlrec_last_error": (C.c_char_p, []),
    "mmlrec_launch_count": (i64, []),
    "mmlrec_struct_size": (i64, [i32]),
    "mmlrec_hyper_advance": (C.c_int, [vp, vp]),
    "mmlrec_hyper_advance_hist": (C.c_int, [vp, vp, i32, vp]),
    "mmlrec_emb_adam_catch_up": (C.c_int, [vp, i64, i32, vp, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp]),
    "mmlrec_emb_adam_catch_up_keys": (C.c_int, [vp, i32, vp, i32, i32, vp, vp, vp, vp, vp, i32, vp, i32, vp]),
    "mmlrec_emb_adam_flush": (C.c_int, [vp, vp, vp, vp, i64, i32, vp, vp, i32, vp]),
    "mmlrec_gather_concat": (C.c_int, [vp, i64, i32, vp, vp, i32, i32, vp, i32, i32, vp, i64, vp, i64, vp, vp]),
    "mmlrec_sort_field_ids": (C.c_int, [vp, i64, i32, vp, i32, vp, vp, vp, vp]),
    "mmlrec_emb_backward_update": (C.c_int, [vp, i64, i32, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
    "mmlrec_emb_stamp_rows": (C.c_int, [vp, vp, i32, i32, i32, vp, vp, vp]),
    "mmlrec_emb_adam_dense_sweep": (C.c_int, [vp, vp, vp, vp, i64, i32, vp, vp]),
    "mmlrec_gemm_grouped_f32": (C.c_int, [vp, vp, i32, i32, vp]),
    "mmlrec_tc_record_bytes": (i64, []),
    "mmlrec_tc_encode_problem": (C.c_int, [C.POINTER(GemmTcDesc), vp]),
    "mmlrec_gemm_grouped_tc": (C.c_int, [vp, vp, i32, i32, vp]),
    "mmlrec_gemm_grouped_tc_scheduled": (C.c_int, [vp, vp, i32, i32, vp, vp, i32, vp]),
    "mmlrec_tc_sm_count": (i32, []),
    "mmlrec_tc2_record_bytes": (C.c_int64, []),
    "mmlrec_tc2_num_tiles": (C.c_int32, [vp]),
    "mmlrec_tc2_encode_problem": (C.c_int, [vp, vp]),
    "mmlrec_gemm_grouped_tc2": (C.c_int, [vp, vp, i32, i32, vp, vp, i32, vp, vp]),
    "mmlrec_gemm_grouped_tc_debug": (C.c_int, [vp, vp, i32, i32, vp, vp, i32, vp, vp]),
    "mmlrec_tc_num_tiles": (i32, [i32, i32]),
    "mmlrec_bn_forward": (C.c_int, [vp, i64, i32, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, i64, vp, i64, i32, i32, vp]),
    "mmlrec_bn_backward": (C.c_int, [vp, i64, vp, i64, i32, i32, vp, vp, vp, vp, i64, vp, i64, vp, vp, vp]),
    "mmlrec_l2_regularize": (C.c_int, [vp, vp, vp, i64, vp, vp, vp]),
    "mmlrec_l2_scratch": (i64, []),
    "mmlrec_bn_stats": (C.c_int, [vp, i64, i32, i32, vp, vp]),
    "mmlrec_bn_combine": (C.c_int, [vp, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp]),
    "mmlrec_bn_backward_sums": (C.c_int, [vp, i64, vp, i64, i32, i32, vp, vp, vp, vp, vp, vp]),
    "mmlrec_bn_backward_synced": (C.c_int, [vp, i64, vp, i64, i32, i32, vp, vp, vp, vp, i64, vp, i64, vp, i32, vp]),
    "mmlrec_gate_mix_forward": (C.c_int, [vp, i32, i32, vp]),
    "mmlrec_gate_mix_backward": (C.c_int, [vp, i32, vp, i32, i32, i32, i32, i32, vp, vp, vp]),
    "mmlrec_gate_mix_backward_scratch": (i64, [i32, i32, i32, i32]),
    "mmlrec_gate_level_forward": (C.c_int, [vp, i32, vp]),
    "mmlrec_gate_level_backward": (C.c_int, [vp, i32, i32, i32, i32, vp, vp, vp]),
    "mmlrec_gate_level_backward_scratch": (i64, [i32, i32]),
    "mmlrec_peer_alloc": (C.c_int, [C.POINTER(C.c_void_p), i64]),
    "mmlrec_peer_free": (C.c_int, [vp]),
    "mmlrec_peer_export": (C.c_int, [vp, C.c_char_p]),
    "mmlrec_peer_import": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "mmlrec_peer_close": (C.c_int, [vp]),
    "mmlrec_peer_fill_u64": (C.c_int, [vp, i64, C.c_uint64, vp]),
    "mmlrec_gather_concat_sharded": (C.c_int, [vp, i64, i32, vp, i32, vp, i32, i32, vp, i32, i32, vp, i64, vp, i64, vp, vp]),
    "mmlrec_peer_barrier": (C.c_int, [vp, vp, i32, i32, vp, vp]),
    "mmlrec_peer_allreduce_f32": (C.c_int, [vp, vp, i64, i32, i32, vp]),
    "mmlrec_emb_push_ids": (C.c_int, [vp, i64, i32, vp, i32, i32, i32, i32, vp, vp, i32, vp, vp]),
    "mmlrec_emb_serve_rows": (C.c_int, [vp, vp, vp, i32, i32, i32, i32, vp, vp, i32, vp]),
    "mmlrec_gather_concat_staged": (C.c_int, [vp, i64, i32, vp, vp, i32, i32, vp, i32, i32, vp, i64, vp, i64, vp]),
    "mmlrec_emb_push_grads": (C.c_int, [vp, i64, i32, vp, i64, vp, i32, i32, i32, i32, i32, vp, vp]),
    "mmlrec_sort_field_keys": (C.c_int, [vp, i32, i32, vp, vp, vp, vp, vp]),
    "mmlrec_emb_backward_update_sharded": (C.c_int, [vp, i64, i32, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]),
    "mmlrec_sum_slices": (C.c_int, [vp, i32, i64, vp, vp, i32, i64, vp]),
    "mmlrec_gate_level_backward_tiled": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp]),
    "mmlrec_gate_level_backward_tiled_scratch": (i64, [i32, i32]),
    "mmlrec_gate_level_forward_tiled": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, vp]),
    "mmlrec_gate_level_forward_tiled_smem": (i64, [i32, i32, i32, i32, i32]),
    "mmlrec_gate_level_backward_tiled_smem": (i64, [i32, i32, i32, i32, i32, i32]),
    "mmlrec_heads_forward_backward": (C.c_int, [vp, i32, i32, vp, i64, vp, i64, vp, i32, i32, vp, i64, vp, vp]),
    "mmlrec_heads_forward_backward_masked": (C.c_int, [vp, i32, i32, vp, i64, vp, i64, vp, i64, vp, i32, i32, vp, i64, vp, vp]),
    "mmlrec_heads_backward_external": (C.c_int, [vp, i32, i32, vp, i64, vp, i64, vp, i32, vp, i64, vp, vp]),
    "mmlrec_heads_scratch": (i64, [i32, i32, i32]),
    "mmlrec_dense_optimizer_step": (C.c_int, [vp, vp, vp, vp, i64, vp, vp, vp]),
    "mmlrec_dense_optimizer_step_sliced": (C.c_int, [vp, vp, vp, vp, i64, vp, vp, i32, i64, vp]),
    "mmlrec_copy_cols": (C.c_int, [vp, i64, vp, i64, vp, i64, i32, i32, vp]),
    "mmlrec_mul_forward": (C.c_int, [vp, i64, vp, i64, vp, i64, vp, i64, i32, i32, vp]),
    "mmlrec_mul_backward": (C.c_int, [vp, i64, vp, i64, vp, i64, vp, vp, i64, i32, i32, vp, vp, i64, i32, i32, i32, i32, vp]),
    "mmlrec_aitm_attention_forward": (C.c_int, [vp, i64, i32, i32, vp, i64, vp, i64, vp, vp]),
    "mmlrec_aitm_attention_backward": (C.c_int, [vp, i64, vp, i64, vp, i32, i32, vp, vp, i64, vp]),
    "mmlrec_apg_mix_forward": (C.c_int, [vp, i64, vp, i64, vp, i64, i32, i32, vp, i64, vp, i64, vp]),
    "mmlrec_apg_mix_backward": (C.c_int, [vp, i64, vp, i64, vp, i64, i32, i32, vp, vp, i64, vp, vp, i64, vp, vp, i64, vp]),
    "mmlrec_colsum": (C.c_int, [vp, vp, i64, i32, i32, vp, vp]),
    "mmlrec_snr_gate_weights": (C.c_int, [vp, vp, vp, i32, i32, i32, i32, vp, i64, vp, vp]),
    "mmlrec_snr_gate_fold": (C.c_int, [vp, i64, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]),
    "mmlrec_star_weights": (C.c_int, [vp, vp, vp, vp, i32, i32, i32, vp, i64, vp, vp, vp]),
    "mmlrec_star_fold": (C.c_int, [vp, i64, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp]),
    "mmlrec_fill_f32": (C.c_int, [vp, i64, f32, vp]),
    "mmlrec_cast_f32_to_bf16": (C.c_int, [vp, i64, vp, i64, i32, i32, i32, vp]),
    "mmlrec_cast_bf16_to_f32": (C.c_int, [vp, i64, vp, i64, i32, i32, vp]),
    "mmlrec_mul_f32": (C.c_int, [vp, i64, vp, i64, vp, i64, i32, i32, i32, vp]),
}

_lib: Optional[C.CDLL] = None


class MmlrecLibraryError(RuntimeError):
    pass


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load the CUDA library (building it in-tree with nvcc if it is absent and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) and build_if_missing:
        try:
            from .csrc.build import build as _build
            _build(verbose=False)
        except Exception as e:  # noqa: BLE001
            raise MmlrecLibraryError(
                f"libmmlrec_b200.so is missing and could not be built ({e}); run __graft_entry__.build()") from e
    if not os.path.exists(LIB_PATH):
        raise MmlrecLibraryError(f"{LIB_PATH} not found: the CUDA library is required, there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.mmlrec_abi_version() != 1:
        raise MmlrecLibraryError("ABI version mismatch between lib.py and libmmlrec_b200.so")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().mmlrec_last_error().decode(errors="replace")
        raise MmlrecLibraryError(f"{what or 'mmlrec call'} failed (rc={rc}): {msg}")


def launch_count() -> int:
    return int(load().mmlrec_launch_count())


def make_hyper(optimizer: str, lr: float) -> Hyper:
    """Defaults of torch.optim.{Adam,Adagrad,SGD,RMSprop} as built by basemodel.py:569-584."""
    h = Hyper()
    h.step, h.optimizer, h.lr, h.lr_d = 0, OPT_CODES[optimizer], lr, lr
    h.beta1, h.beta2, h.beta1_d, h.beta2_d = 0.9, 0.999, 0.9, 0.999
    h.one_minus_beta1, h.one_minus_beta2 = 1 - 0.9, 1 - 0.999
    h.alpha, h.one_minus_alpha = 0.99, 1 - 0.99
    h.eps = {"adam": 1e-8, "adagrad": 1e-10, "rmsprop": 1e-8, "sgd": 0.0}[optimizer]
    h.step_size, h.bc2_sqrt = lr, 1.0
    return h


def struct_bytes(items) -> bytes:
    return b"".join(bytes(it) for it in items)
