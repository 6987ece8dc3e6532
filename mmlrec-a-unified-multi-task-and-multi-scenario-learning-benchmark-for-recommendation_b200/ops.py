"""Tensor-level wrappers over the C ABI: one Python function per kernel family, taking CUDA
``torch.Tensor``s and launching on the current stream.  The step program (``engine/core.py``) talks
to the library directly with pre-built tables; these wrappers are the operator-level API for users
and for the kernel parity tests.  No wrapper has a non-CUDA path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import lib as L


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dev_table(structs, device) -> torch.Tensor:
    return torch.frombuffer(bytearray(L.struct_bytes(structs)), dtype=torch.uint8).to(device)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("mmlrec_b200 ops run only on CUDA tensors (no CPU fallback)")


def field_meta(table_offsets: Sequence[int], vocabs: Sequence[int], x_cols: Sequence[int], D: int, device) -> torch.Tensor:
    meta = []
    for j, (o, v, c) in enumerate(zip(table_offsets, vocabs, x_cols)):
        meta += [o, v, c, j * D]
    return torch.tensor(meta, dtype=torch.int64, device=device)


def gather_concat(X: torch.Tensor, emb: torch.Tensor, meta: torch.Tensor, D: int, dense_cols: Sequence[int] = (),
                  want_f32: bool = True, want_bf16: bool = False, ld_bf16: Optional[int] = None
                  ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], torch.Tensor]:
    """K1.  Returns (dnn_input fp32 [B,in] or None, bf16 copy [B,ld_bf16] or None, oob flag)."""
    _need_cuda(X, emb, meta)
    lib = L.load()
    B, F_s, F_d = X.shape[0], meta.numel() // 4, len(dense_cols)
    in_dim = F_s * D + F_d
    dc = torch.tensor(list(dense_cols) or [0], dtype=torch.int32, device=X.device)
    out = torch.empty(B, in_dim, dtype=torch.float32, device=X.device) if want_f32 else None
    out16 = None
    if want_bf16:
        ld_bf16 = ld_bf16 or (in_dim + 63) // 64 * 64
        out16 = torch.empty(B, ld_bf16, dtype=torch.bfloat16, device=X.device)
    oob = torch.zeros(1, dtype=torch.int32, device=X.device)
    L.check(lib.mmlrec_gather_concat(X.data_ptr(), X.stride(0), B, emb.data_ptr(), meta.data_ptr(), F_s, D,
                                     dc.data_ptr(), F_d, F_s * D, out.data_ptr() if out is not None else None,
                                     out.stride(0) if out is not None else 0,
                                     out16.data_ptr() if out16 is not None else None,
                                     out16.stride(0) if out16 is not None else 0, oob.data_ptr(), _stream()),
            "gather_concat")
    return out, out16, oob


def sort_field_ids(X: torch.Tensor, meta: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    _need_cuda(X, meta)
    lib = L.load()
    B, F_s = X.shape[0], meta.numel() // 4
    n_pad = 32
    while n_pad < B:
        n_pad <<= 1
    ids = torch.empty(F_s, B, dtype=torch.int32, device=X.device)
    pos = torch.empty(F_s, B, dtype=torch.int32, device=X.device)
    ws = torch.empty(F_s * n_pad, dtype=torch.int64, device=X.device)
    L.check(lib.mmlrec_sort_field_ids(X.data_ptr(), X.stride(0), B, meta.data_ptr(), F_s, ids.data_ptr(),
                                      pos.data_ptr(), ws.data_ptr(), _stream()), "sort_field_ids")
    return ids, pos


def hyper_tensor(optimizer: str, lr: float, device) -> torch.Tensor:
    return torch.frombuffer(bytearray(bytes(L.make_hyper(optimizer, lr))), dtype=torch.uint8).to(device)


def hyper_advance(hyper: torch.Tensor) -> None:
    L.check(L.load().mmlrec_hyper_advance(hyper.data_ptr(), _stream()), "hyper_advance")


def emb_backward_update(d_input: torch.Tensor, ids: torch.Tensor, pos: torch.Tensor, meta: torch.Tensor, D: int,
                        hyper: torch.Tensor, emb: Optional[torch.Tensor] = None, s1: Optional[torch.Tensor] = None,
                        s2: Optional[torch.Tensor] = None, row_touch: Optional[torch.Tensor] = None,
                        grad_rows_out: Optional[torch.Tensor] = None) -> None:
    """K2 (segmented reduce + fused row update).  ``grad_rows_out`` [F_s,B,D] receives, at the last
    sorted position of every run, the summed gradient row (test hook)."""
    _need_cuda(d_input, ids, pos, meta, hyper)
    p = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
    F_s, B = ids.shape
    L.check(L.load().mmlrec_emb_backward_update(d_input.data_ptr(), d_input.stride(0), B, ids.data_ptr(), pos.data_ptr(),
                                                meta.data_ptr(), F_s, D, p(emb), p(s1), p(s2), p(row_touch),
                                                hyper.data_ptr(), p(grad_rows_out), _stream()), "emb_backward_update")


def emb_adam_dense_sweep(emb, m, v, row_touch, D: int, hyper) -> None:
    L.check(L.load().mmlrec_emb_adam_dense_sweep(emb.data_ptr(), m.data_ptr(), v.data_ptr(), row_touch.data_ptr(),
                                                 emb.numel() // D, D, hyper.data_ptr(), _stream()), "adam sweep")


def dense_optimizer_step(param, grad, s1, s2, hyper, bf16_shadow=None) -> None:
    p = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
    L.check(L.load().mmlrec_dense_optimizer_step(param.data_ptr(), grad.data_ptr(), p(s1), p(s2), param.numel(),
                                                 hyper.data_ptr(), p(bf16_shadow), _stream()), "dense optimizer")


# ------------------------------------------------------------------------------------------------
# grouped GEMMs
# ------------------------------------------------------------------------------------------------
def gemm_grouped_f32(problems: List[Dict]) -> None:
    """Each problem: dict(A, B, C tensors 2-D fp32 with arbitrary strides, computing C = act(A @ B^T + bias)),
    optional bias, mask, rowsum_a, act (name), accumulate."""
    lib = L.load()
    recs = []
    for q in problems:
        A, Bm, Cm = q["A"], q["B"], q["C"]
        _need_cuda(A, Bm, Cm)
        r = L.GemmF32()
        r.A, r.a_rs, r.a_cs = A.data_ptr(), A.stride(0), A.stride(1)
        r.B, r.b_rs, r.b_cs = Bm.data_ptr(), Bm.stride(0), Bm.stride(1)
        assert Cm.stride(1) == 1
        r.C, r.ldc = Cm.data_ptr(), Cm.stride(0)
        r.M, r.K = A.shape
        r.N = Bm.shape[0]
        assert Bm.shape[1] == r.K and tuple(Cm.shape) == (r.M, r.N)
        if q.get("bias") is not None:
            r.bias = q["bias"].data_ptr()
        if q.get("mask") is not None:
            r.mask, r.ldmask = q["mask"].data_ptr(), q["mask"].stride(0)
        if q.get("rowsum_a") is not None:
            r.rowsum_a = q["rowsum_a"].data_ptr()
        r.act = L.ACT_CODES[q.get("act")]
        r.accumulate = 1 if q.get("accumulate") else 0
        recs.append(r)
    dev = problems[0]["A"].device
    pre, at = [0], 0
    for r in recs:
        at += ((r.M + 63) // 64) * ((r.N + 63) // 64)
        pre.append(at)
    table = _dev_table(recs, dev)
    prefix = torch.tensor(pre, dtype=torch.int32, device=dev)
    L.check(lib.mmlrec_gemm_grouped_f32(table.data_ptr(), prefix.data_ptr(), len(recs), at, _stream()), "gemm_grouped_f32")
    torch.cuda.current_stream().synchronize()  # the temporaries above must outlive the launch


class TcProblemTable:
    """Device-resident table of tensor-core GEMM problems (tensor maps encoded on the host once).
    ``kernel``: 1 = one CTA per 128x128 tile (gemm_tc.cu), 2 = CTA pairs on 256-wide tiles (gemm_tc2.cu);
    default from ``MMLREC_TC_KERNEL`` (2)."""

    def __init__(self, descs: Sequence[L.GemmTcDesc], device, kernel: Optional[int] = None):
        import os
        lib = L.load()
        self.kernel = int(os.environ.get("MMLREC_TC_KERNEL", "2")) if kernel is None else kernel
        pairs = self.kernel == 2
        if pairs:
            descs = L.split_two_output_problems(descs)
        rb = int(lib.mmlrec_tc2_record_bytes() if pairs else lib.mmlrec_tc_record_bytes())
        host = (C.c_uint8 * (rb * len(descs)))()
        pre, at = [0], 0
        for i, d in enumerate(descs):
            if pairs:
                L.check(lib.mmlrec_tc2_encode_problem(C.byref(d), C.addressof(host) + i * rb), "tc2_encode_problem")
                at += int(lib.mmlrec_tc2_num_tiles(C.byref(d)))
            else:
                L.check(lib.mmlrec_tc_encode_problem(C.byref(d), C.addressof(host) + i * rb), "tc_encode_problem")
                at += int(lib.mmlrec_tc_num_tiles(d.M, d.N))
            pre.append(at)
        # 128-byte aligned device copy
        raw = torch.frombuffer(bytearray(bytes(host)), dtype=torch.uint8)
        buf = torch.empty(rb * len(descs) + 128, dtype=torch.uint8, device=device)
        shift = (-buf.data_ptr()) % 128
        self.records = buf[shift:shift + rb * len(descs)]
        self.records.copy_(raw)
        self._buf = buf
        self.prefix = torch.tensor(pre, dtype=torch.int32, device=device)
        self.n, self.tiles = len(descs), at

    def launch(self, stream: Optional[int] = None) -> None:
        st = _stream() if stream is None else stream
        if self.kernel == 2:
            L.check(L.load().mmlrec_gemm_grouped_tc2(self.records.data_ptr(), self.prefix.data_ptr(), self.n, self.tiles,
                                                     None, None, 0, None, st), "gemm_grouped_tc2")
        else:
            L.check(L.load().mmlrec_gemm_grouped_tc(self.records.data_ptr(), self.prefix.data_ptr(), self.n, self.tiles,
                                                    st), "gemm_grouped_tc")


def tc_desc(A: torch.Tensor, B: torch.Tensor, M: int, N: int, K: int, a_mn: bool = False, b_mn: bool = False,
            C_f32: Optional[torch.Tensor] = None, C_bf16: Optional[torch.Tensor] = None,
            bias: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
            rowsum_a: Optional[torch.Tensor] = None, act: Optional[str] = None, accumulate: bool = False,
            relu_bits_out: Optional[torch.Tensor] = None, mask_bits: Optional[torch.Tensor] = None,
            bits_chunks: int = 0, bits_chunk0: int = 0, c_transposed: bool = False,
            colsum_b: Optional[torch.Tensor] = None) -> L.GemmTcDesc:
    """D[M,N] = act(A B^T + bias).  ``A`` is a bf16 array [M,K] (or [K,M] when ``a_mn``); ``B`` is [N,K]
    (or [K,N] when ``b_mn``); both row-major with stride(1) == 1."""
    _need_cuda(A, B)
    d = L.GemmTcDesc()
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.stride(1) == 1 and B.stride(1) == 1
    d.A, d.B, d.lda, d.ldb = A.data_ptr(), B.data_ptr(), A.stride(0), B.stride(0)
    d.a_mn_major, d.b_mn_major = int(a_mn), int(b_mn)
    d.M, d.N, d.K = M, N, K
    if C_f32 is not None:
        d.C_f32, d.ldc_f32 = C_f32.data_ptr(), C_f32.stride(0)
    if C_bf16 is not None:
        d.C_bf16, d.ldc_bf16 = C_bf16.data_ptr(), C_bf16.stride(0)
    if bias is not None:
        d.bias = bias.data_ptr()
    if mask is not None:
        d.mask, d.ldmask = mask.data_ptr(), mask.stride(0)
    if rowsum_a is not None:
        d.colsum = rowsum_a.data_ptr()
    d.act, d.accumulate = L.ACT_CODES[act], int(accumulate)
    d.c_transposed = int(c_transposed)   # C_f32 is then [N, M]: C[n, m] = D[m, n] (pair kernel)
    if colsum_b is not None:             # [N] fp32: sums of B over K (N <= 128, with c_transposed)
        d.colsum_b = colsum_b.data_ptr()
    if relu_bits_out is not None:   # int32 [ceil(M/32), bits_chunks, 32] (pair kernel only)
        d.relu_bits_out, d.bits_out_chunks, d.bits_out_chunk0 = relu_bits_out.data_ptr(), bits_chunks, bits_chunk0
    if mask_bits is not None:
        d.mask_bits, d.mask_bits_chunks, d.mask_bits_chunk0 = mask_bits.data_ptr(), bits_chunks, bits_chunk0
    return d
