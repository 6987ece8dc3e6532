"""Host-side check of the step PLANNER without a GPU (test infrastructure).

``engine/core.py`` turns a model into tables of kernel arguments: grouped-GEMM problem lists (pointers, strides, masks,
accumulate flags, gradient slices), head records, gather metadata.  Everything in those tables is decided on the host,
so it can be checked on the host: here the flat store and every activation buffer are allocated in CPU memory, the
C-ABI entry points the stages call are replaced by numpy restatements of their documented semantics
(``include/mmlrec_b200.h``) that read and write THROUGH THE POINTERS IN THE TABLES, and the planned forward + backward
program is run as written.  What this pins: buffer wiring, operand majors and strides, ReLU-mask sources, which
gradient is assigned / accumulated / sliced, parameter layout -- for both arithmetic modes (the bf16 mode's problem
descriptors are captured before they would be encoded into tensor maps).  What it cannot pin: the CUDA kernels
themselves (the ``-m gpu`` tests do that).

Only dense single-process plans without BatchNorm / gate stages are emulated (what the CPU tests need); anything else
raises ``NotImplementedError`` from ``FakeLib.__getattr__``.
"""
import ctypes as C

import numpy as np
import torch

from mmlrec_b200 import lib as L
from mmlrec_b200.engine import core
from mmlrec_b200.engine.store import FlatStore

_CT = {np.float32: C.c_float, np.uint16: C.c_uint16, np.int32: C.c_int32, np.int64: C.c_int64, np.uint8: C.c_uint8}


def _ptr(p):
    if p is None:
        return 0
    return int(p.value or 0) if isinstance(p, C.c_void_p) else int(p)


def view(ptr, dtype, count):
    """1-D numpy view of `count` elements at host address `ptr`."""
    ptr = _ptr(ptr)
    assert ptr, "null pointer dereferenced by the emulator"
    return np.ctypeslib.as_array((_CT[dtype] * int(count)).from_address(ptr))


def view2(ptr, dtype, rows, cols, rs, cs=1):
    """[rows, cols] strided view: element (r, c) at ptr[r * rs + c * cs]."""
    n = (rows - 1) * rs + (cols - 1) * cs + 1
    flat = view(ptr, dtype, n)
    return np.lib.stride_tricks.as_strided(flat, (rows, cols), (rs * flat.itemsize, cs * flat.itemsize))


def bf16_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16(x):
    """round-to-nearest-even, like cvt.rn.bf16.f32"""
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.bfloat16).view(torch.uint16).numpy()


def _act(v, code):
    if code == L.ACT_RELU:
        return np.maximum(v, 0.0)
    if code == L.ACT_SIGMOID:
        return 1.0 / (1.0 + np.exp(-v))
    if code == L.ACT_SIGMOID2:
        return 2.0 / (1.0 + np.exp(-v))
    return v


class FakeLib:
    """numpy restatements of the entry points a dense, single-process plan calls."""

    def __init__(self, real):
        self._real = real
        self.calls = []

    def __getattr__(self, name):
        if name in ("mmlrec_heads_scratch", "mmlrec_l2_scratch", "mmlrec_tc_sm_count", "mmlrec_last_error"):
            return getattr(self._real, name)
        raise NotImplementedError(f"plan emulator: {name} is not restated")

    # K1: include/mmlrec_b200.h mmlrec_gather_concat
    def mmlrec_gather_concat(self, X, ldx, B, emb, meta, F_s, D, dense_cols, F_d, dense_out_col, o32, ld32, o16, ld16, oob,
                             stream):
        self.calls.append("gather_concat")
        width = F_s * D + F_d
        out = np.zeros((B, width), np.float32)
        x = view2(X, np.float32, B, max(int(ldx), 1), ldx)
        if F_s:
            m = view(meta, np.int64, 4 * F_s).reshape(F_s, 4)
            for off, vocab, xc, oc in m:
                ids = x[:, xc].astype(np.int64)
                assert ids.min() >= 0 and ids.max() < vocab
                rows = view(_ptr(emb) + 4 * int(off), np.float32, int(vocab) * D).reshape(int(vocab), D)
                out[:, oc:oc + D] = rows[ids]
        if F_d:
            dc = view(dense_cols, np.int32, F_d)
            out[:, dense_out_col:dense_out_col + F_d] = x[:, dc]
        if _ptr(o32):
            view2(o32, np.float32, B, width, ld32)[:] = out
        if _ptr(o16):
            view2(o16, np.uint16, B, width, ld16)[:] = f32_to_bf16(out).reshape(B, width)
        return 0

    # K3 fp32: MmlrecGemmF32
    def mmlrec_gemm_grouped_f32(self, table, prefix, n, tiles, stream):
        self.calls.append(f"gemm_f32 x{n}")
        raw = bytes(view(table, np.uint8, n * C.sizeof(L.GemmF32)))
        probs = [L.GemmF32.from_buffer_copy(raw, i * C.sizeof(L.GemmF32)) for i in range(n)]
        results = []
        for p in probs:   # every problem reads the state BEFORE the launch (problems of one launch are independent)
            A = view2(p.A, np.float32, p.M, p.K, p.a_rs, p.a_cs).astype(np.float64)
            Bm = view2(p.B, np.float32, p.N, p.K, p.b_rs, p.b_cs).astype(np.float64)
            v = A @ Bm.T
            if p.bias:
                v = v + view(p.bias, np.float32, p.N).astype(np.float64)[None, :]
            v = _act(v, p.act)
            if p.mask:
                v = np.where(view2(p.mask, np.float32, p.M, p.N, p.ldmask) > 0, v, 0.0)
            results.append((p, v.astype(np.float32), A.sum(1).astype(np.float32)))
        seen = {}
        for p, v, rs in results:
            Cv = view2(p.C, np.float32, p.M, p.N, p.ldc)
            key = (_ptr(p.C), p.ldc)
            assert p.accumulate or key not in seen, "two problems of one launch assign the same output"
            seen[key] = True
            Cv[:] = (Cv + v) if p.accumulate else v
            if p.rowsum_a:
                view(p.rowsum_a, np.float32, p.M)[:] = rs
        return 0

    # K3 bf16: a list of MmlrecGemmTcDesc captured at plan time (see emulate.tc_table)
    def run_tc(self, descs):
        self.calls.append(f"gemm_tc x{len(descs)}")
        results = []
        for d in descs:
            A = bf16_to_f32(view2(d.A, np.uint16, d.M, d.K, *((1, d.lda) if d.a_mn_major else (d.lda, 1)))).astype(np.float64)
            Bm = bf16_to_f32(view2(d.B, np.uint16, d.N, d.K, *((1, d.ldb) if d.b_mn_major else (d.ldb, 1)))).astype(np.float64)
            v = A @ Bm.T
            if d.bias:
                v = v + view(d.bias, np.float32, d.N).astype(np.float64)[None, :]
            v = _act(v, d.act)
            if d.mask:
                keep = bf16_to_f32(view2(d.mask, np.uint16, d.M, d.N, d.ldmask)) > 0
                v = np.where(keep, v, 0.0)
            else:
                assert not d.mask_bits, "bit mask without its bf16 source"
            results.append((d, v.astype(np.float32), A.sum(1).astype(np.float32), Bm.sum(1).astype(np.float32)))
        seen = set()
        for d, v, a_sum, b_sum in results:
            if d.c_transposed:
                assert not (d.bias or d.mask or d.act or d.accumulate or d.colsum or d.C_bf16)
                key = _ptr(d.C_f32)
                assert key not in seen, "two problems of one launch assign the same output"
                seen.add(key)
                view2(d.C_f32, np.float32, d.N, d.M, d.ldc_f32)[:] = v.T
                if d.colsum_b:
                    assert _ptr(d.colsum_b) not in seen
                    seen.add(_ptr(d.colsum_b))
                    view(d.colsum_b, np.float32, d.N)[:] = b_sum
                continue
            if d.C_f32:
                Cv = view2(d.C_f32, np.float32, d.M, d.N, d.ldc_f32)
                assert d.accumulate or _ptr(d.C_f32) not in seen, "two problems of one launch assign the same output"
                seen.add(_ptr(d.C_f32))
                Cv[:] = (Cv + v) if d.accumulate else v
            if d.C_bf16:
                assert _ptr(d.C_bf16) not in seen
                seen.add(_ptr(d.C_bf16))
                view2(d.C_bf16, np.uint16, d.M, d.N, d.ldc_bf16)[:] = f32_to_bf16(v).reshape(d.M, d.N)
            if d.colsum:   # "rowsum of A" (the bias gradient of a wgrad problem), accumulated
                assert _ptr(d.colsum) not in seen
                seen.add(_ptr(d.colsum))
                view(d.colsum, np.float32, d.M)[:] += a_sum
        return 0

    # K4: MmlrecHead, mmlrec_heads_forward_backward (flags 0: independent heads)
    def mmlrec_heads_forward_backward(self, table, T, B, y, ldy, pred, ldp, loss, flags, training, scratch, n_scratch,
                                      counter, stream):
        self.calls.append("heads")
        if flags:
            raise NotImplementedError("plan emulator: plain heads only")
        raw = bytes(view(table, np.uint8, T * C.sizeof(L.Head)))
        heads = [L.Head.from_buffer_copy(raw, i * C.sizeof(L.Head)) for i in range(T)]
        P = view2(pred, np.float32, B, T, ldp)
        total = 0.0
        lossv = view(loss, np.float32, T + 1)
        wrote = set()
        dh_sum = {}   # heads that share an input: the kernel sums their contributions to d(h) (include/mmlrec_b200.h)
        for t, h in enumerate(heads):
            assert h.kind == L.HEAD_SIGMOID_BCE
            hv = view2(h.h, np.float32, B, h.H, h.ld_h).astype(np.float64)
            w = view(h.w, np.float32, h.H).astype(np.float64)
            z = hv @ w + (float(view(h.bias, np.float32, 1)[0]) if h.bias else 0.0)
            p = 1.0 / (1.0 + np.exp(-z))
            P[:, t] = p.astype(np.float32)
            if not training:
                continue
            yt = view2(y, np.float32, B, T, ldy)[:, t].astype(np.float64)
            p32 = P[:, t].astype(np.float64)
            lt = -(yt * np.maximum(np.log(p32), -100.0) + (1 - yt) * np.maximum(np.log1p(-p32), -100.0)).sum()
            lossv[t] = lt
            total += lt
            dz = p32 - yt
            for dst in (h.dw, h.dbias):
                assert not dst or _ptr(dst) not in wrote, "two heads assign the same gradient"
                wrote.add(_ptr(dst))
            view(h.dw, np.float32, h.H)[:] = (dz[:, None] * hv).sum(0).astype(np.float32)
            if h.dbias:
                view(h.dbias, np.float32, 1)[0] = np.float32(dz.sum())
            dh = dz[:, None] * w[None, :]
            if h.relu_mask:
                dh = np.where(hv > 0, dh, 0.0)
            key = (_ptr(h.d_h), _ptr(h.d_h_bf16))
            dh = dh_sum[key] = dh_sum.get(key, 0.0) + dh
            if h.d_h:
                view2(h.d_h, np.float32, B, h.H, h.ld_d_h)[:] = dh.astype(np.float32)
            if h.d_h_bf16:
                view2(h.d_h_bf16, np.uint16, B, h.H, h.ld_d_h_bf16)[:] = f32_to_bf16(dh).reshape(B, h.H)
        if training:
            lossv[T] = total
        return 0

    def mmlrec_copy_cols(self, src, ld_src, d32, ld32, d16, ld16, rows, cols, stream):
        self.calls.append("copy_cols")
        s = view2(src, np.float32, rows, cols, ld_src)
        if _ptr(d32):
            view2(d32, np.float32, rows, cols, ld32)[:] = s
        if _ptr(d16):
            view2(d16, np.uint16, rows, cols, ld16)[:] = f32_to_bf16(s).reshape(rows, cols)
        return 0

    def mmlrec_fill_f32(self, p, n, v, stream):
        view(p, np.float32, n)[:] = v
        return 0

    # aitm.py:82-91
    def mmlrec_aitm_attention_forward(self, vkq, ld, rows, H, o32, ld32, o16, ld16, attn, stream):
        self.calls.append("aitm_attention_forward")
        x = view2(vkq, np.float32, rows, 6 * H, ld).astype(np.float64)
        Vp, Kp, Qp, Vq, Kq, Qq = (x[:, i * H:(i + 1) * H] for i in range(6))
        s = np.stack([(Kp * Qp).sum(1), (Kq * Qq).sum(1)], 1) / np.sqrt(H)
        e = np.exp(s - s.max(1, keepdims=True))
        a = e / e.sum(1, keepdims=True)
        out = a[:, :1] * Vp + a[:, 1:] * Vq
        if _ptr(attn):
            view2(attn, np.float32, rows, 2, 2)[:] = a.astype(np.float32)
        if _ptr(o32):
            view2(o32, np.float32, rows, H, ld32)[:] = out.astype(np.float32)
        if _ptr(o16):
            view2(o16, np.uint16, rows, H, ld16)[:] = f32_to_bf16(out).reshape(rows, H)
        return 0

    def mmlrec_aitm_attention_backward(self, d_out, ld_dout, vkq, ld, attn, rows, H, d32, d16, ld_d, stream):
        self.calls.append("aitm_attention_backward")
        x = view2(vkq, np.float32, rows, 6 * H, ld).astype(np.float64)
        g = view2(d_out, np.float32, rows, H, ld_dout).astype(np.float64)
        a = view2(attn, np.float32, rows, 2, 2).astype(np.float64)
        Vp, Kp, Qp, Vq, Kq, Qq = (x[:, i * H:(i + 1) * H] for i in range(6))
        gp, gq = (g * Vp).sum(1), (g * Vq).sum(1)
        mean = a[:, 0] * gp + a[:, 1] * gq
        dsp = (a[:, 0] * (gp - mean) / np.sqrt(H))[:, None]
        dsq = (a[:, 1] * (gq - mean) / np.sqrt(H))[:, None]
        d = np.concatenate([a[:, :1] * g, dsp * Qp, dsp * Kp, a[:, 1:] * g, dsq * Qq, dsq * Kq], 1)
        if _ptr(d32):
            view2(d32, np.float32, rows, 6 * H, ld_d)[:] = d.astype(np.float32)
        if _ptr(d16):
            view2(d16, np.uint16, rows, 6 * H, ld_d)[:] = f32_to_bf16(d).reshape(rows, 6 * H)
        return 0


    # apg.py:96-99 per-sample product and the column-sum bias gradient (include/mmlrec_b200.h mmlrec_apg_mix_*, mmlrec_colsum)
    def mmlrec_apg_mix_forward(self, nk, ld_nk, wkk, ld_w, bkk, ld_b, B, k, o32, ld32, o16, ld16, stream):
        self.calls.append("apg_mix_forward")
        x = view2(nk, np.float32, B, k, ld_nk).astype(np.float64)
        w = view2(wkk, np.float32, B, k * k, ld_w).astype(np.float64).reshape(B, k, k)
        out = np.einsum("bi,bij->bj", x, w) + view2(bkk, np.float32, B, k, ld_b)
        if _ptr(o32):
            view2(o32, np.float32, B, k, ld32)[:] = out.astype(np.float32)
        if _ptr(o16):
            view2(o16, np.uint16, B, k, ld16)[:] = f32_to_bf16(out).reshape(B, k)
        return 0

    def mmlrec_apg_mix_backward(self, d_kk, ld_dkk, nk, ld_nk, wkk, ld_w, B, k, dnk32, dnk16, ld_dnk, dw32, dw16, ld_dw,
                                db32, db16, ld_db, stream):
        self.calls.append("apg_mix_backward")
        g = view2(d_kk, np.float32, B, k, ld_dkk).astype(np.float64)
        x = view2(nk, np.float32, B, k, ld_nk).astype(np.float64)
        w = view2(wkk, np.float32, B, k * k, ld_w).astype(np.float64).reshape(B, k, k)
        for val, p32, p16, ld, cols in ((np.einsum("bj,bij->bi", g, w), dnk32, dnk16, ld_dnk, k),
                                        (np.einsum("bi,bj->bij", x, g).reshape(B, k * k), dw32, dw16, ld_dw, k * k),
                                        (g, db32, db16, ld_db, k)):
            assert bool(_ptr(p32)) != bool(_ptr(p16)), "exactly one destination precision"
            if _ptr(p32):
                view2(p32, np.float32, B, cols, ld)[:] = val.astype(np.float32)
            else:
                view2(p16, np.uint16, B, cols, ld)[:] = f32_to_bf16(val).reshape(B, cols)
        return 0

    def mmlrec_colsum(self, z32, z16, ld, B, N, out, stream):
        self.calls.append("colsum")
        assert bool(_ptr(z32)) != bool(_ptr(z16))
        z = view2(z32, np.float32, B, N, ld) if _ptr(z32) else bf16_to_f32(view2(z16, np.uint16, B, N, ld))
        view(out, np.float32, N)[:] = z.astype(np.float64).sum(0).astype(np.float32)
        return 0

    # snr_trans.py:36-50 as a derived weight (include/mmlrec_b200.h mmlrec_snr_gate_weights / _fold)
    @staticmethod
    def _snr_s(u, alpha):
        logit = np.log(u) - np.log(1.0 - u) + np.log(alpha) / 0.9
        return 1.0 / (1.0 + np.exp(-logit))

    def mmlrec_snr_gate_weights(self, u, alpha, trans, n_out, n_in, U, zdim, w_eff, ld_w, w16, stream):
        self.calls.append("snr_gate_weights")
        assert zdim in (1, U)
        uu = view(u, np.float32, n_out * n_in * zdim).reshape(n_out, n_in, zdim).astype(np.float64)
        a = float(view(alpha, np.float32, 1)[0])
        M = view(trans, np.float32, n_out * n_in * U * U).reshape(n_out, n_in, U, U).astype(np.float64)
        z = np.clip(self._snr_s(uu, a) * 1.2 - 0.1, 0.0, 1.0)            # [n_out, n_in, zdim] over the output unit v
        # w[i*U+v, j*U+u] = z_ij[v] * M[i, j, u, v]
        w = (z[:, :, None, :] * M).transpose(0, 3, 1, 2).reshape(n_out * U, n_in * U)
        view2(w_eff, np.float32, n_out * U, n_in * U, ld_w)[:] = w.astype(np.float32)
        if _ptr(w16):
            view2(w16, np.uint16, n_out * U, n_in * U, ld_w)[:] = f32_to_bf16(w).reshape(n_out * U, n_in * U)
        return 0

    def mmlrec_snr_gate_fold(self, d_w_eff, ld_w, trans, u, alpha, n_out, n_in, U, zdim, dz_scratch, d_u, d_alpha, stream):
        self.calls.append("snr_gate_fold")
        dW = view2(d_w_eff, np.float32, n_out * U, n_in * U, ld_w).astype(np.float64)
        uu = view(u, np.float32, n_out * n_in * zdim).reshape(n_out, n_in, zdim).astype(np.float64)
        a = float(view(alpha, np.float32, 1)[0])
        M = view(trans, np.float32, n_out * n_in * U * U).reshape(n_out, n_in, U, U).astype(np.float64)
        t = dW.reshape(n_out, U, n_in, U).transpose(0, 2, 3, 1) * M      # [i, j, u, v]
        dz = t.sum((2, 3))[:, :, None] if zdim == 1 else t.sum(2)
        s = self._snr_s(uu, a)
        s_ = s * 1.2 - 0.1
        dlogit = np.where((s_ > 0) & (s_ <= 1), dz * 1.2 * s * (1 - s), 0.0)
        if _ptr(d_u):
            view(d_u, np.float32, n_out * n_in * zdim)[:] = (dlogit * (1 / uu + 1 / (1 - uu))).reshape(-1).astype(np.float32)
        view(d_alpha, np.float32, 1)[0] = np.float32(dlogit.sum() / (a * 0.9))
        return 0


class EmulatedPlan:
    """The planned step of `model` (constructed on the CPU) at batch B, every buffer in host memory."""

    def __init__(self, model, B, precision):
        cpu = torch.device("cpu")
        self.model = model
        dry = core.Builder(2, cpu, None, dry=True, precision=precision)
        model.build_graph(dry)
        model.store = FlatStore(model, dry.param_order, [t[0] for t in model.embedding_layout], cpu,
                                want_bf16=precision == "bf16", ordered_buffers=dry.buffer_order,
                                aux_floats=dry.aux_floats + 64, emb_alloc=None)
        model._index_features()
        self.precision = precision
        self.B = B

    def build(self):
        """after the initial state has been loaded into the store"""
        model, cpu = self.model, torch.device("cpu")
        model.store.refresh_bf16()
        b = core.Builder(self.B, cpu, model.store, dry=False, precision=self.precision)
        fake = FakeLib(b.lib)
        b.lib = fake
        launches = []
        b.tc_table = lambda descs: ("captured", launches.append([_copy_desc(d) for d in descs]) or len(launches) - 1)
        b.tc_launch = lambda tbl, stream, stamps=None: fake.run_tc(launches[tbl[1]])
        b.dp, b.mask_domains = None, 0
        model.build_graph(b)
        b.materialize()
        self.b, self.fake, self.stages = b, fake, b.stages
        self.gather = next(s for s in b.stages if isinstance(s, core.GatherStage))
        self.heads = next(s for s in b.stages if isinstance(s, core.HeadStage))
        for s in reversed(b.stages):
            s.plan_backward()
        self.grad_slices = max([getattr(s, "split_k", 1) for s in b.stages] + [1])
        return self

    def forward_backward(self, X, y):
        self.gather.X.copy_(torch.as_tensor(X, dtype=torch.float32))
        self.heads.y.copy_(torch.as_tensor(y, dtype=torch.float32))
        self.model.store.grad_slices.zero_()
        for s in self.stages:
            s.forward(0, True)
        for s in reversed(self.stages):
            if s is not self.gather:
                s.backward(0)
        self.model.store.live_slices = self.grad_slices
        return self.heads.pred.clone(), self.heads.loss.clone()

    def grad(self, prm):
        return self.model.store.grad_view(prm).clone()

    def input_grad(self):
        return self.gather.out.grad_tensor().clone()


def _check_tc_desc(d):
    """The argument checks of mmlrec_tc2_encode_problem (csrc/gemm_tc2.cu): what the tensor maps need.  Host buffers of
    the emulation are 64-byte aligned like the device allocations, so column offsets show up the same way."""
    assert d.M > 0 and d.N > 0 and d.K > 0, "bad sizes"
    assert _ptr(d.A) % 16 == 0 and _ptr(d.B) % 16 == 0, "operands must be 16-byte aligned"
    assert d.lda % 8 == 0 and d.ldb % 8 == 0, "operand row strides must be multiples of 8 elements"
    assert not d.C_f32 or d.c_transposed or (d.ldc_f32 % 4 == 0 and _ptr(d.C_f32) % 16 == 0), "C_f32 alignment"
    assert not d.C_bf16 or (d.ldc_bf16 % 8 == 0 and _ptr(d.C_bf16) % 16 == 0), "C_bf16 alignment"
    assert not d.mask or (d.ldmask % 8 == 0 and _ptr(d.mask) % 16 == 0), "mask alignment"
    assert not d.colsum_b or d.c_transposed, "colsum_b comes with c_transposed"


def _copy_desc(d):
    _check_tc_desc(d)
    c = L.GemmTcDesc()
    C.memmove(C.byref(c), C.byref(d), C.sizeof(L.GemmTcDesc))
    return c
