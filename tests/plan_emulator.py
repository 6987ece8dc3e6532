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

Single-process plans of every model family are emulated (GEMMs, BatchNorm, gate levels, heads with all their flags,
the element-wise and derived-weight stages); the table update (K2), the optimizer, the L2 kernel and the multi-GPU
exchanges are not: anything else raises ``NotImplementedError`` from ``FakeLib.__getattr__``.
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
        if name in ("mmlrec_l2_scratch", "mmlrec_tc_sm_count", "mmlrec_last_error") or name.endswith("_scratch") \
                or name.endswith("_smem"):
            return getattr(self._real, name)   # host-side size queries: no device work
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

    # K4: MmlrecHead (include/mmlrec_b200.h): flags bit 0 esmm, bit 1 cumulative biases, bit 2 one shared bias
    def _heads(self, table, T, B, y, ldy, mask, ldm, pred, ldp, loss, flags, training, external):
        self.calls.append("heads")
        raw = bytes(view(table, np.uint8, T * C.sizeof(L.Head)))
        heads = [L.Head.from_buffer_copy(raw, i * C.sizeof(L.Head)) for i in range(T)]
        esmm, cum, shared = bool(flags & 1), bool(flags & 2), bool(flags & 4)
        if training or external:
            # the dispatch of heads_launch (csrc/fused_ops.cu): cumulative / shared biases and heads that share an input or
            # a final layer exist only in the one-launch kernel, which is instantiated for these shapes
            hmax = max(h.H for h in heads)
            one_launch = (T <= 4 and hmax <= 64) or (T <= 8 and hmax <= 128) or (bool(flags & 6) and T <= 4 and hmax <= 256)
            sharing = len({_ptr(h.h) for h in heads}) < T or len({_ptr(h.dw) for h in heads}) < T
            if ((flags & 6) or sharing) and not one_launch:
                return -1
        P = view2(pred, np.float32, B, T, ldp)
        lossv = view(loss, np.float32, T + 1)
        scal = lambda p: float(view(p, np.float32, 1)[0]) if p else 0.0   # noqa: E731
        sig32 = lambda v: (1.0 / (1.0 + np.exp(-v.astype(np.float32)))).astype(np.float32).astype(np.float64)   # noqa: E731
        hv = [view2(h.h, np.float32, B, h.H, h.ld_h).astype(np.float64) for h in heads]
        wv = [view(h.w, np.float32, h.H).astype(np.float64) for h in heads]
        z = []
        for t, h in enumerate(heads):
            zt = hv[t] @ wv[t] + scal(h.bias) + scal(h.bias2)
            if cum:
                zt = zt + sum(scal(heads[q].bias) for q in range(t))
            z.append(zt)
        Y = view2(y, np.float32, B, T, ldy).astype(np.float64) if (training or external) else None
        M = view2(mask, np.float32, B, int(ldm), ldm).astype(np.float64) if mask else None
        dz, lt = [None] * T, [0.0] * T
        cross = None
        for t, h in enumerate(heads):
            if h.kind == L.HEAD_SIGMOID_BCE:
                # the probability is an fp32 number in the kernel: a saturated logit gives exactly 1.0f, hence a zero
                # (1 - p) p factor and the reference's zero gradient behind the -100 clamp
                p = sig32(z[t])
                out, scale = p, np.ones(B)
                if esmm and t == 1:
                    p0 = sig32(z[0])
                    out, scale = p0 * p, p0
                if M is not None:
                    mk = M[:, h.mask_col]
                    out, scale = out * mk, scale * mk
                else:
                    mk = 1.0
                P[:, t] = out.astype(np.float32)
                if Y is None:
                    continue
                o32 = P[:, t].astype(np.float64)
                if external:
                    gout = Y[:, t]
                else:
                    with np.errstate(divide="ignore"):
                        lt[t] = float((mk * ((Y[:, t] - 1) * np.maximum(np.log1p(-o32), -100.0)
                                             - Y[:, t] * np.maximum(np.log(o32), -100.0))).sum())
                    gout = mk * (o32 - Y[:, t]) / np.maximum((1 - o32) * o32, 1e-12)
                dz[t] = gout * scale * (1 - p) * p
                if esmm and t == 1:
                    cross = gout * p
            else:
                P[:, t] = z[t].astype(np.float32)
                if Y is None:
                    continue
                if external:
                    dz[t] = Y[:, t]
                else:
                    d = z[t] - Y[:, t]
                    lt[t] = float((d * d).sum())
                    dz[t] = 2 * d
        if Y is None:
            return 0
        if esmm:
            p0 = sig32(z[0])
            dz[0] = dz[0] + cross * (1 - p0) * p0
        for t in range(T):
            lossv[t] = 0.0 if external else lt[t]
        lossv[T] = 0.0 if external else sum(lt)
        # dw: heads that share one final layer share its gradient (sum in task order); d(h): heads that read the same
        # activation share its gradient buffer (the kernel writes the sum once)
        dw_sum, dh_sum = {}, {}
        for t, h in enumerate(heads):
            dw_sum[_ptr(h.dw)] = dw_sum.get(_ptr(h.dw), 0.0) + (dz[t][:, None] * hv[t]).sum(0)
            key = (_ptr(h.d_h), _ptr(h.d_h_bf16))
            if key != (0, 0):
                dh_sum[key] = dh_sum.get(key, 0.0) + dz[t][:, None] * wv[t][None, :]
        for t, h in enumerate(heads):
            view(h.dw, np.float32, h.H)[:] = dw_sum[_ptr(h.dw)].astype(np.float32)
            key = (_ptr(h.d_h), _ptr(h.d_h_bf16))
            if key != (0, 0):
                dh = np.where(hv[t] > 0, dh_sum[key], 0.0) if h.relu_mask else dh_sum[key]
                if h.d_h:
                    view2(h.d_h, np.float32, B, h.H, h.ld_d_h)[:] = dh.astype(np.float32)
                if h.d_h_bf16:
                    view2(h.d_h_bf16, np.uint16, B, h.H, h.ld_d_h_bf16)[:] = f32_to_bf16(dh).reshape(B, h.H)
        tot = [float(d.sum()) for d in dz]
        if esmm or shared:
            if heads[0].dbias:
                view(heads[0].dbias, np.float32, 1)[0] = np.float32(sum(tot))
        elif cum:
            for q, h in enumerate(heads):
                if h.dbias:
                    view(h.dbias, np.float32, 1)[0] = np.float32(sum(tot[q:]))
        else:
            for q, h in enumerate(heads):
                if h.dbias:
                    view(h.dbias, np.float32, 1)[0] = np.float32(tot[q])
        for q, h in enumerate(heads):
            if h.dbias2:
                view(h.dbias2, np.float32, 1)[0] = np.float32(tot[q])
        return 0

    def mmlrec_heads_forward_backward(self, table, T, B, y, ldy, pred, ldp, loss, flags, training, scratch, n_scratch,
                                      counter, stream):
        return self._heads(table, T, B, y, ldy, None, 0, pred, ldp, loss, flags, training, False)

    def mmlrec_heads_forward_backward_masked(self, table, T, B, y, ldy, mask, ldm, pred, ldp, loss, flags, training,
                                             scratch, n_scratch, counter, stream):
        return self._heads(table, T, B, y, ldy, mask, ldm, pred, ldp, loss, flags, training, False)

    def mmlrec_heads_backward_external(self, table, T, B, d_pred, ld_d_pred, pred, ldp, loss, flags, scratch, n_scratch,
                                       counter, stream):
        return self._heads(table, T, B, d_pred, ld_d_pred, None, 0, pred, ldp, loss, flags, True, True)

    # gate head + softmax + mixture of a level (MmlrecGateLevel; the tiled and the plain kernels share these semantics)
    @staticmethod
    def _level(level):
        return L.GateLevel.from_buffer_copy(bytes(view(level, np.uint8, C.sizeof(L.GateLevel))))

    def _gate_level_forward(self, level, B):
        self.calls.append("gate_level_forward")
        r = self._level(level)
        ex = [view2(r.expert[u], np.float32, B, r.H, r.ld_expert).astype(np.float64) for u in range(r.n_experts)]
        for g in range(r.n_gates):
            gi = view2(r.gate_in[g], np.float32, B, r.Hg[g], r.ld_gate_in[g]).astype(np.float64)
            Wg = view2(r.Wg[g], np.float32, r.n_e[g], r.Hg[g], r.ld_Wg[g]).astype(np.float64)
            logit = gi @ Wg.T
            e = np.exp(logit - logit.max(1, keepdims=True))
            p = e / e.sum(1, keepdims=True)
            view2(r.probs[g], np.float32, B, r.n_e[g], r.n_e[g])[:] = p.astype(np.float32)
            mix = np.zeros((B, r.H))
            for u in range(r.n_experts):
                if r.slot[u][g] >= 0:
                    mix += p[:, r.slot[u][g]][:, None] * ex[u]
            view2(r.mix[g], np.float32, B, r.H, r.ld_mix[g])[:] = mix.astype(np.float32)
            if r.mix_bf16[g]:
                view2(r.mix_bf16[g], np.uint16, B, r.H, r.ld_mix_bf16[g])[:] = f32_to_bf16(mix).reshape(B, r.H)
        return 0

    def mmlrec_gate_level_forward(self, level, B, stream):
        return self._gate_level_forward(level, B)

    def mmlrec_gate_level_forward_tiled(self, level, B, n_experts, H, total_wg, total_ne, total_hg, stream):
        # the launcher's limits (csrc/gate_level.cu)
        if not (0 < n_experts <= L.LEVEL_MAX_EXPERTS and H % 4 == 0 and total_wg % 4 == 0 and total_hg % 4 == 0
                and 0 < total_ne <= L.LEVEL_MAX_GATES * L.LEVEL_MAX_EXPERTS
                and self._real.mmlrec_gate_level_forward_tiled_smem(n_experts, H, total_wg, total_ne, total_hg) <= 110 * 1024):
            return -1
        return self._gate_level_forward(level, B)

    def _gate_level_backward(self, level, B, tiled):
        self.calls.append("gate_level_backward")
        r = self._level(level)
        ex = [view2(r.expert[u], np.float32, B, r.H, r.ld_expert).astype(np.float64) for u in range(r.n_experts)]
        d_ex = [np.zeros((B, r.H)) for _ in range(r.n_experts)]
        for g in range(r.n_gates):
            if not r.d_mix[g]:
                continue
            dm = view2(r.d_mix[g], np.float32, B, r.H, r.ld_d_mix[g]).astype(np.float64)
            p = view2(r.probs[g], np.float32, B, r.n_e[g], r.n_e[g]).astype(np.float64)
            dp = np.zeros_like(p)
            for u in range(r.n_experts):
                s_ = r.slot[u][g]
                if s_ < 0:
                    continue
                dp[:, s_] = (dm * ex[u]).sum(1)
                detached = bool(r.detach_mask[u] >> g & 1)
                assert tiled or not detached, "detached pairs need the tiled backward"
                if not detached:
                    d_ex[u] += p[:, s_][:, None] * dm
            dlogit = p * (dp - (p * dp).sum(1, keepdims=True))
            gi = view2(r.gate_in[g], np.float32, B, r.Hg[g], r.ld_gate_in[g]).astype(np.float64)
            Wg = view2(r.Wg[g], np.float32, r.n_e[g], r.Hg[g], r.ld_Wg[g]).astype(np.float64)
            view2(r.dWg[g], np.float32, r.n_e[g], r.Hg[g], r.ld_Wg[g])[:] = (dlogit.T @ gi).astype(np.float32)
            dgi = dlogit @ Wg
            if r.relu_mask_gate_in[g]:
                dgi = np.where(gi > 0, dgi, 0.0)
            if r.d_gate_in[g]:
                dst = view2(r.d_gate_in[g], np.float32, B, r.Hg[g], r.ld_d_gate_in[g])
                dst[:] = (dst + dgi).astype(np.float32) if r.accumulate_d_gate_in[g] else dgi.astype(np.float32)
            if r.d_gate_in_bf16[g]:
                view2(r.d_gate_in_bf16[g], np.uint16, B, r.Hg[g], r.ld_d_gate_in_bf16[g])[:] = \
                    f32_to_bf16(dgi).reshape(B, r.Hg[g])
        for u in range(r.n_experts):
            if not (r.d_expert[u] or r.d_expert_bf16[u]):
                continue
            v = np.where(ex[u] > 0, d_ex[u], 0.0) if r.expert_relu else d_ex[u]
            if r.d_expert[u]:
                view2(r.d_expert[u], np.float32, B, r.H, r.ld_d_expert)[:] = v.astype(np.float32)
            if r.d_expert_bf16[u]:
                view2(r.d_expert_bf16[u], np.uint16, B, r.H, r.ld_d_expert_bf16)[:] = f32_to_bf16(v).reshape(B, r.H)
        return 0

    def mmlrec_gate_level_backward(self, level, B, total_wg, total_ne, total_hg, scratch, counter, stream):
        # the launcher's limits: staged gate-head weights and 8 rows of (dlogits + gate inputs) in shared memory
        if not (0 < total_wg <= L.LEVEL_MAX_WG and 8 * (total_ne + total_hg) * 4 <= 160 * 1024):
            return -1
        return self._gate_level_backward(level, B, False)

    def mmlrec_gate_level_backward_tiled(self, level, B, n_gates, n_experts, H, total_wg, total_ne, total_hg, scratch,
                                         stream):
        if not (0 < n_gates <= L.LEVEL_MAX_GATES and 0 < n_experts <= L.LEVEL_MAX_EXPERTS and H % 4 == 0
                and total_wg % 4 == 0 and total_hg % 4 == 0
                and self._real.mmlrec_gate_level_backward_tiled_smem(n_gates, n_experts, H, total_wg, total_ne,
                                                                     total_hg) <= 110 * 1024):
            return -1
        return self._gate_level_backward(level, B, True)

    # BatchNorm1d + activation (training statistics over the batch, momentum 0.1, eps 1e-5, unbiased running variance)
    def mmlrec_bn_forward(self, Z, ldz, M, N, gamma, beta, rmean, rvar, nbt, n_tracked, smean, sinv, Y, ldy, Y16, ldy16,
                          act, training, stream):
        self.calls.append("bn_forward")
        z = view2(Z, np.float32, M, N, ldz).astype(np.float64)
        g, b_ = view(gamma, np.float32, N).astype(np.float64), view(beta, np.float32, N).astype(np.float64)
        rm, rv = view(rmean, np.float32, N), view(rvar, np.float32, N)
        if training == 1:
            if M <= 1:
                return -1
            mean, var = z.mean(0), z.var(0)
            inv = 1.0 / np.sqrt(var + 1e-5)
            view(smean, np.float32, N)[:] = mean
            view(sinv, np.float32, N)[:] = inv
            rm[:] = (0.9 * rm + 0.1 * mean).astype(np.float32)
            rv[:] = (0.9 * rv + 0.1 * var * M / (M - 1)).astype(np.float32)
            view(nbt, np.int64, n_tracked)[:] += 1
        elif training == 2:
            mean, inv = view(smean, np.float32, N).astype(np.float64), view(sinv, np.float32, N).astype(np.float64)
        else:
            mean, inv = rm.astype(np.float64), 1.0 / np.sqrt(rv.astype(np.float64) + 1e-5)
        out = _act((z - mean) * inv * g + b_, act)
        if _ptr(Y):
            view2(Y, np.float32, M, N, ldy)[:] = out.astype(np.float32)
        if _ptr(Y16):
            view2(Y16, np.uint16, M, N, ldy16)[:] = f32_to_bf16(out).reshape(M, N)
        return 0

    def mmlrec_bn_backward(self, dY, lddy, Z, ldz, M, N, gamma, smean, sinv, dZ, lddz, dZ16, lddz16, dgamma, dbeta, stream):
        self.calls.append("bn_backward")
        dy = view2(dY, np.float32, M, N, lddy).astype(np.float64)
        z = view2(Z, np.float32, M, N, ldz).astype(np.float64)
        g = view(gamma, np.float32, N).astype(np.float64)
        mean, inv = view(smean, np.float32, N).astype(np.float64), view(sinv, np.float32, N).astype(np.float64)
        xhat = (z - mean) * inv
        dg, db = (dy * xhat).sum(0), dy.sum(0)
        view(dgamma, np.float32, N)[:] = dg
        view(dbeta, np.float32, N)[:] = db
        dz = g * inv * (dy - db / M - xhat * dg / M)
        if _ptr(dZ):
            view2(dZ, np.float32, M, N, lddz)[:] = dz.astype(np.float32)
        if _ptr(dZ16):
            view2(dZ16, np.uint16, M, N, lddz16)[:] = f32_to_bf16(dz).reshape(M, N)
        return 0

    # Synchronised BatchNorm (data parallel): per-rank moments -> all-gather -> combination; backward sums -> all-reduce
    def mmlrec_bn_stats(self, Z, ldz, M, N, stats, stream):
        self.calls.append("bn_stats")
        z = view2(Z, np.float32, M, N, ldz).astype(np.float64)
        st = view(stats, np.float32, 2 * N)
        st[:N] = z.mean(0)
        st[N:] = ((z - z.mean(0)) ** 2).sum(0)
        return 0

    def mmlrec_bn_combine(self, all_stats, R, M, N, rmean, rvar, nbt, n_tracked, smean, sinv, stream):
        self.calls.append("bn_combine")
        a = view(all_stats, np.float32, R * 2 * N).reshape(R, 2 * N).astype(np.float64)
        mean = a[:, :N].mean(0)
        m2 = (a[:, N:] + M * (a[:, :N] - mean) ** 2).sum(0)     # Chan's combination over equal-sized ranks
        total = float(M * R)
        view(smean, np.float32, N)[:] = mean
        view(sinv, np.float32, N)[:] = 1.0 / np.sqrt(m2 / total + 1e-5)
        rm, rv = view(rmean, np.float32, N), view(rvar, np.float32, N)
        rm[:] = (0.9 * rm + 0.1 * mean).astype(np.float32)
        rv[:] = (0.9 * rv + 0.1 * m2 / (total - 1)).astype(np.float32)
        view(nbt, np.int64, n_tracked)[:] += 1
        return 0

    def mmlrec_bn_backward_sums(self, dY, lddy, Z, ldz, M, N, smean, sinv, sums, dgamma, dbeta, stream):
        self.calls.append("bn_backward_sums")
        dy = view2(dY, np.float32, M, N, lddy).astype(np.float64)
        z = view2(Z, np.float32, M, N, ldz).astype(np.float64)
        xhat = (z - view(smean, np.float32, N)) * view(sinv, np.float32, N)
        s = view(sums, np.float32, 2 * N)
        s[:N], s[N:] = dy.sum(0), (dy * xhat).sum(0)
        view(dgamma, np.float32, N)[:] = s[N:]
        view(dbeta, np.float32, N)[:] = s[:N]
        return 0

    def mmlrec_bn_backward_synced(self, dY, lddy, Z, ldz, M, N, gamma, smean, sinv, dZ, lddz, dZ16, lddz16, gsums, M_total,
                                  stream):
        self.calls.append("bn_backward_synced")
        dy = view2(dY, np.float32, M, N, lddy).astype(np.float64)
        z = view2(Z, np.float32, M, N, ldz).astype(np.float64)
        inv = view(sinv, np.float32, N).astype(np.float64)
        xhat = (z - view(smean, np.float32, N)) * inv
        gs = view(gsums, np.float32, 2 * N).astype(np.float64)
        dz = view(gamma, np.float32, N) * inv * (dy - gs[:N] / M_total - xhat * gs[N:] / M_total)
        if _ptr(dZ):
            view2(dZ, np.float32, M, N, lddz)[:] = dz.astype(np.float32)
        if _ptr(dZ16):
            view2(dZ16, np.uint16, M, N, lddz16)[:] = f32_to_bf16(dz).reshape(M, N)
        return 0

    # PEPNet: element-wise product with the producers' activation derivatives folded into the gradient writes
    def mmlrec_mul_forward(self, a, lda, b, ldb, o32, ld32, o16, ld16, rows, cols, stream):
        self.calls.append("mul_forward")
        v = view2(a, np.float32, rows, cols, lda).astype(np.float64) * view2(b, np.float32, rows, cols, ldb)
        if _ptr(o32):
            view2(o32, np.float32, rows, cols, ld32)[:] = v.astype(np.float32)
        if _ptr(o16):
            view2(o16, np.uint16, rows, cols, ld16)[:] = f32_to_bf16(v).reshape(rows, cols)
        return 0

    def mmlrec_mul_backward(self, d_out, ld_dout, a, lda, b, ldb, da32, da16, ld_da, dkind_a, acc_a, db32, db16, ld_db,
                            dkind_b, acc_b, rows, cols, stream):
        self.calls.append("mul_backward")
        g = view2(d_out, np.float32, rows, cols, ld_dout).astype(np.float64)
        av = view2(a, np.float32, rows, cols, lda).astype(np.float64)
        bv = view2(b, np.float32, rows, cols, ldb).astype(np.float64)

        def fold(v, own, dkind):
            if dkind == 1:
                return np.where(own > 0, v, 0.0)
            if dkind == 2:
                return v * own * (1 - own / 2)
            return v
        for val, p32, p16, ld, acc in ((fold(g * bv, av, dkind_a), da32, da16, ld_da, acc_a),
                                       (fold(g * av, bv, dkind_b), db32, db16, ld_db, acc_b)):
            if _ptr(p32):
                dst = view2(p32, np.float32, rows, cols, ld)
                dst[:] = (dst + val).astype(np.float32) if acc else val.astype(np.float32)
            if _ptr(p16):
                assert not acc
                view2(p16, np.uint16, rows, cols, ld)[:] = f32_to_bf16(val).reshape(rows, cols)
        return 0

    # STAR: effective per-domain weights shared * specific and the fold of their gradients
    def mmlrec_star_weights(self, spec_ptrs, spec_b_ptrs, shared, shared_b, T, K, N, w_eff, ld_w, w16, b_eff, stream):
        self.calls.append("star_weights")
        sp, sb = view(spec_ptrs, np.int64, T), view(spec_b_ptrs, np.int64, T)
        sh = view(shared, np.float32, K * N).reshape(K, N).astype(np.float64)
        shb = view(shared_b, np.float32, N).astype(np.float64)
        for t in range(T):
            w = (view(int(sp[t]), np.float32, K * N).reshape(K, N) * sh).T      # [N, K]
            view2(_ptr(w_eff) + 4 * t * N * ld_w, np.float32, N, K, ld_w)[:] = w.astype(np.float32)
            if _ptr(w16):
                view2(_ptr(w16) + 2 * t * N * ld_w, np.uint16, N, K, ld_w)[:] = f32_to_bf16(w).reshape(N, K)
            view(_ptr(b_eff) + 4 * t * N, np.float32, N)[:] = (view(int(sb[t]), np.float32, N) + shb).astype(np.float32)
        return 0

    def mmlrec_star_fold(self, d_w_eff, ld_w, d_b_eff, spec_ptrs, shared, live, T, K, N, d_shared, d_shared_b, d_spec_last,
                         d_spec_b_last, stream):
        self.calls.append("star_fold")
        sp, lv = view(spec_ptrs, np.int64, T), view(live, np.int32, T)
        sh = view(shared, np.float32, K * N).reshape(K, N).astype(np.float64)
        acc, bacc = np.zeros((K, N)), np.zeros(N)
        for t in range(T):
            if not lv[t]:
                continue
            g = view2(_ptr(d_w_eff) + 4 * t * N * ld_w, np.float32, N, K, ld_w).astype(np.float64).T   # [K, N]
            acc += g * view(int(sp[t]), np.float32, K * N).reshape(K, N)
            gb = view(_ptr(d_b_eff) + 4 * t * N, np.float32, N).astype(np.float64)
            bacc += gb
            if t == T - 1:
                if _ptr(d_spec_last):
                    view(d_spec_last, np.float32, K * N)[:] = (g * sh).reshape(-1).astype(np.float32)
                if _ptr(d_spec_b_last):
                    view(d_spec_b_last, np.float32, N)[:] = gb.astype(np.float32)
        view(d_shared, np.float32, K * N)[:] = acc.reshape(-1).astype(np.float32)
        view(d_shared_b, np.float32, N)[:] = bacc.astype(np.float32)
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

    def __init__(self, model, B, precision, dp=None):
        cpu = torch.device("cpu")
        self.model, self.dp = model, dp   # dp: a parallel.DataParallelContext (gloo): Sync-BatchNorm + gradient all-reduce
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
        launches = self.launch_tables = []   # descriptor lists of the tensor-core launches, in planning order
        b.tc_table = lambda descs: ("captured", launches.append([_copy_desc(d) for d in descs]) or len(launches) - 1)
        b.tc_launch = lambda tbl, stream, stamps=None: fake.run_tc(launches[tbl[1]])
        b.dp = self.dp
        b.mask_domains = model.num_domains if getattr(model, "use_domain_mask", False) else 0
        model.build_graph(b)
        b.materialize()
        self.b, self.fake, self.stages = b, fake, b.stages
        self.gather = next(s for s in b.stages if isinstance(s, core.GatherStage))
        self.heads = next(s for s in b.stages if isinstance(s, core.HeadStage))
        for s in reversed(b.stages):
            s.plan_backward()
        self.grad_slices = max([getattr(s, "split_k", 1) for s in b.stages] + [1])
        return self

    def forward_backward(self, X, y, mask=None):
        self.gather.X.copy_(torch.as_tensor(X, dtype=torch.float32))
        self.heads.y.copy_(torch.as_tensor(y, dtype=torch.float32))
        if mask is not None:
            self.heads.mask.copy_(torch.as_tensor(mask, dtype=torch.float32))
        self.model.store.grad_slices.zero_()
        for s in self.stages:
            s.forward(0, True)
        for s in reversed(self.stages):
            if s is not self.gather:
                s.backward(0)
        self.model.store.live_slices = self.grad_slices
        if self.dp is not None:   # what StepPlan.train_step does for replicated tables: fold the slices, SUM all-reduce
            st = self.model.store
            for k in range(1, self.grad_slices):
                st.dense_grad.add_(st.grad_slices[k])
            st.live_slices = 1
            self.dp.sum_gradients(st.dense_grad)
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
    if d.c_transposed:
        assert d.C_f32 and not (d.C_bf16 or d.bias or d.mask or d.mask_bits or d.relu_bits_out or d.colsum or d.act
                                or d.accumulate) and (not d.colsum_b or d.N <= 128), \
            "a transposed store takes the plain fp32 product only"
    assert not d.relu_bits_out or d.bits_out_chunks > 0, "bits_out_chunks"
    assert not d.mask_bits or d.mask_bits_chunks > 0, "mask_bits_chunks"
    assert d.C_f32 or d.C_bf16, "a problem without an output"


def _copy_desc(d):
    _check_tc_desc(d)
    c = L.GemmTcDesc()
    C.memmove(C.byref(c), C.byref(d), C.sizeof(L.GemmTcDesc))
    return c
