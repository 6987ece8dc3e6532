"""GPU parity tests, kernel by kernel, through the C ABI (mmlrec_b200.ops -> libmmlrec_b200.so).
Bit-exact where the work is integer / byte / index (gather, sort); within a stated tolerance for
floating point, against the oracle (oracle/mmlrec_oracle.py) or plain fp32 torch on the same inputs."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import rel_err  # noqa: E402


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _tables(vocabs, D, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    offs, at, chunks = [], 0, []
    for v in vocabs:
        offs.append(at)
        chunks.append(torch.randn(v, D, generator=g))
        at += v * D
    emb = torch.cat([c.flatten() for c in chunks]).to(dev)
    return emb, offs, chunks


def _make_X(vocabs, n_dense, B, seed=1):
    g = torch.Generator().manual_seed(seed)
    cols = [torch.randint(0, v, (B,), generator=g).float() for v in vocabs]
    cols += [torch.rand(B, generator=g) for _ in range(n_dense)]
    return torch.stack(cols, 1)


# ------------------------------------------------------------------------------------------------ K1
@pytest.mark.parametrize("D", [4, 8, 16, 32])
@pytest.mark.parametrize("B", [1, 3, 257, 4096])
def test_gather_concat_bit_exact(D, B):
    from mmlrec_b200 import ops
    from oracle.mmlrec_oracle import Column, gather_concat
    dev = _cuda()
    vocabs, n_dense = [2, 53, 1000, 7, 100003], 3
    emb, offs, chunks = _tables(vocabs, D, dev)
    X = _make_X(vocabs, n_dense, B)
    cols = [Column(f"s{j}", "sparse", v, D) for j, v in enumerate(vocabs)] + [Column(f"d{j}", "dense") for j in range(n_dense)]
    params = {f"embedding_dict.s{j}.weight": c for j, c in enumerate(chunks)}
    want = gather_concat(X, params, cols)
    meta = ops.field_meta(offs, vocabs, list(range(len(vocabs))), D, dev)
    dense_cols = list(range(len(vocabs), len(vocabs) + n_dense))
    out, out16, oob = ops.gather_concat(X.to(dev), emb, meta, D, dense_cols, want_f32=True, want_bf16=True)
    assert int(oob.item()) == 0
    assert torch.equal(out.cpu(), want), "gathered rows must be bit-exact"
    in_dim = want.shape[1]
    assert torch.equal(out16[:, :in_dim].cpu(), want.to(torch.bfloat16))
    assert int(out16[:, in_dim:].float().abs().sum().item()) == 0, "bf16 padding columns must be zero"


def test_gather_out_of_range_id_sets_flag():
    from mmlrec_b200 import ops
    dev = _cuda()
    emb, offs, _ = _tables([5], 4, dev)
    X = torch.tensor([[1.0], [7.0], [-1.0]])
    meta = ops.field_meta(offs, [5], [0], 4, dev)
    out, _, oob = ops.gather_concat(X.to(dev), emb, meta, 4)
    assert int(oob.item()) == 1
    assert torch.equal(out[1].cpu(), emb[16:20].cpu()) and torch.equal(out[2].cpu(), emb[0:4].cpu())


def test_gather_fp32_id_carrier_truncates_like_long():
    """ids travel as fp32 and are truncated toward zero (basemodel.py:262, :476)."""
    from mmlrec_b200 import ops
    dev = _cuda()
    emb, offs, chunks = _tables([10], 4, dev)
    X = torch.tensor([[3.0], [3.9999], [0.5]])
    meta = ops.field_meta(offs, [10], [0], 4, dev)
    out, _, _ = ops.gather_concat(X.to(dev), emb, meta, 4)
    assert torch.equal(out.cpu(), chunks[0][X[:, 0].long()])


# ------------------------------------------------------------------------------------------------ sort
@pytest.mark.parametrize("B", [1, 2, 31, 1000, 4096, 5000, 20000])
def test_sort_field_ids(B):
    from mmlrec_b200 import ops
    dev = _cuda()
    vocabs = [2, 17, 100000]
    X = _make_X(vocabs, 0, B, seed=B)
    meta = ops.field_meta([0, 0, 0], vocabs, [0, 1, 2], 4, dev)
    ids, pos = ops.sort_field_ids(X.to(dev), meta)
    ids, pos = ids.cpu().long(), pos.cpu().long()
    for f in range(len(vocabs)):
        col = X[:, f].long()
        order = torch.argsort(col, stable=True)
        assert torch.equal(pos[f], order), "stable order by (id, position)"
        assert torch.equal(ids[f], col[order])


# ------------------------------------------------------------------------------------------------ K2
@pytest.mark.parametrize("D", [4, 8, 16])
@pytest.mark.parametrize("B,vocabs", [(700, [2, 3]), (3000, [2, 50, 100000]), (64, [5]), (4096, [1000, 4])])
def test_emb_segmented_reduce_matches_oracle(D, B, vocabs):
    from mmlrec_b200 import ops
    from oracle.mmlrec_oracle import Column, embedding_row_grads
    dev = _cuda()
    F = len(vocabs)
    X = _make_X(vocabs, 0, B, seed=7)
    g = torch.Generator().manual_seed(3)
    d_input = torch.randn(B, F * D, generator=g)
    cols = [Column(f"s{j}", "sparse", v, D) for j, v in enumerate(vocabs)]
    want = embedding_row_grads(X, d_input, cols)
    meta = ops.field_meta([0] * F, vocabs, list(range(F)), D, dev)
    Xd = X.to(dev)
    ids, pos = ops.sort_field_ids(Xd, meta)
    out = torch.zeros(F, B, D, device=dev)
    hy = ops.hyper_tensor("sgd", 0.1, dev)
    ops.emb_backward_update(d_input.to(dev), ids, pos, meta, D, hy, grad_rows_out=out)
    ids_c, out_c = ids.cpu().long(), out.cpu()
    for f in range(F):
        uniq, rows = want[f"s{f}"]
        tail = torch.ones(B, dtype=torch.bool)
        tail[:-1] = ids_c[f][1:] != ids_c[f][:-1]
        assert torch.equal(ids_c[f][tail], uniq)
        got = out_c[f][tail]
        assert rel_err(got, rows) < 2e-6
        assert float(out_c[f][~tail].abs().sum()) == 0.0, "only the run tails are written"


@pytest.mark.parametrize("opt", ["adagrad", "adam", "sgd"])
def test_emb_fused_update_matches_dense_torch_optimizer(opt):
    """The fused sparse row update (+ Adam's dense sweep) must equal torch.optim on the dense
    [V,D] gradient that embedding_dense_backward would have produced."""
    from mmlrec_b200 import ops
    dev = _cuda()
    D, B, vocabs, lr = 8, 512, [3, 40, 5000], 0.05
    emb, offs, chunks = _tables(vocabs, D, dev, seed=5)
    ref_tables = [torch.nn.Parameter(c.clone()) for c in chunks]
    ref_opt = {"adagrad": torch.optim.Adagrad, "adam": torch.optim.Adam, "sgd": torch.optim.SGD}[opt](ref_tables, lr=lr)
    meta = ops.field_meta(offs, vocabs, list(range(len(vocabs))), D, dev)
    hy = ops.hyper_tensor(opt, lr, dev)
    s1 = torch.zeros_like(emb) if opt != "sgd" else None
    s2 = torch.zeros_like(emb) if opt == "adam" else None
    touch = torch.full((emb.numel() // D,), -1, dtype=torch.int32, device=dev) if opt == "adam" else None
    for step in range(4):
        X = _make_X(vocabs, 0, B, seed=20 + step)
        d_input = torch.randn(B, len(vocabs) * D, generator=torch.Generator().manual_seed(50 + step)) * 0.01
        ref_opt.zero_grad()
        loss = sum((torch.nn.functional.embedding(X[:, f].long(), ref_tables[f]) * d_input[:, f * D:(f + 1) * D]).sum()
                   for f in range(len(vocabs)))
        loss.backward()
        ref_opt.step()
        ops.hyper_advance(hy)
        Xd = X.to(dev)
        ids, pos = ops.sort_field_ids(Xd, meta)
        ops.emb_backward_update(d_input.to(dev), ids, pos, meta, D, hy, emb=emb, s1=s1, s2=s2, row_touch=touch)
        if opt == "adam":
            ops.emb_adam_dense_sweep(emb, s1, s2, touch, D, hy)
    got = emb.cpu()
    for f, v in enumerate(vocabs):
        g = got[offs[f]:offs[f] + v * D].view(v, D)
        assert rel_err(g, ref_tables[f].detach()) < 1e-5, f"table {f}"


def test_lazy_adam_catch_up_is_bit_identical_to_the_dense_sweep():
    """Exact lazy dense-Adam: rows replay the zero-gradient steps they missed when they are next read (catch_up) or at
    a flush; the tables, both moments and every later update must equal -- bit for bit -- the per-step sweep over all
    untouched rows (the reference's nn.Embedding(sparse=False) + torch.optim.Adam semantics)."""
    from mmlrec_b200 import lib as L, ops
    dev = _cuda()
    lib = L.load()
    D, B, vocabs, lr, cap = 8, 256, [3, 40, 5000, 100000], 0.01, 64
    emb0, offs, _ = _tables(vocabs, D, dev, seed=5)
    meta = ops.field_meta(offs, vocabs, list(range(len(vocabs))), D, dev)
    st = torch.cuda.current_stream().cuda_stream
    state = {}
    for mode in ("sweep", "lazy"):
        emb = emb0.clone()
        s1, s2 = torch.zeros_like(emb), torch.zeros_like(emb)
        touch = torch.full((emb.numel() // D,), -1, dtype=torch.int32, device=dev)
        hy = ops.hyper_tensor("adam", lr, dev)
        hist = torch.zeros(2 * cap, device=dev)
        seen = []
        for step in range(150):           # > cap steps: a flush before the ring wraps, as the model does
            X = _make_X(vocabs, 0, B, seed=20 + step).to(dev)
            d_input = (torch.randn(B, len(vocabs) * D, generator=torch.Generator().manual_seed(50 + step)) * 0.01).to(dev)
            if mode == "lazy":
                if step % (cap - 2) == cap - 3:
                    L.check(lib.mmlrec_emb_adam_flush(emb.data_ptr(), s1.data_ptr(), s2.data_ptr(), touch.data_ptr(),
                                                      emb.numel() // D, D, hy.data_ptr(), hist.data_ptr(), cap, st))
                L.check(lib.mmlrec_hyper_advance_hist(hy.data_ptr(), hist.data_ptr(), cap, st))
                L.check(lib.mmlrec_emb_adam_catch_up(X.data_ptr(), X.stride(0), B, meta.data_ptr(), len(vocabs), D,
                                                     emb.data_ptr(), s1.data_ptr(), s2.data_ptr(), touch.data_ptr(),
                                                     hy.data_ptr(), hist.data_ptr(), cap, st))
            else:
                ops.hyper_advance(hy)
            # what the forward gather of this step reads must already be exact
            rows = torch.cat([emb[offs[f]:offs[f] + vocabs[f] * D].view(-1, D)[X[:, f].long()] for f in range(len(vocabs))], 1)
            seen.append(rows.clone())
            ids, pos = ops.sort_field_ids(X, meta)
            ops.emb_backward_update(d_input, ids, pos, meta, D, hy, emb=emb, s1=s1, s2=s2, row_touch=touch)
            if mode == "sweep":
                ops.emb_adam_dense_sweep(emb, s1, s2, touch, D, hy)
        if mode == "lazy":
            L.check(lib.mmlrec_emb_adam_flush(emb.data_ptr(), s1.data_ptr(), s2.data_ptr(), touch.data_ptr(),
                                              emb.numel() // D, D, hy.data_ptr(), hist.data_ptr(), cap, st))
        torch.cuda.synchronize()
        state[mode] = (emb.cpu(), s1.cpu(), s2.cpu(), [r.cpu() for r in seen])
    for a, b, what in zip(state["sweep"][:3], state["lazy"][:3], ("table", "exp_avg", "exp_avg_sq")):
        assert torch.equal(a, b), f"{what} differs between the sweep and the lazy catch-up"
    for i, (a, b) in enumerate(zip(state["sweep"][3], state["lazy"][3])):
        assert torch.equal(a, b), f"rows gathered at step {i} differ"


# ------------------------------------------------------------------------------------------------ K3 fp32
def test_gemm_grouped_f32_forward_dgrad_wgrad():
    from mmlrec_b200 import ops
    dev = _cuda()
    g = torch.Generator().manual_seed(0)
    B, K, N1, N2 = 300, 127, 130, 64
    x = torch.randn(B, K, generator=g).to(dev)
    W = torch.randn(N1 + N2, K, generator=g).to(dev)
    bias = torch.randn(N1 + N2, generator=g).to(dev)
    y = torch.empty(B, N1 + N2, device=dev)
    # forward: two problems reading the same x, writing column slices of one wide buffer
    ops.gemm_grouped_f32([
        dict(A=x, B=W[:N1], C=y[:, :N1], bias=bias[:N1], act="relu"),
        dict(A=x, B=W[N1:], C=y[:, N1:], bias=bias[N1:], act=None)])
    want = torch.cat([torch.relu(x @ W[:N1].T + bias[:N1]), x @ W[N1:].T + bias[N1:]], 1)
    assert rel_err(y, want) < 1e-5
    # dgrad with ReLU mask + accumulate, wgrad with bias gradient (row sums of dZ^T)
    dz = torch.randn(B, N1 + N2, generator=g).to(dev)
    prev = torch.randn(B, K, generator=g).to(dev)
    dx = prev.clone()
    mask = torch.randn(B, K, generator=g).to(dev)
    dW = torch.empty(N1 + N2, K, device=dev)
    db = torch.empty(N1 + N2, device=dev)
    ops.gemm_grouped_f32([
        dict(A=dz, B=W.T, C=dx, mask=mask, accumulate=True),
        dict(A=dz.T, B=x.T, C=dW, rowsum_a=db)])
    assert rel_err(dx, prev + (dz @ W) * (mask > 0)) < 1e-5
    assert rel_err(dW, dz.T @ x) < 1e-5
    assert rel_err(db, dz.sum(0)) < 1e-5


# ------------------------------------------------------------------------------------------------ K3 tensor cores
def _bf16(t):
    return t.to(torch.bfloat16)


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 256), (300, 200, 199), (4096, 24, 130), (70, 520, 1000),
                                   (520, 256, 320), (1000, 300, 72)])
def test_gemm_tc_layouts(a_mn, b_mn, M, N, K, kernel):
    """bf16 tcgen05 GEMM, every combination of K-major / MN-major operands, ragged sizes.  Inputs are
    exact bf16 values, so against fp32 torch only the accumulation order differs."""
    from mmlrec_b200 import ops
    dev = _cuda()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    pad = lambda n: (n + 7) // 8 * 8  # noqa: E731
    A = _bf16(torch.randn(M, K, generator=g)).to(dev)
    Bm = _bf16(torch.randn(N, K, generator=g)).to(dev)
    want = A.float() @ Bm.float().T
    if a_mn:
        store = torch.zeros(K, pad(M), dtype=torch.bfloat16, device=dev)
        store[:, :M] = A.T
        A_arg = store[:, :M]
    else:
        store = torch.zeros(M, pad(K), dtype=torch.bfloat16, device=dev)
        store[:, :K] = A
        A_arg = store[:, :K]
    if b_mn:
        storeb = torch.zeros(K, pad(N), dtype=torch.bfloat16, device=dev)
        storeb[:, :N] = Bm.T
        B_arg = storeb[:, :N]
    else:
        storeb = torch.zeros(N, pad(K), dtype=torch.bfloat16, device=dev)
        storeb[:, :K] = Bm
        B_arg = storeb[:, :K]
    C32 = torch.full((M, pad(N)), float("nan"), device=dev)
    C16 = torch.zeros(M, pad(N), dtype=torch.bfloat16, device=dev)
    rs = torch.full((M,), float("nan"), device=dev)
    d = ops.tc_desc(A_arg, B_arg, M, N, K, a_mn=a_mn, b_mn=b_mn, C_f32=C32, C_bf16=C16, rowsum_a=rs)
    ops.TcProblemTable([d], dev, kernel=kernel).launch()
    torch.cuda.synchronize()
    assert rel_err(C32[:, :N], want) < 1e-5
    assert rel_err(C16[:, :N].float(), want) < 5e-3
    assert rel_err(rs, A.float().sum(1)) < 1e-5, "row sums of A (bias gradient) from the all-ones MMA"
    assert torch.isnan(C32[:, N:]).all(), "columns beyond N are never written"


@pytest.mark.parametrize("kernel", [1, 2])
def test_gemm_tc_epilogue_bias_relu_mask_accumulate_grouped(kernel):
    from mmlrec_b200 import ops
    dev = _cuda()
    g = torch.Generator().manual_seed(11)
    M, N, K = 520, 192, 136
    A = _bf16(torch.randn(M, K, generator=g)).to(dev)
    W = _bf16(torch.randn(N, K, generator=g)).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    mask = _bf16(torch.randn(M, N, generator=g)).to(dev)
    prev = torch.randn(M, N, generator=g).to(dev)
    out1 = torch.empty(M, N, device=dev)
    out2 = prev.clone()
    descs = [ops.tc_desc(A, W, M, N, K, C_f32=out1, bias=bias, act="relu"),
             ops.tc_desc(A, W, M, N, K, C_f32=out2, mask=mask, accumulate=True)]
    tbl = ops.TcProblemTable(descs, dev, kernel=kernel)
    tbl.launch()
    torch.cuda.synchronize()
    base = A.float() @ W.float().T
    assert rel_err(out1, torch.relu(base + bias)) < 1e-5
    assert rel_err(out2, prev + base * (mask.float() > 0)) < 1e-5
    # persistent reuse: a second launch of the same table accumulates again
    tbl.launch()
    torch.cuda.synchronize()
    assert rel_err(out2, prev + 2 * base * (mask.float() > 0)) < 1e-5


@pytest.mark.parametrize("M,N,col0", [(520, 192, 0), (4096, 256, 64), (77, 100, 32)])
def test_gemm_tc_pair_relu_bitmask_roundtrip(M, N, col0):
    """CTA-pair kernel: a ReLU forward problem also emits 1 bit per output ("> 0"), laid out so that the 32 rows of a
    warp are contiguous; a dgrad problem masked by those bits equals the one masked by the bf16 activation itself."""
    from mmlrec_b200 import ops
    dev = _cuda()
    g = torch.Generator().manual_seed(M + N)
    K = 136
    A = _bf16(torch.randn(M, K, generator=g)).to(dev)
    W = _bf16(torch.randn(N, K, generator=g)).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    chunks = (col0 + N + 31) // 32 + 1
    bits = torch.zeros((M + 31) // 32 * chunks * 32, dtype=torch.int32, device=dev)
    y16 = torch.zeros(M, (N + 7) // 8 * 8, dtype=torch.bfloat16, device=dev)
    d = ops.tc_desc(A, W, M, N, K, C_bf16=y16, bias=bias, act="relu", relu_bits_out=bits, bits_chunks=chunks,
                    bits_chunk0=col0 // 32)
    ops.TcProblemTable([d], dev, kernel=2).launch()
    torch.cuda.synchronize()
    want = torch.relu(A.float() @ W.float().T + bias)
    b3 = bits.view(-1, chunks, 32).cpu()
    rows = torch.arange(M)
    got = torch.zeros(M, N, dtype=torch.bool)
    for c in range(N):
        word = b3[rows // 32, (col0 + c) // 32, rows % 32]
        j = (col0 + c) % 32                      # bit position inside the chunk word: even columns low half, odd high
        got[:, c] = ((word >> ((j >> 1) + 16 * (j & 1))) & 1).bool()
    # the bit says "the stored bf16 value is > 0" (= the fp32 result is > 0, up to values that round to bf16 zero)
    near = want.cpu().abs() < 1e-3
    assert torch.equal(got | near, (want.cpu() > 0) | near)
    # dgrad through the ReLU: mask by bits == mask by the bf16 activation (wherever the bf16 value did not round to 0)
    dz = _bf16(torch.randn(M, 72, generator=g)).to(dev)
    W2 = _bf16(torch.randn(N, 72, generator=g)).to(dev)     # dX[M,N] = dz @ W2^T
    out_bits = torch.empty(M, (N + 3) // 4 * 4, device=dev)
    out_mask = torch.empty_like(out_bits)
    d1 = ops.tc_desc(dz, W2, M, N, 72, C_f32=out_bits, mask_bits=bits, bits_chunks=chunks, bits_chunk0=col0 // 32)
    d2 = ops.tc_desc(dz, W2, M, N, 72, C_f32=out_mask, mask=y16)
    ops.TcProblemTable([d1, d2], dev, kernel=2).launch()
    torch.cuda.synchronize()
    base = dz.float() @ W2.float().T
    assert rel_err(out_mask[:, :N], base * (y16[:, :N].float() > 0)) < 1e-5
    assert rel_err(out_bits[:, :N], base * got.to(dev)) < 1e-5
    # the straight-line epilogue variants with bf16 outputs: ReLU forward without the bit array, dgrad masked by bits
    assert rel_err(y16[:, :N].float(), want) < 5e-3
    y16b = torch.zeros_like(y16)
    out16 = torch.zeros(M, (N + 7) // 8 * 8, dtype=torch.bfloat16, device=dev)
    d3 = ops.tc_desc(A, W, M, N, K, C_bf16=y16b, bias=bias, act="relu")
    d4 = ops.tc_desc(dz, W2, M, N, 72, C_bf16=out16, mask_bits=bits, bits_chunks=chunks, bits_chunk0=col0 // 32)
    ops.TcProblemTable([d3, d4], dev, kernel=2).launch()
    torch.cuda.synchronize()
    assert torch.equal(y16b[:, :N], y16[:, :N])
    assert torch.equal(out16[:, :N], out_bits[:, :N].to(torch.bfloat16)), "same fp32 values, rounded once"


@pytest.mark.parametrize("M,N,K", [(256, 128, 1024), (199, 64, 520), (300, 100, 136), (128, 8, 4096), (264, 1792, 512)])
def test_gemm_tc_pair_transposed_store_and_colsum(M, N, K):
    """Weight gradient of a layer with few outputs as the transposed product: D[m, n] = sum_k X[k, m] dZ[k, n] with both
    operands read MN-major, stored as C[n, m] straight from registers; for N <= 128 the bias gradient (column sums of
    dZ) comes out of the same problem through the all-ones A tile."""
    from mmlrec_b200 import ops
    dev = _cuda()
    g = torch.Generator().manual_seed(M + 3 * N + K)
    X = _bf16(torch.randn(K, (M + 7) // 8 * 8, generator=g)).to(dev)
    dZ = _bf16(torch.randn(K, (N + 7) // 8 * 8, generator=g)).to(dev)
    C = torch.full((N, (M + 3) // 4 * 4 + 4), float("nan"), device=dev)
    cs = torch.full((N,), float("nan"), device=dev) if N <= 128 else None
    d = ops.tc_desc(X[:, :M], dZ[:, :N], M, N, K, a_mn=True, b_mn=True, C_f32=C, c_transposed=True, colsum_b=cs)
    ops.TcProblemTable([d], dev, kernel=2).launch()
    torch.cuda.synchronize()
    want = dZ[:, :N].float().T @ X[:, :M].float()
    assert rel_err(C[:, :M], want) < 1e-5
    assert torch.isnan(C[:, M:]).all(), "nothing is written beyond column M"
    if cs is not None:
        assert rel_err(cs, dZ[:, :N].float().sum(0)) < 1e-5


# ------------------------------------------------------------------------------------------------ dense optimizer
@pytest.mark.parametrize("opt", ["adam", "adagrad", "sgd", "rmsprop"])
def test_dense_optimizer_matches_torch(opt):
    from mmlrec_b200 import ops
    dev = _cuda()
    g = torch.Generator().manual_seed(2)
    p0 = torch.randn(1000, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    cls = {"adam": torch.optim.Adam, "adagrad": torch.optim.Adagrad, "sgd": torch.optim.SGD, "rmsprop": torch.optim.RMSprop}[opt]
    ro = cls([ref], lr=0.01)
    p = p0.clone().to(dev)
    s1, s2 = torch.zeros_like(p), torch.zeros_like(p)
    hy = ops.hyper_tensor(opt, 0.01, dev)
    for step in range(5):
        grad = torch.randn(1000, generator=g) * 10 ** (-step)
        ref.grad = grad.clone()
        ro.step()
        ops.hyper_advance(hy)
        ops.dense_optimizer_step(p, grad.to(dev), s1 if opt != "sgd" else None, s2 if opt == "adam" else None, hy)
    assert rel_err(p.cpu(), ref.detach()) < 1e-6
