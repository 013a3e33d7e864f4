"""GPU parity of the tcgen05 GEMM kernels (csrc/gemm.cu) against plain fp32 PyTorch on the same bf16 operands.

What each kernel replaces: the four attention projections and the FFN of detr/model.py:312-314,354,405-411 with the pre-LN
LayerNorm (:221-224,173-182), bias, GELU(tanh), dropout and the residual add fused in.  Tolerances: operands are bf16 on both
sides and accumulation is fp32, so the products agree to fp32 rounding; what remains is the bf16 rounding of the OUTPUT
(2^-9 relative) -- 1e-2 * max|ref| absolute is ~2.5 output ulps.  Dropout masks are checked against the row-op kernels
(detr_epilogue_fwd / _bwd), which generate the same counter-based mask from the same seed."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0, dtype=torch.bfloat16, dev="cuda"):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev, dtype)


def _close(got, ref, rel=1e-2):
    got, ref = got.float(), ref.float()
    tol = rel * ref.abs().max().item() + 1e-6
    err = (got - ref).abs().max().item()
    assert err <= tol, (err, tol)
    return err


# (M, N, K): config 2 encoder (6800 rows) and decoder (800 rows) shapes, ragged M, one and many k-blocks, more tiles than SMs
SHAPES = [(6800, 256, 256), (6800, 512, 256), (6800, 2048, 256), (6800, 256, 2048), (800, 256, 256), (800, 1536, 256),
          (100, 128, 64), (129, 32, 128), (1, 96, 64), (26800, 256, 256)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_bias(cuda, M, N, K):
    from detr_b200 import gemm as G
    a, w = _rand((M, K), 1), _rand((N, K), 2, 0.05)
    bias = _rand((N,), 3, 0.5, torch.float32)
    out = G.gemm(a, w, bias=bias)
    ref = a.float() @ w.float().t() + bias
    _close(out, ref)
    out32 = G.gemm(a, w, bias=None, out_dtype=torch.float32)
    assert out32.dtype == torch.float32
    _close(out32, a.float() @ w.float().t(), rel=2e-5)          # fp32 output: only the accumulation order differs


@pytest.mark.parametrize("M,N,K", [(6800, 256, 512), (800, 256, 2048), (6800, 256, 1536), (77, 64, 64)])
def test_gemm_dgrad_form(cuda, M, N, K):
    """C = A . B with B given as [K][N] (the weight itself in dX = dY . W): MN-major B operand."""
    from detr_b200 import gemm as G
    a, w = _rand((M, K), 4), _rand((K, N), 5, 0.05)
    out = G.gemm(a, w, b_kn=True, out_dtype=torch.float32)
    _close(out, a.float() @ w.float(), rel=2e-5)


def test_gemm_strided_operands(cuda):
    """A and the output as column slices of wider buffers (the fused q|k|v projection output, stacked weights)."""
    from detr_b200 import gemm as G
    big = _rand((800, 768), 6)
    a = big[:, 256:512]
    w = _rand((512, 256), 7, 0.05)
    outbuf = torch.zeros(800, 1024, dtype=torch.bfloat16, device=cuda)
    G.gemm(a, w, out=outbuf[:, 512:])
    _close(outbuf[:, 512:], a.float() @ w.float().t())
    assert outbuf[:, :512].abs().max().item() == 0.0


@pytest.mark.parametrize("res_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("p", [0.0, 0.1])
def test_gemm_residual_dropout(cuda, res_dtype, p):
    from detr_b200 import _lib, gemm as G
    M, N, K = 1700, 256, 256
    a, w = _rand((M, K), 8), _rand((N, K), 9, 0.05)
    bias = _rand((N,), 10, 0.5, torch.float32)
    res = _rand((M, N), 11, 1.0, res_dtype)
    seed = 1234567
    out = G.gemm(a, w, epilogue=G.EPI_RES, bias=bias, res=res, p=p, seed=seed)
    assert out.dtype == res_dtype
    y = (a.float() @ w.float().t() + bias)
    if p == 0.0:
        _close(out, res.float() + y)
        return
    # same mask as the row-op kernel on the same (M, N) index space
    y16 = y.to(torch.bfloat16)
    ref = torch.empty_like(res)
    _lib.call("detr_epilogue_fwd", 0, res.data_ptr(), 0 if res_dtype == torch.float32 else 1, y16.data_ptr(), ref.data_ptr(), M, N, float(p),
              seed, None, _lib.stream_ptr())
    _close(out, ref, rel=1.5e-2)
    kept = ((out.float() - res.float()).abs() > 0).float().mean().item()
    assert abs(kept - 0.9) < 0.004, kept        # p = 3277 / 32768 = 0.100006 (15 random bits per element)


@pytest.mark.parametrize("p", [0.0, 0.1])
def test_gemm_gelu_and_backward(cuda, p):
    from detr_b200 import _lib, gemm as G
    M, N, K = 1700, 2048, 256
    a, w = _rand((M, K), 12), _rand((N, K), 13, 0.06)
    bias = _rand((N,), 14, 0.3, torch.float32)
    seed = 99
    aux = torch.empty(M, N, dtype=torch.bfloat16, device=cuda)
    out = G.gemm(a, w, epilogue=G.EPI_GELU, bias=bias, aux=aux, p=p, seed=seed)
    y = a.float() @ w.float().t() + bias
    _close(aux, y)
    ref = torch.empty(M, N, dtype=torch.bfloat16, device=cuda)
    _lib.call("detr_epilogue_fwd", 1, None, 1, aux.data_ptr(), ref.data_ptr(), M, N, float(p), seed, None, _lib.stream_ptr())
    _close(out, ref, rel=2e-3)   # same bf16 pre-activation, same mask: only the output rounding of equal fp32 values can differ
    if p == 0.0:
        _close(out, F.gelu(aux.float(), approximate="tanh"), rel=6e-3)
    # backward through GELU + dropout: dh = (g @ W2) * mask/(1-p) * gelu'(y) with g (M, 256), W2 (256, 2048)
    g, w2 = _rand((M, 256), 15), _rand((256, N), 16, 0.05)
    dh = G.gemm(g, w2, b_kn=True, epilogue=G.EPI_GELU_BWD, aux=aux, p=p, seed=seed)
    gin = (g.float() @ w2.float()).contiguous()
    ref_dh = torch.empty(M, N, dtype=torch.bfloat16, device=cuda)
    chunks = _lib.load().detr_epilogue_chunks(M, N)
    partial = torch.empty(chunks * N, dtype=torch.float32, device=cuda)
    db = torch.empty(N, dtype=torch.float32, device=cuda)
    _lib.call("detr_epilogue_bwd", 1, gin.data_ptr(), 0, aux.data_ptr(), ref_dh.data_ptr(), partial.data_ptr(), db.data_ptr(),
              _lib.zero_counters(cuda).data_ptr(), M, N, float(p), seed, None, _lib.stream_ptr())
    _close(dh, ref_dh, rel=1e-2)


def _ln_ref(x, gamma, beta, eps):
    return F.layer_norm(x.float(), (x.shape[-1],), gamma, beta, eps)


@pytest.mark.parametrize("x_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,n_pos_end,rpb", [(6800, 768, 512, 850), (800, 768, 512, 100), (800, 256, 256, 100), (6800, 2048, 0, 0),
                                               (130, 768, 512, 65), (26800, 768, 512, 3350)])
def test_gemm_ln_prologue(cuda, x_dtype, M, N, n_pos_end, rpb):
    """LN (+ addend) prologue: q|k columns from LN(x)+pos, v columns from LN(x); side outputs (operands, stats)."""
    from detr_b200 import gemm as G
    x = (_rand((M, 256), 20, 1.0, torch.float32) * 2.0 + 0.3).to(x_dtype)
    gamma = 1.0 + 0.1 * _rand((256,), 21, 1.0, torch.float32)
    beta = 0.1 * _rand((256,), 22, 1.0, torch.float32)
    w = _rand((N, 256), 23, 0.05)
    bias = _rand((N,), 24, 0.2, torch.float32)
    addend, sb, sr = None, 0, 0
    if n_pos_end:
        B = M // rpb
        if rpb == 100:   # decoder: a (Q, C) embedding broadcast over the batch
            emb = _rand((rpb, 256), 25, 1.0, torch.float32)
            addend, sb, sr = emb, 0, emb.stride(0)
            add_full = emb.unsqueeze(0).expand(B, -1, -1).reshape(M, 256)
        else:
            addend = _rand((B, rpb, 256), 25, 1.0, torch.float32)
            sb, sr = addend.stride(0), addend.stride(1)
            add_full = addend.reshape(M, 256)
    out, a_plain, a_pos, stats = G.gemm_ln(x, gamma, beta, 1e-5, w, addend=addend, rows_per_batch=rpb, add_sb=sb, add_sr=sr,
                                           n_pos_end=n_pos_end, bias=bias)
    ln = _ln_ref(x, gamma, beta, 1e-5)
    ref_plain = ln.to(torch.bfloat16)
    cols = []
    if n_pos_end:
        ref_pos = (ln + add_full).to(torch.bfloat16)
        cols.append(ref_pos.float() @ w[:n_pos_end].float().t())
        assert (a_pos.float() - ref_pos.float()).abs().max().item() <= 2 ** -7 * ref_pos.float().abs().max().item()
    if n_pos_end < N:
        cols.append(ref_plain.float() @ w[n_pos_end:].float().t())
        assert (a_plain.float() - ref_plain.float()).abs().max().item() <= 2 ** -7 * ref_plain.float().abs().max().item()
    ref = torch.cat(cols, dim=1) + bias
    _close(out, ref, rel=1.5e-2)     # + one bf16 ulp of operand rounding differences (LN computed in a different order)
    xf = x.float()
    assert (stats[0] - xf.mean(1)).abs().max().item() <= 1e-4
    assert (stats[1] - (xf.var(1, unbiased=False) + 1e-5).rsqrt()).abs().max().item() <= 1e-3


def test_gemm_ln_gelu(cuda):
    from detr_b200 import gemm as G
    M, N = 1700, 2048
    x = _rand((M, 256), 30, 1.5)
    gamma = 1.0 + 0.1 * _rand((256,), 31, 1.0, torch.float32)
    beta = 0.1 * _rand((256,), 32, 1.0, torch.float32)
    w = _rand((N, 256), 33, 0.06)
    bias = _rand((N,), 34, 0.3, torch.float32)
    aux = torch.empty(M, N, dtype=torch.bfloat16, device=cuda)
    out, a_plain, a_pos, stats = G.gemm_ln(x, gamma, beta, 1e-5, w, epilogue=G.EPI_GELU, bias=bias, aux=aux)
    assert a_pos is None
    y = a_plain.float() @ w.float().t() + bias       # on the kernel's own operand: isolates the GEMM + epilogue
    _close(aux, y)
    _close(out, F.gelu(aux.float(), approximate="tanh"), rel=6e-3)


@pytest.mark.parametrize("M,N,K", [(6800, 256, 256), (6800, 2048, 256), (6800, 256, 2048), (800, 768, 256), (6800, 1536, 256), (70, 64, 64),
                                   (26800, 256, 256)])
def test_gemm_wgrad(cuda, M, N, K):
    from detr_b200 import gemm as G
    dy, x = _rand((M, N), 40, 0.1), _rand((M, K), 41)
    dw, db = G.gemm_wgrad(dy, x)
    _close(dw, dy.float().t() @ x.float(), rel=1e-4)
    _close(db, dy.float().sum(0), rel=1e-4)
    dw2, db2 = G.gemm_wgrad(dy, x)
    assert torch.equal(dw, dw2) and torch.equal(db, db2)       # fixed summation order


def test_gemm_wgrad_two_operands(cuda):
    """Fused q|k|v projection: the q/k rows of dW contract with LN(x)+pos, the v rows with LN(x)."""
    from detr_b200 import gemm as G
    M = 6800
    dy, x0, x1 = _rand((M, 768), 42, 0.1), _rand((M, 256), 43), _rand((M, 256), 44)
    dw, db = G.gemm_wgrad(dy, x0, x1, n_switch=512)
    ref = torch.cat([dy[:, :512].float().t() @ x0.float(), dy[:, 512:].float().t() @ x1.float()], 0)
    _close(dw, ref, rel=1e-4)
    _close(db, dy.float().sum(0), rel=1e-4)


def test_wgrad_side_stream(cuda):
    """Weight gradients launched on the second stream inside a backward pass (gemm._SideStream) are joined before backward() returns
    and are bit-identical to the in-line launch; outside a backward pass the switch changes nothing."""
    from detr_b200 import blocks, gemm as G
    torch.manual_seed(3)
    norm = torch.nn.LayerNorm(256).cuda()
    fc1, fc2 = torch.nn.Linear(256, 2048).cuda(), torch.nn.Linear(2048, 256).cuda()
    x0 = torch.randn(2, 400, 256, device="cuda")
    grads = []
    for side in (False, True, True):
        for m in (norm, fc1, fc2):
            m.zero_grad(set_to_none=True)
        x = x0.clone().requires_grad_(True)
        prev = G.wgrad_side_stream(side)
        try:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = blocks.ln_ffn(x * 1.0, norm, fc1, fc2, 0.0, 0.0)
            y.float().square().sum().backward()
            # consumer on the caller's stream right after backward(): must already see the side stream's results
            got = [p.grad.clone() for m in (norm, fc1, fc2) for p in m.parameters()] + [x.grad.clone()]
        finally:
            G.wgrad_side_stream(prev)
        assert not G._SIDE.pending
        grads.append(got)
    for a, b in zip(grads[0], grads[1]):
        assert torch.equal(a, b)
    for a, b in zip(grads[1], grads[2]):
        assert torch.equal(a, b)
    G.wgrad_side_stream(True)
    try:
        dy, xx = _rand((800, 256), 50, 0.1), _rand((800, 256), 51)
        dw, _ = G.gemm_wgrad(dy, xx)           # not inside backward: launched in line
        _close(dw, dy.float().t() @ xx.float(), rel=1e-4)
    finally:
        G.wgrad_side_stream(False)
