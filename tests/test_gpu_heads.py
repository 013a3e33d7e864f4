"""Prediction heads on the GEMM kernels (detr_b200/heads.py) vs the plain modules in fp32 (detr/model.py:45-52,92-93,359-392)."""
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0, dtype=torch.bfloat16, dev="cuda"):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev, dtype)


def _close(got, ref, rel):
    got, ref = got.float(), ref.float()
    tol = rel * ref.abs().max().item() + 1e-7
    err = (got - ref).abs().max().item()
    assert err <= tol, (err, tol)


@pytest.mark.parametrize("M,N", [(4800, 92), (4800, 4), (130, 8), (1, 92), (7200, 92)])
def test_gemm_ragged_fp32_columns(cuda, M, N):
    """fp32 outputs whose width is not a multiple of 32 (the heads' 92 / 4 columns): TMA clips the last box, nothing is written
    beyond column N or row M (canary rows / the dense layout itself prove it)."""
    from detr_b200 import gemm as G
    K = 256
    a, w = _rand((M, K), 1), _rand((N, K), 2, 0.05)
    bias = torch.zeros(128, dtype=torch.float32, device="cuda")
    bias[:N] = _rand((N,), 3, 0.5, torch.float32)
    buf = torch.full((M + 2, N), 7.0, dtype=torch.float32, device="cuda")
    out = G.gemm(a, w, bias=bias, out=buf[1:M + 1])
    ref = a.float() @ w.float().t() + bias[:N]
    _close(out, ref, 2e-5)
    assert (buf[0] == 7.0).all() and (buf[M + 1] == 7.0).all()
    sig = G.gemm(a, w, bias=bias, epilogue=G.EPI_SIGMOID, out=torch.empty(M, N, dtype=torch.float32, device="cuda"))
    _close(sig, torch.sigmoid(ref), 2e-5)


def _modules(n_cls=92, seed=0):
    from detr_b200.harness import _MLP
    torch.manual_seed(seed)
    cls = nn.Linear(256, n_cls).cuda()
    mlp = _MLP(256, 256, 4, 3).cuda()
    with torch.no_grad():      # the reference's init (std 0.02, zero bias) makes every box 0.5: use something with signal
        for m in mlp.net:
            if isinstance(m, nn.Linear):
                m.weight.normal_(0, 0.08)
                m.bias.normal_(0, 0.1)
        cls.bias.normal_(0, 0.1)
    return cls, mlp


@pytest.mark.parametrize("lead,n_cls", [((2, 6, 100), 92), ((8, 6, 100), 92), ((4, 6, 300), 92), ((3, 50), 12)])
def test_heads_forward_backward(cuda, lead, n_cls):
    from detr_b200 import heads
    cls, mlp = _modules(n_cls)
    x0 = _rand((*lead, 256), 5, 1.0, torch.float32)
    gl, gb = _rand((*lead, n_cls), 6, 1.0, torch.float32), _rand((*lead, 4), 7, 1.0, torch.float32)

    def run(fused):
        for m in (cls, mlp):
            m.zero_grad(set_to_none=True)
        x = x0.clone().requires_grad_(True)
        if fused:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                assert heads.supported(x, cls, mlp)
                logits, boxes = heads.predict(x, cls, mlp)
        else:                      # fp32 modules on the bf16-rounded input and weights the kernels see
            logits, boxes = cls(x), mlp(x).sigmoid()
        (logits * gl).sum().add((boxes * gb).sum()).backward()
        return logits.detach(), boxes.detach(), x.grad, [p.grad.clone() for m in (cls, mlp) for p in m.parameters()]

    lf, bf, dxf, gf = run(True)
    assert lf.dtype == torch.float32 and bf.dtype == torch.float32 and lf.is_contiguous() and bf.is_contiguous()
    assert lf.shape == (*lead, n_cls) and bf.shape == (*lead, 4) and bf.data_ptr() % 16 == 0 and lf.data_ptr() % 16 == 0
    lr, br, dxr, gr = run(False)
    # bf16 operands, fp32 accumulation: ~3 bf16 ulps of the operand scale through three layers
    _close(lf, lr, 1e-2)
    _close(bf, br, 1e-2)
    _close(dxf, dxr, 2e-2)
    for a, b in zip(gf, gr):
        assert a.shape == b.shape
        _close(a, b, 2e-2)


def test_heads_feed_matcher_and_criterion_without_copies(cuda):
    """The heads' outputs are what `_rows()` wants: the criterion sees the very same storage (no .float() / .contiguous() copy)."""
    from detr_b200 import heads
    from detr_b200.matcher import _rows
    cls, mlp = _modules()
    x = _rand((2, 6, 100, 256), 8, 1.0, torch.float32)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits, boxes = heads.predict(x, cls, mlp)
    assert _rows(logits, 92).data_ptr() == logits.data_ptr() and _rows(boxes, 4).data_ptr() == boxes.data_ptr()
    assert logits.float().data_ptr() == logits.data_ptr()


def test_heads_fallback_outside_contract(cuda):
    from detr_b200 import heads
    cls, mlp = _modules()
    x = _rand((2, 10, 256), 9, 1.0, torch.float32)
    assert not heads.supported(x, cls, mlp)              # no autocast
    logits, boxes = heads.predict(x, cls, mlp)
    assert torch.equal(logits, cls(x)) and torch.equal(boxes, mlp(x).sigmoid())
