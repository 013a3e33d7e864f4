"""Host-side logic that needs no GPU: padded weight shadows, the contract checks of the prediction heads, the second-stream
bookkeeping outside a backward pass, the product path's refusal to run without CUDA."""
import os
import sys

import pytest
import torch
from torch import nn

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "detr-object-detection_b200"))


def test_shadowed_linears_padded_groups():
    """92- and 4-row weights occupy 128 / 64 rows of their shadows; the padding stays zero across refreshes; fp32 bias shadow."""
    from detr_b200.rowops import ShadowedLinears
    torch.manual_seed(0)
    cls, last, mid = nn.Linear(256, 92), nn.Linear(256, 4), nn.Linear(256, 256)
    sh = ShadowedLinears()
    sh.register("cls", [cls.weight], [cls.bias], pad_rows=[128])
    sh.register("mid", [mid.weight], [mid.bias])
    sh.register("last", [last.weight], [last.bias], pad_rows=[64])
    for _ in range(2):
        sh.refresh(torch.device("cpu"))
        w, b = sh.get_w_b32("cls")
        assert w.shape == (128, 256) and w.dtype == torch.bfloat16 and b.shape == (128,) and b.dtype == torch.float32
        assert torch.equal(w[:92], cls.weight.detach().to(torch.bfloat16)) and not w[92:].any()
        assert torch.equal(b[:92], cls.bias.detach()) and not b[92:].any()
        w3, b3 = sh.get_w_b32("last")
        assert w3.shape == (64, 256) and torch.equal(w3[:4], last.weight.detach().to(torch.bfloat16)) and not w3[4:].any() and not b3[4:].any()
        wm, bm = sh.get_w_b32("mid")
        assert wm.shape == (256, 256) and torch.equal(bm, mid.bias.detach())
        with torch.no_grad():          # an optimizer step between forwards: the next refresh picks it up
            cls.weight.add_(1.0)


def test_heads_contract_checks_and_fallback():
    from detr_b200 import heads
    from detr_b200.harness import _MLP
    cls, mlp = nn.Linear(256, 92), _MLP(256, 256, 4, 3)
    x = torch.randn(2, 6, 10, 256)
    assert not heads.supported(x, cls, mlp)                       # CPU tensor
    logits, boxes = heads.predict(x, cls, mlp)                    # -> the plain modules (detr/model.py:92-93)
    assert torch.equal(logits, cls(x)) and torch.equal(boxes, mlp(x).sigmoid())
    assert heads._mlp_linears(_MLP(256, 256, 4, 2)) is None       # not the reference's three layers
    assert heads._mlp_linears(nn.Sequential(nn.Linear(256, 4))) is None
    assert len(heads._mlp_linears(mlp)) == 3


def test_side_stream_is_inert_outside_backward():
    """`fork` hands out the second stream only inside a backward pass with the switch on; the switch restores."""
    from detr_b200 import gemm as G
    prev = G.wgrad_side_stream(True)
    try:
        assert G._SIDE.fork(torch.device("cpu"), ()) is None      # no graph task: launch in line
        assert not G._SIDE.pending
        G._SIDE.join_now()                                        # nothing pending: a no-op
        G._SIDE.join(12345)
    finally:
        assert G.wgrad_side_stream(prev) is True
    assert G._SIDE.enabled == prev


def test_product_path_refuses_cpu_tensors():
    """No CPU fallback: the kernels' wrappers raise on anything that is not a CUDA tensor."""
    from detr_b200 import gemm as G
    a, w = torch.randn(8, 64).bfloat16(), torch.randn(32, 64).bfloat16()
    with pytest.raises(RuntimeError):
        G.gemm(a, w)
    with pytest.raises(RuntimeError):
        G.gemm_wgrad(a, a)


def test_ln_gemm_partition_model():
    """detr_gemm_ln_partition (host only): row blocks of 32..128 rows, groups within the column tiles, one wave whenever one wave
    is possible, and the documented choices at the BASELINE shapes."""
    import ctypes
    from detr_b200 import _lib
    lib = _lib.load()
    fn = lib.detr_gemm_ln_partition

    def part(M, N, gelu, sms=148):
        r, g = ctypes.c_int(0), ctypes.c_int(0)
        assert fn(M, N, gelu, sms, ctypes.byref(r), ctypes.byref(g)) == 0
        return r.value, g.value

    for M in (1, 70, 800, 1200, 2400, 3400, 6800, 26800, 100000):
        for N in (256, 768, 1536, 2048):
            for gelu in (0, 1):
                rpc, groups = part(M, N, gelu)
                n_tiles = (N + 127) // 128
                assert rpc in (32, 64, 96, 128) and 1 <= groups <= n_tiles
                if -(-M // 128) <= 148:                      # one wave is possible: the model must not pick more
                    assert -(-M // rpc) * groups <= 148, (M, N, rpc, groups)
    assert part(6800, 768, 0) == (96, 2) and part(6800, 2048, 1) == (96, 2)     # 71 x 2 = 142 CTAs on 148 SMs
    assert part(800, 2048, 1) == (96, 16)                                       # decoder FFN: one column tile per CTA
    assert part(800, 256, 0)[0] == 32
