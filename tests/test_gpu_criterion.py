"""GPU parity of the fused SetCriterion (forward + backward) against the reference's golden fixtures and the
oracle.  Bar (fp32): losses rel 1e-5; gradients 1e-6 abs + 1e-4 rel."""
import numpy as np
import pytest
import torch

from oracle import detr_oracle as O
from util import golden_losses, golden_targets, load_golden

pytestmark = pytest.mark.gpu
KEYS = 25


def _run(cuda, logits, boxes, targets, num_classes, w=(1.0, 5.0, 2.0)):
    from detr_b200 import HungarianMatcher, SetCriterion
    crit = SetCriterion(num_classes, HungarianMatcher(cost_class=w[0], cost_bbox=w[1], cost_giou=w[2]), 1.0, 5.0, 2.0, 0.1).to(cuda)
    lg = logits.to(cuda).requires_grad_(True)
    bx = boxes.to(cuda).requires_grad_(True)
    out = crit({"pred_logits": lg, "pred_boxes": bx}, targets)
    total = sum(v for k, v in out.items() if k.startswith("loss"))  # detr/train.py:262
    total.backward()
    crit.check_status()
    return crit, out, lg.grad, bx.grad


@pytest.mark.parametrize("name", ["criterion_q100", "criterion_q20_tall", "criterion_allempty"])
def test_criterion_vs_golden(cuda, name):
    fx = load_golden(name)
    nc = int(fx["num_classes"])
    crit, out, gl, gb = _run(cuda, torch.from_numpy(fx["logits"]), torch.from_numpy(fx["boxes"]), golden_targets(fx, cuda), nc,
                             fx["matcher_w"].tolist())
    ref = golden_losses(fx)
    assert set(out) == set(ref)
    for k, v in ref.items():
        assert out[k].dim() == 0 and out[k].is_cuda
        assert float(out[k]) == pytest.approx(v, rel=1e-5, abs=1e-6), k
    np.testing.assert_allclose(gl.cpu().numpy(), fx["grad_logits"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(gb.cpu().numpy(), fx["grad_boxes"], rtol=1e-4, atol=1e-6)
    # the assignment the criterion used is the reference's
    L = fx["logits"].shape[1]
    lists = crit.indices_as_lists()
    assert len(lists) == L
    for l in range(L):
        for b, (q, g) in enumerate(lists[l]):
            assert np.array_equal(q.cpu().numpy(), fx[f"idx_q/{l}/{b}"]) and np.array_equal(g.cpu().numpy(), fx[f"idx_gt/{l}/{b}"])


def test_criterion_config1_shape_vs_oracle(cuda):
    """BASELINE config 1 criterion inputs: batch 2, 6 layers, 100 queries, 92 logits, 1..20 GT."""
    B, L, Q, NC = 2, 6, 100, 91
    logits, boxes = O.synth_predictions(B, L, Q, NC, seed=4)
    labels, gts = O.synth_targets(B, 20, NC, seed=5)
    tg_cpu = {"class_idx": labels, "boxes_normalized": gts}
    tg = {"class_idx": [l.to(cuda) for l in labels], "boxes_normalized": [g.to(cuda) for g in gts]}
    _, out, gl, gb = _run(cuda, logits, boxes, tg, NC)
    lg = logits.clone().requires_grad_(True); bx = boxes.clone().requires_grad_(True)
    ref = O.set_criterion({"pred_logits": lg, "pred_boxes": bx}, tg_cpu, NC, (1.0, 5.0, 2.0))
    sum(v for k, v in ref.items() if k.startswith("loss")).backward()
    assert len(out) == KEYS and set(out) == set(ref)
    for k in ref:
        assert float(out[k]) == pytest.approx(float(ref[k]), rel=1e-5, abs=1e-6), k
    np.testing.assert_allclose(gl.cpu().numpy(), lg.grad.numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(gb.cpu().numpy(), bx.grad.numpy(), rtol=1e-4, atol=1e-6)


def test_criterion_config3_vs_oracle(cuda):
    """BASELINE config 3 (batch 256, 6 layers) : all 25 values and both gradients against the oracle."""
    B, L, Q, NC = 256, 6, 100, 91
    logits, boxes = O.synth_predictions(B, L, Q, NC, seed=0)
    labels, gts = O.synth_targets(B, 100, NC, seed=1)
    tg = {"class_idx": [l.to(cuda) for l in labels], "boxes_normalized": [g.to(cuda) for g in gts]}
    _, out, gl, gb = _run(cuda, logits, boxes, tg, NC)
    lg = logits.clone().requires_grad_(True); bx = boxes.clone().requires_grad_(True)
    ref = O.set_criterion({"pred_logits": lg, "pred_boxes": bx}, {"class_idx": labels, "boxes_normalized": gts}, NC, (1.0, 5.0, 2.0))
    sum(v for k, v in ref.items() if k.startswith("loss")).backward()
    for k in ref:
        assert float(out[k]) == pytest.approx(float(ref[k]), rel=2e-5, abs=1e-6), k
    np.testing.assert_allclose(gl.cpu().numpy(), lg.grad.numpy(), rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(gb.cpu().numpy(), bx.grad.numpy(), rtol=1e-4, atol=1e-7)


def test_criterion_upstream_gradients_and_weights(cuda):
    """Non-unit upstream gradients per key and non-default loss weights / eos_coef."""
    from detr_b200 import HungarianMatcher, SetCriterion
    B, L, Q, NC = 3, 2, 30, 11
    logits, boxes = O.synth_predictions(B, L, Q, NC, seed=7)
    labels, gts = O.synth_targets(B, 12, NC, seed=8, min_gt=0)
    crit = SetCriterion(NC, HungarianMatcher(2.0, 1.0, 3.0), 0.7, 2.5, 1.5, 0.3).to(cuda)
    lg = logits.to(cuda).requires_grad_(True); bx = boxes.to(cuda).requires_grad_(True)
    out = crit({"pred_logits": lg, "pred_boxes": bx}, {"class_idx": [l.to(cuda) for l in labels], "boxes_normalized": [g.to(cuda) for g in gts]})
    coef = {k: 0.5 + 0.25 * i for i, k in enumerate(sorted(k for k in out if k.startswith("loss")))}
    sum(coef[k] * out[k] for k in coef).backward()
    lg2 = logits.clone().requires_grad_(True); bx2 = boxes.clone().requires_grad_(True)
    ref = O.set_criterion({"pred_logits": lg2, "pred_boxes": bx2}, {"class_idx": labels, "boxes_normalized": gts}, NC,
                          (2.0, 1.0, 3.0), 0.3, 0.7, 2.5, 1.5)
    sum(coef[k] * ref[k] for k in coef).backward()
    for k in ref:
        assert float(out[k]) == pytest.approx(float(ref[k]), rel=1e-5, abs=1e-6), k
    np.testing.assert_allclose(lg.grad.cpu().numpy(), lg2.grad.numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(bx.grad.cpu().numpy(), bx2.grad.numpy(), rtol=1e-4, atol=1e-6)
    assert tuple(crit.state_dict()) == ("empty_weight",) and crit.empty_weight.shape == (NC + 1,)


def test_criterion_fault_poisons_losses(cuda):
    from detr_b200 import HungarianMatcher, SetCriterion
    crit = SetCriterion(2, HungarianMatcher(1.0, 5.0, 2.0)).to(cuda)
    logits = torch.zeros(1, 1, 4, 3, device=cuda); logits[0, 0, 0, 0] = float("nan")
    boxes = torch.full((1, 1, 4, 4), 0.3, device=cuda)
    out = crit({"pred_logits": logits, "pred_boxes": boxes},
               {"class_idx": [torch.tensor([0], device=cuda)], "boxes_normalized": [torch.tensor([[0.1, 0.1, 0.2, 0.2]], device=cuda)]})
    assert torch.isnan(out["loss_label_ce"]) and torch.isnan(out["loss_giou"])
    with pytest.raises(ValueError):
        crit.check_status()


@pytest.mark.parametrize("B,L,Q,NC,pad", [(2, 2, 300, 91, 0),    # config 5: 300 queries -> three register batches per warp
                                          (3, 2, 30, 10, 0),     # K = 11: rows that are not a multiple of 4 floats
                                          (2, 1, 20, 149, 0),    # K = 150: wider than the register path
                                          (2, 3, 100, 91, 5),    # logits are a strided view (row stride K + 5)
                                          (2, 2, 40, 63, 0)])    # K = 64: two full 32-wide chunks
def test_criterion_kernel_paths_vs_oracle(cuda, B, L, Q, NC, pad):
    """Every template instance of the criterion kernels (register chunks 1-4, generic rows, dense float4 / row-wise
    backward) against the oracle."""
    logits, boxes = O.synth_predictions(B, L, Q, NC, seed=11)
    labels, gts = O.synth_targets(B, min(Q, 40), NC, seed=12)
    tg = {"class_idx": [l.to(cuda) for l in labels], "boxes_normalized": [g.to(cuda) for g in gts]}
    from detr_b200 import HungarianMatcher, SetCriterion
    crit = SetCriterion(NC, HungarianMatcher(1.0, 5.0, 2.0)).to(cuda)
    if pad:
        big = torch.randn(B, L, Q, NC + 1 + pad, device=cuda)
        big[..., :NC + 1] = logits.to(cuda)
        big.requires_grad_(True)
        lg_in = big[..., :NC + 1]
        assert not lg_in.is_contiguous()
    else:
        big = logits.to(cuda).requires_grad_(True)
        lg_in = big
    bx = boxes.to(cuda).requires_grad_(True)
    out = crit({"pred_logits": lg_in, "pred_boxes": bx}, tg)
    sum(v for k, v in out.items() if k.startswith("loss")).backward()
    crit.check_status()
    lg2 = logits.clone().requires_grad_(True); bx2 = boxes.clone().requires_grad_(True)
    ref = O.set_criterion({"pred_logits": lg2, "pred_boxes": bx2}, {"class_idx": labels, "boxes_normalized": gts}, NC, (1.0, 5.0, 2.0))
    sum(v for k, v in ref.items() if k.startswith("loss")).backward()
    for k in ref:
        assert float(out[k]) == pytest.approx(float(ref[k]), rel=1e-5, abs=1e-6), k
    np.testing.assert_allclose(big.grad[..., :NC + 1].cpu().numpy(), lg2.grad.numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(bx.grad.cpu().numpy(), bx2.grad.numpy(), rtol=1e-4, atol=1e-6)
