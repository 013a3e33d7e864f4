"""GPU parity: cost matrix, assignment and fused matcher through the C ABI, against the oracle, SciPy and the
reference's golden fixtures.  Bar: indices bit-exact; cost within 2e-6 abs (fp32, values O(1..10))."""
import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment as scipy_lsa

from oracle import detr_oracle as O
from util import golden_targets, load_golden

pytestmark = pytest.mark.gpu
COST_ATOL = 2e-6


def _pairs_equal(got, want):
    return np.array_equal(got[0].cpu().numpy(), np.asarray(want[0])) and np.array_equal(got[1].cpu().numpy(), np.asarray(want[1]))


@pytest.mark.parametrize("name", ["criterion_q100", "criterion_q20_tall"])
def test_lsap_kernel_on_reference_cost_matrices(cuda, name):
    """Feed the REFERENCE's cost matrices to the assignment kernel: indices must equal the reference's."""
    from detr_b200 import linear_sum_assignment_cuda
    fx = load_golden(name)
    keys = sorted(k for k in fx if k.startswith("cost/"))
    costs = [torch.from_numpy(fx[k]).to(cuda) for k in keys]
    got = linear_sum_assignment_cuda(costs)
    for k, g in zip(keys, got):
        want = (fx[k.replace("cost/", "idx_q/")], fx[k.replace("cost/", "idx_gt/")])
        assert g[0].dtype == torch.int64 and _pairs_equal(g, want), k


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("kind", ["uniform", "ties", "zeros", "quarter"])
def test_lsap_kernel_random_and_tie_heavy_vs_scipy(cuda, kind, dtype):
    from detr_b200 import linear_sum_assignment_cuda
    rng = np.random.default_rng(5)
    mats = []
    for _ in range(200):
        nr, nc = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        c = {"uniform": lambda: rng.random((nr, nc)) * 4 - 1, "ties": lambda: rng.integers(0, 3, (nr, nc)),
             "zeros": lambda: np.zeros((nr, nc)), "quarter": lambda: np.round(4 * rng.random((nr, nc))) / 4}[kind]()
        mats.append(c.astype(np.float32))
    got = linear_sum_assignment_cuda([torch.from_numpy(m).to(cuda, dtype) for m in mats])
    bad = [i for i, (m, g) in enumerate(zip(mats, got)) if not _pairs_equal(g, scipy_lsa(m))]
    assert not bad, f"{len(bad)} of {len(mats)} differ, first {bad[:5]}"


@pytest.mark.parametrize("shape", [(100, 1), (100, 100), (100, 130), (300, 100), (300, 300), (20, 600), (600, 40), (1, 1), (7, 0)])
def test_lsap_kernel_shapes(cuda, shape):
    from detr_b200 import linear_sum_assignment_cuda
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    mats = [(rng.random(shape) * 5 - 2).astype(np.float32) for _ in range(3)]
    got = linear_sum_assignment_cuda([torch.from_numpy(m).to(cuda) for m in mats])
    for m, g in zip(mats, got):
        assert _pairs_equal(g, scipy_lsa(m))


def test_lsap_kernel_error_conventions(cuda):
    from detr_b200 import linear_sum_assignment_cuda
    with pytest.raises(ValueError, match="invalid numeric"):
        linear_sum_assignment_cuda([torch.tensor([[float("nan"), 1.0], [1.0, 2.0]], device=cuda)])
    with pytest.raises(ValueError, match="invalid numeric"):
        linear_sum_assignment_cuda([torch.tensor([[float("-inf"), 1.0], [1.0, 2.0]], device=cuda)])
    with pytest.raises(ValueError, match="infeasible"):
        linear_sum_assignment_cuda([torch.tensor([[float("inf"), float("inf")], [1.0, 2.0]], device=cuda)])
    # +inf entries are legal as long as a finite assignment exists (same as SciPy)
    c = np.array([[np.inf, 1.0, 3.0], [2.0, np.inf, 1.0]], dtype=np.float32)
    g = linear_sum_assignment_cuda([torch.from_numpy(c).to(cuda)])[0]
    assert _pairs_equal(g, scipy_lsa(c))


@pytest.mark.parametrize("name", ["criterion_q100", "criterion_q20_tall", "criterion_allempty"])
def test_cost_matrix_and_fused_matcher_vs_golden(cuda, name):
    from detr_b200 import HungarianMatcher
    fx = load_golden(name)
    logits, boxes = torch.from_numpy(fx["logits"]).to(cuda), torch.from_numpy(fx["boxes"]).to(cuda)
    tg = golden_targets(fx, cuda)
    w = fx["matcher_w"].tolist()
    m = HungarianMatcher(cost_class=w[0], cost_bbox=w[1], cost_giou=w[2])
    B, L = logits.shape[:2]
    for l in range(L):
        costs = m.cost_matrices(logits[:, l], boxes[:, l], tg["class_idx"], tg["boxes_normalized"])
        idx = m(logits[:, l], boxes[:, l], tg["class_idx"], tg["boxes_normalized"])  # reference call signature
        assert len(idx) == B
        for b in range(B):
            np.testing.assert_allclose(costs[b].cpu().numpy(), fx[f"cost/{l}/{b}"], rtol=0, atol=COST_ATOL)
            assert idx[b][0].dtype == torch.int64 and idx[b][0].is_cuda
            assert _pairs_equal(idx[b], (fx[f"idx_q/{l}/{b}"], fx[f"idx_gt/{l}/{b}"])), (l, b)
    m.check_status()


def test_fused_matcher_all_layers_config3(cuda):
    """BASELINE config 3: batch 256, 100 queries x 1..100 GT, 6 layers, fp32: 1536 problems in one launch.
    (a) kernel costs vs oracle within tolerance; (b) kernel assignment == SciPy on the KERNEL's own cost
    matrices (bit-exact); (c) agreement with the oracle's end-to-end indices (near-tie flips reported)."""
    from detr_b200 import HungarianMatcher, pack_targets
    B, L, Q, NC = 256, 6, 100, 91
    logits, boxes = O.synth_predictions(B, L, Q, NC, seed=0)
    labels, gts = O.synth_targets(B, 100, NC, seed=1)
    m = HungarianMatcher(1.0, 5.0, 2.0)
    pt = pack_targets([l.to(cuda) for l in labels], [g.to(cuda) for g in gts], Q, cuda)
    iq, ig, cost = m.match_layers(logits.to(cuda), boxes.to(cuda), pt, export_cost=True)
    m.check_status()
    iq, ig, cost = iq.cpu().numpy(), ig.cpu().numpy(), cost.cpu().numpy()
    co = mo = 0
    worst, flips = 0.0, 0
    for b in range(B):
        M, n = pt.counts[b], pt.n_match[b]
        for l in range(L):
            C = cost[co:co + Q * M].reshape(Q, M); co += Q * M
            gq, gg = iq[mo:mo + n], ig[mo:mo + n]; mo += n
            r, c = scipy_lsa(C)
            assert np.array_equal(gq, r) and np.array_equal(gg, c), (b, l)
            if b % 16 == 0:  # oracle end-to-end on a sample (the oracle is a Python loop)
                oc = O.cost_matrix(logits[b, l], boxes[b, l], labels[b], gts[b], 1.0, 5.0, 2.0).numpy()
                worst = max(worst, float(np.abs(oc - C).max()))
                ro, cc = scipy_lsa(oc)
                flips += int(not (np.array_equal(ro, r) and np.array_equal(cc, c)))
    assert worst <= COST_ATOL, worst
    assert flips == 0, f"{flips} problems differ from the oracle end to end (near-tie flips)"


def test_matcher_edge_cases(cuda):
    from detr_b200 import HungarianMatcher
    m = HungarianMatcher(1.0, 5.0, 2.0)
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(3, 10, 6, generator=g).to(cuda)
    boxes = torch.rand(3, 10, 4, generator=g).mul(0.5).add(0.2).to(cuda)
    labels = [torch.zeros(0, dtype=torch.int64, device=cuda), torch.tensor([1, 2, 3], device=cuda), torch.randint(0, 5, (25,), generator=g).to(cuda)]
    gtb = O.synth_targets(3, 25, 5, seed=9)[1]
    gts = [torch.zeros(0, 4, device=cuda), gtb[1][:3].to(cuda) if len(gtb[1]) >= 3 else torch.tensor([[0.1, 0.1, 0.4, 0.4]] * 3, device=cuda),
           torch.cat([gtb[2], gtb[0], gtb[1]])[:25].to(cuda)]
    if gts[2].shape[0] < 25:
        gts[2] = torch.cat([gts[2], gts[2]])[:25]
    # non-contiguous per-layer slices, exactly how detr/loss.py:214-217 calls the matcher
    big_l = torch.randn(3, 4, 10, 6, generator=g).to(cuda); big_l[:, 2] = logits
    big_b = torch.rand(3, 4, 10, 4, generator=g).to(cuda); big_b[:, 2] = boxes
    idx = m(big_l[:, 2], big_b[:, 2], labels, gts)
    want = O.hungarian_match(logits.cpu(), boxes.cpu(), [l.cpu() for l in labels], [b.cpu() for b in gts], 1.0, 5.0, 2.0)
    assert idx[0][0].numel() == 0 and idx[0][1].numel() == 0 and idx[0][0].dtype == torch.int64
    assert idx[2][0].numel() == 10  # M > Q: Q pairs, no transpose
    for a, b in zip(idx, want):
        assert _pairs_equal(a, (b[0].numpy(), b[1].numpy()))
    m.check_status()
    with pytest.raises(AssertionError):
        HungarianMatcher(0, 0, 0)
    # return_cpu mirrors the reference's placement
    idx_cpu = HungarianMatcher(1.0, 5.0, 2.0, return_cpu=True)(logits, boxes, labels, gts)
    assert not idx_cpu[1][0].is_cuda


def test_matcher_fault_channel(cuda):
    """Degenerate boxes assert in the reference (detr/utils.py:87-88); NaN costs make SciPy raise ValueError."""
    from detr_b200 import HungarianMatcher
    m = HungarianMatcher(1.0, 5.0, 2.0)
    logits = torch.zeros(1, 4, 3, device=cuda)
    boxes = torch.full((1, 4, 4), 0.5, device=cuda)
    m(logits, boxes, [torch.tensor([0], device=cuda)], [torch.tensor([[0.5, 0.5, 0.2, 0.2]], device=cuda)])
    with pytest.raises(AssertionError):
        m.check_status()
    logits[0, 0, 0] = float("nan")
    m(logits, boxes, [torch.tensor([0], device=cuda)], [torch.tensor([[0.1, 0.1, 0.2, 0.2]], device=cuda)])
    with pytest.raises(ValueError, match="invalid numeric"):
        m.check_status()
    m.check_status()  # cleared
