"""The CUDA-graph-captured training step (harness.GraphedTrainStep) against the eager step (harness.train_step, the
replay of detr/train.py:258-267) from identical initial weights: same losses step after step (dropout off so the two
are comparable), fresh ground truth between replays, fresh dropout masks per replay in train mode."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(cuda, train):
    from detr_b200 import HungarianMatcher, SetCriterion
    from detr_b200.harness import DetrHarness
    from detr_b200.model import DETRConfig
    torch.manual_seed(0)
    cfg = DETRConfig(num_classes=11, num_object_queries=20, num_encoder_layers=2, num_decoder_layers=2)
    model = DetrHarness(cfg).to(cuda).to(memory_format=torch.channels_last)
    model.train(train)
    crit = SetCriterion(11, HungarianMatcher(1.0, 5.0, 2.0)).to(cuda)
    return model, crit


def test_graphed_step_matches_eager(cuda):
    from detr_b200.harness import GraphedTrainStep, batch_to, make_optimizer, synthetic_batch, train_step
    m1, c1 = _make(cuda, train=False)
    m2, c2 = copy.deepcopy(m1), copy.deepcopy(c1)
    batches = [synthetic_batch(2, 160, 200, 11, 6, seed=s) for s in (1, 2, 3)]
    o1 = make_optimizer(m1, lr=1e-4)
    eager = [float(train_step(m1, c1, o1, batch_to(b, cuda))) for b in batches]
    o2 = make_optimizer(m2, lr=1e-4, capturable=True)
    snapshot = copy.deepcopy(m2.state_dict())
    g = GraphedTrainStep(m2, c2, o2, batches[0], gt_cap=8, warmup=2)
    # the constructor's warm-up iterations are real optimizer steps, but it puts parameters, buffers, moments and step
    # counts back afterwards: training starts from exactly the state that was handed in
    for k, v in m2.state_dict().items():
        assert torch.equal(v, snapshot[k]), k
    assert float(g.fopt.step_t) == 0.0 and float(g.fopt.flat_m.abs().max()) == 0.0 and float(g.fopt.flat_v.abs().max()) == 0.0
    graphed = []
    for b in batches:
        g.load(b)
        graphed.append(float(g.step()))
    for a, b in zip(eager, graphed):
        assert a == pytest.approx(b, rel=2e-2), (eager, graphed)
    assert g.own_launches_per_step > 0
    c2.check_status()


def test_graph_replays_draw_fresh_dropout_masks(cuda):
    from detr_b200.harness import GraphedTrainStep, make_optimizer, synthetic_batch
    m, c = _make(cuda, train=True)
    for p in m.parameters():
        p.requires_grad_(True)
    opt = make_optimizer(m, lr=0.0, weight_decay=0.0, capturable=True)   # lr 0: weights frozen, only the masks change
    b = synthetic_batch(2, 160, 200, 11, 6, seed=4)
    g = GraphedTrainStep(m, c, opt, b, gt_cap=8, warmup=2)
    losses = [float(g.step()) for _ in range(4)]
    assert len(set(round(l, 6) for l in losses)) > 1, losses
    assert int(g.step_counter.item()) == 4        # one per replay; the warm-up's increments were rolled back


def test_prefetch_commit_equals_load(cuda):
    """The double-buffered input path (prefetch on a copy stream + commit) feeds the captured step the same inputs as the
    blocking load(): with frozen weights (lr 0, eval) the loss of a batch is identical either way, in any order."""
    from detr_b200.harness import GraphedTrainStep, make_optimizer, synthetic_batch
    m, c = _make(cuda, train=False)
    opt = make_optimizer(m, lr=0.0, weight_decay=0.0, capturable=True)
    batches = [synthetic_batch(2, 160, 200, 11, 6, seed=s, pin=True) for s in (5, 6, 7)]
    g = GraphedTrainStep(m, c, opt, batches[0], gt_cap=8, warmup=2)
    ref = []
    for b in batches:
        g.load(b)
        ref.append(float(g.step()))
    got = []
    g.prefetch(batches[0])
    for i in range(len(batches)):
        g.commit()
        if i + 1 < len(batches):
            g.prefetch(batches[i + 1])
        got.append(float(g.step()))
    assert got == ref, (got, ref)


def test_flat_adamw_matches_torch(cuda):
    """clip_grad_norm_(1.0) + torch.optim.AdamW against the flat two-launch optimizer (detr_sumsq_f32 + detr_adamw_clip_f32):
    same parameters after several steps, two groups with different learning rates, channels_last conv weights included."""
    from detr_b200.harness import FlatAdamW
    torch.manual_seed(0)
    def make():
        torch.manual_seed(1)
        a = torch.nn.Conv2d(8, 16, 3).to(cuda).to(memory_format=torch.channels_last)
        b = torch.nn.Linear(37, 19).to(cuda)
        return a, b
    (a1, b1), (a2, b2) = make(), make()
    groups = lambda a, b: [{"params": list(b.parameters()), "lr": 3e-3}, {"params": list(a.parameters()), "lr": 3e-4}]
    o1 = torch.optim.AdamW(groups(a1, b1), lr=3e-3, weight_decay=1e-2)
    o2 = torch.optim.AdamW(groups(a2, b2), lr=3e-3, weight_decay=1e-2)
    f = FlatAdamW(o2, cuda)
    p1 = [q for g in o1.param_groups for q in g["params"]]
    for it in range(5):
        if it == 3:   # an LR scheduler step: both optimizers must follow it
            for o in (o1, o2):
                for grp in o.param_groups:
                    grp["lr"] *= 0.5
        gs = [torch.randn_like(q) * (3.0 if it % 2 else 0.01) for q in p1]     # clipped and unclipped steps
        for q, g in zip(p1, gs):
            q.grad = g.clone()
        torch.nn.utils.clip_grad_norm_(p1, 1.0)
        o1.step()
        for v, g in zip(f.grad_views, gs):
            v.copy_(g)
        f.sync_lr()
        f.step(1.0)
    for q1, q2 in zip(p1, f.params):
        assert q1.shape == q2.shape and q1.stride() == q2.stride()
        assert torch.allclose(q1, q2, rtol=1e-4, atol=1e-6), (q1 - q2).abs().max()


def test_load_without_sync_keeps_batches_apart(cuda):
    """load() rewrites pinned staging on the host and enqueues async copies; with the host running steps ahead
    of the device the next load() must not overwrite staging an enqueued copy has yet to read.  Queue several DISTINCT batches
    back to back with no synchronisation in between (a slow kernel holds the stream back) and compare every step's loss with
    the fully synchronised run."""
    from detr_b200.harness import GraphedTrainStep, make_optimizer, synthetic_batch
    m, c = _make(cuda, train=False)
    opt = make_optimizer(m, lr=0.0, weight_decay=0.0, capturable=True)      # frozen weights: a batch's loss identifies it
    batches = [synthetic_batch(2, 160, 200, 11, 6, seed=s, pin=True) for s in range(20, 28)]
    g = GraphedTrainStep(m, c, opt, batches[0], gt_cap=8, warmup=2)
    ref = []
    for b in batches:
        g.load(b)
        ref.append(float(g.step()))
        torch.cuda.synchronize()
    assert len(set(round(r, 5) for r in ref)) > 1
    got = []
    torch.cuda._sleep(int(2e8))                 # ~0.1 s of device work queued first: every load() below runs ahead of the GPU
    for b in batches:
        g.load(b)
        got.append(g.step().clone())
    torch.cuda.synchronize()
    assert [float(x) for x in got] == ref


def test_faulty_batch_is_skipped_and_reported(cuda):
    """a data fault (here a degenerate ground-truth box) poisons that step's losses; the optimizer kernels must
    skip the update (weights, moments and step count untouched), the status word must not stay set, the fault is raised by the
    next step()/poll_faults(), and the following clean batch trains normally."""
    from detr_b200.harness import GraphedTrainStep, make_optimizer, synthetic_batch
    m, c = _make(cuda, train=False)
    opt = make_optimizer(m, lr=1e-4, capturable=True)
    good = synthetic_batch(2, 160, 200, 11, 6, seed=31)
    bad = synthetic_batch(2, 160, 200, 11, 6, seed=32)
    bad["boxes_normalized"][0][0] = torch.tensor([0.6, 0.6, 0.2, 0.2])      # x2 < x1: detr/utils.py:87-88 would assert
    g = GraphedTrainStep(m, c, opt, good, gt_cap=8, warmup=2)
    g.load(good); l0 = float(g.step())
    assert l0 == l0 and float(g.fopt.step_t) == 1.0
    before = g.fopt.flat_p.clone()
    g.load(bad); l1 = float(g.step())
    assert l1 != l1                                                            # NaN: poisoned
    assert torch.equal(g.fopt.flat_p, before) and float(g.fopt.step_t) == 1.0  # update skipped
    with pytest.raises(AssertionError):
        g.poll_faults(wait=True)
    g.load(good); l2 = float(g.step())
    assert l2 == l2 and float(g.fopt.step_t) == 2.0 and not torch.equal(g.fopt.flat_p, before)
    g.poll_faults(wait=True)                                                   # nothing pending


def test_gradient_accumulation_matches_single_step(cuda):
    """detr/train.py:116,258: `accumulate` micro-batches per optimizer step.  The same micro-batch twice with accumulate = 2 gives
    the gradient of that batch again (sum of two equal gradients times grad_div = 1/2): same update as accumulate = 1, no update on
    the non-boundary micro-step."""
    from detr_b200.harness import GraphedTrainStep, make_optimizer, synthetic_batch
    b = synthetic_batch(2, 160, 200, 11, 6, seed=41)
    m1, c1 = _make(cuda, train=False)
    m2, c2 = copy.deepcopy(m1), copy.deepcopy(c1)
    g1 = GraphedTrainStep(m1, c1, make_optimizer(m1, lr=1e-4, capturable=True), b, gt_cap=8, warmup=2)
    g2 = GraphedTrainStep(m2, c2, make_optimizer(m2, lr=1e-4, capturable=True), b, gt_cap=8, warmup=2, accumulate=2)
    p0 = g2.fopt.flat_p.clone()
    g1.load(b); g1.step()
    g2.load(b); g2.step()
    assert torch.equal(g2.fopt.flat_p, p0) and float(g2.fopt.step_t) == 0.0          # micro-step 1 of 2: nothing applied yet
    g2.load(b); g2.step()
    assert float(g2.fopt.step_t) == 1.0
    torch.cuda.synchronize()
    d = (g1.fopt.flat_p - g2.fopt.flat_p).abs()
    assert d.max().item() <= 2.5e-4 and (d > 1e-5).float().mean().item() <= 0.01, (d.max().item(), (d > 1e-5).float().mean().item())


def test_side_stream_weight_gradients_are_identical(cuda, monkeypatch):
    """The captured step launches weight-gradient GEMMs and LayerNorm folds on a second stream (gemm._SideStream).  Every parameter
    gradient must be bit-identical to the single-stream capture: a replay has no launch gaps to hide a missing dependency
    (an AccumulateGrad clone on the main stream once read out_proj's gradient before the side stream had written it)."""
    from detr_b200.harness import GraphedTrainStep, make_optimizer, synthetic_batch
    m0, c0 = _make(cuda, train=False)
    b = synthetic_batch(2, 160, 200, 11, 6, seed=1)
    grads = {}
    for side in ("0", "1"):
        monkeypatch.setenv("DETR_B200_WGRAD_STREAM", side)
        m, c = copy.deepcopy(m0), copy.deepcopy(c0)
        g = GraphedTrainStep(m, c, make_optimizer(m, lr=1e-4, capturable=True), b, gt_cap=8, warmup=2)
        g.load(b)
        for _ in range(2):
            g.step()
        torch.cuda.synchronize()
        names = {id(p): n for n, p in m.named_parameters()}
        grads[side] = {names[id(p)]: v.detach().clone() for p, v in zip(g.fopt.params, g.fopt.grad_views)}
    bad = [n for n, v in grads["1"].items() if not torch.equal(v, grads["0"][n])]
    assert not bad, bad


def test_model_stays_deepcopyable_after_forward(cuda):
    """The second-stream plumbing (streams, events) lives outside the modules: a model that has run forward / backward can still be
    deep-copied (EMA copies, checkpoint tooling), and the copy computes the same thing."""
    from detr_b200.harness import batch_to, synthetic_batch
    m, _ = _make(cuda, train=False)
    b = batch_to(synthetic_batch(2, 160, 200, 11, 6, seed=5), cuda)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m(b["image"], b["height"], b["width"])
    out["pred_logits"].float().sum().backward()
    m2 = copy.deepcopy(m)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        o1 = m(b["image"], b["height"], b["width"])
        o2 = m2(b["image"], b["height"], b["width"])
    assert torch.equal(o1["pred_logits"], o2["pred_logits"]) and torch.equal(o1["pred_boxes"], o2["pred_boxes"])
