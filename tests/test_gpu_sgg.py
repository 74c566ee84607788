"""GPU parity tests (B200 box): SGG pair stage and triplet top-k through the C ABI against the oracle (bit-exact)."""
import numpy as np
import pytest
import torch

from i2vsgg_b200 import synth

pytestmark = pytest.mark.gpu


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def ops():
    from i2vsgg_b200 import ops
    return ops


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


@pytest.mark.parametrize("n", [2, 9, 64])
def test_pair_build_bit_exact(ops, orc, n):
    boxes, _, _ = synth.detections(60 + n, n)
    ixs, ixo, rel, masks = ops.pair_build(cuda(boxes), synth.IM_H, synth.IM_W)
    ws, wo = orc.enumerate_pairs(n)
    assert ixs.dtype == torch.int64 and np.array_equal(ixs.cpu().numpy(), ws) and np.array_equal(ixo.cpu().numpy(), wo)
    assert np.array_equal(rel.cpu().numpy(), orc.union_boxes(boxes, ws, wo, synth.IM_H, synth.IM_W))
    assert np.array_equal(masks.cpu().numpy(), orc.dual_masks(boxes, ws, wo, synth.IM_H, synth.IM_W))


def test_pair_build_degenerate_counts(ops):
    for n in (0, 1):
        ixs, ixo, rel, masks = ops.pair_build(torch.zeros((n, 4), device="cuda"), 600, 1000)
        assert ixs.numel() == 0 and rel.shape == (0, 5) and masks.shape == (0, 2, 32, 32)


@pytest.mark.parametrize("n,r,k", [(64, 132, 100), (9, 26, 100), (3, 5, 100), (20, 40, 7)])
def test_triplet_topk_bit_exact(ops, orc, n, r, k):
    rng = np.random.default_rng(n * 1000 + r)
    boxes, classes, conf = synth.detections(70 + n, n)
    ixs, ixo = orc.enumerate_pairs(n)
    p = len(ixs)
    logits = rng.standard_normal((p, r)).astype(np.float32) * 3
    prob = np.exp(logits - logits.max(1, keepdims=True))
    prob = (prob / prob.sum(1, keepdims=True)).astype(np.float32)
    rec, cnt = ops.triplet_topk(cuda(prob), cuda(conf), cuda(classes), cuda(boxes), cuda(ixs), cuda(ixo), k)
    wconf, wlab, wsub, wobj, wpair = orc.detection_output(prob, conf, classes, boxes, ixs, ixo, k)
    m = min(k, p * r)
    assert int(cnt.item()) == m
    rec = rec.cpu().numpy()
    assert np.array_equal(rec[:m, 0], wconf)
    assert np.array_equal(rec[:m, 1:4], wlab)
    assert np.array_equal(rec[:m, 4:8], wsub) and np.array_equal(rec[:m, 8:12], wobj)
    assert np.array_equal(rec[:m, 12].astype(np.int64), wpair)
    assert not rec[m:].any()


def test_triplet_topk_ties_take_lowest_flat_index(ops, orc):
    n, r = 6, 10
    boxes, classes, _ = synth.detections(5, n)
    conf = np.ones(n, np.float32)
    ixs, ixo = orc.enumerate_pairs(n)
    prob = np.full((len(ixs), r), 0.5, np.float32)
    prob[3, 4] = 0.9
    prob[7, 1] = 0.7
    rec, cnt = ops.triplet_topk(cuda(prob), cuda(conf), cuda(classes), cuda(boxes), cuda(ixs), cuda(ixo), 20)
    rec = rec.cpu().numpy()
    flat = rec[:, 12].astype(int) * r + rec[:, 2].astype(int)
    want = [3 * r + 4, 7 * r + 1] + [i for i in range(len(ixs) * r) if i not in (3 * r + 4, 7 * r + 1)][:18]
    assert list(flat) == want


def test_triplet_topk_crowded_threshold_bin_falls_back(ops, orc):
    """More keys in the threshold bin than the candidate list holds (here: 40 000 equal scores below a handful of distinct
    ones): the select kernel falls back to the exact radix select; ties still go to the lowest flat indices."""
    n, r = 21, 100                                  # 420 pairs x 100 predicates = 42 000 scores
    boxes, classes, _ = synth.detections(9, n)
    conf = np.ones(n, np.float32)
    ixs, ixo = orc.enumerate_pairs(n)
    prob = np.full((len(ixs), r), 0.25, np.float32)
    best = [(17, 3, 0.9), (200, 50, 0.8), (5, 99, 0.7)]
    for p, q, v in best:
        prob[p, q] = v
    rec, cnt = ops.triplet_topk(cuda(prob), cuda(conf), cuda(classes), cuda(boxes), cuda(ixs), cuda(ixo), 100)
    rec = rec.cpu().numpy()
    assert int(cnt.item()) == 100
    flat = rec[:, 12].astype(int) * r + rec[:, 2].astype(int)
    top = [p * r + q for p, q, _ in best]
    want = top + [i for i in range(len(ixs) * r) if i not in top][:97]
    assert list(flat) == want
    wconf, wlab, wsub, wobj, wpair = orc.detection_output(prob, conf, classes, boxes, ixs, ixo, 100)
    assert np.array_equal(rec[:, 0], wconf) and np.array_equal(rec[:, 12].astype(np.int64), wpair)


def test_triplet_topk_config3_size_is_reproducible(ops, orc):
    """4032 pairs x 132 predicates: the candidate list is filled in arbitrary order by many CTAs; the records must not
    depend on it (and equal the oracle's)."""
    n, r = 64, 132
    rng = np.random.default_rng(12)
    boxes, classes, conf = synth.detections(71, n)
    ixs, ixo = orc.enumerate_pairs(n)
    logits = rng.standard_normal((len(ixs), r)).astype(np.float32) * 3
    prob = np.exp(logits - logits.max(1, keepdims=True))
    prob = (prob / prob.sum(1, keepdims=True)).astype(np.float32)
    args = (cuda(prob), cuda(conf), cuda(classes), cuda(boxes), cuda(ixs), cuda(ixo), 100)
    first, _ = ops.triplet_topk(*args)
    for _ in range(3):
        again, _ = ops.triplet_topk(*args)
        assert torch.equal(first, again)
    wconf, wlab, wsub, wobj, wpair = orc.detection_output(prob, conf, classes, boxes, ixs, ixo, 100)
    rec = first.cpu().numpy()
    assert np.array_equal(rec[:, 0], wconf) and np.array_equal(rec[:, 1:4], wlab)
    assert np.array_equal(rec[:, 12].astype(np.int64), wpair)


def test_frame_group_launches_equal_the_per_frame_ones():
    """`pair_build_frames` / `triplet_topk_frames` (one launch chain per frame GROUP) against the per-frame entry points, bit
    for bit: pair lists in group rows, union boxes with the frame number in column 0, one mask per object (= the subject
    channel of that object's first pair), and the top-100 records of every frame -- ties included."""
    import torch
    from i2vsgg_b200 import ops, synth
    F, N, R = 5, 11, 132
    boxes, classes, conf = synth.clip_detections(3, F, N)
    b = torch.from_numpy(boxes).cuda()
    P = N * (N - 1)
    ixs, ixo, rel, om = ops.pair_build_frames(b, synth.IM_H, synth.IM_W)
    assert ixs.shape == (F * P,) and rel.shape == (F * P, 5) and om.shape == (F * N, 32, 32)
    first = torch.arange(N, device="cuda") * (N - 1)
    for f in range(F):
        i1, o1, r1, m1 = ops.pair_build(b[f], synth.IM_H, synth.IM_W)
        assert torch.equal(ixs[f * P:(f + 1) * P], i1 + f * N) and torch.equal(ixo[f * P:(f + 1) * P], o1 + f * N)
        r1[:, 0] = f
        assert torch.equal(rel[f * P:(f + 1) * P], r1)
        assert torch.equal(om[f * N:(f + 1) * N], m1[first, 0])
    g = torch.Generator(device="cuda").manual_seed(9)
    scores = torch.rand((F * P, R), device="cuda", generator=g)
    scores[:, ::7] = 0.5                               # plenty of exact ties
    c = torch.from_numpy(np.tile(classes, (F, 1))).cuda()
    s = torch.from_numpy(np.tile(conf, (F, 1))).cuda() * torch.linspace(0.5, 1.0, F, device="cuda")[:, None]
    rec, cnt = ops.triplet_topk_frames(scores, s, c, b, ixs[:P], ixo[:P], 100)
    for f in range(F):
        want, wc = ops.triplet_topk(scores[f * P:(f + 1) * P], s[f], c[f], b[f], ixs[:P], ixo[:P], 100)
        assert torch.equal(rec[f], want) and int(cnt[f]) == int(wc[0])
