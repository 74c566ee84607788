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
