"""CPU tests: the oracle against the reference's golden vectors (tests/golden, written by make_golden.py from the
reference's own code), against the reference's C sources compiled unmodified (oracle/_ref, when built) and against
torchvision for the ops whose source the reference does not ship."""
import os

import numpy as np
import pytest

from i2vsgg_b200 import synth
from oracle import oracle, ref

SCALE = 1.0 / 16
HERE = os.path.dirname(os.path.abspath(__file__))


def test_anchor_table(golden):
    # generate_anchors.py:12-37 (MATLAB values) minus 1
    table = np.array([[-83, -39, 100, 56], [-175, -87, 192, 104], [-359, -183, 376, 200], [-55, -55, 72, 72],
                      [-119, -119, 136, 136], [-247, -247, 264, 264], [-35, -79, 52, 96], [-79, -167, 96, 184],
                      [-167, -343, 184, 360]], np.float64) - 1
    assert np.array_equal(golden["anchors"], table)
    assert np.array_equal(oracle.generate_anchors(), table)
    assert np.array_equal(synth.BASE_ANCHORS, table.astype(np.float32))
    from i2vsgg_b200.model.rpn.generate_anchors import generate_anchors
    assert np.array_equal(generate_anchors(scales=np.array([8, 16, 32]), ratios=np.array([0.5, 1, 2])), table)


@pytest.mark.parametrize("seed,n,thr", [(1, 1, 0.7), (2, 37, 0.7), (3, 300, 0.7), (4, 2000, 0.7), (5, 2000, 0.3),
                                        (6, 6000, 0.7), (7, 12000, 0.7)])
def test_nms_matches_reference_keep_list(golden, seed, n, thr):
    dets = synth.nms_dets(seed, n)
    assert np.array_equal(oracle.nms(dets, thr), golden[f"nms_keep_{seed}_{n}_{thr}"])


def test_nms_empty_and_max_keep():
    assert oracle.nms(np.zeros((0, 5), np.float32), 0.7).size == 0
    dets = synth.nms_dets(3, 300)
    full = oracle.nms(dets, 0.7)
    order = np.argsort(-dets[:, 4], kind="stable")
    assert np.array_equal(order[oracle.nms_sorted(dets[order], 0.7, 10)], full[:10])


def test_decode_matches_reference(golden):
    cls, reg = synth.rpn_outputs(12, batch=1)
    boxes, scores = oracle.proposal_decode(cls, reg, synth.im_info(1), synth.BASE_ANCHORS)
    # torch's CPU exp (SLEEF, 1 ulp) vs the correctly rounded exp of the oracle: last-bit differences in w/h only
    np.testing.assert_allclose(boxes, golden["decode_b1"], rtol=2e-6, atol=2e-4)
    assert (boxes == golden["decode_b1"]).mean() > 0.9
    fg = cls[:, synth.NUM_ANCHORS:].transpose(0, 2, 3, 1).reshape(1, -1)
    assert np.array_equal(scores, fg)


@pytest.mark.parametrize("key,seed,batch,cfg", [("prop_test_b2", 11, 2, (6000, 300)), ("prop_train_b1", 12, 1, (12000, 2000)),
                                                ("prop_train_target_b1", 12, 1, (12000, 128))])
def test_proposal_layer_matches_reference(golden, key, seed, batch, cfg):
    cls, reg = synth.rpn_outputs(seed, batch=batch)
    out = oracle.proposal_layer(cls, reg, synth.im_info(batch), cfg[0], cfg[1], 0.7)
    want = golden[key]
    assert out.shape == want.shape
    assert np.array_equal(out[:, :, 0], want[:, :, 0])
    # same boxes kept in the same order; coordinates up to the exp() last-bit difference
    np.testing.assert_allclose(out, want, rtol=2e-6, atol=2e-4)


def test_roi_align_forward_bit_exact_vs_reference_c(golden):
    feat = synth.feature_map(21, batch=2, channels=8)
    rois = synth.rois(22, 40, batch=2)
    assert np.array_equal(oracle.roi_align_forward(feat, rois, 8, 8, SCALE), golden["roi_align_ref_8x8"])
    assert np.array_equal(oracle.roi_align_forward(feat, rois, 7, 7, SCALE), golden["roi_align_ref_7x7"])


def test_roi_pool_forward_vs_reference_c(golden):
    feat = synth.feature_map(23, batch=1, channels=8)
    rois = synth.rois(24, 30, batch=1, degenerate=0)
    out, arg = oracle.roi_pool_forward(feat, rois, 7, 7, SCALE)
    # roi_pooling.c:29 seeds the running maximum with -1 (its comment says -inf), the CUDA kernel the reference
    # really runs seeds it with -FLT_MAX (roi_pooling_kernel.cu:68): the CPU file equals max(kernel result, -1)
    assert np.array_equal(np.where(arg >= 0, np.maximum(out, -1.0), out), golden["roi_pool_ref_7x7"])
    assert (out < -1).any()
    flat = feat.ravel()
    assert np.array_equal(np.where(arg >= 0, flat[np.maximum(arg, 0)], 0.0), out)


@pytest.mark.skipif(not ref.have_cpu_ref(), reason="oracle/_ref/libref_cpu.so not built")
def test_roi_align_forward_vs_live_reference_c():
    feat = synth.feature_map(31, batch=3, channels=5, h=20, w=33)
    rois = synth.rois(32, 64, batch=3)
    rois[:, 1:] *= 0.5
    for g in (2, 7, 8):
        assert np.array_equal(oracle.roi_align_forward(feat, rois, g, g, SCALE), ref.cpu_roi_align_forward(feat, rois, g, g, SCALE))


def test_roi_align_backward_is_adjoint_of_forward():
    # <forward(f), g> == <f, backward(g)>: pins the backward restatement to the (reference-pinned) forward
    rng = np.random.default_rng(5)
    feat = synth.feature_map(41, batch=2, channels=3, h=12, w=17)
    rois = synth.rois(42, 25, batch=2)
    rois[:, 1:] *= 0.25
    for mode, p in (("none", 7), ("avg", 7), ("none", 4)):
        g = rng.standard_normal((25, 3, p, p)).astype(np.float32)
        out = oracle.roi_align_pooled_forward(feat, rois, p, p, SCALE, mode)
        gin = oracle.roi_align_pooled_backward(g, feat, rois, p, p, SCALE, mode)
        lhs = float((out.astype(np.float64) * g).sum())
        rhs = float((feat.astype(np.float64) * gin).sum())
        assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))


def test_roi_align_max_backward_routes_to_argmax():
    feat = synth.feature_map(43, batch=1, channels=2, h=12, w=17)
    rois = synth.rois(44, 6, batch=1, degenerate=0)
    rois[:, 1:] *= 0.25
    g = np.ones((6, 2, 7, 7), np.float32)
    gin = oracle.roi_align_pooled_backward(g, feat, rois, 7, 7, SCALE, "max")
    eps = 1e-3
    # directional derivative check on a random direction
    rng = np.random.default_rng(1)
    d = rng.standard_normal(feat.shape).astype(np.float32)
    f0 = oracle.roi_align_pooled_forward(feat, rois, 7, 7, SCALE, "max").astype(np.float64).sum()
    f1 = oracle.roi_align_pooled_forward(feat + eps * d, rois, 7, 7, SCALE, "max").astype(np.float64).sum()
    assert abs((f1 - f0) / eps - float((gin.astype(np.float64) * d).sum())) < 5e-2 * abs((f1 - f0) / eps) + 1e-2


def test_roi_pool_backward_matches_scatter_for_regular_rois():
    rng = np.random.default_rng(7)
    feat = synth.feature_map(45, batch=2, channels=3, h=12, w=17)
    rois = synth.rois(46, 20, batch=2, degenerate=0, edge_frac=0.0)
    rois[:, 1:] *= 0.25
    out, arg = oracle.roi_pool_forward(feat, rois, 7, 7, SCALE)
    g = rng.standard_normal(out.shape).astype(np.float32)
    gin = oracle.roi_pool_backward(g, rois, arg, feat.shape, 7, 7, SCALE)
    want = np.zeros(feat.size, np.float64)
    np.add.at(want, arg[arg >= 0], g[arg >= 0].astype(np.float64))
    np.testing.assert_allclose(gin.ravel(), want, rtol=1e-5, atol=1e-5)


def test_c_ops_match_torchvision():
    tv = pytest.importorskip("torchvision.ops")
    import torch
    feat = synth.feature_map(51, batch=2, channels=4, h=20, w=30)
    rois = synth.rois(52, 30, batch=2, degenerate=0)
    rois[:, 1:] *= 0.5
    tf, tr = torch.from_numpy(feat).requires_grad_(True), torch.from_numpy(rois)
    for sr in (0, 2):
        want = tv.roi_align(tf, tr, (7, 7), SCALE, sr, aligned=False)
        got = oracle.c_roi_align_forward(feat, rois, 7, 7, SCALE, sr)
        np.testing.assert_allclose(got, want.detach().numpy(), rtol=1e-5, atol=1e-5)
        g = torch.from_numpy(np.random.default_rng(3).standard_normal(got.shape).astype(np.float32))
        (gi,) = torch.autograd.grad(want, tf, g)
        np.testing.assert_allclose(oracle.c_roi_align_backward(g.numpy(), rois, feat.shape, 7, 7, SCALE, sr), gi.numpy(),
                                   rtol=1e-4, atol=1e-5)
    want = tv.roi_pool(tf, tr, (7, 7), SCALE)
    got, arg = oracle.c_roi_pool_forward(feat, rois, 7, 7, SCALE)
    np.testing.assert_allclose(got, want.detach().numpy(), rtol=0, atol=0)
    g = torch.from_numpy(np.random.default_rng(4).standard_normal(got.shape).astype(np.float32))
    (gi,) = torch.autograd.grad(want, tf, g)
    np.testing.assert_allclose(oracle.c_roi_pool_backward(g.numpy(), rois, arg, feat.shape, 7, 7), gi.numpy(), rtol=1e-5, atol=1e-5)


def test_pair_stage_matches_python_loops():
    boxes, classes, conf = synth.detections(61, 9)
    ixs, ixo = oracle.enumerate_pairs(9)
    want_s, want_o = [], []
    for i in range(9):                      # faster_rcnn_SGG_emb.py:597-606
        for j in range(9):
            if i != j:
                want_s.append(i)
                want_o.append(j)
    assert np.array_equal(ixs, want_s) and np.array_equal(ixo, want_o)
    rel = oracle.union_boxes(boxes, ixs, ixo, synth.IM_H, synth.IM_W)
    for p in range(len(ixs)):
        a, b = boxes[ixs[p]], boxes[ixo[p]]
        u = [max(0, min(a[0], b[0]) - 10), max(0, min(a[1], b[1]) - 10), min(synth.IM_W, max(a[2], b[2]) + 10),
             min(synth.IM_H, max(a[3], b[3]) + 10)]          # resnet_SGG_emb.py:240-244
        assert np.array_equal(rel[p, 1:], np.array(u, np.float32)) and rel[p, 0] == 0
    masks = oracle.dual_masks(boxes, ixs, ixo, synth.IM_H, synth.IM_W)
    x1, x2, y1, y2 = oracle.dual_mask_extent(boxes[0], synth.IM_H, synth.IM_W)
    assert masks[0, 0].sum() == (y2 - y1) * (x2 - x1)        # resnet_SGG_emb.py:255


@pytest.mark.needs_reference
def test_live_reference_python_agrees_with_golden(golden):
    dets = synth.nms_dets(3, 300)
    assert np.array_equal(ref.py_nms_cpu(dets, 0.7), golden["nms_keep_3_300_0.7"])


def test_roi_crop_forward_bit_exact_vs_reference_c():
    """oracle.roi_crop_forward against the reference's roi_crop.c executed (tests/golden/make_roi_crop_golden.py): one RoI
    per image, where the CPU file and the CUDA kernel the oracle follows describe the same sampling."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_roi_crop_golden", os.path.join(HERE, "golden", "make_roi_crop_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    feat, grids = mod.crop_inputs()
    want = np.load(os.path.join(HERE, "golden", "roi_crop_golden.npz"))["forward"]
    assert np.array_equal(oracle.roi_crop_forward(feat, grids), want)
    if ref.have_cpu_ref():
        live = ref.cpu_roi_crop_forward_bhwd(feat.transpose(0, 2, 3, 1), grids).transpose(0, 3, 1, 2)
        assert np.array_equal(live, want)


def test_roi_crop_backward_is_adjoint_of_forward_and_rois_per_image():
    rng = np.random.default_rng(12)
    B, C, H, W, per = 2, 3, 9, 14, 4
    feat = rng.standard_normal((B, C, H, W)).astype(np.float32)
    grids = rng.uniform(-1.2, 1.2, (B * per, 5, 6, 2)).astype(np.float32)
    g = rng.standard_normal((B * per, C, 5, 6)).astype(np.float32)
    out = oracle.roi_crop_forward(feat, grids)
    gin = oracle.roi_crop_backward(g, grids, feat.shape)
    lhs, rhs = float((out.astype(np.float64) * g).sum()), float((feat.astype(np.float64) * gin).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)
    # RoI n samples frame n // per (roi_crop_cuda_kernel.cu:65): frame 1's RoIs do not see frame 0
    alone = oracle.roi_crop_forward(feat[1:], grids[per:])
    assert np.array_equal(out[per:], alone)
