"""GPU parity tests (B200 box): proposal decode, sort, NMS and the fused proposal layer through the C ABI against the
oracle (bit-exact indices) and against the golden vectors written from the reference's own Python."""
import os

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


@pytest.mark.parametrize("seed,n,thr", [(1, 1, 0.7), (2, 37, 0.7), (3, 300, 0.7), (4, 2000, 0.7), (5, 2000, 0.3),
                                        (6, 6000, 0.7), (7, 12000, 0.7)])
def test_nms_keep_list_bit_exact_vs_reference(ops, golden, seed, n, thr):
    dets = synth.nms_dets(seed, n)
    keep = ops.nms_dets(cuda(dets), thr)
    assert keep.dtype == torch.int32
    assert np.array_equal(keep.cpu().numpy(), golden[f"nms_keep_{seed}_{n}_{thr}"])


def test_nms_wrapper_and_gpu_entry_points(ops, orc, golden):
    import i2vsgg_b200
    i2vsgg_b200.install_as_model()
    from model.nms.nms_wrapper import nms
    from model.nms.nms_gpu import nms_gpu
    from model.roi_layers import nms as c_nms
    dets = synth.nms_dets(3, 300)
    assert nms(torch.zeros((0, 5), device="cuda"), 0.7) == []
    assert np.array_equal(nms(cuda(dets), 0.7).cpu().numpy(), golden["nms_keep_3_300_0.7"])
    order = np.argsort(-dets[:, 4], kind="stable")
    k = nms_gpu(cuda(dets[order]), 0.7)
    assert k.shape[1] == 1
    assert np.array_equal(order[k.view(-1).cpu().numpy()], golden["nms_keep_3_300_0.7"])
    k2 = c_nms(cuda(dets[:, :4]), cuda(dets[:, 4]), 0.7)
    assert k2.dtype == torch.int64 and np.array_equal(k2.cpu().numpy(), golden["nms_keep_3_300_0.7"])


def test_nms_sorted_batched_max_keep_and_spill(ops, orc):
    # batch of independent sets, early stop, and more survivors than the shared-memory cache holds
    sets = [synth.nms_dets(40 + i, 3000) for i in range(3)]
    sets = [d[np.argsort(-d[:, 4], kind="stable")] for d in sets]
    boxes = np.stack(sets)
    for max_keep in (0, 1, 77):
        keep, num = ops.nms_sorted(cuda(boxes), 0.7, max_keep)
        for b in range(3):
            want = orc.nms_sorted(sets[b], 0.7, max_keep)
            assert int(num[b]) == want.size
            assert np.array_equal(keep[b, : want.size].cpu().numpy(), want)
    rng = np.random.default_rng(3)       # 6000 disjoint boxes: every one survives -> spill path
    n = 6000
    gx, gy = np.meshgrid(np.arange(100), np.arange(60))
    b0 = np.stack([gx.ravel() * 10, gy.ravel() * 10, gx.ravel() * 10 + 5, gy.ravel() * 10 + 5,
                   np.linspace(1, 0, n)], 1).astype(np.float32)
    keep, num = ops.nms_sorted(cuda(b0), 0.5)
    assert keep.numel() == n and np.array_equal(keep.cpu().numpy(), np.arange(n))


def test_nms_legacy_entry_point(ops, orc):
    import ctypes
    from i2vsgg_b200 import _lib
    lib = _lib.load()
    d = synth.nms_dets(5, 2000)
    d = np.ascontiguousarray(d[np.argsort(-d[:, 4], kind="stable")])
    keep = torch.zeros(2000, dtype=torch.int32, device="cuda")
    num = torch.zeros(1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    lib.nms_cuda_compute(ctypes.c_void_p(keep.data_ptr()), ctypes.c_void_p(num.data_ptr()),
                         ctypes.c_void_p(d.ctypes.data), 2000, 5, 0.7)        # host boxes, as nms_cuda_kernel.cu:99
    want = orc.nms_sorted(d, 0.7)
    assert int(num.item()) == want.size and np.array_equal(keep[: want.size].cpu().numpy(), want)
    dd = cuda(d)
    lib.nms_cuda_compute(ctypes.c_void_p(keep.data_ptr()), ctypes.c_void_p(num.data_ptr()),
                         ctypes.c_void_p(dd.data_ptr()), 2000, 5, 0.7)        # device boxes work too
    assert int(num.item()) == want.size and np.array_equal(keep[: want.size].cpu().numpy(), want)


def test_decode_and_sort_stages(ops, orc, golden):
    cls, reg = synth.rpn_outputs(12, batch=1)
    info = synth.im_info(1)
    boxes, scores, order = ops.proposal_stages(cuda(cls), cuda(reg), cuda(info), cuda(synth.BASE_ANCHORS), 16)
    wb, wsc = orc.proposal_decode(cls, reg, info, synth.BASE_ANCHORS)
    assert np.array_equal(scores.cpu().numpy(), wsc)
    assert np.array_equal(boxes.cpu().numpy(), wb)                      # same roundings as the oracle: bit-exact
    np.testing.assert_allclose(boxes.cpu().numpy(), golden["decode_b1"], rtol=2e-6, atol=2e-4)   # vs torch (exp ulp)
    assert np.array_equal(order.cpu().numpy()[0], np.argsort(-wsc[0], kind="stable"))


@pytest.mark.parametrize("n", [70, 4000, 4096, 5000])      # <= 4096: the top-K order kernel; above: the full radix sort
def test_sort_is_stable_with_ties_and_negatives(ops, n):
    rng = np.random.default_rng(2)
    score = rng.choice(np.array([-2.5, -0.0, 0.0, 0.25, 0.25, 1e-30, 3.0, -1e-30], np.float32), n)
    dets = np.concatenate([rng.uniform(0, 100, (n, 4)).astype(np.float32), score[:, None]], 1)
    # thresh 2.0 > any IoU: nothing is suppressed, the keep list IS the sort order
    keep = ops.nms_dets(cuda(dets), 2.0).cpu().numpy()
    want = np.argsort(-score.astype(np.float64) + 0.0, kind="stable")
    assert np.array_equal(score[keep], score[want])
    for v in np.unique(score):           # equal scores keep their original order (lower index first)
        idx = keep[score[keep] == v]
        assert np.all(np.diff(idx) > 0)


@pytest.mark.parametrize("key,seed,batch,pre,post", [("prop_test_b2", 11, 2, 6000, 300),
                                                     ("prop_train_b1", 12, 1, 12000, 2000),
                                                     ("prop_train_target_b1", 12, 1, 12000, 128)])
def test_proposal_layer_vs_reference_and_oracle(ops, orc, golden, key, seed, batch, pre, post):
    cls, reg = synth.rpn_outputs(seed, batch=batch)
    info = synth.im_info(batch)
    out, counts = ops.proposal_forward(cuda(cls), cuda(reg), cuda(info), cuda(synth.BASE_ANCHORS), 16, pre, post, 0.7,
                                       return_counts=True)
    want = orc.proposal_layer(cls, reg, info, pre, post, 0.7)
    assert np.array_equal(out.cpu().numpy(), want)                      # bit-exact vs the oracle
    np.testing.assert_allclose(out.cpu().numpy(), golden[key], rtol=2e-6, atol=2e-4)   # vs the reference's Python
    assert np.array_equal(counts.cpu().numpy(), (want[:, :, 1:].any(-1)).sum(1))


def test_proposal_layer_module(ops, orc):
    import i2vsgg_b200
    i2vsgg_b200.install_as_model()
    from model.rpn.proposal_layer import _ProposalLayer
    from model.utils.config import cfg
    cls, reg = synth.rpn_outputs(13, batch=2)
    info = synth.im_info(2)
    layer = _ProposalLayer(cfg.FEAT_STRIDE[0], cfg.ANCHOR_SCALES, cfg.ANCHOR_RATIOS)
    out = layer((cuda(cls), cuda(reg), cuda(info), "TEST"))
    assert out.shape == (2, 300, 5)
    assert np.array_equal(out.cpu().numpy(), orc.proposal_layer(cls, reg, info, 6000, 300, 0.7))
    out = layer((cuda(cls), cuda(reg), cuda(info), "TRAIN"), target=True)
    assert np.array_equal(out.cpu().numpy(), orc.proposal_layer(cls, reg, info, 12000, 128, 0.7))


def _collapse(cls, reg, frame, how_many):
    """The `how_many` best-scored anchors of `frame` regress onto ONE box (so NMS keeps a single one of them)."""
    A = synth.NUM_ANCHORS
    h, w = cls.shape[2:]
    anchors = synth._anchors(h, w).astype(np.float64)
    fg = cls[frame, A:].transpose(1, 2, 0).reshape(-1)
    top = np.argsort(-fg, kind="stable")[:how_many]
    aw, ah = anchors[top, 2] - anchors[top, 0] + 1, anchors[top, 3] - anchors[top, 1] + 1
    acx, acy = anchors[top, 0] + 0.5 * aw, anchors[top, 1] + 0.5 * ah
    d = reg[frame].transpose(1, 2, 0).reshape(-1, 4).copy()
    d[top] = np.stack([(500.0 - acx) / aw, (300.0 - acy) / ah, np.log(200.0 / aw), np.log(150.0 / ah)], 1)
    reg[frame] = d.reshape(h, w, A * 4).transpose(2, 0, 1)


def test_proposal_short_order_runs_out_and_the_frame_is_redone(ops, orc):
    # Frame 1: the 5000 best anchors collapse onto one box, so the 4096 ordered first yield a single kept box and the
    # frame must be redone from the full order; frame 2: every anchor collapses (fewer than 300 boxes exist at all);
    # frames 0 and 3 take the short path.  All four must equal the oracle, which always sorts everything.
    cls, reg = synth.rpn_outputs(31, batch=4)
    info = synth.im_info(4)
    _collapse(cls, reg, 1, 5000)
    _collapse(cls, reg, 2, cls.shape[2] * cls.shape[3] * synth.NUM_ANCHORS)
    for pre, post in ((12000, 300), (6000, 300), (6000, 50)):
        out, counts = ops.proposal_forward(cuda(cls), cuda(reg), cuda(info), cuda(synth.BASE_ANCHORS), 16, pre, post, 0.7,
                                           return_counts=True)
        want = orc.proposal_layer(cls, reg, info, pre, post, 0.7)
        assert np.array_equal(out.cpu().numpy(), want)
        c = counts.cpu().numpy()
        assert c[0] == post and c[3] == post and c[2] < 10


def test_proposal_short_order_with_ties_across_its_boundary(ops, orc):
    # scores quantised to 32 levels: some 670 anchors share each value and one such run straddles rank 4096, where the
    # selection must take exactly the lower-indexed ones (stable order of proposal_layer.py:127)
    cls, reg = synth.rpn_outputs(32, batch=2)
    info = synth.im_info(2)
    A = synth.NUM_ANCHORS
    cls[:, A:] = np.round(cls[:, A:] * 32) / 32
    cls[:, :A] = 1.0 - cls[:, A:]
    cls[1, A:] = 0.5                                                   # one frame with ALL scores equal
    cls[1, :A] = 0.5
    for pre, post in ((12000, 300), (3000, 300)):
        out = ops.proposal_forward(cuda(cls), cuda(reg), cuda(info), cuda(synth.BASE_ANCHORS), 16, pre, post, 0.7)
        assert np.array_equal(out.cpu().numpy(), orc.proposal_layer(cls, reg, info, pre, post, 0.7))


def test_proposal_small_map_and_padding(ops, orc):
    # 5x7 map: 315 anchors < pre_nms_topN, fewer survivors than post_nms_topN -> zero-padded rows keep the frame id
    cls, reg = synth.rpn_outputs(14, batch=3, h=5, w=7, clusters=3)
    info = synth.im_info(3)
    out = ops.proposal_forward(cuda(cls), cuda(reg), cuda(info), cuda(synth.BASE_ANCHORS), 16, 6000, 300, 0.7)
    want = orc.proposal_layer(cls, reg, info, 6000, 300, 0.7)
    assert np.array_equal(out.cpu().numpy(), want)
    assert (out[:, -1, 1:] == 0).all() and (out[:, -1, 0].cpu() == torch.arange(3.0)).all()


def test_nms_properties_at_full_size(ops):
    # size-independent properties at BASELINE config 2 size: kept boxes are pairwise below the threshold, every
    # dropped box overlaps an earlier kept one, and running NMS on the survivors keeps them all (idempotence)
    d = synth.nms_dets(77, 12000)
    d = d[np.argsort(-d[:, 4], kind="stable")]
    keep, _ = ops.nms_sorted(cuda(d), 0.7)
    kept = d[keep.cpu().numpy()]
    again, _ = ops.nms_sorted(cuda(kept), 0.7)
    assert again.numel() == kept.shape[0]
    k = torch.from_numpy(kept[:, :4]).cuda()
    area = (k[:, 2] - k[:, 0] + 1) * (k[:, 3] - k[:, 1] + 1)
    iw = (torch.minimum(k[:, None, 2], k[None, :, 2]) - torch.maximum(k[:, None, 0], k[None, :, 0]) + 1).clamp(min=0)
    ih = (torch.minimum(k[:, None, 3], k[None, :, 3]) - torch.maximum(k[:, None, 1], k[None, :, 1]) + 1).clamp(min=0)
    iou = iw * ih / (area[:, None] + area[None, :] - iw * ih)
    iou.fill_diagonal_(0)
    assert float(iou.max()) <= 0.7 + 1e-6


def test_host_pipeline_chunked_copy_equals_device_step():
    """The end-to-end call (pinned host buffers in and out, frames streamed in chunks over both copy engines) returns
    exactly what the device-resident step computes."""
    from i2vsgg_b200.pipeline import HostPipeline
    frames, ch = 6, 32
    dev = torch.device("cuda", 0)
    cls, reg = synth.rpn_outputs(77, batch=frames)
    info = synth.im_info(frames)
    g = torch.Generator().manual_seed(3)
    feat = torch.randn((frames, ch, 38, 63), generator=g).pin_memory()
    grad = torch.randn((frames * 50, ch, 7, 7), generator=g).pin_memory()
    cls, reg, info = (torch.from_numpy(a).pin_memory() for a in (cls, reg, info))
    pipe = HostPipeline(dev, frames, ch, 38, 63, 7, 1 / 16, 3000, 50, 0.7)
    want = [t.clone() for t in pipe.device_step(*(t.to(dev) for t in (cls, reg, info, feat, grad)))]
    for chunk in (4, 1, 6):
        got = pipe.host_step(cls, reg, info, feat, grad, chunk_frames=chunk)
        for a, b in zip(got, want):
            assert torch.equal(a, b.cpu())


def test_step_backward_on_the_forward_tables_equals_the_plain_call(ops):
    """`HostPipeline` hands the backward of a step rois = NULL (the forward call left this batch's tables and lists in the
    workspace): same gradient, bit for bit, as the stand-alone call that builds its own; and the library refuses NULL when
    the workspace holds no such tables."""
    import ctypes
    from i2vsgg_b200 import _lib
    from i2vsgg_b200.pipeline import HostPipeline
    frames, ch = 3, 32
    dev = torch.device("cuda", 0)
    pipe = HostPipeline(dev, frames, ch, 38, 63, 7, 1 / 16, 3000, 50, 0.7)
    assert pipe._share_tables
    cls, reg = synth.rpn_outputs(61, batch=frames)
    g = torch.Generator().manual_seed(8)
    feat = torch.randn((frames, ch, 38, 63), generator=g).to(dev)
    grad = torch.randn((frames * 50, ch, 7, 7), generator=g).to(dev)
    rois, pooled, grad_in = pipe.device_step(cuda(cls), cuda(reg), cuda(synth.im_info(frames)), feat, grad)
    torch.cuda.synchronize()
    want = ops.roi_align_backward(grad, None, rois.reshape(-1, 5), (frames, ch, 38, 63), 7, 7, 1 / 16, "avg")
    assert torch.equal(grad_in, want)
    lib = _lib.load()
    nb = lib.i2v_roi_align_workspace_bytes(frames, frames * 50)
    fresh = torch.zeros(2 * nb + 512, dtype=torch.uint8, device=dev)[nb + 256:]     # an address no call has prepared
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.i2v_roi_align_backward(p(grad), None, None, p(grad_in), frames, ch, 38, 63, frames * 50, 7, 7, 1 / 16,
                                    _lib.POOL_AVG, _lib.IMPL_AUTO, p(fresh), fresh.numel(),
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc != 0


def test_host_pipeline_two_steps_in_flight():
    """`host_step_async` with step i+1 issued before step i is awaited (alternating result slots, shared device buffers)
    returns, for a sequence of DIFFERENT batches, exactly what the device-resident step computes for each."""
    from i2vsgg_b200.pipeline import HostPipeline
    frames, ch = 6, 32
    dev = torch.device("cuda", 0)
    pipe = HostPipeline(dev, frames, ch, 38, 63, 7, 1 / 16, 3000, 50, 0.7)
    g = torch.Generator().manual_seed(9)
    batches = []
    for k in range(4):
        cls, reg = synth.rpn_outputs(90 + k, batch=frames)
        feat = torch.randn((frames, ch, 38, 63), generator=g)
        grad = torch.randn((frames * 50, ch, 7, 7), generator=g)
        batches.append(tuple(t.pin_memory() for t in (torch.from_numpy(cls), torch.from_numpy(reg),
                                                      torch.from_numpy(synth.im_info(frames)), feat, grad)))
    want = [[t.cpu().clone() for t in pipe.device_step(*(t.to(dev) for t in b))] for b in batches]
    prev, checked = None, 0
    for k, b in enumerate(batches + [None]):
        cur = pipe.host_step_async(*b, chunk_frames=2, slot=k & 1) if b is not None else None
        if prev is not None:
            prev[1][3].synchronize()
            for a, w in zip(prev[1][:3], want[prev[0]]):
                assert torch.equal(a, w)
            checked += 1
        prev = (k, cur) if cur is not None else None
    assert checked == len(batches)


def test_pipelined_step_equals_device_step():
    """The software-pipelined step (proposal of the next batch on a second stream) returns what the plain step returns,
    for a sequence of DIFFERENT batches."""
    from i2vsgg_b200.pipeline import HostPipeline
    frames, ch = 4, 32
    dev = torch.device("cuda", 0)
    pipe = HostPipeline(dev, frames, ch, 38, 63, 7, 1 / 16, 3000, 50, 0.7)
    g = torch.Generator().manual_seed(5)
    feat = torch.randn((frames, ch, 38, 63), generator=g).to(dev)
    grad = torch.randn((frames * 50, ch, 7, 7), generator=g).to(dev)
    batches = []
    for k in range(3):
        cls, reg = synth.rpn_outputs(80 + k, batch=frames)
        batches.append(tuple(torch.from_numpy(a).to(dev) for a in (cls, reg, synth.im_info(frames))))
    want = [[t.clone() for t in pipe.device_step(*b, feat, grad)] for b in batches]
    for k, b in enumerate(batches):
        nxt = batches[k + 1] if k + 1 < len(batches) else None
        got = pipe.pipelined_step(*b, feat, grad, nxt)
        torch.cuda.synchronize()
        for a, w in zip(got, want[k]):
            assert torch.equal(a, w)


def test_rpn_cls_prob_matches_oracle_bit_for_bit(ops, orc):
    """rpn.py:66-68 on the device: the oracle performs the same rounded operations, so the bits agree; the executed
    reference (torch softmax) is within a few ulp (golden fixture)."""
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rpn_golden.npz"))
    for name in ("small", "wide", "a3"):
        score = gold[f"{name}_score"]
        got = ops.rpn_cls_prob(cuda(score)).cpu().numpy()
        assert np.array_equal(got, orc.rpn_cls_prob(score))
        np.testing.assert_allclose(got, gold[f"{name}_prob"], rtol=5e-7, atol=1e-38)
    rng = np.random.default_rng(4)
    big = (rng.standard_normal((4, 18, 38, 63)) * 4).astype(np.float32)
    assert np.array_equal(ops.rpn_cls_prob(cuda(big)).cpu().numpy(), orc.rpn_cls_prob(big))


def test_proposals_from_raw_scores_equal_the_two_step_path(ops, orc):
    """Softmax fused into the decode kernel == rpn_cls_prob followed by the proposal layer, and both equal the oracle's
    proposal layer run on the oracle's probabilities."""
    rng = np.random.default_rng(9)
    B = 3
    score = (rng.standard_normal((B, 18, 38, 63)) * 2.5).astype(np.float32)
    _, reg = synth.rpn_outputs(31, batch=B)
    info, anchors = synth.im_info(B), synth.BASE_ANCHORS
    fused = ops.proposal_forward(cuda(score), cuda(reg), cuda(info), cuda(anchors), 16, 6000, 300, 0.7, from_scores=True)
    prob = ops.rpn_cls_prob(cuda(score))
    two_step = ops.proposal_forward(prob, cuda(reg), cuda(info), cuda(anchors), 16, 6000, 300, 0.7)
    assert torch.equal(fused, two_step)
    want = orc.proposal_layer(orc.rpn_cls_prob(score), reg, info, 6000, 300, 0.7)
    assert np.array_equal(fused.cpu().numpy(), want)


def test_rpn_module_mirror(ops):
    from i2vsgg_b200.model.rpn.rpn import _RPN
    rng = np.random.default_rng(10)
    score = cuda((rng.standard_normal((2, 18, 38, 63)) * 2).astype(np.float32))
    _, reg = synth.rpn_outputs(32, batch=2)
    rpn = _RPN(1024).eval()
    a = rpn.forward_head_outputs(score, cuda(reg), cuda(synth.im_info(2)))
    b = rpn.forward_head_outputs(score, cuda(reg), cuda(synth.im_info(2)), fused=False)
    assert a.shape == (2, 300, 5) and torch.equal(a, b)
    p = _RPN.cls_prob(score)
    ref = torch.softmax(_RPN.reshape(score, 2), 1).view_as(score)
    assert float((p - ref).abs().max()) < 1e-6
