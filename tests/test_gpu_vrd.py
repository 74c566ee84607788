"""GPU parity tests (B200 box): the relation head `vrd.forward` on the tcgen05 path.

Two references:
* the numpy fp32 oracle (pinned to the unmodified reference module by tests/test_oracle_vrd.py) and the reference's own
  outputs in tests/golden/vrd_golden.npz.  The tensor cores read bf16 operands (fp32 accumulation), so the bar here is
  the operand quantisation: features within 2e-2 of their scale, softmax scores within 2e-2 relative;
* the same graph evaluated by torch in float64 on operands rounded to bf16 at exactly the points where the kernels round
  (weights, pooled rows, every bf16 activation): what is left is accumulation order, bar 2e-3 of scale.
"""
import importlib.util
import os

import numpy as np
import pytest
import torch

from i2vsgg_b200 import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _gen():
    spec = importlib.util.spec_from_file_location("make_vrd_golden", os.path.join(HERE, "golden", "make_vrd_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def build(args, params, prd):
    from i2vsgg_b200.model.faster_rcnn.resnet_SGG_emb import vrd
    net = vrd(args, None, prd)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    return net.cuda().eval().prepare()


def frame_inputs(orc, seed, num_det):
    det, classes, _ = synth.detections(seed, num_det)
    ixs, ixo = orc.enumerate_pairs(num_det)
    boxes = np.concatenate([np.zeros((num_det, 1), np.float32), det], 1)
    rel = orc.union_boxes(det, ixs, ixo, synth.IM_H, synth.IM_W)
    masks = orc.dual_masks(det, ixs, ixo, synth.IM_H, synth.IM_W)
    return boxes, rel, masks, classes, ixs, ixo


def bf(t):
    return t.bfloat16().double()


def simulated(params, prd, fmap, boxes, rel, spatial, ixs, ixo, args):
    """vrd.forward in float64 with bf16 rounding wherever the kernels round."""
    import torchvision
    P = {k: torch.from_numpy(v).cuda() for k, v in params.items()}

    def fc(x, name, relu=True, round_out=True):
        y = x @ bf(P[name + ".weight"]).t() + P[name + ".bias"].double()
        y = torch.relu(y) if relu else y
        return bf(y.float()) if round_out else y.float().double()

    fm = torch.from_numpy(fmap).cuda()
    pool = lambda b: bf(torchvision.ops.roi_pool(fm, torch.from_numpy(b).cuda(), 7, 1 / 16).reshape(len(b), -1))
    h_o = fc(fc(pool(boxes), "fc6.fc"), "fc7.fc")
    obj = fc(h_o, "so_vis_embeddings.fc", relu=False, round_out=False)
    x = fc(fc(fc(pool(rel), "fc6.fc"), "fc7.fc"), "fc8.fc")
    parts = [x]
    i1, i2 = torch.from_numpy(ixs).cuda(), torch.from_numpy(ixo).cuda()
    if args.use_obj_visual:
        parts.append(fc(bf(torch.cat([obj[i1], obj[i2]], 1).float()), "fc_so.fc"))
    sp = torch.from_numpy(spatial).cuda()
    if args.spatial_type == 1:
        parts.append(fc(bf(sp.reshape(len(rel), 8)), "fc_lov.fc"))
    elif args.spatial_type == 2:
        lo = sp.reshape(-1, 2, 32, 32).double()
        for i, (s, p) in enumerate(((2, 2), (2, 2), (1, 0))):
            w, b = bf(P[f"conv_lo.{i}.conv.weight"]), P[f"conv_lo.{i}.conv.bias"].double()
            lo = bf(torch.relu(torch.nn.functional.conv2d(lo, w, b, stride=s, padding=p)).float())
        parts.append(fc(lo.reshape(len(rel), -1), "fc_lov.fc"))
    x = fc(fc(torch.cat(parts, 1), "fc_fusion.fc"), "fc_rel.fc", relu=False, round_out=False)
    prdv = torch.from_numpy(prd).cuda()
    e = torch.nn.functional.leaky_relu(prdv @ P["prd_sem_embeddings.0.weight"].t() + P["prd_sem_embeddings.0.bias"], 0.1)
    e = (e @ P["prd_sem_embeddings.2.weight"].t() + P["prd_sem_embeddings.2.bias"]).double()
    sim = torch.nn.functional.normalize(x, dim=1) @ torch.nn.functional.normalize(e, dim=1).t()
    return torch.softmax(sim, 1), x


@pytest.mark.parametrize("kw", [dict(), dict(use_obj_visual=False, spatial_type=1), dict(spatial_type=0)])
def test_vrd_small_config_against_oracle_and_bf16_simulation(orc, kw):
    args = synth.VrdArgs(vrd_in_channels=64, vrd_hidden=512, **kw)
    params = synth.vrd_params(99, args)
    prd = synth.prd_vectors(5, args.num_relations)
    fmap = synth.feature_map(31, 1, 64)
    boxes, rel, masks, classes, ixs, ixo = frame_inputs(orc, 32, 12)
    spatial = masks if args.spatial_type != 1 else np.random.default_rng(3).standard_normal((len(ixs), 8), dtype=np.float32)
    net = build(args, params, prd)
    scores, feat = net(fmap, boxes, rel, spatial, classes, ixs, ixo)
    assert isinstance(feat, np.ndarray) and feat.shape == (132, 300) and scores.shape == (132, 132) and scores.is_cuda
    want_s, want_f = orc.vrd_forward(params, prd, fmap, boxes, rel, spatial, ixs, ixo, args.use_obj_visual,
                                     args.spatial_type, nthreads=orc.default_threads())
    scale = float(np.abs(want_f).max())
    sim_s, sim_f = simulated(params, prd, fmap, boxes, rel, spatial, ixs, ixo, args)
    e_orc = float(np.abs(feat - want_f).max()) / scale
    e_sim = float((torch.from_numpy(feat).cuda().double() - sim_f).abs().max()) / scale
    s_orc = float(np.abs(scores.cpu().numpy() / want_s - 1).max())
    s_sim = float((scores.double() / sim_s - 1).abs().max())
    print(f"vrd small {kw}: feat err vs oracle {e_orc:.2e}, vs bf16 simulation {e_sim:.2e}; "
          f"score rel err vs oracle {s_orc:.2e}, vs simulation {s_sim:.2e}")
    assert e_orc <= 2e-2 and s_orc <= 2e-2
    # a bf16 activation that sits on a rounding boundary may round the other way when the accumulation order differs,
    # which is why this is not at fp32 level
    assert e_sim <= 5e-3 and s_sim <= 5e-3


def test_vrd_reference_golden_full_width(orc):
    """The 12-pair frame the unmodified reference module was run on (1024 channels, 4096 hidden units)."""
    gen = _gen()
    g = np.load(os.path.join(HERE, "golden", "vrd_golden.npz"))
    args = synth.VrdArgs()
    params = synth.vrd_params(gen.PARAM_SEED, args)
    prd = synth.prd_vectors(gen.PRD_SEED, args.num_relations)
    fmap, boxes, rel, masks, classes, ixs, ixo = gen.inputs()
    net = build(args, params, prd)
    scores, feat = net(fmap, boxes, rel, masks, classes, ixs, ixo)
    want_s, want_f = g["full_scores"], g["full_feat"]
    assert float(np.abs(feat - want_f).max()) <= 2e-2 * float(np.abs(want_f).max())
    np.testing.assert_allclose(scores.cpu().numpy(), want_s, rtol=2e-2, atol=1e-6)


def test_vrd_config3_spot_rows(orc):
    """Config 3 at full size (64 detections -> 4032 pairs): 24 sampled pair rows against the oracle."""
    gen = _gen()
    args = synth.VrdArgs()
    params = synth.vrd_params(gen.PARAM_SEED, args)
    prd = synth.prd_vectors(gen.PRD_SEED, args.num_relations)
    fmap = synth.feature_map(41, 1)
    boxes, rel, masks, classes, ixs, ixo = frame_inputs(orc, 42, 64)
    net = build(args, params, prd)
    scores, feat = net(fmap, boxes, rel, masks, classes, ixs, ixo, return_numpy=False)
    assert scores.shape == (4032, 132) and feat.shape == (4032, 300)
    assert float((scores.sum(1) - 1).abs().max()) <= 1e-5
    rows = np.random.default_rng(0).choice(4032, 24, replace=False)
    want_s, want_f = orc.vrd_forward(params, prd, fmap, boxes, rel, masks, ixs, ixo, rows=rows,
                                     nthreads=orc.default_threads())
    got_f = feat[torch.from_numpy(rows).cuda()].cpu().numpy()
    got_s = scores[torch.from_numpy(rows).cuda()].cpu().numpy()
    assert float(np.abs(got_f - want_f).max()) <= 2e-2 * float(np.abs(want_f).max())
    np.testing.assert_allclose(got_s, want_s, rtol=2e-2, atol=1e-6)


def test_vrd_unordered_pair_shortcut_is_bit_identical(orc):
    """(i,j) and (j,i) share their union box: pooling / fc6 / fc7 / fc8 once per unordered pair changes no bit."""
    from i2vsgg_b200 import sgg
    args = synth.VrdArgs(vrd_in_channels=64, vrd_hidden=512)
    params = synth.vrd_params(99, args)
    prd = synth.prd_vectors(5, args.num_relations)
    fmap = synth.feature_map(31, 1, 64)
    boxes, rel, masks, classes, ixs, ixo = frame_inputs(orc, 33, 11)
    net = build(args, params, prd)
    s0, f0 = net(fmap, boxes, rel, masks, classes, ixs, ixo, return_numpy=False)
    rep, inverse = sgg.unordered_pairs(11)
    assert torch.equal(torch.from_numpy(rel).cuda()[rep][inverse], torch.from_numpy(rel).cuda())
    s1, f1 = net(fmap, boxes, rel, masks, classes, ixs, ixo, return_numpy=False, rel_unique=(rep, inverse))
    assert torch.equal(s0, s1) and torch.equal(f0, f1)


def test_vrd_object_mask_path_matches_pair_mask_path(orc):
    """conv_lo's first layer from per-object masks (linear in the input channels) against the im2col of every pair."""
    args = synth.VrdArgs(vrd_in_channels=32, vrd_hidden=256)
    params = synth.vrd_params(21, args)
    net = build(args, params, synth.prd_vectors(5, args.num_relations))
    fmap = synth.feature_map(31, 1, 32)
    boxes, rel, masks, classes, ixs, ixo = frame_inputs(orc, 34, 10)
    obj_masks = np.stack([masks[i * 9, 0] for i in range(10)])
    assert all(np.array_equal(masks[p], np.stack([obj_masks[ixs[p]], obj_masks[ixo[p]]])) for p in range(90))
    s0, f0 = net(fmap, boxes, rel, masks, classes, ixs, ixo, return_numpy=False)
    s1, f1 = net(fmap, boxes, rel, None, classes, ixs, ixo, return_numpy=False, obj_masks=obj_masks)
    # same bf16 operands, the two halves of each sum are accumulated separately: one bf16 rounding flip at most per value
    assert float((f0 - f1).abs().max()) <= 5e-3 * float(f0.abs().max())
    assert float((s0 / s1 - 1).abs().max()) <= 5e-3
    # the first layer itself: equal up to fp32 association before the bf16 rounding
    a = net.conv_lo[0](torch.from_numpy(masks).cuda(), "nchw").float()
    b = net.conv_lo[0].forward_pairs(torch.from_numpy(obj_masks).cuda(), torch.from_numpy(ixs).cuda(),
                                     torch.from_numpy(ixo).cuda()).float()
    assert float((a - b).abs().max()) <= 2 ** -7 * float(a.abs().max())


def test_vrd_refuses_cpu_parameters():
    from i2vsgg_b200.model.faster_rcnn.resnet_SGG_emb import vrd
    args = synth.VrdArgs(vrd_in_channels=16, vrd_hidden=64)
    net = vrd(args, None, synth.prd_vectors(1))
    with pytest.raises(RuntimeError):
        net.eval()(np.zeros((1, 16, 38, 63), np.float32), np.zeros((2, 5)), np.zeros((2, 5)), np.zeros((2, 2, 32, 32)),
                   [1, 1], [0, 1], [1, 0])


def test_vrd_training_mode_dropout_and_raw_scores(orc):
    """train(): inverted dropout (p = 0.5) behind fc6 / fc7 of both branches (resnet_SGG_emb.py:148-149, :161-163), applied in
    the FC epilogue, and raw cosine similarities (:215).  With the four keep masks passed in, the numbers must match the
    oracle's training-mode restatement; with masks from torch's generator the call is reproducible under manual_seed and
    about half of the fc6 activations are dropped."""
    args = synth.VrdArgs(vrd_in_channels=64, vrd_hidden=512)
    params = synth.vrd_params(98, args)
    prd = synth.prd_vectors(5, args.num_relations)
    fmap = synth.feature_map(33, 1, 64)
    boxes, rel, masks, classes, ixs, ixo = frame_inputs(orc, 34, 10)
    n, p, h = len(boxes), len(ixs), 512
    rng = np.random.default_rng(1)
    keep = [(rng.random((r, h)) >= 0.5).astype(np.uint8) for r in (n, n, p, p)]
    net = build(args, params, prd).train()
    scores, feat = net(fmap, boxes, rel, masks, classes, ixs, ixo, dropout_masks=[torch.from_numpy(k).cuda() for k in keep])
    want_s, want_f = orc.vrd_forward(params, prd, fmap, boxes, rel, masks, ixs, ixo, nthreads=orc.default_threads(),
                                     dropout_masks=keep)
    assert float(np.abs(feat - want_f).max()) <= 2e-2 * float(np.abs(want_f).max())
    got = scores.cpu().numpy()
    assert float(np.abs(got - want_s).max()) <= 2e-2 and float(np.abs(got).max()) <= 1.0 + 1e-5     # cosines, not probabilities
    assert float(np.abs(got.sum(1) - 1).min()) > 1e-3
    # an unordered-pair shortcut is ignored in training mode (every ordered pair has its own mask)
    from i2vsgg_b200 import sgg
    s2, _ = net(fmap, boxes, rel, masks, classes, ixs, ixo, dropout_masks=[torch.from_numpy(k).cuda() for k in keep],
                rel_unique=sgg.unordered_pairs(n))
    assert torch.equal(s2, scores)
    torch.manual_seed(7)
    a, _ = net(fmap, boxes, rel, masks, classes, ixs, ixo)
    torch.manual_seed(7)
    b, _ = net(fmap, boxes, rel, masks, classes, ixs, ixo)
    assert torch.equal(a, b) and not torch.equal(a, scores)
    net.eval()
    e, _ = net(fmap, boxes, rel, masks, classes, ixs, ixo)
    assert float((e.sum(1) - 1).abs().max()) <= 1e-5


def test_vrd_training_mode_against_the_executed_reference(orc):
    """Full width, training mode: our forward with the golden run's keep masks against the UNMODIFIED reference module
    executed in training mode (vrd_golden.npz: train_scores / train_feat), in both precisions."""
    gen = _gen()
    g = np.load(os.path.join(HERE, "golden", "vrd_golden.npz"))
    args = synth.VrdArgs()
    net = build(args, synth.vrd_params(gen.PARAM_SEED, args), synth.prd_vectors(gen.PRD_SEED, args.num_relations)).train()
    fmap, boxes, rel, masks, classes, ixs, ixo = gen.inputs()
    keep = [torch.from_numpy(m).cuda() for m in gen.train_masks()]
    want_s, want_f = g["train_scores"], g["train_feat"]
    for precision, bar_f, bar_s in (("bf16", 2e-2, 2e-2), ("tf32", 2e-3, 1e-3)):
        scores, feat = net(fmap, boxes, rel, masks, classes, ixs, ixo, dropout_masks=keep, precision=precision)
        e_f = float(np.abs(feat - want_f).max()) / float(np.abs(want_f).max())
        e_s = float(np.abs(scores.cpu().numpy() - want_s).max())           # cosine similarities in [-1, 1]
        print(f"vrd train vs the executed reference, {precision}: feat {e_f:.2e} scores {e_s:.2e}")
        assert e_f <= bar_f and e_s <= bar_s


def test_vrd_tf32_precision(orc):
    """precision="tf32": fp32 activations and weights through tcgen05 kind::tf32.  Against the unmodified reference module's
    output at full width (vrd_golden.npz) the scores agree to 1e-3 relative (2e-2 is the bf16 bar), and a weight update
    after the first call is picked up (version counters invalidate the derived copies)."""
    gen = _gen()
    g = np.load(os.path.join(HERE, "golden", "vrd_golden.npz"))
    args = synth.VrdArgs()
    params = synth.vrd_params(gen.PARAM_SEED, args)
    prd = synth.prd_vectors(gen.PRD_SEED, args.num_relations)
    fmap, boxes, rel, masks, classes, ixs, ixo = gen.inputs()
    net = build(args, params, prd)
    scores, feat = net(fmap, boxes, rel, masks, classes, ixs, ixo, precision="tf32")
    want_s, want_f = g["full_scores"], g["full_feat"]
    e_f = float(np.abs(feat - want_f).max()) / float(np.abs(want_f).max())
    e_s = float(np.abs(scores.cpu().numpy() / want_s - 1).max())
    s16, f16 = net(fmap, boxes, rel, masks, classes, ixs, ixo)
    b_f = float(np.abs(f16 - want_f).max()) / float(np.abs(want_f).max())
    b_s = float(np.abs(s16.cpu().numpy() / want_s - 1).max())
    print(f"vrd vs the executed reference: tf32 feat {e_f:.2e} scores {e_s:.2e}; bf16 feat {b_f:.2e} scores {b_s:.2e}")
    assert e_s <= 1e-3 and e_f <= 2e-3
    with torch.no_grad():
        net.fc_rel.fc.bias.add_(0.25)
    s2, _ = net(fmap, boxes, rel, masks, classes, ixs, ixo, precision="tf32")
    assert not torch.equal(s2, scores)


def test_clip_runner_groups_frames_without_changing_records(orc):
    """ClipRunner batches frames through one `vrd.forward`; the per-frame records must not depend on the grouping, and
    must be what `sgg.detection_output`'s kernel selects from the per-frame scores."""
    from i2vsgg_b200 import ops, sgg, shard
    from i2vsgg_b200.clip import ClipRunner
    args = synth.VrdArgs(vrd_in_channels=32, vrd_hidden=256)
    head = build(args, synth.vrd_params(3, args), synth.prd_vectors(5, args.num_relations))
    frames, n = 5, 9
    boxes, classes, conf = synth.clip_detections(8, frames, n)
    fmaps = torch.from_numpy(synth.feature_map(50, frames, 32)).cuda()
    b = torch.from_numpy(boxes).cuda()
    c = torch.from_numpy(np.tile(classes, (frames, 1))).cuda()
    s = torch.from_numpy(np.tile(conf, (frames, 1))).cuda()
    rec1, cnt1 = ClipRunner(head, synth.IM_H, synth.IM_W, 1).run(fmaps, b, c, s, frames)
    rec3, cnt3 = ClipRunner(head, synth.IM_H, synth.IM_W, 3).run(fmaps, b, c, s, frames)
    assert torch.equal(rec1, rec3) and torch.equal(cnt1, cnt3) and rec1.shape == (frames, shard.TOP_K, shard.RECORD_WIDTH)
    # frame 2 by hand: pair build -> head -> top-k
    ixs, ixo, rel, masks = sgg.build_pairs(b[2], synth.IM_H, synth.IM_W)
    rois = torch.cat([torch.zeros((n, 1), device="cuda"), b[2]], 1)
    scores, _ = head(fmaps[2:3], rois, rel, masks, None, ixs, ixo, return_numpy=False)
    want, wc = ops.triplet_topk(scores, s[2], c[2], b[2], ixs, ixo, shard.TOP_K)
    assert torch.equal(rec1[2], want) and int(cnt1[2]) == int(wc[0])


def test_clip_runner_full_path_pipelines_the_detector_side_without_changing_results(orc):
    """`run_full` issues the detector side of the next frame group (proposal decode + NMS, RoIAlignAvg) on a side stream
    under the relation stage of the current one: the records must equal `run`'s and the kept-proposal counts those of a
    plain proposal call, whatever the group size."""
    from i2vsgg_b200 import ops
    from i2vsgg_b200.clip import ClipRunner
    args = synth.VrdArgs(vrd_in_channels=32, vrd_hidden=256)
    head = build(args, synth.vrd_params(3, args), synth.prd_vectors(5, args.num_relations))
    frames, n = 7, 9
    boxes, classes, conf = synth.clip_detections(8, frames, n)
    fmaps = torch.from_numpy(synth.feature_map(50, frames, 32)).cuda()
    b = torch.from_numpy(boxes).cuda()
    c = torch.from_numpy(np.tile(classes, (frames, 1))).cuda()
    s = torch.from_numpy(np.tile(conf, (frames, 1))).cuda()
    cls_h, reg_h = synth.rpn_outputs(70, batch=frames)
    cls, reg, info = (torch.from_numpy(a).cuda() for a in (cls_h, reg_h, synth.im_info(frames)))
    want_rec, want_cnt = ClipRunner(head, synth.IM_H, synth.IM_W, 2).run(fmaps, b, c, s, frames)
    _, want_kept = ops.proposal_forward(cls, reg, info, torch.from_numpy(synth.BASE_ANCHORS).cuda(), 16, 6000, 300, 0.7,
                                        return_counts=True)
    for group in (1, 3, 4):
        rec, cnt, kept = ClipRunner(head, synth.IM_H, synth.IM_W, group).run_full(cls, reg, info, fmaps, b, c, s, frames,
                                                                                   pre_nms=6000)
        torch.cuda.synchronize()
        assert torch.equal(rec, want_rec) and torch.equal(cnt, want_cnt)
        assert torch.equal(kept.cpu(), want_kept.cpu())


@pytest.mark.parametrize("n_det", [1, 2])
def test_vrd_tiny_frames(orc, n_det):
    """One detection has no pair (the reference's detection_output returns None there, lib/utils.py:585-586); two have
    two.  The head must come back with empty / tiny tensors instead of tripping over zero-sized launches."""
    from i2vsgg_b200 import sgg
    args = synth.VrdArgs(vrd_in_channels=32, vrd_hidden=128)
    net = build(args, synth.vrd_params(4, args), synth.prd_vectors(5, args.num_relations))
    fmap = synth.feature_map(31, 1, 32)
    det, classes, _ = synth.detections(36, n_det)
    boxes = np.concatenate([np.zeros((n_det, 1), np.float32), det], 1)
    ixs, ixo, rel, masks = sgg.build_pairs(torch.from_numpy(det).cuda(), synth.IM_H, synth.IM_W)
    scores, feat = net(fmap, boxes, rel, masks, classes, ixs, ixo)
    p = n_det * (n_det - 1)
    assert scores.shape == (p, 132) and feat.shape == (p, 300)
    if p:
        want_s, want_f = orc.vrd_forward(synth.vrd_params(4, args), synth.prd_vectors(5, args.num_relations), fmap, boxes,
                                         rel.cpu().numpy(), masks.cpu().numpy(), ixs.cpu().numpy(), ixo.cpu().numpy())
        np.testing.assert_allclose(scores.cpu().numpy(), want_s, rtol=2e-2, atol=1e-6)
        s2, _ = net(fmap, boxes, rel, masks, classes, ixs, ixo, rel_unique=sgg.unordered_pairs(n_det))
        assert torch.equal(s2, scores)
