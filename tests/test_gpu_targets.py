"""GPU parity test (B200 box): `_ProposalTargetLayer` on the device against the oracle (pinned to the reference layer by
tests/test_oracle_targets.py) and the reference's recorded outputs: the sampled RoIs and labels are identical, regression
targets within 2e-6 (one ulp of log)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from i2vsgg_b200 import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def gen():
    spec = importlib.util.spec_from_file_location("make_targets_golden", os.path.join(HERE, "golden", "make_targets_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("name", ["b2", "b1_many_fg", "b3_small"])
def test_proposal_target_layer_equals_reference(name):
    from i2vsgg_b200.model.rpn.proposal_target_layer_cascade import _ProposalTargetLayer
    m = gen()
    g = np.load(os.path.join(HERE, "golden", "targets_golden.npz"))
    rois, gt = synth.proposals_and_gt(**m.CASES[name])
    np.random.seed(m.NP_SEED)
    out = _ProposalTargetLayer(21)(torch.from_numpy(rois).cuda(), torch.from_numpy(gt).cuda(), None)
    out = [o.cpu().numpy() for o in out]
    assert np.array_equal(out[0], g[f"{name}_rois"]) and np.array_equal(out[1], g[f"{name}_labels"])
    np.testing.assert_allclose(out[2], g[f"{name}_targets"], rtol=2e-6, atol=2e-6)
    assert np.array_equal(out[3], g[f"{name}_inside"]) and np.array_equal(out[4], g[f"{name}_outside"])


def test_overlap_reduction_matches_oracle_matrix():
    import ctypes
    from i2vsgg_b200 import _lib
    from oracle import targets
    rois, gt = synth.proposals_and_gt(9, batch=3, num_rois=500, num_gt=7)
    rois[0, 5, 1:] = [10, 10, 10, 10]                  # zero-area RoI -> -1 everywhere
    ov = targets.overlaps_batch(rois[:, :, 1:5], gt)
    r, g = torch.from_numpy(rois).cuda(), torch.from_numpy(gt).cuda()
    mx = torch.empty((3, 500), device="cuda")
    am = torch.empty((3, 500), device="cuda", dtype=torch.int32)
    lb = torch.empty((3, 500), device="cuda")
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(_lib.load().i2v_roi_gt_overlaps(P(r), 5, P(g), 3, 500, gt.shape[1], P(mx), P(am), P(lb),
                                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "overlaps")
    assert np.array_equal(mx.cpu().numpy(), ov.max(2)) and np.array_equal(am.cpu().numpy(), ov.argmax(2))
    assert float(mx[0, 5]) == -1.0


@pytest.mark.parametrize("name", ["a_b2", "a_b3_crowded"])
def test_anchor_target_layer_equals_reference(name):
    from i2vsgg_b200.model.rpn.anchor_target_layer import _AnchorTargetLayer
    from test_oracle_targets import _check_anchor_outputs
    m = gen()
    g = np.load(os.path.join(HERE, "golden", "targets_golden.npz"))
    kw = m.ANCHOR_CASES[name]
    _, gt = synth.proposals_and_gt(num_rois=30, **kw)
    layer = _AnchorTargetLayer(16, [8, 16, 32], [0.5, 1, 2])
    score = torch.zeros((kw["batch"], 18, 38, 63), device="cuda")
    np.random.seed(m.NP_SEED)
    out = layer((score, torch.from_numpy(gt).cuda(), torch.from_numpy(synth.im_info(kw["batch"])).cuda(), None))
    assert out[0].shape == (kw["batch"], 1, 9 * 38, 63) and out[1].shape == (kw["batch"], 36, 38, 63)
    _check_anchor_outputs([o.cpu().numpy() for o in out], g, name, 2e-6)
