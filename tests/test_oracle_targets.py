"""CPU: the numpy restatement of the proposal-target layer (oracle/targets.py) against the reference layer's recorded
outputs (tests/golden/targets_golden.npz, written by executing lib/model/rpn/proposal_target_layer_cascade.py)."""
import importlib.util
import os

import numpy as np
import pytest

from i2vsgg_b200 import synth
from oracle import targets

HERE = os.path.dirname(os.path.abspath(__file__))


def gen():
    spec = importlib.util.spec_from_file_location("make_targets_golden", os.path.join(HERE, "golden", "make_targets_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("name", ["b2", "b1_many_fg", "b3_small"])
def test_proposal_target_oracle_equals_reference_output(name):
    m = gen()
    g = np.load(os.path.join(HERE, "golden", "targets_golden.npz"))
    rois, gt = synth.proposals_and_gt(**m.CASES[name])
    np.random.seed(m.NP_SEED)
    out = targets.proposal_target_layer(rois, gt)
    assert np.array_equal(out[0], g[f"{name}_rois"]) and np.array_equal(out[1], g[f"{name}_labels"])   # the same sample
    np.testing.assert_allclose(out[2], g[f"{name}_targets"], rtol=2e-6, atol=2e-6)                      # log: 1 ulp
    assert np.array_equal(out[3], g[f"{name}_inside"]) and np.array_equal(out[4], g[f"{name}_outside"])


def _check_anchor_outputs(out, g, name, tol):
    assert np.array_equal(out[0].astype(np.int8), g[f"{name}_labels"])                  # labels incl. the sub-sample
    idx = tuple(g[f"{name}_targets_idx"])
    dense = np.zeros_like(out[1])
    dense[idx] = g[f"{name}_targets_val"]
    np.testing.assert_allclose(out[1], dense, rtol=tol, atol=tol)
    assert np.array_equal(out[2].astype(np.int8), g[f"{name}_inside"])
    assert np.array_equal((out[3] > 0).astype(np.int8), g[f"{name}_outside_mask"])
    assert np.array_equal(np.unique(out[3]), g[f"{name}_outside_val"])


@pytest.mark.parametrize("name", ["a_b2", "a_b3_crowded"])
def test_anchor_target_oracle_equals_reference_output(name):
    m = gen()
    g = np.load(os.path.join(HERE, "golden", "targets_golden.npz"))
    kw = m.ANCHOR_CASES[name]
    _, gt = synth.proposals_and_gt(num_rois=30, **kw)
    np.random.seed(m.NP_SEED)
    out = targets.anchor_target_layer(gt, synth.im_info(kw["batch"]), synth.BASE_ANCHORS, 38, 63)
    _check_anchor_outputs(out, g, name, 2e-6)
