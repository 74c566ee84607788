"""The oracle's RPN-head softmax (oracle.rpn_cls_prob) against the executed reference (rpn.py:66-68, golden fixture)."""
import os

import numpy as np

from oracle import oracle

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rpn_golden.npz"))


def test_rpn_cls_prob_matches_the_executed_reference():
    for name in ("small", "wide", "a3"):
        score, want = GOLD[f"{name}_score"], GOLD[f"{name}_prob"]
        got = oracle.rpn_cls_prob(score)
        assert got.shape == want.shape and got.dtype == np.float32
        # torch's vectorised expf against the correctly rounded one: a few ulp of the result
        np.testing.assert_allclose(got, want, rtol=5e-7, atol=1e-38)
        A = score.shape[1] // 2
        np.testing.assert_allclose(got[:, :A] + got[:, A:], 1.0, atol=2e-7)


def test_rpn_cls_prob_saturates_without_nan():
    s = np.zeros((1, 2, 1, 3), np.float32)
    s[0, 0, 0] = [1000.0, -1000.0, 0.0]
    s[0, 1, 0] = [-1000.0, 1000.0, 0.0]
    p = oracle.rpn_cls_prob(s)
    assert np.array_equal(p[0, 0, 0], np.float32([1.0, 0.0, 0.5])) and np.array_equal(p[0, 1, 0], np.float32([0.0, 1.0, 0.5]))
