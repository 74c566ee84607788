"""`generate_anchors` of lib/model/rpn/generate_anchors.py:45-105 (host side, runs once per layer)."""
from __future__ import annotations

import numpy as np


def _whctrs(anchor):
    w = anchor[2] - anchor[0] + 1
    h = anchor[3] - anchor[1] + 1
    return w, h, anchor[0] + 0.5 * (w - 1), anchor[1] + 0.5 * (h - 1)


def _mkanchors(ws, hs, x_ctr, y_ctr):
    ws, hs = np.asarray(ws, np.float64).reshape(-1, 1), np.asarray(hs, np.float64).reshape(-1, 1)
    return np.hstack((x_ctr - 0.5 * (ws - 1), y_ctr - 0.5 * (hs - 1), x_ctr + 0.5 * (ws - 1), y_ctr + 0.5 * (hs - 1)))


def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=2 ** np.arange(3, 6)):
    """Aspect ratios around the (0, 0, base-1, base-1) window, then scales of each: [len(ratios)*len(scales), 4]."""
    ratios, scales = np.asarray(ratios, np.float64), np.asarray(scales, np.float64)
    w, h, xc, yc = _whctrs(np.array([0, 0, base_size - 1, base_size - 1], np.float64))
    ws = np.round(np.sqrt(w * h / ratios))
    hs = np.round(ws * ratios)
    out = []
    for ra in _mkanchors(ws, hs, xc, yc):
        w, h, xc, yc = _whctrs(ra)
        out.append(_mkanchors(w * scales, h * scales, xc, yc))
    return np.vstack(out)
