"""`nms_gpu` of lib/model/nms/nms_gpu.py:7-12: rows already sorted by score, result [K,1] int32 on the device."""
from __future__ import annotations

from ... import ops


def nms_gpu(dets, thresh):
    keep, _ = ops.nms_sorted(dets, float(thresh))
    return keep.view(-1, 1)
