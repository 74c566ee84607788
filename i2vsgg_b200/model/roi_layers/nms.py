"""`nms` of lib/model/roi_layers/nms.py:5 (`_C.nms(dets, scores, threshold)`, maskrcnn-benchmark signature:
boxes [N,4], scores [N] -> kept indices, descending score, int64)."""
from __future__ import annotations

import torch

from ... import ops


def nms(dets, scores, threshold):
    if dets.numel() == 0:
        return torch.empty((0,), dtype=torch.long, device=dets.device)
    return ops.nms_dets(torch.cat((dets[:, :4].float(), scores.float().view(-1, 1)), 1), float(threshold)).long()
