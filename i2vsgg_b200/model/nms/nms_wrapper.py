"""`nms` of lib/model/nms/nms_wrapper.py:13-21.

The reference always detours through `nms_cpu(dets.cpu(), thresh)` (one device->host sync per frame); this runs the
same greedy selection on the device and returns the same IntTensor of kept row indices (nms_cpu.py:6-34), bit-exact
for unique scores.  `force_cpu` is accepted for signature compatibility and ignored: there is no CPU path here."""
from __future__ import annotations

from ... import ops


def nms(dets, thresh, force_cpu=False):
    if dets.shape[0] == 0:
        return []
    return ops.nms_dets(dets, float(thresh))
