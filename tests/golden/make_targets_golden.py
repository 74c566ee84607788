"""Writes tests/golden/targets_golden.npz by EXECUTING the reference's `_ProposalTargetLayer`
(lib/model/rpn/proposal_target_layer_cascade.py) on CPU tensors with numpy's generator seeded (oracle/ref.py).
    python tests/golden/make_targets_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from i2vsgg_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402

CASES = {"b2": dict(seed=3, batch=2, num_rois=300), "b1_many_fg": dict(seed=4, batch=1, num_rois=2000, num_gt=12),
         "b3_small": dict(seed=5, batch=3, num_rois=40, num_gt=2)}
ANCHOR_CASES = {"a_b2": dict(seed=6, batch=2, num_gt=5), "a_b3_crowded": dict(seed=7, batch=3, num_gt=18)}
NP_SEED = 77
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "targets_golden.npz")


def main():
    g = {}
    for name, kw in CASES.items():
        rois, gt = synth.proposals_and_gt(**kw)
        out = ref.py_proposal_target_layer(rois, gt, seed=NP_SEED)
        for k, o in zip(("rois", "labels", "targets", "inside", "outside"), out):
            g[f"{name}_{k}"] = o
        print(name, [o.shape for o in out], int((out[1] > 0).sum()), "foreground")
    for name, kw in ANCHOR_CASES.items():
        _, gt = synth.proposals_and_gt(num_rois=30, **kw)
        out = ref.py_anchor_target_layer(gt, synth.im_info(kw["batch"]), seed=NP_SEED)
        # the four outputs are mostly constant: store labels as int8, the rest only where an anchor is inside the image
        g[f"{name}_labels"] = out[0].astype(np.int8)
        nz = np.nonzero(out[1])
        g[f"{name}_targets_idx"] = np.stack(nz).astype(np.int32)
        g[f"{name}_targets_val"] = out[1][nz]
        g[f"{name}_inside"] = out[2].astype(np.int8)
        g[f"{name}_outside_val"] = np.unique(out[3])
        g[f"{name}_outside_mask"] = (out[3] > 0).astype(np.int8)
        print(name, [o.shape for o in out], int((out[0] == 1).sum()), "positive", int((out[0] == 0).sum()), "negative")
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
