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
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
