"""Writes the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE in the build container.

Run from the repo root:  python tests/golden/make_golden.py
Needs /root/reference (read-only mount) and `make -C oracle ref`.  The fixtures hold only what the reference
produced (keep lists, proposals, decoded boxes, pooled features); inputs are regenerated from the seeds by
i2vsgg_b200/synth.py, so the files stay small.  The GPU box never runs this script.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from i2vsgg_b200 import synth  # noqa: E402
from oracle import oracle, ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    assert ref.have_py_ref(), "/root/reference is not mounted"
    oracle.build()
    g = {}

    # 1. anchors: generate_anchors.py:45-56 executed (and the comment table generate_anchors.py:12-37 minus 1)
    g["anchors"] = ref.py_generate_anchors(scales=np.array([8, 16, 32]), ratios=np.array([0.5, 1, 2]))

    # 2. nms_cpu.py:6-34 executed on unsorted dets of several sizes and thresholds
    for seed, n, thr in [(1, 1, 0.7), (2, 37, 0.7), (3, 300, 0.7), (4, 2000, 0.7), (5, 2000, 0.3), (6, 6000, 0.7),
                         (7, 12000, 0.7)]:
        dets = synth.nms_dets(seed, n)
        keep = ref.py_nms_cpu(dets, thr)
        g[f"nms_keep_{seed}_{n}_{thr}"] = keep.astype(np.int32)

    # 3. _ProposalLayer.forward executed: TEST (6000 -> 300) on two frames, TRAIN (12000 -> 2000) on one,
    #    TRAIN target (-> 128) on one
    cls, reg = synth.rpn_outputs(11, batch=2)
    info = synth.im_info(2)
    g["prop_test_b2"] = ref.py_proposal_layer(cls, reg, info, "TEST")
    cls1, reg1 = synth.rpn_outputs(12, batch=1)
    g["prop_train_b1"] = ref.py_proposal_layer(cls1, reg1, synth.im_info(1), "TRAIN")
    g["prop_train_target_b1"] = ref.py_proposal_layer(cls1, reg1, synth.im_info(1), "TRAIN", target=True)

    # 4. bbox_transform_inv + clip_boxes executed on the anchors/deltas of seed 12 (decode stage on its own)
    A = synth.NUM_ANCHORS
    anchors = synth._anchors().astype(np.float32)[None]
    deltas = reg1.transpose(0, 2, 3, 1).reshape(1, -1, 4).copy()
    g["decode_b1"] = ref.py_bbox_transform_inv_clip(anchors, deltas, synth.im_info(1))

    # 5. the reference's own roi_align.c (compiled unmodified) on a small map: lattice 8x8 and 7x7
    if ref.have_cpu_ref():
        feat = synth.feature_map(21, batch=2, channels=8)
        rois = synth.rois(22, 40, batch=2)
        g["roi_align_ref_8x8"] = ref.cpu_roi_align_forward(feat, rois, 8, 8, 1.0 / 16)
        g["roi_align_ref_7x7"] = ref.cpu_roi_align_forward(feat, rois, 7, 7, 1.0 / 16)
        # roi_pooling.c: NHWC, batch 1, forward only
        feat1 = synth.feature_map(23, batch=1, channels=8)
        rois1 = synth.rois(24, 30, batch=1, degenerate=0)
        g["roi_pool_ref_7x7"] = ref.cpu_roi_pooling_forward_nhwc(feat1.transpose(0, 2, 3, 1).copy(), rois1, 7, 7,
                                                                 1.0 / 16)

    np.savez_compressed(os.path.join(OUT, "reference_golden.npz"), **g)
    print("wrote", os.path.join(OUT, "reference_golden.npz"), {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
