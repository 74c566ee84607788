"""Writes tests/golden/vrd_golden.npz by EXECUTING THE REFERENCE's `vrd.forward` (resnet_SGG_emb.py:128-221) on the CPU
of the build container (see oracle/ref.py::py_vrd_forward for the three stubs that makes possible).

Run from the repo root:  python tests/golden/make_vrd_golden.py
Inputs and the 230 M parameters are regenerated from seeds by i2vsgg_b200/synth.py, so the file holds outputs only.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from i2vsgg_b200 import synth  # noqa: E402
from oracle import oracle, ref  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vrd_golden.npz")
PARAM_SEED, PRD_SEED, FMAP_SEED, DET_SEED, NUM_DET = 1234, 7, 21, 22, 4


def inputs():
    fmap = synth.feature_map(FMAP_SEED, 1)
    det, classes, _ = synth.detections(DET_SEED, NUM_DET)
    ixs, ixo = oracle.enumerate_pairs(NUM_DET)
    boxes = np.concatenate([np.zeros((NUM_DET, 1), np.float32), det], 1)
    rel = oracle.union_boxes(det, ixs, ixo, synth.IM_H, synth.IM_W)
    masks = oracle.dual_masks(det, ixs, ixo, synth.IM_H, synth.IM_W)
    return fmap, boxes, rel, masks, classes, ixs, ixo


def train_masks(hidden: int = 4096):
    """The four keep masks of the training-mode run (F.dropout calls at resnet_SGG_emb.py:148, :149, :162, :163)."""
    rng = np.random.default_rng(77)
    p = NUM_DET * (NUM_DET - 1)
    return [(rng.random((r, hidden)) >= 0.5).astype(np.uint8) for r in (NUM_DET, NUM_DET, p, p)]


def main():
    assert ref.have_py_ref(), "/root/reference is not mounted"
    g = {}
    for tag, kw in (("full", {}), ("novis_loc1", dict(use_obj_visual=False, spatial_type=1))):
        args = synth.VrdArgs(**kw)
        params = synth.vrd_params(PARAM_SEED, args)
        prd = synth.prd_vectors(PRD_SEED, args.num_relations)
        fmap, boxes, rel, masks, classes, ixs, ixo = inputs()
        spatial = masks if args.spatial_type == 2 else \
            np.random.default_rng(3).standard_normal((len(ixs), 8), dtype=np.float32)
        scores, feat = ref.py_vrd_forward(params, args, prd, fmap, boxes, rel, spatial, classes, ixs, ixo)
        g[f"{tag}_scores"], g[f"{tag}_feat"] = scores.astype(np.float32), feat.astype(np.float32)
        print(tag, scores.shape, feat.shape, float(scores.max()), float(np.abs(feat).max()))
        if tag == "full":
            # the same module in training mode: dropout with the supplied keep masks, no softmax (:148-151, :215)
            ts, tf = ref.py_vrd_forward(params, args, prd, fmap, boxes, rel, spatial, classes, ixs, ixo,
                                        train_masks=train_masks())
            g["train_scores"], g["train_feat"] = ts.astype(np.float32), tf.astype(np.float32)
            print("train", ts.shape, float(np.abs(ts).max()), float(np.abs(tf).max()))
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
