"""Writes tests/golden/roi_crop_golden.npz by EXECUTING the reference's own roi_crop.c (lib/model/roi_crop/src/roi_crop.c:7-103,
compiled unmodified into oracle/_ref/libref_cpu.so by `make -C oracle ref`) in the build container.

    python tests/golden/make_roi_crop_golden.py

The CPU file samples image b with grid b from a channel-last tensor; inputs are regenerated from the seed by
`crop_inputs`, so the fixture holds only what the reference produced."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref  # noqa: E402


def crop_inputs(seed: int = 3, B: int = 3, C: int = 5, H: int = 11, W: int = 13, oh: int = 6, ow: int = 7):
    rng = np.random.default_rng(seed)
    feat = rng.standard_normal((B, C, H, W)).astype(np.float32)
    grids = rng.uniform(-1.3, 1.3, (B, oh, ow, 2)).astype(np.float32)      # a good share of samples off the map
    grids[0, 0, 0] = [-1.0, -1.0]
    grids[0, 0, 1] = [1.0, 1.0]
    return feat, grids


def main():
    assert ref.have_cpu_ref(), "run `make -C oracle ref` first"
    feat, grids = crop_inputs()
    out = ref.cpu_roi_crop_forward_bhwd(feat.transpose(0, 2, 3, 1), grids).transpose(0, 3, 1, 2)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "roi_crop_golden.npz"),
                        forward=np.ascontiguousarray(out))
    print("wrote roi_crop_golden.npz", out.shape)


if __name__ == "__main__":
    main()
