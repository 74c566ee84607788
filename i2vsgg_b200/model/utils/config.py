"""The slice of the reference's global `cfg` that the hot path reads (lib/model/utils/config.py:143-150,194-202,
285-305).  Same attribute/item access as the easydict original; values can be overridden in place."""
from __future__ import annotations


class _Cfg(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


cfg = _Cfg(
    TRAIN=_Cfg(RPN_NMS_THRESH=0.7, RPN_PRE_NMS_TOP_N=12000, RPN_POST_NMS_TOP_N=2000, RPN_POST_NMS_TOP_N_TARGET=128,
               RPN_MIN_SIZE=8, BATCH_SIZE=128,
               # proposal targets (config.py:74-118)
               FG_FRACTION=0.25, FG_THRESH=0.5, BG_THRESH_HI=0.5, BG_THRESH_LO=0.1,
               BBOX_NORMALIZE_TARGETS_PRECOMPUTED=True, BBOX_NORMALIZE_MEANS=(0.0, 0.0, 0.0, 0.0),
               BBOX_NORMALIZE_STDS=(0.1, 0.1, 0.2, 0.2), BBOX_INSIDE_WEIGHTS=(1.0, 1.0, 1.0, 1.0),
               # anchor targets (config.py:133-156)
               RPN_POSITIVE_OVERLAP=0.7, RPN_NEGATIVE_OVERLAP=0.3, RPN_CLOBBER_POSITIVES=False, RPN_FG_FRACTION=0.5,
               RPN_BATCHSIZE=256, RPN_BBOX_INSIDE_WEIGHTS=(1.0, 1.0, 1.0, 1.0), RPN_POSITIVE_WEIGHT=-1.0),
    TEST=_Cfg(RPN_NMS_THRESH=0.7, RPN_PRE_NMS_TOP_N=6000, RPN_POST_NMS_TOP_N=300, RPN_POST_NMS_TOP_N_TARGET=128,
              RPN_MIN_SIZE=16),
    USE_GPU_NMS=True,
    POOLING_MODE="align",
    POOLING_SIZE=7,
    ANCHOR_SCALES=[8, 16, 32],
    ANCHOR_RATIOS=[0.5, 1, 2],
    FEAT_STRIDE=[16],
)
