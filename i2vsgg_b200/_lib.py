"""ctypes binding of ``lib/libi2vsgg_b200.so`` (the C ABI declared in ``include/i2vsgg_b200.h``).

There is no fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libi2vsgg_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_WORKSPACE, ERR_UNSUPPORTED = 0, 1, 2, 3, 4
POOL_NONE, POOL_AVG, POOL_MAX = 0, 1, 2
IMPL_AUTO, IMPL_GATHER, IMPL_PLANE, IMPL_ROWS, IMPL_PHASE, IMPL_BAND, IMPL_SLAB, IMPL_EVEN, IMPL_CHAN = 0, 1, 2, 3, 4, 5, 6, 7, 8
ARGMAX_FLAT, ARGMAX_PLANE = 0, 1
DT_F32, DT_BF16, DT_TF32 = 0, 1, 2

_vp, _i, _f, _sz, _ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_longlong

# name -> (restype, argtypes); mirrors include/i2vsgg_b200.h one to one
SIGNATURES = {
    "i2v_last_error": (ctypes.c_char_p, []),
    "i2v_abi_version": (_i, []),
    "ROIAlignForwardLaucher": (_i, [_vp, _f, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "ROIAlignBackwardLaucher": (_i, [_vp, _f, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "ROIPoolForwardLaucher": (_i, [_vp, _f, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "ROIPoolBackwardLaucher": (_i, [_vp, _f, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "nms_cuda_compute": (None, [_vp, _vp, _vp, _i, _i, _f]),
    "i2v_roi_align_workspace_bytes": (_sz, [_i, _i]),
    "i2v_roi_align_forward": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _i, _vp, _sz, _vp]),
    "i2v_roi_align_backward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _i, _vp, _sz, _vp]),
    "i2v_roi_pool_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "i2v_roi_pool_backward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "i2v_c_roi_align_forward": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "i2v_c_roi_align_backward": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "i2v_nms_workspace_bytes": (_sz, [_i, _i]),
    "i2v_nms_sorted": (_i, [_vp, _i, _i, _i, _f, _i, _vp, _i, _vp, _vp, _sz, _vp]),
    "i2v_nms_dets_workspace_bytes": (_sz, [_i]),
    "i2v_nms_dets": (_i, [_vp, _i, _f, _vp, _vp, _vp, _sz, _vp]),
    "i2v_proposal_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "i2v_proposal_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _sz, _vp]),
    "i2v_proposal_forward_chunk": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp, _sz, _vp]),
    "i2v_rois_add_frame": (_i, [_vp, _i, _i, _vp]),
    "i2v_proposal_forward_scores": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _sz, _vp]),
    "i2v_rpn_cls_prob": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "i2v_proposal_stages": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "i2v_roi_crop_forward": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "i2v_roi_crop_backward": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "BilinearSamplerBHWD_updateOutput_cuda_kernel": (_i, [_i] * 8 + [_vp] + [_i] * 4 + [_vp] + [_i] * 4 + [_vp] + [_i] * 4 + [_vp]),
    "BilinearSamplerBHWD_updateGradInput_cuda_kernel": (_i, [_i] * 8 + ([_vp] + [_i] * 4) * 5 + [_vp]),
    "i2v_pair_build": (_i, [_vp, _i, _f, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "i2v_pair_build_frames": (_i, [_vp, _i, _i, _f, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "i2v_triplet_topk_workspace_bytes": (_sz, [_i, _i]),
    "i2v_triplet_topk_frames_workspace_bytes": (_sz, [_i, _i, _i]),
    "i2v_triplet_topk_frames": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "i2v_triplet_topk": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "i2v_roi_pool_rows": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _ll, _i, _vp]),
    "i2v_linear_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _ll, _ll, _ll, _i, _i, _i, _vp]),
    "i2v_round_tf32": (_i, [_vp, _vp, _ll, _ll, _ll, _ll, _vp]),
    "i2v_linear_forward_dropout": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _ll, _ll, _ll, _i, _i, _i, _vp, _ll, _f, _vp]),
    "i2v_cast_bf16": (_i, [_vp, _vp, _ll, _ll, _ll, _ll, _vp]),
    "i2v_im2col_bf16": (_i, [_vp, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _i, _i, _i, _i, _vp, _ll, _vp]),
    "i2v_im2col_f32": (_i, [_vp, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _i, _i, _i, _i, _vp, _ll, _vp]),
    "i2v_pair_rows_bf16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _ll, _vp]),
    "i2v_conv2d_nhwc_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _ll, _ll, _i, _i, _vp]),
    "i2v_pair_conv1_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "i2v_pair_conv1_split_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "i2v_conv2d_nhwc_split_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _ll, _ll, _i, _i, _vp]),
    "i2v_gather_rows_bf16": (_i, [_vp, _vp, _vp, _i, _i, _i, _ll, _ll, _vp]),
    "i2v_roi_gt_overlaps": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "i2v_fg_bg_select": (_i, [_vp, _i, _i, _f, _f, _f, _vp, _vp, _vp, _vp]),
    "i2v_proposal_targets_gather": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp,
                                         _vp, _vp, _vp, _vp, _vp]),
    "i2v_anchor_overlaps": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _f, _f, _f, _vp, _vp, _vp, _vp]),
    "i2v_anchor_labels": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _f, _f, _i, _vp, _vp]),
    "i2v_anchor_disable": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "i2v_anchor_targets_finalize": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _f, _f, _f, _vp, _vp, _vp, _vp,
                                         _vp]),
    "i2v_association_workspace_bytes": (_sz, [_i, _i]),
    "i2v_greedy_association": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "i2v_rel_scores_workspace_bytes": (_sz, [_i, _i]),
    "i2v_rel_scores": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
}

_lib = None


class I2VError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Loads the library (once).  Raises if it has not been built -- there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise I2VError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C i2vsgg_b200/csrc`; i2vsgg_b200 has no CPU fallback")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError here means header and library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != OK:
        msg = load().i2v_last_error().decode("utf-8", "replace")
        raise I2VError(f"{what} failed (status {rc}): {msg}")
