"""i2vsgg_b200 -- sm_100a implementation of the I2VSGG region-level hot path.

``i2vsgg_b200.ops``    tensor-level functions over the C ABI (``include/i2vsgg_b200.h``)
``i2vsgg_b200.model``  the reference's module tree for this path (``model.roi_align``, ``model.roi_pooling``,
                       ``model.nms``, ``model.rpn.proposal_layer``, ``model.roi_layers``) with the same call
                       signatures; ``install_as_model()`` makes it importable under the reference's own
                       dotted names so it drops in over the PyTorch-0.4 cffi extensions.
``i2vsgg_b200.sgg``    pair-feature builder / triplet selection of the SGG stage; the relation head itself is
                       ``i2vsgg_b200.model.faster_rcnn.resnet_SGG_emb.vrd`` (``install_vrd()`` patches it into the reference)
``i2vsgg_b200.shard``  frame sharding over ranks and the triplet all-gather
"""
from __future__ import annotations

import importlib
import sys

__version__ = "0.1.0"

_MODEL_MODULES = [
    "model", "model.utils", "model.utils.config",
    "model.roi_align", "model.roi_align.functions", "model.roi_align.functions.roi_align",
    "model.roi_align.modules", "model.roi_align.modules.roi_align",
    "model.roi_pooling", "model.roi_pooling.functions", "model.roi_pooling.functions.roi_pool",
    "model.roi_pooling.modules", "model.roi_pooling.modules.roi_pool",
    "model.roi_crop", "model.roi_crop.functions", "model.roi_crop.functions.roi_crop",
    "model.roi_crop.modules", "model.roi_crop.modules.roi_crop",
    "model.nms", "model.nms.nms_wrapper", "model.nms.nms_gpu",
    "model.rpn", "model.rpn.generate_anchors", "model.rpn.proposal_layer", "model.rpn.proposal_target_layer_cascade",
    "model.rpn.anchor_target_layer", "model.rpn.rpn",
    "model.roi_layers", "model.roi_layers.roi_align", "model.roi_layers.roi_pool", "model.roi_layers.nms",
]


def install_as_model(force: bool = False) -> None:
    """Registers ``i2vsgg_b200.model.*`` under the reference's names (``model.roi_align.modules.roi_align`` ...).

    After this, ``from model.roi_align.modules.roi_align import RoIAlignAvg`` (faster_rcnn_instance_styleD_bilinear.py:13)
    resolves to this package.  Modules of the reference that are outside the hot path are left alone."""
    for name in _MODEL_MODULES:
        if name in sys.modules and not force:
            continue
        sys.modules[name] = importlib.import_module("i2vsgg_b200." + name)


def install_vrd(reference_module=None):
    """Replaces the class ``vrd`` of the reference's ``model.faster_rcnn.resnet_SGG_emb`` (lib/model/faster_rcnn/
    resnet_SGG_emb.py:64) by the sm_100a relation head.  Only the class is patched: the module also holds the ResNet
    backbone, which is not part of this path.  Returns the class."""
    from .model.faster_rcnn.resnet_SGG_emb import vrd
    if reference_module is None:
        reference_module = sys.modules.get("model.faster_rcnn.resnet_SGG_emb")
    if reference_module is not None:
        reference_module.vrd = vrd
    return vrd
