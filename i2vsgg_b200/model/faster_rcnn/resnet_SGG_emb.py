"""`vrd`, the relation head of lib/model/faster_rcnn/resnet_SGG_emb.py:64-256, on the sm_100a kernels.

Same constructor arguments, sub-module / parameter names (a reference checkpoint loads with `load_state_dict`) and
`forward(fmap, boxes, rel_boxes, SpatialFea, classes, ix1, ix2)` signature and return values.  What changes underneath
(SURVEY.md section 8 a19):

* the feature map stays on the device (the reference bounces it through numpy, faster_rcnn_SGG_emb.py:159 /
  resnet_SGG_emb.py:130);
* the object rows and the union-box rows are pooled into ONE bf16 matrix [N + P, C*7*7] and go through fc6 / fc7
  together (same weights, resnet_SGG_emb.py:147-149 and :161-163), on tcgen05 tensor cores with fp32 accumulation;
* the concatenations of :169-186 are never materialised separately: each branch writes its column slice of the
  fusion input from the FC epilogue;
* conv_lo (:107-110) is three im2col + FC launches, NHWC between layers;
* normalize / cosine similarity / softmax (:207-219) is one small fp32 kernel; the predicate embedding MLP, which
  depends on parameters only, is evaluated once per `prepare()` instead of once per call.

Precision: `precision="bf16"` (default: bf16 operands, fp32 accumulation) or `"tf32"` (fp32 activations and weights
through tcgen05 kind::tf32; the pooled rows stay fp32), per call or as the attribute `self.precision`.

Training mode (`train()`) is the reference's forward in training mode: inverted dropout (p = 0.5) behind fc6 and fc7 of
both branches (:148-149, :161-163), applied in the FC kernel's epilogue with keep masks drawn from torch's CUDA generator
(or passed in), and raw cosine similarities instead of their softmax (:215).  The backward of the head (autograd through
fc6 ... fc_rel) is the trainer's and outside this path.
"""
from __future__ import annotations

import math
import os
import pickle

import numpy as np
import torch
from torch import nn

from ... import ops
from ..utils.config import cfg
from .utils import FC, Conv2d


def _dev(x, device, dtype=torch.float32):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype)
    return torch.as_tensor(np.asarray(x), device=device).to(dtype)


class vrd(nn.Module):
    def __init__(self, args, all_obj_vecs=None, all_prd_vecs=None, bn=False):
        super().__init__()
        self.args = args
        self.n_rel = args.num_relations
        self.n_obj = args.num_classes
        self.emb_dim = args.emb_dim
        self.obj_vecs = all_obj_vecs
        self.prd_vecs = all_prd_vecs
        # resnet_SGG_emb.py:75-80 unpickles three annotation files; they feed debugging code only
        # (faster_rcnn_SGG_emb.py:656), so a missing file leaves the attribute at None instead of failing
        self._so_prior = self._maybe_pickle(getattr(args, "source_so_prior_path", None), as_array=True)
        self.source_gt_rels = self._maybe_pickle(getattr(args, "source_gt_rels_path", None))
        self.target_gt_rels = self._maybe_pickle(getattr(args, "target_gt_rels_path", None))

        self.pool_size = int(getattr(cfg, "POOLING_SIZE", 7))
        self.spatial_scale = 1.0 / 16.0
        self.in_channels = int(getattr(args, "vrd_in_channels", 1024))
        hidden = int(getattr(args, "vrd_hidden", 4096))
        self.fc6 = FC(self.in_channels * self.pool_size * self.pool_size, hidden)
        self.fc7 = FC(hidden, hidden)
        self.so_vis_embeddings = FC(hidden, self.emb_dim, relu=False)
        self.fc8 = FC(hidden, 256)
        self.criterion = torch.nn.BCEWithLogitsLoss()

        n_fusion = 256
        if args.use_obj_visual:
            self.fc_so = FC(self.emb_dim * 2, 256)
            n_fusion += 256
        if args.spatial_type == 1:
            self.fc_lov = FC(8, 256)
            n_fusion += 256
        elif args.spatial_type == 2:
            self.conv_lo = nn.Sequential(Conv2d(2, 96, 5, same_padding=True, stride=2, bn=bn),
                                         Conv2d(96, 128, 5, same_padding=True, stride=2, bn=bn),
                                         Conv2d(128, 64, 8, same_padding=False, bn=bn))
            self.fc_lov = FC(64, 256)
            n_fusion += 256
        self.n_fusion = n_fusion
        self.fc_fusion = FC(n_fusion, 256)
        self.fc_rel = FC(256, self.emb_dim, relu=False)
        self.prd_sem_embeddings = nn.Sequential(nn.Linear(300, 1024), nn.LeakyReLU(0.1), nn.Linear(1024, self.emb_dim))
        self._prd_emb = None
        self._prd_key = None
        self.precision = "bf16"
        self.dropout_p = 0.5          # F.dropout's default, resnet_SGG_emb.py:148

    @staticmethod
    def _maybe_pickle(path, as_array=False):
        if not path or not os.path.exists(path):
            return None
        with open(path, "rb") as fid:
            obj = pickle.load(fid, encoding="bytes")
        return np.array(obj) if as_array else obj

    # ------------------------------------------------------------------ parameter-only work, once per weight update
    def prepare(self):
        """bf16 copies of every FC / conv weight and the predicate embeddings (resnet_SGG_emb.py:203-206)."""
        for m in self.modules():
            if isinstance(m, (FC, Conv2d)):
                m.prepare()
        dev = self.fc6.fc.weight.device
        with torch.no_grad():
            prd = torch.as_tensor(np.asarray(self.prd_vecs, dtype=np.float32), device=dev)
            self._prd_emb = self.prd_sem_embeddings(prd).float().contiguous()
        self._prd_key = self._prd_state()
        return self

    def _prd_state(self):
        return tuple((p._version, p.data_ptr()) for p in self.prd_sem_embeddings.parameters()) + \
            (str(self.fc6.fc.weight.device),)

    # ------------------------------------------------------------------ forward
    def forward(self, fmap, boxes, rel_boxes, SpatialFea, classes, ix1, ix2, return_numpy: bool = True, rel_unique=None,
                obj_masks=None, precision=None, dropout_masks=None):
        """`rel_unique=(rep, inverse)` (see `i2vsgg_b200.sgg.unordered_pairs`) tells the head that rel_boxes[inverse[p]]
        repeats rel_boxes[rep]: the union rows are then pooled and pushed through fc6 / fc7 / fc8 once per distinct box and
        fanned out afterwards -- bit-identical to the full computation, half the work for ordered pairs (eval mode only:
        in training mode every ordered pair draws its own dropout mask, as in the reference).
        `obj_masks` [N,32,32] (spatial_type 2) promises SpatialFea[p] == [obj_masks[ix1[p]], obj_masks[ix2[p]]], which is
        how faster_rcnn_SGG_emb.py:649-656 builds it; conv_lo's first layer then runs once per object instead of once per
        pair (same sums in a different fp32 order) and SpatialFea is not read at all.
        `precision`: "bf16" or "tf32" (default `self.precision`).
        `dropout_masks` (training mode): the four uint8 keep masks [N,h], [N,h], [P,h], [P,h] of the F.dropout calls at
        :148, :149, :162, :163 in that order; drawn from torch's generator when None."""
        dev = self.fc6.fc.weight.device
        if dev.type != "cuda":
            raise RuntimeError("vrd: parameters must live on a CUDA device (there is no CPU path)")
        precision = precision or self.precision
        if precision not in ("bf16", "tf32"):
            raise ValueError(f"vrd: precision must be 'bf16' or 'tf32', not {precision!r}")
        tf32 = precision == "tf32"
        act = torch.float32 if tf32 else torch.bfloat16
        training = self.training
        if self._prd_emb is None or self._prd_key != self._prd_state():
            self.prepare()
        fmap = _dev(fmap, dev).contiguous()
        boxes = _dev(boxes, dev).reshape(-1, 5).contiguous()
        rel_boxes = _dev(rel_boxes, dev).reshape(-1, 5).contiguous()
        ix1 = _dev(ix1, dev, torch.int64).reshape(-1)
        ix2 = _dev(ix2, dev, torch.int64).reshape(-1)
        n_obj, n_pair = boxes.size(0), rel_boxes.size(0)
        ps = self.pool_size
        if rel_unique is not None and not training:
            rep, inverse = rel_unique
            pool_boxes = rel_boxes.index_select(0, rep.to(dev))
            inverse = inverse.to(dev)
        else:
            pool_boxes, inverse = rel_boxes, None
        n_uni = pool_boxes.size(0)
        hidden = self.fc6.fc.out_features

        keep6 = keep7 = None
        scale = 1.0
        if training and self.dropout_p > 0:
            scale = 1.0 / (1.0 - self.dropout_p)
            if dropout_masks is None:
                dropout_masks = [(torch.rand((n, hidden), device=dev) >= self.dropout_p).to(torch.uint8)
                                 for n in (n_obj, n_obj, n_uni, n_uni)]
            m = [_dev(t, dev, torch.uint8).contiguous() for t in dropout_masks]
            if [tuple(t.shape) for t in m] != [(n_obj, hidden), (n_obj, hidden), (n_uni, hidden), (n_uni, hidden)]:
                raise ValueError("vrd: dropout_masks must be [N,h], [N,h], [P,h], [P,h] keep masks")
            keep6, keep7 = torch.cat((m[0], m[2])), torch.cat((m[1], m[3]))

        # roi_pool of objects and unions -> one matrix; fc6 / fc7 over all rows at once (:144-149, :158-163)
        k6 = self.in_channels * ps * ps
        pooled = torch.empty((n_obj + n_uni, k6), dtype=act, device=dev)
        # one launch for both box sets: the feature planes are staged in shared memory once per CTA
        ops.roi_pool_rows(fmap, torch.cat((boxes, pool_boxes)), ps, ps, self.spatial_scale, out=pooled)
        h = self.fc7(self.fc6(pooled, out_dtype=act, keep_mask=keep6, keep_scale=scale), out_dtype=act, keep_mask=keep7,
                     keep_scale=scale)
        obj_feature = self.so_vis_embeddings(h[:n_obj], out_dtype=torch.float32)            # :150

        fusion = torch.empty((n_pair, self.n_fusion), dtype=act, device=dev)
        col = 0
        if inverse is None:
            self.fc8(h[n_obj:], out=fusion[:, col:col + 256])                               # :164
        elif tf32:
            fusion[:, col:col + 256] = self.fc8(h[n_obj:], out_dtype=act).index_select(0, inverse)
        else:
            ops.gather_rows_bf16(self.fc8(h[n_obj:]), inverse, out=fusion[:, col:col + 256])
        col += 256
        if self.args.use_obj_visual:                                                        # :166-170
            if tf32:
                so = torch.cat((obj_feature.index_select(0, ix1), obj_feature.index_select(0, ix2)), 1)
            else:
                so = ops.pair_rows_bf16(obj_feature, ix1, ix2)
            self.fc_so(so, out=fusion[:, col:col + 256])
            col += 256
        if self.args.spatial_type == 1:                                                     # :172-174
            sp = _dev(SpatialFea, dev).reshape(n_pair, 8).contiguous()
            self.fc_lov(sp if tf32 else ops.cast_bf16(sp), out=fusion[:, col:col + 256])
            col += 256
        elif self.args.spatial_type == 2:                                                   # :175-179
            if tf32:
                # fp32 patches and weights, layer by layer (the pair masks are formed from the object masks if need be)
                if obj_masks is not None:
                    om = _dev(obj_masks, dev).reshape(n_obj, 32, 32)
                    sp = torch.stack((om.index_select(0, ix1), om.index_select(0, ix2)), 1).contiguous()
                else:
                    sp = _dev(SpatialFea, dev).reshape(n_pair, 2, 32, 32).contiguous()
                lo = self.conv_lo[0].forward_tf32(sp, "nchw")
                lo = self.conv_lo[1].forward_tf32(lo, "nhwc")
                lo = self.conv_lo[2].forward_tf32(lo, "nhwc").reshape(n_pair, 64)
            elif obj_masks is not None:
                # the first layer writes parity planes when the second (stride 2) can read them as dense TMA boxes
                c0 = self.conv_lo[0].conv
                oh0 = (32 + 2 * c0.padding[0] - c0.kernel_size[0]) // c0.stride[0] + 1
                split = self.conv_lo[1].takes_split(oh0, oh0, c0.out_channels)
                lo = self.conv_lo[0].forward_pairs(_dev(obj_masks, dev).reshape(n_obj, 32, 32).contiguous(), ix1, ix2,
                                                   split=split)
                lo = self.conv_lo[1](lo, "nhwc_split" if split else "nhwc")
                lo = self.conv_lo[2](lo, "nhwc").reshape(n_pair, 64)
            else:
                sp = _dev(SpatialFea, dev).reshape(n_pair, 2, 32, 32).contiguous()
                lo = self.conv_lo[0](sp, "nchw")
                lo = self.conv_lo[1](lo, "nhwc")
                lo = self.conv_lo[2](lo, "nhwc").reshape(n_pair, 64)
            self.fc_lov(lo, out=fusion[:, col:col + 256])
            col += 256
        x = self.fc_rel(self.fc_fusion(fusion, out_dtype=act), out_dtype=torch.float32)     # :190-191
        scores = ops.rel_scores(x, self._prd_emb, softmax=not training)                     # :203-219
        return scores, (x.detach().cpu().numpy() if return_numpy else x)

    def save_semantic_embedding(self, save_path):
        if self._prd_emb is None or self._prd_key != self._prd_state():
            self.prepare()
        np.save(save_path, self._prd_emb.detach().cpu().numpy())

    def _set_trainable(self, model, requires_grad):
        for param in model.parameters():
            param.requires_grad = requires_grad

    # ------------------------------------------------------------------ host helpers kept for callers (:240-256)
    def _getUnionBBox(self, aBB, bBB, ih, iw, margin=10):
        return [max(0, min(aBB[0], bBB[0]) - margin), max(0, min(aBB[1], bBB[1]) - margin),
                min(iw, max(aBB[2], bBB[2]) + margin), min(ih, max(aBB[3], bBB[3]) + margin)]

    def _getDualMask(self, ih, iw, bb):
        rh = 32.0 / ih
        rw = 32.0 / iw
        x1 = max(0, int(math.floor(bb[0] * rw)))
        x2 = min(32, int(math.ceil(bb[2] * rw)))
        y1 = max(0, int(math.floor(bb[1] * rh)))
        y2 = min(32, int(math.ceil(bb[3] * rh)))
        mask = np.zeros((32, 32))
        mask[y1:y2, x1:x2] = 1
        return mask
