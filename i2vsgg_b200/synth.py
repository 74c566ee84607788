"""Seeded synthetic inputs for the region-level hot path.

Shapes follow a 600x1000 VidVRD frame through ResNet-101 conv4: feature map
``[B,1024,38,63]``, A=9 anchors, K*A = 21546 candidates per frame.  Everything is built
with numpy on the host from an integer seed so the CPU oracle and the CUDA path see the
same bits; callers move the arrays to the device.
"""
from __future__ import annotations

import numpy as np

IM_H, IM_W = 600.0, 1000.0
FEAT_H, FEAT_W, FEAT_C = 38, 63, 1024
NUM_ANCHORS = 9
FEAT_STRIDE = 16

# generate_anchors.py:12-37 golden table minus 1 (scales 8,16,32 x ratios .5,1,2, base 16)
BASE_ANCHORS = np.array(
    [[-84., -40., 99., 55.], [-176., -88., 191., 103.], [-360., -184., 375., 199.],
     [-56., -56., 71., 71.], [-120., -120., 135., 135.], [-248., -248., 263., 263.],
     [-36., -80., 51., 95.], [-80., -168., 95., 183.], [-168., -344., 183., 359.]], dtype=np.float32)


def feature_map(seed: int, batch: int = 1, channels: int = FEAT_C, h: int = FEAT_H, w: int = FEAT_W) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.standard_normal((batch, channels, h, w), dtype=np.float32)


def im_info(batch: int = 1) -> np.ndarray:
    return np.tile(np.array([[IM_H, IM_W, 1.0]], np.float32), (batch, 1))


def rois(seed: int, num: int, batch: int = 1, edge_frac: float = 0.05, degenerate: int = 2,
         sort_by_batch: bool = False) -> np.ndarray:
    """Stand-alone RoIs [num,5] = (b, x1, y1, x2, y2) in image pixels.

    ~`edge_frac` of them run past the right/bottom image edge (exercises the lattice op's
    `h >= H -> 0` rule) and `degenerate` of them have x2 < x1 or y2 < y1."""
    rng = np.random.default_rng(seed)
    x1 = rng.uniform(0, 900, num)
    y1 = rng.uniform(0, 500, num)
    bw = rng.uniform(16, 400, num)
    bh = rng.uniform(16, 400, num)
    x2 = np.minimum(x1 + bw, IM_W - 1)
    y2 = np.minimum(y1 + bh, IM_H - 1)
    past = rng.random(num) < edge_frac
    x2 = np.where(past, x1 + bw + 60.0, x2)
    y2 = np.where(past, y1 + bh + 60.0, y2)
    out = np.stack([rng.integers(0, batch, num).astype(np.float64), x1, y1, x2, y2], 1).astype(np.float32)
    for k in range(min(degenerate, num)):
        j = int(rng.integers(0, num))
        if k % 2 == 0:
            out[j, 3] = out[j, 1] - 40.0
        else:
            out[j, 4] = out[j, 2] - 25.0
    if sort_by_batch:
        out = out[np.argsort(out[:, 0], kind="stable")]
    return np.ascontiguousarray(out)


def _anchors(h: int = FEAT_H, w: int = FEAT_W) -> np.ndarray:
    sx, sy = np.meshgrid(np.arange(w) * FEAT_STRIDE, np.arange(h) * FEAT_STRIDE)
    shifts = np.stack([sx.ravel(), sy.ravel(), sx.ravel(), sy.ravel()], 1).astype(np.float32)
    return (BASE_ANCHORS[None, :, :] + shifts[:, None, :]).reshape(-1, 4)


def _iou_matrix(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    aw = a[:, 2] - a[:, 0] + 1
    ah = a[:, 3] - a[:, 1] + 1
    bw = b[:, 2] - b[:, 0] + 1
    bh = b[:, 3] - b[:, 1] + 1
    iw = np.clip(np.minimum(a[:, None, 2], b[None, :, 2]) - np.maximum(a[:, None, 0], b[None, :, 0]) + 1, 0, None)
    ih = np.clip(np.minimum(a[:, None, 3], b[None, :, 3]) - np.maximum(a[:, None, 1], b[None, :, 1]) + 1, 0, None)
    inter = iw * ih
    return inter / ((aw * ah)[:, None] + (bw * bh)[None, :] - inter)


def rpn_outputs(seed: int, batch: int = 1, h: int = FEAT_H, w: int = FEAT_W, clusters: int = 40):
    """(cls_prob [B,2A,h,w], bbox_pred [B,4A,h,w]) with unique fg scores per frame.

    fg scores are a permutation of linspace(0,1,K*A) (unique -> deterministic sort and NMS);
    `clusters` random ground-truth boxes per frame attract every anchor with IoU > 0.5: those anchors
    get the highest scores and deltas that regress onto the box (+ jitter), so NMS really suppresses
    the bulk of the candidates, as with a trained RPN."""
    A = NUM_ANCHORS
    KA = h * w * A
    anchors = _anchors(h, w).astype(np.float64)
    aw = anchors[:, 2] - anchors[:, 0] + 1
    ah = anchors[:, 3] - anchors[:, 1] + 1
    acx = anchors[:, 0] + 0.5 * aw
    acy = anchors[:, 1] + 0.5 * ah
    cls = np.empty((batch, 2 * A, h, w), np.float32)
    reg = np.empty((batch, 4 * A, h, w), np.float32)
    levels = np.linspace(0.0, 1.0, KA, dtype=np.float64).astype(np.float32)
    for b in range(batch):
        rng = np.random.default_rng((seed, b))
        d = np.empty((KA, 4), np.float64)
        d[:, :2] = rng.normal(0, 0.2, (KA, 2))
        d[:, 2:] = rng.normal(0, 0.3, (KA, 2))
        gw = rng.uniform(60, 420, clusters)
        gh = rng.uniform(60, 420, clusters)
        gx = rng.uniform(0, IM_W - 1 - gw)
        gy = rng.uniform(0, IM_H - 1 - gh)
        gt = np.stack([gx, gy, gx + gw, gy + gh], 1)
        iou = _iou_matrix(anchors, gt)
        best = iou.argmax(1)
        hit = iou[np.arange(KA), best] > 0.5
        g = gt[best[hit]]
        gww = g[:, 2] - g[:, 0] + 1
        ghh = g[:, 3] - g[:, 1] + 1
        tgt = np.stack([(g[:, 0] + 0.5 * gww - acx[hit]) / aw[hit], (g[:, 1] + 0.5 * ghh - acy[hit]) / ah[hit],
                        np.log(gww / aw[hit]), np.log(ghh / ah[hit])], 1)
        d[hit] = tgt + rng.normal(0, 0.05, tgt.shape)
        nh = int(hit.sum())
        rank = np.empty(KA, np.int64)
        rank[np.flatnonzero(hit)] = KA - 1 - rng.permutation(nh)        # top levels
        rank[np.flatnonzero(~hit)] = rng.permutation(KA - nh)
        fg = levels[rank]
        # anchor-major [K, A] -> NCHW channel blocks (proposal_layer.py:100-105 inverts this)
        cls[b, A:] = fg.reshape(h, w, A).transpose(2, 0, 1)
        cls[b, :A] = 1.0 - cls[b, A:]
        reg[b] = d.astype(np.float32).reshape(h, w, A * 4).transpose(2, 0, 1)
    return cls, reg


def nms_dets(seed: int, n: int) -> np.ndarray:
    """[n,5] dets (x1,y1,x2,y2,score) in random order: boxes clustered like RPN output, unique scores."""
    rng = np.random.default_rng(seed)
    m = max(4, n // 25)
    cx, cy = rng.uniform(50, 950, m), rng.uniform(50, 550, m)
    w, h = rng.uniform(30, 300, m), rng.uniform(30, 300, m)
    which = rng.integers(0, m, n)
    jit = rng.normal(0, 0.12, (n, 4))
    bw, bh = w[which] * np.exp(jit[:, 2]), h[which] * np.exp(jit[:, 3])
    bx, by = cx[which] + jit[:, 0] * w[which], cy[which] + jit[:, 1] * h[which]
    boxes = np.stack([bx - bw / 2, by - bh / 2, bx + bw / 2, by + bh / 2], 1)
    boxes[:, 0::2] = np.clip(boxes[:, 0::2], 0, 999)
    boxes[:, 1::2] = np.clip(boxes[:, 1::2], 0, 599)
    scores = rng.permutation(np.linspace(0.01, 0.99, n))
    return np.concatenate([boxes, scores[:, None]], 1).astype(np.float32)


def detections(seed: int, num: int = 64, num_classes: int = 35):
    """Config 3: `num` detections of one frame -> (boxes [num,4] fp32, classes [num] int, conf [num] fp32)."""
    rng = np.random.default_rng(seed)
    r = rois(seed, num, batch=1, edge_frac=0.0, degenerate=0)[:, 1:]
    classes = rng.integers(1, num_classes + 1, num).astype(np.int64)
    conf = rng.uniform(0.05, 1.0, num).astype(np.float32)
    return np.ascontiguousarray(r), classes, conf


def clip_detections(seed: int, frames: int, num: int = 64, sigma: float = 4.0):
    """Config 5: a clip whose detections drift by a random walk (sigma px) from frame to frame."""
    rng = np.random.default_rng(seed)
    boxes0, classes, conf = detections(seed, num)
    steps = rng.normal(0, sigma, (frames, num, 2)).astype(np.float32).cumsum(0)
    boxes = np.repeat(boxes0[None], frames, 0)
    boxes[:, :, 0::2] = np.clip(boxes[:, :, 0::2] + steps[:, :, 0:1], 0, IM_W - 1)
    boxes[:, :, 1::2] = np.clip(boxes[:, :, 1::2] + steps[:, :, 1:2], 0, IM_H - 1)
    return boxes, classes, conf


# ------------------------------------------------------------------------------------------------ SGG projection
class VrdArgs:
    """The attributes of the reference's argparse namespace that `vrd` reads (parser_func.py:137-184)."""

    def __init__(self, num_classes=35, num_relations=132, emb_dim=300, use_obj_visual=True, spatial_type=2,
                 vrd_in_channels=1024, vrd_hidden=4096):
        self.num_classes, self.num_relations, self.emb_dim = num_classes, num_relations, emb_dim
        self.use_obj_visual, self.spatial_type = use_obj_visual, spatial_type
        self.vrd_in_channels, self.vrd_hidden = vrd_in_channels, vrd_hidden   # read by i2vsgg_b200 only (small tests)
        self.source_so_prior_path = self.source_gt_rels_path = self.target_gt_rels_path = None


def vrd_params(seed: int = 1234, args: VrdArgs | None = None, pool: int = 7) -> dict:
    """Random-init parameters of `vrd` keyed like its state_dict (resnet_SGG_emb.py:83-127), fp32.
    Weights ~ N(0, 2/fan_in) so activations keep unit scale through the stack, biases ~ N(0, 0.1)."""
    a = args or VrdArgs()
    rng = np.random.default_rng(seed)
    p = {}

    def lin(name, fin, fout):
        p[name + ".weight"] = (rng.standard_normal((fout, fin), dtype=np.float32) * np.float32(np.sqrt(2.0 / fin)))
        p[name + ".bias"] = rng.standard_normal((fout,), dtype=np.float32) * np.float32(0.1)

    def conv(name, cin, cout, k):
        p[name + ".weight"] = (rng.standard_normal((cout, cin, k, k), dtype=np.float32) *
                               np.float32(np.sqrt(2.0 / (cin * k * k))))
        p[name + ".bias"] = rng.standard_normal((cout,), dtype=np.float32) * np.float32(0.1)

    lin("fc6.fc", a.vrd_in_channels * pool * pool, a.vrd_hidden)
    lin("fc7.fc", a.vrd_hidden, a.vrd_hidden)
    lin("so_vis_embeddings.fc", a.vrd_hidden, a.emb_dim)
    lin("fc8.fc", a.vrd_hidden, 256)
    n_fusion = 256
    if a.use_obj_visual:
        lin("fc_so.fc", a.emb_dim * 2, 256)
        n_fusion += 256
    if a.spatial_type == 1:
        lin("fc_lov.fc", 8, 256)
        n_fusion += 256
    elif a.spatial_type == 2:
        conv("conv_lo.0.conv", 2, 96, 5)
        conv("conv_lo.1.conv", 96, 128, 5)
        conv("conv_lo.2.conv", 128, 64, 8)
        lin("fc_lov.fc", 64, 256)
        n_fusion += 256
    lin("fc_fusion.fc", n_fusion, 256)
    lin("fc_rel.fc", 256, a.emb_dim)
    lin("prd_sem_embeddings.0", 300, 1024)
    lin("prd_sem_embeddings.2", 1024, a.emb_dim)
    return p


def prd_vectors(seed: int, num_relations: int = 132) -> np.ndarray:
    """GloVe stand-in [n_rel, 300] ~ N(0,1)."""
    return np.random.default_rng(seed).standard_normal((num_relations, 300), dtype=np.float32)


def clip_records(seed: int, frames: int = 40, tracks: int = 12, clutter: int = 8, top_k: int = 100, empty=(),
                 dropout: float = 0.08):
    """Per-frame triplet records of a clip as `i2v_triplet_topk` writes them: [frames, top_k, 13] fp32 =
    (conf, cls_s, rel, cls_o, sub box, obj box, pair idx) and counts [frames].  `tracks` relation tracks drift by a random
    walk (consecutive boxes overlap well above IoU 0.5), start / stop / drop out, and share a small label set so that label
    equality alone does not decide a match; `clutter` random predictions per frame never line up.  Rows are in random
    order with unique confidences; the frame positions in `empty` have no predictions (lib/utils.py:470-518); a track
    misses a frame with probability `dropout` (0 gives relations as long as the track, hundreds of frames)."""
    rng = np.random.default_rng(seed)
    rec = np.zeros((frames, top_k, 13), np.float32)
    cnt = np.zeros((frames,), np.int32)
    labels = [(int(rng.integers(1, 4)), int(rng.integers(0, 3)), int(rng.integers(1, 4))) for _ in range(tracks)]
    start = rng.integers(0, max(1, frames // 3), tracks)
    stop = np.minimum(frames, start + rng.integers(5, frames, tracks))
    base = rng.uniform(0.2, 0.9, tracks)

    def box():
        x1, y1 = rng.uniform(0, 800), rng.uniform(0, 450)
        return np.array([x1, y1, x1 + rng.uniform(60, 200), y1 + rng.uniform(60, 150)])

    sb = np.stack([box() for _ in range(tracks)])
    ob = np.stack([box() for _ in range(tracks)])
    used = set()
    for f in range(frames):
        sb += rng.normal(0, 4.0, sb.shape)
        ob += rng.normal(0, 4.0, ob.shape)
        if f in empty:
            continue
        rows = []
        for t in range(tracks):
            if start[t] <= f < stop[t] and rng.random() > dropout:
                rows.append((base[t] + rng.normal(0, 0.03), labels[t], sb[t].copy(), ob[t].copy()))
        for _ in range(clutter):
            rows.append((rng.uniform(0.01, 0.95), (int(rng.integers(1, 4)), int(rng.integers(0, 3)), int(rng.integers(1, 4))),
                         box(), box()))
        rng.shuffle(rows)
        for j, (c, (s, p, o), b1, b2) in enumerate(rows[:top_k]):
            c32 = np.float32(min(max(c, 1e-4), 0.9999))
            while float(c32) in used:                      # unique confidences: every sort is unambiguous
                c32 = np.nextafter(c32, np.float32(1))
            used.add(float(c32))
            rec[f, j] = [c32, s, p, o, *b1.astype(np.float32), *b2.astype(np.float32), float(rng.integers(0, 4032))]
        cnt[f] = min(len(rows), top_k)
    return rec, cnt


def records_to_frame_relations(rec, cnt, frame_numbers=None):
    """The list test_net_SGG_emb.py:209 builds for one video: [[fno, [[conf, [s,p,o], [sub box, obj box], rel idx], ...]], ...]."""
    out = []
    for f in range(rec.shape[0]):
        fno = f if frame_numbers is None else int(frame_numbers[f])
        preds = [[float(r[0]), [float(r[1]), float(r[2]), float(r[3])], [[float(v) for v in r[4:8]], [float(v) for v in r[8:12]]],
                  int(r[12])] for r in rec[f, : int(cnt[f])]]
        out.append([fno, preds])
    return out


def proposals_and_gt(seed: int, batch: int = 2, num_rois: int = 300, num_gt: int = 6, max_gt: int = 20, num_classes: int = 21):
    """Inputs of the proposal-target layer: all_rois [B,R,5] (b,x1,y1,x2,y2) and gt_boxes [B,max_gt,5] (x1,y1,x2,y2,label),
    zero-padded past `num_gt` (config.py MAX_NUM_GT_BOXES style).  A third of the proposals are jittered copies of a
    ground-truth box (foreground), the rest are random."""
    rng = np.random.default_rng(seed)
    gt = np.zeros((batch, max_gt, 5), np.float32)
    rois_out = np.zeros((batch, num_rois, 5), np.float32)
    for b in range(batch):
        x1, y1 = rng.uniform(0, 700, num_gt), rng.uniform(0, 400, num_gt)
        w, h = rng.uniform(40, 280, num_gt), rng.uniform(40, 190, num_gt)
        gt[b, :num_gt] = np.stack([x1, y1, x1 + w, y1 + h, rng.integers(1, num_classes, num_gt)], 1)
        r = rois(seed * 31 + b, num_rois, batch=1, edge_frac=0.0, degenerate=0)
        near = rng.integers(0, num_gt, num_rois // 3)
        r[: num_rois // 3, 1:] = gt[b, near, :4] + rng.normal(0, 6.0, (num_rois // 3, 4))
        r[:, 0] = b
        rois_out[b] = r
    return rois_out, gt
