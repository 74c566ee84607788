// Training-side target assignment between NMS and RoIAlign: `_ProposalTargetLayer`
// (lib/model/rpn/proposal_target_layer_cascade.py:33-212) with `bbox_overlaps_batch` / `bbox_transform_batch`
// (lib/model/rpn/bbox_transform.py:168-257, :36-75).  SURVEY.md section 8(f) rank 2.
//
// The reference does this with ~40 elementwise torch launches, a [B, R, K] overlap tensor and a Python loop per image
// with `torch.nonzero` and numpy random draws.  Here: one kernel reduces the overlap matrix on the fly (max / first
// arg-max per RoI, never materialised), one compacts the foreground / background index lists in order, and -- after the
// host has drawn the same numpy random numbers the reference draws (the sample has to be THE reference's sample, so the
// generator stays numpy's) -- one kernel gathers RoIs, labels, regression targets and weights.
// Every comparison that decides an index (overlap against a threshold) is made on fp32 values computed with the
// reference's operation order and no FMA contraction (this file is compiled -fmad=false).
#include "common.cuh"

namespace i2v {
namespace {

// bbox_transform.py:226-257 for one (roi, gt) pair; rois as (x1,y1,x2,y2)
__device__ __forceinline__ float overlap(const float4 a, const float4 g) {
    const float gw = (g.z - g.x) + 1.f, gh = (g.w - g.y) + 1.f;
    const float aw = (a.z - a.x) + 1.f, ah = (a.w - a.y) + 1.f;
    float iw = (fminf(a.z, g.z) - fmaxf(a.x, g.x)) + 1.f;
    float ih = (fminf(a.w, g.w) - fmaxf(a.y, g.y)) + 1.f;
    if (iw < 0.f) iw = 0.f;
    if (ih < 0.f) ih = 0.f;
    const float inter = iw * ih;
    const float ua = (aw * ah + gw * gh) - inter;
    float ov = inter / ua;
    if (gw == 1.f && gh == 1.f) ov = 0.f;      // zero-area (padding) ground truth
    if (aw == 1.f && ah == 1.f) ov = -1.f;     // zero-area RoI
    return ov;
}

// One thread per RoI: max and FIRST arg-max over the K ground-truth boxes of its image (torch.max(dim) semantics),
// plus the label of the assigned box (proposal_target_layer_cascade.py:124-133).
__global__ void __launch_bounds__(256) roi_gt_overlap_kernel(const float* __restrict__ rois, int roi_stride, int roi_off,
                                                             const float* __restrict__ gt, int B, int R, int K,
                                                             float* __restrict__ max_ov, int* __restrict__ assign,
                                                             float* __restrict__ labels) {
    extern __shared__ float s_gt[];     // [K][5] of this image
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < K * 5; i += blockDim.x) s_gt[i] = gt[(size_t)b * K * 5 + i];
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float* p = rois + ((size_t)b * R + r) * roi_stride + roi_off;
    const float4 a = make_float4(p[0], p[1], p[2], p[3]);
    float best = -INFINITY;
    int arg = 0;
    for (int k = 0; k < K; ++k) {
        const float ov = overlap(a, make_float4(s_gt[k * 5], s_gt[k * 5 + 1], s_gt[k * 5 + 2], s_gt[k * 5 + 3]));
        if (ov > best) {
            best = ov;
            arg = k;
        }
    }
    max_ov[(size_t)b * R + r] = best;
    assign[(size_t)b * R + r] = arg;
    if (labels) labels[(size_t)b * R + r] = K > 0 ? s_gt[arg * 5 + 4] : 0.f;
}

// One CTA per image: the indices with max_ov >= fg_thresh, and those with bg_lo <= max_ov < bg_hi, each in ascending
// order (torch.nonzero), and their counts (:139-146).
__global__ void __launch_bounds__(256) fg_bg_select_kernel(const float* __restrict__ max_ov, int R, float fg_thresh,
                                                           float bg_hi, float bg_lo, int* __restrict__ fg_inds,
                                                           int* __restrict__ bg_inds, int* __restrict__ counts) {
    __shared__ int s_warp[2][8];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int base_fg = 0, base_bg = 0;
    for (int i0 = 0; i0 < R; i0 += 256) {
        const int i = i0 + tid;
        const float v = i < R ? max_ov[(size_t)b * R + i] : -2.f;
        const bool fg = i < R && v >= fg_thresh;
        const bool bg = i < R && v < bg_hi && v >= bg_lo;
        const unsigned mf = __ballot_sync(0xffffffffu, fg), mb = __ballot_sync(0xffffffffu, bg);
        __syncthreads();
        if (lane == 0) {
            s_warp[0][warp] = __popc(mf);
            s_warp[1][warp] = __popc(mb);
        }
        __syncthreads();
        int bf = 0, bb = 0, tf = 0, tb = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            if (w < warp) {
                bf += s_warp[0][w];
                bb += s_warp[1][w];
            }
            tf += s_warp[0][w];
            tb += s_warp[1][w];
        }
        const unsigned below = (1u << lane) - 1u;
        if (fg) fg_inds[(size_t)b * R + base_fg + bf + __popc(mf & below)] = i;
        if (bg) bg_inds[(size_t)b * R + base_bg + bb + __popc(mb & below)] = i;
        base_fg += tf;
        base_bg += tb;
    }
    if (tid == 0) {
        counts[b * 2] = base_fg;
        counts[b * 2 + 1] = base_bg;
    }
}

struct TargetNorm {
    float mean[4], stdv[4], inside[4];
    int normalize;
};

// One thread per sampled RoI (:189-209, bbox_transform_batch, _get_bbox_regression_labels_pytorch).
__global__ void __launch_bounds__(256) proposal_target_gather_kernel(
    const float* __restrict__ rois, const float* __restrict__ gt, const int* __restrict__ assign,
    const float* __restrict__ labels, const int* __restrict__ fg_inds, const int* __restrict__ bg_inds,
    const int* __restrict__ positions, const int* __restrict__ fg_this, int B, int R, int K, int S, TargetNorm nm,
    float* __restrict__ rois_out, float* __restrict__ labels_out, float* __restrict__ targets_out,
    float* __restrict__ inside_out, float* __restrict__ outside_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * S) return;
    const int b = idx / S, j = idx - b * S;
    const int nfg = fg_this[b];
    const int pos = positions[idx];
    const int keep = (j < nfg ? fg_inds : bg_inds)[(size_t)b * R + pos];
    const float* r = rois + ((size_t)b * R + keep) * 5;
    const float label = j < nfg ? labels[(size_t)b * R + keep] : 0.f;   // background RoIs are clamped to 0 (:195-196)
    float* ro = rois_out + (size_t)idx * 5;
    ro[0] = (float)b;
    ro[1] = r[1];
    ro[2] = r[2];
    ro[3] = r[3];
    ro[4] = r[4];
    labels_out[idx] = label;
    const float* g = gt + ((size_t)b * K + assign[(size_t)b * R + keep]) * 5;
    // bbox_transform.py:55-68
    const float ew = (r[3] - r[1]) + 1.0f, eh = (r[4] - r[2]) + 1.0f;
    const float ecx = r[1] + 0.5f * ew, ecy = r[2] + 0.5f * eh;
    const float gw = (g[2] - g[0]) + 1.0f, gh = (g[3] - g[1]) + 1.0f;
    const float gcx = g[0] + 0.5f * gw, gcy = g[1] + 0.5f * gh;
    float t[4] = {(gcx - ecx) / ew, (gcy - ecy) / eh, logf(gw / ew), logf(gh / eh)};
    const bool pos_label = label > 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float v = t[c];
        if (nm.normalize) v = (v - nm.mean[c]) / nm.stdv[c];       // :108-111
        const float iw = pos_label ? nm.inside[c] : 0.f;
        targets_out[(size_t)idx * 4 + c] = pos_label ? v : 0.f;
        inside_out[(size_t)idx * 4 + c] = iw;
        outside_out[(size_t)idx * 4 + c] = iw > 0.f ? 1.f : 0.f;   // :56
    }
}


// ------------------------------------------------------------------------------------------ anchor targets
// `_AnchorTargetLayer` (lib/model/rpn/anchor_target_layer.py:48-193).  Anchor j = (y*W + x)*A + a as in the proposal
// layer; an anchor is "inside" when it lies within the FIRST image's bounds (:81-84 reads im_info[0] for the whole batch).
__device__ __forceinline__ float4 grid_anchor(const float* __restrict__ base, int j, int A, int W, int stride) {
    const int a = j % A, k = j / A;
    const float sx = (float)((k % W) * stride), sy = (float)((k / W) * stride);
    return make_float4(base[a * 4] + sx, base[a * 4 + 1] + sy, base[a * 4 + 2] + sx, base[a * 4 + 3] + sy);
}
__device__ __forceinline__ bool anchor_inside(const float4 a, float border, float im_w, float im_h) {
    return a.x >= -border && a.y >= -border && a.z < im_w + border && a.w < im_h + border;
}

// Pass 1: per inside anchor the max / first arg-max overlap over the image's ground truth (:100), and per ground-truth
// box the max over anchors (:101) through an integer atomicMax (overlaps of generated anchors are >= 0, where the float
// order is the order of the bit patterns).  max_ov = -2 marks an outside anchor.
__global__ void __launch_bounds__(256) anchor_overlap_kernel(const float* __restrict__ base, const float* __restrict__ gt,
                                                             int total, int A, int W, int stride, int G, float border,
                                                             float im_w, float im_h, float* __restrict__ max_ov,
                                                             int* __restrict__ argmax, int* __restrict__ gt_max_bits) {
    extern __shared__ float s_gt[];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < G * 5; i += blockDim.x) s_gt[i] = gt[(size_t)b * G * 5 + i];
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= total) return;
    const float4 a = grid_anchor(base, j, A, W, stride);
    if (!anchor_inside(a, border, im_w, im_h)) {
        max_ov[(size_t)b * total + j] = -2.f;
        argmax[(size_t)b * total + j] = 0;
        return;
    }
    float best = -INFINITY;
    int arg = 0;
    for (int g = 0; g < G; ++g) {
        const float ov = overlap(a, make_float4(s_gt[g * 5], s_gt[g * 5 + 1], s_gt[g * 5 + 2], s_gt[g * 5 + 3]));
        if (ov > best) {
            best = ov;
            arg = g;
        }
        if (ov > 0.f) atomicMax(gt_max_bits + b * G + g, __float_as_int(ov));
    }
    max_ov[(size_t)b * total + j] = best;
    argmax[(size_t)b * total + j] = arg;
}

// Pass 2: the label rules of :103-119 (labels as floats: 1 positive, 0 negative, -1 don't care / outside).
__global__ void __launch_bounds__(256) anchor_label_kernel(const float* __restrict__ base, const float* __restrict__ gt,
                                                           int total, int A, int W, int stride, int G,
                                                           const float* __restrict__ max_ov, const int* __restrict__ gt_max_bits,
                                                           float neg_thresh, float pos_thresh, int clobber,
                                                           float* __restrict__ labels) {
    extern __shared__ float s_gt[];      // [G][5] then [G] gt maxima
    float* s_max = s_gt + G * 5;
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < G * 5; i += blockDim.x) s_gt[i] = gt[(size_t)b * G * 5 + i];
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const float m = __int_as_float(gt_max_bits[b * G + g]);
        s_max[g] = (m == 0.f) ? 1e-5f : m;                                   // :106
    }
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= total) return;
    const float mo = max_ov[(size_t)b * total + j];
    float label = -1.f;
    if (mo != -2.f) {
        const float4 a = grid_anchor(base, j, A, W, stride);
        bool is_gt_best = false;
        for (int g = 0; g < G; ++g)
            is_gt_best |= overlap(a, make_float4(s_gt[g * 5], s_gt[g * 5 + 1], s_gt[g * 5 + 2], s_gt[g * 5 + 3])) == s_max[g];
        if (!clobber && mo < neg_thresh) label = 0.f;
        if (is_gt_best) label = 1.f;
        if (mo >= pos_thresh) label = 1.f;
        if (clobber && mo < neg_thresh) label = 0.f;
    }
    labels[(size_t)b * total + j] = label;
}

// labels[b][list[b][positions[b][k]]] = -1 for k < n_disable[b]  (:131-145: the sub-sampled anchors become don't-care)
__global__ void __launch_bounds__(256) anchor_disable_kernel(float* __restrict__ labels, const int* __restrict__ list,
                                                             const int* __restrict__ positions, const int* __restrict__ n_disable,
                                                             int total, int max_disable) {
    const int b = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_disable[b]) return;
    labels[(size_t)b * total + list[(size_t)b * total + positions[(size_t)b * max_disable + k]]] = -1.f;
}

// Pass 3: regression targets against the arg-max box (:150), weights (:153-165) and the output layouts of :167-191:
// labels [B,1,A*H,W], targets / inside / outside [B,4A,H,W].
__global__ void __launch_bounds__(256) anchor_finalize_kernel(const float* __restrict__ base, const float* __restrict__ gt,
                                                              int total, int A, int H, int W, int stride, int G,
                                                              const float* __restrict__ max_ov, const int* __restrict__ argmax,
                                                              const float* __restrict__ labels, float inside_w,
                                                              float pos_w, float neg_w, float* __restrict__ labels_out,
                                                              float* __restrict__ targets_out, float* __restrict__ inside_out,
                                                              float* __restrict__ outside_out) {
    const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= total) return;
    const int a = j % A, k = j / A, x = k % W, y = k / W;
    const float label = labels[(size_t)b * total + j];
    labels_out[((size_t)b * A + a) * H * W + (size_t)y * W + x] = label;
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    if (max_ov[(size_t)b * total + j] != -2.f) {
        const float4 e = grid_anchor(base, j, A, W, stride);
        const float* g = gt + ((size_t)b * G + argmax[(size_t)b * total + j]) * 5;
        const float ew = (e.z - e.x) + 1.0f, eh = (e.w - e.y) + 1.0f;
        const float ecx = e.x + 0.5f * ew, ecy = e.y + 0.5f * eh;
        const float gw = (g[2] - g[0]) + 1.0f, gh = (g[3] - g[1]) + 1.0f;
        const float gcx = g[0] + 0.5f * gw, gcy = g[1] + 0.5f * gh;
        t[0] = (gcx - ecx) / ew;
        t[1] = (gcy - ecy) / eh;
        t[2] = logf(gw / ew);
        t[3] = logf(gh / eh);
    }
    const float iw = label == 1.f ? inside_w : 0.f;
    const float ow = label == 1.f ? pos_w : (label == 0.f ? neg_w : 0.f);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const size_t o = ((size_t)b * 4 * A + a * 4 + c) * H * W + (size_t)y * W + x;
        targets_out[o] = t[c];
        inside_out[o] = iw;
        outside_out[o] = ow;
    }
}

}  // namespace
}  // namespace i2v

using namespace i2v;

extern "C" int i2v_roi_gt_overlaps(const float* rois, int roi_width, const float* gt_boxes, int batch, int num_rois,
                                   int num_gt, float* max_overlaps, int* assignment, float* labels, cudaStream_t stream) {
    I2V_REQUIRE(batch >= 0 && num_rois >= 0 && num_gt >= 0 && (roi_width == 4 || roi_width == 5), "roi_gt_overlaps: bad shape");
    if (batch == 0 || num_rois == 0) return I2V_OK;
    I2V_REQUIRE(rois && max_overlaps && assignment && (gt_boxes || num_gt == 0), "roi_gt_overlaps: null pointer");
    size_t smem = (size_t)(num_gt > 0 ? num_gt : 1) * 5 * sizeof(float);
    I2V_REQUIRE(smem <= 48 * 1024, "roi_gt_overlaps: more than 2457 ground-truth boxes per image");
    dim3 grid((unsigned)ceil_div(num_rois, 256), (unsigned)batch);
    roi_gt_overlap_kernel<<<grid, 256, smem, stream>>>(rois, roi_width, roi_width == 5 ? 1 : 0, gt_boxes, batch, num_rois,
                                                      num_gt, max_overlaps, assignment, labels);
    return check_launch("roi_gt_overlap_kernel");
}

extern "C" int i2v_fg_bg_select(const float* max_overlaps, int batch, int num_rois, float fg_thresh, float bg_thresh_hi,
                                float bg_thresh_lo, int* fg_inds, int* bg_inds, int* counts, cudaStream_t stream) {
    I2V_REQUIRE(batch >= 0 && num_rois >= 0, "fg_bg_select: bad shape");
    if (batch == 0) return I2V_OK;
    I2V_REQUIRE(max_overlaps && fg_inds && bg_inds && counts, "fg_bg_select: null pointer");
    fg_bg_select_kernel<<<batch, 256, 0, stream>>>(max_overlaps, num_rois, fg_thresh, bg_thresh_hi, bg_thresh_lo, fg_inds,
                                                   bg_inds, counts);
    return check_launch("fg_bg_select_kernel");
}

extern "C" int i2v_proposal_targets_gather(const float* rois, const float* gt_boxes, const int* assignment,
                                           const float* labels, const int* fg_inds, const int* bg_inds,
                                           const int* positions, const int* fg_this, int batch, int num_rois, int num_gt,
                                           int rois_per_image, const float* means, const float* stds,
                                           const float* inside_weights, int normalize, float* rois_out, float* labels_out,
                                           float* targets_out, float* inside_out, float* outside_out, cudaStream_t stream) {
    I2V_REQUIRE(batch >= 0 && num_rois >= 1 && num_gt >= 1 && rois_per_image >= 0, "proposal_targets_gather: bad shape");
    if (batch == 0 || rois_per_image == 0) return I2V_OK;
    I2V_REQUIRE(rois && gt_boxes && assignment && labels && fg_inds && bg_inds && positions && fg_this && means && stds &&
                    inside_weights && rois_out && labels_out && targets_out && inside_out && outside_out,
                "proposal_targets_gather: null pointer");
    TargetNorm nm;
    for (int c = 0; c < 4; ++c) {          // host pointers: four floats each
        nm.mean[c] = means[c];
        nm.stdv[c] = stds[c];
        nm.inside[c] = inside_weights[c];
    }
    nm.normalize = normalize;
    int total = batch * rois_per_image;
    proposal_target_gather_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(
        rois, gt_boxes, assignment, labels, fg_inds, bg_inds, positions, fg_this, batch, num_rois, num_gt, rois_per_image, nm,
        rois_out, labels_out, targets_out, inside_out, outside_out);
    return check_launch("proposal_target_gather_kernel");
}

extern "C" int i2v_anchor_overlaps(const float* base_anchors, const float* gt_boxes, int batch, int num_anchors, int height,
                                   int width, int feat_stride, int num_gt, float allowed_border, float im_w, float im_h,
                                   float* max_overlaps, int* argmax, int* gt_max_bits, cudaStream_t stream) {
    I2V_REQUIRE(batch >= 0 && num_anchors >= 1 && height >= 1 && width >= 1 && num_gt >= 1, "anchor_overlaps: bad shape");
    if (batch == 0) return I2V_OK;
    I2V_REQUIRE(base_anchors && gt_boxes && max_overlaps && argmax && gt_max_bits, "anchor_overlaps: null pointer");
    const int total = num_anchors * height * width;
    I2V_CUDA_TRY(cudaMemsetAsync(gt_max_bits, 0, sizeof(int) * (size_t)batch * num_gt, stream));
    size_t smem = (size_t)num_gt * 5 * sizeof(float);
    I2V_REQUIRE(smem <= 40 * 1024, "anchor_overlaps: too many ground-truth boxes per image");
    dim3 grid((unsigned)ceil_div(total, 256), (unsigned)batch);
    anchor_overlap_kernel<<<grid, 256, smem, stream>>>(base_anchors, gt_boxes, total, num_anchors, width, feat_stride, num_gt,
                                                      allowed_border, im_w, im_h, max_overlaps, argmax, gt_max_bits);
    return check_launch("anchor_overlap_kernel");
}

extern "C" int i2v_anchor_labels(const float* base_anchors, const float* gt_boxes, int batch, int num_anchors, int height,
                                 int width, int feat_stride, int num_gt, const float* max_overlaps, const int* gt_max_bits,
                                 float negative_overlap, float positive_overlap, int clobber_positives, float* labels,
                                 cudaStream_t stream) {
    I2V_REQUIRE(batch >= 0 && num_anchors >= 1 && height >= 1 && width >= 1 && num_gt >= 1, "anchor_labels: bad shape");
    if (batch == 0) return I2V_OK;
    I2V_REQUIRE(base_anchors && gt_boxes && max_overlaps && gt_max_bits && labels, "anchor_labels: null pointer");
    const int total = num_anchors * height * width;
    size_t smem = (size_t)num_gt * 6 * sizeof(float);
    dim3 grid((unsigned)ceil_div(total, 256), (unsigned)batch);
    anchor_label_kernel<<<grid, 256, smem, stream>>>(base_anchors, gt_boxes, total, num_anchors, width, feat_stride, num_gt,
                                                    max_overlaps, gt_max_bits, negative_overlap, positive_overlap,
                                                    clobber_positives, labels);
    return check_launch("anchor_label_kernel");
}

extern "C" int i2v_anchor_disable(float* labels, const int* list, const int* positions, const int* n_disable, int batch,
                                  int total_anchors, int max_disable, cudaStream_t stream) {
    I2V_REQUIRE(batch >= 0 && total_anchors >= 0 && max_disable >= 0, "anchor_disable: bad shape");
    if (batch == 0 || max_disable == 0) return I2V_OK;
    I2V_REQUIRE(labels && list && positions && n_disable, "anchor_disable: null pointer");
    dim3 grid((unsigned)ceil_div(max_disable, 256), (unsigned)batch);
    anchor_disable_kernel<<<grid, 256, 0, stream>>>(labels, list, positions, n_disable, total_anchors, max_disable);
    return check_launch("anchor_disable_kernel");
}

extern "C" int i2v_anchor_targets_finalize(const float* base_anchors, const float* gt_boxes, int batch, int num_anchors,
                                           int height, int width, int feat_stride, int num_gt, const float* max_overlaps,
                                           const int* argmax, const float* labels, float inside_weight,
                                           float positive_weight, float negative_weight, float* labels_out,
                                           float* targets_out, float* inside_out, float* outside_out, cudaStream_t stream) {
    I2V_REQUIRE(batch >= 0 && num_anchors >= 1 && height >= 1 && width >= 1 && num_gt >= 1, "anchor_targets_finalize: bad shape");
    if (batch == 0) return I2V_OK;
    I2V_REQUIRE(base_anchors && gt_boxes && max_overlaps && argmax && labels && labels_out && targets_out && inside_out &&
                    outside_out,
                "anchor_targets_finalize: null pointer");
    const int total = num_anchors * height * width;
    dim3 grid((unsigned)ceil_div(total, 256), (unsigned)batch);
    anchor_finalize_kernel<<<grid, 256, 0, stream>>>(base_anchors, gt_boxes, total, num_anchors, height, width, feat_stride,
                                                    num_gt, max_overlaps, argmax, labels, inside_weight, positive_weight,
                                                    negative_weight, labels_out, targets_out, inside_out, outside_out);
    return check_launch("anchor_finalize_kernel");
}
