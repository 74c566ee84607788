/*
 * i2vsgg_b200.h -- C ABI of libi2vsgg_b200.so, the sm_100a implementation of the I2VSGG
 * region-level hot path (RPN proposal decode + NMS, RoIAlign / RoIPool forward+backward,
 * SGG pair enumeration + union boxes + dual masks, triplet top-k, embedding projection).
 *
 * Everything is `extern "C"`, takes plain device pointers + sizes + a cudaStream_t, allocates
 * nothing behind the caller's back (the i2v_* entry points take a caller-provided workspace;
 * only the five legacy-signature launchers keep a small grow-only cache because their
 * signatures have no workspace argument) and never calls exit().
 *
 * All paths below are relative to the reference tree (/root/reference).
 */
#ifndef I2VSGG_B200_H
#define I2VSGG_B200_H

#include <stddef.h>
#include <stdint.h>

#if !defined(__DRIVER_TYPES_H__) && !defined(__CUDA_RUNTIME_H__)
typedef struct CUstream_st* cudaStream_t; /* same opaque type the CUDA runtime declares */
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status ------------------------------------------------------------------------------- */
#define I2V_OK 0
#define I2V_ERR_INVALID 1     /* bad argument (shape, null pointer, unsupported size)            */
#define I2V_ERR_CUDA 2        /* a CUDA call or launch failed; see i2v_last_error()              */
#define I2V_ERR_WORKSPACE 3   /* workspace smaller than the matching *_workspace_bytes() result  */
#define I2V_ERR_UNSUPPORTED 4 /* requested implementation cannot handle this shape               */

/* Message of the last failure on the calling thread ("" if none). */
const char* i2v_last_error(void);
/* ABI version of this library (bumped when a signature changes). */
int i2v_abi_version(void);

/* ---- 1. Drop-in launchers: same names, argument order and meaning as the reference --------- */
/* Replaces lib/model/roi_align/src/roi_align_kernel.h:13-17 (impl. roi_align_kernel.cu:73-91).
 * Returns 1 on success like the reference; 0 (instead of exit(-1), roi_align_kernel.cu:84-88) on error.
 * top_data need not be pre-zeroed. */
int ROIAlignForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois,
                           const int height, const int width, const int channels, const int aligned_height,
                           const int aligned_width, const float* bottom_rois, float* top_data,
                           cudaStream_t stream);
/* Replaces roi_align_kernel.h:24-27 (impl. roi_align_kernel.cu:145-162).  Like the reference it
 * ACCUMULATES into bottom_diff, which the caller zeroes first (functions/roi_align.py:42-43). */
int ROIAlignBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size,
                            const int num_rois, const int height, const int width, const int channels,
                            const int aligned_height, const int aligned_width, const float* bottom_rois,
                            float* bottom_diff, cudaStream_t stream);
/* Replaces lib/model/roi_pooling/src/roi_pooling_kernel.h:8-12 (impl. roi_pooling_kernel.cu:95-125).
 * argmax_data may be NULL (roi_pooling_kernel.cu:89-90). */
int ROIPoolForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois,
                          const int height, const int width, const int channels, const int pooled_height,
                          const int pooled_width, const float* bottom_rois, float* top_data, int* argmax_data,
                          cudaStream_t stream);
/* Replaces roi_pooling_kernel.h:15-18 (impl. roi_pooling_kernel.cu:205-234): overwrites bottom_diff. */
int ROIPoolBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size,
                           const int num_rois, const int height, const int width, const int channels,
                           const int pooled_height, const int pooled_width, const float* bottom_rois,
                           float* bottom_diff, const int* argmax_data, cudaStream_t stream);
/* Replaces lib/model/nms/src/nms_cuda_kernel.h:5-6 (impl. nms_cuda_kernel.cu:87-161).
 * boxes_host: boxes_num x boxes_dim floats (x1,y1,x2,y2[,score]) in descending score order, host OR
 * device memory; keep_out / num_out are DEVICE pointers (nms_cuda_kernel.cu:147,154).  Synchronous on
 * the default stream like the reference, but the greedy scan runs on the device. */
void nms_cuda_compute(int* keep_out, int* num_out, float* boxes_host, int boxes_num, int boxes_dim,
                      float nms_overlap_thresh);

/* ---- 2. Lattice RoIAlign with the 2x2/stride-1 pool of RoIAlignAvg / RoIAlignMax fused ----- */
#define I2V_POOL_NONE 0 /* RoIAlign     (modules/roi_align.py:6-16):  lattice = pooled_h x pooled_w       */
#define I2V_POOL_AVG 1  /* RoIAlignAvg  (modules/roi_align.py:18-29): lattice (ph+1)x(pw+1), avg_pool 2/1 */
#define I2V_POOL_MAX 2  /* RoIAlignMax  (modules/roi_align.py:31-42): lattice (ph+1)x(pw+1), max_pool 2/1 */

#define I2V_IMPL_AUTO 0    /* plane-resident kernel when the shape allows it, else the gather kernel */
#define I2V_IMPL_GATHER 1  /* one thread per output element, L1/L2 gathers (any shape)                */
#define I2V_IMPL_PLANE 2   /* frame plane staged in shared memory; I2V_ERR_UNSUPPORTED if it cannot    */
#define I2V_IMPL_ROWS 3    /* backward only: plane-resident, warps own feature rows (bank-conflict free by
                              construction; forward treats it like I2V_IMPL_PLANE)                                 */
#define I2V_IMPL_PHASE 4   /* backward only: plane-resident, warp = lattice row, lanes = (cell of the bilinear pair,
                              channel); conflict-free, CTA barrier between feature-row phases (forward: like PLANE)    */
#define I2V_IMPL_BAND 5    /* backward only: lanes = 32 channels, a CTA owns a slab of feature rows of (frame, 32
                              channels), each warp a band of those rows: no ordering between warps (forward: like PLANE) */
#define I2V_IMPL_SLAB 6    /* forward only: the 16 planes of a CTA arrive as one TMA bulk copy and are used in place as
                              [channel][row][col] (what AUTO picks for 38x63 / 63x38 maps); I2V_ERR_UNSUPPORTED elsewhere
                              and in the backward                                                                          */
#define I2V_IMPL_EVEN 7    /* forward only: the planes re-pitched in place to an even row pitch, so that the two half-warps
                              of a load are on opposite bank parities by construction (W <= 64, H <= 40; measured 7 %
                              slower than I2V_IMPL_SLAB on config 2); I2V_ERR_UNSUPPORTED elsewhere and in the backward      */
#define I2V_IMPL_CHAN 8    /* forward only: lanes = 32 channels, planes cell-major in two row slabs per (frame, 32 channels);
                              a pooled row whose lattice rows lie in different slabs gets one red.add from either side
                              (C % 32 == 0, W <= 64); I2V_ERR_UNSUPPORTED elsewhere and in the backward                     */

size_t i2v_roi_align_workspace_bytes(int batch, int num_rois);
/* features [B,C,H,W], rois [N,5] = (batch_idx,x1,y1,x2,y2) image px, out [N,C,ph,pw]; all fp32, device.
 * Semantics: functions/roi_align.py:15-35 + roi_align_kernel.cu:15-70 (+ the module's pool).
 * RoIs whose batch index is outside [0,B) produce zeros. */
int i2v_roi_align_forward(const float* features, const float* rois, float* out, int batch, int channels,
                          int height, int width, int num_rois, int pooled_h, int pooled_w, float spatial_scale,
                          int pool_mode, int impl, void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* grad_out [N,C,ph,pw] -> grad_in [B,C,H,W], fully OVERWRITTEN (no pre-zeroing needed).
 * Semantics: functions/roi_align.py:37-51 + roi_align_kernel.cu:94-143 behind the pool's backward.
 * `features` is read only for I2V_POOL_MAX (arg-max routing) and may be NULL otherwise. */
/* `rois` may be NULL when `workspace` still holds what i2v_roi_align_forward left there for the same RoIs, batch, map and
 * pooled size (plane-resident forward kernels; a training step runs the two calls back to back): the backward then skips
 * its own table and list kernels.  Phased kernel only (impl AUTO / PHASE where it applies), I2V_ERR_UNSUPPORTED otherwise. */
int i2v_roi_align_backward(const float* grad_out, const float* features, const float* rois, float* grad_in,
                           int batch, int channels, int height, int width, int num_rois, int pooled_h,
                           int pooled_w, float spatial_scale, int pool_mode, int impl, void* workspace,
                           size_t workspace_bytes, cudaStream_t stream);

/* ---- 3. RoIPool ---------------------------------------------------------------------------- */
#define I2V_ARGMAX_FLAT 0  /* cffi op: index into the whole [B,C,H,W] tensor (roi_pooling_kernel.cu:85) */
#define I2V_ARGMAX_PLANE 1 /* model._C op: h*W+w within the (b,c) plane (roi_layers/roi_pool.py:17-19)  */
int i2v_roi_pool_forward(const float* features, const float* rois, float* out, int* argmax, int batch,
                         int channels, int height, int width, int num_rois, int pooled_h, int pooled_w,
                         float spatial_scale, int argmax_mode, cudaStream_t stream);
/* grad_in [B,C,H,W] is overwritten.  Scatter of grad_out to the recorded arg-max cells. */
int i2v_roi_pool_backward(const float* grad_out, const float* rois, const int* argmax, float* grad_in, int batch,
                          int channels, int height, int width, int num_rois, int pooled_h, int pooled_w,
                          float spatial_scale, int argmax_mode, cudaStream_t stream);

/* ---- 4. The RoIAlign behind model._C (roi_layers/roi_align.py:20,31-42; Mask R-CNN, aligned=False) */
int i2v_c_roi_align_forward(const float* features, const float* rois, float* out, int batch, int channels,
                            int height, int width, int num_rois, int pooled_h, int pooled_w, float spatial_scale,
                            int sampling_ratio, cudaStream_t stream);
int i2v_c_roi_align_backward(const float* grad_out, const float* rois, float* grad_in, int batch, int channels,
                             int height, int width, int num_rois, int pooled_h, int pooled_w, float spatial_scale,
                             int sampling_ratio, cudaStream_t stream);

/* ---- 4b. roi_crop: the bilinear sampler of lib/model/roi_crop (functions/roi_crop.py:8-24) ---------------------- */
/* Replace lib/model/roi_crop/src/roi_crop_cuda_kernel.h:6-34 (impl. roi_crop_cuda_kernel.cu:200-335), argument for
 * argument as roi_crop_cuda.c:21-44,58-95 passes them: sizes, then per tensor the data pointer and its strides in
 * elements (batch, channel, height, width; grids: batch, yx, height, width).  input [B,C,H,W]; grids [N,oh,ow,2] =
 * (y, x) in [-1,1]; RoI n samples frame n / (N / B).  Return 1 = ok, 0 = error (the reference prints and returns 0).
 * The backward ADDS into the caller-zeroed gradInputImages and leaves gradGrids untouched, like the reference
 * (roi_crop_cuda_kernel.cu:155-193 computes the grid dot products and drops them). */
int BilinearSamplerBHWD_updateOutput_cuda_kernel(int oc, int ow, int oh, int ob, int ic, int ih, int iw, int ib,
                                                 float* inputImages, int isb, int isc, int ish, int isw, float* grids,
                                                 int gsb, int gsc, int gsh, int gsw, float* output, int osb, int osc,
                                                 int osh, int osw, cudaStream_t stream);
int BilinearSamplerBHWD_updateGradInput_cuda_kernel(int goc, int gow, int goh, int gob, int ic, int ih, int iw, int ib,
                                                    float* inputImages, int isb, int isc, int ish, int isw, float* grids,
                                                    int gsb, int gsc, int gsh, int gsw, float* gradInputImages, int gisb,
                                                    int gisc, int gish, int gisw, float* gradGrids, int ggsb, int ggsc,
                                                    int ggsh, int ggsw, float* gradOutput, int gosb, int gosc, int gosh,
                                                    int gosw, cudaStream_t stream);
/* The same on contiguous tensors with status codes: out [N,C,oh,ow]; grad_in [B,C,H,W] is OVERWRITTEN. */
int i2v_roi_crop_forward(const float* features, const float* grids, float* out, int batch, int channels, int height,
                         int width, int num_rois, int out_h, int out_w, cudaStream_t stream);
int i2v_roi_crop_backward(const float* grad_out, const float* grids, float* grad_in, int batch, int channels, int height,
                          int width, int num_rois, int out_h, int out_w, cudaStream_t stream);

/* ---- 5. NMS (nms_wrapper.py:13-21 -> nms_cpu.py:6-34; bit-exact keep lists) ------------------ */
size_t i2v_nms_workspace_bytes(int batch, int num_boxes);
/* `batch` independent sets of `num_boxes` rows (row stride `box_stride` floats, x1,y1,x2,y2 first), each
 * already in descending score order.  keep_out [batch, keep_stride] int32 row indices in keep order,
 * num_out [batch].  max_keep > 0 stops after that many survivors (proposal_layer.py:153-154). */
int i2v_nms_sorted(const float* boxes, int batch, int num_boxes, int box_stride, float thresh, int max_keep,
                   int* keep_out, int keep_stride, int* num_out, void* workspace, size_t workspace_bytes,
                   cudaStream_t stream);
/* dets [N,5] (x1,y1,x2,y2,score) in ANY order: sorts by score descending (ties: lower index first), runs the
 * greedy scan and returns ORIGINAL row indices, i.e. what nms_wrapper.nms() returns. */
size_t i2v_nms_dets_workspace_bytes(int num_boxes);
int i2v_nms_dets(const float* dets, int num_boxes, float thresh, int* keep_out, int* num_out, void* workspace,
                 size_t workspace_bytes, cudaStream_t stream);

/* ---- 6. Proposal layer (proposal_layer.py:49-163 + bbox_transform.py:77-103,125-133) -------- */
size_t i2v_proposal_workspace_bytes(int batch, int num_anchors, int height, int width, int pre_nms_top_n);
/* cls_prob [B,2A,H,W], bbox_pred [B,4A,H,W], im_info [B,3]=(h,w,scale), base_anchors [A,4] (device).
 * out_rois [B,post_nms_top_n,5] = (b,x1,y1,x2,y2), zero padded; out_counts [B] (may be NULL). */
int i2v_proposal_forward(const float* cls_prob, const float* bbox_pred, const float* im_info,
                         const float* base_anchors, int batch, int num_anchors, int height, int width,
                         int feat_stride, int pre_nms_top_n, int post_nms_top_n, float nms_thresh,
                         float* out_rois, int* out_counts, void* workspace, size_t workspace_bytes,
                         cudaStream_t stream);
/* The same for a chunk of a larger batch: column 0 of out_rois counts from frame_base. */
int i2v_proposal_forward_chunk(const float* cls_prob, const float* bbox_pred, const float* im_info,
                               const float* base_anchors, int batch, int num_anchors, int height, int width,
                               int feat_stride, int pre_nms_top_n, int post_nms_top_n, float nms_thresh, int frame_base,
                               float* out_rois, int* out_counts, void* workspace, size_t workspace_bytes,
                               cudaStream_t stream);
/* rois [N,5]: column 0 += offset (a chunk processed with chunk-local frame numbers gets the batch's numbering back). */
int i2v_rois_add_frame(float* rois, int num_rois, int offset, cudaStream_t stream);
/* The step before the layer (rpn.py:63-78): rpn_cls_score [B,2A,H,W] -> rpn_cls_prob [B,2A,H,W], the softmax over each
 * anchor's (background, foreground) pair of channels (a, a+A). */
int i2v_rpn_cls_prob(const float* cls_score, float* cls_prob, int batch, int num_anchors, int height, int width,
                     cudaStream_t stream);
/* i2v_proposal_forward fed with the raw scores: softmax and foreground slice fused into the decode kernel; the result
 * equals i2v_rpn_cls_prob followed by i2v_proposal_forward bit for bit. */
int i2v_proposal_forward_scores(const float* cls_score, const float* bbox_pred, const float* im_info,
                                const float* base_anchors, int batch, int num_anchors, int height, int width,
                                int feat_stride, int pre_nms_top_n, int post_nms_top_n, float nms_thresh,
                                float* out_rois, int* out_counts, void* workspace, size_t workspace_bytes,
                                cudaStream_t stream);
/* Stage outputs for tests: decoded+clipped boxes [B,KA,4], scores [B,KA] in anchor-major order, and (after
 * the sort) order [B,KA] int32 = candidate indices by descending score.  Any of the three may be NULL. */
int i2v_proposal_stages(const float* cls_prob, const float* bbox_pred, const float* im_info,
                        const float* base_anchors, int batch, int num_anchors, int height, int width,
                        int feat_stride, float* boxes, float* scores, int* order, void* workspace,
                        size_t workspace_bytes, cudaStream_t stream);

/* ---- 7. SGG pair stage (faster_rcnn_SGG_emb.py:597-606,649-656; resnet_SGG_emb.py:240-256) -- */
/* boxes [N,4] fp32 -> ixs, ixo [P] int64, rel_boxes [P,5] fp32 (col 0 = 0), masks [P,2,32,32] fp32,
 * P = N*(N-1).  masks may be NULL. */
int i2v_pair_build(const float* boxes, int num_boxes, float im_h, float im_w, float margin, int64_t* ixs,
                   int64_t* ixo, float* rel_boxes, float* masks, cudaStream_t stream);

/* The pair stage of a frame GROUP in one launch (the loops of faster_rcnn_SGG_emb.py:597-606,649-656 for every frame):
 * boxes [F,N,4] (16-byte aligned) -> ixs, ixo [F*P] int64 holding GROUP rows f*N+i, rel_boxes [F*P,5] with the frame
 * number f in column 0, obj_masks [F*N,32,32] fp32 = the mask of resnet_SGG_emb.py:246-256 once per OBJECT (a pair's
 * two mask channels are its subject's and its object's mask).  Any output may be NULL. */
int i2v_pair_build_frames(const float* boxes, int frames, int num_boxes, float im_h, float im_w, float margin,
                          int64_t* ixs, int64_t* ixo, float* rel_boxes, float* obj_masks, cudaStream_t stream);

/* ---- 8. Triplet top-k (lib/utils.py:609-626) ------------------------------------------------ */
size_t i2v_triplet_topk_workspace_bytes(int num_pairs, int num_rel);
/* rel_score [P,R]; conf [N]; classes [N] int64; boxes [N,4]; ixs/ixo [P] int64.
 * record_out [top_k, 13] fp32 = (conf, cls_s, rel, cls_o, sub box x4, obj box x4, pair idx), rows past
 * *count_out are zero.  Order: descending score, ties by lower flat index p*R+r. */
int i2v_triplet_topk(const float* rel_score, const float* conf, const int64_t* classes, const float* boxes,
                     const int64_t* ixs, const int64_t* ixo, int num_pairs, int num_rel, int top_k,
                     float* record_out, int* count_out, void* workspace, size_t workspace_bytes,
                     cudaStream_t stream);

/* The same for F frames with the same number of detections in three launches: rel_score [F*P,R], conf / classes
 * [F,N], boxes [F,N,4]; ixs / ixo [P] are the FRAME-LOCAL pair lists (identical for every frame); record_out
 * [F,top_k,13], count_out [F]. */
size_t i2v_triplet_topk_frames_workspace_bytes(int frames, int num_pairs, int num_rel);
int i2v_triplet_topk_frames(const float* rel_score, const float* conf, const int64_t* classes, const float* boxes,
                            const int64_t* ixs, const int64_t* ixo, int frames, int num_boxes, int num_pairs,
                            int num_rel, int top_k, float* record_out, int* count_out, void* workspace,
                            size_t workspace_bytes, cudaStream_t stream);

/* ---- 9. Embedding projection of the SGG stage (resnet_SGG_emb.py:128-221) ---- */
#define I2V_DT_F32 0   /* fp32 storage                                                        */
#define I2V_DT_BF16 1  /* bf16 storage; tensor-core operands under tcgen05 kind::f16           */
#define I2V_DT_TF32 2  /* fp32 storage consumed by the tensor cores as tf32 (kind::tf32)       */
/* roi_pool(fmap, boxes).view(N, -1) of resnet_SGG_emb.py:144-146,158-160: the model._C RoIPool values (no
 * arg-max) written as rows [N, C*ph*pw] with pitch ldo (elements), fp32 or rounded once to bf16.  The bf16 flavour keeps
 * its shared-memory planes in bf16 (two channels per word; rounding is monotone, so the rows equal the fp32 kernel's bit
 * for bit); the environment variable I2V_POOL_F32_PLANES=1 forces fp32 planes (tests compare the two). */
int i2v_roi_pool_rows(const float* features, const float* rois, void* out, int batch, int channels, int height,
                      int width, int num_rois, int pooled_h, int pooled_w, float spatial_scale, long long ldo,
                      int out_dtype, cudaStream_t stream);
/* FC of lib/model/faster_rcnn/utils.py:48-60 (nn.Linear + optional ReLU): y[M,N] = act(x[M,K] . W[N,K]^T + bias[N]) on
 * tcgen05 tensor cores, fp32 accumulation in tensor memory.  in_dtype: I2V_DT_BF16 (x, W bf16) or I2V_DT_TF32
 * (x, W fp32); out_dtype: I2V_DT_F32 or I2V_DT_BF16.  ldx/ldw/ldy are row pitches in elements (x and W need
 * 16-byte aligned bases and pitches); bias may be NULL.  y may be a column slice of a wider buffer. */
int i2v_linear_forward(const void* x, const void* w, const float* bias, void* y, int M, int N, int K,
                       long long ldx, long long ldw, long long ldy, int in_dtype, int out_dtype, int relu,
                       cudaStream_t stream);
/* The same with F.dropout(y, training=True) (resnet_SGG_emb.py:148-151) applied behind the activation in the epilogue:
 * keep_mask [M,N] bytes (row pitch ldm >= N), y = keep ? y * scale : 0.  The caller draws the mask (torch's generator in
 * the Python mirror) and passes scale = 1 / (1 - p). */
int i2v_linear_forward_dropout(const void* x, const void* w, const float* bias, void* y, int M, int N, int K,
                               long long ldx, long long ldw, long long ldy, int in_dtype, int out_dtype, int relu,
                               const unsigned char* keep_mask, long long ldm, float scale, cudaStream_t stream);
/* fp32 -> bf16 (round to nearest even) of a [rows, cols] matrix; pitches in elements. */
int i2v_cast_bf16(const float* src, void* dst, long long rows, long long cols, long long lds, long long ldd,
                  cudaStream_t stream);
/* Patches of a convolution as bf16 rows (conv_lo, resnet_SGG_emb.py:107-110,182-185 -> one FC launch per layer):
 * out[(n*OH+oy)*OW+ox][(ky*KW+kx)*C+c] = in[n,c,oy*stride-pad+ky,ox*stride-pad+kx], zero outside and in the
 * pitch padding [KH*KW*C, ldo).  `in` (fp32 or bf16) is addressed by element strides, so NCHW and NHWC both fit. */
int i2v_im2col_bf16(const void* in, int in_dtype, int n, int channels, int height, int width, long long stride_n,
                    long long stride_c, long long stride_y, long long stride_x, int kernel_h, int kernel_w,
                    int stride, int pad, void* out, long long ldo, cudaStream_t stream);
/* The same patches as fp32 rows (the tf32 precision of the relation head). */
int i2v_im2col_f32(const void* in, int in_dtype, int n, int channels, int height, int width, long long stride_n,
                   long long stride_c, long long stride_y, long long stride_x, int kernel_h, int kernel_w, int stride,
                   int pad, float* out, long long ldo, cudaStream_t stream);
/* fp32 [rows, cols] -> the nearest tf32 value in an fp32 word (in place allowed).  tcgen05 kind::tf32 truncates its
 * operands; the tf32 path of the relation head rounds them first. */
int i2v_round_tf32(const float* src, float* dst, long long rows, long long cols, long long lds, long long ldd,
                   cudaStream_t stream);
/* cat(index_select(obj, 0, ixs), index_select(obj, 0, ixo), 1) of resnet_SGG_emb.py:150-151,169 as bf16 rows
 * [P, 2E] with pitch ldo; obj [N,E] fp32, ixs/ixo [P] int64. */
int i2v_pair_rows_bf16(const float* obj, const int64_t* ixs, const int64_t* ixo, void* out, int num_obj,
                       int num_pairs, int emb_dim, long long ldo, cudaStream_t stream);
/* Strided convolution as an implicit GEMM on tcgen05 (conv_lo's second layer, resnet_SGG_emb.py:108): x [N,H,W,C] bf16
 * NHWC is read through a 4-D TMA map with the convolution's element strides and zero-filled borders, so no patch matrix
 * is written.  w [O, K*K*Cp] bf16, taps in (ky, kx) order with each tap's channels padded to Cp = 64*ceil(C/64), row
 * pitch ldw; y [N*OH*OW, O] with pitch ldy = the next layer's NHWC input.  Needs OH*OW dividing 128, C % 8 == 0 and
 * O <= 128; I2V_ERR_UNSUPPORTED otherwise (use i2v_im2col_bf16 + i2v_linear_forward). */
int i2v_conv2d_nhwc_forward(const void* x, const void* w, const float* bias, void* y, int n, int height, int width,
                            int channels, int out_channels, int kernel, int stride, int pad, long long ldw,
                            long long ldy, int out_dtype, int relu, cudaStream_t stream);
/* The same layer for stride 2 on a PARITY-SPLIT activation: x [N, 2, 2, H/2, W/2, C] bf16, plane (y & 1, x & 1) holding
 * position (y >> 1, x >> 1) of the H x W map (`height`, `width` are the whole map's; both even, pad = (kernel - 1) / 2).
 * The samples of a filter tap are then a dense box of one plane, which the TMA fetches at full rate; the strided box of
 * i2v_conv2d_nhwc_forward makes it walk the skipped positions too.  Same results bit for bit.  The relation head's first
 * conv layer writes this layout directly (i2v_pair_conv1_split_bf16). */
int i2v_conv2d_nhwc_split_forward(const void* x, const void* w, const float* bias, void* y, int n, int height, int width,
                                  int channels, int out_channels, int kernel, int pad, long long ldw, long long ldy,
                                  int out_dtype, int relu, cudaStream_t stream);
/* First layer of conv_lo for ordered pairs (resnet_SGG_emb.py:107,182): a pair's two mask channels are the masks of
 * its subject and object, so conv(pair) = S[subject][.., 0:C] + S[object][.., C:2C] + bias, where obj_maps
 * [N, positions, 2C] fp32 holds the two single-channel convolutions of every OBJECT mask (one small FC launch).
 * out [P, positions, C] bf16 (NHWC), ReLU when `relu`. */
int i2v_pair_conv1_bf16(const float* obj_maps, const int64_t* ixs, const int64_t* ixo, const float* bias, void* out,
                        int num_obj, int num_pairs, int positions, int channels, int relu, cudaStream_t stream);
/* The same rows in the parity-split layout [P, 2, 2, oh/2, ow/2, C] (positions = oh x ow, both even) that
 * i2v_conv2d_nhwc_split_forward reads. */
int i2v_pair_conv1_split_bf16(const float* obj_maps, const int64_t* ixs, const int64_t* ixo, const float* bias, void* out,
                              int num_obj, int num_pairs, int oh, int ow, int channels, int relu, cudaStream_t stream);
/* out[p] = src[idx[p]] for bf16 rows (cols % 8 == 0; pitches in elements, multiples of 8).  Used to fan the rows computed
 * once per UNORDERED pair back out to both orderings: the union boxes of (i,j) and (j,i) are the same box
 * (resnet_SGG_emb.py:240-244 is symmetric), so their pooled / fc6 / fc7 / fc8 rows are identical. */
int i2v_gather_rows_bf16(const void* src, const int64_t* idx, void* out, int num_src, int num_out, int cols,
                         long long lds, long long ldo, cudaStream_t stream);
/* resnet_SGG_emb.py:207-219: scores[P,R] = softmax_R(normalize(x[P,E]) . normalize(prd[R,E])^T); the softmax is
 * the eval-mode branch (apply_softmax != 0).  fp32 throughout. */
size_t i2v_rel_scores_workspace_bytes(int num_rel, int emb_dim);
int i2v_rel_scores(const float* x, const float* prd, float* scores, int num_pairs, int num_rel, int emb_dim,
                   int apply_softmax, void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* ---- 10. Temporal association (lib/utils.py:134-182 greedy_relational_association) ---- */
/* records [F, top_k, 13] fp32 as written by i2v_triplet_topk (all-gathered into video order), counts [F] int32.
 * frame_numbers [F] ascending (NULL: 0..F-1); source_frame [F] (NULL: identity) lets a frame position use another
 * frame's records (-1: none), which is how lib/utils.py:470-518 fills frames without predictions.
 * For frame position f and its j-th prediction in descending confidence: order[f,j] = record row, rel_id[f,j] = the
 * relation (numbered in creation order) it started or extended, -1 past the frame's predictions.
 * rel_info [F*top_k, 6] int32 = (first frame number, end frame number, s, p, o, length), rel_score [F*top_k] double =
 * mean confidence (numpy's pairwise float64 sum), num_rel [1].  top_k <= 128; max_traj = predictions kept per frame. */
size_t i2v_association_workspace_bytes(int frames, int top_k);
int i2v_greedy_association(const float* records, const int* counts, const int* frame_numbers, const int* source_frame,
                           int frames, int top_k, int max_traj, int* rel_id, int* order, int* rel_info,
                           double* rel_score, int* num_rel, void* workspace, size_t workspace_bytes,
                           cudaStream_t stream);

/* ---- 11. Proposal targets (lib/model/rpn/proposal_target_layer_cascade.py:33-212) ---- */
/* bbox_overlaps_batch (bbox_transform.py:215-257) reduced on the fly: rois [B,R,roi_width] (roi_width 5: (b,x1,y1,x2,y2),
 * 4: (x1,y1,x2,y2)), gt_boxes [B,K,5] (x1,y1,x2,y2,label) -> max_overlaps [B,R], assignment [B,R] (first arg-max),
 * labels [B,R] = label of the assigned box (may be NULL).  Zero-area gt -> 0, zero-area roi -> -1, as there. */
int i2v_roi_gt_overlaps(const float* rois, int roi_width, const float* gt_boxes, int batch, int num_rois, int num_gt,
                        float* max_overlaps, int* assignment, float* labels, cudaStream_t stream);
/* Per image, in ascending order (torch.nonzero): fg_inds = {i: max >= fg_thresh}, bg_inds = {i: bg_lo <= max < bg_hi}
 * (both [B,R] int32, only the first counts[b][0] / counts[b][1] entries are written), counts [B,2]. */
int i2v_fg_bg_select(const float* max_overlaps, int batch, int num_rois, float fg_thresh, float bg_thresh_hi,
                     float bg_thresh_lo, int* fg_inds, int* bg_inds, int* counts, cudaStream_t stream);
/* positions [B,S] index INTO fg_inds (first fg_this[b] entries of a row) or bg_inds (the rest), as drawn by the caller
 * (the reference draws them with numpy, :151-180).  Outputs as _sample_rois_pytorch returns them: rois_out [B,S,5],
 * labels_out [B,S], targets_out / inside_out / outside_out [B,S,4] (bbox_transform_batch, optional normalisation by
 * means / stds -- host pointers to four floats each -- and BBOX_INSIDE_WEIGHTS; zero for background). */
int i2v_proposal_targets_gather(const float* rois, const float* gt_boxes, const int* assignment, const float* labels,
                                const int* fg_inds, const int* bg_inds, const int* positions, const int* fg_this,
                                int batch, int num_rois, int num_gt, int rois_per_image, const float* means,
                                const float* stds, const float* inside_weights, int normalize, float* rois_out,
                                float* labels_out, float* targets_out, float* inside_out, float* outside_out,
                                cudaStream_t stream);

/* ---- 12. Anchor targets (lib/model/rpn/anchor_target_layer.py:48-193) ---- */
/* Anchor j = (y*W + x)*A + a over the [height, width] RPN map; "inside" anchors lie within [-border, im_w + border) x
 * [-border, im_h + border) (:81-84, the first image's size for the whole batch).  gt_boxes [B,G,5].
 * max_overlaps [B,KA] (-2 for outside anchors), argmax [B,KA], gt_max_bits [B,G] (float bits of the per-box maxima). */
int i2v_anchor_overlaps(const float* base_anchors, const float* gt_boxes, int batch, int num_anchors, int height,
                        int width, int feat_stride, int num_gt, float allowed_border, float im_w, float im_h,
                        float* max_overlaps, int* argmax, int* gt_max_bits, cudaStream_t stream);
/* The label rules of :103-119 -> labels [B,KA] (1 / 0 / -1 as floats). */
int i2v_anchor_labels(const float* base_anchors, const float* gt_boxes, int batch, int num_anchors, int height, int width,
                      int feat_stride, int num_gt, const float* max_overlaps, const int* gt_max_bits,
                      float negative_overlap, float positive_overlap, int clobber_positives, float* labels,
                      cudaStream_t stream);
/* labels[b][list[b][positions[b][k]]] = -1 for k < n_disable[b] (:131-145; `list` as written by i2v_fg_bg_select on the
 * label array, positions [B,max_disable] drawn by the caller with numpy like the reference). */
int i2v_anchor_disable(float* labels, const int* list, const int* positions, const int* n_disable, int batch,
                       int total_anchors, int max_disable, cudaStream_t stream);
/* Regression targets, weights and the output layouts of :150-191: labels_out [B,1,A*H,W], targets_out / inside_out /
 * outside_out [B,4A,H,W]. */
int i2v_anchor_targets_finalize(const float* base_anchors, const float* gt_boxes, int batch, int num_anchors, int height,
                                int width, int feat_stride, int num_gt, const float* max_overlaps, const int* argmax,
                                const float* labels, float inside_weight, float positive_weight, float negative_weight,
                                float* labels_out, float* targets_out, float* inside_out, float* outside_out,
                                cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* I2VSGG_B200_H */
