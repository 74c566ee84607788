// roi_crop: the bilinear sampler of lib/model/roi_crop (the third pooling mode of the original API) for sm_100a.
//
// Reference semantics: lib/model/roi_crop/src/roi_crop_cuda_kernel.cu:47-108 (forward), :111-195 (backward), as called
// by roi_crop_cuda.c:14-44 / functions/roi_crop.py:8-24:
//   input  [B,C,H,W], grids [N,oh,ow,2] = (y, x) in [-1, 1], output [N,C,oh,ow]; RoI n samples frame n / (N / B);
//   coord = (g + 1) * (extent - 1) / 2, corner = floor(coord), weight of the top-left corner = 1 - (coord - corner);
//   corners off the map contribute zero; an output whose four corners are all off the map keeps its zero;
//   backward: d input = scatter of weight * d output to the corners in range; the reference computes the four
//   grid dot products and drops them (`gradGrids` is never written), so d grids = 0.
//
// Forward: one thread per (RoI, output cell, 4 channels): the sample position, corners and weights are evaluated once
// and used for four planes; consecutive threads take consecutive output cells, so output stores are coalesced and the
// gathers of a warp fall into a few feature rows.  Backward: the same walk with red.global.add.f32 into a zeroed
// gradient (the op is dead code in both reference models -- SURVEY 8(f) rank 4 -- so it keeps the simple scatter).
#include "common.cuh"

namespace i2v {
namespace {

struct Corner {
    int x, y;          // top-left corner
    float wx, wy;      // weight of the left column / upper row
    bool tl, tr, bl, br;
};

// roi_crop_cuda_kernel.cu:11-22 with the operation order of the source: ((g + 1) * (extent - 1)) / 2 in fp32
__device__ __forceinline__ void top_left(float g, int extent, int& point, float& weight) {
    const float coord = __fdiv_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)(extent - 1)), 2.f);
    const float fl = floorf(coord);
    point = (int)fl;
    weight = __fsub_rn(1.f, __fsub_rn(coord, fl));
}

__device__ __forceinline__ Corner corner_of(const float* __restrict__ grid, int H, int W) {
    Corner c;
    top_left(__ldg(grid + 1), W, c.x, c.wx);
    top_left(__ldg(grid), H, c.y, c.wy);
    const bool x0 = c.x >= 0 && c.x <= W - 1, x1 = c.x + 1 >= 0 && c.x + 1 <= W - 1;
    const bool y0 = c.y >= 0 && c.y <= H - 1, y1 = c.y + 1 >= 0 && c.y + 1 <= H - 1;
    c.tl = x0 && y0;
    c.tr = x1 && y0;
    c.bl = x0 && y1;
    c.br = x1 && y1;
    return c;
}

constexpr int kCh = 4;     // channels per thread

__global__ void __launch_bounds__(256) roi_crop_forward_kernel(const float* __restrict__ in, const float* __restrict__ grids,
                                                               float* __restrict__ out, int64_t total, int C, int H, int W,
                                                               int oh, int ow, int rois_per_image, int64_t isb, int64_t isc,
                                                               int64_t ish, int64_t isw, int64_t gsb, int64_t gsh,
                                                               int64_t gsw, int64_t osb, int64_t osc, int64_t osh) {
    const int cgroups = (C + kCh - 1) / kCh;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int xo = (int)(i % ow);
        const int yo = (int)((i / ow) % oh);
        const int cg = (int)((i / ((int64_t)ow * oh)) % cgroups);
        const int n = (int)(i / ((int64_t)ow * oh * cgroups));
        const Corner c = corner_of(grids + n * gsb + yo * gsh + xo * gsw, H, W);
        const bool any = c.tl || c.tr || c.bl || c.br;
        const float w_tl = __fmul_rn(c.wx, c.wy), w_tr = __fmul_rn(__fsub_rn(1.f, c.wx), c.wy);
        const float w_bl = __fmul_rn(c.wx, __fsub_rn(1.f, c.wy)), w_br = __fmul_rn(__fsub_rn(1.f, c.wx), __fsub_rn(1.f, c.wy));
        const float* base = in + (int64_t)(n / rois_per_image) * isb + (int64_t)c.y * ish + (int64_t)c.x * isw;
        float* o = out + (int64_t)n * osb + (int64_t)yo * osh + xo;
#pragma unroll
        for (int k = 0; k < kCh; ++k) {
            const int ch = cg * kCh + k;
            if (ch >= C) break;
            const float* p = base + (int64_t)ch * isc;
            float v = 0.f;
            if (any) {
                const float a = c.tl ? __ldg(p) : 0.f, b = c.tr ? __ldg(p + isw) : 0.f;
                const float d = c.bl ? __ldg(p + ish) : 0.f, e = c.br ? __ldg(p + ish + isw) : 0.f;
                // roi_crop_cuda_kernel.cu:101-104, left to right, every product and sum rounded to fp32
                v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w_tl, a), __fmul_rn(w_tr, b)), __fmul_rn(w_bl, d)),
                              __fmul_rn(w_br, e));
            }
            o[(int64_t)ch * osc] = v;          // the reference pre-zeroes the output (functions/roi_crop.py:11)
        }
    }
}

__global__ void __launch_bounds__(256) roi_crop_backward_kernel(const float* __restrict__ grad_out,
                                                                const float* __restrict__ grids, float* __restrict__ grad_in,
                                                                int64_t total, int C, int H, int W, int oh, int ow,
                                                                int rois_per_image, int64_t gisb, int64_t gisc, int64_t gish,
                                                                int64_t gisw, int64_t gsb, int64_t gsh, int64_t gsw,
                                                                int64_t gosb, int64_t gosc, int64_t gosh) {
    const int cgroups = (C + kCh - 1) / kCh;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int xo = (int)(i % ow);
        const int yo = (int)((i / ow) % oh);
        const int cg = (int)((i / ((int64_t)ow * oh)) % cgroups);
        const int n = (int)(i / ((int64_t)ow * oh * cgroups));
        const Corner c = corner_of(grids + n * gsb + yo * gsh + xo * gsw, H, W);
        if (!(c.tl || c.tr || c.bl || c.br)) continue;
        const float w_tl = __fmul_rn(c.wx, c.wy), w_tr = __fmul_rn(__fsub_rn(1.f, c.wx), c.wy);
        const float w_bl = __fmul_rn(c.wx, __fsub_rn(1.f, c.wy)), w_br = __fmul_rn(__fsub_rn(1.f, c.wx), __fsub_rn(1.f, c.wy));
        float* base = grad_in + (int64_t)(n / rois_per_image) * gisb + (int64_t)c.y * gish + (int64_t)c.x * gisw;
        const float* g = grad_out + (int64_t)n * gosb + (int64_t)yo * gosh + xo;
#pragma unroll
        for (int k = 0; k < kCh; ++k) {
            const int ch = cg * kCh + k;
            if (ch >= C) break;
            const float go = __ldg(g + (int64_t)ch * gosc);
            float* p = base + (int64_t)ch * gisc;
            if (c.tl) atomicAdd(p, __fmul_rn(w_tl, go));
            if (c.tr) atomicAdd(p + gisw, __fmul_rn(w_tr, go));
            if (c.bl) atomicAdd(p + gish, __fmul_rn(w_bl, go));
            if (c.br) atomicAdd(p + gish + gisw, __fmul_rn(w_br, go));
        }
    }
}

int crop_sizes_ok(const char* who, int ob, int oc, int oh, int ow, int ib, int ic, int ih, int iw) {
    I2V_REQUIRE(ob >= 0 && oc >= 0 && oh >= 0 && ow >= 0 && ib >= 0 && ic >= 0 && ih >= 0 && iw >= 0, "%s: negative size", who);
    I2V_REQUIRE(oc == ic, "%s: %d output channels for %d input channels", who, oc, ic);
    I2V_REQUIRE(ob == 0 || (ib > 0 && ob / ib > 0), "%s: %d RoIs for %d frames (roiPerImage = N / B must be positive)", who, ob, ib);
    return I2V_OK;
}

}  // namespace
}  // namespace i2v

using namespace i2v;

// ---- the reference's launchers, argument for argument (roi_crop_cuda_kernel.h:6-34); strides in elements.  The output /
// gradient strides along the width are taken as 1 like roi_crop_cuda_kernel.cu:77,150 do. ----
extern "C" int BilinearSamplerBHWD_updateOutput_cuda_kernel(int oc, int ow, int oh, int ob, int ic, int ih, int iw, int ib,
                                                            float* inputImages, int isb, int isc, int ish, int isw,
                                                            float* grids, int gsb, int gsc, int gsh, int gsw, float* output,
                                                            int osb, int osc, int osh, int osw, cudaStream_t stream) {
    (void)gsc;
    (void)osw;
    if (crop_sizes_ok("roi_crop forward", ob, oc, oh, ow, ib, ic, ih, iw) != I2V_OK) return 0;
    const int64_t total = (int64_t)ob * ((oc + kCh - 1) / kCh) * oh * ow;
    if (total == 0) return 1;
    if (!inputImages || !grids || !output) {
        set_error("roi_crop forward: null pointer");
        return 0;
    }
    roi_crop_forward_kernel<<<grid_for(total, 256), 256, 0, stream>>>(inputImages, grids, output, total, oc, ih, iw, oh, ow,
                                                                     ob / ib, isb, isc, ish, isw, gsb, gsh, gsw, osb, osc, osh);
    return check_launch("roi_crop_forward_kernel") == I2V_OK ? 1 : 0;
}

extern "C" int BilinearSamplerBHWD_updateGradInput_cuda_kernel(int goc, int gow, int goh, int gob, int ic, int ih, int iw, int ib,
                                                               float* inputImages, int isb, int isc, int ish, int isw,
                                                               float* grids, int gsb, int gsc, int gsh, int gsw,
                                                               float* gradInputImages, int gisb, int gisc, int gish, int gisw,
                                                               float* gradGrids, int ggsb, int ggsc, int ggsh, int ggsw,
                                                               float* gradOutput, int gosb, int gosc, int gosh, int gosw,
                                                               cudaStream_t stream) {
    // the features only feed the grid dot products, which the reference computes and drops (roi_crop_cuda_kernel.cu:155-193)
    (void)inputImages; (void)isb; (void)isc; (void)ish; (void)isw; (void)gsc; (void)gosw;
    (void)gradGrids; (void)ggsb; (void)ggsc; (void)ggsh; (void)ggsw;
    if (crop_sizes_ok("roi_crop backward", gob, goc, goh, gow, ib, ic, ih, iw) != I2V_OK) return 0;
    const int64_t total = (int64_t)gob * ((goc + kCh - 1) / kCh) * goh * gow;
    if (total == 0) return 1;
    if (!grids || !gradInputImages || !gradOutput) {
        set_error("roi_crop backward: null pointer");
        return 0;
    }
    roi_crop_backward_kernel<<<grid_for(total, 256), 256, 0, stream>>>(gradOutput, grids, gradInputImages, total, goc, ih, iw,
                                                                      goh, gow, gob / ib, gisb, gisc, gish, gisw, gsb, gsh, gsw,
                                                                      gosb, gosc, gosh);
    return check_launch("roi_crop_backward_kernel") == I2V_OK ? 1 : 0;
}

// ---- contiguous tensors ----
extern "C" int i2v_roi_crop_forward(const float* features, const float* grids, float* out, int batch, int channels, int height,
                                    int width, int num_rois, int out_h, int out_w, cudaStream_t stream) {
    I2V_TRY(crop_sizes_ok("roi_crop_forward", num_rois, channels, out_h, out_w, batch, channels, height, width));
    const int hw = height * width, ohw = out_h * out_w;
    int rc = BilinearSamplerBHWD_updateOutput_cuda_kernel(channels, out_w, out_h, num_rois, channels, height, width, batch,
                                                          const_cast<float*>(features), channels * hw, hw, width, 1,
                                                          const_cast<float*>(grids), ohw * 2, 1, out_w * 2, 2, out,
                                                          channels * ohw, ohw, out_w, 1, stream);
    return rc == 1 ? I2V_OK : I2V_ERR_CUDA;
}

extern "C" int i2v_roi_crop_backward(const float* grad_out, const float* grids, float* grad_in, int batch, int channels,
                                     int height, int width, int num_rois, int out_h, int out_w, cudaStream_t stream) {
    I2V_TRY(crop_sizes_ok("roi_crop_backward", num_rois, channels, out_h, out_w, batch, channels, height, width));
    const size_t in_elems = (size_t)batch * channels * height * width;
    if (in_elems == 0) return I2V_OK;
    I2V_REQUIRE(grad_in, "roi_crop_backward: null grad_in");
    I2V_CUDA_TRY(cudaMemsetAsync(grad_in, 0, in_elems * sizeof(float), stream));
    const int hw = height * width, ohw = out_h * out_w;
    int rc = BilinearSamplerBHWD_updateGradInput_cuda_kernel(
        channels, out_w, out_h, num_rois, channels, height, width, batch, nullptr, 0, 0, 0, 0, const_cast<float*>(grids), ohw * 2, 1,
        out_w * 2, 2, grad_in, channels * hw, hw, width, 1, nullptr, 0, 0, 0, 0, const_cast<float*>(grad_out), channels * ohw, ohw,
        out_w, 1, stream);
    return rc == 1 ? I2V_OK : I2V_ERR_CUDA;
}
