// RoIPool forward / backward for sm_100a, in the two flavours the reference uses:
//   I2V_ARGMAX_FLAT   the cffi op: lib/model/roi_pooling/src/roi_pooling_kernel.cu:24-93 (forward, argmax is an
//                     index into the whole [B,C,H,W] tensor) and :128-203 (backward);
//   I2V_ARGMAX_PLANE  the op behind model._C (lib/model/roi_layers/roi_pool.py:17-19,30-42; maskrcnn-benchmark
//                     ROIPool == torchvision.ops.roi_pool): argmax is h*W+w inside the (b,c) plane.
// plus the RoIAlign behind model._C (lib/model/roi_layers/roi_align.py:20,31-42; Mask R-CNN RoIAlign,
// aligned=False).
//
// Forward: one thread per pooled element (pw fastest, so a warp writes a contiguous run of the output and reads
// neighbouring bins of the same plane).  Backward: scatter of each pooled gradient to its recorded arg-max cell
// (O(N*C*P*P)) instead of the reference's O(B*C*H*W*N) scan; the cffi flavour applies the same feasibility
// tests as roi_pooling_kernel.cu:160-183 so degenerate RoIs drop out exactly as they do there.
#include <cuda_bf16.h>
#include <float.h>
#include <stdlib.h>

#include "common.cuh"

namespace i2v {

struct PoolGeom {
    int rs_w, rs_h, re_w, re_h;
    float bin_h, bin_w;
};

// roi_pooling_kernel.cu:44-53
__device__ __forceinline__ PoolGeom pool_geom(const float* __restrict__ r, float scale, int PH, int PW) {
    PoolGeom g;
    g.rs_w = (int)roundf(__fmul_rn(r[1], scale));
    g.rs_h = (int)roundf(__fmul_rn(r[2], scale));
    g.re_w = (int)roundf(__fmul_rn(r[3], scale));
    g.re_h = (int)roundf(__fmul_rn(r[4], scale));
    int rw = max(g.re_w - g.rs_w + 1, 1), rh = max(g.re_h - g.rs_h + 1, 1);
    g.bin_h = __fdiv_rn((float)rh, (float)PH);
    g.bin_w = __fdiv_rn((float)rw, (float)PW);
    return g;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

template <int MODE>
__global__ void __launch_bounds__(256) roi_pool_fwd_kernel(const float* __restrict__ feat,
                                                           const float* __restrict__ rois, float* __restrict__ out,
                                                           int* __restrict__ argmax, int64_t total, int batch, int C,
                                                           int H, int W, int PH, int PW, float scale) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int pw = (int)(idx % PW);
        int ph = (int)((idx / PW) % PH);
        int c = (int)((idx / ((int64_t)PW * PH)) % C);
        int n = (int)(idx / ((int64_t)PW * PH * C));
        const float* r = rois + (size_t)n * 5;
        int b = (int)r[0];
        float best = 0.f;
        int64_t bi = -1;
        if (b >= 0 && b < batch) {
            PoolGeom g = pool_geom(r, scale, PH, PW);
            int hs = (int)floorf(__fmul_rn((float)ph, g.bin_h)), ws = (int)floorf(__fmul_rn((float)pw, g.bin_w));
            int he = (int)ceilf(__fmul_rn((float)(ph + 1), g.bin_h)), we = (int)ceilf(__fmul_rn((float)(pw + 1), g.bin_w));
            hs = clampi(hs + g.rs_h, 0, H);
            he = clampi(he + g.rs_h, 0, H);
            ws = clampi(ws + g.rs_w, 0, W);
            we = clampi(we + g.rs_w, 0, W);
            bool empty = (he <= hs) || (we <= ws);
            best = empty ? 0.f : -FLT_MAX;
            const int64_t plane_off = ((int64_t)b * C + c) * H * W;
            const float* plane = feat + plane_off;
            int bl = -1;
            for (int h = hs; h < he; ++h)
                for (int w = ws; w < we; ++w) {
                    float v = __ldg(plane + h * W + w);
                    if (v > best) {
                        best = v;
                        bl = h * W + w;
                    }
                }
            if (bl >= 0) bi = (MODE == I2V_ARGMAX_FLAT) ? plane_off + bl : bl;
        }
        out[idx] = best;
        if (argmax) argmax[idx] = (int)bi;
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) roi_pool_bwd_kernel(const float* __restrict__ grad_out,
                                                           const float* __restrict__ rois,
                                                           const int* __restrict__ argmax, float* __restrict__ grad_in,
                                                           int64_t total, int batch, int C, int H, int W, int PH,
                                                           int PW, float scale) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int am = argmax[idx];
        if (am < 0) continue;
        int pw = (int)(idx % PW);
        int ph = (int)((idx / PW) % PH);
        int c = (int)((idx / ((int64_t)PW * PH)) % C);
        int n = (int)(idx / ((int64_t)PW * PH * C));
        const float* r = rois + (size_t)n * 5;
        int b = (int)r[0];
        if (b < 0 || b >= batch) continue;
        if (MODE == I2V_ARGMAX_PLANE) {
            atomicAdd(grad_in + ((size_t)b * C + c) * H * W + am, grad_out[idx]);
        } else {
            // roi_pooling_kernel.cu:143-183: the input element only collects from RoIs of its own frame that contain
            // it and from the pooled cells in its feasible window.
            int64_t plane_off = ((int64_t)b * C + c) * H * W;
            int64_t local = (int64_t)am - plane_off;
            if (local < 0 || local >= (int64_t)H * W) continue;  // argmax of another frame/channel never matches
            int h = (int)(local / W), w = (int)(local % W);
            PoolGeom g = pool_geom(r, scale, PH, PW);
            if (!(w >= g.rs_w && w <= g.re_w && h >= g.rs_h && h <= g.re_h)) continue;
            int p0 = (int)floorf(__fdiv_rn((float)(h - g.rs_h), g.bin_h));
            int p1 = (int)ceilf(__fdiv_rn((float)(h - g.rs_h + 1), g.bin_h));
            int q0 = (int)floorf(__fdiv_rn((float)(w - g.rs_w), g.bin_w));
            int q1 = (int)ceilf(__fdiv_rn((float)(w - g.rs_w + 1), g.bin_w));
            p0 = clampi(p0, 0, PH);
            p1 = clampi(p1, 0, PH);
            q0 = clampi(q0, 0, PW);
            q1 = clampi(q1, 0, PW);
            if (ph < p0 || ph >= p1 || pw < q0 || pw >= q1) continue;
            atomicAdd(grad_in + am, grad_out[idx]);
        }
    }
}

// The pooled rows the SGG projection consumes (resnet_SGG_emb.py:144-146,158-160): the model._C RoIPool value
// (no arg-max), flattened to [N, C*PH*PW] with row pitch `ldo`, as fp32 or rounded once to bf16 for the tensor cores.
__device__ __forceinline__ void store_pooled(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_pooled(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename OutT>
__global__ void __launch_bounds__(256) roi_pool_rows_kernel(const float* __restrict__ feat,
                                                            const float* __restrict__ rois, OutT* __restrict__ out,
                                                            int64_t total, int batch, int C, int H, int W, int PH,
                                                            int PW, float scale, int64_t ldo) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int pw = (int)(idx % PW);
        int ph = (int)((idx / PW) % PH);
        int c = (int)((idx / ((int64_t)PW * PH)) % C);
        int n = (int)(idx / ((int64_t)PW * PH * C));
        const float* r = rois + (size_t)n * 5;
        int b = (int)r[0];
        float best = 0.f;
        if (b >= 0 && b < batch) {
            PoolGeom g = pool_geom(r, scale, PH, PW);
            int hs = (int)floorf(__fmul_rn((float)ph, g.bin_h)), ws = (int)floorf(__fmul_rn((float)pw, g.bin_w));
            int he = (int)ceilf(__fmul_rn((float)(ph + 1), g.bin_h)), we = (int)ceilf(__fmul_rn((float)(pw + 1), g.bin_w));
            hs = clampi(hs + g.rs_h, 0, H);
            he = clampi(he + g.rs_h, 0, H);
            ws = clampi(ws + g.rs_w, 0, W);
            we = clampi(we + g.rs_w, 0, W);
            bool empty = (he <= hs) || (we <= ws);
            best = empty ? 0.f : -FLT_MAX;
            const float* plane = feat + ((int64_t)b * C + c) * H * W;
            for (int h = hs; h < he; ++h)
                for (int w = ws; w < we; ++w) {
                    float v = __ldg(plane + h * W + w);
                    if (v > best) best = v;
                }
        }
        store_pooled(out + (int64_t)n * ldo + (idx - (int64_t)n * C * PH * PW), best);
    }
}

// ---------------------------------------------------------------------------------- model._C RoIAlign
// Mask R-CNN RoIAlign (aligned=False): roi extent max(x2-x1, 1), bin = extent / P, ceil(extent / P) samples per
// bin axis unless sampling_ratio > 0, bilinear with the [-1, size] acceptance band and edge clamp, mean over samples.
struct CAlignGeom {
    float sw, sh, bh, bw;
    int gh, gw;
};
__device__ __forceinline__ CAlignGeom c_align_geom(const float* __restrict__ r, float scale, int PH, int PW,
                                                   int sampling_ratio) {
    CAlignGeom g;
    // separately rounded products and differences: the sample-grid size below is an integer decision
    g.sw = __fmul_rn(r[1], scale);
    g.sh = __fmul_rn(r[2], scale);
    float ew = __fmul_rn(r[3], scale), eh = __fmul_rn(r[4], scale);
    float rw = fmaxf(__fsub_rn(ew, g.sw), 1.f), rh = fmaxf(__fsub_rn(eh, g.sh), 1.f);
    g.bh = __fdiv_rn(rh, (float)PH);
    g.bw = __fdiv_rn(rw, (float)PW);
    g.gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(g.bh);
    g.gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(g.bw);
    return g;
}

struct Bilin {
    int y0, x0, y1, x1;
    float w1, w2, w3, w4;
    bool ok;
};
__device__ __forceinline__ Bilin bilin_setup(int H, int W, float y, float x) {
    Bilin q;
    q.ok = !(y < -1.0f || y > (float)H || x < -1.0f || x > (float)W);
    if (y <= 0) y = 0;
    if (x <= 0) x = 0;
    q.y0 = (int)y;
    q.x0 = (int)x;
    if (q.y0 >= H - 1) {
        q.y1 = q.y0 = H - 1;
        y = (float)q.y0;
    } else {
        q.y1 = q.y0 + 1;
    }
    if (q.x0 >= W - 1) {
        q.x1 = q.x0 = W - 1;
        x = (float)q.x0;
    } else {
        q.x1 = q.x0 + 1;
    }
    float ly = y - (float)q.y0, lx = x - (float)q.x0, hy = 1.f - ly, hx = 1.f - lx;
    q.w1 = hy * hx;
    q.w2 = hy * lx;
    q.w3 = ly * hx;
    q.w4 = ly * lx;
    return q;
}

__global__ void __launch_bounds__(256) c_roi_align_fwd_kernel(const float* __restrict__ feat,
                                                              const float* __restrict__ rois, float* __restrict__ out,
                                                              int64_t total, int batch, int C, int H, int W, int PH,
                                                              int PW, float scale, int sampling_ratio) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int pw = (int)(idx % PW);
        int ph = (int)((idx / PW) % PH);
        int c = (int)((idx / ((int64_t)PW * PH)) % C);
        int n = (int)(idx / ((int64_t)PW * PH * C));
        const float* r = rois + (size_t)n * 5;
        int b = (int)r[0];
        float res = 0.f;
        if (b >= 0 && b < batch) {
            CAlignGeom g = c_align_geom(r, scale, PH, PW, sampling_ratio);
            const float* plane = feat + ((size_t)b * C + c) * H * W;
            float acc = 0.f;
            for (int iy = 0; iy < g.gh; ++iy) {
                float y = g.sh + (float)ph * g.bh + ((float)iy + .5f) * g.bh / (float)g.gh;
                for (int ix = 0; ix < g.gw; ++ix) {
                    float x = g.sw + (float)pw * g.bw + ((float)ix + .5f) * g.bw / (float)g.gw;
                    Bilin q = bilin_setup(H, W, y, x);
                    if (q.ok)
                        acc += q.w1 * __ldg(plane + q.y0 * W + q.x0) + q.w2 * __ldg(plane + q.y0 * W + q.x1) +
                               q.w3 * __ldg(plane + q.y1 * W + q.x0) + q.w4 * __ldg(plane + q.y1 * W + q.x1);
                }
            }
            res = acc / (float)(g.gh * g.gw);
        }
        out[idx] = res;
    }
}

__global__ void __launch_bounds__(256) c_roi_align_bwd_kernel(const float* __restrict__ grad_out,
                                                              const float* __restrict__ rois,
                                                              float* __restrict__ grad_in, int64_t total, int batch,
                                                              int C, int H, int W, int PH, int PW, float scale,
                                                              int sampling_ratio) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int pw = (int)(idx % PW);
        int ph = (int)((idx / PW) % PH);
        int c = (int)((idx / ((int64_t)PW * PH)) % C);
        int n = (int)(idx / ((int64_t)PW * PH * C));
        const float* r = rois + (size_t)n * 5;
        int b = (int)r[0];
        if (b < 0 || b >= batch) continue;
        CAlignGeom g = c_align_geom(r, scale, PH, PW, sampling_ratio);
        float* plane = grad_in + ((size_t)b * C + c) * H * W;
        float go = grad_out[idx];
        float count = (float)(g.gh * g.gw);
        for (int iy = 0; iy < g.gh; ++iy) {
            float y = g.sh + (float)ph * g.bh + ((float)iy + .5f) * g.bh / (float)g.gh;
            for (int ix = 0; ix < g.gw; ++ix) {
                float x = g.sw + (float)pw * g.bw + ((float)ix + .5f) * g.bw / (float)g.gw;
                Bilin q = bilin_setup(H, W, y, x);
                if (!q.ok) continue;
                atomicAdd(plane + q.y0 * W + q.x0, go * q.w1 / count);
                atomicAdd(plane + q.y0 * W + q.x1, go * q.w2 / count);
                atomicAdd(plane + q.y1 * W + q.x0, go * q.w3 / count);
                atomicAdd(plane + q.y1 * W + q.x1, go * q.w4 / count);
            }
        }
    }
}

int roi_pool_rows_plane(const float* features, const float* rois, void* out, int batch, int channels, int height,
                        int width, int num_rois, int pooled_h, int pooled_w, float spatial_scale, long long ldo,
                        int out_dtype, cudaStream_t stream);  // roi_pool_plane.cu

// roi_pool_argmax.cu: plane-resident forward with arg-max, owner-warp backward (no atomics)
int roi_pool_argmax_forward_plane(const float* features, const float* rois, float* out, int* argmax, int batch, int channels,
                                  int height, int width, int num_rois, int pooled_h, int pooled_w, float spatial_scale,
                                  int argmax_mode, cudaStream_t stream);
int roi_pool_backward_owner(const float* grad_out, const float* rois, const int* argmax, float* grad_in, int batch,
                            int channels, int height, int width, int num_rois, int pooled_h, int pooled_w,
                            float spatial_scale, int argmax_mode, cudaStream_t stream);

static int pool_args_ok(const char* who, const void* a, const void* b, const void* c, int batch, int channels,
                        int height, int width, int num_rois, int ph, int pw) {
    I2V_REQUIRE(batch >= 0 && channels >= 0 && num_rois >= 0 && height >= 1 && width >= 1 && ph >= 1 && pw >= 1,
                "%s: bad size", who);
    if (num_rois > 0 && channels > 0) I2V_REQUIRE(a && b && c, "%s: null pointer", who);
    return I2V_OK;
}

}  // namespace i2v

using namespace i2v;

extern "C" int i2v_roi_pool_forward(const float* features, const float* rois, float* out, int* argmax, int batch,
                                    int channels, int height, int width, int num_rois, int pooled_h, int pooled_w,
                                    float spatial_scale, int argmax_mode, cudaStream_t stream) {
    I2V_TRY(pool_args_ok("roi_pool_forward", features, rois, out, batch, channels, height, width, num_rois, pooled_h, pooled_w));
    I2V_REQUIRE(argmax_mode == I2V_ARGMAX_FLAT || argmax_mode == I2V_ARGMAX_PLANE, "roi_pool_forward: bad argmax_mode");
    I2V_REQUIRE(argmax_mode != I2V_ARGMAX_FLAT || (int64_t)batch * channels * height * width <= INT32_MAX,
                "roi_pool_forward: flat arg-max does not fit int32 for this feature tensor");
    int64_t total = (int64_t)num_rois * channels * pooled_h * pooled_w;
    if (total == 0) return I2V_OK;
    // The plane-resident forward (roi_pool_argmax.cu) measured 2.58 ms against 1.98 ms for the per-element kernel at 8 x 300
    // RoIs x 1024 channels (profiles/README.md): it stays selectable (I2V_POOL_PLANE=1, read per call) and parity-tested,
    // the per-element kernel stays the default.
    if (getenv("I2V_POOL_PLANE") != nullptr) {
        int rc = roi_pool_argmax_forward_plane(features, rois, out, argmax, batch, channels, height, width, num_rois, pooled_h,
                                               pooled_w, spatial_scale, argmax_mode, stream);
        if (rc != I2V_ERR_UNSUPPORTED) return rc;
    }
    int grid = grid_for(total, 256);
    if (argmax_mode == I2V_ARGMAX_FLAT)
        roi_pool_fwd_kernel<I2V_ARGMAX_FLAT><<<grid, 256, 0, stream>>>(features, rois, out, argmax, total, batch, channels, height, width, pooled_h, pooled_w, spatial_scale);
    else
        roi_pool_fwd_kernel<I2V_ARGMAX_PLANE><<<grid, 256, 0, stream>>>(features, rois, out, argmax, total, batch, channels, height, width, pooled_h, pooled_w, spatial_scale);
    return check_launch("roi_pool_fwd_kernel");
}

extern "C" int i2v_roi_pool_rows(const float* features, const float* rois, void* out, int batch, int channels,
                                 int height, int width, int num_rois, int pooled_h, int pooled_w, float spatial_scale,
                                 long long ldo, int out_dtype, cudaStream_t stream) {
    I2V_TRY(pool_args_ok("roi_pool_rows", features, rois, out, batch, channels, height, width, num_rois, pooled_h, pooled_w));
    I2V_REQUIRE(out_dtype == I2V_DT_F32 || out_dtype == I2V_DT_BF16, "roi_pool_rows: out_dtype %d", out_dtype);
    I2V_REQUIRE(ldo >= (long long)channels * pooled_h * pooled_w, "roi_pool_rows: row pitch smaller than a row");
    int64_t total = (int64_t)num_rois * channels * pooled_h * pooled_w;
    if (total == 0) return I2V_OK;
    int rc = roi_pool_rows_plane(features, rois, out, batch, channels, height, width, num_rois, pooled_h, pooled_w,
                                 spatial_scale, ldo, out_dtype, stream);
    if (rc != I2V_ERR_UNSUPPORTED) return rc;   // otherwise: shapes the plane kernel does not take
    int grid = grid_for(total, 256);
    if (out_dtype == I2V_DT_BF16)
        roi_pool_rows_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(features, rois, static_cast<__nv_bfloat16*>(out), total, batch, channels, height, width, pooled_h, pooled_w, spatial_scale, ldo);
    else
        roi_pool_rows_kernel<float><<<grid, 256, 0, stream>>>(features, rois, static_cast<float*>(out), total, batch, channels, height, width, pooled_h, pooled_w, spatial_scale, ldo);
    return check_launch("roi_pool_rows_kernel");
}

extern "C" int i2v_roi_pool_backward(const float* grad_out, const float* rois, const int* argmax, float* grad_in,
                                     int batch, int channels, int height, int width, int num_rois, int pooled_h,
                                     int pooled_w, float spatial_scale, int argmax_mode, cudaStream_t stream) {
    I2V_TRY(pool_args_ok("roi_pool_backward", grad_out, rois, argmax, batch, channels, height, width, num_rois, pooled_h, pooled_w));
    I2V_REQUIRE(argmax_mode == I2V_ARGMAX_FLAT || argmax_mode == I2V_ARGMAX_PLANE, "roi_pool_backward: bad argmax_mode");
    size_t in_elems = (size_t)batch * channels * height * width;
    if (in_elems == 0) return I2V_OK;
    I2V_REQUIRE(grad_in, "roi_pool_backward: null grad_in");
    if (getenv("I2V_POOL_PLANE") != nullptr && num_rois > 0) {
        // owner warps: every gradient plane is accumulated in shared memory by one warp and written once -- no atomics,
        // bit-reproducible, but 5.7 ms against 0.66 ms for the atomic scatter at 8 x 300 RoIs x 1024 channels (each warp
        // walks its frame's RoIs one global-load latency at a time): opt-in (I2V_POOL_PLANE=1), not the default
        int rc = roi_pool_backward_owner(grad_out, rois, argmax, grad_in, batch, channels, height, width, num_rois, pooled_h,
                                         pooled_w, spatial_scale, argmax_mode, stream);
        if (rc != I2V_ERR_UNSUPPORTED) return rc;
    }
    I2V_CUDA_TRY(cudaMemsetAsync(grad_in, 0, in_elems * sizeof(float), stream));
    int64_t total = (int64_t)num_rois * channels * pooled_h * pooled_w;
    if (total == 0) return I2V_OK;
    int grid = grid_for(total, 256);
    if (argmax_mode == I2V_ARGMAX_FLAT)
        roi_pool_bwd_kernel<I2V_ARGMAX_FLAT><<<grid, 256, 0, stream>>>(grad_out, rois, argmax, grad_in, total, batch, channels, height, width, pooled_h, pooled_w, spatial_scale);
    else
        roi_pool_bwd_kernel<I2V_ARGMAX_PLANE><<<grid, 256, 0, stream>>>(grad_out, rois, argmax, grad_in, total, batch, channels, height, width, pooled_h, pooled_w, spatial_scale);
    return check_launch("roi_pool_bwd_kernel");
}

extern "C" int i2v_c_roi_align_forward(const float* features, const float* rois, float* out, int batch, int channels,
                                       int height, int width, int num_rois, int pooled_h, int pooled_w,
                                       float spatial_scale, int sampling_ratio, cudaStream_t stream) {
    I2V_TRY(pool_args_ok("c_roi_align_forward", features, rois, out, batch, channels, height, width, num_rois, pooled_h, pooled_w));
    int64_t total = (int64_t)num_rois * channels * pooled_h * pooled_w;
    if (total == 0) return I2V_OK;
    c_roi_align_fwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>(features, rois, out, total, batch, channels, height,
                                                                     width, pooled_h, pooled_w, spatial_scale, sampling_ratio);
    return check_launch("c_roi_align_fwd_kernel");
}

extern "C" int i2v_c_roi_align_backward(const float* grad_out, const float* rois, float* grad_in, int batch,
                                        int channels, int height, int width, int num_rois, int pooled_h, int pooled_w,
                                        float spatial_scale, int sampling_ratio, cudaStream_t stream) {
    I2V_TRY(pool_args_ok("c_roi_align_backward", grad_out, rois, grad_in, batch, channels, height, width, num_rois, pooled_h, pooled_w));
    size_t in_elems = (size_t)batch * channels * height * width;
    if (in_elems == 0) return I2V_OK;
    I2V_REQUIRE(grad_in, "c_roi_align_backward: null grad_in");
    I2V_CUDA_TRY(cudaMemsetAsync(grad_in, 0, in_elems * sizeof(float), stream));
    int64_t total = (int64_t)num_rois * channels * pooled_h * pooled_w;
    if (total == 0) return I2V_OK;
    c_roi_align_bwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>(grad_out, rois, grad_in, total, batch, channels,
                                                                     height, width, pooled_h, pooled_w, spatial_scale, sampling_ratio);
    return check_launch("c_roi_align_bwd_kernel");
}
