// RoIPool with arg-max for the module API (`_RoIPooling`, `ROIPool((7,7), 1/16)`), plane-resident and atomic-free.
//
// Reference semantics: lib/model/roi_pooling/src/roi_pooling_kernel.cu:24-93 (forward; I2V_ARGMAX_FLAT: the arg-max is an
// index into the whole [B,C,H,W] tensor) and :128-203 (backward), and the op behind model._C
// (lib/model/roi_layers/roi_pool.py:17-19,30-42; I2V_ARGMAX_PLANE: h*W+w inside the (b,c) plane).
//
// Forward: one CTA per (frame, 16 channels, RoI slice) keeps the 16 feature planes in shared memory as
// [row][pitch][16 channels] (pitch odd, so the two half-warps -- even / odd rows of a bin at the same column -- hit
// opposite bank halves: every load conflict free, as in roi_pool_plane.cu).  A warp takes one RoI at a time; each
// half-warp keeps (maximum, index of its first occurrence in row-major order) over its rows of the bin, the halves are
// merged with "larger value, then smaller index", which is exactly the reference's row-major scan with a strict '>'.
// Values and indices of a (RoI, 16 channels) tile are contiguous in [N,C,7,7]: they are staged and leave as two TMA
// bulk stores.
//
// Backward: a warp OWNS two (frame, channel) gradient planes in shared memory and walks the frame's RoIs in list order,
// so no two threads of different warps ever add to the same address: no atomics, deterministic.  Lanes are the bins of
// the two channels, taken in four rounds by the parity of (ph, pw): bins two apart cannot share a cell as long as a bin
// is at least one cell in both directions, so the lanes of a round add concurrently; RoIs with smaller bins go bin by
// bin.  The cffi flavour applies the feasibility tests of roi_pooling_kernel.cu:160-183.  The planes are written to HBM
// once (no memset pass, no read-modify-write in global memory).
#include <float.h>

#include "common.cuh"

namespace i2v {
namespace {

constexpr int kK = 16;
__host__ __device__ constexpr int pool_pitch_for(int W) { return W <= 40 ? 41 : 65; }
constexpr int kFwdWarps = 10;
constexpr int kFwdThreads = kFwdWarps * 32;
constexpr int kBins = 49;

__device__ __forceinline__ void cp_async4(void* sdst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void bulk_store_commit(void* gdst, const void* ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template <int MODE, int kPitch>
__global__ void __launch_bounds__(kFwdThreads, 1)
    roi_pool_argmax_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ rois, float* __restrict__ out,
                               int* __restrict__ argmax, int batch, int C, int H, int W, int num_rois, float scale,
                               int split) {
    extern __shared__ __align__(128) float smem[];
    float* planes = smem;                                                   // [H][pitch][16]
    float* stage_v = smem + (size_t)H * kPitch * kK;                        // [warps][16][49]
    int* stage_i = reinterpret_cast<int*>(stage_v + kFwdWarps * kK * kBins);   // [warps][16][49]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ctiles = C / kK;
    const int s = blockIdx.x % split;
    const int ct = (blockIdx.x / split) % ctiles;
    const int b = blockIdx.x / (split * ctiles);
    {
        const int HW = H * W;
        const int tc = lane & 15, tdx = lane >> 4;
        const float* src = feat + ((size_t)b * C + (size_t)ct * kK + tc) * HW + tdx;
        float* dst = planes + tdx * kK + tc;
        for (int row = warp; row < H; row += kFwdWarps) {
            const float* g = src + row * W;
            float* d = dst + (size_t)row * kPitch * kK;
            for (int j = 0; 2 * j < W; ++j)
                if (2 * j + tdx < W) cp_async4(d + j * 2 * kK, g + 2 * j);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    }
    const int half = lane >> 4, c = lane & 15;
    float* my_v = stage_v + (size_t)warp * kK * kBins;
    int* my_i = stage_i + (size_t)warp * kK * kBins;
    const float* lane_base = planes + c;
    const int plane_off = (int)(((int64_t)b * C + ct * kK + c) * H * W);   // fits int32 in the flat flavour (host check)
    for (int n = s * kFwdWarps + warp; n < num_rois; n += split * kFwdWarps) {
        const float* r = rois + (size_t)n * 5;
        const int rb = (int)__ldg(r);
        const bool mine = (rb == b);
        const bool stray = (b == 0) && (rb < 0 || rb >= batch);     // out-of-range frame index: zeros and -1
        if (!mine && !stray) continue;
        int lo = 0, hi = 0;       // lanes 0-6: (hstart, hend) of bin row `lane`; lanes 7-13: (wstart, wend) of bin column
        {
            const int axis = lane >= 7;
            const int p = axis ? lane - 7 : lane;
            const float a0 = __ldg(r + (axis ? 1 : 2)), a1 = __ldg(r + (axis ? 3 : 4));
            const int rs = (int)roundf(__fmul_rn(a0, scale)), re = (int)roundf(__fmul_rn(a1, scale));
            const int extent = max(re - rs + 1, 1);
            const float bin = __fdiv_rn((float)extent, 7.f);
            const int lim = axis ? W : H;
            lo = min(max((int)floorf(__fmul_rn((float)p, bin)) + rs, 0), lim);
            hi = min(max((int)ceilf(__fmul_rn((float)(p + 1), bin)) + rs, 0), lim);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        if (mine) {
#pragma unroll 1
            for (int ph = 0; ph < 7; ++ph) {
                const int hs = __shfl_sync(0xffffffffu, lo, ph), he = __shfl_sync(0xffffffffu, hi, ph);
#pragma unroll 1
                for (int pw = 0; pw < 7; ++pw) {
                    const int ws = __shfl_sync(0xffffffffu, lo, 7 + pw), we = __shfl_sync(0xffffffffu, hi, 7 + pw);
                    float best = -FLT_MAX;
                    int bi = 0x7fffffff;
                    // half 0: rows hs, hs+2, ...; half 1: rows hs+1, hs+3, ... (same column, opposite bank halves)
                    for (int h = hs + half; h < he; h += 2) {
                        const float* p = lane_base + ((size_t)h * kPitch + ws) * kK;
                        for (int w = ws; w < we; ++w, p += kK) {
                            const float v = *p;
                            if (v > best) {             // strict: the first maximum of the scan stays (roi_pooling_kernel.cu:76)
                                best = v;
                                bi = h * W + w;
                            }
                        }
                    }
                    const float ov = __shfl_xor_sync(0xffffffffu, best, 16);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, 16);
                    if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > best || (ov == best && oi < bi))) {
                        best = ov;
                        bi = oi;
                    }
                    if (half == 0) {
                        const bool none = bi == 0x7fffffff;         // empty bin, or nothing above -FLT_MAX (all NaN)
                        const bool empty = he <= hs || we <= ws;
                        my_v[c * kBins + ph * 7 + pw] = empty ? 0.f : (none ? -FLT_MAX : best);
                        my_i[c * kBins + ph * 7 + pw] = none ? -1 : (MODE == I2V_ARGMAX_FLAT ? plane_off + bi : bi);
                    }
                }
            }
        } else {
            for (int i = lane; i < kK * kBins; i += 32) {
                my_v[i] = 0.f;
                my_i[i] = -1;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            bulk_store_commit(out + ((size_t)n * C + (size_t)ct * kK) * kBins, my_v, kK * kBins * 4u);
            if (argmax) bulk_store_commit(argmax + ((size_t)n * C + (size_t)ct * kK) * kBins, my_i, kK * kBins * 4u);
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

size_t fwd_smem_bytes(int H, int W) {
    return (size_t)H * pool_pitch_for(W) * kK * sizeof(float) + (size_t)kFwdWarps * kK * kBins * 8;
}

// ------------------------------------------------------------------------------------------ backward
constexpr int kBwdWarps = 8;            // each owns two channel planes
constexpr int kBwdThreads = kBwdWarps * 32;
constexpr int kChunk = 1024;            // RoIs scanned per round of the CTA-wide list build

struct BwdGeom {       // roi_pooling_kernel.cu:44-53
    int rs_w, rs_h, re_w, re_h;
    float bin_h, bin_w;
};
__device__ __forceinline__ BwdGeom bwd_geom(const float* __restrict__ r, float scale) {
    BwdGeom g;
    g.rs_w = (int)roundf(__fmul_rn(__ldg(r + 1), scale));
    g.rs_h = (int)roundf(__fmul_rn(__ldg(r + 2), scale));
    g.re_w = (int)roundf(__fmul_rn(__ldg(r + 3), scale));
    g.re_h = (int)roundf(__fmul_rn(__ldg(r + 4), scale));
    const int rw = max(g.re_w - g.rs_w + 1, 1), rh = max(g.re_h - g.rs_h + 1, 1);
    g.bin_h = __fdiv_rn((float)rh, 7.f);
    g.bin_w = __fdiv_rn((float)rw, 7.f);
    return g;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

template <int MODE>
__global__ void __launch_bounds__(kBwdThreads, 1)
    roi_pool_owner_bwd_kernel(const float* __restrict__ grad_out, const float* __restrict__ rois, const int* __restrict__ argmax,
                              float* __restrict__ grad_in, int batch, int C, int H, int W, int num_rois, float scale) {
    extern __shared__ __align__(128) float smem[];
    const int HW = H * W;
    float* planes = smem;                                               // [warps][2][HW]
    int* list = reinterpret_cast<int*>(smem + (size_t)kBwdWarps * 2 * HW);  // [kChunk]
    __shared__ int s_warp[kBwdWarps];
    __shared__ int s_count;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ctiles = C / kK;
    const int ct = blockIdx.x % ctiles;
    const int b = blockIdx.x / ctiles;
    const int hl = lane >> 4;                      // which of the warp's two channels
    const int ch = ct * kK + warp * 2 + hl;
    float* plane = planes + ((size_t)warp * 2 + hl) * HW;
    for (int i = tid; i < kBwdWarps * 2 * HW; i += kBwdThreads) planes[i] = 0.f;
    const int64_t plane_off = ((int64_t)b * C + ch) * HW;

    // lane -> bin of round q: the bins with (ph & 1, pw & 1) == (q >> 1, q & 1), 16 / 12 / 12 / 9 of them
    const int j = lane & 15;
    for (int base = 0; base < num_rois; base += kChunk) {
        // ---- the RoIs of this chunk that belong to frame b, in order (block-wide ordered compaction) ----
        __syncthreads();
        int mine[kChunk / kBwdThreads];
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < kChunk / kBwdThreads; ++k) {
            const int n = base + tid * (kChunk / kBwdThreads) + k;
            const bool m = n < num_rois && (int)__ldg(rois + (size_t)n * 5) == b;
            mine[k] = m ? n : -1;
            cnt += m;
        }
        int incl = cnt;
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < kBwdWarps; ++w) {
            if (w < warp) before += s_warp[w];
            total += s_warp[w];
        }
        int pos = before + incl - cnt;
#pragma unroll
        for (int k = 0; k < kChunk / kBwdThreads; ++k)
            if (mine[k] >= 0) list[pos++] = mine[k];
        if (tid == 0) s_count = total;
        __syncthreads();
        const int count = s_count;

        // ---- every warp walks the list for its two channels ----
        for (int i = 0; i < count; ++i) {
            const int n = list[i];
            const float* r = rois + (size_t)n * 5;
            const BwdGeom g = bwd_geom(r, scale);
            const size_t row = ((size_t)n * C + ch) * kBins;
            const bool wide = g.bin_h >= 1.f && g.bin_w >= 1.f;     // uniform per RoI
            auto add_bin = [&](int ph, int pw) {
                const int am = __ldg(argmax + row + ph * 7 + pw);
                if (am < 0) return;
                int cell;
                if (MODE == I2V_ARGMAX_PLANE) {
                    cell = am;
                    if (cell >= HW) return;
                } else {
                    // roi_pooling_kernel.cu:143-183: the input element only collects from RoIs that contain it and from
                    // the pooled cells in its feasible window
                    const int64_t local = (int64_t)am - plane_off;
                    if (local < 0 || local >= HW) return;
                    cell = (int)local;
                    const int h = cell / W, w = cell - h * W;
                    if (!(w >= g.rs_w && w <= g.re_w && h >= g.rs_h && h <= g.re_h)) return;
                    const int p0 = clampi((int)floorf(__fdiv_rn((float)(h - g.rs_h), g.bin_h)), 0, 7);
                    const int p1 = clampi((int)ceilf(__fdiv_rn((float)(h - g.rs_h + 1), g.bin_h)), 0, 7);
                    const int q0 = clampi((int)floorf(__fdiv_rn((float)(w - g.rs_w), g.bin_w)), 0, 7);
                    const int q1 = clampi((int)ceilf(__fdiv_rn((float)(w - g.rs_w + 1), g.bin_w)), 0, 7);
                    if (ph < p0 || ph >= p1 || pw < q0 || pw >= q1) return;
                }
                plane[cell] += __ldg(grad_out + row + ph * 7 + pw);
            };
            if (wide) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int nph = (q >> 1) ? 3 : 4, npw = (q & 1) ? 3 : 4;       // odd: 1,3,5; even: 0,2,4,6
                    if (j < nph * npw) add_bin(2 * (j / npw) + (q >> 1), 2 * (j % npw) + (q & 1));
                    __syncwarp();
                }
            } else {
                for (int bin = 0; bin < kBins; ++bin) {
                    if (j == 0) add_bin(bin / 7, bin % 7);
                    __syncwarp();
                }
            }
        }
    }
    // ---- write-out: the warp's two planes, coalesced ----
    __syncthreads();
    for (int k = 0; k < 2; ++k) {
        const float* src = planes + ((size_t)warp * 2 + k) * HW;
        float* dst = grad_in + ((int64_t)b * C + ct * kK + warp * 2 + k) * HW;
        for (int i = lane; i < HW; i += 32) dst[i] = src[i];
    }
}

size_t bwd_smem_bytes(int H, int W) { return (size_t)kBwdWarps * 2 * H * W * sizeof(float) + kChunk * sizeof(int); }

}  // namespace

// Returns I2V_ERR_UNSUPPORTED for shapes these kernels do not take (the callers in roi_pool.cu then use the per-element kernels).
int roi_pool_argmax_forward_plane(const float* features, const float* rois, float* out, int* argmax, int batch, int channels,
                                  int height, int width, int num_rois, int pooled_h, int pooled_w, float spatial_scale,
                                  int argmax_mode, cudaStream_t stream) {
    const bool ok = pooled_h == 7 && pooled_w == 7 && channels % kK == 0 && width <= 64 && batch >= 1 &&
                    fwd_smem_bytes(height, width) <= (size_t)kMaxSmemPerCta && ((uintptr_t)out & 15) == 0 &&
                    ((uintptr_t)argmax & 15) == 0;
    if (!ok) return I2V_ERR_UNSUPPORTED;
    const int ctiles = channels / kK;
    int split = 1;
    while (batch * ctiles * split < 2 * kNumSMs && split * kFwdWarps < num_rois && split < 16) split *= 2;
    const size_t smem = fwd_smem_bytes(height, width);
    const dim3 grid((unsigned)(batch * ctiles * split));
#define I2V_LAUNCH(MODE, PITCH)                                                                                             \
    do {                                                                                                                    \
        auto kern = roi_pool_argmax_fwd_kernel<MODE, PITCH>;                                                                \
        I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                   \
        kern<<<grid, kFwdThreads, smem, stream>>>(features, rois, out, argmax, batch, channels, height, width, num_rois,    \
                                                  spatial_scale, split);                                                    \
    } while (0)
    const bool narrow = pool_pitch_for(width) == 41;
    if (argmax_mode == I2V_ARGMAX_FLAT) {
        if (narrow) I2V_LAUNCH(I2V_ARGMAX_FLAT, 41);
        else I2V_LAUNCH(I2V_ARGMAX_FLAT, 65);
    } else {
        if (narrow) I2V_LAUNCH(I2V_ARGMAX_PLANE, 41);
        else I2V_LAUNCH(I2V_ARGMAX_PLANE, 65);
    }
#undef I2V_LAUNCH
    return check_launch("roi_pool_argmax_fwd_kernel");
}

int roi_pool_backward_owner(const float* grad_out, const float* rois, const int* argmax, float* grad_in, int batch,
                            int channels, int height, int width, int num_rois, int pooled_h, int pooled_w,
                            float spatial_scale, int argmax_mode, cudaStream_t stream) {
    const bool ok = pooled_h == 7 && pooled_w == 7 && channels % kK == 0 && batch >= 1 &&
                    bwd_smem_bytes(height, width) <= (size_t)kMaxSmemPerCta;
    if (!ok) return I2V_ERR_UNSUPPORTED;
    const size_t smem = bwd_smem_bytes(height, width);
    const dim3 grid((unsigned)(batch * (channels / kK)));
    if (argmax_mode == I2V_ARGMAX_FLAT) {
        auto kern = roi_pool_owner_bwd_kernel<I2V_ARGMAX_FLAT>;
        I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kBwdThreads, smem, stream>>>(grad_out, rois, argmax, grad_in, batch, channels, height, width, num_rois,
                                                  spatial_scale);
    } else {
        auto kern = roi_pool_owner_bwd_kernel<I2V_ARGMAX_PLANE>;
        I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kBwdThreads, smem, stream>>>(grad_out, rois, argmax, grad_in, batch, channels, height, width, num_rois,
                                                  spatial_scale);
    }
    return check_launch("roi_pool_owner_bwd_kernel");
}

}  // namespace i2v
