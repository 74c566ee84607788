// Plane-resident RoIPool rows for the SGG projection (resnet_SGG_emb.py:144-146,158-160): 4032 union boxes pool the
// SAME frame, so the 16 feature planes of a CTA are copied from HBM/L2 into shared memory once and every bin maximum is
// taken from there.
//
//   * planes as [row][65 columns][16 channels] fp32: lanes are channels, a half-warp reads one 64-byte cell; with the odd
//     row pitch the bank half of a cell is (row + column) mod 2, so the two half-warps -- which scan the even and the odd
//     rows of the same bin, same column -- never collide: every shared load is conflict free by construction;
//   * one warp per RoI at a time: bin bounds are computed once per RoI by 14 lanes (roi_pooling_kernel.cu:44-66
//     arithmetic, the model._C flavour has the same rounding) and broadcast with shuffles;
//   * the 16 x 49 results of a (RoI, channel tile) are contiguous in the [N, C*49] output row: they are staged in shared
//     memory and leave as one TMA bulk store (1568 bytes in bf16, 3136 in fp32).
#include <cuda_bf16.h>
#include <float.h>
#include <stdlib.h>

#include "common.cuh"

namespace i2v {
namespace {

constexpr int kK = 16;            // channels per CTA
// cells per shared-memory row, odd (see above): 65 for landscape maps (W <= 64), 41 for portrait ones (W <= 40, H up to 77)
__host__ __device__ constexpr int pool_pitch_for(int W) { return W <= 40 ? 41 : 65; }
constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;
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

__device__ __forceinline__ void put(float* p, float v) { *p = v; }
__device__ __forceinline__ void put(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// One sweep over the columns of a bin row that is NR rows tall (NR <= 12).  Half-warp 0 owns rows 0, 2, ..., half-warp 1
// rows 1, 3, ... of the bin row: NR/2 loads at compile-time offsets for both halves, plus -- for odd NR -- one more at
// `tail` (the extra row of half 0; half 1 re-reads its last row, which cannot change a maximum).  No lane-dependent
// control flow.  Bins tile the columns with at most one shared column, whose value is carried into the next bin: `we[pw]`
// is the end column of bin pw and bit pw of `shared` says that bin pw + 1 starts on bin pw's last column; both are the
// same for the seven bin rows of a RoI and live in registers (the bin loop is unrolled), and the cross-half maxima and
// the stores of the row's seven bins are issued together at the end.
template <int NR, int kPitch, typename OutT>
__device__ __forceinline__ void sweep_bin_row(const float* p, int tail, int x, const int (&we)[7], unsigned shared, int half,
                                              OutT* dst) {
    constexpr int kRow2 = 2 * kPitch * kK;
    float v = -FLT_MAX, carry = -FLT_MAX;
    float best[7];
#pragma unroll
    for (int pw = 0; pw < 7; ++pw) {
        float b = carry;
#pragma unroll 2
        for (; x < we[pw]; ++x, p += kK) {      // the tight part: loads and maxima only
            v = (NR & 1) ? p[tail] : -FLT_MAX;
#pragma unroll
            for (int k = 0; k < NR / 2; ++k) v = fmaxf(v, p[k * kRow2]);
            b = fmaxf(b, v);
        }
        best[pw] = b;
        carry = ((shared >> pw) & 1u) ? v : -FLT_MAX;   // `v` is the bin's last column
    }
#pragma unroll
    for (int pw = 0; pw < 7; ++pw) best[pw] = fmaxf(best[pw], __shfl_xor_sync(0xffffffffu, best[pw], 16));
    if (half == 0) {
#pragma unroll
        for (int pw = 0; pw < 7; ++pw) put(dst + pw, best[pw]);
    }
}

template <typename OutT, int kPitch>
__global__ void __launch_bounds__(kThreads, 1)
    roi_pool_plane_kernel(const float* __restrict__ feat, const float* __restrict__ rois, OutT* __restrict__ out,
                          int batch, int C, int H, int W, int num_rois, float scale, int64_t ldo, int split) {
    extern __shared__ __align__(128) float smem[];
    float* planes = smem;                                             // [H][65][16]
    OutT* stage = reinterpret_cast<OutT*>(smem + (size_t)H * kPitch * kK);   // [warps][16][49]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ctiles = C / kK;
    const int s = blockIdx.x % split;
    const int ct = (blockIdx.x / split) % ctiles;
    const int b = blockIdx.x / (split * ctiles);

    // ---- fill: global [c][row][col] -> shared [row][col][16]; a warp-wide 4-byte cp.async writes two adjacent cells x
    // 16 channels = 128 contiguous shared bytes; everything is in flight before the single wait ----
    {
        const int HW = H * W;
        const int tc = lane & 15, tdx = lane >> 4;
        const float* src = feat + ((size_t)b * C + (size_t)ct * kK + tc) * HW + tdx;
        float* dst = planes + tdx * kK + tc;
        for (int row = warp; row < H; row += kWarps) {
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
    OutT* my_stage = stage + (size_t)warp * kK * kBins;
    const float* lane_base = planes + c;
    for (int n = s * kWarps + warp; n < num_rois; n += split * kWarps) {
        const float* r = rois + (size_t)n * 5;
        const int rb = (int)__ldg(r);
        const bool mine = (rb == b);
        const bool stray = (b == 0) && (rb < 0 || rb >= batch);     // out-of-range frame index: a zero row
        if (!mine && !stray) continue;                               // uniform per warp
        // ---- bin bounds: lanes 0-6 hold (hstart, hend) of bin row `lane`, lanes 7-13 (wstart, wend) of bin column ----
        int lo = 0, hi = 0;
        {
            const int axis = lane >= 7;                   // 0: rows (y), 1: columns (x)
            const int p = axis ? lane - 7 : lane;
            const float a0 = __ldg(r + (axis ? 1 : 2)), a1 = __ldg(r + (axis ? 3 : 4));
            const int rs = (int)roundf(__fmul_rn(a0, scale)), re = (int)roundf(__fmul_rn(a1, scale));
            const int extent = max(re - rs + 1, 1);
            const float bin = __fdiv_rn((float)extent, 7.f);
            const int lim = axis ? W : H;
            lo = min(max((int)floorf(__fmul_rn((float)p, bin)) + rs, 0), lim);
            hi = min(max((int)ceilf(__fmul_rn((float)(p + 1), bin)) + rs, 0), lim);
        }
        // the previous tile of this warp must have left the staging buffer
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        if (mine) {
            // regular RoI: every bin column is non-empty.  Consecutive bins then tile [ws[0], we[6]) with at most one
            // shared column (floor / ceil of the same product), which is what the sweep below relies on.
            const int next_lo = __shfl_down_sync(0xffffffffu, lo, 1);
            const unsigned nonempty = __ballot_sync(0xffffffffu, hi > lo);
            const unsigned tiles = __ballot_sync(0xffffffffu, next_lo >= hi - 1 && next_lo <= hi);
            const bool regular = ((nonempty >> 7) & 0x7fu) == 0x7fu && ((tiles >> 7) & 0x3fu) == 0x3fu;
            int we[7];                                // the bin columns' end, and which bins share a column with the next
#pragma unroll
            for (int pw = 0; pw < 7; ++pw) we[pw] = __shfl_sync(0xffffffffu, hi, 7 + pw);
            const unsigned shared = (__ballot_sync(0xffffffffu, next_lo == hi - 1) >> 7) & 0x3fu;
#pragma unroll 1
            for (int ph = 0; ph < 7; ++ph) {
                const int hs = __shfl_sync(0xffffffffu, lo, ph), he = __shfl_sync(0xffffffffu, hi, ph);
                OutT* dst = my_stage + c * kBins + ph * 7;
                if (he <= hs) {                       // empty bin row: zeros (roi_pooling_kernel.cu:68-70)
                    if (half == 0)
                        for (int pw = 0; pw < 7; ++pw) put(dst + pw, 0.f);
                    continue;
                }
                // half 0 takes rows hs, hs+2, ...; half 1 rows hs+1, hs+3, ...: same column, opposite bank half.
                // The sweep is specialised on the bin height up to 12 rows; taller bins (a RoI reaching far past the map:
                // rs and re are not clipped, roi_pooling_kernel.cu:60-66) take the generic loop below.
                const int nr = he - hs;
                const int myrows = (nr - half + 1) >> 1;
                constexpr int kRow2 = 2 * kPitch * kK;          // two rows further down, in floats
                // a one-row bin leaves half 1 without a row of its own: it re-reads half 0's
                const float* rowp = lane_base + (size_t)(hs + (nr > 1 ? half : 0)) * kPitch * kK;
                if (regular && nr <= 12) {
                    const int x0 = __shfl_sync(0xffffffffu, lo, 7);
                    const float* p = rowp + (size_t)x0 * kK;
                    const int tail = (half == 0 ? nr / 2 : max(nr / 2 - 1, 0)) * kRow2;
                    switch (nr) {   // uniform
                        case 1: sweep_bin_row<1, kPitch>(p, tail, x0, we, shared, half, dst); break;
                        case 2: sweep_bin_row<2, kPitch>(p, tail, x0, we, shared, half, dst); break;
                        case 3: sweep_bin_row<3, kPitch>(p, tail, x0, we, shared, half, dst); break;
                        case 4: sweep_bin_row<4, kPitch>(p, tail, x0, we, shared, half, dst); break;
                        case 5: sweep_bin_row<5, kPitch>(p, tail, x0, we, shared, half, dst); break;
                        case 6: sweep_bin_row<6, kPitch>(p, tail, x0, we, shared, half, dst); break;
                        case 7: sweep_bin_row<7, kPitch>(p, tail, x0, we, shared, half, dst); break;
                        case 8: sweep_bin_row<8, kPitch>(p, tail, x0, we, shared, half, dst); break;
                        case 9: sweep_bin_row<9, kPitch>(p, tail, x0, we, shared, half, dst); break;
                        case 10: sweep_bin_row<10, kPitch>(p, tail, x0, we, shared, half, dst); break;
                        case 11: sweep_bin_row<11, kPitch>(p, tail, x0, we, shared, half, dst); break;
                        default: sweep_bin_row<12, kPitch>(p, tail, x0, we, shared, half, dst); break;
                    }
                } else {
#pragma unroll 1
                    for (int pw = 0; pw < 7; ++pw) {
                        const int ws = __shfl_sync(0xffffffffu, lo, 7 + pw), we = __shfl_sync(0xffffffffu, hi, 7 + pw);
                        float best = -FLT_MAX;
                        for (int h = 0; h < myrows; ++h)
                            for (int w = ws; w < we; ++w) best = fmaxf(best, rowp[(size_t)h * kRow2 + (size_t)w * kK]);
                        best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, 16));
                        if (half == 0) put(dst + pw, we <= ws ? 0.f : best);
                    }
                }
            }
        } else {
            for (int i = lane; i < kK * kBins; i += 32) put(my_stage + i, 0.f);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0)
            bulk_store_commit(out + (size_t)n * ldo + (size_t)ct * kK * kBins, my_stage, kK * kBins * (unsigned)sizeof(OutT));
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------ bf16 planes
// The same kernel for bf16 output with the planes themselves held in bf16: rounding to bf16 is monotone, so
// max(bf16(a), bf16(b)) == bf16(max(a, b)) and the pooled rows are bit-identical to the fp32 kernel's, while one 32-bit
// word of a cell now carries TWO channels and one `max.bf16x2` takes both maxima: a CTA covers 32 channels with the
// instruction count (and the shared-memory traffic) the fp32 kernel spends on 16.  The fill converts on the way in
// (two coalesced-by-L1 global loads, one packed store per word).
constexpr int kKB = 32;                   // channels per CTA: 16 lanes x 2
constexpr unsigned kNegInf2 = 0xFF80FF80u;
__device__ __forceinline__ unsigned hmax2(unsigned a, unsigned b) {
    unsigned r;
    asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ unsigned pack_bf16x2(float lo, float hi) {
    unsigned r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));     // first source -> upper half
    return r;
}

// `p` points at the lane's word (16 words per cell) in column `x`; dst_lo / dst_hi are the staging rows of its two channels.
// `we[pw]` is the end column of bin pw and bit pw of `shared` says that bin pw + 1 starts on bin pw's last column; both are
// the same for the seven bin rows of a RoI and live in registers (the bin loop is unrolled), so a bin costs its loads and
// maxima plus one shuffle and two stores, which are issued together after the row's seven bins.
template <int NR, int kPitch>
__device__ __forceinline__ void sweep_bin_row_bf16(const unsigned* p, int tail, int x, const int (&we)[7], unsigned shared,
                                                   int half, unsigned short* dst_lo, unsigned short* dst_hi) {
    constexpr int kRow2 = 2 * kPitch * 16;
    unsigned v = kNegInf2, carry = kNegInf2;
    unsigned best[7];
#pragma unroll
    for (int pw = 0; pw < 7; ++pw) {
        unsigned b = carry;
#pragma unroll 2
        for (; x < we[pw]; ++x, p += 16) {      // the tight part: loads and maxima only
            v = (NR & 1) ? p[tail] : kNegInf2;
#pragma unroll
            for (int k = 0; k < NR / 2; ++k) v = hmax2(v, p[k * kRow2]);
            b = hmax2(b, v);
        }
        best[pw] = b;
        carry = ((shared >> pw) & 1u) ? v : kNegInf2;   // `v` is the bin's last column
    }
#pragma unroll
    for (int pw = 0; pw < 7; ++pw) best[pw] = hmax2(best[pw], __shfl_xor_sync(0xffffffffu, best[pw], 16));
    if (half == 0) {
#pragma unroll
        for (int pw = 0; pw < 7; ++pw) {
            dst_lo[pw] = (unsigned short)(best[pw] & 0xffffu);
            dst_hi[pw] = (unsigned short)(best[pw] >> 16);
        }
    }
}

template <int kPitch>
__global__ void __launch_bounds__(kThreads, 1)
    roi_pool_plane_bf16_kernel(const float* __restrict__ feat, const float* __restrict__ rois,
                               __nv_bfloat16* __restrict__ out, int batch, int C, int H, int W, int num_rois, float scale,
                               int64_t ldo, int split) {
    extern __shared__ __align__(128) float smem[];
    unsigned* planes = reinterpret_cast<unsigned*>(smem);                             // [H][pitch][16 channel pairs]
    unsigned short* stage = reinterpret_cast<unsigned short*>(planes + (size_t)H * kPitch * 16);   // [warps][32][49]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ctiles = C / kKB;
    const int s = blockIdx.x % split;
    const int ct = (blockIdx.x / split) % ctiles;
    const int b = blockIdx.x / (split * ctiles);

    // ---- fill: global fp32 [c][row][col] -> shared bf16x2 [row][col][16]; lanes are (2 adjacent columns) x (16 channel
    // pairs), so a warp-wide store is 128 contiguous bytes; each global load touches 4 bytes of 32 sectors whose other
    // columns are served from L1 by the next iterations.  Eight iterations are in flight per lane ----
    {
        const int HW = H * W;
        const int tc = lane & 15, tdx = lane >> 4;
        const float* src = feat + ((size_t)b * C + (size_t)ct * kKB + 2 * tc) * HW + tdx;
        unsigned* dst = planes + tdx * 16 + tc;
        const int pairs = (W + 1) / 2;
        for (int row = warp; row < H; row += kWarps) {
            const float* g = src + row * W;
            unsigned* d = dst + (size_t)row * kPitch * 16;
            for (int j0 = 0; j0 < pairs; j0 += 8) {
                float a[8], bq[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j = j0 + u;
                    const bool ok = j < pairs && 2 * j + tdx < W;
                    a[u] = ok ? __ldg(g + 2 * j) : 0.f;
                    bq[u] = ok ? __ldg(g + 2 * j + HW) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j = j0 + u;
                    if (j < pairs && 2 * j + tdx < W) d[j * 32] = pack_bf16x2(a[u], bq[u]);
                }
            }
        }
        __syncthreads();
    }

    const int half = lane >> 4, c = lane & 15;
    unsigned short* my_stage = stage + (size_t)warp * kKB * kBins;
    const unsigned* lane_base = planes + c;
    for (int n = s * kWarps + warp; n < num_rois; n += split * kWarps) {
        const float* r = rois + (size_t)n * 5;
        const int rb = (int)__ldg(r);
        const bool mine = (rb == b);
        const bool stray = (b == 0) && (rb < 0 || rb >= batch);     // out-of-range frame index: a zero row
        if (!mine && !stray) continue;                               // uniform per warp
        // ---- bin bounds: lanes 0-6 hold (hstart, hend) of bin row `lane`, lanes 7-13 (wstart, wend) of bin column ----
        int lo = 0, hi = 0;
        {
            const int axis = lane >= 7;                   // 0: rows (y), 1: columns (x)
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
            const int next_lo = __shfl_down_sync(0xffffffffu, lo, 1);
            const unsigned nonempty = __ballot_sync(0xffffffffu, hi > lo);
            const unsigned tiles = __ballot_sync(0xffffffffu, next_lo >= hi - 1 && next_lo <= hi);
            const bool regular = ((nonempty >> 7) & 0x7fu) == 0x7fu && ((tiles >> 7) & 0x3fu) == 0x3fu;
            int we[7];                                // the bin columns' end, and which bins share a column with the next
#pragma unroll
            for (int pw = 0; pw < 7; ++pw) we[pw] = __shfl_sync(0xffffffffu, hi, 7 + pw);
            const unsigned shared = (__ballot_sync(0xffffffffu, next_lo == hi - 1) >> 7) & 0x3fu;
#pragma unroll 1
            for (int ph = 0; ph < 7; ++ph) {
                const int hs = __shfl_sync(0xffffffffu, lo, ph), he = __shfl_sync(0xffffffffu, hi, ph);
                unsigned short* dst_lo = my_stage + (2 * c) * kBins + ph * 7;
                unsigned short* dst_hi = dst_lo + kBins;
                if (he <= hs) {                       // empty bin row: zeros (roi_pooling_kernel.cu:68-70)
                    if (half == 0)
                        for (int pw = 0; pw < 7; ++pw) dst_lo[pw] = dst_hi[pw] = 0;
                    continue;
                }
                const int nr = he - hs;
                const int myrows = (nr - half + 1) >> 1;
                constexpr int kRow2 = 2 * kPitch * 16;
                const unsigned* rowp = lane_base + (size_t)(hs + (nr > 1 ? half : 0)) * kPitch * 16;
                if (regular && nr <= 12) {
                    const int x0 = __shfl_sync(0xffffffffu, lo, 7);
                    const unsigned* p = rowp + (size_t)x0 * 16;
                    const int tail = (half == 0 ? nr / 2 : max(nr / 2 - 1, 0)) * kRow2;
                    switch (nr) {   // uniform
                        case 1: sweep_bin_row_bf16<1, kPitch>(p, tail, x0, we, shared, half, dst_lo, dst_hi); break;
                        case 2: sweep_bin_row_bf16<2, kPitch>(p, tail, x0, we, shared, half, dst_lo, dst_hi); break;
                        case 3: sweep_bin_row_bf16<3, kPitch>(p, tail, x0, we, shared, half, dst_lo, dst_hi); break;
                        case 4: sweep_bin_row_bf16<4, kPitch>(p, tail, x0, we, shared, half, dst_lo, dst_hi); break;
                        case 5: sweep_bin_row_bf16<5, kPitch>(p, tail, x0, we, shared, half, dst_lo, dst_hi); break;
                        case 6: sweep_bin_row_bf16<6, kPitch>(p, tail, x0, we, shared, half, dst_lo, dst_hi); break;
                        case 7: sweep_bin_row_bf16<7, kPitch>(p, tail, x0, we, shared, half, dst_lo, dst_hi); break;
                        case 8: sweep_bin_row_bf16<8, kPitch>(p, tail, x0, we, shared, half, dst_lo, dst_hi); break;
                        case 9: sweep_bin_row_bf16<9, kPitch>(p, tail, x0, we, shared, half, dst_lo, dst_hi); break;
                        case 10: sweep_bin_row_bf16<10, kPitch>(p, tail, x0, we, shared, half, dst_lo, dst_hi); break;
                        case 11: sweep_bin_row_bf16<11, kPitch>(p, tail, x0, we, shared, half, dst_lo, dst_hi); break;
                        default: sweep_bin_row_bf16<12, kPitch>(p, tail, x0, we, shared, half, dst_lo, dst_hi); break;
                    }
                } else {
#pragma unroll 1
                    for (int pw = 0; pw < 7; ++pw) {
                        const int ws = __shfl_sync(0xffffffffu, lo, 7 + pw), we = __shfl_sync(0xffffffffu, hi, 7 + pw);
                        unsigned best = kNegInf2;
                        for (int h = 0; h < myrows; ++h)
                            for (int w = ws; w < we; ++w) best = hmax2(best, rowp[(size_t)h * kRow2 + (size_t)w * 16]);
                        best = hmax2(best, __shfl_xor_sync(0xffffffffu, best, 16));
                        if (half == 0) {
                            dst_lo[pw] = we <= ws ? (unsigned short)0 : (unsigned short)(best & 0xffffu);
                            dst_hi[pw] = we <= ws ? (unsigned short)0 : (unsigned short)(best >> 16);
                        }
                    }
                }
            }
        } else {
            for (int i = lane; i < kKB * kBins; i += 32) my_stage[i] = 0;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0)
            bulk_store_commit(out + (size_t)n * ldo + (size_t)ct * kKB * kBins, my_stage, kKB * kBins * 2u);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

size_t plane_bf16_smem_bytes(int H, int W) {
    return (size_t)H * pool_pitch_for(W) * 16 * sizeof(unsigned) + (size_t)kWarps * kKB * kBins * 2;
}

size_t plane_smem_bytes(int H, int W, size_t esz) {
    return (size_t)H * pool_pitch_for(W) * kK * sizeof(float) + (size_t)kWarps * kK * kBins * esz;
}

}  // namespace

// Used by i2v_roi_pool_rows (roi_pool.cu) when the shape allows it; returns I2V_ERR_UNSUPPORTED otherwise.
int roi_pool_rows_plane(const float* features, const float* rois, void* out, int batch, int channels, int height,
                        int width, int num_rois, int pooled_h, int pooled_w, float spatial_scale, long long ldo,
                        int out_dtype, cudaStream_t stream) {
    const size_t esz = out_dtype == I2V_DT_BF16 ? 2 : 4;
    const bool ok = pooled_h == 7 && pooled_w == 7 && channels % kK == 0 && width <= 64 && height <= 77 &&
                    batch >= 1 &&
                    plane_smem_bytes(height, width, esz) <= (size_t)kMaxSmemPerCta && ((uintptr_t)out & 15) == 0 &&
                    ((size_t)ldo * esz) % 16 == 0;
    if (!ok) return I2V_ERR_UNSUPPORTED;
    if (out_dtype == I2V_DT_BF16 && channels % kKB == 0 && !getenv("I2V_POOL_F32_PLANES") &&
        plane_bf16_smem_bytes(height, width) <= (size_t)kMaxSmemPerCta) {
        // bf16 planes, two channels per word: half the instructions; CTAs sized to fill whole waves
        // RoI slices per (frame, channel tile): the CTAs should fill whole waves of the 148 SMs (a CTA per SM: the planes
        // take 158 KB) while a slice keeps enough RoIs per warp to pay for its fill
        const int tiles = batch * (channels / kKB);
        int split = 1;
        {
            double best = 0.0;
            const int per_frame = max(1, num_rois / batch);
            for (int sp = 1; sp <= 16; ++sp) {
                if (sp > 1 && per_frame / (sp * kWarps) < 8) break;
                const double waves = (double)tiles * sp / kNumSMs;
                const double eff = waves / ceil(waves) - 0.004 * sp;   // a fill costs about 0.4 % of a two-slice CTA
                if (eff > best) {
                    best = eff;
                    split = sp;
                }
            }
            if (const char* e = getenv("I2V_POOL_SPLIT")) split = max(1, atoi(e));
        }
        const size_t smem = plane_bf16_smem_bytes(height, width);
        dim3 grid((unsigned)(tiles * split));
        if (pool_pitch_for(width) == 41) {
            auto kern = roi_pool_plane_bf16_kernel<41>;
            I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, kThreads, smem, stream>>>(features, rois, static_cast<__nv_bfloat16*>(out), batch, channels, height,
                                                   width, num_rois, spatial_scale, ldo, split);
        } else {
            auto kern = roi_pool_plane_bf16_kernel<65>;
            I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, kThreads, smem, stream>>>(features, rois, static_cast<__nv_bfloat16*>(out), batch, channels, height,
                                                   width, num_rois, spatial_scale, ldo, split);
        }
        return check_launch("roi_pool_plane_bf16_kernel");
    }
    const int ctiles = channels / kK;
    int split = 1;
    while (batch * ctiles * split < 2 * kNumSMs && split * kWarps < num_rois && split < 16) split *= 2;
    const size_t smem = plane_smem_bytes(height, width, esz);
    dim3 grid((unsigned)(batch * ctiles * split));
#define I2V_LAUNCH_POOL(T, PITCH)                                                                                     \
    do {                                                                                                              \
        auto kern = roi_pool_plane_kernel<T, PITCH>;                                                                  \
        I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
        kern<<<grid, kThreads, smem, stream>>>(features, rois, static_cast<T*>(out), batch, channels, height, width,  \
                                               num_rois, spatial_scale, ldo, split);                                  \
    } while (0)
    const bool narrow = pool_pitch_for(width) == 41;
    if (out_dtype == I2V_DT_BF16 && narrow) I2V_LAUNCH_POOL(__nv_bfloat16, 41);
    else if (out_dtype == I2V_DT_BF16) I2V_LAUNCH_POOL(__nv_bfloat16, 65);
    else if (narrow) I2V_LAUNCH_POOL(float, 41);
    else I2V_LAUNCH_POOL(float, 65);
#undef I2V_LAUNCH_POOL
    return check_launch("roi_pool_plane_kernel");
}

}  // namespace i2v
