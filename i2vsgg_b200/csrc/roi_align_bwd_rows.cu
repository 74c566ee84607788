// RoIAlign / RoIAlignAvg backward, row-owner variant (roi_align_kernel.cu:94-143 behind the pool's backward,
// modules/roi_align.py:18-29).
//
// One CTA per (frame, 16 channels) owns those 16 gradient planes in shared memory as [row][column][16 channels] and
// writes them to HBM once (no global atomics, no memset, deterministic).  What is new against lattice_bwd_plane_kernel:
//
//   * lanes are (left/right cell of a bilinear pair) x (16 channels): the two cells a lattice point touches in one
//     feature row are 128 contiguous bytes, so every read-modify-write is ONE conflict-free wavefront each way;
//   * consumer warp w owns the feature rows y = w (mod 8).  A RoI's contribution to row y,
//         G_y[pw] = sum_ph c_y[ph] * dL[ph][pw],   c_y[ph] = (1 - fy) if start_ph == y,  fy if start_ph + 1 == y,
//     is formed in registers first (both vertical neighbours merged), then scattered along x.  No two warps ever touch
//     the same cell, so there is nothing to synchronise between them; inside a warp, lattice columns whose cell pairs
//     may overlap are issued in separate batches (`mode`, decided per RoI by the prep kernel);
//   * two transform warps turn each pooled-gradient tile [16][49] (as TMA delivered it) into lattice gradients
//     dL[ph][pw] -- the 2x2/stride-1 average pool's backward, a few adds -- laid out [ph][pw/4][channel][4] so that a
//     consumer reads its channel's 8 lattice columns with two conflict-free 16-byte loads;
//   * a producer warp streams tiles and per-RoI tables through an 8-stage TMA ring.
#include "common.cuh"

namespace i2v {

struct alignas(16) RowTab {
    int y_lo, y_hi;            // feature rows the RoI touches (y_hi < y_lo: none)
    int mode;                  // 0: all lattice columns independent, 1: even / odd columns, 2: one column at a time
    unsigned valid_x;
    int xoff[8];               // start column * 64 bytes
    float wx0[8], wx1[8];      // weights of the left / right cell (validity and the avg pool's 1/4 folded in)
    int s[8];                  // start row of lattice row ph
    float f[8];                // its vertical fraction
    unsigned char rowrange[48];// per feature row: first | last << 4 lattice row contributing to it, 0xFF: none
};
static_assert(sizeof(RowTab) == 224, "RowTab layout");

namespace {

constexpr int kK = 16;
constexpr int kConsumers = 8;
constexpr int kTransformers = 2;
constexpr int kThreads = (kConsumers + kTransformers + 1) * 32;
constexpr int kStages = 8;
constexpr int kTileBytes = kK * 49 * 4;                 // 3136
constexpr int kTabBytes = (int)sizeof(RowTab);          // 224
constexpr int kDlBytes = 8 * 2 * kK * 16;               // 4096: [ph][pw/4][channel][4 floats]
constexpr int kStageBytes = 7488;                       // 3136 + 224 + 4096 = 7456, padded to 64 (mod 128): see transform
static_assert(kStageBytes >= kTileBytes + kTabBytes + kDlBytes && kStageBytes % 128 == 64, "stage layout");
constexpr int kMaxRows = 48;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------- per-RoI tables
__global__ void __launch_bounds__(128) rows_prep_kernel(const LatticeRoi* __restrict__ tab, RowTab* __restrict__ rtab,
                                                        int num_rois, int G, float wscale) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= num_rois) return;
    const LatticeRoi& t = tab[n];
    RowTab q;
    const unsigned full = (1u << G) - 1u;
    int y_lo = 1 << 30, y_hi = -1;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const bool oky = p < G && ((t.valid_y >> p) & 1u), okx = p < G && ((t.valid_x >> p) & 1u);
        q.s[p] = oky ? t.y.start[p] : -1000;
        q.f[p] = oky ? t.y.frac[p] : 0.f;
        if (oky) {
            y_lo = min(y_lo, t.y.start[p]);
            y_hi = max(y_hi, t.y.start[p] + 1);
        }
        q.xoff[p] = okx ? t.x.start[p] * 64 : 0;
        q.wx0[p] = okx ? (1.f - t.x.frac[p]) * wscale : 0.f;
        q.wx1[p] = okx ? t.x.frac[p] * wscale : 0.f;
    }
    q.valid_x = t.valid_x & full;
    q.y_lo = y_lo;
    q.y_hi = (t.batch >= 0 && q.valid_x != 0u) ? y_hi : -1;     // nothing to scatter: consumers skip the RoI
    // column batches: starts are non-decreasing, a bilinear pair covers [start, start + 1]
    int mode = 2;
    if (q.valid_x == full) {
        int d1 = 1 << 30, d2 = 1 << 30;
        for (int p = 0; p + 1 < G; ++p) d1 = min(d1, t.x.start[p + 1] - t.x.start[p]);
        for (int p = 0; p + 2 < G; ++p) d2 = min(d2, t.x.start[p + 2] - t.x.start[p]);
        mode = d1 >= 2 ? 0 : (d2 >= 2 ? 1 : 2);
    }
    q.mode = mode;
    for (int y = 0; y < kMaxRows; ++y) {
        int pa = -1, pb = -1;
        for (int p = 0; p < G; ++p) {
            if (!((t.valid_y >> p) & 1u)) continue;
            const int st = t.y.start[p];
            if (st == y || st + 1 == y) {
                if (pa < 0) pa = p;
                pb = p;
            }
        }
        q.rowrange[y] = pa < 0 ? (unsigned char)0xFF : (unsigned char)(pa | (pb << 4));
    }
    rtab[n] = q;
}

// ---------------------------------------------------------------------------------------------- the kernel
template <int POOL>
__global__ void __launch_bounds__(kThreads, 1)
    lattice_bwd_rows_kernel(const float* __restrict__ grad_out, const RowTab* __restrict__ rtab,
                            const int* __restrict__ order, const int* __restrict__ starts, float* __restrict__ grad_in,
                            int C, int H, int W) {
    constexpr int P = 7;
    constexpr int G = (POOL == I2V_POOL_NONE) ? P : P + 1;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* ring = smem;                                                        // [stages][kStageBytes]
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);    // [stages]
    uint64_t* dl_full = raw_full + kStages;
    uint64_t* empty = dl_full + kStages;
    float* planes = reinterpret_cast<float*>(empty + kStages);                         // [H][W][16]; 64-byte aligned

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ctiles = C / kK;
    const int ct = blockIdx.x % ctiles;
    const int b = blockIdx.x / ctiles;
    const int list_lo = __ldg(starts + b), list_hi = __ldg(starts + b + 1);
    const int count = list_hi - list_lo;
    const int HW = H * W;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(raw_full + s, 1);
            mbar_init(dl_full + s, 1);
            mbar_init(empty + s, kConsumers);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        float4* z = reinterpret_cast<float4*>(planes);
        for (int i = tid; i < HW * kK / 4; i += kThreads) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (warp == kConsumers + kTransformers) {
        // ---- producer: lane j feeds stage j ----
        if (lane < kStages) {
            unsigned char* dst = ring + lane * kStageBytes;
            unsigned round = 0;
            for (int k = lane; k < count; k += kStages, ++round) {
                const int n = __ldg(order + list_lo + k);
                if (round > 0) mbar_wait(empty + lane, (round - 1) & 1);
                mbar_expect_tx(raw_full + lane, kTileBytes + kTabBytes);
                bulk_load(dst, grad_out + ((size_t)n * C + (size_t)ct * kK) * 49, kTileBytes, raw_full + lane);
                bulk_load(dst + kTileBytes, rtab + n, kTabBytes, raw_full + lane);
            }
        }
    } else if (warp >= kConsumers) {
        // ---- transform: lanes = (stage of a pair) x (channel).  The stage pitch is 64 (mod 128) bytes, so the two halves
        // read opposite bank halves and each of the 49 loads is conflict free (channel pitch 49 words is odd) ----
        const int tw = warp - kConsumers;
        const int hs = lane >> 4, c = lane & 15;
        for (int k0 = 2 * tw; k0 < count; k0 += 2 * kTransformers) {
            const int k = k0 + hs;
            const bool have = k < count;
            const int s = k % kStages;
            const unsigned round = (unsigned)(k / kStages);
            if (have) mbar_wait(raw_full + s, round & 1);
            __syncwarp();
            if (have) {
                const float* tile = reinterpret_cast<const float*>(ring + s * kStageBytes) + c * 49;
                float4* dl = reinterpret_cast<float4*>(ring + s * kStageBytes + kTileBytes + kTabBytes) + c;
                if (POOL == I2V_POOL_NONE) {
#pragma unroll
                    for (int ph = 0; ph < 7; ++ph) {
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 7; ++j) v[j] = tile[ph * 7 + j];
                        v[7] = 0.f;
                        dl[(ph * 2 + 0) * kK] = make_float4(v[0], v[1], v[2], v[3]);
                        dl[(ph * 2 + 1) * kK] = make_float4(v[4], v[5], v[6], v[7]);
                    }
                } else {
                    float prev[7], cur[7];
#pragma unroll
                    for (int j = 0; j < 7; ++j) prev[j] = 0.f;
#pragma unroll
                    for (int ph = 0; ph < 8; ++ph) {
                        // lattice row ph collects pooled rows ph-1 and ph; lattice column pw pooled columns pw-1 and pw
#pragma unroll
                        for (int j = 0; j < 7; ++j) cur[j] = (ph < 7) ? tile[ph * 7 + j] : 0.f;
                        float rs[7], v[8];
#pragma unroll
                        for (int j = 0; j < 7; ++j) rs[j] = prev[j] + cur[j];
                        v[0] = rs[0];
#pragma unroll
                        for (int j = 1; j < 7; ++j) v[j] = rs[j - 1] + rs[j];
                        v[7] = rs[6];
                        dl[(ph * 2 + 0) * kK] = make_float4(v[0], v[1], v[2], v[3]);
                        dl[(ph * 2 + 1) * kK] = make_float4(v[4], v[5], v[6], v[7]);
#pragma unroll
                        for (int j = 0; j < 7; ++j) prev[j] = cur[j];
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(dl_full + (k0 % kStages));
            if (lane == 16 && have) mbar_arrive(dl_full + s);
        }
    } else {
        // ---- consumers: warp w owns the feature rows y = w (mod 8); lanes = (left / right cell) x (channel) ----
        const int dx = lane >> 4, c = lane & 15;
        const int row_bytes = W * 64;
        unsigned char* lane_planes = reinterpret_cast<unsigned char*>(planes) + lane * 4;
        for (int k = 0; k < count; ++k) {
            const int s = k % kStages;
            const unsigned round = (unsigned)(k / kStages);
            mbar_wait(dl_full + s, round & 1);
            const RowTab* t = reinterpret_cast<const RowTab*>(ring + s * kStageBytes + kTileBytes);
            const float4* dl = reinterpret_cast<const float4*>(ring + s * kStageBytes + kTileBytes + kTabBytes) + c;
            const int y_lo = t->y_lo, y_hi = t->y_hi;
            int y = y_lo + ((warp - y_lo) & (kConsumers - 1));      // first owned row >= y_lo
            if (y <= y_hi) {
                const int mode = t->mode;
                const unsigned vx = t->valid_x;
                const float* wxp = dx ? t->wx1 : t->wx0;
                const float4 wa = *reinterpret_cast<const float4*>(wxp), wb = *reinterpret_cast<const float4*>(wxp + 4);
                const int4 oa = *reinterpret_cast<const int4*>(t->xoff), ob = *reinterpret_cast<const int4*>(t->xoff + 4);
                const float wx[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                const int xo[8] = {oa.x, oa.y, oa.z, oa.w, ob.x, ob.y, ob.z, ob.w};
                for (; y <= y_hi; y += kConsumers) {
                    const unsigned rr = t->rowrange[y];
                    if (rr == 0xFFu) continue;
                    const int pa = rr & 15, pb = rr >> 4;
                    float g[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) g[i] = 0.f;
                    for (int ph = pa; ph <= pb; ++ph) {             // one or two lattice rows, more for tiny RoIs
                        const float fy = t->f[ph];
                        const float cy = (t->s[ph] == y) ? 1.f - fy : fy;
                        const float4 a0 = dl[(ph * 2 + 0) * kK], a1 = dl[(ph * 2 + 1) * kK];
                        g[0] = fmaf(cy, a0.x, g[0]); g[1] = fmaf(cy, a0.y, g[1]);
                        g[2] = fmaf(cy, a0.z, g[2]); g[3] = fmaf(cy, a0.w, g[3]);
                        g[4] = fmaf(cy, a1.x, g[4]); g[5] = fmaf(cy, a1.y, g[5]);
                        g[6] = fmaf(cy, a1.z, g[6]); g[7] = fmaf(cy, a1.w, g[7]);
                    }
                    unsigned char* rowp = lane_planes + (size_t)y * row_bytes;
                    if (mode == 0) {
                        float o[G];
#pragma unroll
                        for (int i = 0; i < G; ++i) o[i] = *reinterpret_cast<float*>(rowp + xo[i]);
#pragma unroll
                        for (int i = 0; i < G; ++i) *reinterpret_cast<float*>(rowp + xo[i]) = fmaf(g[i], wx[i], o[i]);
                        __syncwarp();
                    } else if (mode == 1) {
#pragma unroll
                        for (int par = 0; par < 2; ++par) {
                            float o[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (2 * i + par < G) o[i] = *reinterpret_cast<float*>(rowp + xo[2 * i + par]);
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (2 * i + par < G)
                                    *reinterpret_cast<float*>(rowp + xo[2 * i + par]) = fmaf(g[2 * i + par], wx[2 * i + par], o[i]);
                            __syncwarp();
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < G; ++i) {
                            if ((vx >> i) & 1u) {                   // uniform
                                float* q = reinterpret_cast<float*>(rowp + xo[i]);
                                *q = fmaf(g[i], wx[i], *q);
                            }
                            __syncwarp();
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
    }

    // ---- write-out: [cell][16] in shared memory -> [16][cell] in HBM, through a per-warp transpose buffer in the (now
    // idle) ring so that both the shared reads and the global stores are full 128-byte lines ----
    __syncthreads();
    if (warp < kConsumers) {
        constexpr int kTP = 34;                                        // transpose pitch: lanes (dx, c) hit bank 2c + dx
        float* tb = reinterpret_cast<float*>(ring + warp * kStageBytes);   // [16][34]
        const int dx = lane >> 4, c = lane & 15;
        float* dst = grad_in + ((size_t)b * C + (size_t)ct * kK) * HW;
        for (int i0 = warp * 32; i0 < HW; i0 += kConsumers * 32) {
#pragma unroll 4
            for (int j = 0; j < 32; j += 2) {
                const int cell = i0 + j + dx;
                tb[c * kTP + j + dx] = cell < HW ? planes[(size_t)cell * kK + c] : 0.f;
            }
            __syncwarp();
#pragma unroll 4
            for (int ch = 0; ch < kK; ++ch)
                if (i0 + lane < HW) dst[(size_t)ch * HW + i0 + lane] = tb[ch * kTP + lane];
            __syncwarp();
        }
    }
}

}  // namespace

size_t bwd_rows_smem_bytes(int H, int W) {
    return (size_t)kStages * kStageBytes + 3 * kStages * sizeof(uint64_t) + (size_t)H * W * kK * sizeof(float);
}

bool bwd_rows_ok(const float* grad_out, int batch, int C, int H, int W, int PH, int PW, int pool_mode) {
    return batch > 0 && PH == 7 && PW == 7 && pool_mode != I2V_POOL_MAX && C % kK == 0 && H >= 2 && W >= 2 &&
           H <= kMaxRows && bwd_rows_smem_bytes(H, W) <= (size_t)kMaxSmemPerCta && ((uintptr_t)grad_out & 15) == 0;
}

// `tab` holds the LatticeRoi tables of this call; `rtab_space` is the (>= 224 bytes per RoI) slot for the RowTab views.
int launch_bwd_rows(const float* grad_out, const LatticeRoi* tab, void* rtab_space, const int* order, const int* starts,
                    float* grad_in, int batch, int C, int H, int W, int num_rois, int pool_mode, cudaStream_t stream) {
    RowTab* rtab = static_cast<RowTab*>(rtab_space);
    const int G = pool_mode == I2V_POOL_NONE ? 7 : 8;
    rows_prep_kernel<<<ceil_div(num_rois, 128), 128, 0, stream>>>(tab, rtab, num_rois, G,
                                                                   pool_mode == I2V_POOL_AVG ? 0.25f : 1.f);
    I2V_TRY(check_launch("rows_prep_kernel"));
    const size_t smem = bwd_rows_smem_bytes(H, W);
    dim3 grid((unsigned)(batch * (C / kK)));
    if (pool_mode == I2V_POOL_AVG) {
        auto kern = lattice_bwd_rows_kernel<I2V_POOL_AVG>;
        I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kThreads, smem, stream>>>(grad_out, rtab, order, starts, grad_in, C, H, W);
    } else {
        auto kern = lattice_bwd_rows_kernel<I2V_POOL_NONE>;
        I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kThreads, smem, stream>>>(grad_out, rtab, order, starts, grad_in, C, H, W);
    }
    return check_launch("lattice_bwd_rows_kernel");
}

}  // namespace i2v
