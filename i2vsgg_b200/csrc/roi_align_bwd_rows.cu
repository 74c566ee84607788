// RoIAlign / RoIAlignAvg backward, row-owner variant (roi_align_kernel.cu:94-143 behind the pool's backward,
// modules/roi_align.py:18-29).
//
// One CTA per (frame, 16 channels) owns those 16 gradient planes in shared memory as [row][column][16 channels] and
// writes them to HBM once (no global atomics, no memset, deterministic).  What is new against lattice_bwd_plane_kernel:
//
//   * lanes are (lattice column pw or pw+4) x (left/right cell of the bilinear pair) x (channel pair): the two cells a
//     lattice point touches in one feature row are 128 contiguous bytes = one half-warp of 8-byte accesses, so every
//     read-modify-write is conflict free by construction, and one instruction serves two lattice points;
//   * consumer warp w owns the feature rows y = w (mod 16).  A RoI's contribution to row y,
//         G_y[pw] = sum_ph c_y[ph] * dL[ph][pw],   c_y[ph] = (1 - fy) if start_ph == y,  fy if start_ph + 1 == y,
//     is formed in registers first (both vertical neighbours merged), then scattered along x.  No two warps ever touch
//     the same cell, so there is nothing to synchronise between them; inside a warp, lattice columns whose cell pairs
//     may overlap are issued in separate batches (`mode`, decided per RoI by the prep kernel);
//   * four transform warps turn each pooled-gradient tile [16][49] (as TMA delivered it) into lattice gradients
//     dL[ph][pw] -- the 2x2/stride-1 average pool's backward, a few adds -- laid out so that a consumer lane reads its
//     four lattice columns x two channels with two conflict-free 16-byte loads;
//   * a producer warp streams tiles and per-RoI tables through a 9-stage TMA ring; eight lattice slots let the consumer warps drift apart.
#include "common.cuh"

namespace i2v {

struct alignas(16) RowTab {
    int y_lo, y_hi;            // feature rows the RoI touches (y_hi < y_lo: none)
    int mode;                  // 0: all lattice columns independent, 1: even / odd, 2: (i, i+4) pairs in turn, 3: one by one
    unsigned valid_x;
    int xoff[8];               // start column * 64 bytes
    float wx0[8], wx1[8];      // weights of the left / right cell (validity and the avg pool's 1/4 folded in)
    int s[8];                  // start row of lattice row ph
    float f[8];                // its vertical fraction
    unsigned char rowrange[48];// per feature row: first lattice row contributing to it | (number of them) << 4, 0xFF: none
    float rc[48][2];           // the coefficients of the first two of them
};
static_assert(sizeof(RowTab) == 608 && sizeof(RowTab) <= kRoiTabSlotBytes, "RowTab layout");

namespace {

constexpr int kK = 16;
constexpr int kConsumers = 16;
constexpr int kTransformers = 4;
constexpr int kThreads = (kConsumers + kTransformers + 1) * 32;
constexpr int kRawStages = 9;                           // TMA ring: tile + table, 34 KB in flight per SM
constexpr int kDlSlots = 8;                             // lattice gradients + table: how far consumer warps may drift apart
constexpr int kTileBytes = kK * 49 * 4;                 // 3136
constexpr int kTabBytes = (int)sizeof(RowTab);          // 608
constexpr int kRawBytes = kTileBytes + kTabBytes;       // 3744
constexpr int kDlBytes = 8 * 2 * 2 * 8 * 16;            // 4096: [ph][pw/4][(pw%4)/2][channel pair][(pw%2, channel%2)]
constexpr int kSlotBytes = kTabBytes + kDlBytes;        // 4704
constexpr int kRingBytes = kRawStages * kRawBytes + kDlSlots * kSlotBytes;   // 60000
constexpr int kNumBars = 2 * kRawStages + 2 * kDlSlots;
static_assert(kRawBytes % 16 == 0 && kSlotBytes % 16 == 0 && kRingBytes % 16 == 0, "ring layout");
static_assert(kConsumers * 16 * 34 * 4 <= kRingBytes, "transpose buffers");
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
    // column batches: starts are non-decreasing, a bilinear pair covers [start, start + 1].  One instruction serves the
    // lattice columns i and i + 4, so they must be at least two cells apart (d4) for any of the fast modes.
    int mode = 3;
    if (q.valid_x == full) {
        int d1 = 1 << 30, d2 = 1 << 30, d4 = 1 << 30;
        for (int p = 0; p + 1 < G; ++p) d1 = min(d1, t.x.start[p + 1] - t.x.start[p]);
        for (int p = 0; p + 2 < G; ++p) d2 = min(d2, t.x.start[p + 2] - t.x.start[p]);
        for (int p = 0; p + 4 < G; ++p) d4 = min(d4, t.x.start[p + 4] - t.x.start[p]);
        mode = d1 >= 2 ? 0 : (d2 >= 2 ? 1 : (d4 >= 2 ? 2 : 3));
    }
    q.mode = mode;
    for (int y = 0; y < kMaxRows; ++y) {
        int pa = -1, cnt = 0;
        float c0 = 0.f, c1 = 0.f;
        for (int p = 0; p < G; ++p) {
            if (!((t.valid_y >> p) & 1u)) continue;
            const int st = t.y.start[p];
            if (st == y || st + 1 == y) {
                const float cy = (st == y) ? 1.f - t.y.frac[p] : t.y.frac[p];
                if (pa < 0) pa = p;
                if (cnt == 0) c0 = cy;
                if (cnt == 1) c1 = cy;
                ++cnt;
            }
        }
        q.rowrange[y] = pa < 0 ? (unsigned char)0xFF : (unsigned char)(pa | (cnt << 4));
        q.rc[y][0] = c0;
        q.rc[y][1] = c1;
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
    unsigned char* raw = smem;                                                         // [kRawStages][tile | table]
    unsigned char* slots = smem + kRawStages * kRawBytes;                              // [kDlSlots][table | dl]
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(smem + kRingBytes);               // [kRawStages]
    uint64_t* raw_empty = raw_full + kRawStages;
    uint64_t* dl_full = raw_empty + kRawStages;                                        // [kDlSlots]
    uint64_t* dl_empty = dl_full + kDlSlots;
    float* planes = reinterpret_cast<float*>(smem + kRingBytes + ((kNumBars * 8 + 63) / 64) * 64);   // [H][W][16]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ctiles = C / kK;
    const int ct = blockIdx.x % ctiles;
    const int b = blockIdx.x / ctiles;
    const int list_lo = __ldg(starts + b), list_hi = __ldg(starts + b + 1);
    const int count = list_hi - list_lo;
    const int HW = H * W;

    if (tid == 0) {
        for (int s = 0; s < kRawStages; ++s) {
            mbar_init(raw_full + s, 1);
            mbar_init(raw_empty + s, 1);
        }
        for (int s = 0; s < kDlSlots; ++s) {
            mbar_init(dl_full + s, 1);
            mbar_init(dl_empty + s, kConsumers);
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
        // ---- producer: lane j feeds raw stage j ----
        if (lane < kRawStages) {
            unsigned char* dst = raw + lane * kRawBytes;
            unsigned round = 0;
            for (int k = lane; k < count; k += kRawStages, ++round) {
                const int n = __ldg(order + list_lo + k);
                if (round > 0) mbar_wait(raw_empty + lane, (round - 1) & 1);
                mbar_expect_tx(raw_full + lane, kRawBytes);
                bulk_load(dst, grad_out + ((size_t)n * C + (size_t)ct * kK) * 49, kTileBytes, raw_full + lane);
                bulk_load(dst + kTileBytes, rtab + n, kTabBytes, raw_full + lane);
            }
        }
    } else if (warp >= kConsumers) {
        // ---- transform: warp tw serves the items k = tw (mod 4) and owns lattice slot tw; lanes = (lattice rows 0-3 / 4-7)
        // x (channel).  Each lane reads the four pooled rows 3 hf .. 3 hf + 3 of its channel and emits four lattice rows;
        // the RoI's table moves along into the slot, so the raw stage is free for the next TMA load right away ----
        const int tw = warp - kConsumers;
        const int hf = lane >> 4, c = lane & 15;
        for (int k = tw; k < count; k += kTransformers) {
            const int rs_ = k % kRawStages;
            const int sl = k & (kDlSlots - 1);
            const unsigned use = (unsigned)(k / kDlSlots);
            unsigned char* slot = slots + sl * kSlotBytes;
            mbar_wait(raw_full + rs_, (unsigned)(k / kRawStages) & 1u);
            if (use > 0) mbar_wait(dl_empty + sl, (use - 1) & 1u);
            const float* tile = reinterpret_cast<const float*>(raw + rs_ * kRawBytes) + c * 49 + hf * 21;
            {
                const float4* ts = reinterpret_cast<const float4*>(raw + rs_ * kRawBytes + kTileBytes);
                float4* td = reinterpret_cast<float4*>(slot);
                for (int i = lane; i < kTabBytes / 16; i += 32) td[i] = ts[i];
            }
            // dl[ph][pw/4][(pw%4)/2][channel pair][(pw%2) * 2 + channel%2]: what a consumer lane (pt, dx, cp) reads with
            // two 16-byte loads
            float* dlc = reinterpret_cast<float*>(slot + kTabBytes) + (c >> 1) * 4 + (c & 1) + hf * 4 * 128;
            auto put_row = [&](int k4, const float (&v)[8]) {      // lattice row 4 hf + k4
#pragma unroll
                for (int pw = 0; pw < 8; ++pw)
                    dlc[(((k4 * 2 + (pw >> 2)) * 2 + ((pw & 3) >> 1)) * 8) * 4 + (pw & 1) * 2] = v[pw];
            };
            float a[4][7];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < 7; ++j) a[r][j] = tile[r * 7 + j];
            if (POOL == I2V_POOL_NONE) {
                // lattice == pooled grid (7 x 7): lattice row 4 hf + k4 is pooled row 4 hf + k4 = a[k4 + hf] (row 7: none)
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 7; ++j) v[j] = hf ? (k4 < 3 ? a[(k4 + 1) & 3][j] : 0.f) : a[k4][j];
                    v[7] = 0.f;
                    put_row(k4, v);
                }
            } else {
                // lattice row ph collects the pooled rows ph-1 and ph, lattice column pw the pooled columns pw-1 and pw
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    float rs[7], v[8];
#pragma unroll
                    for (int j = 0; j < 7; ++j) {
                        const float lo = (k4 == 0) ? a[0][j] : a[k4 - 1][j] + a[k4][j];                 // hf == 0
                        const float hi = (k4 == 3) ? a[3][j] : a[k4][j] + a[k4 + 1][j];                 // hf == 1
                        rs[j] = hf ? hi : lo;
                    }
                    v[0] = rs[0];
#pragma unroll
                    for (int j = 1; j < 7; ++j) v[j] = rs[j - 1] + rs[j];
                    v[7] = rs[6];
                    put_row(k4, v);
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(raw_empty + rs_);
                mbar_arrive(dl_full + sl);
            }
        }
    } else {
        // ---- consumers: warp w owns the feature rows y = w (mod 16).  lane = (pt, dx, cp): lattice columns 4 pt + i
        // (i = 0..3, one instruction each), left / right cell, channel pair ----
        const int pt = lane >> 4, dx = (lane >> 3) & 1, cp = lane & 7;
        const int row_bytes = W * 64;
        unsigned char* lane_planes = reinterpret_cast<unsigned char*>(planes) + (dx * 16 + cp * 2) * 4;
        for (int k = 0; k < count; ++k) {
            const int s = k & (kDlSlots - 1);
            mbar_wait(dl_full + s, (unsigned)(k / kDlSlots) & 1u);
            const RowTab* t = reinterpret_cast<const RowTab*>(slots + s * kSlotBytes);
            const int2 yy = *reinterpret_cast<const int2*>(&t->y_lo);
            int y = yy.x + ((warp - yy.x) & (kConsumers - 1));      // first owned row >= y_lo
            if (y <= yy.y) {
                // this lane's [ph][pt] slice of the lattice gradients: + ((ph * 2 + pt) * 2 + h) * 8 float4
                const float4* dl = reinterpret_cast<const float4*>(slots + s * kSlotBytes + kTabBytes) + pt * 16 + cp;
                const int2 mv = *reinterpret_cast<const int2*>(&t->mode);
                const int mode = mv.x;
                const unsigned vx = (unsigned)mv.y;
                const float4 w4 = *reinterpret_cast<const float4*>((dx ? t->wx1 : t->wx0) + pt * 4);
                const int4 o4 = *reinterpret_cast<const int4*>(t->xoff + pt * 4);
                const float wx[4] = {w4.x, w4.y, w4.z, w4.w};
                const int xo[4] = {o4.x, o4.y, o4.z, o4.w};
                for (; y <= yy.y; y += kConsumers) {
                    const unsigned rr = t->rowrange[y];
                    if (rr == 0xFFu) continue;
                    const int pa = rr & 15, cnt = rr >> 4;
                    const float2 cc = *reinterpret_cast<const float2*>(t->rc[y]);
                    float2 g[4];
                    {
                        const float4 a0 = dl[pa * 32], a1 = dl[pa * 32 + 8];
                        g[0] = make_float2(cc.x * a0.x, cc.x * a0.y);
                        g[1] = make_float2(cc.x * a0.z, cc.x * a0.w);
                        g[2] = make_float2(cc.x * a1.x, cc.x * a1.y);
                        g[3] = make_float2(cc.x * a1.z, cc.x * a1.w);
                    }
                    if (cnt > 1) {
                        const float4 a0 = dl[(pa + 1) * 32], a1 = dl[(pa + 1) * 32 + 8];
                        g[0].x = fmaf(cc.y, a0.x, g[0].x); g[0].y = fmaf(cc.y, a0.y, g[0].y);
                        g[1].x = fmaf(cc.y, a0.z, g[1].x); g[1].y = fmaf(cc.y, a0.w, g[1].y);
                        g[2].x = fmaf(cc.y, a1.x, g[2].x); g[2].y = fmaf(cc.y, a1.y, g[2].y);
                        g[3].x = fmaf(cc.y, a1.z, g[3].x); g[3].y = fmaf(cc.y, a1.w, g[3].y);
                        for (int ph = pa + 2; ph < pa + cnt; ++ph) {        // tiny RoIs: more lattice rows per cell
                            const float fy = t->f[ph];
                            const float cy = (t->s[ph] == y) ? 1.f - fy : fy;
                            const float4 b0 = dl[ph * 32], b1 = dl[ph * 32 + 8];
                            g[0].x = fmaf(cy, b0.x, g[0].x); g[0].y = fmaf(cy, b0.y, g[0].y);
                            g[1].x = fmaf(cy, b0.z, g[1].x); g[1].y = fmaf(cy, b0.w, g[1].y);
                            g[2].x = fmaf(cy, b1.x, g[2].x); g[2].y = fmaf(cy, b1.y, g[2].y);
                            g[3].x = fmaf(cy, b1.z, g[3].x); g[3].y = fmaf(cy, b1.w, g[3].y);
                        }
                    }
                    unsigned char* rowp = lane_planes + (size_t)y * row_bytes;
                    auto rmw = [&](int i, float2 o) {
                        *reinterpret_cast<float2*>(rowp + xo[i]) = make_float2(fmaf(g[i].x, wx[i], o.x), fmaf(g[i].y, wx[i], o.y));
                    };
                    const bool last_ok = (G == 8) || pt == 0;       // a 7-point lattice has no column 4 + 3
                    if (mode == 0) {
                        float2 o[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) o[i] = *reinterpret_cast<const float2*>(rowp + xo[i]);
#pragma unroll
                        for (int i = 0; i < 3; ++i) rmw(i, o[i]);
                        if (last_ok) rmw(3, o[3]);
                        __syncwarp();
                    } else if (mode == 1) {
                        float2 o0 = *reinterpret_cast<const float2*>(rowp + xo[0]), o2 = *reinterpret_cast<const float2*>(rowp + xo[2]);
                        rmw(0, o0);
                        rmw(2, o2);
                        __syncwarp();
                        float2 o1 = *reinterpret_cast<const float2*>(rowp + xo[1]), o3 = *reinterpret_cast<const float2*>(rowp + xo[3]);
                        rmw(1, o1);
                        if (last_ok) rmw(3, o3);
                        __syncwarp();
                    } else if (mode == 2) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float2 o = *reinterpret_cast<const float2*>(rowp + xo[i]);
                            if (i < 3 || last_ok) rmw(i, o);
                            __syncwarp();
                        }
                    } else {
#pragma unroll
                        for (int pw = 0; pw < G; ++pw) {
                            if (((vx >> pw) & 1u) && pt == (pw >> 2)) {
                                float2 o = *reinterpret_cast<const float2*>(rowp + xo[pw & 3]);
                                rmw(pw & 3, o);
                            }
                            __syncwarp();
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(dl_empty + s);
        }
    }

    // ---- write-out: [cell][16] in shared memory -> [16][cell] in HBM, through a per-warp transpose buffer in the (now
    // idle) ring so that both the shared reads and the global stores are full 128-byte lines ----
    __syncthreads();
    if (warp < kConsumers) {
        constexpr int kTP = 34;                                        // transpose pitch: lanes (dx, c) hit bank 2c + dx
        float* tb = reinterpret_cast<float*>(smem + warp * (16 * 34 * 4));   // [16][34]
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
    return (size_t)kRingBytes + ((kNumBars * 8 + 63) / 64) * 64 + (size_t)H * W * kK * sizeof(float);
}

bool bwd_rows_ok(const float* grad_out, int batch, int C, int H, int W, int PH, int PW, int pool_mode) {
    return batch > 0 && PH == 7 && PW == 7 && pool_mode != I2V_POOL_MAX && C % kK == 0 && H >= 2 && W >= 2 &&
           H <= kMaxRows && bwd_rows_smem_bytes(H, W) <= (size_t)kMaxSmemPerCta && ((uintptr_t)grad_out & 15) == 0;
}

// `tab` holds the LatticeRoi tables of this call; `rtab_space` is the workspace's per-RoI table slot (kRoiTabSlotBytes each).
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
