// RoIAlign / RoIAlignAvg backward, band-owner variant (roi_align_kernel.cu:94-143 behind the pool's backward,
// modules/roi_align.py:18-29).
//
// What the phased kernel (roi_align_bwd_phase.cu) pays for: lanes are (cell, 16 channels), so every pooled value and
// every table word is fetched by two half-warps (half of each shared-memory wavefront is a duplicate), and the RoIs
// scatter strictly one after the other (a named-barrier chain, 760 cycles per RoI against 256 cycles of
// read-modify-write wavefronts).  Here:
//
//   * lanes are 32 CHANNELS.  A cell of the gradient planes is 128 contiguous bytes, so every read-modify-write is one
//     full wavefront whatever the cell, every load of the pooled-gradient tile ([32][49] floats, lane stride 49 words =
//     32 distinct banks) is a full wavefront of distinct bytes, and all index math, table loads and branches are
//     warp-uniform;
//   * 32 planes of a 38x63 map are 306 KB, so a CTA owns a SLAB of feature rows (19 of 38) of one (frame, 32 channels);
//     the prep kernels list, per (frame, slab) and in RoI order, the RoIs whose lattice touches the slab, and the CTA
//     streams exactly those tiles through a TMA ring.  A RoI that spans two slabs is fetched by both CTAs (adjacent
//     block indices, so the second fetch is an L2 hit);
//   * inside the CTA every consumer warp OWNS a band of the slab's rows.  A warp walks the RoIs in list order and adds,
//     for every lattice row whose upper or lower feature row lies in its band, that row's contribution.  No two warps
//     ever touch the same cell, so there is no ordering between warps at all -- no barrier, no hand-off chain; the ring
//     lets a warp run up to `stages` RoIs ahead of the slowest one -- and the accumulation order of every cell is the
//     list order: the result is bit-reproducible;
//   * lattice columns two apart never share a cell unless the RoI is narrower than a cell per bin, so the eight columns
//     of a feature row go as two static passes (even, odd) of hoisted loads / FFMA2 / stores; the rare narrow RoIs take
//     a column-by-column path.  Columns off the map point at two padding cells behind every plane row with weight zero.
#include <stdlib.h>

#include "common.cuh"

namespace i2v {

struct alignas(16) BandTab {
    float2 yw[8];             // per lattice row: {1 - fy, fy} (weights of the upper / lower feature row)
    float2 xw[8];             // per lattice column: weights of the {left, right} cell (avg pool's 1/4 folded in; 0 if off the map)
    unsigned short xoff[8];   // byte offset of the left cell inside a plane row (column * 128; W * 128 = padding if off the map)
    unsigned char ys[8];      // per lattice row: its upper feature row; 0xFF: row off the map
    int xmode;                // 0: even / odd column passes are conflict free; 1: columns one by one
    short row_lo, row_hi;     // feature rows touched by the RoI (row_lo > row_hi: nothing to scatter)
};
static_assert(sizeof(BandTab) == 160 && sizeof(BandTab) <= kRoiTabSlotBytes, "BandTab layout");

constexpr int kBandMaxSlabs = 8;
constexpr int kBandMaxWarps = 16;       // consumer warps (row bands) per CTA

namespace {

constexpr int kK = 32;                                  // channels per CTA = lanes
constexpr int kTileBytes = kK * 49 * 4;                 // 6272
constexpr int kTabBytes = (int)sizeof(BandTab);         // 160
constexpr int kStageBytes = kTileBytes + kTabBytes;     // 6432
constexpr int kMaxStages = 16;
constexpr int kMinStages = 6;
constexpr int kBarBytes = 2 * kMaxStages * 8;           // 256
static_assert(kStageBytes % 16 == 0, "ring layout");

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load(unsigned sdst, const void* gsrc, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst),
                 "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
}
// All shared-memory traffic of the consumers goes through volatile asm so that it is issued in exactly the order
// written below (loads of a pass hoisted above its arithmetic, next loads above the current stores where legal).
template <int OFF>
__device__ __forceinline__ float lds(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts(unsigned addr, float v) {
    asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "f"(v) : "memory");
}
__device__ __forceinline__ void lds_pair2(unsigned addr, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr) : "memory");
}
__device__ __forceinline__ float2 lds_f2(unsigned addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds_u4(unsigned addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned lds_u8(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
// {d0, d1} = w (a packed pair of weights) * {t, t} + {c0, c1}: one packed fp32x2 FMA
__device__ __forceinline__ void ffma2(float& d0, float& d1, unsigned long long w, float t, float c0, float c1) {
    asm("{.reg .b64 rb, rc, rd;\n"
        "mov.b64 rb, {%3, %3};\n"
        "mov.b64 rc, {%4, %5};\n"
        "fma.rn.f32x2 rd, %2, rb, rc;\n"
        "mov.b64 {%0, %1}, rd;}"
        : "=f"(d0), "=f"(d1)
        : "l"(w), "f"(t), "f"(c0), "f"(c1));
}

// ---------------------------------------------------------------------------------------------- per-RoI tables
__global__ void __launch_bounds__(128) band_prep_kernel(const LatticeRoi* __restrict__ tab, unsigned char* __restrict__ tab_space,
                                                        int num_rois, int G, int W, float wscale) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= num_rois) return;
    const LatticeRoi& t = tab[n];
    BandTab q;
    const unsigned full = (1u << G) - 1u;
    const unsigned vx = t.valid_x & full, vy = t.valid_y & full;
    int xmode = 0;
    int prev[2] = {-100, -100};
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const bool okx = (vx >> p) & 1u;
        q.xoff[p] = (unsigned short)((okx ? t.x.start[p] : W) * 128);
        q.xw[p] = okx ? make_float2((1.f - t.x.frac[p]) * wscale, t.x.frac[p] * wscale) : make_float2(0.f, 0.f);
        if (okx) {
            // two columns of one pass may not share a cell: their start cells have to be two apart
            if (t.x.start[p] - prev[p & 1] < 2) xmode = 1;
            prev[p & 1] = t.x.start[p];
        }
    }
    const bool any = t.batch >= 0 && vx != 0u && vy != 0u;
    int lo = 1 << 14, hi = -1;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const bool oky = any && ((vy >> p) & 1u);
        q.yw[p] = oky ? make_float2(1.f - t.y.frac[p], t.y.frac[p]) : make_float2(0.f, 0.f);
        q.ys[p] = (unsigned char)(oky ? t.y.start[p] : 0xFF);
        if (oky) {
            lo = min(lo, t.y.start[p]);
            hi = max(hi, t.y.start[p] + 1);
        }
    }
    q.row_lo = (short)(any ? lo : 1);
    q.row_hi = (short)(any ? hi : 0);
    q.xmode = xmode;
    *reinterpret_cast<BandTab*>(tab_space + (size_t)n * kRoiTabSlotBytes) = q;
}

// One CTA per (frame, slab): the frame's RoIs (ascending, `order`) whose rows meet the slab, in the same order, at
// sorder[slab * num_rois + starts[frame] ...]; their number in scount[frame * nslabs + slab]; and the slab's rows cut
// into `nbands` bands of about equal work (row touches of those RoIs), bounds[(frame * nslabs + slab) * (kBandMaxWarps + 1) ...].
__global__ void __launch_bounds__(256) band_bucket_kernel(const unsigned char* __restrict__ tab_space,
                                                          const int* __restrict__ order, const int* __restrict__ starts,
                                                          int* __restrict__ sorder, int* __restrict__ scount,
                                                          int* __restrict__ bounds, int num_rois, int nslabs, int slab_rows,
                                                          int H, int nbands) {
    __shared__ int s_warp[8];
    __shared__ int s_hits[256];
    const int b = blockIdx.x / nslabs, sl = blockIdx.x % nslabs;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lo = starts[b], hi = starts[b + 1];
    const int r0 = sl * slab_rows, r1 = min(H, r0 + slab_rows);
    const int count = hi - lo;
    const int chunk = (count + 255) / 256;
    const int a = lo + min(tid * chunk, count), e = lo + min(tid * chunk + chunk, count);
    s_hits[tid] = 0;
    __syncthreads();
    int mine = 0;
    for (int i = a; i < e; ++i) {
        const int n = __ldg(order + i);
        const BandTab* t = reinterpret_cast<const BandTab*>(tab_space + (size_t)n * kRoiTabSlotBytes);
        if (t->row_lo < r1 && t->row_hi >= r0) {
            ++mine;
            for (int p = 0; p < 8; ++p) {
                const int y = t->ys[p];
                if (y != 0xFF) {
                    atomicAdd(&s_hits[y], 1);
                    atomicAdd(&s_hits[y + 1], 1);
                }
            }
        }
    }
    int incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < 8; ++w) {
        if (w < warp) before += s_warp[w];
        total += s_warp[w];
    }
    int pos = sl * num_rois + lo + before + incl - mine;
    for (int i = a; i < e; ++i) {
        const int n = __ldg(order + i);
        const BandTab* t = reinterpret_cast<const BandTab*>(tab_space + (size_t)n * kRoiTabSlotBytes);
        if (t->row_lo < r1 && t->row_hi >= r0) sorder[pos++] = n;
    }
    if (tid == 0) {
        scount[b * nslabs + sl] = total;
        // band k = rows [bnd[k], bnd[k+1]): the cut nearest to k / nbands of the slab's row touches (every row costs at
        // least one unit so that an idle slab is cut evenly)
        int* bnd = bounds + (size_t)(b * nslabs + sl) * (kBandMaxWarps + 1);
        int all = 0;
        for (int r = r0; r < r1; ++r) all += s_hits[r] + 1;
        int r = r0, acc = 0;
        bnd[0] = r0;
        for (int k = 1; k < nbands; ++k) {
            const int target = (int)(((long long)all * k) / nbands);
            while (r < r1 && acc + (s_hits[r] + 1) / 2 < target) {
                acc += s_hits[r] + 1;
                ++r;
            }
            bnd[k] = r;
        }
        bnd[nbands] = r1;
    }
}

// ---------------------------------------------------------------------------------------------- the kernel
// The cells of four lattice columns J0, J0+2, J0+4, J0+6 in the upper (U) and / or lower (L) feature row under one
// lattice row += d[j] * wy * {wl, wr}[j].  a[j] is the address of column j's left cell in the upper row; RB (the plane
// row pitch in bytes) is an immediate.  All loads are issued before the first store.
template <int J0, bool U, bool L, int RB>
__device__ __forceinline__ void column_pass(const unsigned (&a)[8], const unsigned long long (&w)[8], const float (&d)[8],
                                            float wy0, float wy1) {
    float ou[4][2], ol[4][2];
    if (U) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ou[k][0] = lds<0>(a[J0 + 2 * k]);
            ou[k][1] = lds<128>(a[J0 + 2 * k]);
        }
    }
    if (L) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ol[k][0] = lds<RB>(a[J0 + 2 * k]);
            ol[k][1] = lds<RB + 128>(a[J0 + 2 * k]);
        }
    }
    if (U) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = J0 + 2 * k;
            float n0, n1;
            ffma2(n0, n1, w[j], d[j] * wy0, ou[k][0], ou[k][1]);
            sts<0>(a[j], n0);
            sts<128>(a[j], n1);
        }
    }
    if (L) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = J0 + 2 * k;
            float n0, n1;
            ffma2(n0, n1, w[j], d[j] * wy1, ol[k][0], ol[k][1]);
            sts<RB>(a[j], n0);
            sts<RB + 128>(a[j], n1);
        }
    }
}
// one column at a time (RoIs narrower than a cell per bin)
template <bool U, bool L, int RB>
__device__ __forceinline__ void column_single(unsigned a, unsigned long long w, float d, float wy0, float wy1) {
    float ou0 = 0.f, ou1 = 0.f, ol0 = 0.f, ol1 = 0.f;
    if (U) {
        ou0 = lds<0>(a);
        ou1 = lds<128>(a);
    }
    if (L) {
        ol0 = lds<RB>(a);
        ol1 = lds<RB + 128>(a);
    }
    if (U) {
        float n0, n1;
        ffma2(n0, n1, w, d * wy0, ou0, ou1);
        sts<0>(a, n0);
        sts<128>(a, n1);
    }
    if (L) {
        float n0, n1;
        ffma2(n0, n1, w, d * wy1, ol0, ol1);
        sts<RB>(a, n0);
        sts<RB + 128>(a, n1);
    }
}
template <bool U, bool L, int RB>
__device__ __forceinline__ void lattice_row(const unsigned (&a)[8], const unsigned long long (&w)[8], const float (&d)[8],
                                            float wy0, float wy1, bool narrow) {
    if (!narrow) {
        column_pass<0, U, L, RB>(a, w, d, wy0, wy1);
        column_pass<1, U, L, RB>(a, w, d, wy0, wy1);
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) column_single<U, L, RB>(a[j], w[j], d[j], wy0, wy1);
    }
}

// WPT: plane row pitch in cells (W + 2) as a compile-time constant, so that the lower row of a lattice row is an
// immediate offset; 0: any width (the two rows are then handled one after the other)
constexpr int kProducers = 4;       // producer warps: one thread each walks every kProducers-th RoI of the list (the
                                    // wait / expect / two bulk copies of one RoI take a thread ~450 cycles)

template <int POOL, int NB, int WPT>
__global__ void __launch_bounds__((NB + kProducers) * 32, 1)
    lattice_bwd_band_kernel(const float* __restrict__ grad_out, const unsigned char* __restrict__ tab_space,
                            const int* __restrict__ sorder, const int* __restrict__ scount, const int* __restrict__ starts,
                            const int* __restrict__ bounds, float* __restrict__ grad_in, int C, int H, int Wrt, int nslabs,
                            int slab_rows, int stages, int num_rois, int debug) {
    constexpr int P = 7;
    constexpr int kThreads = (NB + kProducers) * 32;
    const int W = WPT ? WPT - 2 : Wrt;
    const int WP = W + 2;                                    // two padding cells behind every plane row
    const int RB = WP * 128;                                 // plane row pitch in bytes
    extern __shared__ __align__(128) unsigned char smem[];
    const unsigned ring = smem_u32(smem);                                                // [stages][tile | table]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * kStageBytes);
    const unsigned full = smem_u32(bars), empty = full + kMaxStages * 8;
    float* planes = reinterpret_cast<float*>(smem + (size_t)stages * kStageBytes + kBarBytes);   // [rows][WP][32]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ctiles = C / kK;
    const int sl = blockIdx.x % nslabs;
    const int ct = (blockIdx.x / nslabs) % ctiles;
    const int b = blockIdx.x / (nslabs * ctiles);
    const int s0 = sl * slab_rows, s1 = min(H, s0 + slab_rows), nrows = s1 - s0;
    const int count = __ldg(scount + b * nslabs + sl);
    const int* list = sorder + (size_t)sl * num_rois + __ldg(starts + b);

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(bars + s, 1);
            mbar_init(bars + kMaxStages + s, NB);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        float4* z = reinterpret_cast<float4*>(planes);
        const int n4 = nrows * WP * (kK / 4);
        for (int i = tid; i < n4; i += kThreads) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (warp >= NB) {
        // ---- producers: warp NB + p takes the list entries k = p (mod kProducers); it reads 32 of them at a time and
        // lane 0 issues the two bulk copies of each (slot k % stages, in use for the (k / stages)-th time) ----
        const int pw = warp - NB;
        int st = pw % stages;
        unsigned round = pw / stages;
        for (int base = pw; base < count; base += 32 * kProducers) {
            const int idx = base + lane * kProducers;
            const int mine = idx < count ? __ldg(list + idx) : 0;
            const int lim = min(32, (count - base + kProducers - 1) / kProducers);
            for (int i = 0; i < lim; ++i) {
                const int n = __shfl_sync(0xffffffffu, mine, i);
                if (lane == 0) {
                    if (round > 0) mbar_wait(empty + st * 8, (round - 1) & 1);
                    const unsigned dst = ring + st * kStageBytes;
                    mbar_expect_tx(full + st * 8, kStageBytes);
                    bulk_load(dst, grad_out + ((size_t)n * C + (size_t)ct * kK) * 49, kTileBytes, full + st * 8);
                    bulk_load(dst + kTileBytes, tab_space + (size_t)n * kRoiTabSlotBytes, kTabBytes, full + st * 8);
                }
                st += kProducers;
                while (st >= stages) {
                    st -= stages;
                    ++round;
                }
                __syncwarp();
            }
        }
    } else {
        // ---- consumers: warp = band of feature rows [r0, r1), lane = channel ----
        const int* bnd = bounds + (size_t)(b * nslabs + sl) * (kBandMaxWarps + 1);
        const int r0 = __ldg(bnd + warp), r1 = __ldg(bnd + warp + 1);
        // a lattice row works for this band when its upper row lies in [r0 - 1, r1 - 1]
        const unsigned ylo = (unsigned)max(r0 - 1, 0), yhi = (unsigned)(r1 - 1);
        const bool idle = r1 <= r0;
        const unsigned lane_planes = smem_u32(planes) + lane * 4 - (unsigned)(s0 * RB);   // + row * RB + column * 128
        int s = 0;
        unsigned round = 0;
        for (int k = 0; k < count; ++k) {
            mbar_wait(full + s * 8, round & 1);
            const unsigned stage = ring + s * kStageBytes;
            const unsigned t = stage + kTileBytes;
            // lanes 0-7 test one lattice row each
            const unsigned ysl = lds_u8(t + 144 + (lane & 7));
            unsigned bits = __ballot_sync(0xffffffffu, ysl >= ylo && ysl <= yhi) & 0xffu;
            if (idle || debug == 1) bits = 0;
            if (bits) {
                const uint4 hdr = lds_u4(t + 144);                  // {ys[0..3], ys[4..7], xmode, rows}
                const uint4 xo = lds_u4(t + 128);
                unsigned long long w[8];
                lds_pair2(t + 64, w[0], w[1]);
                lds_pair2(t + 80, w[2], w[3]);
                lds_pair2(t + 96, w[4], w[5]);
                lds_pair2(t + 112, w[6], w[7]);
                const unsigned xa[8] = {lane_planes + (xo.x & 0xffffu), lane_planes + (xo.x >> 16),
                                        lane_planes + (xo.y & 0xffffu), lane_planes + (xo.y >> 16),
                                        lane_planes + (xo.z & 0xffffu), lane_planes + (xo.z >> 16),
                                        lane_planes + (xo.w & 0xffffu), lane_planes + (xo.w >> 16)};
                const bool narrow = hdr.z != 0;
                const unsigned tile = stage + lane * (49 * 4);
                float ga[P];                       // pooled row under the previous lattice row handled (rolling)
                int have = -2;
                do {
                    const int ph = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const float2 wy = lds_f2(t + ph * 8);
                    const unsigned row = ((ph < 4 ? hdr.x : hdr.y) >> ((ph & 3) * 8)) & 0xffu;
                    float d[8];
                    if (POOL == I2V_POOL_NONE) {
                        // lattice == pooled grid (ph < 7: row 7 does not exist and is never listed)
                        const unsigned tr = tile + ph * (P * 4);
                        d[0] = lds<0>(tr); d[1] = lds<4>(tr); d[2] = lds<8>(tr); d[3] = lds<12>(tr);
                        d[4] = lds<16>(tr); d[5] = lds<20>(tr); d[6] = lds<24>(tr);
                        d[7] = 0.f;
                    } else {
                        // lattice row ph collects the pooled rows ph-1 and ph, lattice column j the pooled columns j-1 and j
                        float gb[P];
                        if (ph < P) {
                            const unsigned tr = tile + ph * (P * 4);
                            gb[0] = lds<0>(tr); gb[1] = lds<4>(tr); gb[2] = lds<8>(tr); gb[3] = lds<12>(tr);
                            gb[4] = lds<16>(tr); gb[5] = lds<20>(tr); gb[6] = lds<24>(tr);
                        } else {
#pragma unroll
                            for (int j = 0; j < P; ++j) gb[j] = 0.f;
                        }
                        if (have != ph - 1) {
                            if (ph >= 1) {
                                const unsigned tr = tile + (ph - 1) * (P * 4);
                                ga[0] = lds<0>(tr); ga[1] = lds<4>(tr); ga[2] = lds<8>(tr); ga[3] = lds<12>(tr);
                                ga[4] = lds<16>(tr); ga[5] = lds<20>(tr); ga[6] = lds<24>(tr);
                            } else {
#pragma unroll
                                for (int j = 0; j < P; ++j) ga[j] = 0.f;
                            }
                        }
                        float sj[P];
#pragma unroll
                        for (int j = 0; j < P; ++j) {
                            sj[j] = ga[j] + gb[j];
                            ga[j] = gb[j];
                        }
                        have = ph;
                        d[0] = sj[0];
#pragma unroll
                        for (int j = 1; j < P; ++j) d[j] = sj[j - 1] + sj[j];
                        d[7] = sj[P - 1];
                    }
                    const unsigned ro = row * RB;
                    const unsigned a[8] = {xa[0] + ro, xa[1] + ro, xa[2] + ro, xa[3] + ro,
                                           xa[4] + ro, xa[5] + ro, xa[6] + ro, xa[7] + ro};
                    const bool do_u = (int)row >= r0, do_l = row < yhi;
                    if (WPT) {
                        if (do_u && do_l) lattice_row<true, true, WPT * 128>(a, w, d, wy.x, wy.y, narrow);
                        else if (do_u) lattice_row<true, false, WPT * 128>(a, w, d, wy.x, wy.y, narrow);
                        else lattice_row<false, true, WPT * 128>(a, w, d, wy.x, wy.y, narrow);
                    } else {
                        if (do_u) lattice_row<true, false, 0>(a, w, d, wy.x, wy.y, narrow);
                        if (do_l) {
                            const unsigned al[8] = {a[0] + RB, a[1] + RB, a[2] + RB, a[3] + RB,
                                                    a[4] + RB, a[5] + RB, a[6] + RB, a[7] + RB};
                            lattice_row<true, false, 0>(al, w, d, wy.y, wy.y, narrow);
                        }
                    }
                } while (bits);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s * 8);
            if (++s == stages) {
                s = 0;
                ++round;
            }
        }
    }

    // ---- write-out: [row][col][32] in shared memory -> [32][row][col] in HBM.  A quarter-warp reads the 32 channels of
    // one cell (128 contiguous bytes); a lane stores its four channels to four planes, four lanes cover 16 bytes of a row ----
    __syncthreads();
    {
        const int q = lane & 7, sub = lane >> 3;
        const size_t HW = (size_t)H * W;
        float* dst = grad_in + ((size_t)b * C + (size_t)ct * kK + (size_t)q * 4) * HW + (size_t)s0 * W;
        const float4* src = reinterpret_cast<const float4*>(planes);
        for (int r = warp; r < nrows; r += NB + kProducers) {
            for (int x = sub; x < W; x += 4) {
                const float4 v = src[(r * WP + x) * (kK / 4) + q];
                float* p = dst + (size_t)r * W + x;
                p[0] = v.x;
                p[HW] = v.y;
                p[2 * HW] = v.z;
                p[3 * HW] = v.w;
            }
        }
    }
}

}  // namespace

// slab geometry for a map: the fewest slabs whose 32 planes fit next to a ring of at least kMinStages tiles
static bool band_geometry(int H, int W, int& nslabs, int& slab_rows, int& stages) {
    for (int ns = 1; ns <= kBandMaxSlabs; ++ns) {
        const int rows = ceil_div(H, ns);
        const size_t planes = (size_t)rows * (W + 2) * kK * sizeof(float);
        if (planes + kBarBytes + (size_t)kMinStages * kStageBytes > (size_t)kMaxSmemPerCta) continue;
        nslabs = ceil_div(H, rows);
        slab_rows = rows;
        stages = (int)min((size_t)kMaxStages, ((size_t)kMaxSmemPerCta - planes - kBarBytes) / kStageBytes);
        return true;
    }
    return false;
}

bool bwd_band_ok(const float* grad_out, int batch, int C, int H, int W, int PH, int PW, int pool_mode) {
    int ns, rows, st;
    return batch > 0 && PH == 7 && PW == 7 && pool_mode != I2V_POOL_MAX && C % kK == 0 && H >= 2 && W >= 2 &&
           W * 128 <= 65535 && H <= 250 && ((uintptr_t)grad_out & 15) == 0 && band_geometry(H, W, ns, rows, st);
}

size_t bwd_band_list_ints(int batch, int num_rois) {
    return (size_t)kBandMaxSlabs * num_rois + (size_t)kBandMaxSlabs * (batch + 1) * (kBandMaxWarps + 2);
}

template <int POOL, int NB, int WPT>
static int launch_band(const float* grad_out, const unsigned char* tab_space, const int* sorder, const int* scount,
                       const int* starts, const int* bounds, float* grad_in, int batch, int C, int H, int W, int nslabs,
                       int slab_rows, int stages, int num_rois, cudaStream_t stream) {
    auto kern = lattice_bwd_band_kernel<POOL, NB, WPT>;
    const size_t smem = (size_t)stages * kStageBytes + kBarBytes + (size_t)slab_rows * (W + 2) * kK * sizeof(float);
    I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3((unsigned)(batch * (C / kK) * nslabs)), (NB + kProducers) * 32, smem, stream>>>(
        grad_out, tab_space, sorder, scount, starts, bounds, grad_in, C, H, W, nslabs, slab_rows, stages, num_rois,
        getenv("I2V_BAND_DEBUG") ? atoi(getenv("I2V_BAND_DEBUG")) : 0);
    return check_launch("lattice_bwd_band_kernel");
}
template <int POOL, int NB>
static int launch_band_w(const float* grad_out, const unsigned char* tab_space, const int* sorder, const int* scount,
                         const int* starts, const int* bounds, float* grad_in, int batch, int C, int H, int W, int nslabs,
                         int slab_rows, int stages, int num_rois, cudaStream_t stream) {
    if (W == 63)
        return launch_band<POOL, NB, 65>(grad_out, tab_space, sorder, scount, starts, bounds, grad_in, batch, C, H, W, nslabs,
                                         slab_rows, stages, num_rois, stream);
    return launch_band<POOL, NB, 0>(grad_out, tab_space, sorder, scount, starts, bounds, grad_in, batch, C, H, W, nslabs,
                                    slab_rows, stages, num_rois, stream);
}

constexpr int kBandDefaultWarps = 12;

// `tab` holds the LatticeRoi tables of this call; `tab_space` is the workspace's per-RoI table slot (kRoiTabSlotBytes
// each); `lists` has room for bwd_band_list_ints(batch, num_rois) ints.  `bands` (4, 8, 12 or 16; 0 = default) is the
// number of row bands = consumer warps per CTA.
int launch_bwd_band(const float* grad_out, const LatticeRoi* tab, void* tab_space, const int* order, const int* starts,
                    int* lists, float* grad_in, int batch, int C, int H, int W, int num_rois, int pool_mode, int bands,
                    cudaStream_t stream) {
    int nslabs, slab_rows, stages;
    if (!band_geometry(H, W, nslabs, slab_rows, stages)) {
        set_error("roi_align_backward: no slab geometry for a %dx%d map", H, W);
        return I2V_ERR_UNSUPPORTED;
    }
    if (bands != 4 && bands != 8 && bands != 12 && bands != 16) bands = kBandDefaultWarps;
    unsigned char* ts = static_cast<unsigned char*>(tab_space);
    int* sorder = lists;
    int* scount = lists + (size_t)kBandMaxSlabs * num_rois;
    int* bounds = scount + (size_t)kBandMaxSlabs * (batch + 1);
    const int G = pool_mode == I2V_POOL_NONE ? 7 : 8;
    band_prep_kernel<<<ceil_div(num_rois, 128), 128, 0, stream>>>(tab, ts, num_rois, G, W,
                                                                   pool_mode == I2V_POOL_AVG ? 0.25f : 1.f);
    I2V_TRY(check_launch("band_prep_kernel"));
    band_bucket_kernel<<<batch * nslabs, 256, 0, stream>>>(ts, order, starts, sorder, scount, bounds, num_rois, nslabs,
                                                           slab_rows, H, bands);
    I2V_TRY(check_launch("band_bucket_kernel"));
#define I2V_BAND_LAUNCH(POOL, NB)                                                                                     \
    return launch_band_w<POOL, NB>(grad_out, ts, sorder, scount, starts, bounds, grad_in, batch, C, H, W, nslabs, slab_rows, \
                                   stages, num_rois, stream)
    if (pool_mode == I2V_POOL_AVG) {
        if (bands == 4) I2V_BAND_LAUNCH(I2V_POOL_AVG, 4);
        if (bands == 12) I2V_BAND_LAUNCH(I2V_POOL_AVG, 12);
        if (bands == 16) I2V_BAND_LAUNCH(I2V_POOL_AVG, 16);
        I2V_BAND_LAUNCH(I2V_POOL_AVG, 8);
    }
    if (bands == 4) I2V_BAND_LAUNCH(I2V_POOL_NONE, 4);
    if (bands == 12) I2V_BAND_LAUNCH(I2V_POOL_NONE, 12);
    if (bands == 16) I2V_BAND_LAUNCH(I2V_POOL_NONE, 16);
    I2V_BAND_LAUNCH(I2V_POOL_NONE, 8);
#undef I2V_BAND_LAUNCH
}

}  // namespace i2v
