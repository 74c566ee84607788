// RoIAlign / RoIAlignAvg forward, lane-per-channel variant (roi_align_kernel.cu:15-70 + the module's pool,
// modules/roi_align.py:18-29).
//
// The slab kernel (roi_align_fwd_slab.cu) is bound by shared-memory wavefronts: its lanes are (RoI slot) x (16 channels), so
// the two half-warps of a load hit unrelated cells and collide half of the time (161 M wavefronts for 104 M loads).  Here
// a warp works on ONE RoI and its 32 lanes are 32 channels: the planes sit in shared memory cell-major,
// [row][col][32 channels] with the channel index XOR-swizzled by the column (so that the fill, whose lanes are columns, is
// conflict free as well), and every load of a lattice point is one 128-byte wavefront.
//
// 32 channels of a 38 x 63 map are 306 KB, so a CTA holds a SLAB of rows: rows [0, S0] or [S0, H) with S0 = ceil(H / 2) (one
// row of overlap: a lattice row needs its start row and the next one).  A lattice row belongs to the slab that holds its
// start row; a RoI whose lattice rows fall on both sides is worked on by both CTAs, each producing the pooled rows whose two
// lattice rows it owns.  The one pooled row in between gets a partial sum from either side: the prep kernel zeroes that
// row of the output and both CTAs add their half with red.global.add.f32 -- x + y = y + x bit for bit and 0 + x = x, so
// the result does not depend on who comes first.  No flags, no ordering between CTAs.
#include "common.cuh"
#include <stdio.h>

namespace i2v {

struct alignas(16) ChanTab {
    int xs[8];             // start column of lattice column pw (0 when it is off the map)
    float w0[8], w1[8];    // weights of the left / right cell (validity and the avg pool's 1/4 folded in)
    int ys[8];             // start row of lattice row ph, clamped to [0, H - 2]
    float wy0[8], wy1[8];  // weights of the upper / lower row (validity folded in)
    int split;             // lattice rows [0, split) belong to slab 0, [split, G) to slab 1
    int pad[3];
};
static_assert(sizeof(ChanTab) == 208 && sizeof(ChanTab) <= kRoiTabSlotBytes, "ChanTab layout");

namespace {

constexpr int kCh = 32;
constexpr int kWarps = 10;
constexpr int kThreads = kWarps * 32;
constexpr int kTabV = (int)(sizeof(ChanTab) / 16);      // 13 16-byte pieces
constexpr int kNOut = 49;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store_commit(float* gdst, const float* ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void red_add(float* p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

__host__ __device__ inline int chan_s0(int H) { return (H + 1) / 2; }

// One thread per RoI writes its table; then the block zeroes, for every RoI of its range whose lattice rows fall on both
// slabs, the pooled row that both slabs add into (C x 7 floats per RoI, 7 contiguous per channel).
__global__ void __launch_bounds__(256) chan_prep_kernel(const LatticeRoi* __restrict__ tab, ChanTab* __restrict__ ctab,
                                                        float* __restrict__ out, int num_rois, int G, int H, int C, float wscale,
                                                        int pooled) {
    __shared__ int s_split[256];
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    int split = 0;
    if (n < num_rois) {
        const LatticeRoi& t = tab[n];
        ChanTab q;
        const int S0 = chan_s0(H);
        bool seen_valid = false;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const bool okx = p < G && ((t.valid_x >> p) & 1u), oky = p < G && ((t.valid_y >> p) & 1u);
            q.xs[p] = okx ? t.x.start[p] : 0;
            q.w0[p] = okx ? (1.f - t.x.frac[p]) * wscale : 0.f;
            q.w1[p] = okx ? t.x.frac[p] * wscale : 0.f;
            // a row off the map has weight 0 and may read anywhere; it is placed above / below the valid rows (the sample
            // positions grow with p), so that the rows of slab 0 are a prefix
            seen_valid |= oky;
            const int yc = oky ? min(max(t.y.start[p], 0), H - 2) : (seen_valid || p >= G ? H - 2 : 0);
            q.ys[p] = yc;
            q.wy0[p] = oky ? 1.f - t.y.frac[p] : 0.f;
            q.wy1[p] = oky ? t.y.frac[p] : 0.f;
            if (p < G && yc < S0) ++split;
        }
        q.split = t.batch < 0 ? G : split;      // stray RoIs are handled (zero rows) by the extra bucket of slab 0
        q.pad[0] = q.pad[1] = q.pad[2] = 0;
        split = q.split;
        ctab[n] = q;
        if (t.batch < 0) split = 0;             // nothing to zero
    }
    s_split[threadIdx.x] = (n < num_rois) ? split : 0;
    __syncthreads();
    if (!pooled) return;                        // without the pool every output row has one owner
    for (int r = 0; r < 256; ++r) {
        const int sp = s_split[r];
        if (sp <= 0 || sp >= G) continue;       // uniform
        const int nn = blockIdx.x * 256 + r;
        float* row = out + (size_t)nn * C * kNOut + (size_t)(sp - 1) * 7;
        for (int i = threadIdx.x; i < C * 7; i += 256) {
            const int c = i / 7, j = i - c * 7;
            row[(size_t)c * kNOut + j] = 0.f;
        }
    }
}

template <int POOL, int WT>
__global__ void __launch_bounds__(kThreads, 1)
    lattice_fwd_chan_kernel(const float* __restrict__ feat, const ChanTab* __restrict__ ctab, const int* __restrict__ order,
                            const int* __restrict__ starts, float* __restrict__ out, int batch, int C, int H, int Wrt) {
    constexpr int P = 7;
    constexpr int G = (POOL == I2V_POOL_NONE) ? P : P + 1;
    constexpr int TILE = kCh * kNOut;                      // floats per staged output tile
    extern __shared__ __align__(128) float smem[];
    const int W = WT ? WT : Wrt;
    const int S0 = chan_s0(H);
    const int max_rows = max(S0 + 1, H - S0);
    float* planes = smem;                                  // [rows][W][32]
    float* stage = planes + (size_t)max_rows * W * kCh;    // [warps][TILE]
    ChanTab* tabs = reinterpret_cast<ChanTab*>(stage + (size_t)kWarps * TILE);   // [warps][2 buffers]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ctiles = C / kCh;
    const int slab = blockIdx.x & 1;
    const int ct = (blockIdx.x >> 1) % ctiles;
    const int b = (blockIdx.x >> 1) / ctiles;
    const int list_lo = __ldg(starts + b), list_hi = __ldg(starts + b + 1);
    if (list_lo == list_hi) return;

    if (b == batch) {  // RoIs with an out-of-range batch index: zero rows (slab 0's CTAs only)
        if (slab) return;
        for (int li = list_lo + warp; li < list_hi; li += kWarps) {
            float4* dst = reinterpret_cast<float4*>(out + ((size_t)__ldg(order + li) * C + (size_t)ct * kCh) * kNOut);
            for (int i = lane; i < TILE / 4; i += 32) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }

    const int r0 = slab ? S0 : 0;
    const int nrows = slab ? H - S0 : S0 + 1;

    // ---- fill: lanes are columns (coalesced global rows), the channel index is swizzled by the column so that the 32
    // stores of a row segment fall on 32 banks; four (channel, row) pairs = eight loads are in flight per lane ----
    {
        const float* src = feat + ((size_t)b * C + (size_t)ct * kCh) * H * W + (size_t)r0 * W;
        const int pairs = nrows * kCh;
        for (int p0 = warp * 4; p0 < pairs; p0 += kWarps * 4) {
            float v[4][2];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int p = p0 + q, c = p & 31, y = p >> 5;
                const float* g = src + ((size_t)c * H + y) * W;
                v[q][0] = (p < pairs && lane < W) ? __ldg(g + lane) : 0.f;
                v[q][1] = (p < pairs && lane + 32 < W) ? __ldg(g + lane + 32) : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int p = p0 + q, c = p & 31, y = p >> 5;
                if (p < pairs) {
                    float* d = planes + (size_t)y * W * kCh;
                    if (lane < W) d[lane * kCh + (c ^ lane)] = v[q][0];
                    if (lane + 32 < W) d[(lane + 32) * kCh + (c ^ lane)] = v[q][1];   // (lane + 32) & 31 == lane
                }
            }
        }
    }

    float* my_stage = stage + (size_t)warp * TILE;
    ChanTab* my_tabs = tabs + warp * 2;

    int li = list_lo + warp;
    int n_cur = 0, n_next = 0;
    auto fetch_table = [&](int n, int bufi) {              // 13 lanes copy the 13 x 16 bytes of one table
        if (lane < kTabV)
            cp_async16(reinterpret_cast<char*>(my_tabs + bufi) + lane * 16, reinterpret_cast<const char*>(ctab + n) + lane * 16);
    };
    if (li < list_hi) {
        n_cur = __ldg(order + li);
        fetch_table(n_cur, 0);
        if (li + kWarps < list_hi) n_next = __ldg(order + li + kWarps);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    __syncthreads();                                       // the planes are in place

    int it = 0;
    for (; li < list_hi; li += kWarps, ++it) {             // uniform per warp
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        const ChanTab* t = my_tabs + (it & 1);
        int n_next2 = 0;
        if (li + kWarps < list_hi) {                       // the next RoI's table, and the index after that
            fetch_table(n_next, (it + 1) & 1);
            if (li + 2 * kWarps < list_hi) n_next2 = __ldg(order + li + 2 * kWarps);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        const int n = n_cur;
        n_cur = n_next;
        n_next = n_next2;

        const int split = t->split;
        const int lo = slab ? split : 0, hi = slab ? G : split;      // this slab's lattice rows
#ifdef I2V_CHAN_CHECK
        {
            bool bad = split < 0 || split > G || n < 0;
            for (int p = 0; p < G; ++p) {
                bad |= t->xs[p] < 0 || t->xs[p] > W - 2;
                if (p >= lo && p < hi) bad |= (t->ys[p] - r0) < 0 || (t->ys[p] - r0) > nrows - 2;
            }
            if (bad) {
                if (lane == 0)
                    printf("chan: bad table n=%d slab=%d split=%d lo=%d hi=%d ys=%d..%d xs=%d..%d r0=%d nrows=%d li=%d it=%d\n", n, slab,
                           split, lo, hi, t->ys[0], t->ys[G - 1], t->xs[0], t->xs[G - 1], r0, nrows, li, it);
                continue;
            }
        }
#endif
        if (lo >= hi) continue;

        // word offsets of the lane's channel in the left / right cell of every lattice column, and the column weights
        int cl[8], cr[8];
        float w0[8], w1[8];
#pragma unroll
        for (int pw = 0; pw < G; ++pw) {
            const int x = t->xs[pw];
            cl[pw] = x * kCh + (lane ^ (x & 31));
            cr[pw] = (x + 1) * kCh + (lane ^ ((x + 1) & 31));
            w0[pw] = t->w0[pw];
            w1[pw] = t->w1[pw];
        }
        float part[kNOut];
        float prev[P], bnd[P];
#pragma unroll
        for (int j = 0; j < P; ++j) bnd[j] = 0.f;
#pragma unroll
        for (int ph = 0; ph < G; ++ph) {
            if (ph >= lo && ph < hi) {                     // uniform
                const float* row = planes + (t->ys[ph] - r0) * W * kCh;
                const int down = W * kCh;                  // an immediate when W is a template constant
                const float wy0 = t->wy0[ph], wy1 = t->wy1[ph];
                float curv[G];
#pragma unroll
                for (int pw = 0; pw < G; ++pw) {
                    const float* a = row + cl[pw];
                    const float* c2 = row + cr[pw];
                    curv[pw] = (a[0] * wy0 + a[down] * wy1) * w0[pw] + (c2[0] * wy0 + c2[down] * wy1) * w1[pw];
                }
                if (POOL == I2V_POOL_NONE) {
#pragma unroll
                    for (int pw = 0; pw < G; ++pw) part[ph * P + pw] = curv[pw];
                } else {
#pragma unroll
                    for (int pw = 0; pw < P; ++pw) curv[pw] += curv[pw + 1];   // adjacent columns (x 1/4 in the weights)
                    if (ph > lo) {
#pragma unroll
                        for (int pw = 0; pw < P; ++pw) part[(ph > 0 ? ph - 1 : 0) * P + pw] = prev[pw] + curv[pw];
                    }
                    // the half of a pooled row whose other lattice row belongs to the other slab
                    if ((ph == lo && lo > 0) || (ph == hi - 1 && hi < G)) {
#pragma unroll
                        for (int pw = 0; pw < P; ++pw) bnd[pw] = curv[pw];
                    }
#pragma unroll
                    for (int pw = 0; pw < P; ++pw) prev[pw] = curv[pw];
                }
            }
        }

        // ---- output: rows [ra, rb) are complete here ----
        const int ra = lo, rb = (POOL == I2V_POOL_NONE) ? hi : hi - 1;
        float* dst = out + ((size_t)n * C + (size_t)ct * kCh) * kNOut;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        float* srow = my_stage + lane * kNOut;             // 49 = 17 (mod 32): the lanes' rows start on distinct banks
        if (ra == 0 && rb == P) {                          // the whole tile: contiguous in NCHW, one bulk store
#pragma unroll
            for (int k = 0; k < kNOut; ++k) srow[k] = part[k];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) bulk_store_commit(dst, my_stage, TILE * sizeof(float));
            continue;
        }
#pragma unroll
        for (int r = 0; r < P; ++r) {
            if (r >= ra && r < rb) {
#pragma unroll
                for (int j = 0; j < P; ++j) srow[r * P + j] = part[r * P + j];
            }
        }
        __syncwarp();
        // complete rows: 7 (rb - ra) contiguous floats per channel
        const int len = P * (rb - ra);
        if (len > 0) {
            for (int c = 0; c < kCh; ++c) {
                const float* s = my_stage + c * kNOut + ra * P;
                float* d = dst + (size_t)c * kNOut + ra * P;
                if (lane < len) d[lane] = s[lane];
                if (lane + 32 < len) d[lane + 32] = s[lane + 32];
            }
        }
        if (POOL != I2V_POOL_NONE) {
            // the shared row: 7 floats per channel, added to what the other slab adds (the prep kernel zeroed the row)
            const int rs = slab ? lo - 1 : hi - 1;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < P; ++j) srow[j] = bnd[j];   // the first 7 words of the lane's staging row
            __syncwarp();
            for (int i = lane; i < kCh * P; i += 32) {
                const int c = i / P, j = i - c * P;
                red_add(dst + (size_t)c * kNOut + rs * P + j, my_stage[c * kNOut + j]);
            }
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

size_t chan_smem_bytes(int H, int W) {
    const int S0 = chan_s0(H);
    const int max_rows = S0 + 1 > H - S0 ? S0 + 1 : H - S0;
    return ((size_t)max_rows * W * kCh + (size_t)kWarps * kCh * kNOut) * sizeof(float) + (size_t)kWarps * 2 * sizeof(ChanTab);
}

}  // namespace

bool fwd_chan_ok(const float* features, const float* out, int batch, int C, int H, int W, int PH, int PW, int pool_mode) {
    return batch > 0 && PH == 7 && PW == 7 && pool_mode != I2V_POOL_MAX && C % kCh == 0 && H >= 4 && W >= 2 && W <= 64 &&
           chan_smem_bytes(H, W) <= (size_t)kMaxSmemPerCta && ((uintptr_t)out & 15) == 0 && features != nullptr;
}

template <int POOL, int WT>
static int launch_chan(const float* feat, const ChanTab* ctab, const int* order, const int* starts, float* out, int batch,
                       int C, int H, int W, cudaStream_t stream) {
    auto kern = lattice_fwd_chan_kernel<POOL, WT>;
    const size_t smem = chan_smem_bytes(H, W);
    I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(batch + 1) * (C / kCh) * 2, kThreads, smem, stream>>>(feat, ctab, order, starts, out, batch, C, H, W);
    return check_launch("lattice_fwd_chan_kernel");
}

// `tab` holds the LatticeRoi tables of this call; `tab_space` is the workspace's per-RoI table slot (kRoiTabSlotBytes each).
int launch_fwd_chan(const float* feat, const LatticeRoi* tab, void* tab_space, const int* order, const int* starts, float* out,
                    int batch, int C, int H, int W, int num_rois, int pool_mode, cudaStream_t stream) {
    ChanTab* ctab = static_cast<ChanTab*>(tab_space);
    const int G = pool_mode == I2V_POOL_NONE ? 7 : 8;
    chan_prep_kernel<<<ceil_div(num_rois, 256), 256, 0, stream>>>(tab, ctab, out, num_rois, G, H, C,
                                                                   pool_mode == I2V_POOL_AVG ? 0.25f : 1.f,
                                                                   pool_mode != I2V_POOL_NONE);
    I2V_TRY(check_launch("chan_prep_kernel"));
    if (pool_mode == I2V_POOL_AVG) {
        if (W == 63) return launch_chan<I2V_POOL_AVG, 63>(feat, ctab, order, starts, out, batch, C, H, W, stream);
        return launch_chan<I2V_POOL_AVG, 0>(feat, ctab, order, starts, out, batch, C, H, W, stream);
    }
    return launch_chan<I2V_POOL_NONE, 0>(feat, ctab, order, starts, out, batch, C, H, W, stream);
}

}  // namespace i2v
