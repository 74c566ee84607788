// Lattice RoIAlign (the cffi-era op of lib/model/roi_align) for sm_100a: forward and backward, with the
// 2x2/stride-1 pool of RoIAlignAvg / RoIAlignMax fused.
//
// Reference semantics: lib/model/roi_align/src/roi_align_kernel.cu:15-70 (forward), :94-143 (backward),
// lib/model/roi_align/modules/roi_align.py:18-42 (the pool behind it).
//
// Three kernels per direction:
//   prep    one thread per RoI: the per-axis lattice tables (start cell, fraction, validity), computed with
//           exactly the reference's roundings, plus the per-frame RoI lists the plane kernels walk;
//   plane   one CTA per (frame, 16 channels): the 16 feature planes live in shared memory as [cell][16],
//           every feature byte is read from HBM once, lanes are channels so every shared-memory access is
//           conflict free and all index math is warp-uniform; pooled tiles leave through TMA bulk stores;
//   gather  one thread per output element straight from global memory (any shape; fp64 weights like the
//           reference, so its forward is bit-identical to roi_align.c).
#include <stdlib.h>

#include "common.cuh"

namespace i2v {

// ------------------------------------------------------------------------------------------ prep
// roi_align_kernel.cu:27-57, evaluated once per RoI instead of once per output element.  The products that
// the reference promotes to double (`+ 1.`, `/ (aligned - 1.)`) are promoted here too; `ph * bin + start` is
// kept as a separate multiply and add (the CPU twin roi_align.c:106-107 has no FMA and is what the oracle pins).
__device__ __forceinline__ void lattice_axis(float lo, float hi, int G, int extent, LatticeAxis& ax, unsigned& valid,
                                             bool& strictly_increasing) {
    float span = fmaxf((float)((double)__fsub_rn(hi, lo) + 1.), 0.f);
    float bin = (float)((double)span / ((double)G - 1.));
    unsigned v = 0;
    bool inc = true;
    int prev = -1;
#pragma unroll
    for (int p = 0; p < kMaxLattice; ++p) {
        if (p < G) {
            float pos = __fadd_rn(__fmul_rn((float)p, bin), lo);
            int st = (int)fminf(floorf(pos), (float)(extent - 2));
            bool ok = !(pos < 0.f || pos >= (float)extent);
            if (ok) {
                v |= 1u << p;
                if (st <= prev) inc = false;
                prev = st;
            } else {
                st = 0;
            }
            ax.start[p] = st;
            ax.frac[p] = ok ? __fsub_rn(pos, (float)st) : 0.f;
        } else {
            ax.start[p] = 0;
            ax.frac[p] = 0.f;
        }
    }
    valid = v;
    strictly_increasing = inc;
}

__global__ void lattice_prep_kernel(const float* __restrict__ rois, int num_rois, int batch, int H, int W, int GH,
                                    int GW, float scale, LatticeRoi* __restrict__ tab, int* __restrict__ roi_batch,
                                    int* __restrict__ counts, PlaneTab* __restrict__ ptab, float ptab_scale,
                                    BwdTab* __restrict__ btab, int ptab_row_bytes) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= num_rois) return;
    const float* r = rois + (size_t)n * 5;
    LatticeRoi t;
    int b = (int)r[0];
    bool in_batch = (b >= 0 && b < batch);
    bool incx, incy;
    lattice_axis(__fmul_rn(r[1], scale), __fmul_rn(r[3], scale), GW, W, t.x, t.valid_x, incx);
    lattice_axis(__fmul_rn(r[2], scale), __fmul_rn(r[4], scale), GH, H, t.y, t.valid_y, incy);
    t.batch = in_batch ? b : -1;
    t.flags = (incy ? 1u : 0u) | (incx ? 2u : 0u);
    {
        unsigned runpos = 0, maxrun = 0, same = 0;
        int prev = -1, run = 0, prevx = -1;
#pragma unroll
        for (int p = 0; p < 8; ++p) {   // the plane kernels handle lattices of up to 8 points per axis
            if ((t.valid_y >> p) & 1u) {
                run = (t.y.start[p] == prev) ? run + 1 : 0;
                prev = t.y.start[p];
                runpos |= (unsigned)run << (4 * p);
                maxrun = max(maxrun, (unsigned)run + 1u);
            }
            if ((t.valid_x >> p) & 1u) {
                if (t.x.start[p] == prevx) same |= 1u << p;
                prevx = t.x.start[p];
            }
        }
        t.y_runpos = runpos;
        t.y_maxrun = maxrun;
        t.x_same = same;
        t.pad_ = 0;
    }
    tab[n] = t;
    if (ptab) {  // the forward plane kernel's view: byte offsets into the [row][64][16] planes, ready-to-use weights
        PlaneTab q;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            bool okx = (t.valid_x >> p) & 1u, oky = (t.valid_y >> p) & 1u;
            int sx = t.x.start[p];
            float wl = okx ? (1.f - t.x.frac[p]) * ptab_scale : 0.f, wr = okx ? t.x.frac[p] * ptab_scale : 0.f;
            int even = (sx & 1) ? sx + 1 : sx, odd = (sx & 1) ? sx : sx + 1;
            float we = (sx & 1) ? wr : wl, wo = (sx & 1) ? wl : wr;
            q.xa[p] = make_float4(__int_as_float(even * 64), we, __int_as_float(odd * 64), wo);
            q.xb[p] = make_float4(__int_as_float(odd * 64), wo, __int_as_float(even * 64), we);
            q.y[p] = make_float4(__int_as_float(t.y.start[p] * ptab_row_bytes), oky ? 1.f - t.y.frac[p] : 0.f,
                                 oky ? t.y.frac[p] : 0.f, 0.f);
        }
        ptab[n] = q;
    }
    if (btab) {  // the backward plane kernel's view
        BwdTab q;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            bool okx = (t.valid_x >> p) & 1u;
            q.xoff[p] = t.x.start[p] * 4;
            q.wx0[p] = okx ? (1.f - t.x.frac[p]) * ptab_scale : 0.f;
            q.wx1[p] = okx ? t.x.frac[p] * ptab_scale : 0.f;
            q.yoff[p] = t.y.start[p] * W * 4;
            q.wy0[p] = 1.f - t.y.frac[p];
            q.wy1[p] = t.y.frac[p];
        }
        q.valid_y = t.valid_y;
        q.valid_x = t.valid_x;
        q.y_runpos = t.y_runpos;
        q.y_maxrun = t.y_maxrun;
        q.x_same = t.x_same;
        q.pad_[0] = q.pad_[1] = q.pad_[2] = 0;
        btab[n] = q;
    }
    // RoIs whose batch index is out of range are listed in the extra bucket `batch` (their rows are zero-filled)
    if (roi_batch) roi_batch[n] = in_batch ? b : batch;
    if (counts) atomicAdd(&counts[in_batch ? b : batch], 1);
}

// One CTA per bucket (frames 0..batch-1, then the bucket of out-of-range RoIs): writes the indices of that
// bucket's RoIs, in ascending order, at order[sum(counts[0..b)) ...]; `batch` here counts the extra bucket.  Ordered block compaction (ballot + warp/block prefix), so the lists -- and
// with them the accumulation order of the plane backward -- are deterministic.
__global__ void __launch_bounds__(256) roi_bucket_kernel(const int* __restrict__ roi_batch, int num_rois, int batch,
                                                         const int* __restrict__ counts, int* __restrict__ starts,
                                                         int* __restrict__ order) {
    __shared__ int s_warp[8];
    __shared__ int s_base;
    int b = blockIdx.x;
    int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // start of this frame's list = sum of the earlier counts
    int part = 0;
    for (int i = tid; i < b; i += 256) part += counts[i];
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_warp[warp] = part;
    __syncthreads();
    if (tid == 0) {
        int s = 0;
        for (int w = 0; w < 8; ++w) s += s_warp[w];
        s_base = s;
        starts[b] = s;
        if (b == batch - 1) starts[batch] = s + counts[b];
    }
    __syncthreads();
    // Each thread owns a contiguous chunk of the RoI indices, counts its matches, takes its offset from one block-wide
    // exclusive scan and writes them in ascending order: two barriers in all, instead of two per 256 RoIs.
    const int chunk = (num_rois + 255) / 256;
    const int lo = min(tid * chunk, num_rois), hi = min(lo + chunk, num_rois);
    int mine = 0;
    for (int i = lo; i < hi; ++i) mine += (__ldg(roi_batch + i) == b);
    int incl = mine;                                    // inclusive scan inside the warp
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    __syncthreads();                                    // s_warp reuse
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    int pos = s_base + before + incl - mine;
    for (int i = lo; i < hi; ++i)
        if (__ldg(roi_batch + i) == b) order[pos++] = i;
}

// ------------------------------------------------------------------------------------------ gather forward
// One thread per output element.  Weights and the 4-term sum are double exactly as roi_align_kernel.cu:64-67,
// so POOL_NONE reproduces roi_align.c bit for bit; the fused pools follow ATen's 2x2 window order.
__device__ __forceinline__ float lattice_value(const float* __restrict__ plane, int W, const LatticeRoi& t, int ph,
                                               int pw) {
    if (!((t.valid_y >> ph) & 1u) || !((t.valid_x >> pw) & 1u)) return 0.f;
    const float* p = plane + (size_t)t.y.start[ph] * W + t.x.start[pw];
    float hr = t.y.frac[ph], wr = t.x.frac[pw];
    double v = __ldg(p) * (1. - hr) * (1. - wr) + __ldg(p + 1) * (1. - hr) * wr + __ldg(p + W) * hr * (1. - wr) +
               __ldg(p + W + 1) * hr * wr;
    return (float)v;
}

template <int POOL>
__global__ void __launch_bounds__(256) lattice_fwd_gather_kernel(const float* __restrict__ feat,
                                                                 const LatticeRoi* __restrict__ tab,
                                                                 float* __restrict__ out, int64_t total, int C, int H,
                                                                 int W, int PH, int PW) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int pw = (int)(idx % PW);
        int ph = (int)((idx / PW) % PH);
        int c = (int)((idx / ((int64_t)PW * PH)) % C);
        int n = (int)(idx / ((int64_t)PW * PH * C));
        const LatticeRoi& t = tab[n];
        float r = 0.f;
        if (t.batch >= 0) {
            const float* plane = feat + ((size_t)t.batch * C + c) * H * W;
            if (POOL == I2V_POOL_NONE) {
                r = lattice_value(plane, W, t, ph, pw);
            } else {
                float a = lattice_value(plane, W, t, ph, pw), b = lattice_value(plane, W, t, ph, pw + 1);
                float cc = lattice_value(plane, W, t, ph + 1, pw), d = lattice_value(plane, W, t, ph + 1, pw + 1);
                if (POOL == I2V_POOL_AVG) {
                    r = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a, b), cc), d), 4.f);
                } else {
                    r = a;
                    if (b > r) r = b;
                    if (cc > r) r = cc;
                    if (d > r) r = d;
                }
            }
        }
        out[idx] = r;
    }
}

// ------------------------------------------------------------------------------------------ gather backward
// One thread per LATTICE point (n, c, lh, lw): first the pool's backward (sum of the <=4 windows that contain
// the point -- for MAX only those whose first maximum it is), then the four atomicAdd of
// roi_align_kernel.cu:129-141 with the float-rounded double products.
template <int POOL>
__global__ void __launch_bounds__(256) lattice_bwd_gather_kernel(const float* __restrict__ grad_out,
                                                                 const float* __restrict__ feat,
                                                                 const LatticeRoi* __restrict__ tab,
                                                                 float* __restrict__ grad_in, int64_t total, int C,
                                                                 int H, int W, int GH, int GW) {
    const int PH = (POOL == I2V_POOL_NONE) ? GH : GH - 1, PW = (POOL == I2V_POOL_NONE) ? GW : GW - 1;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int lw = (int)(idx % GW);
        int lh = (int)((idx / GW) % GH);
        int c = (int)((idx / ((int64_t)GW * GH)) % C);
        int n = (int)(idx / ((int64_t)GW * GH * C));
        const LatticeRoi& t = tab[n];
        if (t.batch < 0) continue;
        if (!((t.valid_y >> lh) & 1u) || !((t.valid_x >> lw) & 1u)) continue;
        const float* g = grad_out + ((size_t)n * C + c) * PH * PW;
        float gl = 0.f;
        if (POOL == I2V_POOL_NONE) {
            gl = g[lh * PW + lw];
        } else {
            const float* plane = feat ? feat + ((size_t)t.batch * C + c) * H * W : nullptr;
            for (int i = lh - 1; i <= lh; ++i) {
                if (i < 0 || i >= PH) continue;
                for (int j = lw - 1; j <= lw; ++j) {
                    if (j < 0 || j >= PW) continue;
                    float gv = g[i * PW + j];
                    if (POOL == I2V_POOL_AVG) {
                        gl = __fadd_rn(gl, __fdiv_rn(gv, 4.f));
                    } else {
                        // first maximum of the window in row-major order (ATen max_pool2d)
                        float a = lattice_value(plane, W, t, i, j), b = lattice_value(plane, W, t, i, j + 1);
                        float cc = lattice_value(plane, W, t, i + 1, j), d = lattice_value(plane, W, t, i + 1, j + 1);
                        int best = 0;
                        float m = a;
                        if (b > m) { m = b; best = 1; }
                        if (cc > m) { m = cc; best = 2; }
                        if (d > m) { m = d; best = 3; }
                        int me = (lh - i) * 2 + (lw - j);
                        if (best == me) gl = __fadd_rn(gl, gv);
                    }
                }
            }
        }
        float hr = t.y.frac[lh], wr = t.x.frac[lw];
        float* p = grad_in + ((size_t)t.batch * C + c) * H * W + (size_t)t.y.start[lh] * W + t.x.start[lw];
        atomicAdd(p, (float)(gl * (1. - hr) * (1. - wr)));
        atomicAdd(p + 1, (float)(gl * (1. - hr) * wr));
        atomicAdd(p + W, (float)(gl * hr * (1. - wr)));
        atomicAdd(p + W + 1, (float)(gl * hr * wr));
    }
}

// ------------------------------------------------------------------------------------------ plane forward
constexpr int kPlaneK = 16;        // channels per CTA
constexpr int kPlaneWarps = 9;     // warps per CTA (each works on two RoIs at a time)
constexpr int kPlaneThreads = kPlaneWarps * 32;

__device__ __forceinline__ void bulk_store_commit(float* gdst, const float* ssrc, unsigned bytes) {
    unsigned saddr = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cp_async4(void* sdst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async8(void* sdst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// P = pooled size (output is P x P); lattice G = P (+1 when a pool follows).
// Shared-memory planes: [row][64 columns][16 channels] fp32 (row pitch padded to 64 cells, so the parity of a cell is
// the parity of its column).  Lanes are channels: a warp works on TWO RoIs x 16 channels at a time, lanes 0-15 on one
// RoI and lanes 16-31 on another.  Of the two horizontally adjacent cells of a bilinear sample the low half-warp
// always reads the even-column cell first and the high half-warp the odd-column cell first (the prep kernel stores
// both orders), so the two halves hit disjoint bank groups in every load: conflict free by construction, and all
// index math is uniform per half-warp.  Each lane ends up with the P*P outputs of its (RoI, channel); they are staged
// as [16][P*P] -- contiguous in the NCHW output -- and leave through a TMA bulk store per half.
// Cells per shared-memory row: 64 for landscape maps (W <= 64, H <= 38 at 16 channels), 40 for portrait ones (W <= 40,
// H <= 63: a 1000 x 600 frame gives a 63 x 38 map).  Any even pitch keeps "parity of a cell = parity of its column".
constexpr int kCellBytes = kPlaneK * 4;         // 64 bytes per cell
__host__ __device__ constexpr int plane_pitch_for(int W) { return W <= 40 ? 40 : 64; }

template <int P, int POOL, int kPitch>
__global__ void __launch_bounds__(kPlaneThreads, 1)
    lattice_fwd_plane_kernel(const float* __restrict__ feat, const PlaneTab* __restrict__ ptab,
                             const int* __restrict__ order, const int* __restrict__ starts, float* __restrict__ out,
                             int batch, int C, int H, int W, int split) {
    constexpr int G = (POOL == I2V_POOL_NONE) ? P : P + 1;
    constexpr int NOUT = P * P;
    constexpr int TILE = kPlaneK * NOUT;  // floats per staged output tile
    constexpr int kRowBytes = kPitch * kCellBytes;
    constexpr int TABF = (int)(sizeof(PlaneTab) / sizeof(float));
    constexpr int TABV = (int)(sizeof(PlaneTab) / 16);
    static_assert(G <= 8, "tables hold 8 lattice points per axis");
    extern __shared__ __align__(128) float smem[];
    float* planes = smem;                                    // [H][64][16]
    float* stage = smem + (size_t)H * kPitch * kPlaneK;      // [warps][2][TILE]; doubles as the fill scratch
    float* tabs = stage + (size_t)kPlaneWarps * 2 * TILE;    // [warps][2 halves][2 buffers][PlaneTab]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ctiles = C / kPlaneK;
    const int s = blockIdx.x % split;
    const int ct = (blockIdx.x / split) % ctiles;
    const int b = blockIdx.x / (split * ctiles);
    const int list_lo = __ldg(starts + b), list_hi = __ldg(starts + b + 1);
    if (list_lo == list_hi) return;
    const int role = lane >> 4, c = lane & 15;
    const int ghalf = (s * kPlaneWarps + warp) * 2 + role, gstride = split * kPlaneWarps * 2;

    if (b == batch) {  // RoIs with an out-of-range batch index: zero rows
        for (int li = list_lo + ghalf; li < list_hi; li += gstride) {
            float4* dst = reinterpret_cast<float4*>(out + ((size_t)__ldg(order + li) * C + (size_t)ct * kPlaneK) * NOUT);
            for (int i = c; i < TILE / 4; i += 16) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }

    // ---- fill: 16 planes, global [c][row][col] -> shared [row][col][16], straight into the final layout with 4-byte
    // cp.async: a warp-wide copy moves two adjacent cells x 16 channels = 128 contiguous shared bytes (conflict free);
    // on the global side it touches 8 bytes of 16 sectors, whose remaining bytes are picked up from L1 by the next
    // three copies.  Every copy of the CTA is in flight before the single wait: latency is paid once ----
    {
        const int HW = H * W;
        const int tc = lane & 15, tdx = lane >> 4;
        const float* src = feat + ((size_t)b * C + (size_t)ct * kPlaneK + tc) * HW + tdx;
        float* dst = planes + tdx * kPlaneK + tc;
        for (int row = warp; row < H; row += kPlaneWarps) {
            const float* g = src + row * W;
            float* d = dst + (size_t)row * kPitch * kPlaneK;
#pragma unroll 4
            for (int j = 0; 2 * j < W; ++j) {
                if (2 * j + tdx < W) cp_async4(d + j * 2 * kPlaneK, g + 2 * j);
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
    }

    // shared-window byte address of this lane's channel in cell (0,0)
    const unsigned lane_base = (unsigned)__cvta_generic_to_shared(planes + c);
    float* my_stage = stage + ((size_t)warp * 2 + role) * TILE;
    float* my_tabs = tabs + ((size_t)warp * 2 + role) * 2 * TABF;

    int li = list_lo + ghalf;
    int n_cur = 0, n_next = 0;
    auto fetch_table = [&](int n, int bufi) {  // 16 lanes copy the 24 x 16 bytes of one table
        const float* g = reinterpret_cast<const float*>(ptab + n);
        float* d = my_tabs + bufi * TABF;
        cp_async16(d + c * 4, g + c * 4);
        if (c + 16 < TABV) cp_async16(d + (c + 16) * 4, g + (c + 16) * 4);
    };
    if (li < list_hi) {
        n_cur = __ldg(order + li);
        fetch_table(n_cur, 0);
        if (li + gstride < list_hi) n_next = __ldg(order + li + gstride);
    } else {
        // this half never has work: a zero table keeps its lanes harmless
        for (int i = c; i < 2 * TABF; i += 16) my_tabs[i] = 0.f;
    }
    cp_async_commit();
    int it = 0;
    while (__any_sync(0xffffffffu, li < list_hi)) {
        const bool active = li < list_hi;
        const int lnext = li + gstride;
        cp_async_wait<0>();
        __syncwarp();
        const PlaneTab* t = reinterpret_cast<const PlaneTab*>(my_tabs + (it & 1) * TABF);
        // prefetch the next RoI's table into the other buffer, and the index after that
        int n_next2 = 0;
        if (active) {
            if (lnext < list_hi) {
                fetch_table(n_next, (it + 1) & 1);
                if (lnext + gstride < list_hi) n_next2 = __ldg(order + lnext + gstride);
            } else {
                float* d = my_tabs + ((it + 1) & 1) * TABF;  // the half runs dry after this RoI
                for (int i = c; i < TABF; i += 16) d[i] = 0.f;
            }
        }
        cp_async_commit();

        unsigned xfirst[G], xsecond[G];
        float wfirst[G], wsecond[G];
        {
            // asm volatile: the compiler must keep these 32 values in registers instead of re-reading the table per row
            const unsigned xe = (unsigned)__cvta_generic_to_shared(role ? t->xb : t->xa);
#pragma unroll
            for (int p = 0; p < G; ++p) {
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(xfirst[p]), "=f"(wfirst[p]), "=r"(xsecond[p]), "=f"(wsecond[p])
                             : "r"(xe + p * 16));
            }
        }
        float part[NOUT];
        float prev[G];
#pragma unroll
        for (int ph = 0; ph < G; ++ph) {
            const float4 ye = t->y[ph];
            const unsigned ybase = lane_base + (unsigned)__float_as_int(ye.x);
            const float wy0 = ye.y, wy1 = ye.z;
            float curv[G];
#pragma unroll
            for (int pw = 0; pw < G; ++pw) {
                const unsigned a = ybase + xfirst[pw], bb = ybase + xsecond[pw];
                float a0, a1, b0, b1;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a0) : "r"(a));
                asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(a1) : "r"(a), "n"(kRowBytes));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(b0) : "r"(bb));
                asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(b1) : "r"(bb), "n"(kRowBytes));
                curv[pw] = (a0 * wy0 + a1 * wy1) * wfirst[pw] + (b0 * wy0 + b1 * wy1) * wsecond[pw];
            }
            if (POOL == I2V_POOL_NONE) {
#pragma unroll
                for (int pw = 0; pw < G; ++pw) part[ph * P + pw] = curv[pw];
            } else if (POOL == I2V_POOL_AVG) {
#pragma unroll
                for (int pw = 0; pw < P; ++pw) curv[pw] += curv[pw + 1];  // row sums of adjacent columns (x 1/4 in wx)
                if (ph > 0) {
#pragma unroll
                    for (int pw = 0; pw < P; ++pw) part[(ph - 1) * P + pw] = prev[pw] + curv[pw];
                }
            } else if (ph > 0) {
#pragma unroll
                for (int pw = 0; pw < P; ++pw)
                    part[(ph - 1) * P + pw] = fmaxf(fmaxf(prev[pw], prev[pw + 1]), fmaxf(curv[pw], curv[pw + 1]));
            }
#pragma unroll
            for (int pw = 0; pw < G; ++pw) prev[pw] = curv[pw];
        }

        // ---- stage the [16][NOUT] tile of this half and hand it to the TMA ----
        if (c == 0) bulk_wait_read_all();  // the previous tile of this half has left the staging buffer
        __syncwarp();
        float* row = my_stage + c * NOUT;   // the two halves' tiles are 16 banks apart: conflict free
#pragma unroll
        for (int k = 0; k < NOUT; ++k) row[k] = part[k];
        fence_proxy_async();
        __syncwarp();
        if (c == 0 && active)
            bulk_store_commit(out + ((size_t)n_cur * C + (size_t)ct * kPlaneK) * NOUT, my_stage, TILE * sizeof(float));
        n_cur = n_next;
        n_next = n_next2;
        li = lnext;
        ++it;
    }
    if (c == 0) bulk_wait_all();
}

// ------------------------------------------------------------------------------------------ plane backward
// One CTA per (frame, 16 channels) OWNS those 16 gradient planes: they are accumulated in shared memory and written
// to HBM exactly once, so there is no global atomic, no pre-zeroing pass and the result is deterministic.
// Inside the CTA every consumer warp owns one CHANNEL PAIR, stored interleaved as float2 cells so that one
// ld.shared.v2 / fma.rn.f32x2 / st.shared.v2 updates both channels of a cell; its 32 lanes are (lattice row 0-7) x
// (column pair 0-3) and walk the frame's RoIs in list order.  For one RoI a lane forms its two lattice gradients (the
// pool's backward is a few adds), weights them per column, and adds them to the cells they touch with plain
// load / fma / store -- no atomics are needed because
//   * rows of different lattice rows are distinct cells unless the start cells repeat, and repeated start cells
//     are serialised by their position in the run (y_runpos), one __syncwarp apart;
//   * columns with a repeated start cell are merged in registers first (x_same);
//   * the upper and lower feature row of a lattice row are written in two phases, one __syncwarp apart.
// The pooled-gradient tiles ([16][P*P] floats, contiguous in NCHW) and the RoI tables stream in through a TMA bulk
// copy ring fed by a producer warp, so HBM latency is hidden without occupancy.
constexpr int kBwdK = 16;
constexpr int kBwdConsumerWarps = 8;
constexpr int kBwdThreads = (kBwdConsumerWarps + 1) * 32;
constexpr int kBwdStages = 20;  // 20 x 3.4 KB in flight per SM: enough bytes outstanding to cover HBM latency

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(sdst)),
                 "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

// Packed fp32x2 fused multiply-add (FFMA2): both channels of a cell in one instruction.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd;\n"
        "mov.b64 ra, {%2, %3};\n"
        "mov.b64 rb, {%4, %5};\n"
        "mov.b64 rc, {%6, %7};\n"
        "fma.rn.f32x2 rd, ra, rb, rc;\n"
        "mov.b64 {%0, %1}, rd;}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

__host__ __device__ inline int bwd_plane_pitch(int HW) { return HW + ((8 - HW % 32) + 32) % 32; }  // == 8 (mod 32)

template <int P, int POOL, int WT>
__global__ void __launch_bounds__(kBwdThreads, 1)
    lattice_bwd_plane_kernel(const float* __restrict__ grad_out, const BwdTab* __restrict__ tab,
                             const int* __restrict__ order, const int* __restrict__ starts,
                             float* __restrict__ grad_in, int C, int H, int Wrt) {
    constexpr int G = (POOL == I2V_POOL_NONE) ? P : P + 1;
    constexpr int NOUT = P * P;
    constexpr int TILE_BYTES = kBwdK * NOUT * (int)sizeof(float);
    constexpr int STAGE_BYTES = TILE_BYTES + (int)sizeof(BwdTab);
    static_assert(G <= 8 && TILE_BYTES % 16 == 0 && sizeof(BwdTab) % 16 == 0, "stage layout");
    extern __shared__ __align__(128) unsigned char bsmem[];
    const int W = WT ? WT : Wrt;
    const int HW = H * W;
    const int HWp = bwd_plane_pitch(HW);
    unsigned char* ring = bsmem;                                                     // [stages][STAGE_BYTES]
    uint64_t* full = reinterpret_cast<uint64_t*>(bsmem + kBwdStages * STAGE_BYTES);  // [stages]
    uint64_t* empty = full + kBwdStages;                                             // [stages]
    float* planes = reinterpret_cast<float*>(empty + kBwdStages);                    // [16][HWp]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ctiles = C / kBwdK;
    const int ct = blockIdx.x % ctiles;
    const int b = blockIdx.x / ctiles;
    const int list_lo = __ldg(starts + b), list_hi = __ldg(starts + b + 1);

    if (tid == 0) {
        for (int s = 0; s < kBwdStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, kBwdConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < kBwdK * HWp; i += kBwdThreads) planes[i] = 0.f;
    fence_proxy_async();
    __syncthreads();

    if (warp == kBwdConsumerWarps) {
        // ---- producer warp: lane j feeds ring stage j, so up to kBwdStages bulk copies are issued side by side and
        // the (slow, single-thread) barrier handshake of one stage never holds up another ----
        if (lane < kBwdStages) {
            uint64_t* my_full = full + lane;
            uint64_t* my_empty = empty + lane;
            unsigned char* dst = ring + lane * STAGE_BYTES;
            unsigned round = 0;
            for (int li = list_lo + lane; li < list_hi; li += kBwdStages, ++round) {
                const int n = __ldg(order + li);
                if (round > 0) mbar_wait(my_empty, (round - 1) & 1);
                mbar_expect_tx(my_full, STAGE_BYTES);
                bulk_load(dst, grad_out + ((size_t)n * C + (size_t)ct * kBwdK) * NOUT, TILE_BYTES, my_full);
                bulk_load(dst + TILE_BYTES, tab + n, (unsigned)sizeof(BwdTab), my_full);
            }
        }
        return;
    }

    // ---- consumers: warp w owns the channel pair (2w, 2w+1); lanes = (lattice row 0-7) x (column pair 0-3) ----
    const int ph = lane >> 2, xq = lane & 3;
    const int c0 = warp * 2;
    float2* plane = reinterpret_cast<float2*>(planes) + (size_t)warp * HWp;   // [HWp] cells of (c0, c0+1)
    // which pooled rows / columns this lane's two lattice points (ph, 2xq) and (ph, 2xq+1) collect (pool backward)
    const float m_own = (ph < P) ? 1.f : 0.f, m_up = (ph >= 1) ? 1.f : 0.f;
    const int jl = max(2 * xq - 1, 0), jm = min(2 * xq, P - 1), jr = min(2 * xq + 1, P - 1);
    const float m_l = (2 * xq - 1 >= 0) ? 1.f : 0.f, m_m = (2 * xq < P) ? 1.f : 0.f, m_r = (2 * xq + 1 < P) ? 1.f : 0.f;
    const int row_own = min(ph, P - 1) * P, row_up = max(ph - 1, 0) * P;
    int s = 0;
    unsigned round = 0;
    for (int li = list_lo; li < list_hi; ++li) {
        mbar_wait(full + s, round & 1);
        const float* tile = reinterpret_cast<const float*>(ring + s * STAGE_BYTES);
        const BwdTab* t = reinterpret_cast<const BwdTab*>(ring + s * STAGE_BYTES + TILE_BYTES);

        const unsigned vx = t->valid_x, vy = t->valid_y, xsame = t->x_same;
        const int maxrun = (int)t->y_maxrun;
        constexpr unsigned FULL = (1u << G) - 1u;
        const unsigned ax = vx & ~xsame & FULL;  // columns that still own a pair of cells after merging
        if (maxrun > 0 && ax != 0u) {
            // lattice gradients of this lane's two points, for both channels (x = channel c0, y = channel c0+1)
            float2 gl0, gl1;
            {
                const float* ta = tile + c0 * NOUT;
                const float* tb = ta + NOUT;
                if (POOL == I2V_POOL_NONE) {
                    gl0 = make_float2(ta[row_own + jm] * (m_own * m_m), tb[row_own + jm] * (m_own * m_m));
                    gl1 = make_float2(ta[row_own + jr] * (m_own * m_r), tb[row_own + jr] * (m_own * m_r));
                } else {
                    // sum over the pooled rows (ph-1, ph) first, then over the pooled columns (pw-1, pw)
                    float al = ta[row_own + jl] * m_own + ta[row_up + jl] * m_up, bl = tb[row_own + jl] * m_own + tb[row_up + jl] * m_up;
                    float am = ta[row_own + jm] * m_own + ta[row_up + jm] * m_up, bm = tb[row_own + jm] * m_own + tb[row_up + jm] * m_up;
                    float ar = ta[row_own + jr] * m_own + ta[row_up + jr] * m_up, br = tb[row_own + jr] * m_own + tb[row_up + jr] * m_up;
                    gl0 = make_float2(al * m_l + am * m_m, bl * m_l + bm * m_m);
                    gl1 = make_float2(am * m_m + ar * m_r, bm * m_m + br * m_r);
                }
            }
            // column tables of this lane's two columns: byte offsets (float2 cells) and weights (validity, 1/4 folded in)
            const int2 xo = reinterpret_cast<const int2*>(t->xoff)[xq];
            const float2 w0 = reinterpret_cast<const float2*>(t->wx0)[xq];
            const float2 w1 = reinterpret_cast<const float2*>(t->wx1)[xq];
            float2 t00 = make_float2(gl0.x * w0.x, gl0.y * w0.x), t01 = make_float2(gl0.x * w1.x, gl0.y * w1.x);  // column 2xq: left, right cell
            float2 t10 = make_float2(gl1.x * w0.y, gl1.y * w0.y), t11 = make_float2(gl1.x * w1.y, gl1.y * w1.y);  // column 2xq+1
            const bool oky = (ph < G) && ((vy >> ph) & 1u);
            const float wy0 = t->wy0[ph], wy1 = t->wy1[ph];
            char* row0 = reinterpret_cast<char*>(plane) + 2 * t->yoff[ph];   // tables hold 4-byte cell offsets
            char* pa = row0 + 2 * xo.x;
            char* pb = row0 + 2 * xo.y;
            const int row_bytes = W * 8;
            if (maxrun == 1 && ax == FULL) {
                // ---- common case: every column owns its cells, no two lattice rows share a start row.  Four phases
                // (upper/lower row x left/right cell); inside a phase all 64 cells of the warp are distinct ----
#pragma unroll
                for (int dy = 0; dy < 2; ++dy) {
                    const float2 wy = dy ? make_float2(wy1, wy1) : make_float2(wy0, wy0);
#pragma unroll
                    for (int dxx = 0; dxx < 2; ++dxx) {
                        float2* qa = reinterpret_cast<float2*>(pa + (dy ? row_bytes : 0) + dxx * 8);
                        float2* qb = reinterpret_cast<float2*>(pb + (dy ? row_bytes : 0) + dxx * 8);
                        float2 oa = *qa, ob = *qb;
                        oa = ffma2(dxx ? t01 : t00, wy, oa);
                        ob = ffma2(dxx ? t11 : t10, wy, ob);
                        if (oky) {
                            *qa = oa;
                            if (G % 2 == 0 || 2 * xq + 1 < G) *qb = ob;   // odd lattices have no column G
                        }
                        __syncwarp();
                    }
                }
            } else {
                // ---- general case: merge columns that share a start cell into the first column of their run (the
                // run may cross lanes), serialise lattice rows that share a start row by their position in the run ----
                if (xsame != 0u) {
#pragma unroll
                    for (int pw = G - 1; pw >= 1; --pw) {
                        if ((xsame >> pw) & 1u) {   // uniform: column pw folds into column pw-1
                            if (pw & 1) {            // both columns live in the same lane
                                if (xq == (pw >> 1)) {
                                    t00.x += t10.x; t00.y += t10.y; t01.x += t11.x; t01.y += t11.y;
                                }
                            } else {                 // column pw is the first of lane pw/2, column pw-1 the second of lane pw/2 - 1
                                float sx0 = __shfl_down_sync(0xffffffffu, t00.x, 1), sy0 = __shfl_down_sync(0xffffffffu, t00.y, 1);
                                float sx1 = __shfl_down_sync(0xffffffffu, t01.x, 1), sy1 = __shfl_down_sync(0xffffffffu, t01.y, 1);
                                if (xq == (pw >> 1) - 1) {
                                    t10.x += sx0; t10.y += sy0; t11.x += sx1; t11.y += sy1;
                                }
                            }
                        }
                    }
                }
                const bool own_a = (ax >> (2 * xq)) & 1u, own_b = (ax >> (2 * xq + 1)) & 1u;
                const int myrun = (int)((t->y_runpos >> (4 * ph)) & 15u);
                for (int k = 0; k < maxrun; ++k) {
                    const bool act = oky && (myrun == k);
#pragma unroll
                    for (int dy = 0; dy < 2; ++dy) {
                        const float2 wy = dy ? make_float2(wy1, wy1) : make_float2(wy0, wy0);
#pragma unroll
                        for (int dxx = 0; dxx < 2; ++dxx) {
                            float2* qa = reinterpret_cast<float2*>(pa + (dy ? row_bytes : 0) + dxx * 8);
                            float2* qb = reinterpret_cast<float2*>(pb + (dy ? row_bytes : 0) + dxx * 8);
                            if (act && own_a) *qa = ffma2(dxx ? t01 : t00, wy, *qa);
                            if (act && own_b) *qb = ffma2(dxx ? t11 : t10, wy, *qb);
                            __syncwarp();
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
        if (++s == kBwdStages) {
            s = 0;
            ++round;
        }
    }
    // ---- write this warp's two planes: the only write of these gradient bytes ----
    __syncwarp();
    float* dst = grad_in + ((size_t)b * C + (size_t)ct * kBwdK + (size_t)c0) * HW;
    for (int i = lane; i < HW; i += 32) {
        float2 v = plane[i];
        dst[i] = v.x;
        dst[(size_t)HW + i] = v.y;
    }
}

// ------------------------------------------------------------------------------------------ host side
// roi_align_bwd_rows.cu: the row-owner backward
size_t bwd_rows_smem_bytes(int H, int W);
bool bwd_rows_ok(const float* grad_out, int batch, int C, int H, int W, int PH, int PW, int pool_mode);
int launch_bwd_rows(const float* grad_out, const LatticeRoi* tab, void* rtab_space, const int* order, const int* starts,
                    float* grad_in, int batch, int C, int H, int W, int num_rois, int pool_mode, cudaStream_t stream);

// roi_align_fwd_slab.cu: the forward on planes brought in by one bulk copy
bool fwd_slab_ok(const float* features, const float* out, int batch, int C, int H, int W, int PH, int PW, int pool_mode);
int launch_fwd_slab(const float* feat, const LatticeRoi* tab, void* tab_space, const int* order, const int* starts, float* out,
                    int batch, int C, int H, int W, int num_rois, int pool_mode, cudaStream_t stream);
// roi_align_fwd_even.cu: the forward on planes re-pitched to an even row pitch (conflict free by construction)
bool fwd_chan_ok(const float* features, const float* out, int batch, int C, int H, int W, int PH, int PW, int pool_mode);
int launch_fwd_chan(const float* feat, const LatticeRoi* tab, void* tab_space, const int* order, const int* starts, float* out,
                    int batch, int C, int H, int W, int num_rois, int pool_mode, cudaStream_t stream);
bool fwd_even_ok(const float* features, const float* out, int batch, int C, int H, int W, int PH, int PW, int pool_mode);
int launch_fwd_even(const float* feat, const LatticeRoi* tab, void* tab_space, const int* order, const int* starts, float* out,
                    int batch, int C, int H, int W, int num_rois, int pool_mode, cudaStream_t stream);
// roi_align_bwd_phase.cu: the phased backward (warp = lattice row, CTA barrier between feature-row phases)
size_t bwd_phase_smem_bytes(int H, int W);
bool bwd_phase_ok(const float* grad_out, int batch, int C, int H, int W, int PH, int PW, int pool_mode);
int launch_bwd_phase(const float* grad_out, const LatticeRoi* tab, void* tab_space, const int* order, const int* starts,
                     float* grad_in, int batch, int C, int H, int W, int num_rois, int pool_mode, int accumulate,
                     cudaStream_t stream);

// roi_align_bwd_band.cu: the band-owner backward (lanes = 32 channels, a CTA owns a slab of feature rows, a warp a band)
bool bwd_band_ok(const float* grad_out, int batch, int C, int H, int W, int PH, int PW, int pool_mode);
size_t bwd_band_list_ints(int batch, int num_rois);
int launch_bwd_band(const float* grad_out, const LatticeRoi* tab, void* tab_space, const int* order, const int* starts,
                    int* lists, float* grad_in, int batch, int C, int H, int W, int num_rois, int pool_mode, int bands,
                    cudaStream_t stream);

struct LatticeWs {
    LatticeRoi* tab;
    PlaneTab* ptab;   // the plane tables share one allocation: the forward view or the backward view of a call
    BwdTab* btab;
    int* roi_batch;
    int* counts;
    int* starts;
    int* order;
    int* band_lists;  // per-(frame, slab) RoI lists of the band-owner backward
    size_t bytes;
};

// The tables come first; the per-frame lists (only the plane kernels use them) follow when `lists` is set.
static LatticeWs carve_lattice_ws(void* ws, int batch, int num_rois, bool lists = true) {
    Carver cv(ws);
    LatticeWs w{};
    w.tab = cv.take<LatticeRoi>((size_t)num_rois);
    if (lists) {
        w.roi_batch = cv.take<int>((size_t)num_rois);
        w.counts = cv.take<int>((size_t)batch + 2);
        w.starts = cv.take<int>((size_t)batch + 2);
        w.order = cv.take<int>((size_t)num_rois);
        // one slot per RoI, shared by whichever per-kernel view of the tables a call builds (PlaneTab, BwdTab, or the
        // RowTab / PhaseTab of roi_align_bwd_rows.cu / roi_align_bwd_phase.cu)
        w.ptab = reinterpret_cast<PlaneTab*>(cv.take<unsigned char>((size_t)num_rois * kRoiTabSlotBytes));
        w.btab = reinterpret_cast<BwdTab*>(w.ptab);
        w.band_lists = cv.take<int>(bwd_band_list_ints(batch, num_rois));
        static_assert(sizeof(BwdTab) <= kRoiTabSlotBytes && sizeof(PlaneTab) <= kRoiTabSlotBytes, "table slot");
    }
    w.bytes = cv.used();
    return w;
}

// What the last lattice_prep of this host thread left in which workspace: lets the backward of a step take the forward's
// tables and lists (rois == NULL) and refuse when they are not there.
struct LastPrep {
    const void* tab;
    cudaStream_t stream;
    int batch, num_rois, H, W, GH, GW;
    float scale;
    bool lists;
};
static thread_local LastPrep g_last_prep{};

static int lattice_prep(const float* rois, int batch, int num_rois, int H, int W, int GH, int GW, float scale,
                        const LatticeWs& w, bool lists, cudaStream_t stream, float ptab_scale = 0.f,
                        bool backward = false) {
    if (lists) I2V_CUDA_TRY(cudaMemsetAsync(w.counts, 0, sizeof(int) * ((size_t)batch + 2), stream));
    lattice_prep_kernel<<<ceil_div(num_rois, 128), 128, 0, stream>>>(rois, num_rois, batch, H, W, GH, GW, scale, w.tab,
                                                                     lists ? w.roi_batch : nullptr,
                                                                     lists ? w.counts : nullptr,
                                                                     (lists && ptab_scale != 0.f && !backward) ? w.ptab : nullptr,
                                                                     ptab_scale,
                                                                     (lists && ptab_scale != 0.f && backward) ? w.btab : nullptr,
                                                                     plane_pitch_for(W) * kCellBytes);
    I2V_TRY(check_launch("lattice_prep_kernel"));
    if (lists) {
        roi_bucket_kernel<<<batch + 1, 256, 0, stream>>>(w.roi_batch, num_rois, batch + 1, w.counts, w.starts, w.order);
        I2V_TRY(check_launch("roi_bucket_kernel"));
    }
    g_last_prep = LastPrep{w.tab, stream, batch, num_rois, H, W, GH, GW, scale, lists};
    return I2V_OK;
}

static size_t plane_fwd_smem_bytes(int H, int W, int P) {
    size_t stage = (size_t)kPlaneWarps * 2 * kPlaneK * P * P;
    return ((size_t)H * plane_pitch_for(W) * kPlaneK + stage) * sizeof(float) + (size_t)kPlaneWarps * 4 * sizeof(PlaneTab);
}

static bool plane_forward_ok(const float* out, int batch, int C, int H, int W, int PH, int PW) {
    return batch > 0 && PH == 7 && PW == 7 && C % kPlaneK == 0 && H >= 2 && W >= 2 && W <= 64 &&
           plane_fwd_smem_bytes(H, W, 7) <= (size_t)kMaxSmemPerCta && ((uintptr_t)out & 15) == 0;
}

static int plane_split(int ctas) {
    int split = 1;
    while (ctas * split < 2 * kNumSMs && split < 8) split *= 2;
    return split;
}

template <int POOL, int PITCH>
static int launch_plane_fwd_p(const float* feat, const LatticeWs& w, float* out, int batch, int C, int H, int W,
                              cudaStream_t stream) {
    auto kern = lattice_fwd_plane_kernel<7, POOL, PITCH>;
    size_t smem = plane_fwd_smem_bytes(H, W, 7);
    I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ctiles = C / kPlaneK;
    int split = plane_split(batch * ctiles);
    kern<<<(batch + 1) * ctiles * split, kPlaneThreads, smem, stream>>>(feat, w.ptab, w.order, w.starts, out, batch, C,
                                                                       H, W, split);
    return check_launch("lattice_fwd_plane_kernel");
}
template <int POOL>
static int launch_plane_fwd(const float* feat, const LatticeWs& w, float* out, int batch, int C, int H, int W,
                            cudaStream_t stream) {
    if (plane_pitch_for(W) == 40) return launch_plane_fwd_p<POOL, 40>(feat, w, out, batch, C, H, W, stream);
    return launch_plane_fwd_p<POOL, 64>(feat, w, out, batch, C, H, W, stream);
}

static size_t plane_bwd_smem_bytes(int H, int W, int P) {
    return (size_t)kBwdStages * ((size_t)kBwdK * P * P * sizeof(float) + sizeof(BwdTab)) +
           2 * kBwdStages * sizeof(uint64_t) + (size_t)kBwdK * bwd_plane_pitch(H * W) * sizeof(float);
}

static bool plane_backward_ok(const float* grad_out, int batch, int C, int H, int W, int PH, int PW, int pool_mode) {
    return batch > 0 && PH == 7 && PW == 7 && pool_mode != I2V_POOL_MAX && C % kBwdK == 0 && H >= 2 && W >= 2 &&
           plane_bwd_smem_bytes(H, W, 7) <= (size_t)kMaxSmemPerCta && ((uintptr_t)grad_out & 15) == 0;
}

template <int POOL, int WT>
static int launch_plane_bwd_w(const float* grad_out, const LatticeWs& w, float* grad_in, int batch, int C, int H, int W,
                              cudaStream_t stream) {
    auto kern = lattice_bwd_plane_kernel<7, POOL, WT>;
    size_t smem = plane_bwd_smem_bytes(H, W, 7);
    I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<batch * (C / kBwdK), kBwdThreads, smem, stream>>>(grad_out, w.btab, w.order, w.starts, grad_in, C, H, W);
    return check_launch("lattice_bwd_plane_kernel");
}
template <int POOL>
static int launch_plane_bwd(const float* grad_out, const LatticeWs& w, float* grad_in, int batch, int C, int H, int W,
                            cudaStream_t stream) {
    if (W == 63) return launch_plane_bwd_w<POOL, 63>(grad_out, w, grad_in, batch, C, H, W, stream);
    return launch_plane_bwd_w<POOL, 0>(grad_out, w, grad_in, batch, C, H, W, stream);
}

}  // namespace i2v

using namespace i2v;

extern "C" size_t i2v_roi_align_workspace_bytes(int batch, int num_rois) {
    if (batch < 0 || num_rois < 0) return 0;
    return carve_lattice_ws(nullptr, batch, num_rois).bytes;
}

static int roi_align_check(const char* who, const void* a, const void* b, const void* c, int batch, int channels,
                           int height, int width, int num_rois, int pooled_h, int pooled_w, int pool_mode, int impl,
                           int& GH, int& GW) {
    I2V_REQUIRE(batch >= 0 && channels >= 0 && num_rois >= 0, "%s: negative size", who);
    I2V_REQUIRE(pool_mode >= I2V_POOL_NONE && pool_mode <= I2V_POOL_MAX, "%s: bad pool_mode %d", who, pool_mode);
    I2V_REQUIRE(impl >= I2V_IMPL_AUTO && impl <= I2V_IMPL_CHAN, "%s: bad impl %d", who, impl);
    GH = pooled_h + (pool_mode != I2V_POOL_NONE);
    GW = pooled_w + (pool_mode != I2V_POOL_NONE);
    I2V_REQUIRE(pooled_h >= 1 && pooled_w >= 1 && GH >= 2 && GW >= 2 && GH <= kMaxLattice && GW <= kMaxLattice,
                "%s: lattice %dx%d outside [2,%d]", who, GH, GW, kMaxLattice);
    if (num_rois > 0 && channels > 0) {
        I2V_REQUIRE(a && b && c, "%s: null pointer", who);
        I2V_REQUIRE(height >= 2 && width >= 2, "%s: feature map %dx%d smaller than 2x2", who, height, width);
    }
    return I2V_OK;
}

static int carve_checked(const char* who, void* workspace, size_t workspace_bytes, int batch, int num_rois, bool lists,
                         LatticeWs& w) {
    w = carve_lattice_ws(workspace, batch, num_rois, lists);
    if (!workspace || workspace_bytes < w.bytes) {
        set_error("%s: workspace %zu < %zu bytes", who, workspace_bytes, w.bytes);
        return I2V_ERR_WORKSPACE;
    }
    return I2V_OK;
}

extern "C" int i2v_roi_align_forward(const float* features, const float* rois, float* out, int batch, int channels,
                                     int height, int width, int num_rois, int pooled_h, int pooled_w,
                                     float spatial_scale, int pool_mode, int impl, void* workspace,
                                     size_t workspace_bytes, cudaStream_t stream) {
    LatticeWs w;
    int GH, GW;
    I2V_TRY(roi_align_check("roi_align_forward", features, rois, out, batch, channels, height, width, num_rois,
                            pooled_h, pooled_w, pool_mode, impl, GH, GW));
    if (num_rois == 0 || channels == 0) return I2V_OK;
    bool can_plane = plane_forward_ok(out, batch, channels, height, width, pooled_h, pooled_w);
    const bool can_slab = fwd_slab_ok(features, out, batch, channels, height, width, pooled_h, pooled_w, pool_mode);
    if (impl == I2V_IMPL_SLAB && !can_slab) {
        set_error("roi_align_forward: the slab kernel needs a 7x7 output, pool none/avg, C %% 16 == 0, H*W = 2 (mod 4) and 16 "
                  "planes that fit shared memory");
        return I2V_ERR_UNSUPPORTED;
    }
    const bool can_even = fwd_even_ok(features, out, batch, channels, height, width, pooled_h, pooled_w, pool_mode);
    if (impl == I2V_IMPL_EVEN && !can_even) {
        set_error("roi_align_forward: the even-pitch kernel needs a 7x7 output, pool none/avg, C %% 16 == 0, W <= 64, H <= 40 "
                  "and 16 re-pitched planes that fit shared memory");
        return I2V_ERR_UNSUPPORTED;
    }
    // measured on config 2: 0.725 ms against 0.675 ms for the slab kernel -- the collisions are gone (L1/TEX 63 % instead of
    // 86 %) but the second address per lattice point and the re-pitch pass cost more than they return with ten warps per SM;
    // selectable, not what AUTO takes
    if (impl == I2V_IMPL_EVEN) {
        I2V_TRY(carve_checked("roi_align_forward", workspace, workspace_bytes, batch, num_rois, true, w));
        I2V_TRY(lattice_prep(rois, batch, num_rois, height, width, GH, GW, spatial_scale, w, true, stream));
        return launch_fwd_even(features, w.tab, w.ptab, w.order, w.starts, out, batch, channels, height, width, num_rois,
                               pool_mode, stream);
    }
    if (impl == I2V_IMPL_CHAN) {
        if (!fwd_chan_ok(features, out, batch, channels, height, width, pooled_h, pooled_w, pool_mode)) {
            set_error("roi_align_forward: the lane-per-channel kernel needs a 7x7 output, pool none/avg, C %% 32 == 0, W <= 64 and "
                      "half the rows of 32 planes in shared memory");
            return I2V_ERR_UNSUPPORTED;
        }
        I2V_TRY(carve_checked("roi_align_forward", workspace, workspace_bytes, batch, num_rois, true, w));
        I2V_TRY(lattice_prep(rois, batch, num_rois, height, width, GH, GW, spatial_scale, w, true, stream));
        return launch_fwd_chan(features, w.tab, w.ptab, w.order, w.starts, out, batch, channels, height, width, num_rois,
                               pool_mode, stream);
    }
    if (impl >= I2V_IMPL_PLANE && impl != I2V_IMPL_SLAB && !can_plane) {
        set_error("roi_align_forward: the plane kernel needs a 7x7 output, C %% 16 == 0, a 16-byte aligned output and "
                  "16 planes that fit shared memory");
        return I2V_ERR_UNSUPPORTED;
    }
    // AUTO prefers the slab kernel (one TMA bulk copy fills the planes; 16 warps, one RoI each) where its bank argument
    // holds; `impl = plane` keeps the cell-major plane kernel, which also serves the other shapes and the max pool
    if ((impl == I2V_IMPL_AUTO || impl == I2V_IMPL_SLAB) && can_slab) {
        I2V_TRY(carve_checked("roi_align_forward", workspace, workspace_bytes, batch, num_rois, true, w));
        I2V_TRY(lattice_prep(rois, batch, num_rois, height, width, GH, GW, spatial_scale, w, true, stream));
        return launch_fwd_slab(features, w.tab, w.ptab, w.order, w.starts, out, batch, channels, height, width, num_rois,
                               pool_mode, stream);
    }
    bool plane = can_plane && impl != I2V_IMPL_GATHER;
    I2V_TRY(carve_checked("roi_align_forward", workspace, workspace_bytes, batch, num_rois, plane, w));
    I2V_TRY(lattice_prep(rois, batch, num_rois, height, width, GH, GW, spatial_scale, w, plane, stream,
                         pool_mode == I2V_POOL_AVG ? 0.25f : 1.f));
    if (plane) {
        if (pool_mode == I2V_POOL_AVG) return launch_plane_fwd<I2V_POOL_AVG>(features, w, out, batch, channels, height, width, stream);
        if (pool_mode == I2V_POOL_MAX) return launch_plane_fwd<I2V_POOL_MAX>(features, w, out, batch, channels, height, width, stream);
        return launch_plane_fwd<I2V_POOL_NONE>(features, w, out, batch, channels, height, width, stream);
    }
    int64_t total = (int64_t)num_rois * channels * pooled_h * pooled_w;
    int grid = grid_for(total, 256);
    if (pool_mode == I2V_POOL_AVG)
        lattice_fwd_gather_kernel<I2V_POOL_AVG><<<grid, 256, 0, stream>>>(features, w.tab, out, total, channels, height, width, pooled_h, pooled_w);
    else if (pool_mode == I2V_POOL_MAX)
        lattice_fwd_gather_kernel<I2V_POOL_MAX><<<grid, 256, 0, stream>>>(features, w.tab, out, total, channels, height, width, pooled_h, pooled_w);
    else
        lattice_fwd_gather_kernel<I2V_POOL_NONE><<<grid, 256, 0, stream>>>(features, w.tab, out, total, channels, height, width, pooled_h, pooled_w);
    return check_launch("lattice_fwd_gather_kernel");
}

static int roi_align_backward_impl(const float* grad_out, const float* features, const float* rois, float* grad_in,
                                   int batch, int channels, int height, int width, int num_rois, int pooled_h,
                                   int pooled_w, float spatial_scale, int pool_mode, int impl, void* workspace,
                                   size_t workspace_bytes, bool zero_first, cudaStream_t stream) {
    LatticeWs w;
    int GH, GW;
    // rois == NULL: the workspace still holds the tables and per-frame lists that i2v_roi_align_forward left there for the
    // same RoIs and geometry (a training step runs the two back to back); only the phased kernel takes them as they are
    const bool reuse = rois == nullptr && num_rois > 0;
    I2V_TRY(roi_align_check("roi_align_backward", grad_out, reuse ? grad_out : rois, grad_in, batch, channels, height, width,
                            num_rois, pooled_h, pooled_w, pool_mode, impl, GH, GW));
    I2V_REQUIRE(pool_mode != I2V_POOL_MAX || features || num_rois == 0, "roi_align_backward: POOL_MAX needs the features");
    size_t in_elems = (size_t)batch * channels * height * width;
    if (in_elems == 0) return I2V_OK;
    I2V_REQUIRE(grad_in, "roi_align_backward: null grad_in");
    // the plane kernel overwrites grad_in, so it cannot serve the accumulate-into-caller's-buffer launcher
    bool can_plane = zero_first && num_rois > 0 &&
                     plane_backward_ok(grad_out, batch, channels, height, width, pooled_h, pooled_w, pool_mode);
    bool can_rows = zero_first && num_rois > 0 &&
                    bwd_rows_ok(grad_out, batch, channels, height, width, pooled_h, pooled_w, pool_mode);
    // the phased kernel can also ADD its planes to grad_in (the reference launcher's contract)
    bool can_phase = num_rois > 0 && bwd_phase_ok(grad_out, batch, channels, height, width, pooled_h, pooled_w, pool_mode);
    bool can_band = zero_first && num_rois > 0 &&
                    bwd_band_ok(grad_out, batch, channels, height, width, pooled_h, pooled_w, pool_mode);
    if (impl == I2V_IMPL_SLAB || impl == I2V_IMPL_EVEN || impl == I2V_IMPL_CHAN) {
        set_error("roi_align_backward: I2V_IMPL_SLAB / I2V_IMPL_EVEN / I2V_IMPL_CHAN are forward kernels");
        return I2V_ERR_UNSUPPORTED;
    }
    if ((impl == I2V_IMPL_PLANE && !can_plane) || (impl == I2V_IMPL_ROWS && !can_rows) ||
        (impl == I2V_IMPL_PHASE && !can_phase) || (impl == I2V_IMPL_BAND && !can_band)) {
        set_error("roi_align_backward: the plane kernels need a 7x7 pooled size, pool none/avg, C %% 16 == 0, a 16-byte "
                  "aligned gradient and 16 planes that fit shared memory");
        return I2V_ERR_UNSUPPORTED;
    }
    // AUTO takes the phased kernel: 1.59 ms on config 2 against 2.28 ms for the warp-per-channel-pair plane kernel and
    // 2.35 ms for the row-owner kernel (profiles/README.md); the other two stay selectable and serve as cross-checks
    if (reuse && !(can_phase && (impl == I2V_IMPL_PHASE || impl == I2V_IMPL_AUTO))) {
        set_error("roi_align_backward: rois == NULL (tables of the preceding forward call) needs the phased kernel");
        return I2V_ERR_UNSUPPORTED;
    }
    if (can_phase && (impl == I2V_IMPL_PHASE || impl == I2V_IMPL_AUTO)) {
        I2V_TRY(carve_checked("roi_align_backward", workspace, workspace_bytes, batch, num_rois, true, w));
        if (reuse) {
            const LastPrep& lp = g_last_prep;
            if (!(lp.lists && lp.tab == w.tab && lp.stream == stream && lp.batch == batch && lp.num_rois == num_rois &&
                  lp.H == height && lp.W == width && lp.GH == GH && lp.GW == GW && lp.scale == spatial_scale)) {
                set_error("roi_align_backward: rois == NULL, but this workspace does not hold the tables of a forward call with "
                          "the same RoI count, batch, map, pooled size, scale and stream");
                return I2V_ERR_INVALID;
            }
        } else {
            I2V_TRY(lattice_prep(rois, batch, num_rois, height, width, GH, GW, spatial_scale, w, true, stream));
        }
        return launch_bwd_phase(grad_out, w.tab, w.ptab, w.order, w.starts, grad_in, batch, channels, height, width,
                                num_rois, pool_mode, zero_first ? 0 : 1, stream);
    }
    // the band-owner kernel (roi_align_bwd_band.cu) keeps maps of any size plane-resident (a CTA owns a slab of rows):
    // AUTO takes it where the phased kernel cannot hold 16 whole planes (1.7 ms against 1.49 ms on config 2 otherwise)
    if (can_band && (impl == I2V_IMPL_BAND || impl == I2V_IMPL_AUTO)) {
        I2V_TRY(carve_checked("roi_align_backward", workspace, workspace_bytes, batch, num_rois, true, w));
        I2V_TRY(lattice_prep(rois, batch, num_rois, height, width, GH, GW, spatial_scale, w, true, stream));
        static const int bands = [] {
            const char* e = getenv("I2V_BAND_WARPS");
            return e ? atoi(e) : 0;
        }();
        return launch_bwd_band(grad_out, w.tab, w.ptab, w.order, w.starts, w.band_lists, grad_in, batch, channels, height,
                               width, num_rois, pool_mode, bands, stream);
    }
    if (can_rows && (impl == I2V_IMPL_ROWS || (impl == I2V_IMPL_AUTO && !can_plane))) {
        I2V_TRY(carve_checked("roi_align_backward", workspace, workspace_bytes, batch, num_rois, true, w));
        I2V_TRY(lattice_prep(rois, batch, num_rois, height, width, GH, GW, spatial_scale, w, true, stream));
        return launch_bwd_rows(grad_out, w.tab, w.ptab, w.order, w.starts, grad_in, batch, channels, height, width,
                               num_rois, pool_mode, stream);
    }
    if (can_plane && impl != I2V_IMPL_GATHER) {
        I2V_TRY(carve_checked("roi_align_backward", workspace, workspace_bytes, batch, num_rois, true, w));
        I2V_TRY(lattice_prep(rois, batch, num_rois, height, width, GH, GW, spatial_scale, w, true, stream,
                             pool_mode == I2V_POOL_AVG ? 0.25f : 1.f, true));
        if (pool_mode == I2V_POOL_AVG) return launch_plane_bwd<I2V_POOL_AVG>(grad_out, w, grad_in, batch, channels, height, width, stream);
        return launch_plane_bwd<I2V_POOL_NONE>(grad_out, w, grad_in, batch, channels, height, width, stream);
    }
    if (zero_first) I2V_CUDA_TRY(cudaMemsetAsync(grad_in, 0, in_elems * sizeof(float), stream));
    if (num_rois == 0 || channels == 0) return I2V_OK;
    I2V_TRY(carve_checked("roi_align_backward", workspace, workspace_bytes, batch, num_rois, false, w));
    I2V_TRY(lattice_prep(rois, batch, num_rois, height, width, GH, GW, spatial_scale, w, false, stream));
    int64_t total = (int64_t)num_rois * channels * GH * GW;
    int grid = grid_for(total, 256);
    if (pool_mode == I2V_POOL_AVG)
        lattice_bwd_gather_kernel<I2V_POOL_AVG><<<grid, 256, 0, stream>>>(grad_out, features, w.tab, grad_in, total, channels, height, width, GH, GW);
    else if (pool_mode == I2V_POOL_MAX)
        lattice_bwd_gather_kernel<I2V_POOL_MAX><<<grid, 256, 0, stream>>>(grad_out, features, w.tab, grad_in, total, channels, height, width, GH, GW);
    else
        lattice_bwd_gather_kernel<I2V_POOL_NONE><<<grid, 256, 0, stream>>>(grad_out, features, w.tab, grad_in, total, channels, height, width, GH, GW);
    return check_launch("lattice_bwd_gather_kernel");
}

extern "C" int i2v_roi_align_backward(const float* grad_out, const float* features, const float* rois, float* grad_in,
                                      int batch, int channels, int height, int width, int num_rois, int pooled_h,
                                      int pooled_w, float spatial_scale, int pool_mode, int impl, void* workspace,
                                      size_t workspace_bytes, cudaStream_t stream) {
    return roi_align_backward_impl(grad_out, features, rois, grad_in, batch, channels, height, width, num_rois, pooled_h,
                                   pooled_w, spatial_scale, pool_mode, impl, workspace, workspace_bytes, true, stream);
}

// ---- the reference-signature launchers (roi_align_kernel.h:13-27) ----
// roi_align_kernel.h:13-17 does not pass the batch size, and the plane-resident kernels are launched per frame: the
// largest frame index of the RoI list is found on the device and read back (one 4-byte copy and one stream
// synchronisation per call).
__global__ void max_frame_kernel(const float* __restrict__ rois, int num_rois, int* __restrict__ out) {
    int m = -1;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < num_rois; n += gridDim.x * blockDim.x)
        m = max(m, (int)fminf(fmaxf(__ldg(rois + (size_t)n * 5), -1.f), 1.0e9f));
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m >= 0) atomicMax(out, m);
}
constexpr int kLegacyMaxFrames = 4096;   // beyond this the launcher keeps to the gather kernel (no per-frame lists)

int i2v::legacy_frame_count(const float* rois, int num_rois, int* scratch, cudaStream_t stream, int* frames) {
    *frames = 0;
    if (num_rois <= 0) return I2V_OK;
    int h_max = -1;
    I2V_CUDA_TRY(cudaMemsetAsync(scratch, 0xff, sizeof(int), stream));
    max_frame_kernel<<<grid_for(num_rois, 256, 1), 256, 0, stream>>>(rois, num_rois, scratch);
    I2V_TRY(check_launch("max_frame_kernel"));
    I2V_CUDA_TRY(cudaMemcpyAsync(&h_max, scratch, sizeof(int), cudaMemcpyDeviceToHost, stream));
    I2V_CUDA_TRY(cudaStreamSynchronize(stream));
    *frames = h_max + 1;
    return I2V_OK;
}

extern "C" int ROIAlignForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois,
                                      const int height, const int width, const int channels, const int aligned_height,
                                      const int aligned_width, const float* bottom_rois, float* top_data,
                                      cudaStream_t stream) {
    if (num_rois < 0) return 0;
    size_t have = 0;
    const size_t gather_bytes = carve_lattice_ws(nullptr, 0, num_rois, false).bytes;
    char* ws = static_cast<char*>(legacy_scratch(gather_bytes + 256, &have, stream));
    if (!ws) return 0;
    int frames = INT32_MAX;     // roi_align_kernel.h:13-17: every non-negative frame index is accepted
    if (num_rois > 0 && aligned_height == 7 && aligned_width == 7 && channels % 16 == 0) {
        int found = 0;
        if (legacy_frame_count(bottom_rois, num_rois, reinterpret_cast<int*>(ws), stream, &found) != I2V_OK) return 0;
        if (found >= 1 && found <= kLegacyMaxFrames) {
            frames = found;
            const size_t need = i2v_roi_align_workspace_bytes(frames, num_rois);
            ws = static_cast<char*>(legacy_scratch(need + 256, &have, stream));
            if (!ws) return 0;
            int rc = i2v_roi_align_forward(bottom_data, bottom_rois, top_data, frames, channels, height, width, num_rois,
                                           aligned_height, aligned_width, spatial_scale, I2V_POOL_NONE, I2V_IMPL_AUTO,
                                           ws + 256, have - 256, stream);
            return rc == I2V_OK ? 1 : 0;
        }
    }
    int rc = i2v_roi_align_forward(bottom_data, bottom_rois, top_data, INT32_MAX, channels, height, width, num_rois,
                                   aligned_height, aligned_width, spatial_scale, I2V_POOL_NONE, I2V_IMPL_GATHER, ws + 256,
                                   have - 256, stream);
    return rc == I2V_OK ? 1 : 0;
}

extern "C" int ROIAlignBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size,
                                       const int num_rois, const int height, const int width, const int channels,
                                       const int aligned_height, const int aligned_width, const float* bottom_rois,
                                       float* bottom_diff, cudaStream_t stream) {
    if (num_rois < 0 || batch_size < 0) return 0;
    size_t have = 0;
    // accumulates into the caller-zeroed bottom_diff like roi_align_kernel.cu:129-141: AUTO takes the phased kernel in
    // its adding mode where the shape allows it (7x7 lattice, C % 16 == 0, 16 planes in shared memory), else the atomic
    // gather kernel
    void* ws = legacy_scratch(i2v_roi_align_workspace_bytes(batch_size, num_rois), &have, stream);
    if (!ws) return 0;
    int rc = roi_align_backward_impl(top_diff, nullptr, bottom_rois, bottom_diff, batch_size, channels, height, width,
                                     num_rois, aligned_height, aligned_width, spatial_scale, I2V_POOL_NONE,
                                     I2V_IMPL_AUTO, ws, have, false, stream);
    return rc == I2V_OK ? 1 : 0;
}
