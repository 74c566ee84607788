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
                                    int* __restrict__ counts) {
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
    tab[n] = t;
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
    int base = s_base;
    for (int i0 = 0; i0 < num_rois; i0 += 256) {
        int i = i0 + tid;
        bool mine = (i < num_rois) && (roi_batch[i] == b);
        unsigned m = __ballot_sync(0xffffffffu, mine);
        __syncthreads();  // s_warp reuse
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            int c = s_warp[w];
            if (w < warp) before += c;
            total += c;
        }
        if (mine) order[base + before + __popc(m & ((1u << lane) - 1u))] = i;
        base += total;
    }
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
constexpr int kPlaneWarps = 8;     // warps per CTA
constexpr int kPlaneThreads = kPlaneWarps * 32;
constexpr int kFillCells = 768;    // cells per fill round
constexpr int kFillPitch = 770;    // == 2 (mod 32): the transposing read of a round is bank-conflict free

__device__ __forceinline__ void bulk_store_commit(float* gdst, const float* ssrc, unsigned bytes) {
    unsigned saddr = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// The table of one RoI as it sits in registers while the previous RoI is being computed (G <= 8).
struct RawLattice {
    int4 ys[2], yf[2], xs[2], xf[2], tail;
    int n;
};
__device__ __forceinline__ void load_raw(RawLattice& r, const LatticeRoi* __restrict__ tab,
                                         const int* __restrict__ order, int li) {
    r.n = __ldg(order + li);
    const int4* q = reinterpret_cast<const int4*>(tab + r.n);
    r.ys[0] = __ldg(q + 0);  r.ys[1] = __ldg(q + 1);
    r.yf[0] = __ldg(q + 4);  r.yf[1] = __ldg(q + 5);
    r.xs[0] = __ldg(q + 8);  r.xs[1] = __ldg(q + 9);
    r.xf[0] = __ldg(q + 12); r.xf[1] = __ldg(q + 13);
    r.tail = __ldg(q + 16);
}
__device__ __forceinline__ int pick(const int4 (&a)[2], int p) {  // p is a compile-time constant after unrolling
    const int4& v = a[p >> 2];
    return (p & 3) == 0 ? v.x : (p & 3) == 1 ? v.y : (p & 3) == 2 ? v.z : v.w;
}

// P = pooled size (output is P x P); lattice G = P (+1 when a pool follows); WT = compile-time map width (0: runtime).
// A warp works on one RoI x 16 channels at a time: lanes 0-15 read the left cell of every bilinear pair, lanes
// 16-31 the right one (adjacent 64-byte rows of the [cell][16] layout: 32 distinct banks), each half carries the
// partial sums of its column, the halves are combined with one shuffle per output and the finished
// [16][P*P] tile -- which is contiguous in the NCHW output -- leaves through a TMA bulk store.
template <int P, int POOL, int WT>
__global__ void __launch_bounds__(kPlaneThreads, 1)
    lattice_fwd_plane_kernel(const float* __restrict__ feat, const LatticeRoi* __restrict__ tab,
                             const int* __restrict__ order, const int* __restrict__ starts, float* __restrict__ out,
                             int batch, int C, int H, int Wrt, int split) {
    constexpr int G = (POOL == I2V_POOL_NONE) ? P : P + 1;
    constexpr int NOUT = P * P;
    constexpr int TILE = kPlaneK * NOUT;  // floats per staged output tile
    static_assert(G <= 8, "register tables hold 8 lattice points per axis");
    extern __shared__ __align__(128) float smem[];
    const int W = WT ? WT : Wrt;
    const int HW = H * W;
    float* planes = smem;                        // [HW][16]
    float* stage = smem + (size_t)HW * kPlaneK;  // [warps][2][TILE]; doubles as the fill scratch

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ctiles = C / kPlaneK;
    const int s = blockIdx.x % split;
    const int ct = (blockIdx.x / split) % ctiles;
    const int b = blockIdx.x / (split * ctiles);
    const int list_lo = __ldg(starts + b), list_hi = __ldg(starts + b + 1);
    if (list_lo == list_hi) return;
    const int gwarp = s * kPlaneWarps + warp, gstride = split * kPlaneWarps;

    if (b == batch) {  // RoIs with an out-of-range batch index: zero rows
        for (int li = list_lo + gwarp; li < list_hi; li += gstride) {
            float4* dst = reinterpret_cast<float4*>(out + ((size_t)__ldg(order + li) * C + (size_t)ct * kPlaneK) * NOUT);
            for (int i = lane; i < TILE / 4; i += 32) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }

    // ---- fill: 16 planes, global [c][cell] -> shared [cell][16], through a [16][770] scratch ----
    {
        const float* src = feat + ((size_t)b * C + (size_t)ct * kPlaneK) * HW;
        for (int c0 = 0; c0 < HW; c0 += kFillCells) {
            const int ncell = min(kFillCells, HW - c0);
            for (int i = tid; i < kPlaneK * kFillCells; i += kPlaneThreads) {
                int c = i / kFillCells, x = i - c * kFillCells;
                if (x < ncell) stage[c * kFillPitch + x] = __ldg(src + (size_t)c * HW + c0 + x);
            }
            __syncthreads();
            for (int i = tid; i < kPlaneK * kFillCells; i += kPlaneThreads) {
                int c = i & (kPlaneK - 1), x = i >> 4;
                if (x < ncell) planes[(size_t)(c0 + x) * kPlaneK + c] = stage[c * kFillPitch + x];
            }
            __syncthreads();
        }
    }

    const int c = lane & 15, dx = lane >> 4;
    const char* pl0 = reinterpret_cast<const char*>(planes + c + dx * kPlaneK);  // row hs, this lane's column
    const int row_bytes = W * kPlaneK * (int)sizeof(float);
    float* my_stage = stage + (size_t)warp * 2 * TILE;
    int buf = 0;

    int li = list_lo + gwarp;
    RawLattice cur;
    if (li < list_hi) load_raw(cur, tab, order, li);
    while (li < list_hi) {
        RawLattice nxt;
        const int lnext = li + gstride;
        if (lnext < list_hi) load_raw(nxt, tab, order, lnext);

        // per-lane tables: column byte offset / x weight per lattice column, row byte offset / y weights per row
        int xoff[G], yoff[G];
        float wxl[G], wy0[G], wy1[G];
        {
            const unsigned vy = (unsigned)cur.tail.y, vx = (unsigned)cur.tail.z;
#pragma unroll
            for (int p = 0; p < G; ++p) {
                float xf = __int_as_float(pick(cur.xf, p)), yf = __int_as_float(pick(cur.yf, p));
                bool okx = (vx >> p) & 1u, oky = (vy >> p) & 1u;
                xoff[p] = pick(cur.xs, p) * (kPlaneK * (int)sizeof(float));
                float wx = dx ? xf : 1.f - xf;
                if (POOL == I2V_POOL_AVG) wx *= 0.25f;  // the pool's divide, folded into the column weight
                wxl[p] = okx ? wx : 0.f;
                yoff[p] = pick(cur.ys, p) * row_bytes;
                wy0[p] = oky ? 1.f - yf : 0.f;
                wy1[p] = oky ? yf : 0.f;
            }
        }
        float part[NOUT];  // AVG: this half-warp's partial sums; MAX / NONE: the finished values
        float prev[G];
#pragma unroll
        for (int ph = 0; ph < G; ++ph) {
            float curv[G];
#pragma unroll
            for (int pw = 0; pw < G; ++pw) {
                const char* a = pl0 + (yoff[ph] + xoff[pw]);
                float f0 = *reinterpret_cast<const float*>(a);
                float f1 = *reinterpret_cast<const float*>(a + (WT ? WT * kPlaneK * 4 : row_bytes));
                curv[pw] = (f0 * wy0[ph] + f1 * wy1[ph]) * wxl[pw];
            }
            if (POOL != I2V_POOL_AVG) {
                // max is not linear and NONE writes lattice values: combine the two columns now
#pragma unroll
                for (int pw = 0; pw < G; ++pw) curv[pw] += __shfl_xor_sync(0xffffffffu, curv[pw], 16);
            }
            if (POOL == I2V_POOL_NONE) {
#pragma unroll
                for (int pw = 0; pw < G; ++pw) part[ph * P + pw] = curv[pw];
            } else if (POOL == I2V_POOL_AVG) {
#pragma unroll
                for (int pw = 0; pw < P; ++pw) curv[pw] += curv[pw + 1];  // row sums of adjacent columns
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

        // ---- stage the [16][NOUT] tile and hand it to the TMA ----
        float* st = my_stage + buf * TILE;
        if (lane == 0) bulk_wait_read_1();  // the store that last read this buffer (two RoIs ago) is done
        __syncwarp();
        float* row = st + c * NOUT;
        // outputs [0,16) are written by the low half-warp while the high one writes [16,32): 16 floats apart, so
        // the two halves use disjoint banks; the rest is written by the low half alone.
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (k + 16 < NOUT) {
                float mine = dx ? part[k + 16] : part[k];
                if (POOL == I2V_POOL_AVG) {
                    float send = dx ? part[k] : part[k + 16];
                    mine += __shfl_xor_sync(0xffffffffu, send, 16);
                }
                row[k + dx * 16] = mine;
            } else if (k < NOUT) {
                float mine = part[k];
                if (POOL == I2V_POOL_AVG) mine += __shfl_xor_sync(0xffffffffu, mine, 16);
                if (!dx) row[k] = mine;
            }
        }
#pragma unroll
        for (int k = 32; k < NOUT; ++k) {
            float mine = part[k];
            if (POOL == I2V_POOL_AVG) mine += __shfl_xor_sync(0xffffffffu, mine, 16);
            if (!dx) row[k] = mine;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0)
            bulk_store_commit(out + ((size_t)cur.n * C + (size_t)ct * kPlaneK) * NOUT, st, TILE * sizeof(float));
        buf ^= 1;
        cur = nxt;
        li = lnext;
    }
    if (lane == 0) bulk_wait_all();
}

// ------------------------------------------------------------------------------------------ host side
struct LatticeWs {
    LatticeRoi* tab;
    int* roi_batch;
    int* counts;
    int* starts;
    int* order;
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
    }
    w.bytes = cv.used();
    return w;
}

static int lattice_prep(const float* rois, int batch, int num_rois, int H, int W, int GH, int GW, float scale,
                        const LatticeWs& w, bool lists, cudaStream_t stream) {
    if (lists) I2V_CUDA_TRY(cudaMemsetAsync(w.counts, 0, sizeof(int) * ((size_t)batch + 2), stream));
    lattice_prep_kernel<<<ceil_div(num_rois, 128), 128, 0, stream>>>(rois, num_rois, batch, H, W, GH, GW, scale, w.tab,
                                                                     lists ? w.roi_batch : nullptr,
                                                                     lists ? w.counts : nullptr);
    I2V_TRY(check_launch("lattice_prep_kernel"));
    if (lists) {
        roi_bucket_kernel<<<batch + 1, 256, 0, stream>>>(w.roi_batch, num_rois, batch + 1, w.counts, w.starts, w.order);
        I2V_TRY(check_launch("roi_bucket_kernel"));
    }
    return I2V_OK;
}

static size_t plane_fwd_smem_bytes(int H, int W, int P) {
    size_t stage = (size_t)kPlaneWarps * 2 * kPlaneK * P * P;
    size_t scratch = (size_t)kPlaneK * kFillPitch;
    return ((size_t)H * W * kPlaneK + (stage > scratch ? stage : scratch)) * sizeof(float);
}

static bool plane_forward_ok(const float* out, int batch, int C, int H, int W, int PH, int PW) {
    return batch > 0 && PH == 7 && PW == 7 && C % kPlaneK == 0 && H >= 2 && W >= 2 &&
           plane_fwd_smem_bytes(H, W, 7) <= (size_t)kMaxSmemPerCta && ((uintptr_t)out & 15) == 0;
}

static int plane_split(int ctas) {
    int split = 1;
    while (ctas * split < 2 * kNumSMs && split < 8) split *= 2;
    return split;
}

template <int POOL, int WT>
static int launch_plane_fwd_w(const float* feat, const LatticeWs& w, float* out, int batch, int C, int H, int W,
                              cudaStream_t stream) {
    auto kern = lattice_fwd_plane_kernel<7, POOL, WT>;
    size_t smem = plane_fwd_smem_bytes(H, W, 7);
    I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ctiles = C / kPlaneK;
    int split = plane_split(batch * ctiles);
    kern<<<(batch + 1) * ctiles * split, kPlaneThreads, smem, stream>>>(feat, w.tab, w.order, w.starts, out, batch, C,
                                                                       H, W, split);
    return check_launch("lattice_fwd_plane_kernel");
}
template <int POOL>
static int launch_plane_fwd(const float* feat, const LatticeWs& w, float* out, int batch, int C, int H, int W,
                            cudaStream_t stream) {
    if (W == 63) return launch_plane_fwd_w<POOL, 63>(feat, w, out, batch, C, H, W, stream);
    return launch_plane_fwd_w<POOL, 0>(feat, w, out, batch, C, H, W, stream);
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
    I2V_REQUIRE(impl >= I2V_IMPL_AUTO && impl <= I2V_IMPL_PLANE, "%s: bad impl %d", who, impl);
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
    if (impl == I2V_IMPL_PLANE && !can_plane) {
        set_error("roi_align_forward: the plane kernel needs a 7x7 output, C %% 16 == 0, a 16-byte aligned output and "
                  "16 planes that fit shared memory");
        return I2V_ERR_UNSUPPORTED;
    }
    bool plane = can_plane && impl != I2V_IMPL_GATHER;
    I2V_TRY(carve_checked("roi_align_forward", workspace, workspace_bytes, batch, num_rois, plane, w));
    I2V_TRY(lattice_prep(rois, batch, num_rois, height, width, GH, GW, spatial_scale, w, plane, stream));
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
    I2V_TRY(roi_align_check("roi_align_backward", grad_out, rois, grad_in, batch, channels, height, width, num_rois,
                            pooled_h, pooled_w, pool_mode, impl, GH, GW));
    I2V_REQUIRE(pool_mode != I2V_POOL_MAX || features || num_rois == 0, "roi_align_backward: POOL_MAX needs the features");
    size_t in_elems = (size_t)batch * channels * height * width;
    if (in_elems == 0) return I2V_OK;
    I2V_REQUIRE(grad_in, "roi_align_backward: null grad_in");
    if (impl == I2V_IMPL_PLANE) {
        set_error("roi_align_backward: no plane kernel for this shape");
        return I2V_ERR_UNSUPPORTED;
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
extern "C" int ROIAlignForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois,
                                      const int height, const int width, const int channels, const int aligned_height,
                                      const int aligned_width, const float* bottom_rois, float* top_data,
                                      cudaStream_t stream) {
    if (num_rois < 0) return 0;
    size_t have = 0;
    void* ws = legacy_scratch(carve_lattice_ws(nullptr, 0, num_rois, false).bytes, &have);
    if (!ws) return 0;
    // roi_align_kernel.h:13-17 does not pass the batch size: every non-negative frame index is accepted
    int rc = i2v_roi_align_forward(bottom_data, bottom_rois, top_data, INT32_MAX, channels, height, width, num_rois,
                                   aligned_height, aligned_width, spatial_scale, I2V_POOL_NONE, I2V_IMPL_GATHER, ws, have,
                                   stream);
    return rc == I2V_OK ? 1 : 0;
}

extern "C" int ROIAlignBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size,
                                       const int num_rois, const int height, const int width, const int channels,
                                       const int aligned_height, const int aligned_width, const float* bottom_rois,
                                       float* bottom_diff, cudaStream_t stream) {
    if (num_rois < 0) return 0;
    size_t have = 0;
    void* ws = legacy_scratch(carve_lattice_ws(nullptr, 0, num_rois, false).bytes, &have);
    if (!ws) return 0;
    // accumulates into the caller-zeroed bottom_diff like roi_align_kernel.cu:129-141
    int rc = roi_align_backward_impl(top_diff, nullptr, bottom_rois, bottom_diff, batch_size, channels, height, width,
                                     num_rois, aligned_height, aligned_width, spatial_scale, I2V_POOL_NONE,
                                     I2V_IMPL_GATHER, ws, have, false, stream);
    return rc == I2V_OK ? 1 : 0;
}
