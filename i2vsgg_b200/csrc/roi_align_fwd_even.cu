// RoIAlign / RoIAlignAvg forward, even-pitch slab variant (roi_align_kernel.cu:15-70 + the module's pool,
// modules/roi_align.py:18-29).
//
// The slab kernel (roi_align_fwd_slab.cu) uses the 16 planes of a CTA where one TMA bulk copy drops them, as
// [channel][row][63 columns]: a lane's bank is decided by (row + column) of its cell, the two half-warps (two different
// RoIs) collide whenever those parities agree, and that costs 57 M of its 161 M shared-memory wavefronts at 88 % L1/TEX
// utilisation.  Here the planes are re-pitched in place to an EVEN row pitch (64 cells; 40 for portrait maps) with a plane
// stride of 2 (mod 32) words:
//
//   bank(channel c, row y, column x) = (2 c + pitch * y + x) mod 32  ->  its parity is the parity of x alone,
//
// the 16 channels of a half-warp still fall on 16 distinct banks of that parity, and a bilinear pair (x, x + 1) always has
// one even and one odd column.  The per-RoI table lists each pair twice -- even cell first for half-warp 0, odd cell first
// for half-warp 1 -- so in every load instruction the two halves are on opposite parities: conflict free by construction,
// with per-AXIS tables only (the 63-pitch layout needs row-and-column parity, i.e. selects per lattice point).
//
// Fill: one TMA bulk copy of the 16 contiguous planes into the front of the buffer (as in the slab kernel), then every
// warp reads its share of the dense rows into registers, a barrier, and the rows are written back at the even pitch
// (the destination of a row lies behind its source, so the move cannot be done in place without the register stage).
#include "common.cuh"

namespace i2v {

struct alignas(16) EvenTab {
    // per half-warp h and lattice column: word offset of the cell read first / second inside a plane row and their weights
    // (validity and the avg pool's 1/4 folded in).  h = 0 reads the even column first, h = 1 the odd one.
    int xf[2][8];
    float wf[2][8];
    int xs[2][8];
    float ws[2][8];
    int yoff[8];            // start row * pitch (words)
    float wy0[8], wy1[8];   // weights of the upper / lower row (validity folded in)
};
static_assert(sizeof(EvenTab) == 352 && sizeof(EvenTab) <= kRoiTabSlotBytes, "EvenTab layout");

namespace {

constexpr int kK = 16;
constexpr int kWarps = 10;          // each works on two RoIs at a time (one per half-warp)
constexpr int kThreads = kWarps * 32;
constexpr int kHalfTabBytes = 128 + 96;                  // one half's x arrays + the y arrays
constexpr int kHalfTabV = kHalfTabBytes / 16;            // 14 16-byte pieces

__host__ __device__ constexpr int even_pitch_for(int W) { return W <= 40 ? 40 : 64; }
// words between planes: H * pitch rounded up to 2 (mod 32)
__host__ __device__ inline int even_plane_stride(int H, int W) {
    const int n = H * even_pitch_for(W);
    return n + ((34 - (n & 31)) & 31);
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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
__device__ __forceinline__ void bulk_store_commit(float* gdst, const float* ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// plain (non-volatile) shared loads through 32-bit addresses: one IADD per address, the lower row as an immediate
template <int OFF>
__device__ __forceinline__ float lds32(unsigned addr) {
    float v;
    asm("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
    return v;
}
// {d0, d1} = {a0, a1} * {b, b} + {c0, c1} and {d0, d1} = {a0, a1} * {b, b}: packed fp32x2
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b, float c0, float c1) {
    asm("{.reg .b64 ra, rb, rc, rd;\n"
        "mov.b64 ra, {%2, %3};\n"
        "mov.b64 rb, {%4, %4};\n"
        "mov.b64 rc, {%5, %6};\n"
        "fma.rn.f32x2 rd, ra, rb, rc;\n"
        "mov.b64 {%0, %1}, rd;}"
        : "=f"(d0), "=f"(d1)
        : "f"(a0), "f"(a1), "f"(b), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void mul2(float& d0, float& d1, float a0, float a1, float b) {
    asm("{.reg .b64 ra, rb, rd;\n"
        "mov.b64 ra, {%2, %3};\n"
        "mov.b64 rb, {%4, %4};\n"
        "mul.rn.f32x2 rd, ra, rb;\n"
        "mov.b64 {%0, %1}, rd;}"
        : "=f"(d0), "=f"(d1)
        : "f"(a0), "f"(a1), "f"(b));
}
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}

__global__ void __launch_bounds__(128) even_prep_kernel(const LatticeRoi* __restrict__ tab, unsigned char* __restrict__ tab_space,
                                                        int num_rois, int G, int pitch, float wscale) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= num_rois) return;
    const LatticeRoi& t = tab[n];
    EvenTab q;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const bool okx = p < G && ((t.valid_x >> p) & 1u), oky = p < G && ((t.valid_y >> p) & 1u);
        const int x = okx ? t.x.start[p] : 0;
        const float wl = okx ? (1.f - t.x.frac[p]) * wscale : 0.f, wr = okx ? t.x.frac[p] * wscale : 0.f;
        const bool left_even = (x & 1) == 0;
        const int xe = left_even ? x : x + 1, xo = left_even ? x + 1 : x;
        const float we = left_even ? wl : wr, wo = left_even ? wr : wl;
        q.xf[0][p] = xe; q.wf[0][p] = we; q.xs[0][p] = xo; q.ws[0][p] = wo;
        q.xf[1][p] = xo; q.wf[1][p] = wo; q.xs[1][p] = xe; q.ws[1][p] = we;
        q.yoff[p] = oky ? t.y.start[p] * pitch : 0;
        q.wy0[p] = oky ? 1.f - t.y.frac[p] : 0.f;
        q.wy1[p] = oky ? t.y.frac[p] : 0.f;
    }
    *reinterpret_cast<EvenTab*>(tab_space + (size_t)n * kRoiTabSlotBytes) = q;
}

// what one half-warp keeps of a table in shared memory: its x arrays and the y arrays, contiguous
struct alignas(16) HalfTab {
    int xf[8];
    float wf[8];
    int xs[8];
    float ws[8];
    int yoff[8];
    float wy0[8], wy1[8];
};
static_assert(sizeof(HalfTab) == kHalfTabBytes, "HalfTab layout");

template <int POOL, int PITCH>
__global__ void __launch_bounds__(kThreads, 1)
    lattice_fwd_even_kernel(const float* __restrict__ feat, const unsigned char* __restrict__ tab_space,
                            const int* __restrict__ order, const int* __restrict__ starts, float* __restrict__ out, int batch,
                            int C, int H, int W, int PS, int split) {
    constexpr int P = 7;
    constexpr int G = (POOL == I2V_POOL_NONE) ? P : P + 1;
    constexpr int NOUT = P * P;
    constexpr int TILE = kK * NOUT;                        // floats per staged output tile
    constexpr int kMaxRowsPerWarp = 64;                    // dense rows a warp re-pitches (16 * H / kWarps, H <= 40)
    extern __shared__ __align__(128) float smem[];
    const int HW = H * W;
    float* slab = smem;                                    // [16][PS], row pitch PITCH
    float* stage = slab + (size_t)kK * PS;                 // [warps][2][TILE]
    HalfTab* tabs = reinterpret_cast<HalfTab*>(stage + (size_t)kWarps * 2 * TILE);   // [warps][2 halves][2 buffers]
    uint64_t* bar = reinterpret_cast<uint64_t*>(tabs + kWarps * 4);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ctiles = C / kK;
    const int s = blockIdx.x % split;
    const int ct = (blockIdx.x / split) % ctiles;
    const int b = blockIdx.x / (split * ctiles);
    const int list_lo = __ldg(starts + b), list_hi = __ldg(starts + b + 1);
    if (list_lo == list_hi) return;
    const int role = lane >> 4, c = lane & 15;
    const int ghalf = (s * kWarps + warp) * 2 + role, gstride = split * kWarps * 2;

    if (b == batch) {  // RoIs with an out-of-range batch index: zero rows
        for (int li = list_lo + ghalf; li < list_hi; li += gstride) {
            float4* dst = reinterpret_cast<float4*>(out + ((size_t)__ldg(order + li) * C + (size_t)ct * kK) * NOUT);
            for (int i = c; i < TILE / 4; i += 16) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }

    // ---- fill: one bulk copy of the CTA's 16 contiguous planes into the front of the slab ----
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar, (unsigned)(kK * HW * sizeof(float)));
        bulk_load(slab, feat + ((size_t)b * C + (size_t)ct * kK) * HW, (unsigned)(kK * HW * sizeof(float)), bar);
    }
    __syncthreads();                                       // the barrier is initialised for everyone

    float* my_stage = stage + ((size_t)warp * 2 + role) * TILE;
    HalfTab* my_tabs = tabs + (warp * 2 + role) * 2;

    int li = list_lo + ghalf;
    int n_cur = 0, n_next = 0;
    auto fetch_table = [&](int n, int bufi) {              // 14 lanes of the half copy its 8 + 6 sixteen-byte pieces
        if (c < kHalfTabV) {
            const unsigned char* src = tab_space + (size_t)n * kRoiTabSlotBytes;
            // pieces 0-7: the half's x arrays (xf, wf, xs, ws are [2][8]: 32 bytes per half inside each 64-byte array)
            const int piece = c;
            const unsigned char* g = piece < 8 ? src + (piece >> 1) * 64 + role * 32 + (piece & 1) * 16
                                               : src + 256 + (piece - 8) * 16;
            cp_async16(reinterpret_cast<char*>(my_tabs + bufi) + piece * 16, g);
        }
    };
    if (li < list_hi) {
        n_cur = __ldg(order + li);
        fetch_table(n_cur, 0);
        if (li + gstride < list_hi) n_next = __ldg(order + li + gstride);
    } else {
        for (int i = c; i < (int)(2 * sizeof(HalfTab) / 4); i += 16) reinterpret_cast<float*>(my_tabs)[i] = 0.f;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    mbar_wait(bar, 0);                                     // the dense planes have landed

    // ---- re-pitch: dense row r = c * H + y (W words at r * W) -> c * PS + y * PITCH.  A warp takes the rows r = warp,
    // warp + kWarps, ...; lanes the columns lane and lane + 32.  Everything is read before anything is written. ----
    {
        const int rows = kK * H;
        float va[kMaxRowsPerWarp], vb[kMaxRowsPerWarp];
        const bool hi_col = lane + 32 < W;
#pragma unroll
        for (int k = 0; k < kMaxRowsPerWarp; ++k) {
            const int r = warp + k * kWarps;
            if (r < rows) {
                va[k] = lane < W ? slab[(size_t)r * W + lane] : 0.f;
                vb[k] = hi_col ? slab[(size_t)r * W + lane + 32] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kMaxRowsPerWarp; ++k) {
            const int r = warp + k * kWarps;
            if (r < rows) {
                const int ch = r / H, y = r - ch * H;
                float* d = slab + (size_t)ch * PS + y * PITCH;
                if (lane < W) d[lane] = va[k];
                if (hi_col) d[lane + 32] = vb[k];
                else if (lane + 32 < PITCH) d[lane + 32] = 0.f;      // the padding columns: finite, never weighted
            }
        }
        __syncthreads();
    }

    const unsigned base = smem_u32(slab) + (unsigned)c * (unsigned)PS * 4u;     // byte address of the lane's plane
    int it = 0;
    while (__any_sync(0xffffffffu, li < list_hi)) {
        const bool active = li < list_hi;
        const int lnext = li + gstride;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        const HalfTab* t = my_tabs + (it & 1);
        int n_next2 = 0;
        if (active) {
            if (lnext < list_hi) {                         // the next RoI's table, and the index after that
                fetch_table(n_next, (it + 1) & 1);
                if (lnext + gstride < list_hi) n_next2 = __ldg(order + lnext + gstride);
            } else {                                       // the half runs dry after this RoI
                float* d = reinterpret_cast<float*>(my_tabs + ((it + 1) & 1));
                for (int i = c; i < (int)(sizeof(HalfTab) / 4); i += 16) d[i] = 0.f;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");

        unsigned pf[8], ps[8];                             // byte offsets of the first / second cell inside a plane row
        float wf[8], ws[8];
        {
            const int4 f0 = *reinterpret_cast<const int4*>(t->xf), f1 = *reinterpret_cast<const int4*>(t->xf + 4);
            const int4 s0 = *reinterpret_cast<const int4*>(t->xs), s1 = *reinterpret_cast<const int4*>(t->xs + 4);
            const float4 a0 = *reinterpret_cast<const float4*>(t->wf), a1 = *reinterpret_cast<const float4*>(t->wf + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(t->ws), b1 = *reinterpret_cast<const float4*>(t->ws + 4);
            pf[0] = f0.x * 4; pf[1] = f0.y * 4; pf[2] = f0.z * 4; pf[3] = f0.w * 4;
            pf[4] = f1.x * 4; pf[5] = f1.y * 4; pf[6] = f1.z * 4; pf[7] = f1.w * 4;
            ps[0] = s0.x * 4; ps[1] = s0.y * 4; ps[2] = s0.z * 4; ps[3] = s0.w * 4;
            ps[4] = s1.x * 4; ps[5] = s1.y * 4; ps[6] = s1.z * 4; ps[7] = s1.w * 4;
            wf[0] = a0.x; wf[1] = a0.y; wf[2] = a0.z; wf[3] = a0.w; wf[4] = a1.x; wf[5] = a1.y; wf[6] = a1.z; wf[7] = a1.w;
            ws[0] = b0.x; ws[1] = b0.y; ws[2] = b0.z; ws[3] = b0.w; ws[4] = b1.x; ws[5] = b1.y; ws[6] = b1.z; ws[7] = b1.w;
        }
        float part[NOUT];
        float prev[G];
#pragma unroll
        for (int ph = 0; ph < G; ++ph) {
            const unsigned row = base + (unsigned)t->yoff[ph] * 4u;
            const float wy0 = t->wy0[ph], wy1 = t->wy1[ph];
            float curv[G];
#pragma unroll
            for (int pw = 0; pw < G; ++pw) {
                // first the cell whose column parity belongs to this half-warp (even for half 0, odd for half 1), then the other
                const unsigned p = row + pf[pw], q = row + ps[pw];
                const float p0 = lds32<0>(p), p1 = lds32<PITCH * 4>(p);
                const float q0 = lds32<0>(q), q1 = lds32<PITCH * 4>(q);
                float u0, u1;
                mul2(u0, u1, p0, q0, wy0);                 // {p0, q0} * wy0
                fma2(u0, u1, p1, q1, wy1, u0, u1);         // + {p1, q1} * wy1
                curv[pw] = fmaf(u1, ws[pw], u0 * wf[pw]);
            }
            if (POOL == I2V_POOL_NONE) {
#pragma unroll
                for (int pw = 0; pw < G; ++pw) part[ph * P + pw] = curv[pw];
            } else {
#pragma unroll
                for (int pw = 0; pw < P; ++pw) curv[pw] += curv[pw + 1];   // adjacent columns (x 1/4 in the weights)
                if (ph > 0) {
#pragma unroll
                    for (int pw = 0; pw < P; ++pw) part[(ph - 1) * P + pw] = prev[pw] + curv[pw];
                }
            }
#pragma unroll
            for (int pw = 0; pw < G; ++pw) prev[pw] = curv[pw];
        }

        // ---- stage the [16][49] tile of this half and hand it to the TMA ----
        if (c == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        float* row = my_stage + c * NOUT;                  // the two halves' tiles are 16 banks apart
#pragma unroll
        for (int k = 0; k < NOUT; ++k) row[k] = part[k];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (c == 0 && active)
            bulk_store_commit(out + ((size_t)n_cur * C + (size_t)ct * kK) * NOUT, my_stage, TILE * sizeof(float));
        n_cur = n_next;
        n_next = n_next2;
        li = lnext;
        ++it;
    }
    if (c == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

size_t even_smem_bytes(int H, int W) {
    return ((size_t)kK * even_plane_stride(H, W) + (size_t)kWarps * 2 * kK * 49) * sizeof(float) +
           (size_t)kWarps * 4 * sizeof(HalfTab) + 16;
}

}  // namespace

bool fwd_even_ok(const float* features, const float* out, int batch, int C, int H, int W, int PH, int PW, int pool_mode) {
    const long long hw = (long long)H * W;
    return batch > 0 && PH == 7 && PW == 7 && pool_mode != I2V_POOL_MAX && C % kK == 0 && H >= 2 && W >= 2 && W <= 64 &&
           kK * H <= 64 * kWarps &&                         // the register stage of the re-pitch
           (kK * hw * 4) % 16 == 0 && kK * hw * 4 < (1 << 20) &&   // one bulk copy, one mbarrier transaction
           even_smem_bytes(H, W) <= (size_t)kMaxSmemPerCta && ((uintptr_t)features & 15) == 0 && ((uintptr_t)out & 15) == 0;
}

template <int POOL, int PITCH>
static int launch_even(const float* feat, const unsigned char* ts, const int* order, const int* starts, float* out, int batch,
                       int C, int H, int W, cudaStream_t stream) {
    auto kern = lattice_fwd_even_kernel<POOL, PITCH>;
    const size_t smem = even_smem_bytes(H, W);
    I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ctiles = C / kK;
    int split = 1;
    while (batch * ctiles * split < 2 * kNumSMs && split < 8) split *= 2;
    kern<<<(batch + 1) * ctiles * split, kThreads, smem, stream>>>(feat, ts, order, starts, out, batch, C, H, W,
                                                                   even_plane_stride(H, W), split);
    return check_launch("lattice_fwd_even_kernel");
}

// `tab` holds the LatticeRoi tables of this call; `tab_space` is the workspace's per-RoI table slot (kRoiTabSlotBytes each).
int launch_fwd_even(const float* feat, const LatticeRoi* tab, void* tab_space, const int* order, const int* starts, float* out,
                    int batch, int C, int H, int W, int num_rois, int pool_mode, cudaStream_t stream) {
    unsigned char* ts = static_cast<unsigned char*>(tab_space);
    const int G = pool_mode == I2V_POOL_NONE ? 7 : 8;
    const int pitch = even_pitch_for(W);
    even_prep_kernel<<<ceil_div(num_rois, 128), 128, 0, stream>>>(tab, ts, num_rois, G, pitch,
                                                                   pool_mode == I2V_POOL_AVG ? 0.25f : 1.f);
    I2V_TRY(check_launch("even_prep_kernel"));
    if (pool_mode == I2V_POOL_AVG) {
        if (pitch == 64) return launch_even<I2V_POOL_AVG, 64>(feat, ts, order, starts, out, batch, C, H, W, stream);
        return launch_even<I2V_POOL_AVG, 40>(feat, ts, order, starts, out, batch, C, H, W, stream);
    }
    if (pitch == 64) return launch_even<I2V_POOL_NONE, 64>(feat, ts, order, starts, out, batch, C, H, W, stream);
    return launch_even<I2V_POOL_NONE, 40>(feat, ts, order, starts, out, batch, C, H, W, stream);
}

}  // namespace i2v
