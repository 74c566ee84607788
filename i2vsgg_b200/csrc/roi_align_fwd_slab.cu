// RoIAlign / RoIAlignAvg forward, slab variant (roi_align_kernel.cu:15-70 + the module's pool, modules/roi_align.py:18-29).
//
// One CTA per (frame, 16 channels, RoI slice).  The 16 feature planes of the CTA are CONTIGUOUS in the NCHW tensor, so they
// are brought into shared memory by ONE TMA bulk copy (150 KB for a 38x63 map) and used in place, as [channel][row][col]:
// the fill that costs the cell-major plane kernel of roi_align.cu a fifth of its run time (38 K 4-byte cp.async per CTA)
// takes one instruction here.
//
// Lanes are (RoI slot) x (channel) as in the plane kernel: a warp works on two RoIs at a time, one per half-warp, and a
// lane reads the four cells of a lattice point of its channel (one address add, three immediate offsets).  With H*W = 2
// (mod 4) the word address c*H*W + cell of the 16 channels of a half-warp falls on 16 distinct banks of one parity, so a
// half-warp never conflicts with itself; the two half-warps (different RoIs, unrelated cells) collide when their cells
// have the same parity, i.e. half of the time, which costs this kernel 1.5 wavefronts per load where the cell-major
// layout pays one -- a good trade against its fill (38x63 = 2394 and its portrait twin qualify; other shapes and the
// max pool stay on the plane kernel).
#include "common.cuh"

namespace i2v {

struct alignas(16) SlabTab {
    int xs[8];             // start column of lattice column pw (0 when it is off the map)
    float w0[8], w1[8];    // weights of the left / right cell (validity and the avg pool's 1/4 folded in)
    int yoff[8];           // start row * W (words)
    float wy0[8], wy1[8];  // weights of the upper / lower row (validity folded in)
};
static_assert(sizeof(SlabTab) == 192 && sizeof(SlabTab) <= kRoiTabSlotBytes, "SlabTab layout");

namespace {

constexpr int kK = 16;
constexpr int kWarps = 11;          // each works on two RoIs at a time (one per half-warp); 11 is what fits beside the planes
constexpr int kThreads = kWarps * 32;
constexpr int kTabV = (int)(sizeof(SlabTab) / 16);      // 12 16-byte pieces

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
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}

__global__ void __launch_bounds__(128) slab_prep_kernel(const LatticeRoi* __restrict__ tab, SlabTab* __restrict__ stab,
                                                        int num_rois, int G, int W, float wscale) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= num_rois) return;
    const LatticeRoi& t = tab[n];
    SlabTab q;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const bool okx = p < G && ((t.valid_x >> p) & 1u), oky = p < G && ((t.valid_y >> p) & 1u);
        q.xs[p] = okx ? t.x.start[p] : 0;
        q.w0[p] = okx ? (1.f - t.x.frac[p]) * wscale : 0.f;
        q.w1[p] = okx ? t.x.frac[p] * wscale : 0.f;
        q.yoff[p] = oky ? t.y.start[p] * W : 0;
        q.wy0[p] = oky ? 1.f - t.y.frac[p] : 0.f;
        q.wy1[p] = oky ? t.y.frac[p] : 0.f;
    }
    stab[n] = q;
}

template <int POOL, int WT>
__global__ void __launch_bounds__(kThreads, 1)
    lattice_fwd_slab_kernel(const float* __restrict__ feat, const SlabTab* __restrict__ stab, const int* __restrict__ order,
                            const int* __restrict__ starts, float* __restrict__ out, int batch, int C, int H, int Wrt,
                            int split) {
    constexpr int P = 7;
    constexpr int G = (POOL == I2V_POOL_NONE) ? P : P + 1;
    constexpr int NOUT = P * P;
    constexpr int TILE = kK * NOUT;                        // floats per staged output tile
    extern __shared__ __align__(128) float smem[];
    const int W = WT ? WT : Wrt;
    const int HW = H * W;
    float* slab = smem;                                    // [16][H*W]
    float* stage = slab + (size_t)kK * HW;                 // [warps][2][TILE]
    SlabTab* tabs = reinterpret_cast<SlabTab*>(stage + (size_t)kWarps * 2 * TILE);   // [warps][2 halves][2 buffers]
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

    // ---- fill: one bulk copy of the CTA's 16 contiguous planes ----
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar, (unsigned)(kK * HW * sizeof(float)));
        bulk_load(slab, feat + ((size_t)b * C + (size_t)ct * kK) * HW, (unsigned)(kK * HW * sizeof(float)), bar);
    }
    __syncthreads();                                       // the barrier is initialised for everyone

    const float* base = slab + (size_t)c * HW;
    float* my_stage = stage + ((size_t)warp * 2 + role) * TILE;
    SlabTab* my_tabs = tabs + (warp * 2 + role) * 2;

    int li = list_lo + ghalf;
    int n_cur = 0, n_next = 0;
    auto fetch_table = [&](int n, int bufi) {              // 12 lanes of the half copy the 12 x 16 bytes of one table
        if (c < kTabV)
            cp_async16(reinterpret_cast<char*>(my_tabs + bufi) + c * 16, reinterpret_cast<const char*>(stab + n) + c * 16);
    };
    if (li < list_hi) {
        n_cur = __ldg(order + li);
        fetch_table(n_cur, 0);
        if (li + gstride < list_hi) n_next = __ldg(order + li + gstride);
    } else {
        // this half never has work: a zero table keeps its lanes harmless
        for (int i = c; i < (int)(2 * sizeof(SlabTab) / 4); i += 16) reinterpret_cast<float*>(my_tabs)[i] = 0.f;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    mbar_wait(bar, 0);                                     // the planes have landed

    int it = 0;
    while (__any_sync(0xffffffffu, li < list_hi)) {
        const bool active = li < list_hi;
        const int lnext = li + gstride;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        const SlabTab* t = my_tabs + (it & 1);
        int n_next2 = 0;
        if (active) {
            if (lnext < list_hi) {                         // the next RoI's table, and the index after that
                fetch_table(n_next, (it + 1) & 1);
                if (lnext + gstride < list_hi) n_next2 = __ldg(order + lnext + gstride);
            } else {                                       // the half runs dry after this RoI
                float* d = reinterpret_cast<float*>(my_tabs + ((it + 1) & 1));
                for (int i = c; i < (int)(sizeof(SlabTab) / 4); i += 16) d[i] = 0.f;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");

        int xs[8];
        float w0[8], w1[8];
        {
            const int4 x0 = *reinterpret_cast<const int4*>(t->xs), x1 = *reinterpret_cast<const int4*>(t->xs + 4);
            const float4 a0 = *reinterpret_cast<const float4*>(t->w0), a1 = *reinterpret_cast<const float4*>(t->w0 + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(t->w1), b1 = *reinterpret_cast<const float4*>(t->w1 + 4);
            xs[0] = x0.x; xs[1] = x0.y; xs[2] = x0.z; xs[3] = x0.w; xs[4] = x1.x; xs[5] = x1.y; xs[6] = x1.z; xs[7] = x1.w;
            w0[0] = a0.x; w0[1] = a0.y; w0[2] = a0.z; w0[3] = a0.w; w0[4] = a1.x; w0[5] = a1.y; w0[6] = a1.z; w0[7] = a1.w;
            w1[0] = b0.x; w1[1] = b0.y; w1[2] = b0.z; w1[3] = b0.w; w1[4] = b1.x; w1[5] = b1.y; w1[6] = b1.z; w1[7] = b1.w;
        }
        float part[NOUT];
        float prev[G];
#pragma unroll
        for (int ph = 0; ph < G; ++ph) {
            const float* row = base + t->yoff[ph];
            const float wy0 = t->wy0[ph], wy1 = t->wy1[ph];
            float curv[G];
#pragma unroll
            for (int pw = 0; pw < G; ++pw) {
                const float* p = row + xs[pw];
                curv[pw] = (p[0] * wy0 + p[W] * wy1) * w0[pw] + (p[1] * wy0 + p[W + 1] * wy1) * w1[pw];
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

size_t slab_smem_bytes(int H, int W) {
    return ((size_t)kK * H * W + (size_t)kWarps * 2 * kK * 49) * sizeof(float) + (size_t)kWarps * 4 * sizeof(SlabTab) + 16;
}

}  // namespace

bool fwd_slab_ok(const float* features, const float* out, int batch, int C, int H, int W, int PH, int PW, int pool_mode) {
    const long long hw = (long long)H * W;
    return batch > 0 && PH == 7 && PW == 7 && pool_mode != I2V_POOL_MAX && C % kK == 0 && H >= 2 && W >= 2 &&
           hw % 4 == 2 &&                                   // the bank argument in the header
           kK * hw * 4 < (1 << 20) &&                       // one mbarrier transaction
           slab_smem_bytes(H, W) <= (size_t)kMaxSmemPerCta && ((uintptr_t)features & 15) == 0 && ((uintptr_t)out & 15) == 0;
}

template <int POOL, int WT>
static int launch_slab(const float* feat, const SlabTab* stab, const int* order, const int* starts, float* out, int batch,
                       int C, int H, int W, cudaStream_t stream) {
    auto kern = lattice_fwd_slab_kernel<POOL, WT>;
    const size_t smem = slab_smem_bytes(H, W);
    I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ctiles = C / kK;
    int split = 1;
    while (batch * ctiles * split < 2 * kNumSMs && split < 8) split *= 2;
    kern<<<(batch + 1) * ctiles * split, kThreads, smem, stream>>>(feat, stab, order, starts, out, batch, C, H, W, split);
    return check_launch("lattice_fwd_slab_kernel");
}

// `tab` holds the LatticeRoi tables of this call; `tab_space` is the workspace's per-RoI table slot (kRoiTabSlotBytes each).
int launch_fwd_slab(const float* feat, const LatticeRoi* tab, void* tab_space, const int* order, const int* starts, float* out,
                    int batch, int C, int H, int W, int num_rois, int pool_mode, cudaStream_t stream) {
    SlabTab* stab = static_cast<SlabTab*>(tab_space);
    const int G = pool_mode == I2V_POOL_NONE ? 7 : 8;
    slab_prep_kernel<<<ceil_div(num_rois, 128), 128, 0, stream>>>(tab, stab, num_rois, G, W,
                                                                   pool_mode == I2V_POOL_AVG ? 0.25f : 1.f);
    I2V_TRY(check_launch("slab_prep_kernel"));
    if (pool_mode == I2V_POOL_AVG) {
        if (W == 63) return launch_slab<I2V_POOL_AVG, 63>(feat, stab, order, starts, out, batch, C, H, W, stream);
        return launch_slab<I2V_POOL_AVG, 0>(feat, stab, order, starts, out, batch, C, H, W, stream);
    }
    return launch_slab<I2V_POOL_NONE, 0>(feat, stab, order, starts, out, batch, C, H, W, stream);
}

}  // namespace i2v
