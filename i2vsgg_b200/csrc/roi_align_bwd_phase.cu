// RoIAlign / RoIAlignAvg backward, phased variant (roi_align_kernel.cu:94-143 behind the pool's backward,
// modules/roi_align.py:18-29).
//
// One CTA per (frame, 16 channels) owns those 16 gradient planes in shared memory as [cell][16 channels] (64 bytes per
// cell) and writes them to HBM once: no global atomics, no memset, deterministic.  The frame's RoIs stream through a TMA
// ring (pooled-gradient tile [16][49] + the RoI's table) fed by a producer warp.
//
//   * lanes are (parity of the cell's column) x (channel): the two cells a lattice point touches in one feature row are
//     128 contiguous bytes, one owned by each half-warp, so every read-modify-write of the planes is one conflict-free
//     wavefront (the warp-per-channel-pair kernel in roi_align.cu pays 2.6 wavefronts per access for its scattered
//     8-byte cells).  Because a lane always owns the cells of one parity, two lattice columns that fall into the same
//     cell (RoIs narrower than two cells per bin) meet in the SAME lane: the lane carries a running sum along such a
//     run and stores every column in order, the last store holding the whole sum.  The eight read-modify-writes of a
//     feature row are therefore always independent -- all loads first, then all stores -- whatever the RoI's width;
//   * a consumer warp owns one lattice row of the RoI in flight.  It forms the row's eight lattice gradients for its
//     lane's channel straight from the tile (the 2x2 / stride-1 average pool's backward is a few adds over the pooled
//     rows i-1 and i), scales them by the column weights, and adds them into the row's upper feature row (phase A) and
//     lower feature row (phase B).  Inside a phase the warps of one RoI touch distinct feature rows; a barrier of the
//     eight warps separates the phases (not needed when the RoI is tall enough for all sixteen rows to differ);
//   * lattice rows that share a start row (RoIs less than a cell high per bin) take turns by their position in the run;
//   * three groups of eight warps take the RoIs in turn.  While one group scatters, the others fetch and reduce theirs;
//     a named barrier (the finishing group arrives, the next group waits) keeps the scatters in list order, so the
//     result is bit-reproducible.
//
// Measured on config 2 (profiles/README.md): 1.54 ms against 2.28 ms; what binds it is the serial chain of scatter
// phases (about 760 cycles per RoI against 256 cycles of shared-memory wavefronts) and, right behind it, the ring feed.
#include "common.cuh"

namespace i2v {

struct alignas(16) PhaseTab {
    // per half-warp h (= parity of the cells it owns) and lattice column j: the cell of the bilinear pair
    // (start, start + 1) that has parity h (a 16-bit cell index), and its weight (validity and the avg pool's 1/4 folded in)
    unsigned short xcell[2][8];
    float w[2][8];
    // per lattice row: {start row * W * 64 bytes (int bits), 1 - fy, fy, packed (int bits)}.  packed: bits 0-3 position
    // of the row in its run of equal start rows (15: row invalid), bits 4-7 the longest such run of the RoI (0: nothing
    // to scatter), bit 8: consecutive valid lattice rows start at least two feature rows apart (all 16 rows distinct),
    // bits 16-31: bit 8 h + j set when, for half-warp h, column j falls into the same cell as column j-1 (the two
    // contributions are then summed in registers).  The RoI-wide fields are repeated in every row so that a warp needs
    // one 16-byte load.
    float4 yrow[8];
};
static_assert(sizeof(PhaseTab) == 224 && sizeof(PhaseTab) <= kRoiTabSlotBytes, "PhaseTab layout");

namespace {

constexpr int kK = 16;
constexpr int kGroups = 3;                              // groups of eight consumer warps that take the RoIs in turn
constexpr int kConsumers = 8 * kGroups;
constexpr int kThreads = (kConsumers + 1) * 32;
constexpr int kStages = 20;
constexpr int kTileBytes = kK * 49 * 4;                 // 3136
constexpr int kTabBytes = (int)sizeof(PhaseTab);        // 224
constexpr int kStageBytes = kTileBytes + kTabBytes;     // 3360
constexpr int kRingBytes = kStages * kStageBytes;
constexpr int kBarBytes = ((2 * kStages * 8 + 32 + 127) / 128) * 128;    // barriers + 32 bytes of zeros (a pooled row outside the grid)
static_assert(kStageBytes % 16 == 0, "ring layout");

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
// named barriers: 1 = all consumer warps, 2 + g = "group g has finished its RoI" (group g arrives, the next group
// waits), 2 + kGroups + g = the warps of group g between their two scatter phases
__device__ __forceinline__ void bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int threads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---------------------------------------------------------------------------------------------- per-RoI tables
__global__ void __launch_bounds__(128) phase_prep_kernel(const LatticeRoi* __restrict__ tab, PhaseTab* __restrict__ ptab,
                                                         int num_rois, int G, int W, float wscale) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= num_rois) return;
    const LatticeRoi& t = tab[n];
    PhaseTab q;
    const unsigned full = (1u << G) - 1u;
    const unsigned vx = t.valid_x & full, vy = t.valid_y & full;
    // Columns: starts are non-decreasing, so the cell of parity h under column j is non-decreasing in j too and equal
    // cells are consecutive.  Each lane sums a run of equal cells in registers (merge) and stores every column in order:
    // the last store of a run carries the whole sum and overwrites the partial ones before it (same thread, same
    // address), so the eight read-modify-writes of a feature row never lose an update, whatever the RoI's width.
    // Columns off the map (a prefix or a suffix) join the nearest valid column's run with weight zero.
    int fv = -1, lv = -1;
#pragma unroll
    for (int p = 0; p < 8; ++p)
        if ((vx >> p) & 1u) {
            if (fv < 0) fv = p;
            lv = p;
        }
    unsigned mergebits = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int prev = -1;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const bool okx = (vx >> p) & 1u;
            const int q_ = okx ? p : (fv < 0 ? 0 : (p < fv ? fv : lv));
            const int st = t.x.start[q_];
            const int cell = st + ((st & 1) ^ h);
            const bool left = (st & 1) == h;                      // this half-warp owns the pair's left cell
            q.xcell[h][p] = (unsigned short)cell;
            q.w[h][p] = okx ? (left ? 1.f - t.x.frac[p] : t.x.frac[p]) * wscale : 0.f;
            if (p > 0 && cell == prev) mergebits |= 1u << (8 * h + p);
            prev = cell;
        }
    }
    const bool any = t.batch >= 0 && vx != 0u && vy != 0u;
    const unsigned maxrun = any ? min(t.y_maxrun, 15u) : 0u;
    unsigned apart = 1;
    {
        int prev = -100;
#pragma unroll
        for (int p = 0; p < 8; ++p)
            if ((vy >> p) & 1u) {
                if (t.y.start[p] - prev < 2) apart = 0;
                prev = t.y.start[p];
            }
    }
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const bool oky = (vy >> p) & 1u;
        const unsigned packed = ((oky && any) ? ((t.y_runpos >> (4 * p)) & 15u) : 15u) | (maxrun << 4) |
                                ((any ? apart : 0u) << 8) | (mergebits << 16);
        q.yrow[p] = make_float4(__int_as_float(oky ? t.y.start[p] * W * 64 : 0), oky ? 1.f - t.y.frac[p] : 0.f,
                                oky ? t.y.frac[p] : 0.f, __int_as_float((int)packed));
    }
    ptab[n] = q;
}

// ---------------------------------------------------------------------------------------------- the kernel
__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(unsigned addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// One feature row of one lattice row: the lane's eight cells += val[j] * wy.  All eight loads are issued before the
// first store; columns that share a cell hold running sums, so storing in column order leaves the complete one.
// Packed fp32x2 fused multiply-add (FFMA2): two lattice columns per instruction.
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b, float c0, float c1) {
    asm("{.reg .b64 ra, rb, rc, rd;\n"
        "mov.b64 ra, {%2, %3};\n"
        "mov.b64 rb, {%4, %4};\n"
        "mov.b64 rc, {%5, %6};\n"
        "fma.rn.f32x2 rd, ra, rb, rc;\n"
        "mov.b64 {%0, %1}, rd;}"
        : "=f"(d0), "=f"(d1)
        : "f"(a0), "f"(a1), "f"(b), "f"(c0), "f"(c1));
}
template <int G>
__device__ __forceinline__ void scatter_row(const unsigned (&addr)[8], unsigned row_off, const float (&val)[8], float wy) {
    float o[8], n[8];
#pragma unroll
    for (int j = 0; j < G; ++j) o[j] = lds_f32(addr[j] + row_off);
#pragma unroll
    for (int j = 0; j + 1 < G; j += 2) ffma2(n[j], n[j + 1], val[j], val[j + 1], wy, o[j], o[j + 1]);
    if (G & 1) n[G - 1] = fmaf(val[G - 1], wy, o[G - 1]);
#pragma unroll
    for (int j = 0; j < G; ++j) sts_f32(addr[j] + row_off, n[j]);
}

template <int POOL, int WT>
__global__ void __launch_bounds__(kThreads, 1)
    lattice_bwd_phase_kernel(const float* __restrict__ grad_out, const PhaseTab* __restrict__ ptab,
                             const int* __restrict__ order, const int* __restrict__ starts, float* __restrict__ grad_in,
                             int C, int H, int Wrt, int accumulate) {
    constexpr int P = 7;
    constexpr int G = (POOL == I2V_POOL_NONE) ? P : P + 1;
    extern __shared__ __align__(128) unsigned char smem[];
    const int W = WT ? WT : Wrt;
    const int HW = H * W;
    unsigned char* ring = smem;                                                    // [kStages][tile | table]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kRingBytes);               // [kStages]
    uint64_t* empty = full + kStages;
    float* planes = reinterpret_cast<float*>(smem + kRingBytes + kBarBytes);       // [H * W][16]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ctiles = C / kK;
    const int ct = blockIdx.x % ctiles;
    const int b = blockIdx.x / ctiles;
    const int list_lo = __ldg(starts + b), list_hi = __ldg(starts + b + 1);
    const int count = list_hi - list_lo;
    // `accumulate`: add the planes to what grad_in holds (the reference launcher's contract, roi_align_kernel.cu:129-141:
    // the caller zeroes bottom_diff and the kernel adds) instead of overwriting it; a frame without RoIs is left alone
    if (accumulate && count == 0) return;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 8);        // the eight warps of the group that owns the stage's RoI
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 8) reinterpret_cast<float*>(empty + kStages)[tid] = 0.f;
    {
        float4* z = reinterpret_cast<float4*>(planes);
        for (int i = tid; i < HW * kK / 4; i += kThreads) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (warp == kConsumers) {
        // ---- producer: the warp reads the frame's RoI list 32 entries at a time; lane 0 walks the ring in order (the
        // consumers release the stages in list order too) and issues the two bulk copies of each RoI ----
        int st = 0;
        unsigned round = 0;
        for (int base = 0; base < count; base += 32) {
            const int mine = (base + lane < count) ? __ldg(order + list_lo + base + lane) : 0;
            const int lim = min(32, count - base);
            for (int i = 0; i < lim; ++i) {
                const int n = __shfl_sync(0xffffffffu, mine, i);
                if (lane == 0) {
                    if (round > 0) mbar_wait(empty + st, (round - 1) & 1);
                    unsigned char* dst = ring + st * kStageBytes;
                    mbar_expect_tx(full + st, kStageBytes);
                    bulk_load(dst, grad_out + ((size_t)n * C + (size_t)ct * kK) * 49, kTileBytes, full + st);
                    bulk_load(dst + kTileBytes, ptab + n, kTabBytes, full + st);
                }
                if (++st == kStages) {
                    st = 0;
                    ++round;
                }
                __syncwarp();
            }
        }
        return;
    }

    // ---- consumers: group = RoIs k = g (mod kGroups), warp in group = lattice row, lane = (dx, channel).  While one
    // group scatters its RoI the others fetch and reduce theirs; the scatters themselves stay in list order (the done
    // barrier of the previous group), so the result does not depend on timing ----
    const int grp = warp >> 3, row = warp & 7;
    const int hx = lane >> 4, c = lane & 15;     // hx: parity of the cells this half-warp owns
    const int row_bytes = W * 64;
    const unsigned lane_planes = smem_u32(planes) + c * 4;
    // pooled rows under lattice row `row`: rows i-1 and i for the avg pool, row i without a pool; a row outside the 7x7
    // grid (and the unused upper row of the pool-less lattice) reads the zero row instead
    const int ra_off = (c * 49 + max(row - 1, 0) * P) * 4, rb_off = (c * 49 + min(row, P - 1) * P) * 4;
    const bool has_a = (POOL != I2V_POOL_NONE) && row >= 1, has_b = row < P;
    const float* zero_row = reinterpret_cast<const float*>(empty + kStages);
    const int bar_prev = 2 + (grp + kGroups - 1) % kGroups, bar_mine = 2 + grp, bar_intra = 2 + kGroups + grp;

    int s = grp % kStages;
    unsigned round = 0;
    for (int k = grp; k < count; k += kGroups) {
        // ---- fetch + reduce: the lane's eight weighted lattice gradients of lattice row `row` ----
        mbar_wait(full + s, round & 1);
        const unsigned char* stage = ring + s * kStageBytes;
        const PhaseTab* t = reinterpret_cast<const PhaseTab*>(stage + kTileBytes);
        const float4 yr = t->yrow[row];
        const unsigned packed = (unsigned)__float_as_int(yr.w);
        const int maxrun = (int)((packed >> 4) & 15u);
        const bool rows_apart = (packed >> 8) & 1u;
        const uint4 xc = *reinterpret_cast<const uint4*>(t->xcell[hx]);         // eight 16-bit cell indices
        const float4 w0 = *reinterpret_cast<const float4*>(t->w[hx]), w1 = *reinterpret_cast<const float4*>(t->w[hx] + 4);
        const float* ra = has_a ? reinterpret_cast<const float*>(stage + ra_off) : zero_row;
        const float* rb = has_b ? reinterpret_cast<const float*>(stage + rb_off) : zero_row;
        const int runpos = (row < G) ? (int)(packed & 15u) : 15;      // 15: no such lattice row
        const float wy0 = yr.y, wy1 = yr.z;
        const unsigned rowa = lane_planes + (unsigned)__float_as_int(yr.x);
        const unsigned addr[8] = {rowa + (xc.x & 0xffffu) * 64u, rowa + (xc.x >> 16) * 64u,
                                  rowa + (xc.y & 0xffffu) * 64u, rowa + (xc.y >> 16) * 64u,
                                  rowa + (xc.z & 0xffffu) * 64u, rowa + (xc.z >> 16) * 64u,
                                  rowa + (xc.w & 0xffffu) * 64u, rowa + (xc.w >> 16) * 64u};
        const unsigned mbits = packed >> (16 + 8 * hx);
        const float wl[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        float val[8];
        if (POOL == I2V_POOL_NONE) {
            // lattice == pooled grid: lattice row i is pooled row i
#pragma unroll
            for (int j = 0; j < P; ++j) val[j] = rb[j] * wl[j];
            val[7] = 0.f;
        } else {
            // lattice row i collects the pooled rows i-1 and i, lattice column j the pooled columns j-1 and j
            float sj[P];
#pragma unroll
            for (int j = 0; j < P; ++j) sj[j] = ra[j] + rb[j];
            val[0] = sj[0] * wl[0];
#pragma unroll
            for (int j = 1; j < P; ++j) val[j] = (sj[j - 1] + sj[j]) * wl[j];
            val[7] = sj[P - 1] * wl[7];
        }
        // columns that fall into the cell of the column before them carry that column's sum along
        if (mbits & 0xfeu) {                          // uniform per half-warp; most RoIs are wide enough to have none
#pragma unroll
            for (int j = 1; j < G; ++j)
                if ((mbits >> j) & 1u) val[j] += val[j - 1];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);      // the stage's bytes are in registers now
        s += kGroups;
        if (s >= kStages) {
            s -= kStages;
            ++round;
        }
        // ---- scatter, after RoI k-1 ----
        if (k > 0) bar_sync(bar_prev, 512);
        if (rows_apart) {
            // tall RoI: the sixteen feature rows under the eight lattice rows are all different, one phase will do
            if (runpos == 0) {
                float oa[8], ob[8], na[8], nb[8];
#pragma unroll
                for (int j = 0; j < G; ++j) oa[j] = lds_f32(addr[j]);
#pragma unroll
                for (int j = 0; j < G; ++j) ob[j] = lds_f32(addr[j] + (unsigned)row_bytes);
#pragma unroll
                for (int j = 0; j + 1 < G; j += 2) {
                    ffma2(na[j], na[j + 1], val[j], val[j + 1], wy0, oa[j], oa[j + 1]);
                    ffma2(nb[j], nb[j + 1], val[j], val[j + 1], wy1, ob[j], ob[j + 1]);
                }
                if (G & 1) {
                    na[G - 1] = fmaf(val[G - 1], wy0, oa[G - 1]);
                    nb[G - 1] = fmaf(val[G - 1], wy1, ob[G - 1]);
                }
#pragma unroll
                for (int j = 0; j < G; ++j) sts_f32(addr[j], na[j]);
#pragma unroll
                for (int j = 0; j < G; ++j) sts_f32(addr[j] + (unsigned)row_bytes, nb[j]);
            }
        } else
        for (int r = 0; r < maxrun; ++r) {          // maxrun == 0: nothing to scatter (uniform over the CTA)
            const bool mine = runpos == r;
            if (r > 0) bar_sync(bar_intra, 256);
            if (mine) scatter_row<G>(addr, 0u, val, wy0);
            bar_sync(bar_intra, 256);
            if (mine) scatter_row<G>(addr, (unsigned)row_bytes, val, wy1);
        }
        // bar.arrive orders this thread's earlier shared-memory accesses before the barrier's completion (PTX ISA,
        // barrier.cta: "prior memory accesses requested by this thread are performed relative to all participants")
        if (k + 1 < count) bar_arrive(bar_mine, 512);
    }

    // ---- write-out: [cell][16] in shared memory -> [16][cell] in HBM.  A lane reads four channels of one cell (16 bytes,
    // the warp 512 contiguous bytes) and stores them to four planes; eight lanes cover one 32-byte sector of a plane ----
    bar_sync(1, kConsumers * 32);
    {
        const int q = lane & 3;
        float* dst = grad_in + ((size_t)b * C + (size_t)ct * kK + (size_t)q * 4) * HW;
        const float4* src = reinterpret_cast<const float4*>(planes);
        for (int cell = warp * 8 + (lane >> 2); cell < HW; cell += kConsumers * 8) {
            float4 v = src[cell * 4 + q];
            if (accumulate) {
                v.x += dst[cell];
                v.y += dst[(size_t)HW + cell];
                v.z += dst[(size_t)2 * HW + cell];
                v.w += dst[(size_t)3 * HW + cell];
            }
            dst[cell] = v.x;
            dst[(size_t)HW + cell] = v.y;
            dst[(size_t)2 * HW + cell] = v.z;
            dst[(size_t)3 * HW + cell] = v.w;
        }
    }
}

}  // namespace

size_t bwd_phase_smem_bytes(int H, int W) { return (size_t)kRingBytes + kBarBytes + (size_t)H * W * kK * sizeof(float); }

bool bwd_phase_ok(const float* grad_out, int batch, int C, int H, int W, int PH, int PW, int pool_mode) {
    return batch > 0 && PH == 7 && PW == 7 && pool_mode != I2V_POOL_MAX && C % kK == 0 && H >= 2 && W >= 2 &&
           bwd_phase_smem_bytes(H, W) <= (size_t)kMaxSmemPerCta && ((uintptr_t)grad_out & 15) == 0;
}

template <int POOL, int WT>
static int launch_phase(const float* grad_out, const PhaseTab* ptab, const int* order, const int* starts, float* grad_in,
                        int batch, int C, int H, int W, int accumulate, cudaStream_t stream) {
    auto kern = lattice_bwd_phase_kernel<POOL, WT>;
    const size_t smem = bwd_phase_smem_bytes(H, W);
    I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3((unsigned)(batch * (C / kK))), kThreads, smem, stream>>>(grad_out, ptab, order, starts, grad_in, C, H, W,
                                                                            accumulate);
    return check_launch("lattice_bwd_phase_kernel");
}

// `tab` holds the LatticeRoi tables of this call; `tab_space` is the workspace's per-RoI table slot (kRoiTabSlotBytes each).
int launch_bwd_phase(const float* grad_out, const LatticeRoi* tab, void* tab_space, const int* order, const int* starts,
                     float* grad_in, int batch, int C, int H, int W, int num_rois, int pool_mode, int accumulate,
                     cudaStream_t stream) {
    PhaseTab* ptab = static_cast<PhaseTab*>(tab_space);
    const int G = pool_mode == I2V_POOL_NONE ? 7 : 8;
    phase_prep_kernel<<<ceil_div(num_rois, 128), 128, 0, stream>>>(tab, ptab, num_rois, G, W,
                                                                    pool_mode == I2V_POOL_AVG ? 0.25f : 1.f);
    I2V_TRY(check_launch("phase_prep_kernel"));
    if (pool_mode == I2V_POOL_AVG) {
        if (W == 63) return launch_phase<I2V_POOL_AVG, 63>(grad_out, ptab, order, starts, grad_in, batch, C, H, W, accumulate, stream);
        return launch_phase<I2V_POOL_AVG, 0>(grad_out, ptab, order, starts, grad_in, batch, C, H, W, accumulate, stream);
    }
    return launch_phase<I2V_POOL_NONE, 0>(grad_out, ptab, order, starts, grad_in, batch, C, H, W, accumulate, stream);
}

}  // namespace i2v
