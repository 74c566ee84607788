// The dense projection of the SGG stage on the 5th-generation tensor cores of sm_100a:
//     y[M,N] = act(x[M,K] . W[N,K]^T + bias[N])
// which is every FC layer of vrd.forward (lib/model/faster_rcnn/resnet_SGG_emb.py:144-177: fc6 50176->4096, fc7,
// fc8, so_vis_embeddings, fc_so, fc_lov, fc_fusion, fc_rel; FC = nn.Linear + optional ReLU, lib/model/faster_rcnn/utils.py:48-60).
// x is K-major (rows of activations), W is nn.Linear's [out, in] layout, i.e. K-major as well.
//
// Wide layers (N > 128, M > 128) run on CTA pairs (linear_tcgen05_pair_kernel below, cta_group::2, 256 x 256 tiles); the
// single-CTA kernel serves the rest and is the fallback (I2V_LINEAR_1CTA=1).  One CTA computes one 128 x 256 output tile
// (128 x 128 when N <= 128):
//   warp 0      TMA producer: cp.async.bulk.tensor loads of the x tile (128 rows) and the W tile (256 rows), one
//               128-byte swizzled k-block (64 bf16 / 32 tf32) per pipeline stage, completion on an mbarrier;
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma (M=128, N=256, K=32 bytes per instruction) straight
//               from the swizzled shared-memory tiles; the fp32 accumulator lives in tensor memory (256 columns);
//               tcgen05.commit hands each stage back to the producer and finally signals the epilogue;
//   warps 2-5   epilogue: tcgen05.ld of the accumulator (lane = tile row), + bias, ReLU, fp32 or bf16 stores.
// Operands are bf16 (kind::f16) or fp32 consumed as tf32 (kind::tf32); accumulation is fp32 in both.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace i2v {
namespace {

constexpr int kBM = 128;
constexpr int kRowBytes = 128;                    // one swizzle row of a k-block
constexpr int kABytes = kBM * kRowBytes;          // 16 KB
constexpr int kThreads = 192;
// Output tile width BN = 256 (4 stages x 48 KB) for the wide layers, 128 (3 stages x 32 KB) when N <= 128 so that half
// of every MMA is not spent on zero padding (conv_lo's 128-channel layer has M = 258 048 rows).
template <int BN>
struct Tile {
    static constexpr int kBN = BN;
    static constexpr int kBBytes = BN * kRowBytes;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = BN == 256 ? 4 : 3;    // 3 x 32 KB: two CTAs per SM, one's epilogue under the other's MMAs
    static constexpr int kTmemCols = BN;          // power of two >= 32
    static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /* alignment slack */ + 256 /* barriers */;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
__device__ __forceinline__ void tma_load_2d(void* sdst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(sdst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* sdst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
            smem_u32(sdst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* sdst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
            smem_u32(sdst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// Shared-memory matrix descriptor of a K-major tile stored as 128-byte rows with the 128B swizzle (what the TMA
// wrote): 8-row groups are 1024 bytes apart (SBO), the leading-dimension offset is unused for swizzled K-major
// tiles, descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}

template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (KIND == 0) {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}


// Accumulator rows -> y: warp quarter q owns TMEM lanes [32 q, +32) = tile rows; 32 columns per tcgen05.ld.
// `mask` (optional, [M, N] bytes, pitch ldm): inverted dropout behind the activation -- y = mask ? y * mscale : 0, the
// F.dropout(x, training=True) of resnet_SGG_emb.py:148-151 with the keep mask drawn by the caller.
struct DropMask {
    const unsigned char* mask = nullptr;
    long long ldm = 0;
    float scale = 1.f;
};
template <int BN>
__device__ __forceinline__ void epilogue_rows(uint32_t tmem_base, int q, int lane, int row, int n0, const float* __restrict__ bias,
                                              void* __restrict__ y, int M, int N, long long ldy, int y_bf16, int relu,
                                              const DropMask dm = DropMask{}) {
        const bool vec_ok = y_bf16 ? ((ldy & 7) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0)
                                   : ((ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0);
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
            const int col0 = n0 + ch * 32;
            if (col0 >= N) break;                      // uniform
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), v);
            if (row < M) {
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float t = __uint_as_float(v[j]);
                    if (bias != nullptr && col0 + j < N) t += __ldg(bias + col0 + j);
                    f[j] = relu ? fmaxf(t, 0.f) : t;
                }
                if (dm.mask != nullptr) {
                    const unsigned char* mrow = dm.mask + (size_t)row * dm.ldm + col0;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + j < N) f[j] = __ldg(mrow + j) ? f[j] * dm.scale : 0.f;
                }
                if (y_bf16) {
                    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(y) + (size_t)row * ldy + col0;
                    if (vec_ok && col0 + 32 <= N) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 pk;
                            __nv_bfloat162 p0 = __floats2bfloat162_rn(f[j], f[j + 1]), p1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
                            __nv_bfloat162 p2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]), p3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
                            pk.x = *reinterpret_cast<uint32_t*>(&p0);
                            pk.y = *reinterpret_cast<uint32_t*>(&p1);
                            pk.z = *reinterpret_cast<uint32_t*>(&p2);
                            pk.w = *reinterpret_cast<uint32_t*>(&p3);
                            *reinterpret_cast<uint4*>(dst + j) = pk;
                        }
                    } else {
                        for (int j = 0; j < 32 && col0 + j < N; ++j) dst[j] = __float2bfloat16_rn(f[j]);
                    }
                } else {
                    float* dst = reinterpret_cast<float*>(y) + (size_t)row * ldy + col0;
                    if (vec_ok && col0 + 32 <= N) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                    } else {
                        for (int j = 0; j < 32 && col0 + j < N; ++j) dst[j] = f[j];
                    }
                }
            }
        }
}

// KIND 0: bf16 operands (64 elements per k-block), KIND 1: fp32 operands read as tf32 (32 elements per k-block).
// Implicit-GEMM convolution (CONV): x is an NHWC activation seen through a 4-D tensor map (channels, x, y, image) whose
// element strides are the convolution's stride; k-block kb is channel chunk kb % chunks of filter tap kb / chunks, and the
// 128 tile rows are `rows_per_image` output positions of 128 / rows_per_image consecutive images.  The TMA zero-fills the
// padding border and the channel padding, so no patch matrix is ever written.
// SPLIT (stride 2 only): the activation is stored as four parity planes per image, [N, 2, 2, H/2, W/2, C] (plane
// (y & 1, x & 1), position (y >> 1, x >> 1)), seen through a 5-D map with unit element strides; the samples 2 o + k - pad of
// tap k are then the DENSE box of plane (k - pad) & 1 at offset (k - pad) >> 1.  The strided 4-D box makes the TMA walk
// every position it skips (the layer ran at 42 % of the tensor peak, waiting for its A tiles); the dense one does not.
struct ConvGeom {
    int chunks;          // 64-channel chunks per tap
    int kw;              // filter width (taps per filter row)
    int pad;
    int rows_per_image;  // OH * OW
    int split;           // 1: parity-split activation
};

template <int KIND, int BN, bool CONV = false>
__global__ void __launch_bounds__(kThreads, 1)
    linear_tcgen05_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                          const float* __restrict__ bias, void* __restrict__ y, int M, int N, int K, long long ldy,
                          int y_bf16, int relu, ConvGeom cg = ConvGeom{}, const DropMask dm = DropMask{}) {
    constexpr int ELEMS = (KIND == 0) ? 64 : 32;
    constexpr int kBN = Tile<BN>::kBN, kStages = Tile<BN>::kStages, kStageBytes = Tile<BN>::kStageBytes;
    constexpr int kTmemCols = Tile<BN>::kTmemCols;
    // instruction descriptor: D = fp32, A/B format (1 = bf16 under kind::f16, 2 = tf32), both K-major, N >> 3, M >> 4
    constexpr uint32_t FMT = (KIND == 0) ? 1u : 2u;
    constexpr uint32_t IDESC = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

    extern __shared__ uint8_t raw_smem[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw_smem) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes);
    uint64_t* empty = full + kStages;
    uint64_t* accum = empty + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
    const int kblocks = (K + ELEMS - 1) / ELEMS;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {  // tensor memory for the 128 x 256 fp32 accumulator
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kStages;
                const unsigned round = (unsigned)(kb / kStages);
                mbar_wait(empty + s, (round & 1u) ^ 1u);
                mbar_expect_tx(full + s, kStageBytes);
                uint8_t* a = smem + (size_t)s * kStageBytes;
                if (CONV) {
                    const int tap = kb / cg.chunks, chunk = kb - tap * cg.chunks;
                    const int ky = tap / cg.kw, kx = tap - ky * cg.kw;
                    const int dx = kx - cg.pad, dy = ky - cg.pad;
                    if (cg.split)
                        tma_load_5d(a, &map_x, chunk * ELEMS, dx >> 1, dy >> 1, ((dy & 1) << 1) | (dx & 1),
                                    m0 / cg.rows_per_image, full + s);
                    else
                        tma_load_4d(a, &map_x, chunk * ELEMS, dx, dy, m0 / cg.rows_per_image, full + s);
                } else {
                    tma_load_2d(a, &map_x, kb * ELEMS, m0, full + s);
                }
                tma_load_2d(a + kABytes, &map_w, kb * ELEMS, n0, full + s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kStages;
                const unsigned round = (unsigned)(kb / kStages);
                mbar_wait(full + s, round & 1u);
                tc_fence_after();
                const uint32_t a = smem_u32(smem + (size_t)s * kStageBytes);
                const uint64_t adesc = umma_desc(a), bdesc = umma_desc(a + kABytes);
#pragma unroll
                for (int k = 0; k < kRowBytes / 32; ++k) {
                    // +32 bytes along K inside the swizzle row = +2 in the 16-byte units of the descriptor
                    umma<KIND>(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), IDESC,
                               (kb | k) != 0 ? 1u : 0u);
                }
                tc_commit(empty + s);   // the stage is free once these MMAs have read it
            }
            tc_commit(accum);           // all MMAs done: the accumulator is complete
        }
    } else {
        // ---- epilogue: warp w may touch TMEM lanes [32 (w % 4), +32) = tile rows ----
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        mbar_wait(accum, 0);
        tc_fence_after();
        epilogue_rows<kBN>(tmem_base, q, lane, row, n0, bias, y, M, N, ldy, y_bf16, relu, dm);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols)
                     : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- CTA-pair variant
// Two CTAs of a cluster (one TPC) compute a 256 x 256 tile with tcgen05.mma.cta_group::2: CTA r holds the x rows
// [128 r, +128) and the W rows [128 r, +128) of the tile, the leader's single thread issues one M = 256 MMA that reads A
// from both shared memories and either half of B from its owner, and each CTA ends up with its own 128 accumulator rows
// in its own tensor memory.  Per k-block a CTA moves 32 KB instead of 48 KB through L2 and shared memory for the same
// flops, which is what the single-CTA kernel is bound by (19 GB of tile traffic for fc6 = 14.5 TB/s at 78 % of peak).
constexpr size_t pair_smem_bytes(int stages, int bn = 256) {       // a stage: 128 x rows + bn / 2 W rows (32 KB at bn = 256)
    return (size_t)stages * (kABytes + (bn / 2) * kRowBytes) + 1024 + 256;
}
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;                        // shared::cluster address of the same offset in CTA 0

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* sdst, const CUtensorMap* map, int c0, int c1, uint64_t* leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(sdst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(leader_bar) & kPeerMask)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* sdst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                                 uint64_t* leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, "
        "%5}], [%6];" ::"r"(smem_u32(sdst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(leader_bar) & kPeerMask)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(void* sdst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                                 uint64_t* leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, "
        "%5, %6}], [%7];" ::"r"(smem_u32(sdst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4),
        "r"(smem_u32(leader_bar) & kPeerMask)
        : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {     // arrives on `bar` in both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

// BN = 128 (one tile column of at most 128 outputs: conv_lo's layers) halves the W half to 64 rows; CONV reads x through
// the implicit-GEMM maps of linear_tcgen05_kernel.
template <int KIND, int kPairStages, int BN = 256, bool CONV = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    linear_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                               const float* __restrict__ bias, void* __restrict__ y, int M, int N, int K, long long ldy,
                               int y_bf16, int relu, const DropMask dm, ConvGeom cg) {
    constexpr int ELEMS = (KIND == 0) ? 64 : 32;
    constexpr uint32_t FMT = (KIND == 0) ? 1u : 2u;
    // D = fp32, A/B format, K-major, N = BN (>> 3), M = 256 (>> 4)
    constexpr uint32_t IDESC = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    constexpr int kTmemCols = BN;
    constexpr int kPairStageBytes = kABytes + (BN / 2) * kRowBytes;    // this CTA's x rows + its half of the W tile

    extern __shared__ uint8_t raw_smem[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw_smem) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kPairStages * kPairStageBytes);
    uint64_t* empty = full + kPairStages;
    uint64_t* accum = empty + kPairStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const bool leader = rank == 0;
    const int m0 = (int)(blockIdx.x >> 1) * 256 + (int)rank * 128;   // this CTA's 128 rows of x
    const int n0 = blockIdx.y * BN;                                  // the pair's BN columns
    const int kblocks = (K + ELEMS - 1) / ELEMS;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
        for (int s = 0; s < kPairStages; ++s) {
            mbar_init(full + s, 1);      // the leader's arrive.expect_tx covers both CTAs' bytes (used in CTA 0 only)
            mbar_init(empty + s, 1);     // one multicast commit per use
        }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                  // both CTAs' barriers are initialised before any remote arrive / TMA completion
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kPairStages;
                const unsigned round = (unsigned)(kb / kPairStages);
                mbar_wait(empty + s, (round & 1u) ^ 1u);
                // Only the leader arrives: its expect_tx announces the bytes of BOTH CTAs.  The peer's loads can complete
                // first (the transaction count goes negative for a while), but the phase cannot flip before the leader's
                // arrival, and the peer cannot run a stage ahead because its empty barrier follows the leader's MMAs.  A
                // remote arrive per k-block from the peer would cost a cluster-scope release fence (~1500 cycles here).
                if (leader) mbar_expect_tx(full + s, 2 * kPairStageBytes);
                uint8_t* a = smem + (size_t)s * kPairStageBytes;
                if (CONV) {
                    const int tap = kb / cg.chunks, chunk = kb - tap * cg.chunks;
                    const int ky = tap / cg.kw, kx = tap - ky * cg.kw;
                    const int dx = kx - cg.pad, dy = ky - cg.pad;
                    if (cg.split)
                        tma_load_5d_pair(a, &map_x, chunk * ELEMS, dx >> 1, dy >> 1, ((dy & 1) << 1) | (dx & 1),
                                         m0 / cg.rows_per_image, full + s);
                    else
                        tma_load_4d_pair(a, &map_x, chunk * ELEMS, dx, dy, m0 / cg.rows_per_image, full + s);
                } else {
                    tma_load_2d_pair(a, &map_x, kb * ELEMS, m0, full + s);
                }
                tma_load_2d_pair(a + kABytes, &map_w, kb * ELEMS, n0 + (int)rank * (BN / 2), full + s);
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kPairStages;
                const unsigned round = (unsigned)(kb / kPairStages);
                mbar_wait(full + s, round & 1u);
                tc_fence_after();
                const uint32_t a = smem_u32(smem + (size_t)s * kPairStageBytes);
                const uint64_t adesc = umma_desc(a), bdesc = umma_desc(a + kABytes);
#pragma unroll
                for (int k = 0; k < kRowBytes / 32; ++k) {
                    const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
                    if (KIND == 0) {
                        asm volatile(
                            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_base),
                            "l"(adesc + (uint64_t)(2 * k)), "l"(bdesc + (uint64_t)(2 * k)), "r"(IDESC), "r"(acc)
                            : "memory");
                    } else {
                        asm volatile(
                            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_base),
                            "l"(adesc + (uint64_t)(2 * k)), "l"(bdesc + (uint64_t)(2 * k)), "r"(IDESC), "r"(acc)
                            : "memory");
                    }
                }
                tc_commit_pair(empty + s);
            }
            tc_commit_pair(accum);
        }
    } else {
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        mbar_wait(accum, 0);
        tc_fence_after();
        epilogue_rows<BN>(tmem_base, q, lane, row, n0, bias, y, M, N, ldy, y_bf16, relu, dm);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                  // the peer may still be reading this CTA's shared memory / signalling its barriers
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols)
                     : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- small kernels
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                        int64_t rows, int64_t cols, int64_t lds, int64_t ldd) {
    int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / cols, c = i - r * cols;
        dst[r * ldd + c] = __float2bfloat16_rn(src[r * lds + c]);
    }
}

// resnet_SGG_emb.py:207-219: scores = softmax(normalize(x) . normalize(prd)^T) (softmax only when `apply_softmax`).
// F.normalize divides by max(||v||_2, 1e-12).  One warp per row of x; the predicate embeddings are normalised by the
// first kernel.  Plain fp32 CUDA-core arithmetic: P x R x E = 4032 x 132 x 300 is 0.3 GFLOP.
__global__ void __launch_bounds__(256) l2_normalize_rows_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                int rows, int cols) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* s = src + (size_t)row * cols;
    float acc = 0.f;
    for (int i = lane; i < cols; i += 32) acc += s[i] * s[i];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    float inv = 1.f / fmaxf(sqrtf(acc), 1e-12f);
    for (int i = lane; i < cols; i += 32) dst[(size_t)row * cols + i] = s[i] * inv;
}

constexpr int kScoreWarps = 8;
constexpr int kScoreRows = 4;       // pair rows per warp: every predicate-embedding load feeds four dot products
// STAGED: the normalised predicate embeddings [R][E] sit in shared memory for the whole CTA (R*E*4 = 158 KB at 132 x 300);
// otherwise every dot product pulls them from L2 and the kernel is latency-bound on those loads.
template <bool STAGED>
__global__ void __launch_bounds__(kScoreWarps * 32)
    rel_score_kernel(const float* __restrict__ x, const float* __restrict__ prdn, float* __restrict__ scores, int P, int R,
                     int E, int apply_softmax) {
    extern __shared__ __align__(16) float sc_smem[];   // [warps][rows][E] x rows, [warps][rows][R] similarities, ([R][E])
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* pr = prdn;
    if (STAGED) {
        float* ps = sc_smem + (size_t)kScoreWarps * kScoreRows * (E + R);
        for (int i = threadIdx.x; i < R * E; i += kScoreWarps * 32) ps[i] = __ldg(prdn + i);
        __syncthreads();
        pr = ps;
    }
    const int row0 = (blockIdx.x * kScoreWarps + warp) * kScoreRows;
    if (row0 >= P) return;
    float* xs = sc_smem + (size_t)warp * kScoreRows * E;
    float* sim = sc_smem + (size_t)kScoreWarps * kScoreRows * E + (size_t)warp * kScoreRows * R;
#pragma unroll
    for (int q = 0; q < kScoreRows; ++q) {
        const int row = min(row0 + q, P - 1);            // a ragged tail recomputes the last row (never stored)
        const float* xr = x + (size_t)row * E;
        float acc = 0.f;
        for (int i = lane; i < E; i += 32) {
            float v = xr[i];
            xs[q * E + i] = v;
            acc += v * v;
        }
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        const float inv = 1.f / fmaxf(sqrtf(acc), 1e-12f);
        for (int i = lane; i < E; i += 32) xs[q * E + i] *= inv;
    }
    __syncwarp();
    for (int r = 0; r < R; ++r) {
        const float* p = pr + (size_t)r * E;
        float d[kScoreRows];
#pragma unroll
        for (int q = 0; q < kScoreRows; ++q) d[q] = 0.f;
        for (int i = lane; i < E; i += 32) {
            const float w = p[i];
#pragma unroll
            for (int q = 0; q < kScoreRows; ++q) d[q] = fmaf(xs[q * E + i], w, d[q]);
        }
#pragma unroll
        for (int q = 0; q < kScoreRows; ++q) {
            for (int o = 16; o; o >>= 1) d[q] += __shfl_xor_sync(0xffffffffu, d[q], o);
            if (lane == 0) sim[q * R + r] = d[q];
        }
    }
    __syncwarp();
    for (int q = 0; q < kScoreRows; ++q) {
        const int row = row0 + q;
        if (row >= P) break;
        float* out = scores + (size_t)row * R;
        float* sq = sim + q * R;
        if (!apply_softmax) {
            for (int r = lane; r < R; r += 32) out[r] = sq[r];
            continue;
        }
        float mx = -INFINITY;
        for (int r = lane; r < R; r += 32) mx = fmaxf(mx, sq[r]);
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int r = lane; r < R; r += 32) {
            float e = expf(sq[r] - mx);
            sq[r] = e;
            sum += e;
        }
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        for (int r = lane; r < R; r += 32) out[r] = sq[r] / sum;
    }
}


// Tiled variant: lanes own predicates (r = lane + 32 j), a warp works on eight rows of x at a time, and the dot products
// run over E with the predicate table TRANSPOSED in shared memory ([E][RP], RP = R rounded up to 32, zero padded, written
// that way by the normalisation kernel): per element of E a warp issues RP/32 conflict-free shared loads that feed
// 8 x RP/32 FMAs, the eight x values arriving as broadcast 16-byte loads.  No cross-lane reduction is left in the dot
// products; only the norms and the softmax use shuffles.  One persistent CTA per SM keeps the table for all its rows.
constexpr int kTileRows = 8;
constexpr int kTileWarps = 16;
template <int J>    // RP / 32
__global__ void __launch_bounds__(kTileWarps * 32, 1)
    rel_score_tile_kernel(const float* __restrict__ x, const float* __restrict__ prdT, float* __restrict__ scores, int P, int R,
                          int E, int apply_softmax) {
    extern __shared__ __align__(16) float sc_smem[];   // [E][32 J]
    constexpr int RP = 32 * J;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < E * RP / 4; i += kTileWarps * 32)
        reinterpret_cast<float4*>(sc_smem)[i] = __ldg(reinterpret_cast<const float4*>(prdT) + i);
    __syncthreads();
    const int tiles = (P + kTileRows - 1) / kTileRows;
    for (int tile = blockIdx.x * kTileWarps + warp; tile < tiles; tile += gridDim.x * kTileWarps) {
        const int row0 = tile * kTileRows;
        const float* xr[kTileRows];
        float inv[kTileRows];
#pragma unroll
        for (int q = 0; q < kTileRows; ++q) {
            xr[q] = x + (size_t)min(row0 + q, P - 1) * E;            // a ragged tail recomputes the last row (never stored)
            float ss = 0.f;
            for (int i = lane; i < E; i += 32) {
                const float v = __ldg(xr[q] + i);
                ss = fmaf(v, v, ss);
            }
            for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            inv[q] = 1.f / fmaxf(sqrtf(ss), 1e-12f);
        }
        float acc[kTileRows][J];
#pragma unroll
        for (int q = 0; q < kTileRows; ++q)
#pragma unroll
            for (int j = 0; j < J; ++j) acc[q][j] = 0.f;
        const float* ps = sc_smem + lane;
#pragma unroll 1
        for (int e4 = 0; e4 < E / 4; ++e4) {
            float4 xv[kTileRows];
#pragma unroll
            for (int q = 0; q < kTileRows; ++q) xv[q] = __ldg(reinterpret_cast<const float4*>(xr[q]) + e4);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float pv[J];
#pragma unroll
                for (int j = 0; j < J; ++j) pv[j] = ps[(e4 * 4 + t) * RP + 32 * j];
#pragma unroll
                for (int q = 0; q < kTileRows; ++q) {
                    const float xq = t == 0 ? xv[q].x : t == 1 ? xv[q].y : t == 2 ? xv[q].z : xv[q].w;
#pragma unroll
                    for (int j = 0; j < J; ++j) acc[q][j] = fmaf(xq, pv[j], acc[q][j]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < kTileRows; ++q) {
            const int row = row0 + q;
            if (row >= P) break;
            float* out = scores + (size_t)row * R;
            float sv[J];
#pragma unroll
            for (int j = 0; j < J; ++j) sv[j] = acc[q][j] * inv[q];
            if (apply_softmax) {
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < J; ++j)
                    if (lane + 32 * j < R) mx = fmaxf(mx, sv[j]);
                for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    sv[j] = lane + 32 * j < R ? expf(sv[j] - mx) : 0.f;
                    sum += sv[j];
                }
                for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
#pragma unroll
                for (int j = 0; j < J; ++j) sv[j] = sv[j] / sum;
            }
#pragma unroll
            for (int j = 0; j < J; ++j)
                if (lane + 32 * j < R) out[lane + 32 * j] = sv[j];
        }
    }
}

// normalised predicate embeddings, transposed and zero padded: dst [cols][rp], rp >= rows
__global__ void __launch_bounds__(256) l2_normalize_rows_t_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                  int rows, int cols, int rp) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rp) return;
    if (row >= rows) {
        for (int i = lane; i < cols; i += 32) dst[(size_t)i * rp + row] = 0.f;
        return;
    }
    const float* s = src + (size_t)row * cols;
    float acc = 0.f;
    for (int i = lane; i < cols; i += 32) acc += s[i] * s[i];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    float inv = 1.f / fmaxf(sqrtf(acc), 1e-12f);
    for (int i = lane; i < cols; i += 32) dst[(size_t)i * rp + row] = s[i] * inv;
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows, cols] K-major matrix with `ld` elements between rows; box = one 128-byte k-block x `box_rows` rows.
int make_map(CUtensorMap* map, const void* base, int kind, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) {
        set_error("linear_forward: cuTensorMapEncodeTiled is not available from this driver");
        return I2V_ERR_CUDA;
    }
    const size_t esz = kind == 0 ? 2 : 4;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
    cuuint32_t box[2] = {(cuuint32_t)(kRowBytes / esz), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("linear_forward: cuTensorMapEncodeTiled failed (%d) for a [%lld x %lld] matrix, ld %lld", (int)r,
                  (long long)rows, (long long)cols, (long long)ld);
        return I2V_ERR_INVALID;
    }
    return I2V_OK;
}

}  // namespace
}  // namespace i2v

using namespace i2v;

static int linear_impl(const void* x, const void* w, const float* bias, void* y, int M, int N, int K, long long ldx,
                       long long ldw, long long ldy, int in_dtype, int out_dtype, int relu, const DropMask dm,
                       cudaStream_t stream) {
    I2V_REQUIRE(M >= 0 && N >= 0 && K >= 1, "linear_forward: bad shape %d x %d x %d", M, N, K);
    I2V_REQUIRE(in_dtype == I2V_DT_BF16 || in_dtype == I2V_DT_TF32, "linear_forward: in_dtype %d", in_dtype);
    I2V_REQUIRE(out_dtype == I2V_DT_BF16 || out_dtype == I2V_DT_F32, "linear_forward: out_dtype %d", out_dtype);
    if (M == 0 || N == 0) return I2V_OK;
    I2V_REQUIRE(x && w && y, "linear_forward: null pointer");
    const int kind = in_dtype == I2V_DT_BF16 ? 0 : 1;
    const size_t esz = kind == 0 ? 2 : 4;
    I2V_REQUIRE(ldx >= K && ldw >= K && ldy >= N, "linear_forward: leading dimension smaller than the row");
    I2V_REQUIRE((ldx * esz) % 16 == 0 && (ldw * esz) % 16 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0,
                "linear_forward: x and W need 16-byte aligned bases and row pitches (TMA)");
    alignas(64) CUtensorMap map_x, map_w;
    I2V_TRY(make_map(&map_x, x, kind, M, K, ldx, kBM));
    const int bn = N <= 128 ? 128 : 256;
    static const bool pair_off = getenv("I2V_LINEAR_1CTA") != nullptr;
    if (bn == 256 && M > 128 && !pair_off) {
        // CTA pairs: every CTA loads its own 128 x rows and 128 of the tile's 256 W rows
        I2V_TRY(make_map(&map_w, w, kind, N, K, ldw, 128));
        dim3 grid2(2u * (unsigned)ceil_div(M, 256), (unsigned)ceil_div(N, 256));
        const int yb2 = out_dtype == I2V_DT_BF16;
        constexpr int kPairStages = 6;
        if (kind == 0) {
            auto kern = linear_tcgen05_pair_kernel<0, kPairStages>;
            I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pair_smem_bytes(kPairStages)));
            kern<<<grid2, kThreads, pair_smem_bytes(kPairStages), stream>>>(map_x, map_w, bias, y, M, N, K, ldy, yb2, relu, dm, ConvGeom{});
        } else {
            auto kern = linear_tcgen05_pair_kernel<1, kPairStages>;
            I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pair_smem_bytes(kPairStages)));
            kern<<<grid2, kThreads, pair_smem_bytes(kPairStages), stream>>>(map_x, map_w, bias, y, M, N, K, ldy, yb2, relu, dm, ConvGeom{});
        }
        return check_launch("linear_tcgen05_pair_kernel");
    }
    I2V_TRY(make_map(&map_w, w, kind, N, K, ldw, bn));
    dim3 grid((unsigned)ceil_div(N, bn), (unsigned)ceil_div(M, kBM));
    const int yb = out_dtype == I2V_DT_BF16;
#define I2V_LAUNCH_LINEAR(KIND, BN)                                                                                   \
    do {                                                                                                              \
        auto kern = linear_tcgen05_kernel<KIND, BN>;                                                                  \
        I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile<BN>::kSmemBytes)); \
        kern<<<grid, kThreads, Tile<BN>::kSmemBytes, stream>>>(map_x, map_w, bias, y, M, N, K, ldy, yb, relu, ConvGeom{}, dm); \
    } while (0)
    if (kind == 0 && bn == 256) I2V_LAUNCH_LINEAR(0, 256);
    else if (kind == 0) I2V_LAUNCH_LINEAR(0, 128);
    else if (bn == 256) I2V_LAUNCH_LINEAR(1, 256);
    else I2V_LAUNCH_LINEAR(1, 128);
#undef I2V_LAUNCH_LINEAR
    return check_launch("linear_tcgen05_kernel");
}

extern "C" int i2v_linear_forward(const void* x, const void* w, const float* bias, void* y, int M, int N, int K,
                                  long long ldx, long long ldw, long long ldy, int in_dtype, int out_dtype, int relu,
                                  cudaStream_t stream) {
    return linear_impl(x, w, bias, y, M, N, K, ldx, ldw, ldy, in_dtype, out_dtype, relu, DropMask{}, stream);
}

extern "C" int i2v_linear_forward_dropout(const void* x, const void* w, const float* bias, void* y, int M, int N, int K,
                                          long long ldx, long long ldw, long long ldy, int in_dtype, int out_dtype, int relu,
                                          const unsigned char* keep_mask, long long ldm, float scale, cudaStream_t stream) {
    I2V_REQUIRE(keep_mask && ldm >= N, "linear_forward_dropout: the keep mask is [M, N] bytes with pitch >= N");
    DropMask dm;
    dm.mask = keep_mask;
    dm.ldm = ldm;
    dm.scale = scale;
    return linear_impl(x, w, bias, y, M, N, K, ldx, ldw, ldy, in_dtype, out_dtype, relu, dm, stream);
}

extern "C" int i2v_cast_bf16(const float* src, void* dst, long long rows, long long cols, long long lds, long long ldd,
                             cudaStream_t stream) {
    I2V_REQUIRE(rows >= 0 && cols >= 0 && lds >= cols && ldd >= cols, "cast_bf16: bad shape");
    if (rows == 0 || cols == 0) return I2V_OK;
    I2V_REQUIRE(src && dst, "cast_bf16: null pointer");
    cast_bf16_kernel<<<grid_for(rows * cols, 256), 256, 0, stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), rows,
                                                                      cols, lds, ldd);
    return check_launch("cast_bf16_kernel");
}

extern "C" size_t i2v_rel_scores_workspace_bytes(int num_rel, int emb_dim) {
    if (num_rel < 0 || emb_dim < 0) return 0;
    return align_up((size_t)align_up((size_t)num_rel, 32) * emb_dim * sizeof(float), 256);   // rows padded to 32 (tiled kernel)
}

extern "C" int i2v_rel_scores(const float* x, const float* prd, float* scores, int num_pairs, int num_rel, int emb_dim,
                              int apply_softmax, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    I2V_REQUIRE(num_pairs >= 0 && num_rel >= 1 && emb_dim >= 1, "rel_scores: bad shape");
    if (num_pairs == 0) return I2V_OK;
    I2V_REQUIRE(x && prd && scores, "rel_scores: null pointer");
    size_t need = i2v_rel_scores_workspace_bytes(num_rel, emb_dim);
    if (!workspace || workspace_bytes < need) {
        set_error("rel_scores: workspace %zu < %zu bytes", workspace_bytes, need);
        return I2V_ERR_WORKSPACE;
    }
    float* prdn = static_cast<float*>(workspace);
    {   // tiled kernel: transposed table in shared memory, as many predicates per lane as the template holds
        const int rp = (int)align_up((size_t)num_rel, 32), J = rp / 32;
        const size_t table = (size_t)emb_dim * rp * sizeof(float);
        if (J <= 5 && emb_dim % 4 == 0 && table <= (size_t)kMaxSmemPerCta && ((uintptr_t)x & 15) == 0 &&
            num_pairs >= 64 && !getenv("I2V_REL_SCORE_ROWS")) {
            l2_normalize_rows_t_kernel<<<ceil_div(rp, 8), 256, 0, stream>>>(prd, prdn, num_rel, emb_dim, rp);
            I2V_TRY(check_launch("l2_normalize_rows_t_kernel"));
            const int tiles = ceil_div(num_pairs, kTileRows);
            const int grid = min(kNumSMs, ceil_div(tiles, kTileWarps));
#define I2V_SCORE_TILE(JJ)                                                                                              \
    do {                                                                                                                \
        auto kern = rel_score_tile_kernel<JJ>;                                                                          \
        I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)table));              \
        kern<<<grid, kTileWarps * 32, table, stream>>>(x, prdn, scores, num_pairs, num_rel, emb_dim, apply_softmax);    \
    } while (0)
            switch (J) {
                case 1: I2V_SCORE_TILE(1); break;
                case 2: I2V_SCORE_TILE(2); break;
                case 3: I2V_SCORE_TILE(3); break;
                case 4: I2V_SCORE_TILE(4); break;
                default: I2V_SCORE_TILE(5); break;
            }
#undef I2V_SCORE_TILE
            return check_launch("rel_score_tile_kernel");
        }
    }
    l2_normalize_rows_kernel<<<ceil_div(num_rel, 8), 256, 0, stream>>>(prd, prdn, num_rel, emb_dim);
    I2V_TRY(check_launch("l2_normalize_rows_kernel"));
    size_t smem = (size_t)kScoreWarps * kScoreRows * ((size_t)emb_dim + num_rel) * sizeof(float);
    I2V_REQUIRE(smem <= (size_t)kMaxSmemPerCta, "rel_scores: emb_dim + num_rel too large for shared memory");
    size_t staged = smem + (size_t)num_rel * emb_dim * sizeof(float);
    int grid = ceil_div(num_pairs, kScoreWarps * kScoreRows);
    if (staged <= (size_t)kMaxSmemPerCta && grid >= 8) {
        I2V_CUDA_TRY(cudaFuncSetAttribute(rel_score_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged));
        rel_score_kernel<true><<<grid, kScoreWarps * 32, staged, stream>>>(x, prdn, scores, num_pairs, num_rel, emb_dim,
                                                                          apply_softmax);
    } else {
        I2V_CUDA_TRY(cudaFuncSetAttribute(rel_score_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rel_score_kernel<false><<<grid, kScoreWarps * 32, smem, stream>>>(x, prdn, scores, num_pairs, num_rel, emb_dim,
                                                                         apply_softmax);
    }
    return check_launch("rel_score_kernel");
}

// conv_lo's strided layers as an implicit GEMM (resnet_SGG_emb.py:107-110): x [N,H,W,C] bf16 NHWC, w [O, KH*KW*Cp] bf16
// with every tap's channels padded to Cp = 64 * ceil(C / 64), y [N*OH*OW, O] (NHWC of the next layer).
static int conv2d_nhwc_impl(const void* x, const void* w, const float* bias, void* y, int n, int height, int width,
                            int channels, int out_channels, int kernel, int stride, int pad, long long ldw, long long ldy,
                            int out_dtype, int relu, bool split, cudaStream_t stream) {
    I2V_REQUIRE(n >= 0 && height >= 1 && width >= 1 && channels >= 1 && out_channels >= 1 && kernel >= 1 && stride >= 1 &&
                    pad >= 0,
                "conv2d_nhwc: bad shape");
    I2V_REQUIRE(out_dtype == I2V_DT_BF16 || out_dtype == I2V_DT_F32, "conv2d_nhwc: out_dtype %d", out_dtype);
    const int OH = (height + 2 * pad - kernel) / stride + 1, OW = (width + 2 * pad - kernel) / stride + 1;
    const int rows = OH * OW, chunks = ceil_div(channels, 64), K = kernel * kernel * chunks * 64;
    bool ok = OH >= 1 && OW >= 1 && rows <= kBM && kBM % rows == 0 && channels % 8 == 0 && out_channels <= 128 &&
              OW * stride <= 256 && OH * stride <= 256 && ldw >= K && (ldw * 2) % 16 == 0 && ((uintptr_t)x & 15) == 0 &&
              ((uintptr_t)w & 15) == 0;
    if (split)   // parity planes: stride 2, even map, "same" padding of an odd kernel, output = one plane's size
        ok = ok && stride == 2 && height % 2 == 0 && width % 2 == 0 && OH == height / 2 && OW == width / 2;
    if (!ok) {
        set_error("conv2d_nhwc: needs OH*OW dividing 128, C %% 8 == 0, at most 128 output channels and 16-byte aligned rows");
        return I2V_ERR_UNSUPPORTED;
    }
    if (n == 0) return I2V_OK;
    I2V_REQUIRE(x && w && y, "conv2d_nhwc: null pointer");
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) {
        set_error("conv2d_nhwc: cuTensorMapEncodeTiled is not available from this driver");
        return I2V_ERR_CUDA;
    }
    alignas(64) CUtensorMap map_x, map_w;
    if (split) {
        const int per_tile = kBM / rows;
        const cuuint64_t plane = (cuuint64_t)OH * OW * channels * 2;
        cuuint64_t dims[5] = {(cuuint64_t)channels, (cuuint64_t)OW, (cuuint64_t)OH, 4, (cuuint64_t)n};
        cuuint64_t strides[4] = {(cuuint64_t)channels * 2, (cuuint64_t)OW * channels * 2, plane, 4 * plane};
        cuuint32_t box[5] = {64, (cuuint32_t)OW, (cuuint32_t)OH, 1, (cuuint32_t)per_tile};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = fn(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("conv2d_nhwc: cuTensorMapEncodeTiled failed (%d)", (int)r);
            return I2V_ERR_INVALID;
        }
    } else {
        const int per_tile = kBM / rows;
        cuuint64_t dims[4] = {(cuuint64_t)channels, (cuuint64_t)width, (cuuint64_t)height, (cuuint64_t)n};
        cuuint64_t strides[3] = {(cuuint64_t)channels * 2, (cuuint64_t)width * channels * 2,
                                 (cuuint64_t)height * width * channels * 2};
        // with an element stride s the box extent counts traversed positions: OW samples span OW * s of them
        cuuint32_t box[4] = {64, (cuuint32_t)(OW * stride), (cuuint32_t)(OH * stride), (cuuint32_t)per_tile};
        cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        CUresult r = fn(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("conv2d_nhwc: cuTensorMapEncodeTiled failed (%d)", (int)r);
            return I2V_ERR_INVALID;
        }
    }
    I2V_TRY(make_map(&map_w, w, 0, out_channels, K, ldw, 128));
    const int M = n * rows;
    ConvGeom cg{chunks, kernel, pad, rows, split ? 1 : 0};
    static const bool pair_off = getenv("I2V_LINEAR_1CTA") != nullptr;
    if (M > kBM && !pair_off) {
        // CTA pairs: a 256-row tile per pair, every CTA loading its own 128 rows of patches and 64 of the 128 W rows.  With a
        // 128-wide tile an SM's shared memory moves 64 KB per k-block (32 KB written by the TMA, 32 KB read by the MMAs) in
        // the 256 cycles the MMAs take -- twice its 128 B/clk -- and the pair brings that to 48 KB (0.93 -> 0.86 ms for
        // conv_lo's second layer); four stages, so that two CTAs share an SM as in the 1-CTA kernel
        alignas(64) CUtensorMap map_w2;
        I2V_TRY(make_map(&map_w2, w, 0, out_channels, K, ldw, 64));
        constexpr int kConvStages = 4;
        auto kern2 = linear_tcgen05_pair_kernel<0, kConvStages, 128, true>;
        I2V_CUDA_TRY(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)pair_smem_bytes(kConvStages, 128)));
        dim3 grid2(2u * (unsigned)ceil_div(M, 256), 1u);
        kern2<<<grid2, kThreads, pair_smem_bytes(kConvStages, 128), stream>>>(map_x, map_w2, bias, y, M, out_channels, K, ldy,
                                                                             out_dtype == I2V_DT_BF16, relu, DropMask{}, cg);
        return check_launch("linear_tcgen05_pair_kernel<conv>");
    }
    dim3 grid(1u, (unsigned)ceil_div(M, kBM));
    auto kern = linear_tcgen05_kernel<0, 128, true>;
    I2V_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile<128>::kSmemBytes));
    kern<<<grid, kThreads, Tile<128>::kSmemBytes, stream>>>(map_x, map_w, bias, y, M, out_channels, K, ldy,
                                                           out_dtype == I2V_DT_BF16, relu, cg, DropMask{});
    return check_launch("linear_tcgen05_kernel<conv>");
}

extern "C" int i2v_conv2d_nhwc_forward(const void* x, const void* w, const float* bias, void* y, int n, int height, int width,
                                       int channels, int out_channels, int kernel, int stride, int pad, long long ldw,
                                       long long ldy, int out_dtype, int relu, cudaStream_t stream) {
    return conv2d_nhwc_impl(x, w, bias, y, n, height, width, channels, out_channels, kernel, stride, pad, ldw, ldy, out_dtype,
                            relu, false, stream);
}

// The same stride-2 layer on a parity-split activation x [N, 2, 2, H/2, W/2, C] (what i2v_pair_conv1_split_bf16 writes);
// `height` and `width` are those of the whole map.  y as above.
extern "C" int i2v_conv2d_nhwc_split_forward(const void* x, const void* w, const float* bias, void* y, int n, int height,
                                             int width, int channels, int out_channels, int kernel, int pad, long long ldw,
                                             long long ldy, int out_dtype, int relu, cudaStream_t stream) {
    return conv2d_nhwc_impl(x, w, bias, y, n, height, width, channels, out_channels, kernel, 2, pad, ldw, ldy, out_dtype, relu,
                            true, stream);
}
