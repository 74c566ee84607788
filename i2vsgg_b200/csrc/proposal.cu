// RPN proposal path for sm_100a: anchor decode + clip, per-frame descending sort, greedy NMS, pad/pack.
//
// Reference semantics:
//   decode + clip   lib/model/rpn/proposal_layer.py:67,81-111 + lib/model/rpn/bbox_transform.py:77-103,125-133
//   sort / top-N    lib/model/rpn/proposal_layer.py:127,138-144
//   NMS             lib/model/nms/nms_wrapper.py:13-21 -> lib/model/nms/nms_cpu.py:6-34 (fp32, no FMA, `ovr <= thresh`)
//   pad / pack      lib/model/rpn/proposal_layer.py:129,153-161
//
// This file is compiled with -fmad=false and uses the _rn intrinsics on every operation whose rounding decides an
// integer result (sort order, keep list): the keep lists are bit-exact against the numpy reference.
//
// Kernel chain (one launch each, all frames of the batch at once):
//   proposal_decode_kernel   one thread per anchor
//   segment_sort_kernel      one CTA per frame: stable LSD radix sort (4 x 8 bit) of (score, index), warp-private
//                            histograms ranked with match.any so no atomics are needed
//   nms_scan_kernel          one CTA per frame: candidates stream through in chunks of 1024; a chunk is first tested
//                            against the boxes kept so far (parallel), then resolved warp by warp with a 32x32
//                            warp-ballot IoU bitmask and a register-resident greedy scan; stops as soon as
//                            post_nms_topN boxes are kept and writes the padded [post,5] rows itself.
#include <limits.h>

#include "common.cuh"

namespace i2v {

// ------------------------------------------------------------------------------------------ decode
// bbox_transform.py:77-103 with every operation rounded separately; exp evaluated in double and rounded once
// (CUDA's double exp is < 1 ulp in double, so the float result is the correctly rounded one, which is what the
// oracle computes with glibc).
// The RPN head's 2-way softmax (rpn.py:63-69: scores [B,2A,H,W] viewed as [B,2,A*H,W], softmax over dim 1): channel a
// is an anchor's background score, channel a + A its foreground score.  Every step is one rounded fp32 operation and the
// exponentials are the correctly rounded ones (double exp rounded once), so the oracle reproduces the bits.
__device__ __forceinline__ void softmax2(float s_bg, float s_fg, float& p_bg, float& p_fg) {
    const float m = fmaxf(s_bg, s_fg);
    const float e_bg = (float)exp((double)__fsub_rn(s_bg, m)), e_fg = (float)exp((double)__fsub_rn(s_fg, m));
    const float sum = __fadd_rn(e_bg, e_fg);
    p_bg = __fdiv_rn(e_bg, sum);
    p_fg = __fdiv_rn(e_fg, sum);
}

__global__ void __launch_bounds__(256) rpn_cls_prob_kernel(const float* __restrict__ score, float* __restrict__ prob,
                                                           int B, int A, int HW) {
    const int64_t total = (int64_t)B * A * HW;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = idx / ((int64_t)A * HW), r = idx - b * (int64_t)A * HW;      // r = a * HW + pixel
        const size_t bg = (size_t)b * 2 * A * HW + (size_t)r, fg = bg + (size_t)A * HW;
        float p0, p1;
        softmax2(__ldg(score + bg), __ldg(score + fg), p0, p1);
        prob[bg] = p0;
        prob[fg] = p1;
    }
}

// SCORES: `cls_prob` holds the raw RPN scores and the foreground probability is formed here (rpn.py:63-69 fused in)
template <bool SCORES>
__global__ void __launch_bounds__(256) proposal_decode_kernel(const float* __restrict__ cls_prob,
                                                              const float* __restrict__ bbox_pred,
                                                              const float* __restrict__ im_info,
                                                              const float* __restrict__ base_anchors, int B, int A,
                                                              int H, int W, int feat_stride, float* __restrict__ boxes,
                                                              float* __restrict__ scores) {
    const int HW = H * W;
    const int64_t total = (int64_t)B * A * HW;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        // threads run over pixels fastest so the NCHW reads coalesce; the anchor-major writes are 16-byte rows
        int pix = (int)(idx % HW);
        int a = (int)((idx / HW) % A);
        int b = (int)(idx / ((int64_t)HW * A));
        int y = pix / W, x = pix - y * W;
        float sx = (float)(x * feat_stride), sy = (float)(y * feat_stride);
        float ax1 = __fadd_rn(__ldg(base_anchors + a * 4 + 0), sx), ay1 = __fadd_rn(__ldg(base_anchors + a * 4 + 1), sy);
        float ax2 = __fadd_rn(__ldg(base_anchors + a * 4 + 2), sx), ay2 = __fadd_rn(__ldg(base_anchors + a * 4 + 3), sy);
        const float* d = bbox_pred + ((size_t)b * 4 * A + 4 * a) * HW + pix;
        float dx = __ldg(d), dy = __ldg(d + HW), dw = __ldg(d + 2 * (size_t)HW), dh = __ldg(d + 3 * (size_t)HW);
        float w = __fadd_rn(__fsub_rn(ax2, ax1), 1.0f), h = __fadd_rn(__fsub_rn(ay2, ay1), 1.0f);
        float cx = __fadd_rn(ax1, __fmul_rn(0.5f, w)), cy = __fadd_rn(ay1, __fmul_rn(0.5f, h));
        float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
        float pw = __fmul_rn((float)exp((double)dw), w), ph = __fmul_rn((float)exp((double)dh), h);
        float x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw)), y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
        float x2 = __fadd_rn(pcx, __fmul_rn(0.5f, pw)), y2 = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
        float xmax = __fsub_rn(__ldg(im_info + b * 3 + 1), 1.f), ymax = __fsub_rn(__ldg(im_info + b * 3 + 0), 1.f);
        size_t j = (size_t)b * HW * A + (size_t)pix * A + a;
        if (boxes) {
            float4 o;
            o.x = fminf(fmaxf(x1, 0.f), xmax);
            o.y = fminf(fmaxf(y1, 0.f), ymax);
            o.z = fminf(fmaxf(x2, 0.f), xmax);
            o.w = fminf(fmaxf(y2, 0.f), ymax);
            reinterpret_cast<float4*>(boxes)[j] = o;
        }
        if (scores) {
            const float* fg = cls_prob + ((size_t)b * 2 * A + A + a) * HW + pix;
            if (SCORES) {
                float p0, p1;
                softmax2(__ldg(fg - (size_t)A * HW), __ldg(fg), p0, p1);
                scores[j] = p1;
            } else {
                scores[j] = __ldg(fg);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ segmented sort
constexpr int kSortThreads = 1024;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortBatch = 8;  // groups of 32 keys in flight per warp

// Monotone map float -> uint32 whose ASCENDING order is the DESCENDING order of the floats.
__device__ __forceinline__ unsigned desc_key(float f) {
    unsigned u = __float_as_uint(f);
    if (u == 0x80000000u) u = 0u;  // -0 sorts as +0 (they compare equal; index order breaks the tie)
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending-orderable
    return ~u;
}

// One CTA per segment.  keys: [segments][n] floats (stride key_stride).  Result: order_out[seg][i] = index of the
// i-th largest key, ties by lower index (stable).  buf_k / buf_i: two ping-pong arrays of n uint32 per segment each.
__global__ void __launch_bounds__(kSortThreads) segment_sort_kernel(const float* __restrict__ keys, int n,
                                                                    int key_stride, int64_t seg_stride,
                                                                    unsigned* buf_k, unsigned* buf_i,
                                                                    int* __restrict__ order_out, int64_t order_stride,
                                                                    const int* __restrict__ gate) {
    __shared__ unsigned hist[kSortWarps][256];
    __shared__ unsigned digit_base[256];
    const int seg = blockIdx.x;
    if (gate && gate[seg] == 0) return;  // the short path already finished this segment
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* kin = keys + (size_t)seg * seg_stride;
    unsigned* k0 = buf_k + (size_t)seg * 2 * n;
    unsigned* k1 = k0 + n;
    unsigned* i0 = buf_i + (size_t)seg * 2 * n;
    unsigned* i1 = i0 + n;
    int* out = order_out + (size_t)seg * order_stride;
    // contiguous slice of the input per warp, walked 32 items at a time in index order
    const int per = ceil_div(ceil_div(n, kSortWarps), 32) * 32;
    const int lo = min(n, warp * per), hi = min(n, lo + per);

    for (int pass = 0; pass < 4; ++pass) {
        const int shift = pass * 8;
        const unsigned* src_k = (pass & 1) ? k0 : k1;  // pass 0 reads the floats; 1: k0 -> k1; 2: k1 -> k0; 3: k0 -> out
        const unsigned* src_i = (pass & 1) ? i0 : i1;
        unsigned* dst_k = (pass & 1) ? k1 : k0;
        unsigned* dst_i = (pass & 1) ? i1 : i0;
        for (int i = tid; i < kSortWarps * 256; i += kSortThreads) (&hist[0][0])[i] = 0;
        __syncthreads();
        // sweep 1: warp-private digit counts.  Eight groups of 32 keys are fetched at once so that the global-memory
        // latency is paid once per 256 keys, not once per group.
        for (int b0 = lo; b0 < hi; b0 += 32 * kSortBatch) {
            unsigned keyv[kSortBatch];
#pragma unroll
            for (int q = 0; q < kSortBatch; ++q) {
                int i = b0 + q * 32 + lane;
                keyv[q] = (i < hi) ? (pass == 0 ? desc_key(kin[(size_t)i * key_stride]) : src_k[i]) : 0u;
            }
#pragma unroll
            for (int q = 0; q < kSortBatch; ++q) {
                if (b0 + q * 32 >= hi) break;
                bool ok = b0 + q * 32 + lane < hi;
                unsigned d = ok ? ((keyv[q] >> shift) & 255u) : (256u + lane);
                unsigned peers = __match_any_sync(0xffffffffu, d);
                if (ok && lane == (__ffs(peers) - 1)) hist[warp][d] += __popc(peers);
                __syncwarp();
            }
        }
        __syncthreads();
        // offsets: digit-major, then warp
        if (tid < 256) {
            unsigned tot = 0;
            for (int w = 0; w < kSortWarps; ++w) tot += hist[w][tid];
            digit_base[tid] = tot;
        }
        __syncthreads();
        if (warp == 0) {  // exclusive scan of 256 totals: 8 per lane
            unsigned v[8], sum = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                v[q] = digit_base[lane * 8 + q];
                sum += v[q];
            }
            unsigned incl = sum;
            for (int o = 1; o < 32; o <<= 1) {
                unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            unsigned run = incl - sum;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                digit_base[lane * 8 + q] = run;
                run += v[q];
            }
        }
        __syncthreads();
        if (tid < 256) {
            unsigned run = digit_base[tid];
            for (int w = 0; w < kSortWarps; ++w) {
                unsigned c = hist[w][tid];
                hist[w][tid] = run;
                run += c;
            }
        }
        __syncthreads();
        // sweep 2: stable scatter (same batching)
        for (int b0 = lo; b0 < hi; b0 += 32 * kSortBatch) {
            unsigned keyv[kSortBatch], idxv[kSortBatch];
#pragma unroll
            for (int q = 0; q < kSortBatch; ++q) {
                int i = b0 + q * 32 + lane;
                bool ok = i < hi;
                keyv[q] = ok ? (pass == 0 ? desc_key(kin[(size_t)i * key_stride]) : src_k[i]) : 0u;
                idxv[q] = ok ? (pass == 0 ? (unsigned)i : src_i[i]) : 0u;
            }
#pragma unroll
            for (int q = 0; q < kSortBatch; ++q) {
                if (b0 + q * 32 >= hi) break;
                bool ok = b0 + q * 32 + lane < hi;
                unsigned d = ok ? ((keyv[q] >> shift) & 255u) : (256u + lane);
                unsigned peers = __match_any_sync(0xffffffffu, d);
                unsigned pos = 0;
                if (ok) pos = hist[warp][d] + __popc(peers & ((1u << lane) - 1u));
                __syncwarp();
                if (ok && lane == (__ffs(peers) - 1)) hist[warp][d] += __popc(peers);
                __syncwarp();
                if (ok) {
                    if (pass == 3) {
                        out[pos] = (int)idxv[q];
                    } else {
                        dst_k[pos] = keyv[q];
                        dst_i[pos] = idxv[q];
                    }
                }
            }
        }
        __syncthreads();  // global writes of this pass are visible to the whole CTA before the next pass reads them
    }
}


// ------------------------------------------------------------------------------------------ top-K order
// NMS with a small post_nms_topN stops after a few thousand candidates (2.2-2.9 K on the bench workload), so ordering all
// K*A = 21.5 K scores of a frame is mostly wasted.  These CTAs order only the kTopK best, kTopKPart ranks per CTA and with
// no communication between the CTAs of a frame: CTA c finds, by radix SELECT over the composite (desc_key, index) --
// unique, so ties need no special case -- the composites of rank c*kTopKPart and (c+1)*kTopKPart (16-bit digits, one
// counting sweep per digit, the first sweep shared by both ranks), compacts what lies between them into registers, one
// composite per thread, and orders them with a bitonic network (shuffles below distance 32, shared memory above); the low
// word of a composite is the row.  The result equals the first kTopK entries of the stable full sort bit for bit.  If NMS
// runs out of these candidates before it has post_nms_topN boxes it raises the frame's retry flag and the full sort + a
// second NMS run, gated on that flag, redo the frame (rare; see proposal_forward_impl).
constexpr int kTopK = 4096;
constexpr int kTopKPart = 1024;
constexpr int kTopKThreads = 1024;
constexpr int kTopKBatch = 8;
constexpr int kTopKHistWords = 32768;  // 65536 bins of 16 bits, two per word: needs n <= 65535
constexpr size_t kTopKSmemBytes = (size_t)kTopKHistWords * 4 + (size_t)2 * kTopKPart * 8;
static_assert(kTopKPart == kTopKThreads, "one composite per thread");

__global__ void __launch_bounds__(kTopKThreads, 1) segment_topk_kernel(const float* __restrict__ keys, int n, int key_stride,
                                                                       int64_t seg_stride, int* __restrict__ order_out,
                                                                       int64_t order_stride, int need) {
    extern __shared__ __align__(16) unsigned char tk_smem[];
    unsigned* hist = reinterpret_cast<unsigned*>(tk_smem);
    unsigned long long* xch = reinterpret_cast<unsigned long long*>(tk_smem + (size_t)kTopKHistWords * 4);   // [2][kTopKPart]
    __shared__ unsigned s_wtot[32];
    __shared__ unsigned s_owner, s_excl, s_digit, s_krem, s_binc, s_cnt;

    const int seg = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank_lo = (int)blockIdx.x * kTopKPart;               // this CTA orders ranks (rank_lo, rank_hi], 1-based
    const int rank_hi = min(min(n, need), rank_lo + kTopKPart);
    if (rank_hi <= rank_lo) return;
    const float* kin = keys + (size_t)seg * seg_stride;
    int* out = order_out + (size_t)seg * order_stride;

    // one counting sweep: histogram of a 16-bit digit over the keys that match the digits fixed so far
    auto sweep = [&](int pass, unsigned p0, unsigned p1) {
        for (int i = tid; i < kTopKHistWords / 4; i += kTopKThreads) reinterpret_cast<uint4*>(hist)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        // eight loads in flight per thread: the sweep pays the L2 latency once per 8 K keys, not once per 1 K
        for (int base = tid; base < n; base += kTopKBatch * kTopKThreads) {
            unsigned kv[kTopKBatch];
#pragma unroll
            for (int q = 0; q < kTopKBatch; ++q) {
                const int i = base + q * kTopKThreads;
                kv[q] = i < n ? desc_key(kin[(size_t)i * key_stride]) : 0u;
            }
#pragma unroll
            for (int q = 0; q < kTopKBatch; ++q) {
                const int i = base + q * kTopKThreads;
                const unsigned kk = kv[q];
                unsigned d;
                bool match;
                if (pass == 0) {
                    d = kk >> 16;
                    match = true;
                } else if (pass == 1) {
                    d = kk & 0xffffu;
                    match = (kk >> 16) == p0;
                } else {
                    d = (unsigned)i;        // n <= 65535: the index is one digit
                    match = kk == ((p0 << 16) | p1);
                }
                if (match && i < n) atomicAdd(&hist[d >> 1], (d & 1u) ? 65536u : 1u);
            }
        }
        __syncthreads();
    };
    // which bin of the current histogram holds its krem-th entry (1-based): digit, rank inside the bin, size of the bin.
    // Thread t owns bins [64t, 64t + 64) (words rotated by t: no bank conflicts).
    auto locate = [&](unsigned krem, unsigned& digit, unsigned& krem_in, unsigned& binc) {
        unsigned sum = 0;
#pragma unroll 8
        for (int w = 0; w < 32; ++w) {
            const unsigned v = hist[32 * tid + ((w + tid) & 31)];
            sum += (v & 0xffffu) + (v >> 16);
        }
        unsigned incl = sum;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wtot[warp] = incl;
        __syncthreads();
        unsigned before = 0;
        for (int w = 0; w < warp; ++w) before += s_wtot[w];
        const unsigned excl = before + incl - sum;
        if (excl < krem && krem <= excl + sum) {
            s_owner = (unsigned)tid;
            s_excl = excl;
        }
        __syncthreads();
        if (warp == 0) {
            const unsigned word = 32u * s_owner + (unsigned)lane;
            const unsigned v = hist[word], c0 = v & 0xffffu, c1 = v >> 16, sw = c0 + c1;
            unsigned in2 = sw;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, in2, o);
                if (lane >= o) in2 += t;
            }
            const unsigned e = s_excl + in2 - sw;
            if (e < krem && krem <= e + sw) {
                const unsigned r = krem - e;
                if (r <= c0) {
                    s_digit = 2u * word;
                    s_krem = r;
                    s_binc = c0;
                } else {
                    s_digit = 2u * word + 1u;
                    s_krem = r - c0;
                    s_binc = c1;
                }
            }
        }
        __syncthreads();
        digit = s_digit;
        krem_in = s_krem;
        binc = s_binc;
        __syncthreads();                                           // s_* are rewritten by the next call
    };

    // the composites of rank rank_hi and rank_lo; a bin that is wanted whole ends the descent (bound = its upper edge)
    unsigned long long t_hi = ~0ull, t_lo = 0ull;
    const bool want_hi = rank_hi < n, want_lo = rank_lo > 0;       // otherwise: everything up to the end / from the start
    if (want_hi || want_lo) {
        sweep(0, 0, 0);
        unsigned dh = 0, kh = 0, bh = 0, dl = 0, kl = 0, bl = 0;
        if (want_hi) locate((unsigned)rank_hi, dh, kh, bh);
        if (want_lo) locate((unsigned)rank_lo, dl, kl, bl);
        unsigned long long* bound[2] = {&t_hi, &t_lo};
        const bool want[2] = {want_hi, want_lo};
        const unsigned d0[2] = {dh, dl}, k0[2] = {kh, kl}, b0[2] = {bh, bl};
        unsigned swept_p0 = 0xffffffffu;                           // the pass-1 histogram in shared memory belongs to this digit
        for (int w = 0; w < 2; ++w) {
            if (!want[w]) continue;
            const unsigned p0 = d0[w];
            unsigned krem = k0[w];
            *bound[w] = ((unsigned long long)p0 << 48) | 0xffffffffffffull;
            if (b0[w] == krem) continue;
            if (swept_p0 != p0) sweep(1, p0, 0);
            swept_p0 = p0;
            unsigned p1, binc;
            locate(krem, p1, krem, binc);
            *bound[w] = ((unsigned long long)((p0 << 16) | p1) << 32) | 0xffffffffull;
            if (binc == krem) continue;
            sweep(2, p0, p1);
            swept_p0 = 0xffffffffu;
            unsigned di;
            locate(krem, di, krem, binc);
            *bound[w] = ((unsigned long long)((p0 << 16) | p1) << 32) | (unsigned long long)di;
        }
    }

    // compaction into shared memory (unordered: the network orders them, and they are unique)
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    for (int base = 0; base < n; base += kTopKBatch * kTopKThreads) {
        unsigned kv[kTopKBatch];
#pragma unroll
        for (int q = 0; q < kTopKBatch; ++q) {
            const int i = base + q * kTopKThreads + tid;
            kv[q] = i < n ? desc_key(kin[(size_t)i * key_stride]) : 0u;
        }
#pragma unroll
        for (int q = 0; q < kTopKBatch; ++q) {
            const int i = base + q * kTopKThreads + tid;
            const unsigned long long c = ((unsigned long long)kv[q] << 32) | (unsigned)i;
            const bool sel = i < n && c <= t_hi && (!want_lo || c > t_lo);
            const unsigned m = __ballot_sync(0xffffffffu, sel);
            if (m) {
                const int leader = __ffs(m) - 1;
                unsigned b = 0;
                if (lane == leader) b = atomicAdd(&s_cnt, (unsigned)__popc(m));
                b = __shfl_sync(0xffffffffu, b, leader);
                if (sel) xch[b + __popc(m & ((1u << lane) - 1u))] = c;
            }
        }
    }
    __syncthreads();
    const int cnt = (int)s_cnt;                                    // rank_hi - rank_lo
    unsigned long long v = tid < cnt ? xch[tid] : ~0ull;
    __syncthreads();

    // bitonic network over the CTA's registers, ascending: thread t ends up with the composite of rank rank_lo + t + 1
    int buf = 0;
    for (int k = 2; k <= kTopKPart; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            unsigned long long o;
            if (j < 32) {
                o = __shfl_xor_sync(0xffffffffu, v, j);
            } else {
                unsigned long long* x = xch + buf * kTopKPart;
                x[tid] = v;
                __syncthreads();
                o = x[tid ^ j];
                buf ^= 1;                                          // the next exchange writes the other buffer: one barrier each
            }
            const bool keep_min = ((tid & j) == 0) == ((tid & k) == 0);
            if ((o < v) == keep_min) v = o;
        }
    }
    if (tid < cnt) out[rank_lo + tid] = (int)(unsigned)(v & 0xffffffffull);
}

static bool topk_order_ok(int n) { return n >= 1 && n <= 65535; }

// orders the min(n, need, kTopK) best keys of every segment
static int launch_topk_order(const float* keys, int segments, int n, int key_stride, int64_t seg_stride, int* order,
                             int64_t order_stride, int need, cudaStream_t stream) {
    I2V_CUDA_TRY(cudaFuncSetAttribute(segment_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTopKSmemBytes));
    const int want = min(min(n, need), kTopK);
    dim3 grid(ceil_div(want, kTopKPart), segments);
    segment_topk_kernel<<<grid, kTopKThreads, kTopKSmemBytes, stream>>>(keys, n, key_stride, seg_stride, order, order_stride, want);
    return check_launch("segment_topk_kernel");
}

// ------------------------------------------------------------------------------------------ NMS
constexpr int kNmsThreads = 1024;
constexpr int kNmsWarps = kNmsThreads / 32;
constexpr int kNmsKeptSmem = 4096;  // kept boxes cached in shared memory; later ones spill to the workspace

// nms_cpu.py:14,20-31 -- true iff box b must be removed because of kept box a.
__device__ __forceinline__ bool nms_suppressed(const float4 a, const float aa, const float4 b, const float ab,
                                               const float thr) {
    float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y), xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    float w = fmaxf(0.f, __fadd_rn(__fsub_rn(xx2, xx1), 1.f));
    float h = fmaxf(0.f, __fadd_rn(__fsub_rn(yy2, yy1), 1.f));
    float inter = __fmul_rn(w, h);
    float sum = __fadd_rn(aa, ab);
    if (inter == 0.f && sum > 0.f) return thr < 0.f;  // ovr == 0 exactly: skip the divide
    float ovr = __fdiv_rn(inter, __fsub_rn(sum, inter));
    return !(ovr <= thr);
}

struct NmsArgs {
    const float* boxes;      // [sets][rows][box_stride]
    int64_t set_stride;      // floats between sets
    int box_stride;
    const int* order;        // optional [sets][order_stride]: candidate i is row order[i]
    int64_t order_stride;
    int n;                   // candidates per set
    float thresh;
    int max_keep;            // <= 0: unlimited
    int* keep_out;           // optional [sets][keep_stride]
    int keep_stride;
    int emit_rows;           // keep_out holds row numbers (order[i]) instead of candidate ranks
    int* num_out;            // optional [sets]
    float4* spill_box;       // [sets][n]
    float* spill_area;       // [sets][n]
    float* out_rois;         // optional [sets][post][5]
    int post;
    int frame_base;          // column 0 of out_rois = frame_base + set (a chunk of a larger batch keeps the batch's numbering)
    const int* gate;         // optional [sets]: a set whose entry is 0 is skipped (its result is already final)
    int* retry;              // optional [sets]: set to 1 when the candidates ran out before max_keep boxes were kept and
    int more;                //   `more` says that the caller holds further candidates (the order was only a prefix)
};

__global__ void __launch_bounds__(kNmsThreads, 1) nms_scan_kernel(const NmsArgs a) {
    extern __shared__ __align__(16) unsigned char nms_smem[];
    float4* kbox = reinterpret_cast<float4*>(nms_smem);          // [kNmsKeptSmem]
    float4* cbox = kbox + kNmsKeptSmem;                          // [kNmsThreads]
    float4* nbox = cbox + kNmsThreads;                           // [2][32]
    float* karea = reinterpret_cast<float*>(nbox + 64);          // [kNmsKeptSmem]
    float* carea = karea + kNmsKeptSmem;                         // [kNmsThreads]
    float* narea = carea + kNmsThreads;                          // [2][32]
    __shared__ int s_nk;
    __shared__ int s_ncnt[2];
    __shared__ int s_nkv[2];   // the running count as published in turn w (double-buffered like s_ncnt: warp w + 1 may
                               // already be writing s_nk for its own turn while slower warps still test turn w's value)

    const int set = blockIdx.x;
    if (a.gate && a.gate[set] == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* boxes = a.boxes + (size_t)set * a.set_stride;
    const int* order = a.order ? a.order + (size_t)set * a.order_stride : nullptr;
    float4* sp_box = a.spill_box + (size_t)set * a.n;
    float* sp_area = a.spill_area + (size_t)set * a.n;
    int* keep_out = a.keep_out ? a.keep_out + (size_t)set * a.keep_stride : nullptr;
    const float thr = a.thresh;
    const int limit = a.max_keep > 0 ? a.max_keep : INT_MAX;

    __shared__ unsigned s_live[2];   // per chunk (double-buffered): bit w <=> warp w still has a candidate after step (a)
    if (tid == 0) {
        s_nk = 0;
        s_live[0] = s_live[1] = 0u;
    }
    __syncthreads();
    bool done = (a.n == 0);

    for (int base = 0, chunk = 0; base < a.n && !done; base += kNmsThreads, ++chunk) {
        const int i = base + tid;
        const bool have = i < a.n;
        int row = 0;
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        float area = 0.f;
        if (have) {
            row = order ? order[i] : i;
            const float* p = boxes + (size_t)row * a.box_stride;
            box = make_float4(p[0], p[1], p[2], p[3]);
            area = __fmul_rn(__fadd_rn(__fsub_rn(box.z, box.x), 1.f), __fadd_rn(__fsub_rn(box.w, box.y), 1.f));
        }
        cbox[tid] = box;
        carea[tid] = area;
        bool alive = have;
        const int nk0 = s_nk;
        // (a) against everything kept before this chunk
        for (int k = 0; k < nk0; ++k) {
            if (__ballot_sync(0xffffffffu, alive) == 0u) break;
            float4 kb;
            float ka;
            if (k < kNmsKeptSmem) {
                kb = kbox[k];
                ka = karea[k];
            } else {
                kb = sp_box[k];
                ka = sp_area[k];
            }
            if (alive && nms_suppressed(kb, ka, box, area, thr)) alive = false;
        }
        const unsigned am0 = __ballot_sync(0xffffffffu, alive);
        if (lane == 0 && am0) atomicOr(&s_live[chunk & 1], 1u << warp);
        __syncthreads();
        const unsigned live = s_live[chunk & 1];
        if (tid == 0) s_live[(chunk + 1) & 1] = 0u;        // nobody touches the other buffer before the chunk ends
        // (b) this warp's 32x32 suppression bitmask: bit j of `mask` <=> this lane's box removes lane j's (j > lane);
        //     only candidates that are still there matter
        unsigned mask = 0;
        for (unsigned todo = am0; todo; todo &= todo - 1u) {
            const int j = __ffs(todo) - 1;
            float4 ob = cbox[warp * 32 + j];
            float oa = carea[warp * 32 + j];
            if (alive && j > lane && nms_suppressed(box, area, ob, oa, thr)) mask |= 1u << j;
        }
        // (c) warps take their turn in candidate order; a warp with nothing left after (a) has no turn
        int turn = 0;
        for (int w = 0; w < kNmsWarps; ++w) {
            if (!((live >> w) & 1u)) continue;
            const int pb = turn++ & 1;
            if (warp == w) {
                const int nk = s_nk;
                unsigned rem = __ballot_sync(0xffffffffu, alive), keepm = 0;
                while (rem) {                              // one step per KEPT box: the lowest survivor stays, its mask goes
                    const int l = __ffs(rem) - 1;
                    keepm |= 1u << l;
                    rem &= ~(__shfl_sync(0xffffffffu, mask, l) | (1u << l));
                }
                int cnt = __popc(keepm);
                const int room = limit - nk;
                while (cnt > room) {  // drop the last survivors beyond the requested count
                    keepm &= ~(1u << (31 - __clz(keepm)));
                    --cnt;
                }
                if ((keepm >> lane) & 1u) {
                    int rank = __popc(keepm & ((1u << lane) - 1u));
                    int slot = nk + rank;
                    if (slot < kNmsKeptSmem) {
                        kbox[slot] = box;
                        karea[slot] = area;
                    } else {
                        sp_box[slot] = box;
                        sp_area[slot] = area;
                    }
                    if (keep_out) keep_out[slot] = a.emit_rows ? row : i;
                    nbox[pb * 32 + rank] = box;
                    narea[pb * 32 + rank] = area;
                }
                if (lane == 0) {
                    s_ncnt[pb] = cnt;
                    s_nkv[pb] = nk + cnt;
                    s_nk = nk + cnt;
                }
            }
            __syncthreads();
            if (s_nkv[pb] >= limit) {
                done = true;
                break;
            }
            const int cnt = s_ncnt[pb];
            if (warp > w && alive) {
                for (int s = 0; s < cnt; ++s) {
                    if (nms_suppressed(nbox[pb * 32 + s], narea[pb * 32 + s], box, area, thr)) {
                        alive = false;
                        break;
                    }
                }
            }
        }
        __syncthreads();
    }

    const int nk = s_nk;
    if (tid == 0 && a.num_out) a.num_out[set] = nk;
    if (tid == 0 && a.retry) a.retry[set] = (a.more && nk < limit) ? 1 : 0;
    if (a.out_rois) {  // proposal_layer.py:129,153-161: [post,5] rows (frame, x1, y1, x2, y2), zero padded
        float* o = a.out_rois + (size_t)set * a.post * 5;
        for (int k = tid; k < a.post; k += kNmsThreads) {
            float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < nk) bx = (k < kNmsKeptSmem) ? kbox[k] : sp_box[k];
            o[k * 5 + 0] = (float)(a.frame_base + set);
            o[k * 5 + 1] = bx.x;
            o[k * 5 + 2] = bx.y;
            o[k * 5 + 3] = bx.z;
            o[k * 5 + 4] = bx.w;
        }
    }
}


constexpr size_t kNmsSmemBytes = (size_t)(kNmsKeptSmem + kNmsThreads + 64) * (sizeof(float4) + sizeof(float));

static int launch_nms(const NmsArgs& a, int sets, cudaStream_t stream) {
    I2V_CUDA_TRY(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNmsSmemBytes));
    nms_scan_kernel<<<sets, kNmsThreads, kNmsSmemBytes, stream>>>(a);
    return check_launch("nms_scan_kernel");
}

struct NmsWs {
    float4* spill_box;
    float* spill_area;
    unsigned* sort_k;
    unsigned* sort_i;
    int* order;
    float* scores;
    float* boxes;
    int* retry;
    size_t bytes;
};
// `sort_n` > 0 adds the sort buffers (sort_n items per set); `decode` adds decoded boxes + scores.
static NmsWs carve_nms_ws(void* ws, int sets, int n, int sort_n, bool decode) {
    Carver cv(ws);
    NmsWs w{};
    w.spill_box = cv.take<float4>((size_t)sets * n);
    w.spill_area = cv.take<float>((size_t)sets * n);
    if (sort_n > 0) {
        w.sort_k = cv.take<unsigned>((size_t)sets * 2 * sort_n);
        w.sort_i = cv.take<unsigned>((size_t)sets * 2 * sort_n);
        w.order = cv.take<int>((size_t)sets * sort_n);
    }
    if (decode) {
        w.scores = cv.take<float>((size_t)sets * sort_n);
        w.boxes = cv.take<float>((size_t)sets * sort_n * 4);
    }
    w.retry = cv.take<int>((size_t)sets);
    w.bytes = cv.used();
    return w;
}

// `gate`: per-segment flags; a segment whose flag is 0 is left alone
static int launch_sort(const float* keys, int segments, int n, int key_stride, int64_t seg_stride, const NmsWs& w,
                       cudaStream_t stream, const int* gate = nullptr) {
    segment_sort_kernel<<<segments, kSortThreads, 0, stream>>>(keys, n, key_stride, seg_stride, w.sort_k, w.sort_i,
                                                              w.order, n, gate);
    return check_launch("segment_sort_kernel");
}

}  // namespace i2v

using namespace i2v;

extern "C" size_t i2v_nms_workspace_bytes(int batch, int num_boxes) {
    if (batch < 0 || num_boxes < 0) return 0;
    return carve_nms_ws(nullptr, batch, num_boxes, 0, false).bytes;
}

extern "C" int i2v_nms_sorted(const float* boxes, int batch, int num_boxes, int box_stride, float thresh, int max_keep,
                              int* keep_out, int keep_stride, int* num_out, void* workspace, size_t workspace_bytes,
                              cudaStream_t stream) {
    I2V_REQUIRE(batch >= 0 && num_boxes >= 0 && box_stride >= 4, "nms_sorted: bad size");
    if (batch == 0) return I2V_OK;
    I2V_REQUIRE(num_boxes == 0 || boxes, "nms_sorted: null boxes");
    I2V_REQUIRE(!keep_out || keep_stride >= (max_keep > 0 && max_keep < num_boxes ? max_keep : num_boxes),
                "nms_sorted: keep_stride too small");
    NmsWs w = carve_nms_ws(workspace, batch, num_boxes, 0, false);
    if (num_boxes > 0 && (!workspace || workspace_bytes < w.bytes)) {
        set_error("nms_sorted: workspace %zu < %zu bytes", workspace_bytes, w.bytes);
        return I2V_ERR_WORKSPACE;
    }
    NmsArgs a{};
    a.boxes = boxes;
    a.set_stride = (int64_t)num_boxes * box_stride;
    a.box_stride = box_stride;
    a.n = num_boxes;
    a.thresh = thresh;
    a.max_keep = max_keep;
    a.keep_out = keep_out;
    a.keep_stride = keep_stride;
    a.num_out = num_out;
    a.spill_box = w.spill_box;
    a.spill_area = w.spill_area;
    return launch_nms(a, batch, stream);
}

extern "C" size_t i2v_nms_dets_workspace_bytes(int num_boxes) {
    if (num_boxes < 0) return 0;
    return carve_nms_ws(nullptr, 1, num_boxes, num_boxes, false).bytes;
}

extern "C" int i2v_nms_dets(const float* dets, int num_boxes, float thresh, int* keep_out, int* num_out,
                            void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    I2V_REQUIRE(num_boxes >= 0, "nms_dets: bad size");
    I2V_REQUIRE(num_out, "nms_dets: null num_out");
    if (num_boxes == 0) {
        I2V_CUDA_TRY(cudaMemsetAsync(num_out, 0, sizeof(int), stream));
        return I2V_OK;
    }
    I2V_REQUIRE(dets && keep_out, "nms_dets: null pointer");
    NmsWs w = carve_nms_ws(workspace, 1, num_boxes, num_boxes, false);
    if (!workspace || workspace_bytes < w.bytes) {
        set_error("nms_dets: workspace %zu < %zu bytes", workspace_bytes, w.bytes);
        return I2V_ERR_WORKSPACE;
    }
    if (num_boxes <= kTopK)   // the whole order in one short kernel
        I2V_TRY(launch_topk_order(dets + 4, 1, num_boxes, 5, 0, w.order, num_boxes, num_boxes, stream));
    else
        I2V_TRY(launch_sort(dets + 4, 1, num_boxes, 5, 0, w, stream));
    NmsArgs a{};
    a.boxes = dets;
    a.box_stride = 5;
    a.order = w.order;
    a.order_stride = num_boxes;
    a.n = num_boxes;
    a.thresh = thresh;
    a.keep_out = keep_out;
    a.keep_stride = num_boxes;
    a.emit_rows = 1;
    a.num_out = num_out;
    a.spill_box = w.spill_box;
    a.spill_area = w.spill_area;
    return launch_nms(a, 1, stream);
}

extern "C" size_t i2v_proposal_workspace_bytes(int batch, int num_anchors, int height, int width, int pre_nms_top_n) {
    if (batch < 0 || num_anchors < 0 || height < 0 || width < 0) return 0;
    int64_t ka = (int64_t)num_anchors * height * width;
    if (ka > INT_MAX) return 0;
    (void)pre_nms_top_n;
    return carve_nms_ws(nullptr, batch, (int)ka, (int)ka, true).bytes;
}

static int proposal_common(const float* cls_prob, const float* bbox_pred, const float* im_info,
                           const float* base_anchors, int batch, int num_anchors, int height, int width,
                           int feat_stride, void* workspace, size_t workspace_bytes, NmsWs& w, int& ka) {
    I2V_REQUIRE(batch >= 0 && num_anchors >= 1 && height >= 1 && width >= 1, "proposal: bad size");
    int64_t ka64 = (int64_t)num_anchors * height * width;
    I2V_REQUIRE(ka64 <= (1 << 24), "proposal: %lld anchors per frame is more than this kernel chain supports", (long long)ka64);
    ka = (int)ka64;
    if (batch == 0) return I2V_OK;
    I2V_REQUIRE(cls_prob && bbox_pred && im_info && base_anchors, "proposal: null pointer");
    w = carve_nms_ws(workspace, batch, ka, ka, true);
    if (!workspace || workspace_bytes < w.bytes) {
        set_error("proposal: workspace %zu < %zu bytes", workspace_bytes, w.bytes);
        return I2V_ERR_WORKSPACE;
    }
    (void)feat_stride;
    return I2V_OK;
}

static int proposal_forward_impl(bool from_scores, const float* cls_prob, const float* bbox_pred, const float* im_info,
                                 const float* base_anchors, int batch, int num_anchors, int height, int width,
                                 int feat_stride, int pre_nms_top_n, int post_nms_top_n, float nms_thresh,
                                 float* out_rois, int* out_counts, void* workspace, size_t workspace_bytes,
                                 cudaStream_t stream, int frame_base = 0) {
    NmsWs w;
    int ka;
    I2V_TRY(proposal_common(cls_prob, bbox_pred, im_info, base_anchors, batch, num_anchors, height, width, feat_stride,
                            workspace, workspace_bytes, w, ka));
    I2V_REQUIRE(post_nms_top_n >= 1, "proposal_forward: post_nms_top_n must be positive");
    if (batch == 0) return I2V_OK;
    I2V_REQUIRE(out_rois, "proposal_forward: null out_rois");
    int64_t total = (int64_t)batch * ka;
    if (from_scores)
        proposal_decode_kernel<true><<<grid_for(total, 256), 256, 0, stream>>>(cls_prob, bbox_pred, im_info, base_anchors, batch,
                                                                               num_anchors, height, width, feat_stride, w.boxes, w.scores);
    else
        proposal_decode_kernel<false><<<grid_for(total, 256), 256, 0, stream>>>(cls_prob, bbox_pred, im_info, base_anchors, batch,
                                                                                num_anchors, height, width, feat_stride, w.boxes, w.scores);
    I2V_TRY(check_launch("proposal_decode_kernel"));
    // proposal_layer.py:140-141: keep the pre_nms_topN best (the guard compares against the batch-wide numel, which
    // for any batch reduces to min(pre_nms_topN, K*A) per frame)
    const int n_cand = (pre_nms_top_n > 0 && pre_nms_top_n < ka) ? pre_nms_top_n : ka;
    // Short path: order only the kTopK best scores of each frame and run NMS on those.  It is complete when they are all
    // the candidates there are, and otherwise whenever NMS keeps post_nms_topN boxes before they run out (test mode: 300
    // boxes are found within the first ~3 K candidates); a frame where it does not raises its retry flag and is redone
    // by the full sort + NMS below, which skip every other frame.
    static const bool no_short = getenv("I2V_PROPOSAL_FULL_SORT") != nullptr;
    const bool short_path = !no_short && topk_order_ok(ka) && (n_cand <= kTopK || post_nms_top_n * 8 <= kTopK);
    NmsArgs a{};
    a.boxes = w.boxes;
    a.set_stride = (int64_t)ka * 4;
    a.box_stride = 4;
    a.order = w.order;
    a.order_stride = ka;
    a.thresh = nms_thresh;
    a.max_keep = post_nms_top_n;
    a.num_out = out_counts;
    a.spill_box = w.spill_box;
    a.spill_area = w.spill_area;
    a.out_rois = out_rois;
    a.post = post_nms_top_n;
    a.frame_base = frame_base;
    // the spill arrays are indexed by kept slot < a.n <= ka: carved for ka per frame
    if (short_path) {
        I2V_TRY(launch_topk_order(w.scores, batch, ka, 1, ka, w.order, ka, n_cand, stream));
        a.n = n_cand < kTopK ? n_cand : kTopK;
        if (n_cand <= kTopK) return launch_nms(a, batch, stream);
        a.more = 1;
        a.retry = w.retry;
        I2V_TRY(launch_nms(a, batch, stream));
        a.more = 0;
        a.retry = nullptr;
        a.gate = w.retry;
        I2V_TRY(launch_sort(w.scores, batch, ka, 1, ka, w, stream, w.retry));
    } else {
        I2V_TRY(launch_sort(w.scores, batch, ka, 1, ka, w, stream));
    }
    a.n = n_cand;
    return launch_nms(a, batch, stream);
}

extern "C" int i2v_proposal_forward(const float* cls_prob, const float* bbox_pred, const float* im_info,
                                    const float* base_anchors, int batch, int num_anchors, int height, int width,
                                    int feat_stride, int pre_nms_top_n, int post_nms_top_n, float nms_thresh,
                                    float* out_rois, int* out_counts, void* workspace, size_t workspace_bytes,
                                    cudaStream_t stream) {
    return proposal_forward_impl(false, cls_prob, bbox_pred, im_info, base_anchors, batch, num_anchors, height, width,
                                 feat_stride, pre_nms_top_n, post_nms_top_n, nms_thresh, out_rois, out_counts, workspace,
                                 workspace_bytes, stream);
}

// A chunk of a larger batch: the frames are numbered frame_base, frame_base + 1, ... in column 0 of out_rois
// (proposal_layer.py:160 writes the index inside the batch the layer was called with).
extern "C" int i2v_proposal_forward_chunk(const float* cls_prob, const float* bbox_pred, const float* im_info,
                                          const float* base_anchors, int batch, int num_anchors, int height, int width,
                                          int feat_stride, int pre_nms_top_n, int post_nms_top_n, float nms_thresh,
                                          int frame_base, float* out_rois, int* out_counts, void* workspace,
                                          size_t workspace_bytes, cudaStream_t stream) {
    return proposal_forward_impl(false, cls_prob, bbox_pred, im_info, base_anchors, batch, num_anchors, height, width,
                                 feat_stride, pre_nms_top_n, post_nms_top_n, nms_thresh, out_rois, out_counts, workspace,
                                 workspace_bytes, stream, frame_base);
}

// rois[n, 0] += offset: a chunk that was processed with chunk-local frame numbers (what the RoI kernels of the chunk need)
// gets the numbering of the whole batch back before it leaves the device.
__global__ void rois_add_frame_kernel(float* __restrict__ rois, int num_rois, float offset) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < num_rois) rois[(size_t)n * 5] += offset;
}
extern "C" int i2v_rois_add_frame(float* rois, int num_rois, int offset, cudaStream_t stream) {
    I2V_REQUIRE(num_rois >= 0, "rois_add_frame: bad size");
    if (num_rois == 0 || offset == 0) return I2V_OK;
    I2V_REQUIRE(rois, "rois_add_frame: null pointer");
    rois_add_frame_kernel<<<ceil_div(num_rois, 256), 256, 0, stream>>>(rois, num_rois, (float)offset);
    return check_launch("rois_add_frame_kernel");
}

// rpn.py:63-78: the same layer fed with the RPN head's raw class scores; the 2-way softmax and the foreground slice are
// folded into the decode kernel (no [B,2A,H,W] probability tensor is written or re-read)
extern "C" int i2v_proposal_forward_scores(const float* cls_score, const float* bbox_pred, const float* im_info,
                                           const float* base_anchors, int batch, int num_anchors, int height, int width,
                                           int feat_stride, int pre_nms_top_n, int post_nms_top_n, float nms_thresh,
                                           float* out_rois, int* out_counts, void* workspace, size_t workspace_bytes,
                                           cudaStream_t stream) {
    return proposal_forward_impl(true, cls_score, bbox_pred, im_info, base_anchors, batch, num_anchors, height, width,
                                 feat_stride, pre_nms_top_n, post_nms_top_n, nms_thresh, out_rois, out_counts, workspace,
                                 workspace_bytes, stream);
}

// rpn.py:66-68: rpn_cls_prob [B,2A,H,W] from rpn_cls_score [B,2A,H,W] (softmax over (a, a + A) pairs)
extern "C" int i2v_rpn_cls_prob(const float* cls_score, float* cls_prob, int batch, int num_anchors, int height,
                                int width, cudaStream_t stream) {
    I2V_REQUIRE(batch >= 0 && num_anchors >= 1 && height >= 1 && width >= 1, "rpn_cls_prob: bad size");
    if (batch == 0) return I2V_OK;
    I2V_REQUIRE(cls_score && cls_prob, "rpn_cls_prob: null pointer");
    const int64_t total = (int64_t)batch * num_anchors * height * width;
    rpn_cls_prob_kernel<<<grid_for(total, 256), 256, 0, stream>>>(cls_score, cls_prob, batch, num_anchors, height * width);
    return check_launch("rpn_cls_prob_kernel");
}

extern "C" int i2v_proposal_stages(const float* cls_prob, const float* bbox_pred, const float* im_info,
                                   const float* base_anchors, int batch, int num_anchors, int height, int width,
                                   int feat_stride, float* boxes, float* scores, int* order, void* workspace,
                                   size_t workspace_bytes, cudaStream_t stream) {
    NmsWs w;
    int ka;
    I2V_TRY(proposal_common(cls_prob, bbox_pred, im_info, base_anchors, batch, num_anchors, height, width, feat_stride,
                            workspace, workspace_bytes, w, ka));
    if (batch == 0) return I2V_OK;
    int64_t total = (int64_t)batch * ka;
    proposal_decode_kernel<false><<<grid_for(total, 256), 256, 0, stream>>>(cls_prob, bbox_pred, im_info, base_anchors, batch,
                                                                     num_anchors, height, width, feat_stride, w.boxes, w.scores);
    I2V_TRY(check_launch("proposal_decode_kernel"));
    if (boxes) I2V_CUDA_TRY(cudaMemcpyAsync(boxes, w.boxes, sizeof(float) * 4 * (size_t)total, cudaMemcpyDeviceToDevice, stream));
    if (scores) I2V_CUDA_TRY(cudaMemcpyAsync(scores, w.scores, sizeof(float) * (size_t)total, cudaMemcpyDeviceToDevice, stream));
    if (order) {
        I2V_TRY(launch_sort(w.scores, batch, ka, 1, ka, w, stream));
        I2V_CUDA_TRY(cudaMemcpyAsync(order, w.order, sizeof(int) * (size_t)total, cudaMemcpyDeviceToDevice, stream));
    }
    return I2V_OK;
}
