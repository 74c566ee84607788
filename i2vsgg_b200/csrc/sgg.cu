// SGG pair stage and triplet selection for sm_100a.
//
// Reference semantics:
//   pair enumeration  lib/model/faster_rcnn/faster_rcnn_SGG_emb.py:597-606   (all ordered i != j, i-major)
//   union boxes       lib/model/faster_rcnn/resnet_SGG_emb.py:240-244 applied per pair at faster_rcnn_SGG_emb.py:649-653
//   dual masks        lib/model/faster_rcnn/resnet_SGG_emb.py:246-256, stacked at faster_rcnn_SGG_emb.py:654-655
//   triplet top-100   lib/utils.py:609-626
//
// One launch builds everything a frame's pair list needs (indices, union boxes, 32x32 dual masks); one launch
// selects the top-k triplets of a frame (radix select over the descending-order keys in shared-memory histograms,
// ordered tie handling, bitonic sort of the winners, record gather).
#include <algorithm>

#include "common.cuh"

namespace i2v {

// ------------------------------------------------------------------------------------------ pair build
__device__ __forceinline__ void mask_extent(const float* __restrict__ bb, double rh, double rw, int& x1, int& x2,
                                            int& y1, int& y2) {
    // resnet_SGG_emb.py:249-252, evaluated in double like the Python floats
    x1 = max(0, (int)floor((double)bb[0] * rw));
    x2 = min(32, (int)ceil((double)bb[2] * rw));
    y1 = max(0, (int)floor((double)bb[1] * rh));
    y2 = min(32, (int)ceil((double)bb[3] * rh));
}

__global__ void __launch_bounds__(256) pair_build_kernel(const float* __restrict__ boxes, int N, float im_h,
                                                         float im_w, float margin, int64_t* __restrict__ ixs,
                                                         int64_t* __restrict__ ixo, float* __restrict__ rel_boxes,
                                                         float* __restrict__ masks) {
    const int p = blockIdx.x;
    const int i = p / (N - 1);
    const int jj = p - i * (N - 1);
    const int j = jj + (jj >= i ? 1 : 0);
    const float* s = boxes + (size_t)i * 4;
    const float* o = boxes + (size_t)j * 4;
    if (threadIdx.x == 0) {
        if (ixs) ixs[p] = i;
        if (ixo) ixo[p] = j;
        if (rel_boxes) {
            double m = (double)margin;
            float* r = rel_boxes + (size_t)p * 5;
            r[0] = 0.f;
            r[1] = (float)fmax(0.0, fmin((double)s[0], (double)o[0]) - m);
            r[2] = (float)fmax(0.0, fmin((double)s[1], (double)o[1]) - m);
            r[3] = (float)fmin((double)im_w, fmax((double)s[2], (double)o[2]) + m);
            r[4] = (float)fmin((double)im_h, fmax((double)s[3], (double)o[3]) + m);
        }
    }
    if (!masks) return;
    const double rh = 32.0 / (double)im_h, rw = 32.0 / (double)im_w;
    // threads 0-127 paint the subject mask, 128-255 the object mask; 8 cells (two float4) per thread
    const int which = threadIdx.x >> 7, t = threadIdx.x & 127;
    int x1, x2, y1, y2;
    mask_extent(which ? o : s, rh, rw, x1, x2, y1, y2);
    float4* dst = reinterpret_cast<float4*>(masks + ((size_t)p * 2 + which) * 1024);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        int v = t * 2 + q;          // float4 index inside the 32x32 mask
        int y = v >> 3, xb = (v & 7) * 4;
        bool row = (y >= y1 && y < y2);
        float4 m;
        m.x = (row && xb + 0 >= x1 && xb + 0 < x2) ? 1.f : 0.f;
        m.y = (row && xb + 1 >= x1 && xb + 1 < x2) ? 1.f : 0.f;
        m.z = (row && xb + 2 >= x1 && xb + 2 < x2) ? 1.f : 0.f;
        m.w = (row && xb + 3 >= x1 && xb + 3 < x2) ? 1.f : 0.f;
        dst[v] = m;
    }
}

// The pair stage of a whole frame group in ONE launch (faster_rcnn_SGG_emb.py:597-606,649-656 for every frame of the group):
// blockIdx.y = frame; the first `pair_blocks` blocks of a frame write 256 pairs each (indices into the group's row
// space f * N + i, union box with the frame number in column 0); the next N blocks paint one 32x32 object mask each.
// The relation head forms a pair's two mask channels from its subject's and object's masks (conv_lo is linear in its
// input channels), so the [P,2,32,32] pair masks -- 33 MB per 64-detection frame, of which 64 rows were read -- are not
// written at all.
__global__ void __launch_bounds__(256) pair_build_frames_kernel(const float* __restrict__ boxes, int N, int pair_blocks,
                                                                float im_h, float im_w, float margin,
                                                                int64_t* __restrict__ ixs, int64_t* __restrict__ ixo,
                                                                float* __restrict__ rel_boxes,
                                                                float* __restrict__ obj_masks) {
    const int f = blockIdx.y;
    const int P = N * (N - 1);
    const float* fb = boxes + (size_t)f * N * 4;
    if ((int)blockIdx.x < pair_blocks) {
        const int p = blockIdx.x * 256 + threadIdx.x;
        if (p >= P) return;
        const int i = p / (N - 1);
        const int jj = p - i * (N - 1);
        const int j = jj + (jj >= i ? 1 : 0);
        const float4 s = *reinterpret_cast<const float4*>(fb + (size_t)i * 4);
        const float4 o = *reinterpret_cast<const float4*>(fb + (size_t)j * 4);
        const size_t g = (size_t)f * P + p;
        if (ixs) ixs[g] = (int64_t)f * N + i;
        if (ixo) ixo[g] = (int64_t)f * N + j;
        if (rel_boxes) {
            const double m = (double)margin;
            float* r = rel_boxes + g * 5;
            r[0] = (float)f;
            r[1] = (float)fmax(0.0, fmin((double)s.x, (double)o.x) - m);
            r[2] = (float)fmax(0.0, fmin((double)s.y, (double)o.y) - m);
            r[3] = (float)fmin((double)im_w, fmax((double)s.z, (double)o.z) + m);
            r[4] = (float)fmin((double)im_h, fmax((double)s.w, (double)o.w) + m);
        }
        return;
    }
    if (!obj_masks) return;
    const int i = blockIdx.x - pair_blocks;
    const double rh = 32.0 / (double)im_h, rw = 32.0 / (double)im_w;
    int x1, x2, y1, y2;
    mask_extent(fb + (size_t)i * 4, rh, rw, x1, x2, y1, y2);
    float4* dst = reinterpret_cast<float4*>(obj_masks + ((size_t)f * N + i) * 1024);
    const int v = threadIdx.x;                 // float4 index inside the 32x32 mask
    const int y = v >> 3, xb = (v & 7) * 4;
    const bool row = (y >= y1 && y < y2);
    float4 mk;
    mk.x = (row && xb + 0 >= x1 && xb + 0 < x2) ? 1.f : 0.f;
    mk.y = (row && xb + 1 >= x1 && xb + 1 < x2) ? 1.f : 0.f;
    mk.z = (row && xb + 2 >= x1 && xb + 2 < x2) ? 1.f : 0.f;
    mk.w = (row && xb + 3 >= x1 && xb + 3 < x2) ? 1.f : 0.f;
    dst[v] = mk;
}

// ------------------------------------------------------------------------------------------ triplet top-k
constexpr int kTopThreads = 1024;
constexpr int kTopMax = 1024;  // largest supported top_k

__device__ __forceinline__ unsigned topk_desc_key(float f) {
    unsigned u = __float_as_uint(f);
    if (u == 0x80000000u) u = 0u;  // -0 ranks as +0
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ~u;
}
__device__ __forceinline__ float topk_key_value(unsigned k) {
    unsigned u = ~k;
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(u);
}

// The selection runs in three launches:
//   triplet_keys_kernel    (many CTAs) keys[i] = descending-order key of (rel * conf_s) * conf_o, and a histogram of the
//                          keys' top 11 bits (shared-memory counts flushed with integer atomics: deterministic);
//   triplet_filter_kernel  (many CTAs) finds the histogram bin d* in which the K-th best key lies and appends every
//                          (key, flat index) whose top bits are <= d* to a candidate list (a few hundred to a few
//                          thousand of the 532 K scores of a config-3 frame; list order is arbitrary);
//   triplet_select_kernel  (one CTA) sorts the candidates as 64-bit (key << 32 | index) values -- best score first, ties
//                          by lower flat index, whatever the list order -- and writes the K records.  If the list
//                          overflowed (more than kCandCap keys share the threshold bin, e.g. all scores equal) it falls
//                          back to a radix select over all keys.
constexpr int kCandCap = 16384;
constexpr int kHistBins = 2048;

__global__ void __launch_bounds__(256) triplet_keys_kernel(const float* __restrict__ rel_score,
                                                           const float* __restrict__ conf,
                                                           const int64_t* __restrict__ ixs,
                                                           const int64_t* __restrict__ ixo, int total, int R,
                                                           unsigned* __restrict__ keys, unsigned* __restrict__ ghist,
                                                           int num_boxes) {
    __shared__ unsigned hist[kHistBins];
    // blockIdx.y = frame of a group of frames with identical pair lists (ixs / ixo are frame-local)
    rel_score += (size_t)blockIdx.y * total;
    conf += (size_t)blockIdx.y * num_boxes;
    keys += (size_t)blockIdx.y * total;
    ghist += (size_t)blockIdx.y * (kHistBins + 64);
    for (int i = threadIdx.x; i < kHistBins; i += 256) hist[i] = 0;
    __syncthreads();
    for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
        const int p = i / R;
        // (rel * conf[ixs]) * conf[ixo], each product rounded to fp32 (lib/utils.py:611)
        const float v = __fmul_rn(__fmul_rn(__ldg(rel_score + i), __ldg(conf + ixs[p])), __ldg(conf + ixo[p]));
        const unsigned k = topk_desc_key(v);
        keys[i] = k;
        atomicAdd(&hist[k >> 21], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kHistBins; i += 256)
        if (hist[i]) atomicAdd(ghist + i, hist[i]);
}

// the bin where the running count of the histogram reaches `need` (bins in ascending key order = best scores first)
__device__ __forceinline__ unsigned threshold_bin(const unsigned* __restrict__ hist, unsigned need, int lane) {
    constexpr int per = kHistBins / 32;
    unsigned sum = 0;
    for (int q = 0; q < per; ++q) sum += hist[lane * per + q];
    unsigned incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const unsigned excl = incl - sum;
    unsigned bin = kHistBins - 1;
    if (excl < need && incl >= need) {
        unsigned run = excl;
        for (int q = 0; q < per; ++q) {
            run += hist[lane * per + q];
            if (run >= need) {
                bin = lane * per + q;
                break;
            }
        }
    } else {
        bin = 0xffffffffu;
    }
    const unsigned owner = __ballot_sync(0xffffffffu, bin != 0xffffffffu);
    return owner ? __shfl_sync(0xffffffffu, bin, __ffs(owner) - 1) : kHistBins - 1;
}

__global__ void __launch_bounds__(256) triplet_filter_kernel(const unsigned* __restrict__ keys, int total, int K,
                                                             const unsigned* __restrict__ ghist,
                                                             unsigned* __restrict__ counter,
                                                             unsigned long long* __restrict__ cand) {
    __shared__ unsigned s_bin;
    keys += (size_t)blockIdx.y * total;
    ghist += (size_t)blockIdx.y * (kHistBins + 64);
    counter += (size_t)blockIdx.y * (kHistBins + 64);
    cand += (size_t)blockIdx.y * kCandCap;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {
        const unsigned b = threshold_bin(ghist, (unsigned)K, lane);
        if (lane == 0) s_bin = b;
    }
    __syncthreads();
    const unsigned dstar = s_bin;
    const int n_iter = (total + gridDim.x * 256 - 1) / (gridDim.x * 256);
    for (int it = 0; it < n_iter; ++it) {
        const int i = (it * gridDim.x + blockIdx.x) * 256 + threadIdx.x;
        const unsigned k = i < total ? keys[i] : 0xffffffffu;
        const bool take = i < total && (k >> 21) <= dstar;
        const unsigned m = __ballot_sync(0xffffffffu, take);
        if (m) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(counter, (unsigned)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            const unsigned slot = base + __popc(m & ((1u << lane) - 1u));
            if (take && slot < (unsigned)kCandCap) cand[slot] = ((unsigned long long)k << 32) | (unsigned)i;
        }
    }
}

// Radix select over all keys by one CTA (the fallback of triplet_select_kernel): leaves the K winners in cand[0..K).
__device__ void select_all_keys(const unsigned* __restrict__ keys, int total, int K, unsigned* hist,
                                unsigned long long* cand) {
    __shared__ unsigned s_prefix, s_need, s_cnt, s_warp[32], s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        s_prefix = 0;
        s_need = K;
    }
    __syncthreads();

    // radix select, most significant digit first (11 + 11 + 10 bits): afterwards s_prefix is the key of the K-th
    // element and s_need the number of elements equal to it that still belong to the top K
    const int shifts[3] = {21, 10, 0};
    const int bits[3] = {11, 11, 10};
    unsigned known_mask = 0;
    for (int pass = 0; pass < 3 && K > 0; ++pass) {
        for (int i = tid; i < 2048; i += kTopThreads) hist[i] = 0;
        __syncthreads();
        const unsigned prefix = s_prefix;
        const unsigned dmask = (1u << bits[pass]) - 1u;
        for (int i = tid; i < total; i += kTopThreads) {
            unsigned k = keys[i];
            if ((k & known_mask) == prefix) atomicAdd(&hist[(k >> shifts[pass]) & dmask], 1u);
        }
        __syncthreads();
        if (warp == 0) {
            // find the digit where the running count reaches s_need: 64 bins per lane
            const int per = (1 << bits[pass]) / 32;
            const unsigned need = s_need;  // read by every lane before the owning lane rewrites it below
            unsigned sum = 0;
            for (int q = 0; q < per; ++q) sum += hist[lane * per + q];
            unsigned incl = sum;
            for (int o = 1; o < 32; o <<= 1) {
                unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            unsigned excl = incl - sum;
            bool mine = (excl < need) && (incl >= need);
            __syncwarp();
            if (mine) {
                unsigned run = excl;
                for (int q = 0; q < per; ++q) {
                    unsigned c = hist[lane * per + q];
                    if (run + c >= need) {
                        s_prefix = prefix | ((unsigned)(lane * per + q) << shifts[pass]);
                        s_need = need - run;
                        break;
                    }
                    run += c;
                }
            }
        }
        known_mask |= dmask << shifts[pass];
        __syncthreads();
    }

    // collect: everything strictly better than the threshold (any order), then the first s_need ties in index order
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    const unsigned thr = s_prefix;
    const unsigned need_ties = s_need;
    if (K > 0) {
        for (int i = tid; i < total; i += kTopThreads) {
            unsigned k = keys[i];
            if (k < thr) {
                unsigned slot = atomicAdd(&s_cnt, 1u);
                cand[slot] = ((unsigned long long)k << 32) | (unsigned)i;
            }
        }
        __syncthreads();
        if (tid == 0) s_base = s_cnt;  // == K - need_ties
        __syncthreads();
        unsigned taken = 0;
        for (int i0 = 0; i0 < total && taken < need_ties; i0 += kTopThreads) {
            int i = i0 + tid;
            bool tie = (i < total) && (keys[i] == thr);
            unsigned m = __ballot_sync(0xffffffffu, tie);
            if (lane == 0) s_warp[warp] = __popc(m);
            __syncthreads();
            unsigned before = 0, all = 0;
            for (int w = 0; w < 32; ++w) {
                unsigned c = s_warp[w];
                if (w < warp) before += c;
                all += c;
            }
            unsigned rank = taken + before + __popc(m & ((1u << lane) - 1u));
            if (tie && rank < need_ties) cand[s_base + rank] = ((unsigned long long)thr << 32) | (unsigned)i;
            taken += all;
            __syncthreads();
        }
    }
    __syncthreads();
}

// One CTA.  cand_list / counter come from triplet_filter_kernel; dynamic shared memory: max(kCandCap, kTopMax) 64-bit slots.
__global__ void __launch_bounds__(kTopThreads) triplet_select_kernel(
    const int64_t* __restrict__ classes, const float* __restrict__ boxes, const int64_t* __restrict__ ixs,
    const int64_t* __restrict__ ixo, int P, int R, int top_k, const unsigned* __restrict__ keys,
    const unsigned* __restrict__ counter, const unsigned long long* __restrict__ cand_list,
    float* __restrict__ record_out, int* __restrict__ count_out, int num_boxes) {
    extern __shared__ __align__(16) unsigned long long cand[];
    __shared__ unsigned hist[kHistBins];
    const int tid = threadIdx.x;
    const int total = P * R;
    {   // blockIdx.x = frame
        const size_t f = blockIdx.x;
        classes += f * num_boxes;
        boxes += f * num_boxes * 4;
        keys += f * total;
        counter += f * (kHistBins + 64);
        cand_list += f * kCandCap;
        record_out += f * top_k * 13;
        if (count_out) count_out += f;
    }
    const int K = min(top_k, total);
    const unsigned M = K > 0 ? *counter : 0u;
    int n_sort = K;
    if (M > (unsigned)kCandCap) {
        select_all_keys(keys, total, K, hist, cand);          // threshold bin too crowded: the slow, exact way
    } else {
        for (int i = tid; i < (int)M; i += kTopThreads) cand[i] = cand_list[i];
        n_sort = (int)M;
        __syncthreads();
    }
    // bitonic sort by (key, index) ascending == (score descending, flat index ascending); the first K are the winners
    int n2 = 1;
    while (n2 < n_sort) n2 <<= 1;
    for (int i = n_sort + tid; i < n2; i += kTopThreads) cand[i] = ~0ull;
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < n2 / 2; t += kTopThreads) {
                int lo = (t / stride) * stride * 2 + (t % stride);
                int hi = lo + stride;
                bool up = ((lo & size) == 0);
                unsigned long long a = cand[lo], b = cand[hi];
                if ((a > b) == up) {
                    cand[lo] = b;
                    cand[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    // records: (conf, cls_s, rel, cls_o, sub box x4, obj box x4, pair idx); zero rows past K
    for (int t = tid; t < top_k; t += kTopThreads) {
        float* rec = record_out + (size_t)t * 13;
        if (t < K) {
            unsigned long long e = cand[t];
            unsigned flat = (unsigned)(e & 0xffffffffu);
            int p = flat / R, r = flat - p * R;
            int64_t si = ixs[p], oi = ixo[p];
            rec[0] = topk_key_value((unsigned)(e >> 32));
            rec[1] = (float)classes[si];
            rec[2] = (float)r;
            rec[3] = (float)classes[oi];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                rec[4 + q] = boxes[si * 4 + q];
                rec[8 + q] = boxes[oi * 4 + q];
            }
            rec[12] = (float)p;
        } else {
#pragma unroll
            for (int q = 0; q < 13; ++q) rec[q] = 0.f;
        }
    }
    if (tid == 0 && count_out) *count_out = K;
}

}  // namespace i2v

using namespace i2v;

extern "C" int i2v_pair_build(const float* boxes, int num_boxes, float im_h, float im_w, float margin, int64_t* ixs,
                              int64_t* ixo, float* rel_boxes, float* masks, cudaStream_t stream) {
    I2V_REQUIRE(num_boxes >= 0, "pair_build: bad size");
    if (num_boxes < 2) return I2V_OK;  // faster_rcnn_SGG_emb.py:590-595: no pairs
    I2V_REQUIRE(boxes, "pair_build: null boxes");
    I2V_REQUIRE((int64_t)num_boxes * (num_boxes - 1) <= INT32_MAX, "pair_build: too many pairs");
    I2V_REQUIRE(!masks || ((uintptr_t)masks & 15) == 0, "pair_build: masks must be 16-byte aligned");
    int P = num_boxes * (num_boxes - 1);
    pair_build_kernel<<<P, 256, 0, stream>>>(boxes, num_boxes, im_h, im_w, margin, ixs, ixo, rel_boxes, masks);
    return check_launch("pair_build_kernel");
}

struct TopkWs {
    unsigned* keys;             // [frames][pairs * rel]
    unsigned* hist;             // [frames][kHistBins + 64]: the histogram followed by the candidate counter
    unsigned long long* cand;   // [frames][kCandCap]
    size_t bytes;
};
static TopkWs carve_topk_ws(void* ws, int frames, int num_pairs, int num_rel) {
    Carver cv(ws);
    TopkWs w{};
    w.keys = cv.take<unsigned>((size_t)frames * num_pairs * num_rel);
    w.hist = cv.take<unsigned>((size_t)frames * (kHistBins + 64));
    w.cand = cv.take<unsigned long long>((size_t)frames * kCandCap);
    w.bytes = cv.used();
    return w;
}

extern "C" size_t i2v_triplet_topk_workspace_bytes(int num_pairs, int num_rel) {
    if (num_pairs < 0 || num_rel < 0) return 0;
    return carve_topk_ws(nullptr, 1, num_pairs, num_rel).bytes;
}
extern "C" size_t i2v_triplet_topk_frames_workspace_bytes(int frames, int num_pairs, int num_rel) {
    if (frames < 0 || num_pairs < 0 || num_rel < 0) return 0;
    return carve_topk_ws(nullptr, frames, num_pairs, num_rel).bytes;
}

extern "C" int i2v_triplet_topk_frames(const float* rel_score, const float* conf, const int64_t* classes, const float* boxes,
                                       const int64_t* ixs, const int64_t* ixo, int frames, int num_boxes, int num_pairs,
                                       int num_rel, int top_k, float* record_out, int* count_out, void* workspace,
                                       size_t workspace_bytes, cudaStream_t stream) {
    I2V_REQUIRE(frames >= 0 && num_boxes >= 0 && num_pairs >= 0 && num_rel >= 0 && top_k >= 1 && top_k <= kTopMax,
                "triplet_topk: bad size (top_k <= %d)", kTopMax);
    I2V_REQUIRE((int64_t)num_pairs * num_rel <= INT32_MAX && frames <= 65535, "triplet_topk: too many scores or frames");
    if (frames == 0) return I2V_OK;
    I2V_REQUIRE(record_out, "triplet_topk: null record_out");
    const int total = num_pairs * num_rel;
    const size_t need = i2v_triplet_topk_frames_workspace_bytes(frames, num_pairs, num_rel);
    if (total > 0) I2V_REQUIRE(rel_score && conf && classes && boxes && ixs && ixo, "triplet_topk: null pointer");
    if (!workspace || workspace_bytes < need) {
        set_error("triplet_topk: workspace %zu < %zu bytes", workspace_bytes, need);
        return I2V_ERR_WORKSPACE;
    }
    const TopkWs w = carve_topk_ws(workspace, frames, num_pairs, num_rel);
    I2V_CUDA_TRY(cudaMemsetAsync(w.hist, 0, (size_t)frames * (kHistBins + 64) * sizeof(unsigned), stream));
    if (total > 0) {
        const int per_frame = std::max(1, 2 * kNumSMs / frames);
        const int gx = (int)std::min<int64_t>(per_frame, ((int64_t)total + 255) / 256);
        const dim3 grid((unsigned)gx, (unsigned)frames);
        triplet_keys_kernel<<<grid, 256, 0, stream>>>(rel_score, conf, ixs, ixo, total, num_rel, w.keys, w.hist, num_boxes);
        I2V_TRY(check_launch("triplet_keys_kernel"));
        triplet_filter_kernel<<<grid, 256, 0, stream>>>(w.keys, total, std::min(top_k, total), w.hist, w.hist + kHistBins,
                                                        w.cand);
        I2V_TRY(check_launch("triplet_filter_kernel"));
    }
    const size_t smem = (size_t)kCandCap * sizeof(unsigned long long);
    I2V_CUDA_TRY(cudaFuncSetAttribute(triplet_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    triplet_select_kernel<<<frames, kTopThreads, smem, stream>>>(classes, boxes, ixs, ixo, num_pairs, num_rel, top_k, w.keys,
                                                                w.hist + kHistBins, w.cand, record_out, count_out, num_boxes);
    return check_launch("triplet_select_kernel");
}

extern "C" int i2v_triplet_topk(const float* rel_score, const float* conf, const int64_t* classes, const float* boxes,
                                const int64_t* ixs, const int64_t* ixo, int num_pairs, int num_rel, int top_k,
                                float* record_out, int* count_out, void* workspace, size_t workspace_bytes,
                                cudaStream_t stream) {
    // one frame; the strides of the frame dimension are never used
    return i2v_triplet_topk_frames(rel_score, conf, classes, boxes, ixs, ixo, 1, 0, num_pairs, num_rel, top_k, record_out,
                                   count_out, workspace, workspace_bytes, stream);
}

extern "C" int i2v_pair_build_frames(const float* boxes, int frames, int num_boxes, float im_h, float im_w, float margin,
                                     int64_t* ixs, int64_t* ixo, float* rel_boxes, float* obj_masks, cudaStream_t stream) {
    I2V_REQUIRE(frames >= 0 && num_boxes >= 0 && frames <= 65535, "pair_build_frames: bad size");
    if (frames == 0 || num_boxes == 0) return I2V_OK;
    I2V_REQUIRE(boxes && ((uintptr_t)boxes & 15) == 0, "pair_build_frames: boxes must be a 16-byte aligned device pointer");
    I2V_REQUIRE((int64_t)frames * num_boxes * (num_boxes - 1) <= INT32_MAX, "pair_build_frames: too many pairs");
    I2V_REQUIRE(!obj_masks || ((uintptr_t)obj_masks & 15) == 0, "pair_build_frames: masks must be 16-byte aligned");
    const int P = num_boxes * (num_boxes - 1);
    const int pair_blocks = num_boxes >= 2 ? (P + 255) / 256 : 0;
    const dim3 grid((unsigned)(pair_blocks + (obj_masks ? num_boxes : 0)), (unsigned)frames);
    if (grid.x == 0) return I2V_OK;
    pair_build_frames_kernel<<<grid, 256, 0, stream>>>(boxes, num_boxes, pair_blocks, im_h, im_w, margin, ixs, ixo, rel_boxes,
                                                       obj_masks);
    return check_launch("pair_build_frames_kernel");
}
