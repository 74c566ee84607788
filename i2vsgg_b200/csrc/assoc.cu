// Greedy temporal association of per-frame triplets into video relations on the device
// (lib/utils.py:134-182 `greedy_relational_association` with `VideoRelation`, :37-98, and `_iou`, :20-32).
//
// The algorithm is sequential in the frames and, inside a frame, in the predictions (a relation that has been extended
// is taken off the candidate list, :168-169), so one CTA walks the clip.  Per frame everything that does not depend on
// that order runs in parallel: the stable sorts (rank sorts of <= 128 keys); the STATIC part of every (prediction,
// candidate) test -- labels, end frame, both IoUs -- as one bitmask of candidates per prediction (thread = candidate);
// then one thread walks the predictions and only intersects each mask with the candidates still alive (first set bit =
// the reference's first match); then thread = prediction builds the extended / new relation (new ids and slots from a
// prefix count), thread = candidate returns the slots of relations that ended, and the mean confidence of every live
// relation is recomputed (numpy's pairwise float64 summation, reproduced term for term because the mean is a sort key).
// Three barriers per PREDICTION became about ten per FRAME: 48 -> 6 us per frame.
// All comparisons that decide the result are made in double on float32-exact inputs, like the reference's Python floats.
//
// Outputs describe the relations without materialising trajectories: for every frame position f and every prediction j
// in descending-confidence order, `order[f][j]` is the record row and `rel_id[f][j]` the relation (numbered in creation
// order) it started or extended; `rel_info[r] = (first frame number, end frame number, s, p, o, length)` and
// `rel_score[r]` = mean confidence.  The host turns that into the reference's dictionaries (i2vsgg_b200/sgg.py).
#include "common.cuh"

namespace i2v {
namespace {

constexpr int kMaxK = 128;          // predictions per frame (the reference keeps 100, lib/utils.py:141-142)
constexpr int kThreads = 128;
constexpr int kRec = 13;            // conf, cls_s, rel, cls_o, sub box x4, obj box x4, pair idx

// numpy's pairwise_sum (loops_utils.h.src) for a contiguous float64 reduction, fed with float32 confidences.  numpy
// recurses on halves (the left one rounded down to a multiple of 8) above 128 terms; the recursion is unrolled onto a small
// explicit stack here: a recursive device function spills its caller's registers once per level and overran the default
// 1 KB thread stack as soon as a relation was longer than 128 frames.
__device__ __forceinline__ double pairwise_leaf(const float* a, int n) {     // n <= 128
    if (n < 8) {
        double res = 0.;
        for (int i = 0; i < n; ++i) res += (double)a[i];
        return res;
    }
    double r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = (double)a[k];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] += (double)a[i + k];
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += (double)a[i];
    return res;
}

__device__ double pairwise_sum(const float* a, int n) {
    if (n <= 128) return pairwise_leaf(a, n);
    constexpr int kDepth = 24;                  // halving from 2^31 terms down to 128 needs fewer levels than this
    int off[kDepth], len[kDepth], stage[kDepth];
    double left[kDepth];
    int sp = 0;
    off[0] = 0;
    len[0] = n;
    stage[0] = 0;
    double ret = 0.;
    while (sp >= 0) {
        if (len[sp] <= 128) {
            ret = pairwise_leaf(a + off[sp], len[sp]);
            --sp;
            continue;
        }
        int n2 = len[sp] / 2;
        n2 -= n2 % 8;
        if (stage[sp] == 0) {                   // descend into the left half
            stage[sp] = 1;
            off[sp + 1] = off[sp];
            len[sp + 1] = n2;
            stage[sp + 1] = 0;
            ++sp;
        } else if (stage[sp] == 1) {            // left half done: keep it, descend into the right half
            left[sp] = ret;
            stage[sp] = 2;
            off[sp + 1] = off[sp] + n2;
            len[sp + 1] = len[sp] - n2;
            stage[sp + 1] = 0;
            ++sp;
        } else {                                // both halves done
            ret = left[sp] + ret;
            --sp;
        }
    }
    return ret;
}

// lib/utils.py:20-32
__device__ __forceinline__ double box_iou(const float4 a, const float4 b) {
    const double left = fmax((double)a.x, (double)b.x), right = fmin((double)a.z, (double)b.z);
    const double up = fmax((double)a.y, (double)b.y), down = fmin((double)a.w, (double)b.w);
    if (left >= right || down <= up) return 0.;
    const double s1 = ((double)a.z - (double)a.x) * ((double)a.w - (double)a.y);
    const double s2 = ((double)b.z - (double)b.x) * ((double)b.w - (double)b.y);
    const double sc = (down - up) * (right - left);
    return sc / (s1 + s2 - sc);
}

struct Entry {          // a relation that was started or extended in a frame
    int slot, uid, s, p, o, fend, len, alive;
    float4 sb, ob;      // its last subject / object box
    double mean;
};

__global__ void __launch_bounds__(kThreads, 1)
    greedy_association_kernel(const float* __restrict__ records, const int* __restrict__ counts,
                              const int* __restrict__ frame_numbers, const int* __restrict__ source_frame, int frames,
                              int top_k, int max_traj, int* __restrict__ rel_id, int* __restrict__ order,
                              int* __restrict__ rel_info, double* __restrict__ rel_score, int* __restrict__ num_rel,
                              float* __restrict__ slot_confs /* [2 * kMaxK][frames] */) {
    __shared__ Entry last[kMaxK], cur[kMaxK], tmp[kMaxK];
    __shared__ float p_conf[kMaxK];
    __shared__ int p_s[kMaxK], p_p[kMaxK], p_o[kMaxK];
    __shared__ float4 p_sb[kMaxK], p_ob[kMaxK];
    __shared__ int free_slots[2 * kMaxK];
    __shared__ int n_last, n_cur, n_free, n_rel;
    __shared__ unsigned match_mask[kMaxK][4];   // per prediction: the candidates that match it statically
    __shared__ unsigned alive_mask[4];          // candidates that were not extended in this frame
    __shared__ int matched[kMaxK];              // per prediction: its candidate, or -1
    __shared__ int warp_new[4];
    const int tid = threadIdx.x;

    for (int i = tid; i < 2 * kMaxK; i += kThreads) free_slots[i] = 2 * kMaxK - 1 - i;
    if (tid == 0) {
        n_last = 0;
        n_free = 2 * kMaxK;
        n_rel = 0;
    }
    for (int i = tid; i < frames * top_k; i += kThreads) {
        rel_id[i] = -1;
        order[i] = -1;
    }
    __syncthreads();

    for (int f = 0; f < frames; ++f) {
        const int src = source_frame ? source_frame[f] : f;
        const int fno = frame_numbers ? frame_numbers[f] : f;
        int cnt = src >= 0 ? min(counts[src], top_k) : 0;
        const float* rec = records + (size_t)max(src, 0) * top_k * kRec;

        // ---- A. the frame's predictions in descending confidence, stable (lib/utils.py:140-142) ----
        float my_conf = 0.f;
        if (tid < cnt) my_conf = rec[tid * kRec];
        if (tid < kMaxK) p_conf[tid] = my_conf;          // staging for the rank computation
        __syncthreads();
        int rank = 0;
        if (tid < cnt) {
            for (int k = 0; k < cnt; ++k) {
                const float c = p_conf[k];
                rank += (c > my_conf) || (c == my_conf && k < tid);
            }
        }
        __syncthreads();
        const int n_pred = min(cnt, max_traj);
        if (tid < cnt && rank < n_pred) {
            const float* r = rec + tid * kRec;
            p_conf[rank] = my_conf;
            p_s[rank] = (int)r[1];
            p_p[rank] = (int)r[2];
            p_o[rank] = (int)r[3];
            p_sb[rank] = make_float4(r[4], r[5], r[6], r[7]);
            p_ob[rank] = make_float4(r[8], r[9], r[10], r[11]);
            order[(size_t)f * top_k + rank] = tid;
        }
        if (tid == 0) n_cur = 0;
        __syncthreads();

        // ---- B. candidates by mean confidence, descending, stable (:159; the list only loses elements while the
        // frame's predictions are walked, so one sort per frame is what the per-prediction sort amounts to) ----
        const int nl = n_last;
        if (f > 0 && n_pred > 0 && nl > 1) {
            if (tid < nl) {
                const double m = last[tid].mean;
                int rk = 0;
                for (int k = 0; k < nl; ++k) {
                    const double mk = last[k].mean;
                    rk += (mk > m) || (mk == m && k < tid);
                }
                tmp[rk] = last[tid];
            }
            __syncthreads();
            if (tid < nl) last[tid] = tmp[tid];
            __syncthreads();
        }

        // ---- C1. static matches: bit k of match_mask[j] <=> candidate k (in candidate order) has prediction j's labels,
        // ends where this frame begins and overlaps both boxes by >= 0.5 (:160-167 without the "still a candidate" part)
        for (int i = tid; i < kMaxK * 4; i += kThreads) (&match_mask[0][0])[i] = 0u;
        __syncthreads();
        if (f > 0 && tid < nl) {
            const Entry e = last[tid];
            if (e.fend == fno) {
                for (int j = 0; j < n_pred; ++j) {
                    if (e.s == p_s[j] && e.p == p_p[j] && e.o == p_o[j] && box_iou(e.sb, p_sb[j]) >= 0.5 &&
                        box_iou(e.ob, p_ob[j]) >= 0.5)
                        atomicOr(&match_mask[j][tid >> 5], 1u << (tid & 31));
                }
            }
        }
        __syncthreads();
        // ---- C2. the predictions in turn (:144-178): the first candidate that matches and has not been extended yet ----
        if (tid == 0) {
            unsigned alive[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
            for (int j = 0; j < n_pred; ++j) {
                int hit = -1;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const unsigned m = match_mask[j][w] & alive[w];
                    if (hit < 0 && m) {
                        const int b = __ffs(m) - 1;
                        hit = w * 32 + b;
                        alive[w] &= ~(1u << b);
                    }
                }
                matched[j] = hit;
            }
#pragma unroll
            for (int w = 0; w < 4; ++w) alive_mask[w] = alive[w];
        }
        __syncthreads();
        // ---- C3. thread = prediction: extend the matched relation (:76-81) or start a new one (:170-173) ----
        {
            const bool mine = tid < n_pred;
            const int hit = mine ? matched[tid] : 0;
            const bool fresh = mine && hit < 0;
            const unsigned fb = __ballot_sync(0xffffffffu, fresh);
            if ((tid & 31) == 0) warp_new[tid >> 5] = __popc(fb);
            __syncthreads();
            int before = __popc(fb & ((1u << (tid & 31)) - 1u));
            for (int w = 0; w < (tid >> 5); ++w) before += warp_new[w];
            if (mine) {
                Entry e;
                if (!fresh) {
                    e = last[hit];
                    e.fend += 1;
                    rel_info[e.uid * 6 + 1] = e.fend;
                } else {
                    e.uid = n_rel + before;
                    e.slot = free_slots[n_free - 1 - before];
                    e.s = p_s[tid];
                    e.p = p_p[tid];
                    e.o = p_o[tid];
                    e.fend = fno + 1;
                    e.len = 0;
                    int* info = rel_info + e.uid * 6;
                    info[0] = fno;
                    info[1] = fno + 1;
                    info[2] = e.s;
                    info[3] = e.p;
                    info[4] = e.o;
                }
                slot_confs[(size_t)e.slot * frames + e.len] = p_conf[tid];
                e.len += 1;
                rel_info[e.uid * 6 + 5] = e.len;
                e.sb = p_sb[tid];
                e.ob = p_ob[tid];
                e.alive = 1;
                rel_id[(size_t)f * top_k + tid] = e.uid;
                cur[tid] = e;
            }
            __syncthreads();
            if (tid == 0) {
                const int total_new = warp_new[0] + warp_new[1] + warp_new[2] + warp_new[3];
                n_rel += total_new;
                n_free -= total_new;
                n_cur = n_pred;
            }
            __syncthreads();
        }

        // ---- D. relations that were not extended are closed (their slots return in candidate order); the rest carry over
        // with a fresh mean (:65-66) ----
        {
            const bool ended = f > 0 && tid < nl && ((alive_mask[tid >> 5] >> (tid & 31)) & 1u);
            const unsigned eb = __ballot_sync(0xffffffffu, ended);
            if ((tid & 31) == 0) warp_new[tid >> 5] = __popc(eb);
            __syncthreads();
            int before = __popc(eb & ((1u << (tid & 31)) - 1u));
            for (int w = 0; w < (tid >> 5); ++w) before += warp_new[w];
            if (ended) free_slots[n_free + before] = last[tid].slot;
            __syncthreads();
            if (tid == 0) n_free += warp_new[0] + warp_new[1] + warp_new[2] + warp_new[3];
        }
        __syncthreads();
        const int nc = n_cur;
        if (tid < nc) {
            Entry e = cur[tid];
            e.mean = pairwise_sum(slot_confs + (size_t)e.slot * frames, e.len) / (double)e.len;
            rel_score[e.uid] = e.mean;
            last[tid] = e;
        }
        if (tid == 0) n_last = nc;
        __syncthreads();
    }
    if (tid == 0) *num_rel = n_rel;
}

}  // namespace
}  // namespace i2v

using namespace i2v;

extern "C" size_t i2v_association_workspace_bytes(int frames, int top_k) {
    if (frames < 0 || top_k < 0) return 0;
    return align_up((size_t)2 * kMaxK * (size_t)(frames > 0 ? frames : 1) * sizeof(float), 256);
}

extern "C" int i2v_greedy_association(const float* records, const int* counts, const int* frame_numbers,
                                      const int* source_frame, int frames, int top_k, int max_traj, int* rel_id,
                                      int* order, int* rel_info, double* rel_score, int* num_rel, void* workspace,
                                      size_t workspace_bytes, cudaStream_t stream) {
    I2V_REQUIRE(frames >= 0 && top_k >= 1 && top_k <= kMaxK && max_traj >= 1, "greedy_association: bad shape (top_k <= %d)",
                kMaxK);
    I2V_REQUIRE(num_rel, "greedy_association: null num_rel");
    if (frames == 0) {
        I2V_CUDA_TRY(cudaMemsetAsync(num_rel, 0, sizeof(int), stream));
        return I2V_OK;
    }
    I2V_REQUIRE(records && counts && rel_id && order && rel_info && rel_score, "greedy_association: null pointer");
    size_t need = i2v_association_workspace_bytes(frames, top_k);
    if (!workspace || workspace_bytes < need) {
        set_error("greedy_association: workspace %zu < %zu bytes", workspace_bytes, need);
        return I2V_ERR_WORKSPACE;
    }
    greedy_association_kernel<<<1, kThreads, 0, stream>>>(records, counts, frame_numbers, source_frame, frames, top_k,
                                                         max_traj, rel_id, order, rel_info, rel_score, num_rel,
                                                         static_cast<float*>(workspace));
    return check_launch("greedy_association_kernel");
}
