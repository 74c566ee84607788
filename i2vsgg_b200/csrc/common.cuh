// Shared helpers for the sm_100a kernels of the I2VSGG region-level hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/i2vsgg_b200.h"

namespace i2v {

// Records a human-readable message for i2v_last_error(); thread-local.
void set_error(const char* fmt, ...);

// Grow-only per-(thread, device, stream) device buffer for the reference-signature launchers, which have no workspace
// argument.  Returns nullptr (and sets the error) when the allocation fails; *have receives the buffer size.
void* legacy_scratch(size_t bytes, size_t* have, cudaStream_t stream);

// For the reference-signature forward launchers, whose argument lists lack the batch size: the largest frame index of
// the RoI list + 1, found on the device and read back (one 4-byte copy and one synchronisation of `stream`).  *frames = 0
// when no RoI has a non-negative index.  `scratch` is at least 4 bytes of device memory owned by the caller.
int legacy_frame_count(const float* rois, int num_rois, int* scratch, cudaStream_t stream, int* frames);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return I2V_ERR_CUDA;
    }
    return I2V_OK;
}

#define I2V_CUDA_TRY(expr)                                                        \
    do {                                                                          \
        cudaError_t _e = (expr);                                                  \
        if (_e != cudaSuccess) {                                                  \
            ::i2v::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));     \
            return I2V_ERR_CUDA;                                                  \
        }                                                                         \
    } while (0)

#define I2V_REQUIRE(cond, ...)              \
    do {                                    \
        if (!(cond)) {                      \
            ::i2v::set_error(__VA_ARGS__);  \
            return I2V_ERR_INVALID;         \
        }                                   \
    } while (0)

#define I2V_TRY(expr)                 \
    do {                              \
        int _rc = (expr);             \
        if (_rc != I2V_OK) return _rc; \
    } while (0)

constexpr int kNumSMs = 148;            // B200: 2 dies x 74 SMs
constexpr int kMaxSmemPerCta = 232448;  // 227 KB opt-in limit on sm_100a

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Grid size for a grid-stride loop over `total` items: enough CTAs to cover the work, capped at a few
// resident waves of the 148 SMs.
inline int grid_for(int64_t total, int threads, int ctas_per_sm = 8) {
    int64_t need = ceil_div64(total, threads);
    int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// Carves aligned sub-buffers out of a caller-provided workspace.
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* p) : base(static_cast<char*>(p)) {}
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 256);
        T* p = reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return p;
    }
    size_t used() const { return align_up(off, 256); }
};

// Streaming (evict-first) 16-byte store: outputs of the RoI ops are written once and not re-read by us.
__device__ __forceinline__ void st_global_cs_v4(float* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float4 ld_global_cs_v4(const float* p) {
    float4 v;
    asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// ---- lattice RoIAlign axis tables (one per RoI, built by the prep kernel in roi_align.cu) -------------
constexpr int kMaxLattice = 16;  // lattice points per axis supported by the table-driven kernels

struct LatticeAxis {
    int start[kMaxLattice];   // hstart / wstart  (roi_align_kernel.cu:47-48)
    float frac[kMaxLattice];  // h_ratio / w_ratio (roi_align_kernel.cu:56-57)
};
struct alignas(16) LatticeRoi {
    LatticeAxis y, x;
    int batch;            // -1 when the RoI's batch index is outside [0,B)
    unsigned valid_y;     // bit p set <=> !(h < 0 || h >= H) for lattice row p (roi_align_kernel.cu:54)
    unsigned valid_x;
    unsigned flags;       // bit0: y starts strictly increasing over valid rows; bit1: same for x
    // duplicate structure of the start cells (starts are non-decreasing), used by the plane backward:
    unsigned y_runpos;    // nibble p: how many earlier valid rows share row p's start cell (position in its run)
    unsigned y_maxrun;    // 1 + largest nibble of y_runpos (0 when no row is valid)
    unsigned x_same;      // bit p: valid column p has the same start cell as valid column p-1
    unsigned pad_;
};

constexpr size_t kRoiTabSlotBytes = 640;  // workspace bytes per RoI for the kernel-specific view of its tables

// What the forward plane kernel needs of one RoI (lattices of up to 8 points per axis): byte offsets into the
// [row][64 columns][16 channels] shared-memory planes and weights with validity (and the avg pool's 1/4) folded in.
// A bilinear sample reads two horizontally adjacent cells; xa lists them even column first, xb odd column first.
struct alignas(16) PlaneTab {
    float4 xa[8];  // {byte offset of the even-column cell (int bits), its weight, offset of the odd-column cell, its weight}
    float4 xb[8];  // the same pair in the opposite order
    float4 y[8];   // {byte offset of the upper row (int bits), weight of the upper row, weight of the lower row, 0}
};

// What the backward plane kernel needs of one RoI: offsets into a [H][W] fp32 plane and weights with the avg pool's
// 1/4 folded in (zero for out-of-range lattice columns), plus the duplicate structure of the start cells.
struct alignas(16) BwdTab {
    int xoff[8];    // start column * 4 bytes
    float wx0[8];   // weight of the left cell
    float wx1[8];   // weight of the right cell
    int yoff[8];    // start row * W * 4 bytes
    float wy0[8];   // weight of the upper row
    float wy1[8];   // weight of the lower row
    unsigned valid_y, valid_x, y_runpos, y_maxrun, x_same, pad_[3];
};

}  // namespace i2v
