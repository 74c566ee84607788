// Error reporting, ABI version and the five drop-in launchers that keep the reference's names and argument
// lists (lib/model/roi_align/src/roi_align_kernel.h:13-27, lib/model/roi_pooling/src/roi_pooling_kernel.h:8-18,
// lib/model/nms/src/nms_cuda_kernel.h:5-6).  The launchers have no workspace argument, so they keep one grow-only
// device buffer per (thread, device, stream); everything else in the library takes the caller's workspace.
#include <limits.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <utility>

#include "common.cuh"

namespace i2v {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

struct Scratch {
    void* ptr = nullptr;
    size_t bytes = 0;
    int device = -1;
    cudaStream_t stream = nullptr;
};
constexpr int kMaxScratch = 16;
static thread_local Scratch g_scratch[kMaxScratch];
static thread_local int g_scratch_used = 0;

// Grow-only scratch for the legacy launchers, one buffer per (thread, device, stream): two launcher calls of one thread
// on different streams never share tables, like the stateless reference launchers.  Growing a buffer waits for the work
// queued on its stream (the old buffer may still be in use there), which happens a handful of times per process.  With
// more than kMaxScratch live (device, stream) pairs in one thread the oldest entry is recycled the same way.
void* legacy_scratch(size_t bytes, size_t* have, cudaStream_t stream) {
    int dev = 0;
    if (have) *have = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    Scratch* e = nullptr;
    for (int i = 0; i < g_scratch_used; ++i)
        if (g_scratch[i].device == dev && g_scratch[i].stream == stream) e = &g_scratch[i];
    if (e && e->bytes >= bytes) {
        if (have) *have = e->bytes;
        return e->ptr;
    }
    if (!e) {
        if (g_scratch_used < kMaxScratch) {
            e = &g_scratch[g_scratch_used++];
        } else {
            e = &g_scratch[0];
            for (int i = 1; i < kMaxScratch; ++i) std::swap(g_scratch[i - 1], g_scratch[i]);   // recycle the oldest
            e = &g_scratch[kMaxScratch - 1];
        }
    }
    if (e->ptr) {
        int cur = dev;
        if (e->device != dev) cudaSetDevice(e->device);
        cudaStreamSynchronize(e->stream);
        cudaFree(e->ptr);
        if (e->device != cur) cudaSetDevice(cur);
        cudaGetLastError();
    }
    *e = Scratch{};
    size_t want = bytes < (1u << 20) ? (1u << 20) : bytes + bytes / 2;
    void* p = nullptr;
    if (cudaMalloc(&p, want) != cudaSuccess) {
        set_error("legacy launcher: cudaMalloc(%zu) failed", want);
        cudaGetLastError();
        e->device = -1;
        return nullptr;
    }
    e->ptr = p;
    e->bytes = want;
    e->device = dev;
    e->stream = stream;
    if (have) *have = want;
    return p;
}
}  // namespace i2v

using namespace i2v;

extern "C" const char* i2v_last_error(void) { return g_error; }
extern "C" int i2v_abi_version(void) { return 1; }

extern "C" int ROIPoolForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois,
                                     const int height, const int width, const int channels, const int pooled_height,
                                     const int pooled_width, const float* bottom_rois, float* top_data,
                                     int* argmax_data, cudaStream_t stream) {
    // roi_pooling_kernel.h:8-12 does not pass the batch size: accept every frame index whose flat arg-max
    // (roi_pooling_kernel.cu:85) still fits the int32 output
    int64_t frame = (int64_t)channels * height * width;
    int max_batch = (int)(INT32_MAX / (frame > 0 ? frame : 1));
    if (max_batch < 1) {
        set_error("ROIPoolForwardLaucher: one frame does not fit the int32 arg-max");
        return 0;
    }
    // the plane-resident kernel loads whole frames, so it needs the real frame count: read it back from the RoI list
    int frames = max_batch;
    if (getenv("I2V_POOL_PLANE") != nullptr && num_rois > 0 && pooled_height == 7 && pooled_width == 7 && channels % 16 == 0) {
        size_t have = 0;
        int* scratch = static_cast<int*>(legacy_scratch(256, &have, stream));
        int found = 0;
        if (!scratch || legacy_frame_count(bottom_rois, num_rois, scratch, stream, &found) != I2V_OK) return 0;
        if (found > max_batch) {
            set_error("ROIPoolForwardLaucher: frame index %d does not fit the int32 arg-max", found - 1);
            return 0;
        }
        frames = found > 0 ? found : 1;
    }   // other shapes stay on the per-element kernels, which never touch a frame no RoI names
    int rc = i2v_roi_pool_forward(bottom_data, bottom_rois, top_data, argmax_data, frames, channels, height, width, num_rois,
                              pooled_height, pooled_width, spatial_scale, I2V_ARGMAX_FLAT, stream);
    return rc == I2V_OK ? 1 : 0;
}

extern "C" int ROIPoolBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size,
                                      const int num_rois, const int height, const int width, const int channels,
                                      const int pooled_height, const int pooled_width, const float* bottom_rois,
                                      float* bottom_diff, const int* argmax_data, cudaStream_t stream) {
    int rc = i2v_roi_pool_backward(top_diff, bottom_rois, argmax_data, bottom_diff, batch_size, channels, height, width,
                                   num_rois, pooled_height, pooled_width, spatial_scale, I2V_ARGMAX_FLAT, stream);
    return rc == I2V_OK ? 1 : 0;
}

extern "C" void nms_cuda_compute(int* keep_out, int* num_out, float* boxes_host, int boxes_num, int boxes_dim,
                                 float nms_overlap_thresh) {
    if (boxes_num < 0 || boxes_dim < 4 || !keep_out || !num_out) {
        set_error("nms_cuda_compute: bad argument");
        return;
    }
    size_t box_bytes = align_up((size_t)boxes_num * boxes_dim * sizeof(float), 256);
    size_t need = box_bytes + i2v_nms_workspace_bytes(1, boxes_num);
    size_t have = 0;
    char* ws = static_cast<char*>(legacy_scratch(need, &have, nullptr));
    if (!ws) return;
    const float* dev_boxes = boxes_host;
    cudaPointerAttributes attr;
    bool on_device = (cudaPointerGetAttributes(&attr, boxes_host) == cudaSuccess) &&
                     (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    if (!on_device && boxes_num > 0) {
        // nms_cuda_kernel.cu:99-101: a blocking host-to-device copy of the box list
        if (cudaMemcpy(ws, boxes_host, (size_t)boxes_num * boxes_dim * sizeof(float), cudaMemcpyHostToDevice) !=
            cudaSuccess) {
            set_error("nms_cuda_compute: copying boxes to the device failed");
            cudaGetLastError();
            return;
        }
        dev_boxes = reinterpret_cast<const float*>(ws);
    }
    int rc = i2v_nms_sorted(dev_boxes, 1, boxes_num, boxes_dim, nms_overlap_thresh, 0, keep_out, boxes_num, num_out,
                            ws + box_bytes, have - box_bytes, 0);
    if (rc != I2V_OK) return;
    if (cudaStreamSynchronize(0) != cudaSuccess) {  // synchronous like the reference (nms_cuda_kernel.cu:117-160)
        set_error("nms_cuda_compute: %s", cudaGetErrorString(cudaGetLastError()));
    }
}
