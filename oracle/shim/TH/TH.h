/*
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Minimal stand-in for the PyTorch-0.4 `TH/TH.h` C tensor API so that the
 * reference's own CPU sources
 *   /root/reference/lib/model/roi_align/src/roi_align.c      (includes <TH/TH.h> at :1)
 *   /root/reference/lib/model/roi_pooling/src/roi_pooling.c  (includes <TH/TH.h> at :1)
 *   /root/reference/lib/model/roi_crop/src/roi_crop.c        (includes <TH/TH.h> at :1)
 * compile *unmodified, from where they lie* into oracle/_ref/libref_cpu.so.
 * Only the six accessors those files call are provided.  The struct is
 * filled from Python (ctypes) by oracle/ref.py.
 */
#ifndef I2V_ORACLE_TH_SHIM_H
#define I2V_ORACLE_TH_SHIM_H

#include <stddef.h>

typedef struct THFloatStorage {
    float *data;
    long numel;
} THFloatStorage;

typedef struct THFloatTensor {
    float *data;
    long size[4];
    int ndim;
    THFloatStorage storage;
} THFloatTensor;

static inline float *THFloatTensor_data(THFloatTensor *t) { return t->data; }
static inline long THFloatTensor_size(THFloatTensor *t, int d) { return t->size[d]; }
/* contiguous tensors only (what oracle/ref.py hands over): roi_crop.c:17-27 reads the strides */
static inline long THFloatTensor_stride(THFloatTensor *t, int d) {
    long s = 1;
    for (int k = t->ndim - 1; k > d; --k) s *= t->size[k];
    return s;
}
static inline THFloatStorage *THFloatTensor_storage(THFloatTensor *t) { return &t->storage; }
static inline void THFloatStorage_fill(THFloatStorage *s, float v) {
    for (long i = 0; i < s->numel; ++i) s->data[i] = v;
}

#endif
