// dlpack_export.cu -- hands borrowed device buffers to the consumer as DLPack tensors
// (DLPack ABI v0.8 layout: DLManagedTensor).  The handle that owns the memory outlives the
// tensor by contract (include/gymcity_b200.h), so the deleter only frees the descriptor -- or, with
// gcb_dlpack_wrap_owned, first tells the owner that this view is gone, so that the owner can keep the memory until then.
#include <stdint.h>
#include <stdlib.h>
#include "../../include/gymcity_b200.h"

namespace {
struct DLDevice { int32_t device_type; int32_t device_id; };
struct DLDataType { uint8_t code; uint8_t bits; uint16_t lanes; };
struct DLTensor {
    void *data; DLDevice device; int32_t ndim; DLDataType dtype;
    int64_t *shape; int64_t *strides; uint64_t byte_offset;
};
struct DLManagedTensor {
    DLTensor dl_tensor; void *manager_ctx; void (*deleter)(DLManagedTensor *);
};
struct Owned { DLManagedTensor t; int64_t shape[8]; void (*release)(void *); void *release_ctx; };
void drop(DLManagedTensor *t)
{
    Owned *o = (Owned *)t->manager_ctx;
    if (o->release) o->release(o->release_ctx);
    free(o);
}
}

extern "C" void *gcb_dlpack_wrap(void *dev_ptr, int dtype_code, int bits, int ndim,
                                 const int64_t *shape, int device)
{
    if (!dev_ptr || ndim < 1 || ndim > 8) return NULL;
    Owned *o = (Owned *)calloc(1, sizeof(Owned));
    for (int i = 0; i < ndim; i++) o->shape[i] = shape[i];
    o->t.dl_tensor.data = dev_ptr;
    o->t.dl_tensor.device.device_type = 2;          /* kDLCUDA */
    o->t.dl_tensor.device.device_id = device;
    o->t.dl_tensor.ndim = ndim;
    o->t.dl_tensor.dtype.code = (uint8_t)dtype_code; /* 0 int, 1 uint, 2 float */
    o->t.dl_tensor.dtype.bits = (uint8_t)bits;
    o->t.dl_tensor.dtype.lanes = 1;
    o->t.dl_tensor.shape = o->shape;
    o->t.dl_tensor.strides = NULL;
    o->t.manager_ctx = o;
    o->t.deleter = drop;
    return &o->t;
}

extern "C" void *gcb_dlpack_wrap_owned(void *dev_ptr, int dtype_code, int bits, int ndim, const int64_t *shape, int device,
                                       void (*release)(void *), void *release_ctx)
{
    void *mt = gcb_dlpack_wrap(dev_ptr, dtype_code, bits, ndim, shape, device);
    if (!mt) return NULL;
    Owned *o = (Owned *)((DLManagedTensor *)mt)->manager_ctx;
    o->release = release;
    o->release_ctx = release_ctx;
    return mt;
}
