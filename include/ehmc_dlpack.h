/*
 * Minimal DLPack (v0.8 ABI) structure definitions used by the ehmc C-ABI.
 *
 * Written from the public DLPack specification so the library has no external
 * header dependency.  A torch tensor is handed over by calling
 * ``torch.utils.dlpack.to_dlpack(t)`` and passing the ``DLManagedTensor*`` held by
 * the returned PyCapsule; ``DLManagedTensor`` starts with its ``DLTensor``, which is
 * what the entry points of ``ehmc.h`` take (borrowed, never retained).
 */
#ifndef EHMC_DLPACK_H_
#define EHMC_DLPACK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  kDLCPU = 1,
  kDLCUDA = 2,
  kDLCUDAHost = 3,
  kDLCUDAManaged = 13
} DLDeviceType;

typedef struct {
  int32_t device_type; /* DLDeviceType */
  int32_t device_id;
} DLDevice;

typedef enum { kDLInt = 0, kDLUInt = 1, kDLFloat = 2, kDLBfloat = 4, kDLBool = 6 } DLDataTypeCode;

typedef struct {
  uint8_t code;
  uint8_t bits;
  uint16_t lanes;
} DLDataType;

typedef struct {
  void* data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t* shape;
  int64_t* strides; /* in elements; NULL = compact row-major */
  uint64_t byte_offset;
} DLTensor;

typedef struct DLManagedTensor {
  DLTensor dl_tensor;
  void* manager_ctx;
  void (*deleter)(struct DLManagedTensor* self);
} DLManagedTensor;

#ifdef __cplusplus
}
#endif
#endif /* EHMC_DLPACK_H_ */
