/*
 * Arrow C Data Interface + Arrow C Device Data Interface struct definitions.
 *
 * These are the ABI-stable structs published by the Apache Arrow specification
 * (https://arrow.apache.org/docs/format/CDataInterface.html and
 * CDeviceDataInterface.html).  arrow-rs 28 — the crate every sqlrs operator is
 * built on (reference Cargo.toml:10) — exposes the same layout as
 * `arrow::ffi::FFI_ArrowArray` / `FFI_ArrowSchema`, which is what makes a
 * RecordBatch hand-off between the Rust host and this library zero-copy.
 *
 * Guarded with the spec's own macros so that including another copy is harmless.
 */
#ifndef SQLRS_B200_ARROW_C_DATA_H
#define SQLRS_B200_ARROW_C_DATA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE

#define ARROW_FLAG_DICTIONARY_ORDERED 1
#define ARROW_FLAG_NULLABLE 2
#define ARROW_FLAG_MAP_KEYS_SORTED 4

struct ArrowSchema {
  const char* format;
  const char* name;
  const char* metadata;
  int64_t flags;
  int64_t n_children;
  struct ArrowSchema** children;
  struct ArrowSchema* dictionary;
  void (*release)(struct ArrowSchema*);
  void* private_data;
};

struct ArrowArray {
  int64_t length;
  int64_t null_count;
  int64_t offset;
  int64_t n_buffers;
  int64_t n_children;
  const void** buffers;
  struct ArrowArray** children;
  struct ArrowArray* dictionary;
  void (*release)(struct ArrowArray*);
  void* private_data;
};

#endif /* ARROW_C_DATA_INTERFACE */

#ifndef ARROW_C_DEVICE_DATA_INTERFACE
#define ARROW_C_DEVICE_DATA_INTERFACE

typedef int32_t ArrowDeviceType;
#define ARROW_DEVICE_CPU 1
#define ARROW_DEVICE_CUDA 2
#define ARROW_DEVICE_CUDA_HOST 3

struct ArrowDeviceArray {
  struct ArrowArray array;   /* buffers[] hold DEVICE pointers when device_type == CUDA */
  int64_t device_id;
  ArrowDeviceType device_type;
  void* sync_event;          /* cudaEvent_t* the consumer must wait on, or NULL */
  int64_t reserved[3];
};

#endif /* ARROW_C_DEVICE_DATA_INTERFACE */

#ifdef __cplusplus
}
#endif

#endif /* SQLRS_B200_ARROW_C_DATA_H */
