// sqlrs_b200 — shared host-side plumbing of the CUDA library: status/error type, CUDA error
// checks, the launch counter behind sqlrs_kernel_launches(), small helpers.
#pragma once
#include <cuda_runtime.h>

#include <time.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sqlrs_b200.h"

namespace sq {

// ExecutorError (reference src/executor/mod.rs:67-85) carried as status code + message.
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
[[noreturn]] inline void fail(int code, const std::string& m) { throw Error(code, m); }

#define SQ_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      ::sq::fail(SQLRS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));             \
  } while (0)

#ifdef __CUDACC__
#define SQ_HD __host__ __device__
#else
#define SQ_HD
#endif

// stream-ordered scratch memory from the library's block cache (device.cpp): freed blocks are reused by later
// allocations of the same size class on the same stream without a driver call
void* scratch_alloc(size_t bytes, cudaStream_t stream);
void scratch_free(void* p, cudaStream_t stream);

extern std::atomic<int64_t> g_kernel_launches;
inline void count_launch(int64_t n = 1) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }

// SQLRS_B200_TRACE=1: wall-clock phase timings on stderr (each phase end synchronises the stream, so the
// numbers attribute device work to the phase that enqueued it; tracing slows the run down)
struct Trace {
  const char* name;
  cudaStream_t stream;
  bool on;
  double t0 = 0;
  static bool enabled() {
    static int e = -1;
    if (e < 0) e = getenv("SQLRS_B200_TRACE") ? 1 : 0;
    return e == 1;
  }
  static double now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  }
  Trace(const char* n, cudaStream_t s) : name(n), stream(s), on(enabled()) {
    if (on) t0 = now();
  }
  ~Trace() {
    if (!on) return;
    const double t1 = now();
    cudaStreamSynchronize(stream);
    const double t2 = now();
    fprintf(stderr, "[sqlrs trace] %-28s host %8.3f ms  +drain %8.3f ms\n", name, t1 - t0, t2 - t1);
  }
};

// SQLRS_FLAG_KERNEL_EVENTS: CUDA events around the hot kernels, recorded on the launching stream and resolved later
// (kernel_events_collect, after the caller synchronised) — no synchronisation is added to the run.  device.cpp.
void kernel_event_begin(cudaStream_t stream, const char* name);
void kernel_event_end(cudaStream_t stream);
std::string kernel_events_collect_json();  // {"name": {"ms": total, "launches": n}, ...}; clears the record
struct KernelEvent {
  cudaStream_t stream;
  bool on;
  KernelEvent(int flags, cudaStream_t s, const char* name) : stream(s), on((flags & SQLRS_FLAG_KERNEL_EVENTS) != 0) {
    if (on) kernel_event_begin(stream, name);
  }
  ~KernelEvent() {
    if (on) kernel_event_end(stream);
  }
};

// device copy of the string pool's byte-wise rank table (int32 per pool id) on the current device, brought up to date with the
// pool on `stream` (device.cpp); what Utf8 ordering comparisons inside expressions read (csrc/jit/strrank.cuh)
const void* string_rank_table(cudaStream_t stream);

inline const char* dtype_name(int dt) {
  switch (dt) {
    case SQLRS_DT_NULL: return "Null";
    case SQLRS_DT_BOOL: return "Boolean";
    case SQLRS_DT_INT32: return "Int32";
    case SQLRS_DT_INT64: return "Int64";
    case SQLRS_DT_FLOAT64: return "Float64";
    case SQLRS_DT_UTF8: return "Utf8";
  }
  return "?";
}
inline bool is_numeric(int dt) { return dt == SQLRS_DT_INT32 || dt == SQLRS_DT_INT64 || dt == SQLRS_DT_FLOAT64; }
// bytes per value of the fixed-width device layout (Boolean is bit-packed: 0 here)
inline int dtype_width(int dt) {
  switch (dt) {
    case SQLRS_DT_INT32: return 4;
    case SQLRS_DT_INT64:
    case SQLRS_DT_FLOAT64:
    case SQLRS_DT_UTF8: return 8;  // Utf8 on the device: the string's id in the library's string pool (device.hpp)
  }
  return 0;
}
inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t bitmap_words(int64_t n) { return div_up(n, 32); }  // device bitmaps are u32-word granular

// thrown by an operator whose state was sized from a previous run's numbers and turned out too small; the plan re-runs
// without the hint (Plan::run_validated)
struct RetrySizingError {};

struct Options {
  int count_mode = SQLRS_COUNT_REFERENCE_OVERWRITE;
  int match_mode = SQLRS_MATCH_HASH_ONLY;
  int device_id = -1;
  int flags = 0;
  cudaStream_t stream = nullptr;  // user stream, or nullptr = library-owned
};
inline Options copy_options(const sqlrs_options* o) {
  Options r;
  if (o) {
    r.count_mode = o->count_mode;
    r.match_mode = o->match_mode;
    r.device_id = o->device_id;
    r.flags = o->flags;
    r.stream = (cudaStream_t)o->stream;
  }
  return r;
}

}  // namespace sq
