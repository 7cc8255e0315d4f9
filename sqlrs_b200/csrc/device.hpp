// sqlrs_b200 — device-resident columns / batches and the Arrow C Data Interface bridge.
//
// Layout in HBM (DESIGN.md "Data layout"): one contiguous value buffer per column
// (int64/float64: 8 B, int32: 4 B per row; Boolean: bit-packed, LSB first — Arrow's own
// layouts, so a device-resident Arrow array is used in place) plus an optional validity bitmap
// (bit-packed, u32-word granular, bit offset 0, bits past `n` zero).  A column without nulls
// has no bitmap at all, and kernels are specialised on that.
#pragma once
#include <functional>
#include <list>

#include "common.hpp"

namespace sq {

// Execution context of one handle: device + stream + stream-ordered allocations + deferred
// release of moved-in host arrays (their buffers must outlive the async H2D copy).
struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  bool offline = false;  // device_id == -2: code generation only, no CUDA calls (diagnostics / build check)
  struct Pending {
    cudaEvent_t ev;
    std::function<void()> fn;
  };
  std::list<Pending> pending;

  explicit Ctx(const Options& o);
  ~Ctx();
  Ctx(const Ctx&) = delete;
  Ctx& operator=(const Ctx&) = delete;
  void activate() { SQ_CUDA(cudaSetDevice(device)); }
  void defer(std::function<void()> fn);  // run fn once everything enqueued so far has completed
  void reap();                           // run the deferred functions whose work is done
  void sync();                           // cudaStreamSynchronize + run all deferred functions
};

// Utf8 on the device: every distinct string the library has seen lives once in a process-wide,
// append-only pool; a Utf8 column in HBM is an 8-byte column of pool ids (dictionary encoding on ingest).  Equal strings
// have equal ids, so group keys, join keys, DISTINCT and = / <> work on the ids as they stand; what depends on the ORDER
// of strings (ORDER BY, MIN / MAX, < ... >=) goes through a rank table — the byte-wise rank of every id, as Rust's `str`
// ordering (min_string / max_string, arrow sort) compares — built on the host when an operator needs it.
// Result columns are decoded back to Arrow Utf8 on export.
class StringPool {
 public:
  static StringPool& instance();
  int64_t intern(const char* data, size_t len);
  std::string get(int64_t id) const;
  int64_t size() const;
  std::vector<int32_t> ranks() const;  // ranks()[id] = position of string `id` in byte-wise order

 private:
  struct Impl;
  StringPool();
  Impl* impl_;
};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaStream_t stream = nullptr;
  int device = 0;
  ~DevBuf();
};
using BufPtr = std::shared_ptr<DevBuf>;
BufPtr dev_alloc(Ctx& ctx, size_t bytes);        // stream-ordered (cudaMallocAsync)
BufPtr dev_alloc_zero(Ctx& ctx, size_t bytes);

struct Field {
  std::string name;
  int dtype = SQLRS_DT_NULL;
  bool nullable = true;
};

struct DCol {
  int dtype = SQLRS_DT_NULL;
  int64_t n = 0;
  const void* data = nullptr;      // device pointer (nullptr for dtype Null)
  const uint32_t* valid = nullptr; // device bitmap or nullptr (= no nulls)
  int64_t null_count = 0;          // -1 = unknown (bitmap present)
  std::shared_ptr<void> keep_data, keep_valid;  // owners (DevBuf or foreign release hook)
  bool all_null() const { return dtype == SQLRS_DT_NULL; }
};

struct DBatch {
  std::vector<Field> fields;
  std::vector<DCol> cols;
  int64_t n = 0;
};

int dtype_from_format(const char* fmt);
const char* format_of_dtype(int dt);
std::vector<Field> import_fields(const ArrowSchema* schema);

// Host Arrow struct array -> device batch.  MOVES `array` (ABI ownership rule): the struct is
// copied, the caller's copy is marked released, the original release callback runs once the
// H2D copies have completed.
DBatch import_batch_host(Ctx& ctx, ArrowArray* array, const ArrowSchema* schema);
// Device-resident Arrow struct array (buffers[] are device pointers) -> zero-copy batch.
DBatch import_batch_device(Ctx& ctx, ArrowDeviceArray* array, const ArrowSchema* schema);
// Device batch -> host Arrow struct array with our release callbacks (synchronises ctx).
void export_batch_host(Ctx& ctx, const DBatch& b, ArrowArray* out, ArrowSchema* out_schema);
void export_schema(const std::vector<Field>& fields, ArrowSchema* out);

// result columns assembled on the host (aggregate finalisation) -> host Arrow struct array
struct HostCol {
  int dtype = SQLRS_DT_NULL;
  std::vector<int64_t> i;   // Int32 / Int64 / Boolean (0/1)
  std::vector<double> f;    // Float64
  std::vector<uint8_t> valid;  // empty = no nulls, else one byte per row
  int64_t size() const { return dtype == SQLRS_DT_FLOAT64 ? (int64_t)f.size() : (int64_t)i.size(); }
};
void export_host_columns(const std::vector<Field>& fields, const std::vector<HostCol>& cols, int64_t n, ArrowArray* out,
                         ArrowSchema* out_schema);

// null_count of a column (computes and caches it when unknown); synchronises when it must count
int64_t null_count_of(Ctx& ctx, DCol& c);

// column helpers used by the operators
DCol make_col(Ctx& ctx, int dtype, int64_t n, bool with_validity);  // uninitialised values
DCol null_col(Ctx& ctx, int dtype, int64_t n);                      // arrow new_null_array
inline void* col_data(DCol& c) { return const_cast<void*>(c.data); }
inline uint32_t* col_valid(DCol& c) { return const_cast<uint32_t*>(c.valid); }
size_t col_value_bytes(int dtype, int64_t n);

}  // namespace sq
