// sqlrs_b200 — device memory, device batches and the Arrow C Data bridge (see device.hpp).
#include "device.hpp"

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <unordered_map>

#include "kernels_aot.hpp"

namespace sq {

std::atomic<int64_t> g_kernel_launches{0};

static void flush_stream_cache(int device, cudaStream_t stream);  // block cache below

// ------------------------------------------------------------------ string pool (Utf8 columns are columns of pool ids)
struct StringPool::Impl {
  mutable std::mutex mu;
  std::unordered_map<std::string, int64_t> ids;
  std::vector<const std::string*> by_id;  // points into `ids` (node-based: stable addresses)
};
StringPool::StringPool() : impl_(new Impl()) {}
StringPool& StringPool::instance() {
  static StringPool pool;
  return pool;
}
int64_t StringPool::intern(const char* data, size_t len) {
  std::lock_guard<std::mutex> lock(impl_->mu);
  auto it = impl_->ids.emplace(std::string(data, len), (int64_t)impl_->by_id.size());
  if (it.second) impl_->by_id.push_back(&it.first->first);
  return it.first->second;
}
std::string StringPool::get(int64_t id) const {
  std::lock_guard<std::mutex> lock(impl_->mu);
  if (id < 0 || id >= (int64_t)impl_->by_id.size()) fail(SQLRS_ERR_INTERNAL, "string pool: id out of range");
  return *impl_->by_id[(size_t)id];
}
int64_t StringPool::size() const {
  std::lock_guard<std::mutex> lock(impl_->mu);
  return (int64_t)impl_->by_id.size();
}
std::vector<int32_t> StringPool::ranks() const {
  std::lock_guard<std::mutex> lock(impl_->mu);
  const size_t n = impl_->by_id.size();
  std::vector<int32_t> order(n), rank(n);
  for (size_t i = 0; i < n; i++) order[i] = (int32_t)i;
  std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return *impl_->by_id[(size_t)a] < *impl_->by_id[(size_t)b]; });  // std::string <: byte-wise
  for (size_t r = 0; r < n; r++) rank[(size_t)order[r]] = (int32_t)r;
  return rank;
}

// One rank table per device, rebuilt when the pool has grown since the last call.  The upload is ordered on the caller's stream
// and waited for (the staging vector is reused); kernels already enqueued on that stream read the previous contents first.
// A table that outgrows its buffer moves to a new one; old buffers are kept (a kernel on another stream may still read them).
const void* string_rank_table(cudaStream_t stream) {
  struct Entry {
    void* buf = nullptr;
    size_t capacity = 0;
    int64_t ranked = -1;
  };
  static std::mutex mu;
  static std::map<int, Entry> tables;
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  SQ_CUDA(cudaGetDevice(&dev));
  Entry& e = tables[dev];
  const int64_t n = StringPool::instance().size();
  if (e.ranked == n && e.buf) return e.buf;
  const std::vector<int32_t> ranks = StringPool::instance().ranks();
  if (!e.buf || ranks.size() > e.capacity) {
    e.capacity = std::max<size_t>(2 * ranks.size(), 1024);
    SQ_CUDA(cudaMalloc(&e.buf, e.capacity * 4));
  }
  if (!ranks.empty()) {
    SQ_CUDA(cudaMemcpyAsync(e.buf, ranks.data(), ranks.size() * 4, cudaMemcpyHostToDevice, stream));
    SQ_CUDA(cudaStreamSynchronize(stream));
  }
  e.ranked = (int64_t)ranks.size();
  return e.buf;
}

// ------------------------------------------------------------------ kernel events (SQLRS_FLAG_KERNEL_EVENTS)
namespace {
struct EventRec {
  std::string name;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  bool closed = false;
};
std::mutex g_events_mu;
std::vector<EventRec> g_events;
}  // namespace
void kernel_event_begin(cudaStream_t stream, const char* name) {
  EventRec r;
  r.name = name;
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, stream);
  std::lock_guard<std::mutex> lock(g_events_mu);
  g_events.push_back(r);
}
void kernel_event_end(cudaStream_t stream) {  // closes the innermost open record (scopes nest)
  std::lock_guard<std::mutex> lock(g_events_mu);
  for (size_t i = g_events.size(); i-- > 0;) {
    if (g_events[i].closed) continue;
    cudaEventRecord(g_events[i].e1, stream);
    g_events[i].closed = true;
    break;
  }
}
std::string kernel_events_collect_json() {
  std::lock_guard<std::mutex> lock(g_events_mu);
  std::map<std::string, std::pair<double, int64_t>> acc;
  for (EventRec& r : g_events) {
    if (r.closed) {
      float ms = 0.f;
      if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
        auto& a = acc[r.name];
        a.first += ms;
        a.second += 1;
      }
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_events.clear();
  cudaGetLastError();
  std::string out = "{";
  bool first = true;
  for (auto& kv : acc) {
    char buf[256];
    std::snprintf(buf, sizeof buf, "%s\"%s\": {\"ms\": %.6f, \"launches\": %lld}", first ? "" : ", ", kv.first.c_str(), kv.second.first, (long long)kv.second.second);
    out += buf;
    first = false;
  }
  return out + "}";
}

// ------------------------------------------------------------------ Ctx
static void tune_pool(int device) {
  static std::mutex mu;
  static bool done[64] = {false};
  std::lock_guard<std::mutex> lock(mu);
  if (device < 0 || device >= 64 || done[device]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t threshold = ~0ULL;  // keep freed blocks in the pool: operators re-allocate the same sizes every batch
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  }
  done[device] = true;
}

Ctx::Ctx(const Options& o) {
  if (o.device_id == -2) {
    offline = true;
    return;
  }
  if (o.device_id >= 0) device = o.device_id;
  else SQ_CUDA(cudaGetDevice(&device));
  SQ_CUDA(cudaSetDevice(device));
  tune_pool(device);
  if (o.stream) {
    stream = o.stream;
  } else {
    SQ_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    own_stream = true;
  }
}

Ctx::~Ctx() {
  if (offline) return;
  cudaSetDevice(device);
  cudaStreamSynchronize(stream);
  for (auto& p : pending) {
    p.fn();
    cudaEventDestroy(p.ev);
  }
  pending.clear();
  if (own_stream) {
    flush_stream_cache(device, stream);  // cached blocks are keyed by the stream that is about to disappear
    cudaStreamSynchronize(stream);
    cudaStreamDestroy(stream);
  }
}

void Ctx::defer(std::function<void()> fn) {
  Pending p;
  SQ_CUDA(cudaEventCreateWithFlags(&p.ev, cudaEventDisableTiming));
  SQ_CUDA(cudaEventRecord(p.ev, stream));
  p.fn = std::move(fn);
  pending.push_back(std::move(p));
}

void Ctx::reap() {
  while (!pending.empty()) {
    if (cudaEventQuery(pending.front().ev) != cudaSuccess) {
      cudaGetLastError();
      break;
    }
    pending.front().fn();
    cudaEventDestroy(pending.front().ev);
    pending.pop_front();
  }
}

void Ctx::sync() {
  SQ_CUDA(cudaStreamSynchronize(stream));
  for (auto& p : pending) {
    p.fn();
    cudaEventDestroy(p.ev);
  }
  pending.clear();
}

// ------------------------------------------------------------------ block caches
// A plan re-allocates the same buffer sizes on the same stream every run.  Going to the driver's stream-ordered pool
// for each of them costs 2-5 us per call and — measured on Q3' SF100 — sporadic 0.5 s stalls when the pool has to
// re-map physical memory to satisfy a multi-GB request (profiles/r01e_q3_sf100_trace.txt).  Freed blocks are therefore
// kept per (device, stream, size class) and handed out again in stream order: a block's next user is enqueued on the
// same stream after its previous user, so no synchronisation is needed.  Size classes: powers of two up to 1 MiB,
// then 16 steps per octave (<= 12.5 % slack).
namespace {

size_t size_class(size_t bytes) {
  if (bytes <= 512) return 512;
  size_t p = 512;
  while (p < bytes) p <<= 1;
  if (p <= (1u << 20)) return p;
  const size_t step = p >> 4;
  return (bytes + step - 1) / step * step;
}

struct DeviceCache {
  std::mutex mu;
  std::map<std::tuple<int, cudaStream_t, size_t>, std::vector<void*>> free_blocks;
  std::unordered_map<void*, size_t> live;  // blocks handed out through scratch_alloc: their class
  size_t cached = 0;
  size_t cap = 64ULL << 30;
  DeviceCache() {
    if (const char* e = std::getenv("SQLRS_B200_CACHE_GB")) cap = (size_t)std::max(0, atoi(e)) << 30;
  }
  void* get(int device, cudaStream_t stream, size_t cls) {
    {
      std::lock_guard<std::mutex> lock(mu);
      auto it = free_blocks.find(std::make_tuple(device, stream, cls));
      if (it != free_blocks.end() && !it->second.empty()) {
        void* p = it->second.back();
        it->second.pop_back();
        cached -= cls;
        return p;
      }
    }
    void* p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, cls, stream);
    if (e == cudaErrorMemoryAllocation) {  // give everything cached back to the driver and retry once
      cudaGetLastError();
      flush(-1, nullptr, true);
      e = cudaMallocAsync(&p, cls, stream);
    }
    if (e != cudaSuccess) fail(SQLRS_ERR_CUDA, std::string("cudaMallocAsync(") + std::to_string(cls) + " bytes): " + cudaGetErrorString(e));
    return p;
  }
  void put(int device, cudaStream_t stream, void* p, size_t cls) {
    {
      std::lock_guard<std::mutex> lock(mu);
      if (cached + cls <= cap) {
        free_blocks[std::make_tuple(device, stream, cls)].push_back(p);
        cached += cls;
        return;
      }
    }
    cudaFreeAsync(p, stream);
  }
  // returns cached blocks to the driver: those of one (device, stream), or all of them
  void flush(int device, cudaStream_t stream, bool all) {
    std::lock_guard<std::mutex> lock(mu);
    for (auto it = free_blocks.begin(); it != free_blocks.end();) {
      if (all || (std::get<0>(it->first) == device && std::get<1>(it->first) == stream)) {
        for (void* p : it->second) {
          cudaFreeAsync(p, std::get<1>(it->first));
          cached -= std::get<2>(it->first);
        }
        it = free_blocks.erase(it);
      } else {
        ++it;
      }
    }
  }
};
DeviceCache& device_cache() {
  static DeviceCache* c = new DeviceCache();  // never destroyed: release callbacks may run during process exit
  return *c;
}

// pinned host blocks for results leaving the device: D2H at PCIe speed straight into the exported Arrow buffers
// (measured: 36 MB of Q3' SF100 groups took 17 ms into pageable memory)
struct PinnedCache {
  std::mutex mu;
  std::map<size_t, std::vector<void*>> free_blocks;
  std::unordered_map<void*, size_t> live;
  size_t cached = 0;
  size_t cap = 8ULL << 30;
  void* get(size_t bytes) {
    const size_t cls = size_class(bytes);
    void* p = nullptr;
    {
      std::lock_guard<std::mutex> lock(mu);
      auto it = free_blocks.find(cls);
      if (it != free_blocks.end() && !it->second.empty()) {
        p = it->second.back();
        it->second.pop_back();
        cached -= cls;
      }
    }
    if (!p && cudaHostAlloc(&p, cls, cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;  // the caller falls back to pageable memory
    }
    std::lock_guard<std::mutex> lock(mu);
    live[p] = cls;
    return p;
  }
  bool put(void* p) {  // false: not one of ours
    size_t cls = 0;
    {
      std::lock_guard<std::mutex> lock(mu);
      auto it = live.find(p);
      if (it == live.end()) return false;
      cls = it->second;
      live.erase(it);
      if (cached + cls <= cap) {
        free_blocks[cls].push_back(p);
        cached += cls;
        return true;
      }
    }
    cudaFreeHost(p);
    return true;
  }
};
PinnedCache& pinned_cache() {
  static PinnedCache* c = new PinnedCache();
  return *c;
}

}  // namespace

static void flush_stream_cache(int device, cudaStream_t stream) { device_cache().flush(device, stream, false); }

void* scratch_alloc(size_t bytes, cudaStream_t stream) {
  int device = 0;
  SQ_CUDA(cudaGetDevice(&device));
  DeviceCache& c = device_cache();
  const size_t cls = size_class(bytes);
  void* p = c.get(device, stream, cls);
  std::lock_guard<std::mutex> lock(c.mu);
  c.live[p] = cls;
  return p;
}
void scratch_free(void* p, cudaStream_t stream) {
  if (!p) return;
  int device = 0;
  cudaGetDevice(&device);
  DeviceCache& c = device_cache();
  size_t cls = 0;
  {
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.live.find(p);
    if (it == c.live.end()) return;
    cls = it->second;
    c.live.erase(it);
  }
  c.put(device, stream, p, cls);
}

// ------------------------------------------------------------------ buffers
DevBuf::~DevBuf() {
  if (p) device_cache().put(device, stream, p, size_class(bytes ? bytes : 16));
}

BufPtr dev_alloc(Ctx& ctx, size_t bytes) {
  auto b = std::make_shared<DevBuf>();
  b->bytes = bytes;
  b->stream = ctx.stream;
  b->device = ctx.device;
  b->p = device_cache().get(ctx.device, ctx.stream, size_class(bytes ? bytes : 16));
  return b;
}
BufPtr dev_alloc_zero(Ctx& ctx, size_t bytes) {
  BufPtr b = dev_alloc(ctx, bytes);
  SQ_CUDA(cudaMemsetAsync(b->p, 0, bytes ? bytes : 16, ctx.stream));
  return b;
}

size_t col_value_bytes(int dtype, int64_t n) {
  if (dtype == SQLRS_DT_BOOL) return (size_t)bitmap_words(n) * 4;
  return (size_t)dtype_width(dtype) * (size_t)n;
}

DCol make_col(Ctx& ctx, int dtype, int64_t n, bool with_validity) {
  DCol c;
  c.dtype = dtype;
  c.n = n;
  if (dtype == SQLRS_DT_NULL) {
    c.null_count = n;
    return c;
  }
  BufPtr d = dtype == SQLRS_DT_BOOL ? dev_alloc_zero(ctx, col_value_bytes(dtype, n)) : dev_alloc(ctx, col_value_bytes(dtype, n));
  c.data = d->p;
  c.keep_data = d;
  if (with_validity) {
    BufPtr v = dev_alloc_zero(ctx, (size_t)bitmap_words(n) * 4);
    c.valid = (const uint32_t*)v->p;
    c.keep_valid = v;
    c.null_count = -1;
  }
  return c;
}

DCol null_col(Ctx& ctx, int dtype, int64_t n) {
  DCol c;
  c.dtype = dtype;
  c.n = n;
  c.null_count = n;
  if (dtype == SQLRS_DT_NULL) return c;
  BufPtr d = dev_alloc_zero(ctx, col_value_bytes(dtype, n));
  BufPtr v = dev_alloc_zero(ctx, (size_t)bitmap_words(n) * 4);
  c.data = d->p;
  c.keep_data = d;
  c.valid = (const uint32_t*)v->p;
  c.keep_valid = v;
  if (n == 0) {
    c.valid = nullptr;
    c.keep_valid.reset();
    c.null_count = 0;
  }
  return c;
}

int64_t null_count_of(Ctx& ctx, DCol& c) {
  if (c.dtype == SQLRS_DT_NULL) return c.n;
  if (!c.valid) return 0;
  if (c.null_count >= 0) return c.null_count;
  BufPtr cnt = dev_alloc_zero(ctx, 8);
  launch_count_bits(c.valid, c.n, (unsigned long long*)cnt->p, ctx.stream);
  unsigned long long set = 0;
  SQ_CUDA(cudaMemcpyAsync(&set, cnt->p, 8, cudaMemcpyDeviceToHost, ctx.stream));
  SQ_CUDA(cudaStreamSynchronize(ctx.stream));
  c.null_count = c.n - (int64_t)set;
  return c.null_count;
}

// ------------------------------------------------------------------ schema
int dtype_from_format(const char* fmt) {
  if (!fmt) fail(SQLRS_ERR_INVALID_ARG, "ArrowSchema.format is NULL");
  std::string f(fmt);
  if (f == "n") return SQLRS_DT_NULL;
  if (f == "b") return SQLRS_DT_BOOL;
  if (f == "i") return SQLRS_DT_INT32;
  if (f == "l") return SQLRS_DT_INT64;
  if (f == "g") return SQLRS_DT_FLOAT64;
  if (f == "u") return SQLRS_DT_UTF8;
  fail(SQLRS_ERR_UNSUPPORTED, "unsupported Arrow format '" + f + "' (v1 type universe: n,b,i,l,g,u)");
}
const char* format_of_dtype(int dt) {
  switch (dt) {
    case SQLRS_DT_NULL: return "n";
    case SQLRS_DT_BOOL: return "b";
    case SQLRS_DT_INT32: return "i";
    case SQLRS_DT_INT64: return "l";
    case SQLRS_DT_FLOAT64: return "g";
    case SQLRS_DT_UTF8: return "u";
  }
  return "n";
}

std::vector<Field> import_fields(const ArrowSchema* schema) {
  if (!schema || !schema->format || std::string(schema->format) != "+s")
    fail(SQLRS_ERR_INVALID_ARG, "expected a struct ('+s') schema describing a RecordBatch");
  std::vector<Field> out;
  for (int64_t c = 0; c < schema->n_children; c++) {
    const ArrowSchema* cs = schema->children[c];
    Field f;
    f.name = cs->name ? cs->name : "";
    f.dtype = dtype_from_format(cs->format);
    f.nullable = (cs->flags & ARROW_FLAG_NULLABLE) != 0;
    out.push_back(f);
  }
  return out;
}

// ------------------------------------------------------------------ import (host)
namespace {

struct MovedArray {
  ArrowArray a;
  ~MovedArray() {
    if (a.release) a.release(&a);
  }
};

// n bits starting at bit `off` of src -> freshly allocated device bitmap (u32 words, tail zero)
BufPtr upload_bits(Ctx& ctx, const uint8_t* src, int64_t off, int64_t n, std::vector<std::shared_ptr<std::vector<uint8_t>>>& staging) {
  size_t bytes = (size_t)bitmap_words(n) * 4;
  BufPtr d = dev_alloc_zero(ctx, bytes);
  if (n == 0) return d;
  size_t nb = (size_t)((n + 7) / 8);
  if ((off & 7) == 0 && (n & 7) == 0) {
    SQ_CUDA(cudaMemcpyAsync(d->p, src + (off >> 3), nb, cudaMemcpyHostToDevice, ctx.stream));
    return d;
  }
  auto tmp = std::make_shared<std::vector<uint8_t>>(nb, 0);
  const int sh = (int)(off & 7);
  const uint8_t* s = src + (off >> 3);
  const size_t src_bytes = (size_t)((sh + n + 7) / 8);
  for (size_t i = 0; i < nb; i++) {
    unsigned lo = s[i], hi = (i + 1 < src_bytes) ? s[i + 1] : 0;
    (*tmp)[i] = (uint8_t)(((lo >> sh) | (hi << (8 - sh))) & 0xff);
  }
  if (n & 7) (*tmp)[nb - 1] &= (uint8_t)((1u << (n & 7)) - 1u);
  SQ_CUDA(cudaMemcpyAsync(d->p, tmp->data(), nb, cudaMemcpyHostToDevice, ctx.stream));
  staging.push_back(tmp);
  return d;
}

DCol import_column_host(Ctx& ctx, const ArrowArray* a, int dtype, int64_t parent_offset, int64_t length,
                        std::vector<std::shared_ptr<std::vector<uint8_t>>>& staging) {
  DCol col;
  col.dtype = dtype;
  col.n = length;
  if (dtype == SQLRS_DT_NULL) {
    col.null_count = length;
    return col;
  }
  if (a->length < parent_offset + length) fail(SQLRS_ERR_INVALID_ARG, "child array shorter than batch");
  if (a->n_buffers < 2) fail(SQLRS_ERR_INVALID_ARG, "primitive array needs 2 buffers");
  const int64_t off = a->offset + parent_offset;
  const uint8_t* validity = (const uint8_t*)a->buffers[0];
  if (dtype == SQLRS_DT_UTF8) {  // dictionary-encode on ingest: the column becomes its strings' pool ids
    if (a->n_buffers < 3) fail(SQLRS_ERR_INVALID_ARG, "utf8 array needs 3 buffers");
    const int32_t* offsets = (const int32_t*)a->buffers[1];
    const char* chars = (const char*)a->buffers[2];
    auto ids = std::make_shared<std::vector<uint8_t>>((size_t)std::max<int64_t>(length, 1) * 8);
    int64_t* out = (int64_t*)ids->data();
    StringPool& pool = StringPool::instance();
    for (int64_t r = 0; r < length; r++) {
      const bool is_null = validity && a->null_count != 0 && !((validity[(off + r) >> 3] >> ((off + r) & 7)) & 1);
      const int32_t b = offsets[off + r], e = offsets[off + r + 1];
      out[r] = is_null ? 0 : pool.intern(chars ? chars + b : "", (size_t)(e - b));
    }
    if (validity && a->null_count != 0 && length > 0) {
      BufPtr v = upload_bits(ctx, validity, off, length, staging);
      col.valid = (const uint32_t*)v->p;
      col.keep_valid = v;
      col.null_count = a->null_count > 0 && parent_offset == 0 && a->length == length ? a->null_count : -1;
    }
    BufPtr d = dev_alloc(ctx, (size_t)std::max<int64_t>(length, 1) * 8);
    if (length) SQ_CUDA(cudaMemcpyAsync(d->p, ids->data(), (size_t)length * 8, cudaMemcpyHostToDevice, ctx.stream));
    staging.push_back(ids);
    col.data = d->p;
    col.keep_data = d;
    return col;
  }
  if (validity && a->null_count != 0 && length > 0) {
    BufPtr v = upload_bits(ctx, validity, off, length, staging);
    col.valid = (const uint32_t*)v->p;
    col.keep_valid = v;
    col.null_count = a->null_count > 0 && parent_offset == 0 && a->length == length ? a->null_count : -1;
  }
  if (dtype == SQLRS_DT_BOOL) {
    BufPtr d = upload_bits(ctx, (const uint8_t*)a->buffers[1], off, length, staging);
    col.data = d->p;
    col.keep_data = d;
  } else {
    const int w = dtype_width(dtype);
    BufPtr d = dev_alloc(ctx, (size_t)w * length);
    if (length) SQ_CUDA(cudaMemcpyAsync(d->p, (const uint8_t*)a->buffers[1] + (size_t)off * w, (size_t)w * length, cudaMemcpyHostToDevice, ctx.stream));
    col.data = d->p;
    col.keep_data = d;
  }
  return col;
}

}  // namespace

DBatch import_batch_host(Ctx& ctx, ArrowArray* array, const ArrowSchema* schema) {
  if (!array || !array->release) fail(SQLRS_ERR_INVALID_ARG, "input ArrowArray is NULL or already released");
  auto moved = std::make_shared<MovedArray>();
  moved->a = *array;
  array->release = nullptr;  // moved
  DBatch b;
  b.fields = import_fields(schema);
  if (moved->a.n_children != (int64_t)b.fields.size()) fail(SQLRS_ERR_INVALID_ARG, "array/schema children mismatch");
  b.n = moved->a.length;
  std::vector<std::shared_ptr<std::vector<uint8_t>>> staging;
  for (int64_t c = 0; c < moved->a.n_children; c++)
    b.cols.push_back(import_column_host(ctx, moved->a.children[c], b.fields[c].dtype, moved->a.offset, moved->a.length, staging));
  // the producer's buffers (and our staging copies) must stay alive until the H2D copies are done
  ctx.defer([moved, staging]() {});
  return b;
}

// ------------------------------------------------------------------ import (device)
DBatch import_batch_device(Ctx& ctx, ArrowDeviceArray* darray, const ArrowSchema* schema) {
  if (!darray || !darray->array.release) fail(SQLRS_ERR_INVALID_ARG, "input ArrowDeviceArray is NULL or already released");
  if (darray->device_type != ARROW_DEVICE_CUDA) fail(SQLRS_ERR_INVALID_ARG, "plan_push_table_device expects ARROW_DEVICE_CUDA buffers");
  if (darray->device_id != ctx.device) fail(SQLRS_ERR_INVALID_ARG, "device array lives on another GPU than the plan");
  if (darray->sync_event) SQ_CUDA(cudaStreamWaitEvent(ctx.stream, *(cudaEvent_t*)darray->sync_event, 0));
  auto moved = std::make_shared<MovedArray>();
  moved->a = darray->array;
  darray->array.release = nullptr;
  DBatch b;
  b.fields = import_fields(schema);
  const ArrowArray& a = moved->a;
  if (a.n_children != (int64_t)b.fields.size()) fail(SQLRS_ERR_INVALID_ARG, "array/schema children mismatch");
  b.n = a.length;
  for (int64_t c = 0; c < a.n_children; c++) {
    const ArrowArray* ch = a.children[c];
    int dt = b.fields[c].dtype;
    DCol col;
    col.dtype = dt;
    col.n = a.length;
    if (dt == SQLRS_DT_NULL) {
      col.null_count = a.length;
      b.cols.push_back(col);
      continue;
    }
    if (dt == SQLRS_DT_UTF8) fail(SQLRS_ERR_UNSUPPORTED, "device-resident Utf8 columns are not supported (push host batches: strings are dictionary-encoded on ingest)");
    if (ch->length < a.offset + a.length) fail(SQLRS_ERR_INVALID_ARG, "child array shorter than batch");
    const int64_t off = ch->offset + a.offset;
    if (ch->buffers[0] && ch->null_count != 0) {
      if (off & 31) fail(SQLRS_ERR_UNSUPPORTED, "device-resident validity bitmaps need an offset that is a multiple of 32");
      col.valid = (const uint32_t*)ch->buffers[0] + (off >> 5);
      col.keep_valid = moved;
      col.null_count = -1;
    }
    if (dt == SQLRS_DT_BOOL) {
      if (off & 31) fail(SQLRS_ERR_UNSUPPORTED, "device-resident Boolean columns need an offset that is a multiple of 32");
      col.data = (const uint32_t*)ch->buffers[1] + (off >> 5);
    } else {
      col.data = (const uint8_t*)ch->buffers[1] + (size_t)off * dtype_width(dt);
    }
    col.keep_data = moved;
    b.cols.push_back(col);
  }
  return b;
}

// ------------------------------------------------------------------ export
namespace {

struct ExportPriv {
  std::vector<void*> owned;
  std::vector<const void*> buffers;
  std::vector<ArrowArray*> children;
  std::vector<ArrowArray> child_store;
};
struct SchemaPriv {
  std::string name;
  std::vector<ArrowSchema*> children;
  std::vector<ArrowSchema> child_store;
};

void release_array(ArrowArray* a) {
  if (!a || !a->release) return;
  for (int64_t c = 0; c < a->n_children; c++)
    if (a->children[c] && a->children[c]->release) a->children[c]->release(a->children[c]);
  auto* p = (ExportPriv*)a->private_data;
  for (void* m : p->owned)
    if (!pinned_cache().put(m)) std::free(m);
  delete p;
  a->release = nullptr;
}
void release_schema(ArrowSchema* s) {
  if (!s || !s->release) return;
  for (int64_t c = 0; c < s->n_children; c++)
    if (s->children[c] && s->children[c]->release) s->children[c]->release(s->children[c]);
  delete (SchemaPriv*)s->private_data;
  s->release = nullptr;
}
void* xmalloc(size_t bytes) {
  void* p = std::malloc(bytes ? bytes : 1);
  if (!p) fail(SQLRS_ERR_INTERNAL, "out of host memory");
  return p;
}
// destination of a D2H copy: pinned (recycled through the release callback) when it is large enough to matter
void* result_alloc(size_t bytes) {
  if (bytes >= (64u << 10)) {
    if (void* p = pinned_cache().get(bytes)) return p;
  }
  return xmalloc(bytes);
}
void export_field(const Field& f, ArrowSchema* out) {
  auto* p = new SchemaPriv();
  p->name = f.name;
  std::memset(out, 0, sizeof(*out));
  out->format = format_of_dtype(f.dtype);
  out->name = p->name.c_str();
  out->flags = f.nullable ? ARROW_FLAG_NULLABLE : 0;
  out->private_data = p;
  out->release = release_schema;
}

}  // namespace

void export_schema(const std::vector<Field>& fields, ArrowSchema* out) {
  auto* p = new SchemaPriv();
  p->child_store.resize(fields.size());
  for (size_t c = 0; c < fields.size(); c++) {
    export_field(fields[c], &p->child_store[c]);
    p->children.push_back(&p->child_store[c]);
  }
  std::memset(out, 0, sizeof(*out));
  out->format = "+s";
  out->name = "";
  out->n_children = (int64_t)fields.size();
  out->children = p->children.data();
  out->private_data = p;
  out->release = release_schema;
}

void export_batch_host(Ctx& ctx, const DBatch& b, ArrowArray* out, ArrowSchema* out_schema) {
  // null counts first (one device reduction per bitmap), then all D2H copies, then one sync
  std::vector<DCol> cols = b.cols;
  std::vector<BufPtr> counters(cols.size());
  std::vector<unsigned long long> set_bits(cols.size(), 0);
  for (size_t c = 0; c < cols.size(); c++) {
    if (cols[c].valid && cols[c].null_count < 0 && cols[c].n > 0) {
      counters[c] = dev_alloc_zero(ctx, 8);
      launch_count_bits(cols[c].valid, cols[c].n, (unsigned long long*)counters[c]->p, ctx.stream);
      SQ_CUDA(cudaMemcpyAsync(&set_bits[c], counters[c]->p, 8, cudaMemcpyDeviceToHost, ctx.stream));
    }
  }
  auto* p = new ExportPriv();
  std::unique_ptr<ExportPriv> hold(p);
  p->child_store.resize(cols.size());
  std::vector<ExportPriv*> child_priv(cols.size());
  std::vector<uint8_t*> validity_host(cols.size(), nullptr);
  for (size_t c = 0; c < cols.size(); c++) {
    const DCol& col = cols[c];
    auto* cp = new ExportPriv();
    child_priv[c] = cp;
    ArrowArray* a = &p->child_store[c];
    std::memset(a, 0, sizeof(*a));
    a->length = col.n;
    a->private_data = cp;
    a->release = release_array;
    p->children.push_back(a);
    if (col.dtype == SQLRS_DT_NULL) {
      a->null_count = col.n;
      continue;
    }
    if (col.valid && col.n > 0) {
      size_t nb = (size_t)bitmap_words(col.n) * 4;
      validity_host[c] = (uint8_t*)result_alloc(nb);
      cp->owned.push_back(validity_host[c]);
      SQ_CUDA(cudaMemcpyAsync(validity_host[c], col.valid, nb, cudaMemcpyDeviceToHost, ctx.stream));
    }
    cp->buffers.push_back(nullptr);
    size_t vb = col_value_bytes(col.dtype, col.n);
    void* v = result_alloc(vb);
    cp->owned.push_back(v);
    if (vb) SQ_CUDA(cudaMemcpyAsync(v, col.data, vb, cudaMemcpyDeviceToHost, ctx.stream));
    cp->buffers.push_back(v);
  }
  ctx.sync();
  for (size_t c = 0; c < cols.size(); c++) {
    const DCol& col = cols[c];
    if (col.dtype == SQLRS_DT_NULL) continue;
    ArrowArray* a = &p->child_store[c];
    ExportPriv* cp = child_priv[c];
    int64_t nulls = 0;
    if (col.valid && col.n > 0) nulls = col.null_count >= 0 ? col.null_count : col.n - (int64_t)set_bits[c];
    a->null_count = nulls;
    if (nulls > 0) cp->buffers[0] = validity_host[c];
    a->n_buffers = 2;
    if (col.dtype == SQLRS_DT_UTF8) {  // decode the pool ids: Arrow Utf8 = [validity, int32 offsets, bytes]
      const int64_t* ids = (const int64_t*)cp->buffers[1];
      int32_t* offsets = (int32_t*)xmalloc(sizeof(int32_t) * (size_t)(col.n + 1));
      std::string chars;
      StringPool& pool = StringPool::instance();
      offsets[0] = 0;
      for (int64_t r = 0; r < col.n; r++) {
        const bool is_null = nulls > 0 && !((validity_host[c][r >> 3] >> (r & 7)) & 1);
        if (!is_null) chars += pool.get(ids[r]);
        if (chars.size() > 0x7fffffffu) fail(SQLRS_ERR_UNSUPPORTED, "Utf8 result column exceeds 2 GiB");
        offsets[r + 1] = (int32_t)chars.size();
      }
      char* bytes = (char*)xmalloc(chars.size());
      if (!chars.empty()) std::memcpy(bytes, chars.data(), chars.size());
      cp->owned.push_back(offsets);
      cp->owned.push_back(bytes);
      cp->buffers[1] = offsets;
      cp->buffers.push_back(bytes);
      a->n_buffers = 3;
    }
    a->buffers = cp->buffers.data();
  }
  p->buffers.push_back(nullptr);
  std::memset(out, 0, sizeof(*out));
  out->length = b.n;
  out->n_buffers = 1;
  out->buffers = p->buffers.data();
  out->n_children = (int64_t)cols.size();
  out->children = p->children.data();
  out->private_data = hold.release();
  out->release = release_array;
  if (out_schema) export_schema(b.fields, out_schema);
}

void export_host_columns(const std::vector<Field>& fields, const std::vector<HostCol>& cols, int64_t n, ArrowArray* out,
                         ArrowSchema* out_schema) {
  auto* p = new ExportPriv();
  std::unique_ptr<ExportPriv> hold(p);
  p->child_store.resize(cols.size());
  for (size_t c = 0; c < cols.size(); c++) {
    const HostCol& col = cols[c];
    auto* cp = new ExportPriv();
    ArrowArray* a = &p->child_store[c];
    std::memset(a, 0, sizeof(*a));
    a->length = n;
    a->private_data = cp;
    a->release = release_array;
    p->children.push_back(a);
    if (col.dtype == SQLRS_DT_NULL) {
      a->null_count = n;
      continue;
    }
    int64_t nulls = 0;
    uint8_t* validity = nullptr;
    if (!col.valid.empty()) {
      for (uint8_t v : col.valid) nulls += v == 0;
      if (nulls) {
        size_t nb = (size_t)((n + 7) / 8);
        validity = (uint8_t*)xmalloc(nb);
        std::memset(validity, 0, nb);
        for (int64_t r = 0; r < n; r++)
          if (col.valid[r]) validity[r >> 3] |= (uint8_t)(1u << (r & 7));
        cp->owned.push_back(validity);
      }
    }
    cp->buffers.push_back(validity);
    void* v = nullptr;
    switch (col.dtype) {
      case SQLRS_DT_BOOL: {
        size_t nb = (size_t)((n + 7) / 8);
        uint8_t* b = (uint8_t*)xmalloc(nb);
        std::memset(b, 0, nb);
        for (int64_t r = 0; r < n; r++)
          if (col.i[r]) b[r >> 3] |= (uint8_t)(1u << (r & 7));
        v = b;
        break;
      }
      case SQLRS_DT_INT32: {
        int32_t* b = (int32_t*)xmalloc(sizeof(int32_t) * n);
        for (int64_t r = 0; r < n; r++) b[r] = (int32_t)col.i[r];
        v = b;
        break;
      }
      case SQLRS_DT_INT64: {
        int64_t* b = (int64_t*)xmalloc(sizeof(int64_t) * n);
        if (n) std::memcpy(b, col.i.data(), sizeof(int64_t) * n);
        v = b;
        break;
      }
      case SQLRS_DT_FLOAT64: {
        double* b = (double*)xmalloc(sizeof(double) * n);
        if (n) std::memcpy(b, col.f.data(), sizeof(double) * n);
        v = b;
        break;
      }
      default: fail(SQLRS_ERR_UNSUPPORTED, "unsupported result column type");
    }
    cp->owned.push_back(v);
    cp->buffers.push_back(v);
    a->null_count = nulls;
    a->n_buffers = 2;
    a->buffers = cp->buffers.data();
  }
  p->buffers.push_back(nullptr);
  std::memset(out, 0, sizeof(*out));
  out->length = n;
  out->n_buffers = 1;
  out->buffers = p->buffers.data();
  out->n_children = (int64_t)cols.size();
  out->children = p->children.data();
  out->private_data = hold.release();
  out->release = release_array;
  if (out_schema) export_schema(fields, out_schema);
}

}  // namespace sq
