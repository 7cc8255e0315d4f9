// ORACLE — TEST INFRASTRUCTURE ONLY.
// Arrow C Data Interface <-> oracle::Batch (host memory only).  Import copies the
// data out of the producer's buffers (honouring offset / null_count / absent validity,
// SURVEY.md §8b "Data ownership"); export hands out malloc'ed buffers with our release.
#pragma once
#include <cstdlib>

#include "columns.hpp"

namespace oracle {

inline int dtype_from_format(const char* fmt) {
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
inline const char* format_of_dtype(int dt) {
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

inline bool bit_get(const uint8_t* bits, int64_t i) { return (bits[i >> 3] >> (i & 7)) & 1; }

inline std::vector<Field> import_fields(const ArrowSchema* schema) {
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

inline ColPtr import_column(const ArrowArray* a, int dtype, int64_t parent_offset, int64_t length) {
  auto col = std::make_shared<Column>();
  col->alloc(dtype, length);
  if (dtype == SQLRS_DT_NULL) return col;
  int64_t off = a->offset + parent_offset;
  if (a->length < parent_offset + length) fail(SQLRS_ERR_INVALID_ARG, "child array shorter than batch");
  const uint8_t* validity = a->n_buffers > 0 ? (const uint8_t*)a->buffers[0] : nullptr;
  if (validity && a->null_count != 0) {
    col->valid.assign(length, 1);
    for (int64_t r = 0; r < length; r++) col->valid[r] = bit_get(validity, off + r);
  }
  switch (dtype) {
    case SQLRS_DT_BOOL: {
      const uint8_t* v = (const uint8_t*)a->buffers[1];
      for (int64_t r = 0; r < length; r++) col->i[r] = bit_get(v, off + r);
      break;
    }
    case SQLRS_DT_INT32: {
      const int32_t* v = (const int32_t*)a->buffers[1];
      for (int64_t r = 0; r < length; r++) col->i[r] = v[off + r];
      break;
    }
    case SQLRS_DT_INT64: {
      const int64_t* v = (const int64_t*)a->buffers[1];
      if (length) std::memcpy(col->i.data(), v + off, sizeof(int64_t) * length);
      break;
    }
    case SQLRS_DT_FLOAT64: {
      const double* v = (const double*)a->buffers[1];
      if (length) std::memcpy(col->f.data(), v + off, sizeof(double) * length);
      break;
    }
    case SQLRS_DT_UTF8: {
      const int32_t* o = (const int32_t*)a->buffers[1];
      const char* d = (const char*)a->buffers[2];
      for (int64_t r = 0; r < length; r++)
        col->s[r].assign(d + o[off + r], (size_t)(o[off + r + 1] - o[off + r]));
      break;
    }
  }
  col->normalize();
  return col;
}

// Import a struct array as a RecordBatch.  Does NOT release `array` (the ABI layer does).
inline Batch import_batch(const ArrowArray* array, const ArrowSchema* schema) {
  Batch b;
  b.fields = import_fields(schema);
  if (!array) fail(SQLRS_ERR_INVALID_ARG, "ArrowArray is NULL");
  if (array->n_children != (int64_t)b.fields.size())
    fail(SQLRS_ERR_INVALID_ARG, "array/schema children mismatch");
  b.n = array->length;
  for (int64_t c = 0; c < array->n_children; c++)
    b.cols.push_back(import_column(array->children[c], b.fields[c].dtype, array->offset, array->length));
  return b;
}

// ------------------------------------------------------------------ export
struct ExportPriv {
  std::vector<void*> owned;              // malloc'ed buffers
  std::vector<const void*> buffers;      // buffers[] storage
  std::vector<ArrowArray*> children;     // children[] storage
  std::vector<ArrowArray> child_store;
};
struct SchemaPriv {
  std::string name;
  std::vector<ArrowSchema*> children;
  std::vector<ArrowSchema> child_store;
};

inline void release_array(ArrowArray* a) {
  if (!a || !a->release) return;
  for (int64_t c = 0; c < a->n_children; c++)
    if (a->children[c] && a->children[c]->release) a->children[c]->release(a->children[c]);
  auto* p = (ExportPriv*)a->private_data;
  for (void* m : p->owned) std::free(m);
  delete p;
  a->release = nullptr;
}
inline void release_schema(ArrowSchema* s) {
  if (!s || !s->release) return;
  for (int64_t c = 0; c < s->n_children; c++)
    if (s->children[c] && s->children[c]->release) s->children[c]->release(s->children[c]);
  delete (SchemaPriv*)s->private_data;
  s->release = nullptr;
}

inline void* xmalloc(size_t bytes) {
  void* p = std::malloc(bytes ? bytes : 1);
  if (!p) fail(SQLRS_ERR_INTERNAL, "out of memory");
  return p;
}

inline void export_column(const Column& c, ArrowArray* out) {
  auto* p = new ExportPriv();
  std::memset(out, 0, sizeof(*out));
  out->length = c.n;
  out->offset = 0;
  out->null_count = c.null_count();
  out->private_data = p;
  out->release = release_array;
  if (c.dtype == SQLRS_DT_NULL) {
    out->n_buffers = 0;
    out->buffers = nullptr;
    return;
  }
  uint8_t* validity = nullptr;
  if (out->null_count > 0) {
    size_t nb = (size_t)((c.n + 7) / 8);
    validity = (uint8_t*)xmalloc(nb);
    std::memset(validity, 0, nb);
    for (int64_t r = 0; r < c.n; r++)
      if (c.valid[r]) validity[r >> 3] |= (uint8_t)(1u << (r & 7));
    p->owned.push_back(validity);
  }
  p->buffers.push_back(validity);
  switch (c.dtype) {
    case SQLRS_DT_BOOL: {
      size_t nb = (size_t)((c.n + 7) / 8);
      uint8_t* v = (uint8_t*)xmalloc(nb);
      std::memset(v, 0, nb);
      for (int64_t r = 0; r < c.n; r++)
        if (c.i[r]) v[r >> 3] |= (uint8_t)(1u << (r & 7));
      p->owned.push_back(v);
      p->buffers.push_back(v);
      break;
    }
    case SQLRS_DT_INT32: {
      int32_t* v = (int32_t*)xmalloc(sizeof(int32_t) * c.n);
      for (int64_t r = 0; r < c.n; r++) v[r] = (int32_t)c.i[r];
      p->owned.push_back(v);
      p->buffers.push_back(v);
      break;
    }
    case SQLRS_DT_INT64: {
      int64_t* v = (int64_t*)xmalloc(sizeof(int64_t) * c.n);
      if (c.n) std::memcpy(v, c.i.data(), sizeof(int64_t) * c.n);
      p->owned.push_back(v);
      p->buffers.push_back(v);
      break;
    }
    case SQLRS_DT_FLOAT64: {
      double* v = (double*)xmalloc(sizeof(double) * c.n);
      if (c.n) std::memcpy(v, c.f.data(), sizeof(double) * c.n);
      p->owned.push_back(v);
      p->buffers.push_back(v);
      break;
    }
    case SQLRS_DT_UTF8: {
      int32_t* o = (int32_t*)xmalloc(sizeof(int32_t) * (c.n + 1));
      size_t total = 0;
      for (int64_t r = 0; r < c.n; r++) total += c.s[r].size();
      char* d = (char*)xmalloc(total);
      size_t pos = 0;
      for (int64_t r = 0; r < c.n; r++) {
        o[r] = (int32_t)pos;
        std::memcpy(d + pos, c.s[r].data(), c.s[r].size());
        pos += c.s[r].size();
      }
      o[c.n] = (int32_t)pos;
      p->owned.push_back(o);
      p->owned.push_back(d);
      p->buffers.push_back(o);
      p->buffers.push_back(d);
      break;
    }
  }
  out->n_buffers = (int64_t)p->buffers.size();
  out->buffers = p->buffers.data();
}

inline void export_field(const Field& f, ArrowSchema* out) {
  auto* p = new SchemaPriv();
  p->name = f.name;
  std::memset(out, 0, sizeof(*out));
  out->format = format_of_dtype(f.dtype);
  out->name = p->name.c_str();
  out->flags = f.nullable ? ARROW_FLAG_NULLABLE : 0;
  out->private_data = p;
  out->release = release_schema;
}

inline void export_schema(const std::vector<Field>& fields, ArrowSchema* out) {
  auto* p = new SchemaPriv();
  p->child_store.resize(fields.size());
  for (size_t c = 0; c < fields.size(); c++) {
    export_field(fields[c], &p->child_store[c]);
    p->children.push_back(&p->child_store[c]);
  }
  std::memset(out, 0, sizeof(*out));
  out->format = "+s";
  out->name = "";
  out->flags = 0;
  out->n_children = (int64_t)fields.size();
  out->children = p->children.data();
  out->private_data = p;
  out->release = release_schema;
}

inline void export_batch(const Batch& b, ArrowArray* out, ArrowSchema* out_schema) {
  auto* p = new ExportPriv();
  p->child_store.resize(b.cols.size());
  for (size_t c = 0; c < b.cols.size(); c++) {
    export_column(*b.cols[c], &p->child_store[c]);
    p->children.push_back(&p->child_store[c]);
  }
  p->buffers.push_back(nullptr);  // struct validity
  std::memset(out, 0, sizeof(*out));
  out->length = b.n;
  out->null_count = 0;
  out->offset = 0;
  out->n_buffers = 1;
  out->buffers = p->buffers.data();
  out->n_children = (int64_t)b.cols.size();
  out->children = p->children.data();
  out->private_data = p;
  out->release = release_array;
  if (out_schema) export_schema(b.fields, out_schema);
}

}  // namespace oracle
