// sqlrs_b200 — expression evaluation into columns, Filter, take/concat helpers.
#include "kernels_aot.hpp"
#include "ops.hpp"

namespace sq {

std::vector<ColInfo> col_infos(const DBatch& b) {
  std::vector<ColInfo> out;
  for (size_t k = 0; k < b.cols.size(); k++) {
    const DCol& c = b.cols[k];
    out.push_back(ColInfo{c.dtype, c.valid != nullptr, k < b.fields.size() ? b.fields[k].nullable : true});
  }
  return out;
}

std::string gen_input_decls(const std::vector<ColInfo>& cols) {
  std::ostringstream s;
  size_t n = cols.size() ? cols.size() : 1;
  s << "#define SQ_NCOLS " << cols.size() << "\n";
  s << "struct SqIn { const void* col[" << n << "]; const u32* val[" << n << "]; };\n";
  s << "#define SQ_LD_I64(c, r) sq_ld_i64(in.col[c], r)\n";
  s << "#define SQ_LD_I32(c, r) sq_ld_i32(in.col[c], r)\n";
  s << "#define SQ_LD_F64(c, r) sq_ld_f64(in.col[c], r)\n";
  s << "#define SQ_LD_BOOL(c, r) sq_ld_bit(in.col[c], r)\n";
  s << "#define SQ_VALID(c, r) sq_ld_bit(in.val[c], r)\n";
  return s.str();
}

namespace {
struct SqInHost {
  std::vector<const void*> blob;  // col[N] then val[N]
  explicit SqInHost(const DBatch& b) {
    size_t n = b.cols.size() ? b.cols.size() : 1;
    blob.assign(2 * n, nullptr);
    for (size_t c = 0; c < b.cols.size(); c++) {
      blob[c] = b.cols[c].data;
      blob[n + c] = b.cols[c].valid;
    }
  }
  void* ptr() { return blob.data(); }
};
}  // namespace

void check_error_flag(Ctx& ctx, const BufPtr& err, const char* what) {
  uint32_t flag = 0;
  SQ_CUDA(cudaMemcpyAsync(&flag, err->p, 4, cudaMemcpyDeviceToHost, ctx.stream));
  SQ_CUDA(cudaStreamSynchronize(ctx.stream));
  if (flag & 1u) fail(SQLRS_ERR_ARROW, std::string(arithmetic_error_text()) + (what ? std::string(" (") + what + ")" : ""));
}

// ------------------------------------------------------------------ EvalProgram
std::string EvalProgram::source_for(const std::vector<ColInfo>& cols, std::vector<int>* out_dtypes, std::vector<bool>* out_nullable,
                                    std::vector<int>* expr_dtypes) {
  RowProgram prog(cols);
  std::vector<Val> vals;
  for (const ExprCopy& e : req_.exprs) vals.push_back(prog.compile(e, 0));
  expr_dtypes->clear();
  for (const Val& v : vals) expr_dtypes->push_back(v.dtype);
  std::vector<Val> keys;
  for (size_t k = 0; k < vals.size(); k++)
    if (k < req_.is_key.size() && req_.is_key[k]) keys.push_back(vals[k]);

  struct OutGen {
    std::string ctype, value, valid;
    int dtype;
    bool nullable, bits;
  };
  std::vector<OutGen> outs;
  for (const EvalRequest::Out& o : req_.outs) {
    OutGen g;
    g.bits = false;
    g.nullable = false;
    switch (o.kind) {
      case OUT_VALUE: {
        const Val& v = vals.at(o.expr);
        g.dtype = v.dtype;
        g.ctype = ctype_of(v.dtype);
        g.value = "v" + std::to_string(v.id);
        g.valid = "n" + std::to_string(v.id);
        g.nullable = v.maybe_null;
        g.bits = v.dtype == SQLRS_DT_BOOL;
        break;
      }
      case OUT_KEEP: {
        const Val& v = vals.at(o.expr);
        if (v.dtype != SQLRS_DT_BOOL) fail(SQLRS_ERR_INTERNAL, "filter executor expected evaluate boolean array");
        g.dtype = SQLRS_DT_BOOL;
        g.ctype = "bool";
        g.value = "(n" + std::to_string(v.id) + " && v" + std::to_string(v.id) + ")";
        g.valid = "true";
        g.bits = true;
        break;
      }
      case OUT_HASH: {
        int id = prog.emit_row_hash(keys);
        g.dtype = SQLRS_DT_INT64;
        g.ctype = "u64";
        g.value = "v" + std::to_string(id);
        g.valid = "true";
        break;
      }
      case OUT_MIXHASH: {
        std::vector<int> raw;
        for (const Val& k : keys) raw.push_back(prog.emit_raw_bits(k));
        int id = prog.emit_mix_hash(raw, keys);
        g.dtype = SQLRS_DT_INT64;
        g.ctype = "u64";
        g.value = "v" + std::to_string(id);
        g.valid = "true";
        break;
      }
      case OUT_RAWBITS: {
        int id = prog.emit_raw_bits(vals.at(o.expr));
        g.dtype = SQLRS_DT_INT64;
        g.ctype = "u64";
        g.value = "v" + std::to_string(id);
        g.valid = "true";
        break;
      }
      case OUT_NULLMASK: {
        std::string m = "0u";
        for (size_t k = 0; k < keys.size(); k++) m += " | (n" + std::to_string(keys[k].id) + " ? 0u : " + std::to_string(1u << k) + "u)";
        g.dtype = SQLRS_DT_INT32;
        g.ctype = "u32";
        g.value = "(" + m + ")";
        g.valid = "true";
        break;
      }
      default: fail(SQLRS_ERR_INVALID_ARG, "bad eval output kind");
    }
    outs.push_back(g);
  }
  out_dtypes->clear();
  out_nullable->clear();
  std::ostringstream s;
  s << gen_input_decls(cols);
  size_t m = outs.size() ? outs.size() : 1;
  s << "#define SQ_NOUT " << outs.size() << "\n";
  s << "struct SqOut { void* col[" << m << "]; u32* val[" << m << "]; };\n";
  s << "struct SqRow {\n";
  for (size_t j = 0; j < outs.size(); j++) s << "  " << outs[j].ctype << " v" << j << "; bool n" << j << ";\n";
  s << "};\n";
  s << "__device__ __forceinline__ void sq_row(const SqIn& in, i64 r, SqRow& o, bool& e0, bool& e1) {\n";
  s << prog.body_str();
  for (size_t j = 0; j < outs.size(); j++) s << "  o.v" << j << " = " << outs[j].value << "; o.n" << j << " = " << outs[j].valid << ";\n";
  s << "}\n";
  s << "__device__ __forceinline__ void sq_store(const SqOut& out, i64 r, bool inb, int lane, const SqRow& o) {\n";
  for (size_t j = 0; j < outs.size(); j++) {
    const OutGen& g = outs[j];
    out_dtypes->push_back(g.dtype);
    out_nullable->push_back(g.nullable);
    if (g.dtype == SQLRS_DT_NULL) continue;
    if (g.bits) {
      s << "  { const u32 w = __ballot_sync(SQ_FULL, inb && o.v" << j << "); if (lane == 0) ((u32*)out.col[" << j << "])[r >> 5] = w; }\n";
    } else {
      s << "  if (inb) ((" << g.ctype << "*)out.col[" << j << "])[r] = o.v" << j << ";\n";
    }
    if (g.nullable)
      s << "  { const u32 w = __ballot_sync(SQ_FULL, inb && o.n" << j << "); if (lane == 0) out.val[" << j << "][r >> 5] = w; }\n";
  }
  s << "}\n";
  return s.str();
}

EvalResult EvalProgram::run(Ctx& ctx, const DBatch& batch, const char* what) {
  Trace tr("eval.run", ctx.stream);
  std::vector<ColInfo> cols = col_infos(batch);
  std::string sig = RowProgram(cols).signature();
  auto it = cache_.find(sig);
  if (it == cache_.end()) {
    Compiled c;
    std::string src = source_for(cols, &c.out_dtypes, &c.out_nullable, &c.expr_dtypes);
    c.kernel = jit_get("eval", src, "sq_eval_kernel");
    it = cache_.emplace(sig, std::move(c)).first;
  }
  Compiled& c = it->second;
  EvalResult res;
  res.expr_dtypes = c.expr_dtypes;
  const int64_t n = batch.n;
  size_t m = c.out_dtypes.size() ? c.out_dtypes.size() : 1;
  std::vector<void*> out_blob(2 * m, nullptr);
  {
    Trace tr_alloc("  eval.alloc", ctx.stream);
    for (size_t j = 0; j < c.out_dtypes.size(); j++) {
      DCol col = make_col(ctx, c.out_dtypes[j], n, c.out_nullable[j]);
      out_blob[j] = col_data(col);
      out_blob[m + j] = col_valid(col);
      res.cols.push_back(col);
    }
  }
  if (n == 0) return res;
  BufPtr err = dev_alloc_zero(ctx, 4);
  SqInHost in(batch);
  int64_t n_arg = n;
  void* errp = err->p;
  void* args[] = {in.ptr(), out_blob.data(), &n_arg, &errp};
  const int sms = device_sm_count(ctx.device);
  int64_t want = div_up(n, 256 * 4);
  unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)sms * 16);
  jit_launch(c.kernel, grid, 256, 0, ctx.stream, args);
  check_error_flag(ctx, err, what);
  return res;
}

// ------------------------------------------------------------------ take / compaction / concat
DCol gather_col_u32(Ctx& ctx, const DCol& src, const uint32_t* idx, int64_t m) {
  DCol out = make_col(ctx, src.dtype, m, src.valid != nullptr);
  if (src.dtype == SQLRS_DT_NULL || m == 0) return out;
  launch_gather_u32idx(dtype_width(src.dtype), src.data, src.valid, idx, m, col_data(out), col_valid(out), ctx.stream);
  return out;
}
DCol gather_col_i64(Ctx& ctx, const DCol& src, const int64_t* idx, int64_t m, bool idx_may_be_null) {
  DCol out = make_col(ctx, src.dtype, m, src.valid != nullptr || idx_may_be_null);
  if (src.dtype == SQLRS_DT_NULL || m == 0) return out;
  launch_gather_i64idx(dtype_width(src.dtype), src.data, src.valid, idx, m, col_data(out), col_valid(out), ctx.stream);
  return out;
}

BufPtr compact_indices(Ctx& ctx, const uint32_t* keep, int64_t n, int64_t* out_count) {
  *out_count = 0;
  if (n == 0) return dev_alloc(ctx, 4);
  size_t chunks = compact_num_chunks(n);
  BufPtr counts = dev_alloc(ctx, chunks * 4);
  BufPtr offsets = dev_alloc(ctx, chunks * 8 + 8);
  unsigned long long* total_d = (unsigned long long*)offsets->p + chunks;
  launch_compact_count(keep, n, (uint32_t*)counts->p, ctx.stream);
  launch_scan_u32((const uint32_t*)counts->p, (int64_t)chunks, (unsigned long long*)offsets->p, total_d, ctx.stream);
  unsigned long long total = 0;
  SQ_CUDA(cudaMemcpyAsync(&total, total_d, 8, cudaMemcpyDeviceToHost, ctx.stream));
  SQ_CUDA(cudaStreamSynchronize(ctx.stream));
  *out_count = (int64_t)total;
  BufPtr idx = dev_alloc(ctx, (size_t)total * 4);
  if (total) launch_compact_write(keep, n, (const unsigned long long*)offsets->p, (uint32_t*)idx->p, ctx.stream);
  return idx;
}

DCol concat_cols(Ctx& ctx, const std::vector<DCol>& parts, int dtype) {
  int64_t total = 0;
  bool any_valid = false;
  for (const DCol& p : parts) {
    if (p.dtype != dtype) fail(SQLRS_ERR_ARROW, "concat_batches: column type mismatch");
    total += p.n;
    any_valid |= p.valid != nullptr;
  }
  if (parts.size() == 1) return parts[0];
  DCol out = make_col(ctx, dtype, total, any_valid);
  if (dtype == SQLRS_DT_NULL) return out;
  int64_t off = 0;
  for (const DCol& p : parts) {
    if (p.n == 0) continue;
    if (dtype == SQLRS_DT_BOOL) {
      launch_bitmap_append((uint32_t*)col_data(out), off, (const uint32_t*)p.data, p.n, ctx.stream);
    } else {
      const int w = dtype_width(dtype);
      SQ_CUDA(cudaMemcpyAsync((uint8_t*)col_data(out) + (size_t)off * w, p.data, (size_t)p.n * w, cudaMemcpyDeviceToDevice, ctx.stream));
    }
    if (any_valid) launch_bitmap_append(col_valid(out), off, p.valid, p.n, ctx.stream);
    off += p.n;
  }
  return out;
}

// ------------------------------------------------------------------ Filter
static EvalRequest keep_request(const ExprCopy& predicate) {
  EvalRequest r;
  r.exprs.push_back(predicate);
  r.is_key.push_back(false);
  r.outs.push_back({OUT_KEEP, 0});
  return r;
}

// filter.rs:16-25 — mask = eval; keep rows whose mask is valid and true; row order preserved
DBatch filter_batch(Ctx& ctx, EvalProgram& prog, const DBatch& in, const std::vector<bool>* needed) {
  EvalResult mask = prog.run(ctx, in, "filter predicate");
  DBatch out;
  out.fields = in.fields;
  int64_t kept = 0;
  BufPtr idx = compact_indices(ctx, (const uint32_t*)mask.cols[0].data, in.n, &kept);
  out.n = kept;
  for (size_t k = 0; k < in.cols.size(); k++) {
    if (needed && !needed->empty() && !(k < needed->size() && (*needed)[k])) {
      DCol ph;  // nobody above reads this column
      ph.dtype = SQLRS_DT_NULL;
      ph.n = kept;
      ph.null_count = kept;
      out.cols.push_back(ph);
      continue;
    }
    out.cols.push_back(gather_col_u32(ctx, in.cols[k], (const uint32_t*)idx->p, kept));
  }
  // the index list must outlive the enqueued gathers
  ctx.defer([idx]() {});
  return out;
}

FilterOp::FilterOp(const ExprCopy& predicate, const Options& opt) : ctx_(opt), prog_(keep_request(predicate)) {
  if (predicate.empty()) fail(SQLRS_ERR_INVALID_ARG, "filter needs a predicate");
}
DBatch FilterOp::execute(const DBatch& in) { return filter_batch(ctx_, prog_, in); }

}  // namespace sq
