// sqlrs_b200 — CSV ingest onto the device: host side (see csv.hpp, kernels_csv.cu).
// Reference: CsvTable / CsvTransaction (src/storage/csv.rs:99-235) over arrow-csv 28's Reader [ext]: header row, ',' delimiter,
// schema inferred from the first 10 records (Boolean / Int64 / Float64 / Utf8 by arrow-csv's regular expressions, an
// Int64+Float64 mix is Float64, anything else Utf8; every field nullable), 1024-row batches, bounds over the whole file,
// projection.  An empty field is NULL in a numeric / Boolean column and the empty string in a Utf8 column.
#include "csv.hpp"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "kernels_aot.hpp"

namespace sq {
namespace {

// one record of `buf` starting at `pos` split into fields (quotes honoured, "" unescaped); returns the position after it
size_t split_record(const std::string& buf, size_t pos, char delim, std::vector<std::string>* fields, std::vector<bool>* quoted) {
  fields->clear();
  quoted->clear();
  std::string cur;
  bool in_q = false, was_q = false;
  size_t i = pos;
  for (; i < buf.size(); i++) {
    const char c = buf[i];
    if (in_q) {
      if (c == '"') {
        if (i + 1 < buf.size() && buf[i + 1] == '"') {
          cur += '"';
          i++;
        } else {
          in_q = false;
        }
      } else {
        cur += c;
      }
    } else if (c == '"' && cur.empty() && !was_q) {
      in_q = true;
      was_q = true;
    } else if (c == delim) {
      fields->push_back(cur);
      quoted->push_back(was_q);
      cur.clear();
      was_q = false;
    } else if (c == '\n') {
      break;
    } else if (c != '\r' || (i + 1 < buf.size() && buf[i + 1] != '\n')) {
      cur += c;
    }
  }
  fields->push_back(cur);
  quoted->push_back(was_q);
  return i < buf.size() ? i + 1 : i;
}

bool all_digits(const std::string& s, size_t from) {
  if (from >= s.size()) return false;
  for (size_t i = from; i < s.size(); i++)
    if (!std::isdigit((unsigned char)s[i])) return false;
  return true;
}
// arrow-csv 28 infer_field_schema [ext]: BOOLEAN (?i)^(true)$|^(false)$, DECIMAL ^-?(\d+\.\d+)$, INTEGER ^-?(\d+)$; quoted -> Utf8
int infer_cell(const std::string& s, bool quoted) {
  if (quoted) return SQLRS_DT_UTF8;
  std::string low;
  for (char c : s) low += (char)std::tolower((unsigned char)c);
  if (low == "true" || low == "false") return SQLRS_DT_BOOL;
  const size_t from = !s.empty() && s[0] == '-' ? 1 : 0;
  if (all_digits(s, from)) return SQLRS_DT_INT64;
  const size_t dot = s.find('.');
  if (dot != std::string::npos && all_digits(s.substr(0, dot), from) && all_digits(s.substr(dot + 1), 0)) return SQLRS_DT_FLOAT64;
  return SQLRS_DT_UTF8;  // (dates infer as Date32 / Date64 in arrow-csv; the v1 engine has no such type: kept as text)
}

}  // namespace

std::vector<DBatch> read_csv_device(Ctx& ctx, const std::string& path, const CsvOptions& opt) {
  std::ifstream f(path, std::ios::binary);
  if (!f) fail(SQLRS_ERR_STORAGE, "io error: cannot open " + path);
  std::string buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (!buf.empty() && buf.back() != '\n') buf += '\n';
  const int64_t n = (int64_t)buf.size();
  if (n >= (1LL << 40)) fail(SQLRS_ERR_UNSUPPORTED, "csv: file too large");

  // ---- schema: names from the header, types from the first records (host; a few hundred bytes)
  std::vector<std::string> names, cells;
  std::vector<bool> quoted;
  size_t pos = 0;
  if (opt.has_header && n > 0) {
    pos = split_record(buf, 0, opt.delimiter, &names, &quoted);
  }
  std::vector<int> seen_mask;  // bit per dtype seen in the sampled records
  {
    size_t p = pos;
    for (int64_t r = 0; r < opt.infer_max_records && p < buf.size(); r++) {
      p = split_record(buf, p, opt.delimiter, &cells, &quoted);
      if (seen_mask.size() < cells.size()) seen_mask.resize(cells.size(), 0);
      for (size_t c = 0; c < cells.size(); c++)
        if (!cells[c].empty() || quoted[c]) seen_mask[c] |= 1 << infer_cell(cells[c], quoted[c]);
    }
  }
  const size_t n_fields = std::max(names.size(), seen_mask.size());
  seen_mask.resize(n_fields, 0);
  std::vector<Field> fields(n_fields);
  for (size_t c = 0; c < n_fields; c++) {
    fields[c].name = c < names.size() && opt.has_header ? names[c] : "column_" + std::to_string(c + 1);
    fields[c].nullable = true;
    const int m = seen_mask[c];
    if (m == (1 << SQLRS_DT_BOOL)) fields[c].dtype = SQLRS_DT_BOOL;
    else if (m == (1 << SQLRS_DT_INT64)) fields[c].dtype = SQLRS_DT_INT64;
    else if (m == (1 << SQLRS_DT_FLOAT64) || m == ((1 << SQLRS_DT_INT64) | (1 << SQLRS_DT_FLOAT64))) fields[c].dtype = SQLRS_DT_FLOAT64;
    else fields[c].dtype = SQLRS_DT_UTF8;
  }
  std::vector<int> proj = opt.projection;
  if (proj.empty())
    for (size_t c = 0; c < n_fields; c++) proj.push_back((int)c);
  if (proj.size() > (size_t)kCsvMaxColumns) fail(SQLRS_ERR_UNSUPPORTED, "csv: more than 64 projected columns");
  for (int c : proj)
    if (c < 0 || c >= (int)n_fields) fail(SQLRS_ERR_ARROW, "csv: projection index out of bounds");

  // ---- index the records on the device
  std::vector<DBatch> out;
  DBatch whole;
  for (int c : proj) whole.fields.push_back(fields[(size_t)c]);
  int64_t n_rows = 0;
  BufPtr d_buf, line_end;
  int64_t first_record = 0;
  if (n > 0) {
    d_buf = dev_alloc(ctx, (size_t)n);
    SQ_CUDA(cudaMemcpyAsync(d_buf->p, buf.data(), (size_t)n, cudaMemcpyHostToDevice, ctx.stream));
    const int64_t chunks = div_up(n, kCsvChunk);
    BufPtr quotes = dev_alloc(ctx, (size_t)chunks * 4), q_before = dev_alloc(ctx, (size_t)chunks * 8 + 8), scratch = dev_alloc(ctx, scan_scratch_entries(chunks) * 8);
    BufPtr nl_counts = dev_alloc(ctx, (size_t)chunks * 4), nl_offsets = dev_alloc(ctx, (size_t)chunks * 8 + 8);
    launch_csv_count_quotes((const char*)d_buf->p, n, (uint32_t*)quotes->p, ctx.stream);
    launch_scan_u32_large((const uint32_t*)quotes->p, chunks, (unsigned long long*)q_before->p, (unsigned long long*)q_before->p + chunks,
                          (unsigned long long*)scratch->p, ctx.stream);
    launch_csv_newlines((const char*)d_buf->p, n, (const unsigned long long*)q_before->p, (uint32_t*)nl_counts->p, nullptr, nullptr, ctx.stream);
    launch_scan_u32_large((const uint32_t*)nl_counts->p, chunks, (unsigned long long*)nl_offsets->p, (unsigned long long*)nl_offsets->p + chunks,
                          (unsigned long long*)scratch->p, ctx.stream);
    unsigned long long records = 0;
    SQ_CUDA(cudaMemcpyAsync(&records, (unsigned long long*)nl_offsets->p + chunks, 8, cudaMemcpyDeviceToHost, ctx.stream));
    SQ_CUDA(cudaStreamSynchronize(ctx.stream));
    line_end = dev_alloc(ctx, std::max<size_t>((size_t)records, 1) * 8);
    launch_csv_newlines((const char*)d_buf->p, n, (const unsigned long long*)q_before->p, nullptr, (const unsigned long long*)nl_offsets->p, (int64_t*)line_end->p,
                        ctx.stream);
    const int64_t header = opt.has_header ? 1 : 0;
    int64_t data_rows = std::max<int64_t>((int64_t)records - header, 0);
    // bounds (csv.rs:196-206 + arrow-csv's line counting [ext]): `offset` data rows are skipped, then `limit` rows are read —
    // one more when the file has no header (the reader's end line is offset + limit + 1 and its line counter starts at offset)
    int64_t skip = 0, take = data_rows;
    if (opt.bounds_offset >= 0) {
      skip = std::min(opt.bounds_offset, data_rows);
      take = data_rows - skip;
      if (opt.bounds_limit >= 0) take = std::min(take, opt.bounds_limit + (opt.has_header ? 0 : 1));
    }
    first_record = header + skip;
    n_rows = take;
  }
  whole.n = n_rows;

  // ---- parse the projected columns, one thread per record
  CsvColumns cols{};
  cols.n = (int)proj.size();
  std::vector<BufPtr> data(proj.size()), aux(proj.size()), valid_bytes(proj.size());
  const size_t rows1 = (size_t)std::max<int64_t>(n_rows, 1);
  for (size_t k = 0; k < proj.size(); k++) {
    const int dt = fields[(size_t)proj[k]].dtype;
    cols.field[k] = proj[k];
    cols.dtype[k] = dt;
    data[k] = dev_alloc(ctx, rows1 * (dt == SQLRS_DT_BOOL ? 1 : 8));
    valid_bytes[k] = dev_alloc(ctx, rows1);
    if (dt == SQLRS_DT_UTF8) aux[k] = dev_alloc(ctx, rows1 * 4);
    cols.data[k] = data[k]->p;
    cols.aux[k] = aux[k] ? aux[k]->p : nullptr;
    cols.valid[k] = valid_bytes[k]->p;
  }
  BufPtr flags = dev_alloc_zero(ctx, 4);
  uint32_t hflags = 0;
  if (n_rows > 0) {
    launch_csv_parse((const char*)d_buf->p, (const int64_t*)line_end->p, first_record, n_rows, opt.delimiter, cols, (uint32_t*)flags->p, ctx.stream);
    SQ_CUDA(cudaMemcpyAsync(&hflags, flags->p, 4, cudaMemcpyDeviceToHost, ctx.stream));
    SQ_CUDA(cudaStreamSynchronize(ctx.stream));
    if (hflags & 1u) fail(SQLRS_ERR_ARROW, "Parser error: a value of " + path + " does not parse as its column's inferred type");
  }
  for (size_t k = 0; k < proj.size(); k++) {
    const int dt = fields[(size_t)proj[k]].dtype;
    DCol col;
    col.dtype = dt;
    col.n = n_rows;
    std::vector<uint8_t> hvalid((size_t)n_rows);
    if (n_rows) SQ_CUDA(cudaMemcpy(hvalid.data(), valid_bytes[k]->p, (size_t)n_rows, cudaMemcpyDeviceToHost));
    int64_t nulls = 0;
    for (uint8_t v : hvalid) nulls += v == 0;
    if (dt == SQLRS_DT_UTF8) {  // strings: intern on the host (offset, length) -> pool ids
      std::vector<int64_t> offs((size_t)n_rows);
      std::vector<int32_t> lens((size_t)n_rows);
      if (n_rows) {
        SQ_CUDA(cudaMemcpy(offs.data(), data[k]->p, (size_t)n_rows * 8, cudaMemcpyDeviceToHost));
        SQ_CUDA(cudaMemcpy(lens.data(), aux[k]->p, (size_t)n_rows * 4, cudaMemcpyDeviceToHost));
      }
      StringPool& pool = StringPool::instance();
      std::string tmp;
      for (int64_t r = 0; r < n_rows; r++) {
        if (!hvalid[(size_t)r]) {
          offs[(size_t)r] = 0;
          continue;
        }
        const char* s = buf.data() + offs[(size_t)r];
        const size_t len = (size_t)lens[(size_t)r];
        if (memmem(s, len, "\"\"", 2)) {  // escaped quotes inside a quoted field
          tmp.clear();
          for (size_t i = 0; i < len; i++) {
            tmp += s[i];
            if (s[i] == '"' && i + 1 < len && s[i + 1] == '"') i++;
          }
          offs[(size_t)r] = pool.intern(tmp.data(), tmp.size());
        } else {
          offs[(size_t)r] = pool.intern(s, len);
        }
      }
      if (n_rows) SQ_CUDA(cudaMemcpy(data[k]->p, offs.data(), (size_t)n_rows * 8, cudaMemcpyHostToDevice));
      col.data = data[k]->p;
      col.keep_data = data[k];
    } else if (dt == SQLRS_DT_BOOL) {
      BufPtr words = dev_alloc_zero(ctx, (size_t)bitmap_words(std::max<int64_t>(n_rows, 1)) * 4);
      launch_pack_bytes((const uint8_t*)data[k]->p, n_rows, (uint32_t*)words->p, ctx.stream);
      col.data = words->p;
      col.keep_data = words;
    } else {
      if (dt == SQLRS_DT_FLOAT64 && (hflags & 2u)) {  // some value is outside the exact fast path: the host parses this column (strtod)
        std::vector<double> vals((size_t)n_rows, 0.0);
        size_t p = 0;
        for (int64_t rec = 0; rec < first_record + n_rows && p < buf.size(); rec++) {
          p = split_record(buf, p, opt.delimiter, &cells, &quoted);
          if (rec >= first_record && (size_t)proj[k] < cells.size() && !cells[(size_t)proj[k]].empty())
            vals[(size_t)(rec - first_record)] = std::strtod(cells[(size_t)proj[k]].c_str(), nullptr);
        }
        if (n_rows) SQ_CUDA(cudaMemcpy(data[k]->p, vals.data(), (size_t)n_rows * 8, cudaMemcpyHostToDevice));
      }
      col.data = data[k]->p;
      col.keep_data = data[k];
    }
    if (nulls > 0) {
      BufPtr words = dev_alloc_zero(ctx, (size_t)bitmap_words(n_rows) * 4);
      launch_pack_bytes((const uint8_t*)valid_bytes[k]->p, n_rows, (uint32_t*)words->p, ctx.stream);
      col.valid = (const uint32_t*)words->p;
      col.keep_valid = words;
      col.null_count = nulls;
    }
    whole.cols.push_back(col);
  }
  SQ_CUDA(cudaStreamSynchronize(ctx.stream));
  // ---- batches of batch_rows rows: zero-copy slices where the 32-row bitmap granularity allows, i.e. batch_rows % 32 == 0
  const int64_t br = opt.batch_rows > 0 ? opt.batch_rows : 1024;
  if (br % 32 != 0) fail(SQLRS_ERR_INVALID_ARG, "csv: batch_rows must be a multiple of 32");
  if (n_rows == 0) return out;  // the reference's reader yields no batch for an empty range
  for (int64_t off = 0; off < n_rows; off += br) {
    DBatch b;
    b.fields = whole.fields;
    b.n = std::min(br, n_rows - off);
    for (const DCol& c : whole.cols) {
      DCol s = c;
      s.n = b.n;
      if (c.dtype == SQLRS_DT_BOOL) s.data = (const uint32_t*)c.data + (off >> 5);
      else s.data = (const uint8_t*)c.data + (size_t)off * dtype_width(c.dtype);
      if (c.valid) {
        s.valid = c.valid + (off >> 5);
        s.null_count = n_rows == b.n ? c.null_count : -1;
      }
      b.cols.push_back(s);
    }
    out.push_back(std::move(b));
  }
  return out;
}

}  // namespace sq
