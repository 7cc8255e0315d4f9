// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's CSV scan: CsvTable / CsvTransaction
// (src/storage/csv.rs:99-235) over arrow-csv 28.0.0's Reader [ext: not under /root/reference; Cargo.lock:43-181].
// Behaviour restated from the reference's call site and arrow-csv's documented algorithm:
//   * CsvConfig::default (csv.rs:99-109): header row, ',' delimiter, schema inferred from the first 10 records, 1024-row batches;
//   * inference per field (arrow-csv infer_field_schema): quoted -> Utf8; (?i)^(true)$|^(false)$ -> Boolean; ^-?(\d+)$ -> Int64;
//     ^-?(\d+\.\d+)$ -> Float64; otherwise Utf8; empty cells do not vote; {Int64, Float64} -> Float64; any other mix / no vote -> Utf8;
//     every field nullable;
//   * an empty cell is NULL in a Boolean / Int64 / Float64 column and the empty string in a Utf8 column (pinned by
//     tests/slt/aggregation.slt:19-34: employee 4 has salary NULL and state '');
//   * bounds (csv.rs:196-206): offset data rows are skipped, then limit rows are read (limit + 1 without a header: the reader's end
//     line is offset + limit + 1 while its line counter starts at offset, or offset + 1 after a header);
//   * projection selects and orders the output columns.
#pragma once
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <iterator>

#include "columns.hpp"

namespace oracle {

struct CsvRecord {
  std::vector<std::string> cells;
  std::vector<bool> quoted;
};

inline bool csv_next_record(const std::string& text, size_t& pos, char delim, CsvRecord& rec) {
  if (pos >= text.size()) return false;
  rec.cells.clear();
  rec.quoted.clear();
  std::string cur;
  bool inside = false, was_quoted = false;
  while (pos < text.size()) {
    const char c = text[pos++];
    if (inside) {
      if (c == '"') {
        if (pos < text.size() && text[pos] == '"') {
          cur.push_back('"');
          pos++;
        } else {
          inside = false;
        }
      } else {
        cur.push_back(c);
      }
      continue;
    }
    if (c == '"' && cur.empty() && !was_quoted) {
      inside = was_quoted = true;
    } else if (c == delim) {
      rec.cells.push_back(cur);
      rec.quoted.push_back(was_quoted);
      cur.clear();
      was_quoted = false;
    } else if (c == '\n') {
      break;
    } else if (c == '\r' && pos < text.size() && text[pos] == '\n') {
      // CRLF: dropped with the line feed
    } else {
      cur.push_back(c);
    }
  }
  rec.cells.push_back(cur);
  rec.quoted.push_back(was_quoted);
  return true;
}

inline int csv_vote(const std::string& s, bool quoted) {
  if (quoted) return SQLRS_DT_UTF8;
  std::string low;
  for (char c : s) low.push_back((char)std::tolower((unsigned char)c));
  if (low == "true" || low == "false") return SQLRS_DT_BOOL;
  size_t i = (!s.empty() && s[0] == '-') ? 1 : 0;
  size_t digits = 0, dots = 0, after = 0;
  for (size_t k = i; k < s.size(); k++) {
    if (std::isdigit((unsigned char)s[k])) {
      digits++;
      if (dots) after++;
    } else if (s[k] == '.' && dots == 0 && digits > 0) {
      dots++;
    } else {
      return SQLRS_DT_UTF8;
    }
  }
  if (digits == 0) return SQLRS_DT_UTF8;
  if (dots == 0) return SQLRS_DT_INT64;
  return after > 0 ? SQLRS_DT_FLOAT64 : SQLRS_DT_UTF8;
}

inline std::vector<Batch> read_csv(const std::string& path, bool has_header, char delim, int64_t batch_rows, int64_t bounds_offset, int64_t bounds_limit,
                                   const std::vector<int>& projection_in) {
  std::ifstream f(path, std::ios::binary);
  if (!f) fail(SQLRS_ERR_STORAGE, "io error: cannot open " + path);
  const std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  size_t pos = 0;
  CsvRecord rec;
  std::vector<std::string> names;
  if (has_header && csv_next_record(text, pos, delim, rec)) names = rec.cells;
  const size_t data_begin = pos;
  std::vector<int> votes;
  for (int r = 0; r < 10 && csv_next_record(text, pos, delim, rec); r++) {
    if (votes.size() < rec.cells.size()) votes.resize(rec.cells.size(), 0);
    for (size_t c = 0; c < rec.cells.size(); c++)
      if (!rec.cells[c].empty() || rec.quoted[c]) votes[c] |= 1 << csv_vote(rec.cells[c], rec.quoted[c]);
  }
  const size_t n_fields = std::max(names.size(), votes.size());
  votes.resize(n_fields, 0);
  std::vector<Field> fields(n_fields);
  for (size_t c = 0; c < n_fields; c++) {
    fields[c].name = has_header && c < names.size() ? names[c] : "column_" + std::to_string(c + 1);
    fields[c].nullable = true;
    const int v = votes[c];
    fields[c].dtype = v == (1 << SQLRS_DT_BOOL) ? SQLRS_DT_BOOL
                      : v == (1 << SQLRS_DT_INT64) ? SQLRS_DT_INT64
                      : (v == (1 << SQLRS_DT_FLOAT64) || v == ((1 << SQLRS_DT_INT64) | (1 << SQLRS_DT_FLOAT64))) ? SQLRS_DT_FLOAT64
                                                                                                                  : SQLRS_DT_UTF8;
  }
  std::vector<int> projection = projection_in;
  if (projection.empty())
    for (size_t c = 0; c < n_fields; c++) projection.push_back((int)c);
  for (int c : projection)
    if (c < 0 || c >= (int)n_fields) fail(SQLRS_ERR_ARROW, "csv: projection index out of bounds");
  int64_t skip = 0, take = -1;
  if (bounds_offset >= 0) {
    skip = bounds_offset;
    if (bounds_limit >= 0) take = bounds_limit + (has_header ? 0 : 1);
  }
  std::vector<Batch> out;
  pos = data_begin;
  int64_t row = 0, taken = 0;
  std::vector<std::shared_ptr<Column>> cur;
  auto flush = [&]() {
    if (cur.empty() || cur[0]->n == 0) return;
    Batch b;
    for (size_t k = 0; k < projection.size(); k++) {
      b.fields.push_back(fields[(size_t)projection[k]]);
      cur[k]->normalize();
      b.cols.push_back(cur[k]);
    }
    b.n = cur[0]->n;
    out.push_back(std::move(b));
    cur.clear();
  };
  while (csv_next_record(text, pos, delim, rec)) {
    if (row++ < skip) continue;
    if (take >= 0 && taken >= take) break;
    taken++;
    if (cur.empty())
      for (int c : projection) {
        auto col = std::make_shared<Column>();
        col->dtype = fields[(size_t)c].dtype;
        cur.push_back(col);
      }
    for (size_t k = 0; k < projection.size(); k++) {
      Column& col = *cur[k];
      const size_t c = (size_t)projection[k];
      const bool present = c < rec.cells.size();
      const std::string cell = present ? rec.cells[c] : std::string();
      const bool quoted = present && rec.quoted[c];
      Scalar v = Scalar::null_of(col.dtype);
      if (col.dtype == SQLRS_DT_UTF8) {
        if (present) {
          v.is_null = false;
          v.s = cell;
        }
      } else if (!cell.empty() || quoted) {
        v.is_null = false;
        char* end = nullptr;
        if (col.dtype == SQLRS_DT_INT64) {
          v.i = std::strtoll(cell.c_str(), &end, 10);
          if (end == cell.c_str() || *end) fail(SQLRS_ERR_ARROW, "Parser error: Error while parsing value " + cell + " for column " + std::to_string(c));
        } else if (col.dtype == SQLRS_DT_FLOAT64) {
          v.f = std::strtod(cell.c_str(), &end);
          if (end == cell.c_str() || *end) fail(SQLRS_ERR_ARROW, "Parser error: Error while parsing value " + cell + " for column " + std::to_string(c));
        } else {
          std::string low;
          for (char ch : cell) low.push_back((char)std::tolower((unsigned char)ch));
          if (low != "true" && low != "false") fail(SQLRS_ERR_ARROW, "Parser error: Error while parsing value " + cell + " for column " + std::to_string(c));
          v.i = low == "true";
        }
      }
      append_scalar(col, v);
    }
    if (cur[0]->n == batch_rows) flush();
  }
  flush();
  return out;
}

}  // namespace oracle
