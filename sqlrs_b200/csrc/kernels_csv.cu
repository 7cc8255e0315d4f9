// sqlrs_b200 — CSV -> Arrow columns on the device (the step before the hot path).
// Reference: src/storage/csv.rs:99-109 (has_header, ',' delimiter, 1024-row batches) and :190-235 (arrow-csv Reader with
// bounds and projection).  The file's bytes are copied to HBM once; three passes index the records — quotes per chunk ->
// (scan) quote parity at every chunk start -> record-terminating newlines per chunk -> (scan) -> their positions — and one
// thread per record then splits its fields and parses the projected ones.  HBM-bound byte work: the file is read 3 times.
#include "csv.hpp"

namespace sq {
namespace {

constexpr int kBlock = 256;
inline unsigned grid_for(int64_t items) {
  int64_t g = div_up(items, kBlock);
  return (unsigned)std::min<int64_t>(std::max<int64_t>(g, 1), 148 * 16);
}

__global__ void __launch_bounds__(kBlock) k_csv_count_quotes(const char* __restrict__ buf, int64_t n, int64_t n_chunks, uint32_t* __restrict__ quotes) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += stride) {
    const int64_t lo = c * kCsvChunk, hi = lo + kCsvChunk < n ? lo + kCsvChunk : n;
    uint32_t q = 0;
    for (int64_t i = lo; i < hi; i++) q += buf[i] == '"';
    quotes[c] = q;
  }
}

// newlines outside quotes: counted (positions == nullptr) or written at line_offsets[c] + k
__global__ void __launch_bounds__(kBlock) k_csv_newlines(const char* __restrict__ buf, int64_t n, int64_t n_chunks, const unsigned long long* __restrict__ quotes_before,
                                                         uint32_t* __restrict__ counts, const unsigned long long* __restrict__ line_offsets,
                                                         int64_t* __restrict__ positions) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += stride) {
    const int64_t lo = c * kCsvChunk, hi = lo + kCsvChunk < n ? lo + kCsvChunk : n;
    bool inside = (quotes_before[c] & 1ULL) != 0;
    uint32_t k = 0;
    unsigned long long at = positions ? line_offsets[c] : 0;
    for (int64_t i = lo; i < hi; i++) {
      const char ch = buf[i];
      if (ch == '"') inside = !inside;
      else if (ch == '\n' && !inside) {
        if (positions) positions[at + k] = i;
        k++;
      }
    }
    if (!positions) counts[c] = k;
  }
}

__device__ __forceinline__ bool csv_is_space(char c) { return c == ' ' || c == '\t'; }

// one thread per record: split into fields (quotes honoured), parse the projected columns
__global__ void __launch_bounds__(kBlock) k_csv_parse(const char* __restrict__ buf, const int64_t* __restrict__ line_end, int64_t first_record, int64_t n_rows,
                                                      char delimiter, const __grid_constant__ CsvColumns cols, uint32_t* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
    const int64_t rec = first_record + r;
    int64_t pos = rec == 0 ? 0 : line_end[rec - 1] + 1;
    int64_t end = line_end[rec];
    if (end > pos && buf[end - 1] == '\r') end--;
    int field = 0;
    while (true) {
      // field [b, e): quoted fields lose their quotes; `escaped` when it contains "" (the host unescapes strings)
      int64_t b = pos, e;
      bool quoted = false;
      if (pos < end && buf[pos] == '"') {
        quoted = true;
        b = pos + 1;
        int64_t i = b;
        while (i < end) {
          if (buf[i] == '"') {
            if (i + 1 < end && buf[i + 1] == '"') i += 2;
            else break;
          } else i++;
        }
        e = i;
        pos = i < end ? i + 1 : end;
        while (pos < end && buf[pos] != delimiter) pos++;
      } else {
        while (pos < end && buf[pos] != delimiter) pos++;
        e = pos;
      }
      // which projected columns want this field?
      for (int c = 0; c < cols.n; c++) {
        if (cols.field[c] != field) continue;
        const int dt = cols.dtype[c];
        bool valid = e > b || quoted;
        if (dt == SQLRS_DT_UTF8) {
          ((int64_t*)cols.data[c])[r] = b;                 // byte offset; the host interns the strings
          ((int32_t*)cols.aux[c])[r] = (int32_t)(e - b);
          valid = true;                                     // arrow-csv [ext]: an empty Utf8 field is the empty string, not NULL
        } else if (dt == SQLRS_DT_INT64) {
          int64_t i = b;
          bool neg = false;
          if (i < e && (buf[i] == '-' || buf[i] == '+')) neg = buf[i++] == '-';
          unsigned long long v = 0;
          bool ok = i < e;
          for (; i < e; i++) {
            const unsigned d = (unsigned)(buf[i] - '0');
            if (d > 9u) {
              ok = false;
              break;
            }
            v = v * 10ULL + d;
          }
          if (valid && !ok) atomicOr(flags, 1u);  // not an integer: the host reports the parse error
          ((int64_t*)cols.data[c])[r] = valid ? (neg ? (int64_t)(0ULL - v) : (int64_t)v) : 0;
        } else if (dt == SQLRS_DT_FLOAT64) {
          // fast path of a correctly rounded decimal -> double conversion: <= 15 significant digits and |exponent| <= 22 is ONE
          // exact operation on exactly representable operands; anything else is flagged and parsed on the host (strtod)
          int64_t i = b;
          bool neg = false;
          if (i < e && (buf[i] == '-' || buf[i] == '+')) neg = buf[i++] == '-';
          unsigned long long m = 0;
          int digits = 0, exp10 = 0;
          bool ok = i < e, seen_dot = false, any = false;
          for (; i < e; i++) {
            const char ch = buf[i];
            if (ch == '.' && !seen_dot) {
              seen_dot = true;
              continue;
            }
            const unsigned d = (unsigned)(ch - '0');
            if (d > 9u) break;
            any = true;
            if (digits < 18) {
              m = m * 10ULL + d;
              if (m) digits++;
              if (seen_dot) exp10--;
            } else if (!seen_dot) {
              exp10++;
            }
          }
          if (i < e && (buf[i] == 'e' || buf[i] == 'E')) {
            i++;
            bool eneg = false;
            if (i < e && (buf[i] == '-' || buf[i] == '+')) eneg = buf[i++] == '-';
            int x = 0;
            bool edig = false;
            for (; i < e && (unsigned)(buf[i] - '0') <= 9u; i++) {
              x = x < 10000 ? x * 10 + (buf[i] - '0') : x;
              edig = true;
            }
            ok = ok && edig;
            exp10 += eneg ? -x : x;
          }
          ok = ok && any && i == e;
          double v = 0.0;
          if (ok && valid) {
            if (m == 0) v = 0.0;
            else if (digits <= 15 && exp10 >= -22 && exp10 <= 22) {
              const double p10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
              v = exp10 >= 0 ? __dmul_rn((double)m, p10[exp10]) : __ddiv_rn((double)m, p10[-exp10]);
            } else {
              atomicOr(flags, 2u);  // slow path: the host re-parses this column
            }
          }
          if (valid && !ok) atomicOr(flags, 1u);
          ((double*)cols.data[c])[r] = valid ? (neg ? -v : v) : 0.0;
        } else if (dt == SQLRS_DT_BOOL) {
          const int64_t len = e - b;
          const bool t = len == 4 && (buf[b] | 32) == 't' && (buf[b + 1] | 32) == 'r' && (buf[b + 2] | 32) == 'u' && (buf[b + 3] | 32) == 'e';
          const bool f = len == 5 && (buf[b] | 32) == 'f' && (buf[b + 1] | 32) == 'a' && (buf[b + 2] | 32) == 'l' && (buf[b + 3] | 32) == 's' && (buf[b + 4] | 32) == 'e';
          if (valid && !t && !f) atomicOr(flags, 1u);
          ((uint8_t*)cols.data[c])[r] = t ? 1 : 0;  // one byte per row here; packed by the host side of the reader
        }
        ((uint8_t*)cols.valid[c])[r] = valid ? 1 : 0;
      }
      field++;
      if (pos >= end) break;
      pos++;  // the delimiter
      if (pos == end) {  // trailing delimiter: one more, empty field
        for (int c = 0; c < cols.n; c++)
          if (cols.field[c] == field) {
            const bool str = cols.dtype[c] == SQLRS_DT_UTF8;
            ((uint8_t*)cols.valid[c])[r] = str ? 1 : 0;
            if (str) {
              ((int64_t*)cols.data[c])[r] = end;
              ((int32_t*)cols.aux[c])[r] = 0;
            }
          }
        field++;
        break;
      }
    }
    // missing trailing fields are NULL
    for (int c = 0; c < cols.n; c++)
      if (cols.field[c] >= field) ((uint8_t*)cols.valid[c])[r] = 0;
  }
}

// byte-per-row flags -> packed bitmap words (validity, Boolean values)
__global__ void __launch_bounds__(kBlock) k_pack_bytes(const uint8_t* __restrict__ bytes, int64_t n, uint32_t* __restrict__ words) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_up = (n + 31) & ~(int64_t)31;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_up; i += stride) {
    const uint32_t w = __ballot_sync(0xffffffffu, i < n && bytes[i] != 0);
    if (lane == 0) words[i >> 5] = w;
  }
}

}  // namespace

void launch_csv_count_quotes(const char* buf, int64_t n, uint32_t* quotes, cudaStream_t stream) {
  const int64_t chunks = div_up(n, kCsvChunk);
  if (chunks <= 0) return;
  k_csv_count_quotes<<<grid_for(chunks), kBlock, 0, stream>>>(buf, n, chunks, quotes);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}
void launch_csv_newlines(const char* buf, int64_t n, const unsigned long long* quotes_before, uint32_t* counts, const unsigned long long* line_offsets,
                         int64_t* positions, cudaStream_t stream) {
  const int64_t chunks = div_up(n, kCsvChunk);
  if (chunks <= 0) return;
  k_csv_newlines<<<grid_for(chunks), kBlock, 0, stream>>>(buf, n, chunks, quotes_before, counts, line_offsets, positions);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}
void launch_csv_parse(const char* buf, const int64_t* line_end, int64_t first_record, int64_t n_rows, char delimiter, const CsvColumns& cols, uint32_t* flags,
                      cudaStream_t stream) {
  if (n_rows <= 0 || cols.n <= 0) return;
  k_csv_parse<<<grid_for(n_rows), kBlock, 0, stream>>>(buf, line_end, first_record, n_rows, delimiter, cols, flags);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}
void launch_pack_bytes(const uint8_t* bytes, int64_t n, uint32_t* words, cudaStream_t stream) {
  if (n <= 0) return;
  k_pack_bytes<<<grid_for(n), kBlock, 0, stream>>>(bytes, n, words);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}

}  // namespace sq
