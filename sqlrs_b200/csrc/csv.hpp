// sqlrs_b200 — CSV ingest onto the device (kernels_csv.cu, csv.cpp).  Reference: src/storage/csv.rs.
#pragma once
#include "device.hpp"

namespace sq {

constexpr int kCsvChunk = 256;      // bytes one thread indexes
constexpr int kCsvMaxColumns = 64;  // projected columns per read

struct CsvColumns {  // kernel parameter block of k_csv_parse
  int n;
  int field[kCsvMaxColumns];   // field number in the record
  int dtype[kCsvMaxColumns];
  void* data[kCsvMaxColumns];  // Int64 / Float64: values; Utf8: int64 byte offsets; Boolean: one byte per row
  void* aux[kCsvMaxColumns];   // Utf8: int32 lengths
  void* valid[kCsvMaxColumns]; // one byte per row
};

void launch_csv_count_quotes(const char* buf, int64_t n, uint32_t* quotes, cudaStream_t stream);
void launch_csv_newlines(const char* buf, int64_t n, const unsigned long long* quotes_before, uint32_t* counts, const unsigned long long* line_offsets,
                         int64_t* positions, cudaStream_t stream);
void launch_csv_parse(const char* buf, const int64_t* line_end, int64_t first_record, int64_t n_rows, char delimiter, const CsvColumns& cols, uint32_t* flags,
                      cudaStream_t stream);
void launch_pack_bytes(const uint8_t* bytes, int64_t n, uint32_t* words, cudaStream_t stream);

struct CsvOptions {
  bool has_header = true;            // csv.rs:102
  char delimiter = ',';              // :103
  int64_t infer_max_records = 10;    // :104
  int64_t batch_rows = 1024;         // :105
  int64_t bounds_offset = -1, bounds_limit = -1;  // Bounds = Option<(offset, limit)> over the whole table (:196-206); -1 = none
  std::vector<int> projection;       // empty = every column
};
// the file as device-resident batches of batch_rows rows (zero-copy slices of whole columns) + its inferred fields
std::vector<DBatch> read_csv_device(Ctx& ctx, const std::string& path, const CsvOptions& opt);

}  // namespace sq
