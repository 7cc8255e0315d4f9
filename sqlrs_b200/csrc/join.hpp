// sqlrs_b200 — HashJoinExecutor on the GPU (reference src/executor/join/hash_join.rs:16-323).
// Build side = left child, fully drained and kept resident; one output batch per probe batch in the
// reference's row order (probe-row order, per probe row the build rows in insertion order); Left/Full
// tail of unmatched build rows at finish.
#pragma once
#include "kernels_aot.hpp"
#include "ops.hpp"

namespace sq {

// generated CUDA of the probe side's row program (after gen_input_decls): SQ_JKEYS, SQ_JMATCH, struct SqProbe and
// sq_probe_row(in, r, p, e0, e1) = Filter fused below the join on the probe side (optional) + the join key
// expressions, their create_hashes row hash, raw key bits and null mask.  Shared by csrc/jit/joinprobe.cuh and joinagg.cuh.
// When every column that program reads is an 8-byte column (at most 3 of them), `tile_cols` lists them and the source
// also contains sq_probe_row_tile(in, tile, t, r, ...): the same program reading those columns from a shared-memory
// tile staged by TMA bulk copies (SQ_TMA 1, SQ_TILE_NCOLS, SQ_TILE_COLS) — csrc/jit/joinagg.cuh: sq_joinagg_tma_kernel.
struct ProbeProgram {
  std::string src;
  std::vector<int> tile_cols;
};
ProbeProgram gen_probe_program(const std::vector<ColInfo>& cols, const std::vector<ExprCopy>& right_keys, const ExprCopy& probe_pred, bool jmatch);

class JoinOp {
 public:
  JoinOp(int join_type, std::vector<ExprCopy> left_keys, std::vector<ExprCopy> right_keys, ExprCopy filter,
         std::vector<Field> out_fields, const Options& opt);
  ~JoinOp();
  void build_push(const DBatch& batch);
  bool probe(const DBatch& right, DBatch* out);  // false where the reference yields nothing (empty build side)
  bool finish(DBatch* out);                      // Left/Full tail
  // plan executor only: Filters directly below the join evaluated inside the key kernels (rows keep their
  // original ids, results are identical), and output columns nobody above reads left out of the gathers
  void set_side_predicates(ExprCopy build_pred, ExprCopy probe_pred);
  void set_needed_columns(std::vector<bool> needed);
  Ctx& ctx() { return ctx_; }
  // for consumers fused onto the probe (AggOp::push_join): the sealed build side
  void seal();
  bool empty_build() const;                 // no build batch at all: the reference yields nothing (hash_join.rs:183-185)
  const struct JoinTableView& table_view() const;
  const DBatch& build_side() const;
  int join_type() const { return join_type_; }
  const std::vector<ExprCopy>& right_keys() const { return right_keys_; }
  const ExprCopy& join_filter() const { return filter_; }
  bool match_keys() const { return opt_.match_mode == SQLRS_MATCH_HASH_AND_KEY; }
  // generated CUDA of the fused probe kernel for probe batches of that schema (diagnostics / build check, no GPU)
  std::string debug_probe_source(const std::vector<ColInfo>& probe_cols, const ExprCopy& probe_pred) const;

 private:
  struct Impl;
  DBatch build_batch(const DBatch& right, const int64_t* li, bool li_nullable, const uint32_t* ri, int64_t m);
  void check_schema(DBatch& b);

  Ctx ctx_;
  Options opt_;
  int join_type_;
  std::vector<ExprCopy> left_keys_, right_keys_;
  ExprCopy filter_;
  std::vector<Field> out_fields_;
  std::unique_ptr<Impl> impl_;
};

}  // namespace sq
