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
// shape of the TMA ring of csrc/jit/joinagg.cuh: sq_joinagg_tma_kernel — rows per tile (8 B x rows per staged column), stages,
// consumer warps per CTA.  Defaults 1024 / 4 / 8; SQLRS_B200_TMA_TROWS / _STAGES / _CONSUMERS override them (experiments).
struct TmaShape {
  int tile_rows, stages, consumers;
};
const TmaShape& tma_shape();
struct ProbeProgram {
  std::string src;
  std::vector<int> tile_cols;
};
ProbeProgram gen_probe_program(const std::vector<ColInfo>& cols, const std::vector<ExprCopy>& right_keys, const ExprCopy& probe_pred, bool jmatch);
// struct SqInB + the SQ_LDB_* / SQ_VALIDB macros: the build side of a joined-mode row program (RowProgram(build_cols, probe_cols))
std::string gen_build_decls(const std::vector<ColInfo>& build_cols);

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
  const std::vector<Field>& out_fields() const { return out_fields_; }
  // a build side produced elsewhere (JoinChainOp: the previous join's probe built this table directly): `build` is the
  // (virtual) build batch the table's row ids index, `keep` the buffers the view points into
  void adopt_build(DBatch build, const struct JoinTableView& view, std::vector<BufPtr> keep);
  // A caller that validates the run afterwards (JoinChainOp) may pass what the previous run of the same plan learned: the
  // number of build rows its fused Filter kept.  seal() then sizes the table from it, assumes unique keys and does NOT
  // synchronise; it copies {flags u32[6] (kernels_aot.hpp: launch_join_insert_kv), pad, kept rows u64} to `pinned`
  // (8 x u32 + u64, pinned host memory) on the stream for the caller to check.  Only taken for the kv layout.
  void set_build_hint(int64_t kept_rows, uint32_t* pinned) { hint_kept_ = kept_rows; hint_pinned_ = pinned; }
  bool sealed_deferred() const { return sealed_deferred_; }
  // plan executor: one build batch, one compared key, no Left / Full tail -> scan + Filter + hash + insert in ONE kernel
  // (csrc/jit/joinbuild.cuh) at seal(); anything else (several batches, repeated keys, ...) takes the materialising path
  void enable_fused_build();
  int64_t build_rows() const;
  // generated CUDA of the fused probe kernel for probe batches of that schema (diagnostics / build check, no GPU)
  std::string debug_probe_source(const std::vector<ColInfo>& probe_cols, const ExprCopy& probe_pred) const;

 private:
  struct Impl;
  DBatch build_batch(const DBatch& right, const int64_t* li, bool li_nullable, const uint32_t* ri, int64_t m);
  void eval_push(const DBatch& batch);  // the materialising build: key / hash / keep columns of one batch
  bool seal_fused();
  void check_schema(DBatch& b);

  Ctx ctx_;
  Options opt_;
  int join_type_;
  std::vector<ExprCopy> left_keys_, right_keys_;
  ExprCopy filter_;
  std::vector<Field> out_fields_;
  std::unique_ptr<Impl> impl_;
  int64_t hint_kept_ = -1;
  uint32_t* hint_pinned_ = nullptr;
  bool sealed_deferred_ = false;
};

// Left-deep join chains: join 1's probe builds join 2's table directly (csrc/jit/joinchain.cuh), nothing of join 1's output
// is materialised.  Applies to: INNER joins with key comparison, join 1 without non-equi filter and with unique build keys
// (found at seal), ONE probe batch, join 2 with a single key over join 1's probe-side columns, and no build-1 column read
// above.  Table 2 is sized from a strided sample count of join 1's matches, or — on repeated runs of the same plan over
// tables of the same size — from the previous run's exact count; in that case nothing synchronises and the run is
// validated afterwards (validate()): a table that turned out too small, a repeated or unrepresentable key make the
// plan fall back to the operator-at-a-time path.
class JoinChainOp {
 public:
  explicit JoinChainOp(const Options& opt);
  ~JoinChainOp();
  // false = the chain does not apply (caller builds join 2 the ordinary way)
  bool run(JoinOp& j1, const DBatch& probe1, const ExprCopy& probe_pred1, const ExprCopy& key2, JoinOp& j2);
  bool pending() const { return pending_; }
  bool validate();  // after the stream has been synchronised; false = results of this run must be discarded
  bool disabled() const { return disabled_; }
  std::string debug_source(const std::vector<ColInfo>& build_cols, const std::vector<ColInfo>& probe_cols, const std::vector<ExprCopy>& right_keys1,
                           const ExprCopy& probe_pred1, const ExprCopy& key2, int* key_dtype = nullptr);

 private:
  bool check_flags();
  Ctx ctx_;
  Options opt_;
  std::map<std::string, std::pair<JitKernel*, int>> kernels_;  // by schema signature: kernel + dtype of the chain key
  struct Host {  // pinned
    uint32_t flags[4];
    unsigned long long inserted;
    uint32_t j1_flags[8];          // join 1's deferred seal (JoinOp::set_build_hint)
    unsigned long long j1_kept;
  };
  Host* host_ = nullptr;
  int64_t hint_j1_kept_ = -1;
  bool j1_deferred_ = false;
  int64_t hint_inserted_ = -1, hint_probe_rows_ = -1, hint_build_rows_ = -1;
  uint64_t cap_used_ = 0;
  bool pending_ = false, disabled_ = false;
};

}  // namespace sq
