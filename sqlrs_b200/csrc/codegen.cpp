// sqlrs_b200 — row-program code generator (see codegen.hpp).
#include "codegen.hpp"

#include <cstring>

#include "device.hpp"

#include <cstdio>

namespace sq {

ExprCopy copy_expr(const sqlrs_expr* e) {
  ExprCopy out;
  if (!e) return out;
  if (e->n_nodes > 0 && !e->nodes) fail(SQLRS_ERR_INVALID_ARG, "sqlrs_expr.nodes is NULL");
  for (int32_t k = 0; k < e->n_nodes; k++) {
    ExprNodeCopy n;
    n.op = e->nodes[k].op;
    n.dtype = e->nodes[k].dtype;
    n.index = e->nodes[k].index;
    n.is_null = e->nodes[k].is_null;
    n.imm_bits = e->nodes[k].imm_bits;
    n.has_str = e->nodes[k].str != nullptr;
    if (n.op == SQLRS_OP_CONSTANT && n.dtype == SQLRS_DT_UTF8 && !n.is_null) {  // a string literal becomes its id in the string pool
      const char* str = e->nodes[k].str ? e->nodes[k].str : "";
      n.imm_bits = StringPool::instance().intern(str, std::strlen(str));
    }
    out.push_back(n);
  }
  return out;
}

const char* ctype_of(int dtype) {
  switch (dtype) {
    case SQLRS_DT_BOOL: return "bool";
    case SQLRS_DT_INT32: return "int";
    case SQLRS_DT_INT64: return "long long";
    case SQLRS_DT_FLOAT64: return "double";
  }
  return "long long";  // Null-typed values carry a dummy
}

std::string lit_i64(int64_t v) {
  char buf[48];
  // as an unsigned hex pattern: avoids the LLONG_MIN literal problem
  std::snprintf(buf, sizeof buf, "((long long)0x%llxULL)", (unsigned long long)v);
  return buf;
}
std::string lit_f64_bits(int64_t bits) {
  char buf[64];
  std::snprintf(buf, sizeof buf, "__longlong_as_double((long long)0x%llxULL)", (unsigned long long)bits);
  return buf;
}

static std::string vname(int id) { return "v" + std::to_string(id); }
static std::string nname(int id) { return "n" + std::to_string(id); }

std::string RowProgram::signature() const {
  std::string s;
  for (const ColInfo& c : cols_) {
    s += (char)('0' + c.dtype);
    s += c.has_valid ? 'v' : '-';
    s += c.nullable ? 'n' : '-';
  }
  if (joined_) {
    s += '|';
    for (const ColInfo& c : build_cols_) {
      s += (char)('0' + c.dtype);
      s += c.has_valid ? 'v' : '-';
      s += c.nullable ? 'n' : '-';
    }
  }
  return s;
}

Val RowProgram::define(int dtype, const std::string& value_expr, const std::string& valid_expr, bool maybe_null, bool decl_null,
                       const std::string& cse_key) {
  if (!cse_key.empty()) {
    auto it = cse_.find(cse_key);
    if (it != cse_.end()) return it->second;
  }
  Val v;
  v.id = fresh();
  v.dtype = dtype;
  v.maybe_null = maybe_null;
  v.decl_null = decl_null || maybe_null;
  body_ << "  const " << ctype_of(dtype) << " " << vname(v.id) << " = " << value_expr << ";\n";
  body_ << "  const bool " << nname(v.id) << " = " << (maybe_null ? valid_expr : std::string("true")) << ";\n";
  if (!cse_key.empty()) cse_[cse_key] = v;
  return v;
}

Val RowProgram::load_column(int index) {
  const int nb = joined_ ? (int)build_cols_.size() : 0;
  if (index < 0 || index >= nb + (int)cols_.size()) fail(SQLRS_ERR_INTERNAL, "InputRef index out of bounds");
  if (index < nb) {  // build side of the joined row: gathered at the matched build row
    const ColInfo& c = build_cols_[index];
    std::string key = "colB" + std::to_string(index);
    auto it = cse_.find(key);
    if (it != cse_.end()) return it->second;
    std::string ci = std::to_string(index);
    switch (c.dtype) {
      case SQLRS_DT_NULL: {
        Val v = define(SQLRS_DT_NULL, "0", "false", true, true, key);
        v.always_null = true;
        cse_[key] = v;
        return v;
      }
      case SQLRS_DT_BOOL: return define(c.dtype, "SQ_LDB_BOOL(" + ci + ", b)", "SQ_VALIDB(" + ci + ", b)", c.has_valid, c.nullable, key);
      case SQLRS_DT_INT32: return define(c.dtype, "SQ_LDB_I32(" + ci + ", b)", "SQ_VALIDB(" + ci + ", b)", c.has_valid, c.nullable, key);
      case SQLRS_DT_INT64: return define(c.dtype, "SQ_LDB_I64(" + ci + ", b)", "SQ_VALIDB(" + ci + ", b)", c.has_valid, c.nullable, key);
      case SQLRS_DT_FLOAT64: return define(c.dtype, "SQ_LDB_F64(" + ci + ", b)", "SQ_VALIDB(" + ci + ", b)", c.has_valid, c.nullable, key);
      case SQLRS_DT_UTF8: return define(c.dtype, "SQ_LDB_I64(" + ci + ", b)", "SQ_VALIDB(" + ci + ", b)", c.has_valid, c.nullable, key);  // string pool id
      default: fail(SQLRS_ERR_INVALID_ARG, "unknown column dtype");
    }
  }
  index -= nb;
  const ColInfo& c = cols_[index];
  std::string key = "col" + std::to_string(index);
  auto it = cse_.find(key);
  if (it != cse_.end()) return it->second;
  std::string ci = std::to_string(index);
  Val v;
  switch (c.dtype) {
    case SQLRS_DT_NULL:
      v = define(SQLRS_DT_NULL, "0", "false", true, true, key);
      v.always_null = true;
      cse_[key] = v;
      return v;
    case SQLRS_DT_BOOL: return define(c.dtype, "SQ_LD_BOOL(" + ci + ", r)", "SQ_VALID(" + ci + ", r)", c.has_valid, c.nullable, key);
    case SQLRS_DT_INT32: return define(c.dtype, "SQ_LD_I32(" + ci + ", r)", "SQ_VALID(" + ci + ", r)", c.has_valid, c.nullable, key);
    case SQLRS_DT_INT64: return define(c.dtype, "SQ_LD_I64(" + ci + ", r)", "SQ_VALID(" + ci + ", r)", c.has_valid, c.nullable, key);
    case SQLRS_DT_FLOAT64: return define(c.dtype, "SQ_LD_F64(" + ci + ", r)", "SQ_VALID(" + ci + ", r)", c.has_valid, c.nullable, key);
    case SQLRS_DT_UTF8: return define(c.dtype, "SQ_LD_I64(" + ci + ", r)", "SQ_VALID(" + ci + ", r)", c.has_valid, c.nullable, key);  // string pool id
  }
  fail(SQLRS_ERR_INVALID_ARG, "unknown column dtype");
}

Val RowProgram::constant(const ExprNodeCopy& n) {
  bool is_null = n.dtype == SQLRS_DT_NULL || n.is_null;
  std::string key = "const" + std::to_string(n.dtype) + (is_null ? "N" : "V") + std::to_string(n.imm_bits);
  std::string value;
  switch (n.dtype) {
    case SQLRS_DT_NULL: value = "0"; break;
    case SQLRS_DT_BOOL: value = (!is_null && n.imm_bits != 0) ? "true" : "false"; break;
    case SQLRS_DT_INT32: value = is_null ? "0" : "((int)" + lit_i64((int32_t)n.imm_bits) + ")"; break;
    case SQLRS_DT_INT64: value = is_null ? "0" : lit_i64(n.imm_bits); break;
    case SQLRS_DT_FLOAT64: value = is_null ? "0.0" : lit_f64_bits(n.imm_bits); break;
    case SQLRS_DT_UTF8: value = is_null ? "0" : lit_i64(n.imm_bits); break;  // copy_expr interned the literal: its pool id
    default: fail(SQLRS_ERR_INVALID_ARG, "unknown constant dtype");
  }
  Val v = define(n.dtype, value, "false", is_null, is_null, key);
  v.always_null = is_null;
  cse_[key] = v;
  return v;
}

// arrow compute::cast, the subset reachable from the v1 types (evaluator.rs:23, sum.rs:54):
// numeric casts are "safe" (out of range -> NULL), Null -> T gives all-NULL.
Val RowProgram::cast(const Val& a, int to) {
  if (a.dtype == to) return a;
  std::string key = "cast" + std::to_string(to) + "_" + std::to_string(a.id);
  std::string av = vname(a.id), an = nname(a.id);
  int from = a.dtype;
  auto is_intlike = [](int d) { return d == SQLRS_DT_INT32 || d == SQLRS_DT_INT64 || d == SQLRS_DT_BOOL; };
  if (to == SQLRS_DT_UTF8 || from == SQLRS_DT_UTF8)
    fail(SQLRS_ERR_UNSUPPORTED, "Utf8 casts are not supported by the CUDA backend yet");
  if (from == SQLRS_DT_NULL) {
    if (to == SQLRS_DT_NULL) return a;
    Val v = define(to, to == SQLRS_DT_FLOAT64 ? "0.0" : (to == SQLRS_DT_BOOL ? "false" : "0"), "false", true, true, key);
    v.always_null = true;
    cse_[key] = v;
    return v;
  }
  if (is_intlike(from) && to == SQLRS_DT_INT64) return define(to, "(long long)" + av, an, a.maybe_null, a.decl_null, key);
  if ((from == SQLRS_DT_INT64 || from == SQLRS_DT_BOOL) && to == SQLRS_DT_INT32) {
    if (from == SQLRS_DT_BOOL) return define(to, "(int)" + av, an, a.maybe_null, a.decl_null, key);
    std::string inr = "(" + av + " >= -2147483648LL && " + av + " <= 2147483647LL)";
    return define(to, "(" + inr + " ? (int)" + av + " : 0)", "(" + an + " && " + inr + ")", true, true, key);
  }
  if (is_intlike(from) && to == SQLRS_DT_FLOAT64) return define(to, "(double)" + av, an, a.maybe_null, a.decl_null, key);
  if (from == SQLRS_DT_FLOAT64 && (to == SQLRS_DT_INT64 || to == SQLRS_DT_INT32)) {
    std::string t = "trunc(" + av + ")";
    std::string lo = to == SQLRS_DT_INT64 ? "-9223372036854775808.0" : "-2147483648.0";
    std::string hi = to == SQLRS_DT_INT64 ? "9223372036854775808.0" : "2147483648.0";
    std::string inr = "(" + t + " >= " + lo + " && " + t + " < " + hi + ")";  // false for NaN
    std::string ct = to == SQLRS_DT_INT64 ? "(long long)" : "(int)";
    return define(to, "(" + inr + " ? " + ct + t + " : 0)", "(" + an + " && " + inr + ")", true, true, key);
  }
  if ((from == SQLRS_DT_INT32 || from == SQLRS_DT_INT64) && to == SQLRS_DT_BOOL)
    return define(to, "(" + av + " != 0)", an, a.maybe_null, a.decl_null, key);
  if (from == SQLRS_DT_FLOAT64 && to == SQLRS_DT_BOOL) return define(to, "(" + av + " != 0.0)", an, a.maybe_null, a.decl_null, key);
  fail(SQLRS_ERR_ARROW, std::string("Casting from ") + dtype_name(from) + " to " + dtype_name(to) + " not supported");
}

// arithmetic_op!, array_compute.rs:37-46 — arrow add/subtract/multiply/divide: integers wrap,
// NULL in either operand -> NULL, a valid zero divisor -> Err(DivideByZero).
bool g_checked_arithmetic_compiled = false;

Val RowProgram::arithmetic(const Val& l, const Val& r, int op_in, int err_class) {
  // *_checked (the v2 engine, arithmetic_function.rs:66-71,147-152,219-224,244-249): the same value, and an integer result
  // that does not fit raises the row's error flag (reported as SQLRS_ERR_ARROW like arrow's "Overflow happened on ...")
  const bool checked = op_in >= SQLRS_OP_ADD_CHECKED;
  const int op = checked ? op_in - (SQLRS_OP_ADD_CHECKED - SQLRS_OP_ADD) : op_in;
  if (!is_numeric(l.dtype)) fail(SQLRS_ERR_UNSUPPORTED, "todo!: unsupported data type");
  if (r.dtype != l.dtype) fail(SQLRS_ERR_INTERNAL, "compute_op failed to downcast array");
  std::string a = vname(l.id), b = vname(r.id);
  bool mn = l.maybe_null || r.maybe_null;
  std::string valid = "(" + nname(l.id) + " && " + nname(r.id) + ")";
  std::string key = "ar" + std::to_string(op_in) + "_" + std::to_string(l.id) + "_" + std::to_string(r.id);
  std::string expr;
  if (checked && l.dtype != SQLRS_DT_FLOAT64 && cse_.find(key) == cse_.end()) {
    g_checked_arithmetic_compiled = true;
    std::string ovf;
    if (l.dtype == SQLRS_DT_INT32) {
      const char* o = op == SQLRS_OP_ADD ? "+" : op == SQLRS_OP_SUB ? "-" : op == SQLRS_OP_MUL ? "*" : nullptr;
      if (o) ovf = "((long long)" + a + " " + o + " (long long)" + b + " != (long long)(int)((long long)" + a + " " + o + " (long long)" + b + "))";
      else ovf = "(" + a + " == (-2147483647 - 1) && " + b + " == -1)";
    } else {
      const std::string mx = "9223372036854775807LL", mn64 = "(-9223372036854775807LL - 1)";
      switch (op) {
        case SQLRS_OP_ADD: ovf = "((" + b + " > 0 && " + a + " > " + mx + " - " + b + ") || (" + b + " < 0 && " + a + " < " + mn64 + " - " + b + "))"; break;
        case SQLRS_OP_SUB: ovf = "((" + b + " < 0 && " + a + " > " + mx + " + " + b + ") || (" + b + " > 0 && " + a + " < " + mn64 + " + " + b + "))"; break;
        case SQLRS_OP_MUL:
          ovf = "(__mul64hi(" + a + ", " + b + ") != ((long long)((unsigned long long)" + a + " * (unsigned long long)" + b + ") >> 63))";
          break;
        default: ovf = "(" + a + " == " + mn64 + " && " + b + " == -1)"; break;
      }
    }
    body_ << "  e" << err_class << " |= (" << valid << " && " << ovf << ");\n";
    err_used_[err_class] = true;
  }
  if (l.dtype == SQLRS_DT_FLOAT64) {
    const char* o = op == SQLRS_OP_ADD ? "+" : op == SQLRS_OP_SUB ? "-" : op == SQLRS_OP_MUL ? "*" : "/";
    // __d*_rn: keep the reference's separate multiply/add roundings (no FMA contraction)
    const char* fn = op == SQLRS_OP_ADD ? "__dadd_rn" : op == SQLRS_OP_SUB ? "__dsub_rn" : op == SQLRS_OP_MUL ? "__dmul_rn" : "__ddiv_rn";
    (void)o;
    expr = std::string(fn) + "(" + a + ", " + b + ")";
    if (op == SQLRS_OP_DIV) {
      if (cse_.find(key) == cse_.end()) {
        body_ << "  e" << err_class << " |= (" << valid << " && " << b << " == 0.0);\n";
        err_used_[err_class] = true;
      }
    }
  } else {
    bool i32 = l.dtype == SQLRS_DT_INT32;
    std::string ut = i32 ? "unsigned" : "unsigned long long";
    std::string st = i32 ? "int" : "long long";
    auto wrap = [&](const char* o) { return "(" + st + ")((" + ut + ")" + a + " " + o + " (" + ut + ")" + b + ")"; };
    switch (op) {
      case SQLRS_OP_ADD: expr = wrap("+"); break;
      case SQLRS_OP_SUB: expr = wrap("-"); break;
      case SQLRS_OP_MUL: expr = wrap("*"); break;
      case SQLRS_OP_DIV:
        if (cse_.find(key) == cse_.end()) {
          body_ << "  e" << err_class << " |= (" << valid << " && " << b << " == 0);\n";
          err_used_[err_class] = true;
        }
        // div_wrapping: MIN / -1 wraps; NULL rows and the (flagged) zero divisor produce 0
        expr = "((!" + valid + " || " + b + " == 0) ? (" + st + ")0 : (" + b + " == -1 ? (" + st + ")((" + ut + ")0 - (" + ut + ")" + a +
               ") : " + a + " / " + b + "))";
        break;
    }
  }
  return define(l.dtype, expr, valid, mn, l.decl_null || r.decl_null, key);
}

// gt_dyn / lt_dyn / gt_eq_dyn / lt_eq_dyn / eq_dyn / neq_dyn, array_compute.rs:80-85
Val RowProgram::comparison(const Val& l, const Val& r, int op) {
  if (l.dtype != r.dtype)
    fail(SQLRS_ERR_ARROW, std::string("Invalid argument error: comparing ") + dtype_name(l.dtype) + " with " + dtype_name(r.dtype));
  if (l.dtype == SQLRS_DT_NULL) fail(SQLRS_ERR_ARROW, "comparison of Null arrays is not supported");
  const char* o = op == SQLRS_OP_GT ? ">" : op == SQLRS_OP_LT ? "<" : op == SQLRS_OP_GE ? ">=" : op == SQLRS_OP_LE ? "<=" : op == SQLRS_OP_EQ ? "==" : "!=";
  std::string key = "cmp" + std::to_string(op) + "_" + std::to_string(l.id) + "_" + std::to_string(r.id);
  std::string a = vname(l.id), b = vname(r.id);
  // Utf8 values are string-pool ids: equal strings have equal ids; the ORDER of two strings is the order of their byte-wise
  // ranks (csrc/jit/strrank.cuh; a NULL row's id is not looked up)
  if (l.dtype == SQLRS_DT_UTF8 && op != SQLRS_OP_EQ && op != SQLRS_OP_NE) {
    a = "sq_str_rank(" + nname(l.id) + " ? " + a + " : 0)";
    b = "sq_str_rank(" + nname(r.id) + " ? " + b + " : 0)";
  }
  if (l.dtype == SQLRS_DT_BOOL) {
    a = "(int)" + a;
    b = "(int)" + b;
  }
  return define(SQLRS_DT_BOOL, "(" + a + " " + o + " " + b + ")", "(" + nname(l.id) + " && " + nname(r.id) + ")",
                l.maybe_null || r.maybe_null, l.decl_null || r.decl_null, key);
}

// boolean_op! + and_kleene / or_kleene, array_compute.rs:48-68,86-87
Val RowProgram::kleene(const Val& l, const Val& r, int op) {
  if (l.dtype != SQLRS_DT_BOOL || r.dtype != SQLRS_DT_BOOL)
    fail(SQLRS_ERR_INTERNAL, std::string("Cannot evaluate binary expression with types ") + dtype_name(l.dtype) + " and " +
                                 dtype_name(r.dtype) + ", only Boolean supported");
  std::string key = "kl" + std::to_string(op) + "_" + std::to_string(l.id) + "_" + std::to_string(r.id);
  std::string a = vname(l.id), b = vname(r.id), la = nname(l.id), lb = nname(r.id);
  bool mn = l.maybe_null || r.maybe_null;
  std::string value, valid;
  if (op == SQLRS_OP_AND) {
    // valid iff both valid, or one side is a valid FALSE
    valid = "((" + la + " && " + lb + ") || (" + la + " && !" + a + ") || (" + lb + " && !" + b + "))";
    value = "((" + la + " ? " + a + " : true) && (" + lb + " ? " + b + " : true) && " + valid + ")";
  } else {
    valid = "((" + la + " && " + lb + ") || (" + la + " && " + a + ") || (" + lb + " && " + b + "))";
    value = "(((" + la + " && " + a + ") || (" + lb + " && " + b + ")))";
  }
  return define(SQLRS_DT_BOOL, value, valid, mn, l.decl_null || r.decl_null, key);
}

// BoundExpr::eval_column over the flattened (postfix) tree, evaluator.rs:13-28
Val RowProgram::compile(const ExprCopy& e, int err_class) {
  if (e.empty()) fail(SQLRS_ERR_INVALID_ARG, "empty expression");
  std::vector<Val> stack;
  for (const ExprNodeCopy& n : e) {
    switch (n.op) {
      case SQLRS_OP_INPUT_REF: stack.push_back(load_column(n.index)); break;
      case SQLRS_OP_CONSTANT: stack.push_back(constant(n)); break;
      case SQLRS_OP_CAST: {
        if (stack.empty()) fail(SQLRS_ERR_INVALID_ARG, "malformed expression");
        Val a = stack.back();
        stack.pop_back();
        stack.push_back(cast(a, n.dtype));
        break;
      }
      default: {
        if (stack.size() < 2) fail(SQLRS_ERR_INVALID_ARG, "malformed expression");
        Val r = stack.back();
        stack.pop_back();
        Val l = stack.back();
        stack.pop_back();
        if (n.op >= SQLRS_OP_ADD && n.op <= SQLRS_OP_DIV_CHECKED) stack.push_back(arithmetic(l, r, n.op, err_class));
        else if (n.op >= SQLRS_OP_GT && n.op <= SQLRS_OP_NE) stack.push_back(comparison(l, r, n.op));
        else if (n.op == SQLRS_OP_AND || n.op == SQLRS_OP_OR) stack.push_back(kleene(l, r, n.op));
        else fail(SQLRS_ERR_UNSUPPORTED, "todo!: unsupported binary operator");
      }
    }
  }
  if (stack.size() != 1) fail(SQLRS_ERR_INVALID_ARG, "malformed expression");
  return stack.back();
}

int RowProgram::emit_raw_bits(const Val& v) {
  int id = fresh();
  std::string e;
  switch (v.dtype) {
    case SQLRS_DT_FLOAT64: e = "(unsigned long long)__double_as_longlong(" + vname(v.id) + ")"; break;
    case SQLRS_DT_BOOL: e = "(unsigned long long)(" + vname(v.id) + " ? 1 : 0)"; break;
    case SQLRS_DT_INT32: e = "(unsigned long long)(long long)" + vname(v.id); break;
    case SQLRS_DT_INT64:
    case SQLRS_DT_UTF8: e = "(unsigned long long)" + vname(v.id); break;  // Utf8: the string pool id identifies the string
    default: e = "0ULL"; break;
  }
  // NULL cells compare by the null mask alone: force their payload to 0
  body_ << "  const unsigned long long " << vname(id) << " = " << nname(v.id) << " ? " << e << " : 0ULL;\n";
  return id;
}

std::string RowProgram::mix_hash_of_bits_source(const std::vector<int>& dtypes) {
  std::ostringstream s;
  s << "  unsigned long long h = 0x9e3779b97f4a7c15ULL;\n";
  for (size_t k = 0; k < dtypes.size(); k++)
    s << "  h = (h ^ kb[" << k << "] ^ 0x" << std::hex << (0x1000193ULL * (unsigned)(dtypes[k] + 1)) << std::dec << "ULL) * 0xff51afd7ed558ccdULL;\n  h ^= h >> 32;\n";
  s << "  return h;\n";
  return s.str();
}

int RowProgram::emit_mix_hash(const std::vector<int>& raw_ids, const std::vector<Val>& keys) {
  int id = fresh();
  body_ << "  unsigned long long " << vname(id) << " = 0x9e3779b97f4a7c15ULL;\n";
  for (size_t k = 0; k < raw_ids.size(); k++) {
    // the key's type is part of the identity (the reference's hash_one differs between Int32 and Int64 too)
    body_ << "  " << vname(id) << " = (" << vname(id) << " ^ " << vname(raw_ids[k]) << " ^ 0x" << std::hex << (0x1000193ULL * (unsigned)(keys[k].dtype + 1)) << std::dec << "ULL"
          << (keys[k].maybe_null ? " ^ (" + nname(keys[k].id) + " ? 0ULL : 0x5bd1e995ULL)" : std::string("")) << ") * 0xff51afd7ed558ccdULL;\n";
    body_ << "  " << vname(id) << " ^= " << vname(id) << " >> 32;\n";
  }
  return id;
}

// create_hashes, hash_utils.rs:161-220, RandomState::with_seeds(0,0,0,0): a single key column is
// hash_one(v); several fold combine_hashes from 0; a NULL cell leaves the running hash untouched
// (quirk K3); a Null-typed column hashes the constant 1 (hash_null, :18-29).
int RowProgram::emit_row_hash(const std::vector<Val>& keys) {
  int id = fresh();
  bool multi = keys.size() > 1;
  body_ << "  unsigned long long " << vname(id) << " = 0ULL;\n";
  for (const Val& k : keys) {
    std::string cell;
    switch (k.dtype) {
      case SQLRS_DT_NULL:
        body_ << "  " << vname(id) << " = " << (multi ? "sq_combine(sq_hash_one(1ULL), " + vname(id) + ")" : std::string("sq_hash_one(1ULL)")) << ";\n";
        continue;
      case SQLRS_DT_INT32: cell = "(unsigned long long)(unsigned)" + vname(k.id); break;
      case SQLRS_DT_INT64: cell = "(unsigned long long)" + vname(k.id); break;
      // Utf8: hash_one over the pool id, not over the bytes as the reference does (hash_utils.rs:199-208) — hashes never leave
      // the operators and ids identify strings, so only the collision pattern of the hash-only identity mode differs (unpinned)
      case SQLRS_DT_UTF8: cell = "(unsigned long long)" + vname(k.id) + " ^ 0x7574663875746638ULL"; break;
      case SQLRS_DT_BOOL: cell = "(unsigned long long)(" + vname(k.id) + " ? 1 : 0)"; break;
      case SQLRS_DT_FLOAT64: cell = "(unsigned long long)__double_as_longlong(" + vname(k.id) + ")"; break;
      default: fail(SQLRS_ERR_INTERNAL, std::string("Unsupported data type in hasher: ") + dtype_name(k.dtype));
    }
    std::string hv = "sq_hash_one(" + cell + ")";
    std::string upd = multi ? "sq_combine(" + hv + ", " + vname(id) + ")" : hv;
    if (k.maybe_null) body_ << "  if (" << nname(k.id) << ") " << vname(id) << " = " << upd << ";\n";
    else body_ << "  " << vname(id) << " = " << upd << ";\n";
  }
  return id;
}

}  // namespace sq
