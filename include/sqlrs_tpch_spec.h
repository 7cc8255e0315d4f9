/*
 * Synthetic TPC-H-shaped tables — the row functions.
 *
 * Counter-based: every cell is a pure function of (seed, column id, row), so a table
 * of any scale factor can be produced in place in HBM (CUDA generator), in host RAM
 * (CPU generator used by the oracle / CPU baseline) or shard by shard on N GPUs and
 * is bit-identical everywhere.  This header is the single definition of those
 * functions; it is plain C99/C++/CUDA and has no dependency on either library.
 *
 * The reference (sqlrs v1) has no Date/Decimal types (src/types/mod.rs:23-36), so dates
 * are int64 days since 1970-01-01 and flags are int64 codes.
 *
 *  customer  c_custkey = row+1, c_mktsegment = u mod 5              (1 = BUILDING)
 *  orders    o_orderkey = row+1 (unique, ascending), o_custkey uniform over customers with
 *            custkey mod 3 != 0 (TPC-H: a third of customers place no orders),
 *            o_orderdate uniform in [8035 (1992-01-01), 10440 (1998-08-02)], o_shippriority = 0
 *  lineitem  clustered on l_orderkey; every block of 7 consecutive orders owns 28 rows, order j of
 *            block b has 1 + ((j + s_b) mod 7) lines (1..7 lines per order, 4.0 on average), so
 *            row -> order is closed-form;  l_quantity in 1..50, price(part) in [900.00, 2100.00],
 *            l_extendedprice = quantity * price (2 decimals), l_discount in 0.00..0.10,
 *            l_tax in 0.00..0.08, l_shipdate = o_orderdate + 1 + u mod 121,
 *            l_returnflag / l_linestatus per flags_mode, l_quantity_i64 = (int64) l_quantity.
 */
#ifndef SQLRS_TPCH_SPEC_H
#define SQLRS_TPCH_SPEC_H

#include <stdint.h>

#if defined(__CUDACC__)
#define SQLRS_HD __host__ __device__ __forceinline__
#else
#define SQLRS_HD static inline
#endif

#define SQLRS_TPCH_SEED 0x51512025ULL

/* column ids (the `col` argument of sqlrs_tpch_u) == column positions in each table */
enum { SQLRS_C_CUSTKEY = 0, SQLRS_C_MKTSEGMENT = 1, SQLRS_CUSTOMER_NCOLS = 2 };
enum { SQLRS_O_ORDERKEY = 0, SQLRS_O_CUSTKEY = 1, SQLRS_O_ORDERDATE = 2, SQLRS_O_SHIPPRIORITY = 3, SQLRS_ORDERS_NCOLS = 4 };
enum {
  SQLRS_L_ORDERKEY = 0,
  SQLRS_L_QUANTITY = 1,
  SQLRS_L_EXTENDEDPRICE = 2,
  SQLRS_L_DISCOUNT = 3,
  SQLRS_L_TAX = 4,
  SQLRS_L_RETURNFLAG = 5,
  SQLRS_L_LINESTATUS = 6,
  SQLRS_L_SHIPDATE = 7,
  SQLRS_L_QUANTITY_I64 = 8,
  SQLRS_LINEITEM_NCOLS = 9
};

SQLRS_HD uint64_t sqlrs_splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
/* table: 0 customer, 1 orders, 2 lineitem; salt separates independent draws of one cell */
SQLRS_HD uint64_t sqlrs_tpch_u(int table, int col, int salt, int64_t row) {
  uint64_t tag = ((uint64_t)table << 60) ^ ((uint64_t)col << 52) ^ ((uint64_t)salt << 48);
  return sqlrs_splitmix64(SQLRS_TPCH_SEED ^ tag ^ (uint64_t)row);
}

SQLRS_HD int64_t sqlrs_tpch_lineitem_rows(int64_t n_orders) {
  int64_t full = n_orders / 7, rem = n_orders % 7;
  int64_t rows = full * 28;
  if (rem) {
    int s = (int)(sqlrs_tpch_u(2, 0, 1, full) % 7);
    for (int j = 0; j < (int)rem; j++) rows += 1 + ((j + s) % 7);
  }
  return rows;
}

/* ---- customer ---- */
SQLRS_HD int64_t sqlrs_c_custkey(int64_t row) { return row + 1; }
SQLRS_HD int64_t sqlrs_c_mktsegment(int64_t row) { return (int64_t)(sqlrs_tpch_u(0, SQLRS_C_MKTSEGMENT, 0, row) % 5); }

/* ---- orders ---- */
SQLRS_HD int64_t sqlrs_o_orderkey(int64_t row) { return row + 1; }
SQLRS_HD int64_t sqlrs_o_custkey(int64_t row, int64_t n_customer) {
  int64_t eligible = n_customer - n_customer / 3; /* custkeys in 1..n with key mod 3 != 0 */
  if (eligible <= 0) return 1;
  int64_t k = (int64_t)(sqlrs_tpch_u(1, SQLRS_O_CUSTKEY, 0, row) % (uint64_t)eligible);
  return 3 * (k / 2) + 1 + (k % 2);
}
SQLRS_HD int64_t sqlrs_o_orderdate(int64_t row) { return 8035 + (int64_t)(sqlrs_tpch_u(1, SQLRS_O_ORDERDATE, 0, row) % 2406); }
SQLRS_HD int64_t sqlrs_o_shippriority(int64_t row) {
  (void)row;
  return 0;
}

/* ---- lineitem ---- */
/* row -> 0-based index of the owning order */
SQLRS_HD int64_t sqlrs_l_order_index(int64_t row) {
  int64_t b = row / 28;
  int k = (int)(row % 28);
  int s = (int)(sqlrs_tpch_u(2, 0, 1, b) % 7);
  int j = 0;
  for (; j < 6; j++) {
    int lines = 1 + ((j + s) % 7);
    if (k < lines) break;
    k -= lines;
  }
  return 7 * b + j;
}
SQLRS_HD int64_t sqlrs_l_orderkey(int64_t row) { return sqlrs_o_orderkey(sqlrs_l_order_index(row)); }
SQLRS_HD int64_t sqlrs_l_quantity_i64(int64_t row) { return 1 + (int64_t)(sqlrs_tpch_u(2, SQLRS_L_QUANTITY, 0, row) % 50); }
SQLRS_HD double sqlrs_l_quantity(int64_t row) { return (double)sqlrs_l_quantity_i64(row); }
SQLRS_HD double sqlrs_l_extendedprice(int64_t row) {
  int64_t cents = 90000 + (int64_t)(sqlrs_tpch_u(2, SQLRS_L_EXTENDEDPRICE, 0, row) % 120001);
  return (double)(sqlrs_l_quantity_i64(row) * cents) / 100.0;
}
SQLRS_HD double sqlrs_l_discount(int64_t row) { return (double)(sqlrs_tpch_u(2, SQLRS_L_DISCOUNT, 0, row) % 11) / 100.0; }
SQLRS_HD double sqlrs_l_tax(int64_t row) { return (double)(sqlrs_tpch_u(2, SQLRS_L_TAX, 0, row) % 9) / 100.0; }
SQLRS_HD int64_t sqlrs_l_shipdate(int64_t row) {
  return sqlrs_o_orderdate(sqlrs_l_order_index(row)) + 1 + (int64_t)(sqlrs_tpch_u(2, SQLRS_L_SHIPDATE, 0, row) % 121);
}
/* flags_mode 0 (8 groups): returnflag = u mod 4, linestatus = (u >> 2) mod 2
 * flags_mode 1 (TPC-H spec): receipt = ship + 1 + u mod 30; receipt <= 9298 (1995-06-17) ? (R=2 | A=0) : N=1;
 *                            linestatus = ship > 9298 ? O=1 : F=0      (4 populated groups) */
SQLRS_HD int64_t sqlrs_l_returnflag(int64_t row, int flags_mode) {
  uint64_t u = sqlrs_tpch_u(2, SQLRS_L_RETURNFLAG, 0, row);
  if (flags_mode == 0) return (int64_t)(u % 4);
  int64_t receipt = sqlrs_l_shipdate(row) + 1 + (int64_t)((u >> 8) % 30);
  if (receipt <= 9298) return (u & 1) ? 2 : 0;
  return 1;
}
SQLRS_HD int64_t sqlrs_l_linestatus(int64_t row, int flags_mode) {
  if (flags_mode == 0) return (int64_t)((sqlrs_tpch_u(2, SQLRS_L_RETURNFLAG, 0, row) >> 2) % 2);
  return sqlrs_l_shipdate(row) > 9298 ? 1 : 0;
}

/* one cell as raw 8 bytes (int64 value, or the bit pattern of the double) */
SQLRS_HD uint64_t sqlrs_tpch_cell(int table, int col, int64_t row, int64_t n_customer, int flags_mode) {
  union {
    double d;
    int64_t i;
    uint64_t u;
  } v;
  v.u = 0;
  if (table == 0) {
    v.i = col == SQLRS_C_CUSTKEY ? sqlrs_c_custkey(row) : sqlrs_c_mktsegment(row);
  } else if (table == 1) {
    switch (col) {
      case SQLRS_O_ORDERKEY: v.i = sqlrs_o_orderkey(row); break;
      case SQLRS_O_CUSTKEY: v.i = sqlrs_o_custkey(row, n_customer); break;
      case SQLRS_O_ORDERDATE: v.i = sqlrs_o_orderdate(row); break;
      default: v.i = sqlrs_o_shippriority(row); break;
    }
  } else {
    switch (col) {
      case SQLRS_L_ORDERKEY: v.i = sqlrs_l_orderkey(row); break;
      case SQLRS_L_QUANTITY: v.d = sqlrs_l_quantity(row); break;
      case SQLRS_L_EXTENDEDPRICE: v.d = sqlrs_l_extendedprice(row); break;
      case SQLRS_L_DISCOUNT: v.d = sqlrs_l_discount(row); break;
      case SQLRS_L_TAX: v.d = sqlrs_l_tax(row); break;
      case SQLRS_L_RETURNFLAG: v.i = sqlrs_l_returnflag(row, flags_mode); break;
      case SQLRS_L_LINESTATUS: v.i = sqlrs_l_linestatus(row, flags_mode); break;
      case SQLRS_L_SHIPDATE: v.i = sqlrs_l_shipdate(row); break;
      default: v.i = sqlrs_l_quantity_i64(row); break;
    }
  }
  return v.u;
}

#endif /* SQLRS_TPCH_SPEC_H */
