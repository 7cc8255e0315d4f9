"""Ad-hoc timing of Q1'-shaped plan variants on one GPU (device-resident table): prints scan-kernel ms."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sqlrs_b200.host import ffi, tpch
from sqlrs_b200.host.expr import AggFunc, BinaryOp, Constant, InputRef, bind_binary_op
from sqlrs_b200.host.plan import ExecutorBuilder, PhysicalFilter, PhysicalHashAgg, PhysicalSimpleAgg, PhysicalTableScan

sf = float(sys.argv[1]) if len(sys.argv) > 1 else 10
which = sys.argv[2].split(",") if len(sys.argv) > 2 else None
lib = ffi.load()
d = tpch.dims(sf)
table = tpch.device_table(lib, d, tpch.LINEITEM, columns=tpch.Q1_COLUMNS)
n = table.n_rows
plan, schemas = tpch.q1_plan()
s = schemas[0]
col = {f.name: InputRef(i, ffi.dtype_of(f.type)) for i, f in enumerate(s)}
pred = plan.child.expr
variants = {
    "q1": plan,
    "q1_nofilter": PhysicalHashAgg(plan.agg_funcs, plan.group_by, PhysicalTableScan(0)),
    "simple_same_aggs": PhysicalSimpleAgg(plan.agg_funcs, PhysicalFilter(pred, PhysicalTableScan(0))),
    "one_key": PhysicalHashAgg(plan.agg_funcs, [col["l_returnflag"]], PhysicalFilter(pred, PhysicalTableScan(0))),
    "count_only": PhysicalHashAgg([AggFunc("Count", [col["l_quantity"]])], plan.group_by, PhysicalFilter(pred, PhysicalTableScan(0))),
    "by_qty_50g": PhysicalHashAgg(plan.agg_funcs, [col["l_quantity_i64"]], PhysicalFilter(pred, PhysicalTableScan(0))),
    "by_shipdate_2500g": PhysicalHashAgg(plan.agg_funcs, [col["l_shipdate"]], PhysicalFilter(pred, PhysicalTableScan(0))),
    "by_qty_ship_125kg": PhysicalHashAgg(plan.agg_funcs, [col["l_quantity_i64"], col["l_shipdate"]], PhysicalFilter(pred, PhysicalTableScan(0))),
    "sum1": PhysicalHashAgg([AggFunc("Sum", [col["l_quantity"]])], plan.group_by, PhysicalFilter(pred, PhysicalTableScan(0))),
}
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    for name, root in variants.items():
        if which and name not in which:
            continue
        for mm in (ffi.MATCH_HASH_AND_KEY, ffi.MATCH_HASH_ONLY):
            opts = lib.options(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=mm, flags=ffi.FLAG_TIMING, stream=C.c_void_p(stream.cuda_stream))
            p = ExecutorBuilder(lib, opts).build(root, schemas)
            best = 1e9
            for it in range(4):
                p.push_table_device(0, table)
                p.execute(); p.collect()
                ms, nl = p.scan_kernel_ms()
                p.reset()
                if it: best = min(best, ms)
            print(f"{name:18s} match_mode={mm} kernel_ms={best:8.3f}  GB/s(64B/row)={n*64/best/1e6:8.1f}  {p.describe()[:100]}", flush=True)
            p.close()
