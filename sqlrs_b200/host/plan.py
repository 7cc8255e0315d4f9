"""Host-side mirror of the reference's physical plan nodes and ExecutorBuilder.

Reference: `PhysicalTableScan`, `PhysicalFilter`, `PhysicalSimpleAgg`, `PhysicalHashAgg`,
`PhysicalHashJoin`, `PhysicalOrder`, `PhysicalProject`, `PhysicalLimit`
(src/optimizer/plan_node/physical_*.rs) and
`ExecutorBuilder::build(plan) -> BoxedExecutor` (src/executor/mod.rs:36-56, visit_* :87-200).
Handing the library the whole sub-plan (sqlrs_plan_* of include/sqlrs_b200.h) keeps tables and
intermediates in HBM and lets it fuse Filter into the aggregate above it.  This file contains no
compute: it flattens the tree into `sqlrs_plan_node` records and moves Arrow data across the ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import pyarrow as pa

from . import ffi
from .executor import JOIN_TYPES, BoundOrderBy, JoinCondition
from .expr import AggArray, BoundExpr, ExprArray, NameArray, project_field_names


class PlanNode:
    def output_schema(self, tables: Dict[int, pa.Schema]) -> pa.Schema:
        raise NotImplementedError


@dataclass
class PhysicalTableScan(PlanNode):
    """PhysicalTableScan (src/optimizer/plan_node/physical_table_scan.rs): batches are pushed per table slot."""
    table_slot: int

    def output_schema(self, tables):
        return tables[self.table_slot]


@dataclass
class PhysicalFilter(PlanNode):
    """PhysicalFilter{expr, input} (physical_filter.rs:8-20)"""
    expr: BoundExpr
    child: PlanNode

    def output_schema(self, tables):
        return self.child.output_schema(tables)


@dataclass
class PhysicalSimpleAgg(PlanNode):
    """PhysicalSimpleAgg{agg_funcs, input} (physical_simple_agg.rs)"""
    agg_funcs: Sequence[BoundExpr]
    child: PlanNode

    def output_schema(self, tables):
        s = self.child.output_schema(tables)
        return pa.schema([a.eval_field(s) for a in self.agg_funcs])


@dataclass
class PhysicalHashAgg(PlanNode):
    """PhysicalHashAgg{agg_funcs, group_by, input} (physical_hash_agg.rs:8-20)"""
    agg_funcs: Sequence[BoundExpr]
    group_by: Sequence[BoundExpr]
    child: PlanNode

    def output_schema(self, tables):
        s = self.child.output_schema(tables)
        return pa.schema([g.eval_field(s) for g in self.group_by] + [a.eval_field(s) for a in self.agg_funcs])


@dataclass
class PhysicalHashJoin(PlanNode):
    """PhysicalHashJoin{left, right, join_type, join_condition, join_output_columns} (physical_hash_join.rs:9-41)"""
    left: PlanNode
    right: PlanNode
    join_type: str
    join_condition: JoinCondition
    join_output_schema: pa.Schema

    def output_schema(self, tables):
        return self.join_output_schema


@dataclass
class PhysicalCrossJoin(PlanNode):
    """PhysicalCrossJoin{left, right, join_output_columns} (physical_cross_join.rs)"""
    left: PlanNode
    right: PlanNode
    join_output_schema: pa.Schema

    def output_schema(self, tables):
        return self.join_output_schema


@dataclass
class PhysicalProject(PlanNode):
    """PhysicalProject{exprs, input} (physical_project.rs)"""
    exprs: Sequence[BoundExpr]
    child: PlanNode

    def output_schema(self, tables):
        s = self.child.output_schema(tables)
        return pa.schema([e.eval_field(s) for e in self.exprs])


@dataclass
class PhysicalOrder(PlanNode):
    """PhysicalOrder{order_by, input} (physical_order.rs)"""
    order_by: Sequence[BoundOrderBy]
    child: PlanNode

    def output_schema(self, tables):
        return self.child.output_schema(tables)


@dataclass
class PhysicalLimit(PlanNode):
    """PhysicalLimit{limit, offset, input} (physical_limit.rs); None = not given"""
    limit: Optional[int]
    offset: Optional[int]
    child: PlanNode

    def output_schema(self, tables):
        return self.child.output_schema(tables)


class GpuPlan:
    """A built plan: push tables, execute, collect (try_collect, src/executor/mod.rs:58-64)."""

    def __init__(self, lib: ffi.Library, root: PlanNode, table_schemas: Dict[int, pa.Schema], options=None):
        self.lib = lib
        self.options = options if options is not None else lib.options()
        self._keep: list = []
        nodes: List[ffi.PlanNode] = []

        def add(node: PlanNode) -> int:
            rec = ffi.PlanNode()
            rec.child0 = rec.child1 = -1
            rec.limit = rec.offset = -1
            if isinstance(node, PhysicalTableScan):
                rec.kind, rec.table_slot = ffi.NODE_SCAN, node.table_slot
            elif isinstance(node, PhysicalFilter):
                rec.kind = ffi.NODE_FILTER
                rec.child0 = add(node.child)
                flat = node.expr.flatten()
                self._keep.append(flat)
                rec.predicate = flat.c
            elif isinstance(node, (PhysicalSimpleAgg, PhysicalHashAgg)):
                child_schema = node.child.output_schema(table_schemas)
                rec.child0 = add(node.child)
                aggs = AggArray(node.agg_funcs, child_schema)
                self._keep.append(aggs)
                rec.aggs, rec.n_aggs = aggs.ptr, aggs.n
                if isinstance(node, PhysicalHashAgg):
                    rec.kind = ffi.NODE_HASH_AGG
                    groups = ExprArray(node.group_by)
                    names = NameArray([g.eval_field(child_schema).name for g in node.group_by])
                    self._keep += [groups, names]
                    rec.group_by, rec.group_names, rec.n_group_by = groups.ptr, names.ptr, groups.n
                else:
                    rec.kind = ffi.NODE_SIMPLE_AGG
            elif isinstance(node, PhysicalHashJoin):
                if node.join_type not in JOIN_TYPES:
                    raise ffi.ExecutorError(ffi.ERR_INTERNAL, "Cross join should not be in HashJoinExecutor")
                rec.kind = ffi.NODE_HASH_JOIN
                rec.child0 = add(node.left)
                rec.child1 = add(node.right)
                rec.join_type = JOIN_TYPES[node.join_type]
                lk = ExprArray([l for l, _ in node.join_condition.on])
                rk = ExprArray([r for _, r in node.join_condition.on])
                self._keep += [lk, rk]
                rec.left_keys, rec.right_keys, rec.n_keys = lk.ptr, rk.ptr, lk.n
                if node.join_condition.filter is not None:
                    flat = node.join_condition.filter.flatten()
                    self._keep.append(flat)
                    rec.predicate = flat.c
                sch = ffi.export_schema(node.join_output_schema)
                self._keep.append(sch)
                rec.join_output_schema = C.pointer(sch)
            elif isinstance(node, PhysicalCrossJoin):
                rec.kind = ffi.NODE_CROSS_JOIN
                rec.child0 = add(node.left)
                rec.child1 = add(node.right)
                sch = ffi.export_schema(node.join_output_schema)
                self._keep.append(sch)
                rec.join_output_schema = C.pointer(sch)
            elif isinstance(node, PhysicalProject):
                child_schema = node.child.output_schema(table_schemas)
                rec.kind = ffi.NODE_PROJECT
                rec.child0 = add(node.child)
                exprs = ExprArray(node.exprs)
                names = NameArray(project_field_names(node.exprs, child_schema))
                self._keep += [exprs, names]
                rec.exprs, rec.expr_names, rec.n_exprs = exprs.ptr, names.ptr, exprs.n
            elif isinstance(node, PhysicalOrder):
                rec.kind = ffi.NODE_ORDER
                rec.child0 = add(node.child)
                exprs = ExprArray([o.expr for o in node.order_by])
                asc = (C.c_int32 * max(1, len(node.order_by)))(*[int(o.asc) for o in node.order_by])
                self._keep += [exprs, asc]
                rec.exprs, rec.n_exprs = exprs.ptr, exprs.n
                rec.order_asc = C.cast(asc, C.POINTER(C.c_int32))
            elif isinstance(node, PhysicalLimit):
                rec.kind = ffi.NODE_LIMIT
                rec.child0 = add(node.child)
                rec.limit = -1 if node.limit is None else int(node.limit)
                rec.offset = -1 if node.offset is None else int(node.offset)
            else:
                raise TypeError(f"unknown plan node {node!r}")
            nodes.append(rec)
            return len(nodes) - 1

        root_idx = add(root)
        arr = (ffi.PlanNode * len(nodes))(*nodes)
        self._keep.append(arr)
        self.handle = C.c_void_p()
        try:
            lib.check(lib.plan_create(arr, len(nodes), root_idx, C.byref(self.options), C.byref(self.handle)))
        finally:
            for k in self._keep:
                if isinstance(k, ffi.ArrowSchema):
                    ffi.release_schema(k)

    def push_table(self, table_slot: int, batch: pa.RecordBatch):
        arr, sch = ffi.export_batch(batch)
        try:
            self.lib.check(self.lib.plan_push_table(self.handle, table_slot, C.byref(arr), C.byref(sch)))
        finally:
            ffi.release_schema(sch)

    def push_table_batched(self, table_slot: int, batch: pa.RecordBatch, batch_rows: int = 1024):
        """ONE host batch that the plan scans in `batch_rows`-row slices (the reference's scan batch size, csv.rs:105):
        the slicing happens inside the library, not in a Python loop."""
        arr, sch = ffi.export_batch(batch)
        try:
            self.lib.check(self.lib.plan_push_table_batched(self.handle, table_slot, C.byref(arr), C.byref(sch), batch_rows))
        finally:
            ffi.release_schema(sch)

    def result_shape(self):
        """(rows, columns) of the next pending DEVICE result, or None."""
        n, c, has = C.c_int64(0), C.c_int32(0), C.c_int32(0)
        self.lib.check(self.lib.plan_result_shape(self.handle, C.byref(n), C.byref(c), C.byref(has)))
        return (int(n.value), int(c.value)) if has.value else None

    def next_to_device(self, column_ptrs: Sequence[int]):
        """copies the next pending result's columns into device buffers (raw pointers), device to device"""
        arr = (C.c_void_p * max(1, len(column_ptrs)))(*column_ptrs)
        self.lib.check(self.lib.plan_next_to_device(self.handle, arr, len(column_ptrs)))

    def kernel_events(self) -> dict:
        """{kernel: {ms, launches}} recorded since the last call (plans built with FLAG_KERNEL_EVENTS)"""
        import json

        out = C.c_void_p()
        self.lib.check(self.lib.kernel_events_collect(C.byref(out)))
        try:
            return json.loads(C.string_at(out).decode()) if out.value else {}
        finally:
            self.lib.free(out)

    def push_table_device(self, table_slot: int, device_batch):
        """`device_batch`: tpch.DeviceTable (columns resident in HBM); zero copy."""
        darr, sch = device_batch.export()
        try:
            self.lib.check(self.lib.plan_push_table_device(self.handle, table_slot, C.byref(darr), C.byref(sch)))
        finally:
            ffi.release_schema(sch)

    def push_table_resident(self, table_slot: int, table):
        """`table`: storage.InMemoryTable — its batches already live in HBM (zero copy, no PCIe traffic)."""
        self.lib.check(self.lib.plan_push_table_resident(self.handle, table_slot, table.handle))

    def execute(self):
        self.lib.check(self.lib.plan_execute(self.handle))

    def collect(self) -> List[pa.RecordBatch]:
        out = []
        has = C.c_int32(0)
        while True:
            arr, sch = ffi.ArrowArray(), ffi.ArrowSchema()
            self.lib.check(self.lib.plan_next(self.handle, C.byref(arr), C.byref(sch), C.byref(has)))
            if not has.value:
                return out
            out.append(ffi.import_batch(arr, sch))

    def run(self) -> List[pa.RecordBatch]:
        self.execute()
        return self.collect()

    def reset(self):
        self.lib.check(self.lib.plan_reset(self.handle))

    def clear_table(self, table_slot: int):
        """forget the batches of one table slot; the other slots keep theirs (re-run with one input replaced)"""
        self.lib.check(self.lib.plan_clear_table(self.handle, table_slot))

    def scan_kernel_ms(self):
        """(device ms, launches) of the dominant scan kernel(s) in the last execute (needs FLAG_TIMING)."""
        n = C.c_int64(0)
        ms = self.lib.plan_scan_kernel_ms(self.handle, C.byref(n))
        return float(ms), int(n.value)

    def describe(self) -> str:
        s = self.lib.plan_describe(self.handle)
        return s.decode() if s else ""

    def close(self):
        if self.handle:
            self.lib.plan_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ExecutorBuilder:
    """ExecutorBuilder::new(storage).build(plan) (src/executor/mod.rs:36-56)"""

    def __init__(self, lib: Optional[ffi.Library] = None, options=None):
        self.lib = lib if lib is not None else ffi.load()
        self.options = options

    def build(self, plan: PlanNode, table_schemas: Dict[int, pa.Schema]) -> GpuPlan:
        return GpuPlan(self.lib, plan, table_schemas, self.options)
