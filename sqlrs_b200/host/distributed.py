"""Multi-GPU execution of a plan whose root is an aggregate.

One process per GPU (torch.distributed; NCCL over NVLink on the GPU box, gloo in the CPU tests).
Scan, filter and the partial aggregation stay GPU-local; the only exchange is the group-by's:

  1. every rank aggregates its contiguous row shard           sqlrs_plan_execute_partial
  (few groups, NCCL: the packed partial tables are all-gathered device-to-device and rank 0 folds them —
   sqlrs_plan_export_partials_device / sqlrs_plan_merge_partials_device; otherwise:)
  2. partial groups are radix-partitioned by their identity hash (owner = hash mod world) and
     exchanged all-to-all; the owner folds them               sqlrs_plan_merge_partials
  3. the (now disjoint) owner-merged groups are gathered on rank 0, which finalises them in the
     reference's first-appearance order (minimum global row id) sqlrs_plan_finish_partial

Joins (Q3'): `broadcast_build_join_aggregate` (any sharding of the probe sides, build sides exchanged) and
`copartitioned_topk` (fact tables range-partitioned on the join key that is also a group key, dimension table
replicated: every rank runs the WHOLE query on its shards, no data-path collective at all, only the final
LIMIT rows are gathered).

The reference has no distributed execution; results equal the single-process ones (integers
bit-exact, float sums up to summation order).  This module contains no compute — partial groups
cross the C ABI as opaque Arrow batches whose column 0 is the partitioning hash.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import pyarrow as pa

from . import ffi


def _to_bytes(batch: pa.RecordBatch) -> bytes:
    sink = pa.BufferOutputStream()
    with pa.ipc.new_stream(sink, batch.schema) as w:
        w.write_batch(batch)
    return sink.getvalue().to_pybytes()


def _from_bytes(data: bytes) -> pa.RecordBatch:
    with pa.ipc.open_stream(pa.py_buffer(data)) as r:
        batches = [b for b in r]
    if len(batches) == 1:
        return batches[0]
    return pa.Table.from_batches(batches).combine_chunks().to_batches()[0]


class TorchGroup:
    """Byte-level all-to-all / gather on top of torch.distributed (works with nccl and gloo)."""

    def __init__(self, dist, device):
        import torch

        self.torch, self.dist, self.device = torch, dist, device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.native_a2a = dist.get_backend() == "nccl"

    def _tensor(self, data: bytes):
        t = self.torch.frombuffer(bytearray(data), dtype=self.torch.uint8) if data else self.torch.empty(0, dtype=self.torch.uint8)
        return t.to(self.device)

    def all_to_all_bytes(self, payloads: List[bytes]) -> List[bytes]:
        torch, dist, W = self.torch, self.dist, self.world
        send_sizes = torch.tensor([len(p) for p in payloads], dtype=torch.int64, device=self.device)
        if self.native_a2a:
            recv_sizes = torch.empty(W, dtype=torch.int64, device=self.device)
            dist.all_to_all_single(recv_sizes, send_sizes)
            rs = recv_sizes.tolist()
            send = self._tensor(b"".join(payloads))
            recv = torch.empty(sum(rs), dtype=torch.uint8, device=self.device)
            dist.all_to_all_single(recv, send, output_split_sizes=rs, input_split_sizes=[len(p) for p in payloads])
            data = recv.cpu().numpy().tobytes()
            out, off = [], 0
            for s in rs:
                out.append(data[off:off + s])
                off += s
            return out
        # gloo has no all_to_all: all-gather the size matrix and the (padded) send buffers, slice locally
        sizes = [torch.empty(W, dtype=torch.int64, device=self.device) for _ in range(W)]
        dist.all_gather(sizes, send_sizes)
        matrix = [s.tolist() for s in sizes]  # matrix[src][dst]
        max_total = max(sum(row) for row in matrix)
        send = torch.zeros(max(max_total, 1), dtype=torch.uint8, device=self.device)
        joined = b"".join(payloads)
        if joined:
            send[:len(joined)] = self._tensor(joined)
        bufs = [torch.empty_like(send) for _ in range(W)]
        dist.all_gather(bufs, send)
        out = []
        for src in range(W):
            off = sum(matrix[src][:self.rank])
            out.append(bufs[src][off:off + matrix[src][self.rank]].cpu().numpy().tobytes())
        return out

    def all_gather_bytes(self, payload: bytes) -> List[bytes]:
        torch, dist, W = self.torch, self.dist, self.world
        size = torch.tensor([len(payload)], dtype=torch.int64, device=self.device)
        sizes = [torch.empty(1, dtype=torch.int64, device=self.device) for _ in range(W)]
        dist.all_gather(sizes, size)
        sz = [int(s.item()) for s in sizes]
        send = torch.zeros(max(max(sz), 1), dtype=torch.uint8, device=self.device)
        if payload:
            send[:len(payload)] = self._tensor(payload)
        bufs = [torch.empty_like(send) for _ in range(W)]
        dist.all_gather(bufs, send)
        return [bufs[r][:sz[r]].cpu().numpy().tobytes() for r in range(W)]

    def gather_small(self, payload: bytes, dst: int = 0, cap: int = 8192) -> Optional[List[bytes]]:
        """gather_bytes for payloads known to be small (a LIMIT's worth of rows): ONE fixed-size all-gather
        [u32 length | payload | padding] instead of a size exchange + a data exchange."""
        import struct

        torch, dist, W = self.torch, self.dist, self.world
        if len(payload) + 4 > cap:
            # rare: tell every rank to take the two-phase path (length prefix 0xffffffff)
            framed = struct.pack("<I", 0xFFFFFFFF)
        else:
            framed = struct.pack("<I", len(payload)) + payload
        send = torch.zeros(cap, dtype=torch.uint8)
        send[:len(framed)] = torch.frombuffer(bytearray(framed), dtype=torch.uint8)
        send = send.to(self.device, non_blocking=True)
        recv = torch.empty(cap * W, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(recv, send) if self.native_a2a else dist.all_gather(list(recv.view(W, cap).unbind(0)), send)
        data = recv.cpu().numpy().tobytes()
        lens = [struct.unpack_from("<I", data, r * cap)[0] for r in range(W)]
        if any(n == 0xFFFFFFFF for n in lens):
            return self.gather_bytes(payload, dst)
        if self.rank != dst:
            return None
        return [data[r * cap + 4:r * cap + 4 + lens[r]] for r in range(W)]

    def gather_bytes(self, payload: bytes, dst: int = 0) -> Optional[List[bytes]]:
        torch, dist, W = self.torch, self.dist, self.world
        size = torch.tensor([len(payload)], dtype=torch.int64, device=self.device)
        sizes = [torch.empty(1, dtype=torch.int64, device=self.device) for _ in range(W)]
        dist.all_gather(sizes, size)
        sz = [int(s.item()) for s in sizes]
        send = torch.zeros(max(max(sz), 1), dtype=torch.uint8, device=self.device)
        if payload:
            send[:len(payload)] = self._tensor(payload)
        bufs = [torch.empty_like(send) for _ in range(W)]
        dist.all_gather(bufs, send)
        if self.rank != dst:
            return None
        return [bufs[r][:sz[r]].cpu().numpy().tobytes() for r in range(W)]


def all_gather_batches(group: "TorchGroup", batches: List[pa.RecordBatch], schema: pa.Schema) -> List[pa.RecordBatch]:
    """Every rank receives every rank's batches, in rank order (the build-side broadcast of a join)."""
    table = pa.Table.from_batches(batches, schema=schema).combine_chunks()
    payload = _to_bytes(table.to_batches()[0]) if table.num_rows else _to_bytes(pa.RecordBatch.from_pylist([], schema=schema))
    out = []
    for data in group.all_gather_bytes(payload):
        b = _from_bytes(data)
        if b.num_rows:
            out.append(b)
    return out


def _export_partials(plan) -> pa.RecordBatch:
    arr, sch = ffi.ArrowArray(), ffi.ArrowSchema()
    plan.lib.check(plan.lib.plan_export_partials(plan.handle, C.byref(arr), C.byref(sch)))
    return ffi.import_batch(arr, sch)


def _merge_partials(plan, batch: pa.RecordBatch):
    arr, sch = ffi.export_batch(batch)
    try:
        plan.lib.check(plan.lib.plan_merge_partials(plan.handle, C.byref(arr), C.byref(sch)))
    finally:
        ffi.release_schema(sch)


def partition_by_owner(partials: pa.RecordBatch, world: int) -> List[pa.RecordBatch]:
    """Radix partition on the identity hash (column 0, the u64 row hash carried as int64)."""
    h = np.asarray(partials.column(0).to_numpy(zero_copy_only=False)).view(np.uint64)
    owner = (h % np.uint64(world)).astype(np.int64)
    return [partials.filter(pa.array(owner == r)) for r in range(world)]


def _plan_stream(plan) -> int:
    return int(getattr(plan.options, "stream", None) or 0)


def _device_exchange(plan, group: TorchGroup, cap_rows: int):
    """Fast path for few groups (NCCL): partial groups never leave HBM.  Every rank packs its groups into a
    fixed-size device buffer, ONE all-gather over NVLink hands all of them to every rank, rank 0 folds them.
    Returns None when some rank has more than `cap_rows` groups (the caller takes the radix all-to-all path)."""
    torch, lib = group.torch, plan.lib
    words = C.c_int32(0)
    lib.check(lib.plan_partials_row_words(plan.handle, C.byref(words)))
    key = (words.value, cap_rows)
    bufs = getattr(group, "_bufs", {})
    if key not in bufs:
        n = (cap_rows + 1) * words.value
        bufs[key] = (torch.empty(n, dtype=torch.int64, device=group.device), torch.empty(n * group.world, dtype=torch.int64, device=group.device))
        group._bufs = bufs
    send, recv = bufs[key]
    # stream contract (include/sqlrs_b200.h): a plan on the caller's stream packs in stream order with the collective;
    # a plan that owns its stream synchronises inside export, and needs the gathered buffer complete before the merge
    shared_stream = _plan_stream(plan) == torch.cuda.current_stream(group.device).cuda_stream
    if _plan_stream(plan) and not shared_stream:
        raise ValueError("sharded_aggregate: options.stream must be torch's current stream (or NULL)")
    lib.check(lib.plan_export_partials_device(plan.handle, C.c_void_p(send.data_ptr()), cap_rows))
    group.dist.all_gather_into_tensor(recv, send)
    if not shared_stream:
        torch.cuda.current_stream(group.device).synchronize()
    counts = recv.view(group.world, cap_rows + 1, words.value)[:, 0, 0]
    if int(counts.max().item()) > cap_rows:
        return None
    lib.check(lib.plan_clear_partials(plan.handle))
    if group.rank == 0:
        lib.check(lib.plan_merge_partials_device(plan.handle, C.c_void_p(recv.data_ptr()), group.world, cap_rows))
    lib.check(lib.plan_finish_partial(plan.handle))
    result = plan.collect()
    return result if group.rank == 0 else []


def sharded_aggregate(plan, group: TorchGroup, row_base: int = 0, device_cap_rows: int = 256) -> List[pa.RecordBatch]:
    """Runs `plan` (root = aggregate) over the shard pushed on this rank; rank 0 returns the final batches."""
    lib = plan.lib
    lib.check(lib.plan_execute_partial(plan.handle, row_base))
    if group.native_a2a and device_cap_rows > 0 and lib.prefix == "sqlrs_":
        result = _device_exchange(plan, group, device_cap_rows)
        if result is not None:
            return result
    local = _export_partials(plan)
    received = group.all_to_all_bytes([_to_bytes(p) for p in partition_by_owner(local, group.world)])
    lib.check(lib.plan_clear_partials(plan.handle))
    for data in received:
        _merge_partials(plan, _from_bytes(data))
    owned = _export_partials(plan)
    gathered = group.gather_bytes(_to_bytes(owned), dst=0)
    lib.check(lib.plan_clear_partials(plan.handle))
    if gathered is not None:
        for data in gathered:
            _merge_partials(plan, _from_bytes(data))
    lib.check(lib.plan_finish_partial(plan.handle))
    result = plan.collect()
    return result if group.rank == 0 else []


def broadcast_build_join_aggregate(builder, group: TorchGroup, stage1, stage1_schemas, stage1_tables, stage2, stage2_schemas, stage2_tables,
                                   build_slot: int) -> List[pa.RecordBatch]:
    """Broadcast-build / partitioned-probe execution of a left-deep join tree under an aggregate (SURVEY.md §8e, Q3').

    stage1: a plan whose root is a join probed by this rank's shard (its build side is small and present in full on
    every rank); the join output of all ranks — all-gathered in rank order, which is the single-process output
    order because shards are contiguous — becomes table `build_slot` of stage2 on every rank.
    stage2: root = aggregate over a join whose build side is that table and whose probe side is this rank's shard;
    it runs through `sharded_aggregate` (partial aggregation + exchange of the groups).
    `stage*_tables`: {slot: RecordBatch | tpch.DeviceTable}."""
    def push(plan, tables):
        for slot, t in tables.items():
            if isinstance(t, pa.RecordBatch):
                plan.push_table(slot, t)
            else:
                plan.push_table_device(slot, t)

    p1 = builder.build(stage1, stage1_schemas)
    push(p1, stage1_tables)
    local = p1.run()
    p1.close()
    build_side = all_gather_batches(group, local, stage1.output_schema(stage1_schemas))
    p2 = builder.build(stage2, stage2_schemas)
    for b in build_side:
        p2.push_table(build_slot, b)
    if not build_side:
        p2.push_table(build_slot, pa.RecordBatch.from_pylist([], schema=stage2_schemas[build_slot]))
    push(p2, stage2_tables)
    # rank-major first-appearance order: join output rows of rank r come after those of rank r-1
    result = sharded_aggregate(p2, group, row_base=group.rank << 40)
    p2.close()
    return result


def copartitioned_shard(n_orders: int, rank: int, world: int):
    """Row ranges of rank `rank` for the synthetic orders / lineitem tables range-partitioned on orderkey: the generator
    lays out the lines of orders [7b, 7b+7) in lineitem rows [28b, 28b+28) (include/sqlrs_tpch_spec.h), so cutting
    orders at multiples of 7 and lineitem at the matching multiples of 28 puts every order next to all its lines.
    Returns ((orders_lo, orders_hi), (lineitem_lo, lineitem_hi or None = to the end))."""
    blocks = n_orders // 7
    b_lo, b_hi = blocks * rank // world, blocks * (rank + 1) // world
    if rank == world - 1:
        return (7 * b_lo, n_orders), (28 * b_lo, None)
    return (7 * b_lo, 7 * b_hi), (28 * b_lo, 28 * b_hi)


def copartitioned_topk(plan, group: "TorchGroup", order_by, limit: int, offset: Optional[int] = None) -> List[pa.RecordBatch]:
    """Multi-GPU execution of  Limit(Project(Order(Aggregate(joins...))))  over tables that are co-partitioned on a
    join key which is also a group-by key ("scan stays GPU-local").

    Every rank has pushed its shards of the partitioned tables and the whole of the replicated (dimension) tables
    into `plan` — the full query plan, tail included.  Because the partition key is a group key, the groups of
    different ranks are disjoint, so the global top rows are among the per-rank top rows: each rank runs the plan
    locally (no collective on the data path), the <= offset+limit local rows are gathered on rank 0, which orders
    them again with the library's own Order / Limit operators.  `order_by`: BoundOrderBy list over the plan's OUTPUT
    columns.  Ties between ranks resolve in rank order = global row order, as in the single-process run."""
    import os
    import time

    from . import executor as ex

    trace = os.environ.get("SQLRS_B200_DIST_TRACE") == "1"
    t0 = time.perf_counter()
    local = plan.run()
    t1 = time.perf_counter()
    schema = local[0].schema if local else None
    table = pa.Table.from_batches(local).combine_chunks() if local else None
    payload = _to_bytes(table.to_batches()[0]) if table is not None and table.num_rows else b""
    gathered = group.gather_small(payload, dst=0)
    t2 = time.perf_counter()
    if gathered is None:
        if trace:
            print(f"[dist trace] rank {group.rank}: local plan {1e3 * (t1 - t0):.3f} ms, gather {1e3 * (t2 - t1):.3f} ms", flush=True)
        return []
    batches = [_from_bytes(data) for data in gathered if data]
    if not batches:
        return [pa.RecordBatch.from_pylist([], schema=schema)] if schema is not None else []
    ordered = ex.OrderExecutor(order_by, batches, lib=plan.lib, options=plan.options).execute()
    out = ex.try_collect(ex.LimitExecutor(limit, offset, ordered, lib=plan.lib, options=plan.options).execute())
    if trace:
        print(f"[dist trace] rank {group.rank}: local plan {1e3 * (t1 - t0):.3f} ms, gather {1e3 * (t2 - t1):.3f} ms, merge {1e3 * (time.perf_counter() - t2):.3f} ms",
              flush=True)
    return out
