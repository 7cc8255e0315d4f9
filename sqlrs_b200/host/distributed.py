"""Multi-GPU execution of plans over row-sharded tables.

One process per GPU (torch.distributed; NCCL over NVLink on the GPU box, gloo in the CPU tests).  Scan, filter, join
probe and the partial aggregation stay GPU-local; what crosses NVLink are the exchange steps of the stateful operators,
and with NCCL + the CUDA library they cross it as DEVICE buffers — no Arrow IPC, no host copy of the data plane:

  group-by   sqlrs_plan_execute_partial, then
             few groups : every rank packs its group table into a fixed-size device buffer, ONE all-gather, rank 0 folds
                          them (sqlrs_plan_export_partials_device / _merge_partials_device); the header counts are read
                          back asynchronously and checked at the END of the step, not between collective and merge
             many groups: a partition kernel packs the partial groups by identity hash mod world into per-owner regions
                          (sqlrs_plan_export_partials_partitioned), ONE all_to_all_single with equal splits, every owner
                          folds what it received, the owner-merged (disjoint) groups are all-gathered and finalised on
                          rank 0 in the reference's first-appearance order
  join       broadcast-build / partitioned-probe: every rank runs the build side's sub-plan on its shard, the surviving
             rows are all-gathered device-to-device in rank order (= the single-process row order: shards are contiguous)
             and every rank builds the whole table (`broadcast_rows`); the probe side stays sharded.
             When the caller's table statistics (min / max of the join key per shard, computed once at load time like a
             zone map) show that a build shard covers every key its probe shard can hold and the shards' key ranges are
             disjoint, the join needs no exchange at all (`key_ranges_copartitioned`): `distributed_join_topk`.

The reference has no distributed execution; results equal the single-process ones (integers bit-exact, float sums up to
summation order).  With gloo or the CPU checker the same steps run through host batches (`_export_partials` etc.).
This module contains no compute.
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


def _check_stream(plan, group) -> bool:
    """stream contract of the *_device calls (include/sqlrs_b200.h): a plan on the caller's stream works in stream order with
    the collectives; a plan that owns its stream synchronises inside its exports, and the caller synchronises before merges"""
    torch = group.torch
    shared = _plan_stream(plan) == torch.cuda.current_stream(group.device).cuda_stream
    if _plan_stream(plan) and not shared:
        raise ValueError("distributed execution: options.stream must be torch's current stream (or NULL)")
    return shared


def _buffers(group, key, make):
    bufs = getattr(group, "_bufs", None)
    if bufs is None:
        bufs = group._bufs = {}
    if key not in bufs:
        bufs[key] = make()
    return bufs[key]


def _row_words(plan) -> int:
    words = C.c_int32(0)
    plan.lib.check(plan.lib.plan_partials_row_words(plan.handle, C.byref(words)))
    return int(words.value)


def _device_exchange(plan, group: TorchGroup, cap_rows: int):
    """Few groups (NCCL): partial groups never leave HBM.  Every rank packs its groups into a fixed-size device buffer, ONE
    all-gather over NVLink hands all of them to every rank, rank 0 folds them and finalises.  Nothing synchronises between
    the collective and the merge: the per-rank group counts (buffer headers) are copied to pinned memory asynchronously and
    looked at when the step's work has been enqueued.  Returns None when some rank had more than `cap_rows` groups (every
    rank sees the same headers, so every rank takes the radix path together)."""
    torch, lib, W = group.torch, plan.lib, group.world
    words = _row_words(plan)

    def make():
        n = (cap_rows + 1) * words
        return (torch.empty(n, dtype=torch.int64, device=group.device), torch.empty(n * W, dtype=torch.int64, device=group.device),
                torch.empty(W, dtype=torch.int64, pin_memory=True))

    send, recv, counts_host = _buffers(group, ("few", words, cap_rows), make)
    shared = _check_stream(plan, group)
    lib.check(lib.plan_export_partials_device(plan.handle, C.c_void_p(send.data_ptr()), cap_rows))
    group.dist.all_gather_into_tensor(recv, send)
    counts_host.copy_(recv.view(W, cap_rows + 1, words)[:, 0, 0], non_blocking=True)
    if not shared:
        torch.cuda.current_stream(group.device).synchronize()
    lib.check(lib.plan_clear_partials(plan.handle))
    result = []
    if group.rank == 0:
        lib.check(lib.plan_merge_partials_device(plan.handle, C.c_void_p(recv.data_ptr()), W, cap_rows))
        lib.check(lib.plan_finish_partial(plan.handle))
        result = plan.collect()
    else:
        lib.check(lib.plan_finish_partial(plan.handle))
        plan.collect()
    torch.cuda.current_stream(group.device).synchronize()  # counts_host is complete (rank 0 synchronised for its result anyway)
    if int(counts_host.max()) > cap_rows:
        return None
    return result


def _radix_exchange(plan, group: TorchGroup, finish: bool = True):
    """Many groups (NCCL): radix-partition the partial groups by identity hash mod world into per-owner regions of a device
    send buffer, exchange them with ONE all_to_all_single (equal splits: every region carries its own count), fold on the
    owner, then all-gather the owner-merged groups — disjoint across ranks — and finalise on rank 0.
    `finish=False`: stop once rank 0's selected table holds the merged groups (the caller exchanges further tables)."""
    torch, lib, W = group.torch, plan.lib, group.world
    words = _row_words(plan)
    shared = _check_stream(plan, group)
    groups = C.c_int64(0)
    lib.check(lib.plan_export_partials_partitioned(plan.handle, None, W, 0, C.byref(groups)))  # learn the local group count
    t = torch.tensor([groups.value], dtype=torch.int64, device=group.device)
    group.dist.all_reduce(t, op=group.dist.ReduceOp.MAX)
    most = int(t.item())
    for attempt in range(3):
        # a hash partition holds ~1/W of a rank's groups; head-room for skew, then x4 per retry
        cap = max(256, int(most / W * 1.25 * (4 ** attempt)) + 1024)
        cap = min(cap, max(most, 1))
        region = (cap + 1) * words
        send = torch.empty(region * W, dtype=torch.int64, device=group.device)
        recv = torch.empty(region * W, dtype=torch.int64, device=group.device)
        lib.check(lib.plan_export_partials_partitioned(plan.handle, C.c_void_p(send.data_ptr()), W, cap, C.byref(groups)))
        group.dist.all_to_all_single(recv, send)
        worst = recv.view(W, cap + 1, words)[:, 0, 0].max().reshape(1).clone()
        group.dist.all_reduce(worst, op=group.dist.ReduceOp.MAX)
        if int(worst.item()) <= cap:
            break
    else:
        raise RuntimeError("radix exchange: a hash partition kept overflowing its region")
    if not shared:
        torch.cuda.current_stream(group.device).synchronize()
    lib.check(lib.plan_clear_partials(plan.handle))
    lib.check(lib.plan_merge_partials_device(plan.handle, C.c_void_p(recv.data_ptr()), W, cap))
    # the owner-merged groups of all ranks -> rank 0
    owned = C.c_int64(0)
    lib.check(lib.plan_export_partials_partitioned(plan.handle, None, 1, 0, C.byref(owned)))
    t = torch.tensor([owned.value], dtype=torch.int64, device=group.device)
    group.dist.all_reduce(t, op=group.dist.ReduceOp.MAX)
    cap2 = max(int(t.item()), 1)
    send2 = torch.empty((cap2 + 1) * words, dtype=torch.int64, device=group.device)
    recv2 = torch.empty((cap2 + 1) * words * W, dtype=torch.int64, device=group.device)
    lib.check(lib.plan_export_partials_device(plan.handle, C.c_void_p(send2.data_ptr()), cap2))
    group.dist.all_gather_into_tensor(recv2, send2)
    if not shared:
        torch.cuda.current_stream(group.device).synchronize()
    lib.check(lib.plan_clear_partials(plan.handle))
    if group.rank == 0:
        lib.check(lib.plan_merge_partials_device(plan.handle, C.c_void_p(recv2.data_ptr()), W, cap2))
    if not finish:
        return None
    lib.check(lib.plan_finish_partial(plan.handle))
    result = plan.collect()
    return result if group.rank == 0 else []


def _partials_tables(plan) -> int:
    n = C.c_int32(1)
    plan.lib.check(plan.lib.plan_partials_tables(plan.handle, C.byref(n)))
    return int(n.value)


def _select(plan, table: int):
    plan.lib.check(plan.lib.plan_select_partials_table(plan.handle, table))


def sharded_aggregate(plan, group: TorchGroup, row_base: int = 0, device_cap_rows: int = 256) -> List[pa.RecordBatch]:
    """Runs `plan` (root = aggregate) over the shard pushed on this rank; rank 0 returns the final batches.
    `row_base`: global row number of the shard's first row (first-appearance order across ranks).
    DISTINCT aggregates: the partial state is several tables (the plain aggregates' groups + the set elements of every
    DISTINCT aggregate); each is exchanged like a group table — partitioned by its own identity hash — and rank 0 finalises
    once all of them are merged."""
    lib = plan.lib
    lib.check(lib.plan_execute_partial(plan.handle, row_base))
    tables = _partials_tables(plan)
    if group.native_a2a and lib.prefix == "sqlrs_":  # NCCL + the CUDA library: the exchange stays in HBM
        if tables > 1:
            for t in range(tables):
                _select(plan, t)
                _radix_exchange(plan, group, finish=False)
            if group.rank != 0:
                return []  # nothing to finalise here: the merged tables live on rank 0
            lib.check(lib.plan_finish_partial(plan.handle))
            return plan.collect()
        if device_cap_rows > 0:
            if not getattr(plan, "_many_groups", False):
                result = _device_exchange(plan, group, device_cap_rows)
                if result is not None:
                    return result
                plan._many_groups = True  # remembered: later runs of this plan go straight to the radix exchange
                lib.check(lib.plan_execute_partial(plan.handle, row_base))  # the few-groups attempt consumed the partial state
        return _radix_exchange(plan, group)
    # host path (gloo / the CPU checker): the same steps through Arrow batches.  Every table is exported before any is
    # cleared and merged in table order (table 0 = the groups themselves)
    def export_all():
        out = []
        for t in range(tables):
            _select(plan, t)
            out.append(_export_partials(plan))
        return out

    def replace_all(received):  # received[t] = batches to fold into table t
        for t in reversed(range(tables)):
            _select(plan, t)
            lib.check(lib.plan_clear_partials(plan.handle))
        for t in range(tables):
            _select(plan, t)
            for batch in received[t]:
                _merge_partials(plan, batch)

    local = export_all()
    received = [[_from_bytes(d) for d in group.all_to_all_bytes([_to_bytes(p) for p in partition_by_owner(local[t], group.world)])]
                for t in range(tables)]
    replace_all(received)
    owned = export_all()
    gathered = [group.gather_bytes(_to_bytes(owned[t]), dst=0) for t in range(tables)]
    replace_all([[_from_bytes(d) for d in g] if g is not None else [] for g in gathered])
    if tables > 1 and group.rank != 0:
        return []
    lib.check(lib.plan_finish_partial(plan.handle))
    result = plan.collect()
    return result if group.rank == 0 else []


# ---------------------------------------------------------------------------------------------- joins
_MARK_T = [None]


def _mark(group, label):
    """SQLRS_B200_DIST_TRACE=1: wall time per phase with a device synchronisation at every mark (attributes device work to
    the phase that enqueued it); =host: the host's own timeline, nothing added (shows where the host waits)."""
    import os
    import time

    mode = os.environ.get("SQLRS_B200_DIST_TRACE")
    if mode not in ("1", "host"):
        return
    if mode == "1":
        group.torch.cuda.synchronize(group.device)
    now = time.perf_counter()
    if _MARK_T[0] is not None and label:
        print(f"[dist trace] rank {group.rank}: {label} {1e3 * (now - _MARK_T[0]):.3f} ms", flush=True)
    _MARK_T[0] = now


class DeviceBatch:
    """Columns resident in HBM as torch tensors (int64 storage; Float64 columns as their bit patterns), exportable as an
    ArrowDeviceArray for `GpuPlan.push_table_device` (same interface as tpch.DeviceTable)."""

    def __init__(self, schema: pa.Schema, tensors: list, n_rows: int, device_index: int):
        self._schema, self.tensors, self.n_rows, self.device_index = schema, tensors, n_rows, device_index
        self._keep: list = []

    @property
    def schema(self) -> pa.Schema:
        return self._schema

    def export(self):
        from .tpch import DeviceTable

        return DeviceTable.export(self)


def merge_topk_device(plan, group: TorchGroup, schema: pa.Schema, order_by, limit: int, state: dict, builder) -> List[pa.RecordBatch]:
    """`merge_topk` without leaving HBM until the final rows: the <= limit rows pending in `plan` (a DEVICE result) are copied
    into one fixed-size tensor [count | limit x columns], all-gathered, and rank 0 runs Limit(Order(Scan)) over the W x limit
    candidate rows on the device — rank-major row order, so ties resolve as in the single-process run.
    Rank 0 does not wait for the per-rank counts before it orders: it assumes the usual case (every rank returned `limit` rows),
    enqueues the tail plan on the W x limit candidates, and looks at the counts — copied to pinned memory ahead of the tail —
    when the tail's result has arrived; only if some rank had fewer rows is the tail run again over the exact candidates."""
    from .plan import PhysicalLimit, PhysicalOrder, PhysicalTableScan

    torch, W = group.torch, group.world
    ncols = len(schema)
    shape = plan.result_shape()
    n = shape[0] if shape is not None else 0
    if n > limit:
        raise ValueError("merge_topk_device: the local plan returned more rows than the limit")
    key = ("topk", ncols, limit)
    if key not in state:
        state[key] = (torch.zeros(1 + ncols * limit, dtype=torch.int64, device=group.device),
                      torch.empty(W * (1 + ncols * limit), dtype=torch.int64, device=group.device),
                      torch.empty((ncols, W * limit), dtype=torch.int64, device=group.device),
                      torch.empty(W, dtype=torch.int64, pin_memory=True))
    send, recv, cand, counts_host = state[key]
    send[:1].fill_(n)
    if shape is not None:
        body = send[1:].view(ncols, limit)
        plan.next_to_device([body[c].data_ptr() for c in range(ncols)])
    shared = _check_stream(plan, group)
    group.dist.all_gather_into_tensor(recv, send)
    if group.rank != 0:
        return []
    every = recv.view(W, 1 + ncols * limit)
    counts_host.copy_(every[:, 0], non_blocking=True)
    cand.view(ncols, W, limit).copy_(every[:, 1:].view(W, ncols, limit).permute(1, 0, 2))  # rank-major rows, one strided copy
    if "p_tail" not in state:
        tail = PhysicalLimit(limit, None, PhysicalOrder(order_by, PhysicalTableScan(0)))
        state["p_tail"] = builder.build(tail, {0: schema})
    p_tail = state["p_tail"]
    dev = group.device.index if group.device.index is not None else 0
    if not shared:
        torch.cuda.current_stream(group.device).synchronize()
    if not state.get("tail_pushed"):  # the candidate buffer is the same memory every step: the tail plan keeps scanning it
        p_tail.reset()
        p_tail.push_table_device(0, DeviceBatch(schema, [cand[c] for c in range(ncols)], W * limit, dev))
        state["tail_pushed"] = True
    out = p_tail.run()
    torch.cuda.current_stream(group.device).synchronize()  # counts_host is complete (a no-op after the tail's own result copy)
    counts = counts_host.tolist()
    _mark(group, "  top-k: tail plan over W x limit candidates")
    if all(c == limit for c in counts):
        return out
    # some rank had fewer than `limit` rows: order exactly the rows that exist
    cols = [torch.cat([every[r, 1:].view(ncols, limit)[c, :counts[r]] for r in range(W)]) for c in range(ncols)]
    state["tail_pushed"] = False
    p_tail.reset()
    p_tail.push_table_device(0, DeviceBatch(schema, cols, int(sum(counts)), dev))
    return p_tail.run()


def packing_index(torch, counts, ncols: int, most: int):
    """Flat positions in an all-gathered buffer recv[W, ncols, most] (rank r filled its first counts[r] rows of every column) of
    the packed layout [ncols, sum(counts)] whose rows are the ranks' rows in rank order: torch.take(recv, index) packs them."""
    W, total = len(counts), int(sum(counts))
    rank_of = torch.repeat_interleave(torch.arange(W), torch.tensor(counts, dtype=torch.int64))
    starts = torch.cumsum(torch.tensor((0,) + tuple(counts[:-1]), dtype=torch.int64), 0)
    within = torch.arange(total, dtype=torch.int64) - starts[rank_of]
    return (rank_of * (ncols * most) + within).unsqueeze(0) + (torch.arange(ncols, dtype=torch.int64) * most).unsqueeze(1)


def broadcast_rows(plan, group: TorchGroup, schema: pa.Schema, state: Optional[dict] = None, while_waiting=None) -> DeviceBatch:
    """The rows `plan` (already executed on this rank's shard; ONE pending device result of fixed-width, NULL-free columns)
    produced on every rank, concatenated in rank order, on every rank — all-gathered device to device.  This is the
    build-side broadcast of a broadcast-build / partitioned-probe join.
    `state`: a dict the caller keeps across calls (buffers and the packing index are reused while the row counts repeat);
    `while_waiting`: host work that does not depend on the exchange, done while the row counts travel."""
    torch, W = group.torch, group.world
    state = state if state is not None else {}
    shape = plan.result_shape()
    n, ncols = shape if shape is not None else (0, len(schema))
    if "bc_counts" not in state:
        state["bc_counts"] = (torch.empty(1, dtype=torch.int64, device=group.device), torch.empty(W, dtype=torch.int64, device=group.device),
                              torch.empty(W, dtype=torch.int64, pin_memory=True))
    mine, counts_dev, counts_host = state["bc_counts"]
    mine.fill_(n)
    group.dist.all_gather_into_tensor(counts_dev, mine)
    counts_host.copy_(counts_dev, non_blocking=True)
    arrived = torch.cuda.Event()
    arrived.record(torch.cuda.current_stream(group.device))
    if while_waiting is not None:
        while_waiting()
    arrived.synchronize()  # sizes of the receive buffers: the one host round trip of this step
    counts = tuple(counts_host.tolist())
    _mark(group, "  broadcast: row counts known on the host")
    most, total = max(max(counts), 1), int(sum(counts))
    key = ("bc_bufs", ncols, counts)
    if state.get("bc_key") != key:
        flat = packing_index(torch, counts, ncols, most)
        state["bc_key"] = key
        state["bc_bufs"] = (torch.empty((ncols, most), dtype=torch.int64, device=group.device),
                            torch.empty((W, ncols, most), dtype=torch.int64, device=group.device),
                            torch.empty((ncols, max(total, 1)), dtype=torch.int64, device=group.device),
                            flat.to(group.device))
    send, recv, packed, flat = state["bc_bufs"]
    if shape is not None:
        plan.next_to_device([send[c].data_ptr() for c in range(ncols)])
    _check_stream(plan, group)
    group.dist.all_gather_into_tensor(recv, send)
    if total:
        torch.take(recv, flat, out=packed)  # one gather kernel: packed [ncols, total]
    dev = group.device.index if group.device.index is not None else 0
    _mark(group, "  broadcast: all-gather + packing enqueued")
    return DeviceBatch(schema, [packed[c, :total] for c in range(ncols)], total, dev)


def key_ranges_copartitioned(group: TorchGroup, build_range, probe_range) -> bool:
    """`build_range` / `probe_range`: (min, max) of the join key in this rank's shard of the build-side / probe-side table
    (table statistics, computed once at load time).  True on every rank iff on every rank the probe shard's keys lie inside
    the build shard's range and the build shards' ranges are pairwise disjoint — then every probe row finds all its matches
    on its own rank and the join needs no exchange."""
    torch, W = group.torch, group.world
    mine = torch.tensor([build_range[0], build_range[1], probe_range[0], probe_range[1]], dtype=torch.int64, device=group.device)
    every = torch.empty(W * 4, dtype=torch.int64, device=group.device)
    group.dist.all_gather_into_tensor(every, mine) if group.native_a2a else group.dist.all_gather(list(every.view(W, 4).unbind(0)), mine)
    r = every.view(W, 4).tolist()
    inside = all(b_lo <= p_lo and p_hi <= b_hi for b_lo, b_hi, p_lo, p_hi in r if p_lo <= p_hi)
    spans = sorted((b_lo, b_hi) for b_lo, b_hi, _, _ in r if b_lo <= b_hi)
    disjoint = all(spans[i][1] < spans[i + 1][0] for i in range(len(spans) - 1))
    return inside and disjoint


def merge_topk(plan, group: TorchGroup, local: List[pa.RecordBatch], order_by, limit: int) -> List[pa.RecordBatch]:
    """Global ORDER BY ... LIMIT over per-rank results whose groups are disjoint across ranks: the <= limit local rows of
    every rank are gathered on rank 0 (a few hundred bytes per rank) and ordered again with the library's own Order / Limit
    operators.  Ties between ranks resolve in rank order = global row order, as in the single-process run."""
    from . import executor as ex

    schema = local[0].schema if local else None
    table = pa.Table.from_batches(local).combine_chunks() if local else None
    payload = _to_bytes(table.to_batches()[0]) if table is not None and table.num_rows else b""
    gathered = group.gather_small(payload, dst=0)
    if gathered is None:
        return []
    batches = [_from_bytes(data) for data in gathered if data]
    if not batches:
        return [pa.RecordBatch.from_pylist([], schema=schema)] if schema is not None else []
    ordered = ex.OrderExecutor(order_by, batches, lib=plan.lib, options=plan.options).execute()
    return ex.try_collect(ex.LimitExecutor(limit, None, ordered, lib=plan.lib, options=plan.options).execute())


def _push(plan, tables):
    for slot, t in tables.items():
        if isinstance(t, pa.RecordBatch):
            plan.push_table(slot, t)
        else:
            plan.push_table_device(slot, t)


def broadcast_build_join_aggregate(builder, group: TorchGroup, stage1, stage1_schemas, stage1_tables, stage2, stage2_schemas, stage2_tables,
                                   build_slot: int, row_base: Optional[int] = None) -> List[pa.RecordBatch]:
    """Broadcast-build / partitioned-probe execution of a left-deep join tree under an aggregate (SURVEY.md §8e, Q3').

    stage1: a plan whose root is a join probed by this rank's shard (its build side is present in full on every rank);
    the join output of all ranks — all-gathered in rank order, which is the single-process output order because shards
    are contiguous — becomes table `build_slot` of stage2 on every rank.  With NCCL and the CUDA library the rows move
    device to device (`broadcast_rows`); otherwise as Arrow batches through the host.
    stage2: root = aggregate over a join whose build side is that table and whose probe side is this rank's shard;
    it runs through `sharded_aggregate` (partial aggregation + exchange of the groups).
    `stage*_tables`: {slot: RecordBatch | DeviceBatch | tpch.DeviceTable}.  `row_base`: global row number of the first row
    of this rank's stage-2 probe shard (default: rank << 40, enough to order ranks; < 16 ranks)."""
    device_path = group.native_a2a and builder.lib.prefix == "sqlrs_"
    p1 = builder.build(stage1, stage1_schemas)
    _push(p1, stage1_tables)
    p2 = builder.build(stage2, stage2_schemas)
    if device_path:
        p1.execute()
        build_side = broadcast_rows(p1, group, stage1.output_schema(stage1_schemas))
        p1.close()
        p2.push_table_device(build_slot, build_side)
    else:
        local = p1.run()
        p1.close()
        build_side = all_gather_batches(group, local, stage1.output_schema(stage1_schemas))
        for b in build_side:
            p2.push_table(build_slot, b)
        if not build_side:
            p2.push_table(build_slot, pa.RecordBatch.from_pylist([], schema=stage2_schemas[build_slot]))
    _push(p2, stage2_tables)
    if row_base is None:
        if group.world > 16:
            raise ValueError("broadcast_build_join_aggregate: pass row_base for more than 16 ranks (ordinals are 44 bits)")
        row_base = group.rank << 40  # rank-major first-appearance order: rows of rank r come after those of rank r-1
    result = sharded_aggregate(p2, group, row_base=row_base)
    p2.close()
    return result


def _same_tables(prev: dict, now: dict) -> bool:
    """the same DEVICE-resident tables as last time (a host batch is copied to the device on every call: that is the call's input)"""
    return list(prev) == list(now) and all(prev[k] is now[k] and not isinstance(now[k], pa.RecordBatch) for k in now)


def distributed_join_topk(builder, group: TorchGroup, *, build_plan, build_schemas, build_tables, query_plan, query_schemas, query_tables,
                          build_slot: int, order_by, limit: int, state: Optional[dict] = None) -> List[pa.RecordBatch]:
    """ORDER BY ... LIMIT query over a join tree whose fact tables are sharded so that the joins between them need no
    exchange (the caller established that with `key_ranges_copartitioned`) and whose first build side is small:

      1. `build_plan` (e.g. Filter(Scan) over this rank's shard of the dimension table) runs locally; its surviving rows are
         all-gathered device to device, so every rank holds the whole filtered build side — never the whole table;
      2. `query_plan` — the complete query, tail included — runs locally with that table in `build_slot`; the partition key
         is a group key, so groups of different ranks are disjoint and the global top rows are among the per-rank top rows;
      3. the <= `limit` rows per rank are gathered on rank 0 and ordered again (`merge_topk`).

    `state`: a dict kept by the caller across calls; the built plans live there so that repeated runs reuse compiled kernels,
    device buffers and sizing hints."""
    def mark(label):
        _mark(group, label)

    state = state if state is not None else {}
    if "p_build" not in state:
        state["p_build"] = builder.build(build_plan, build_schemas)
        state["p_query"] = builder.build(query_plan, query_schemas)
    p_build, p_query = state["p_build"], state["p_query"]
    mark("")
    # tables that are the same objects as in the previous (completed) call stay pushed: only the exchanged build side changes
    same = state.get("pushed") is not None and all(_same_tables(x, y) for x, y in zip(state["pushed"], (build_tables, query_tables)))
    state["pushed"] = None  # (set again when this call completes: a failed call leaves the plans to a full reset)
    if same:
        p_query.clear_table(build_slot)
        later = None
    else:
        p_build.reset()
        p_query.reset()
        _push(p_build, build_tables)
        later = lambda: _push(p_query, query_tables)  # noqa: E731 — pushed while the row counts of the exchange travel
    device_path = group.native_a2a and builder.lib.prefix == "sqlrs_"
    if device_path:
        p_build.execute()
        mark("build-side sub-plan (local shard)")
        build_side = broadcast_rows(p_build, group, build_plan.output_schema(build_schemas), state, while_waiting=later)
        mark("all-gather of its rows")
        p_query.push_table_device(build_slot, build_side)
    else:  # gloo / the CPU checker: the same steps through host batches
        schema = build_plan.output_schema(build_schemas)
        rows = all_gather_batches(group, p_build.run(), schema)
        if later is not None:
            later()
        for b in rows or [pa.RecordBatch.from_pylist([], schema=schema)]:
            p_query.push_table(build_slot, b)
    mark("  query: tables pushed")
    if device_path:
        p_query.execute()
        mark("query plan (local shards)")
        out = merge_topk_device(p_query, group, query_plan.output_schema(query_schemas), order_by, limit, state, builder)
    else:
        local = p_query.run()
        mark("query plan (local shards)")
        out = merge_topk(p_query, group, local, order_by, limit)
    mark("top-k merge")
    state["pushed"] = (dict(build_tables), dict(query_tables))
    return out


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
    if offset:
        # every rank's plan carries the query's own Limit: with an OFFSET each rank would drop ITS first rows, not the global ones
        raise ValueError("copartitioned_topk: build the per-rank plan with Limit(limit + offset) and apply the OFFSET to the merged rows")
    return _copartitioned_topk(plan, group, order_by, limit)


def _copartitioned_topk(plan, group: "TorchGroup", order_by, limit: int, offset: Optional[int] = None) -> List[pa.RecordBatch]:
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
