"""The N>1 path on CPU: two gloo ranks, each aggregating its row shard with the CPU
checker behind the same C ABI, partial groups exchanged all-to-all by key hash, merged by their owner,
finalised on rank 0 — must equal the single-process result (first-appearance order included).
The same host code (sqlrs_b200/host/distributed.py) runs over NCCL with the CUDA library in bench.py."""
import os
import sys

import pyarrow as pa
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, plan_name, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch

    from sqlrs_b200.host import distributed as sqdist
    from sqlrs_b200.host import ffi, tpch
    from sqlrs_b200.host.plan import ExecutorBuilder

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lib = ffi.Library(os.path.join(ROOT, "oracle", "liboracle.so"), "sqlrs_oracle_")
    d = tpch.dims(0.02)
    plan, schemas = tpch.q1_plan()
    if plan_name == "distinct":  # DISTINCT aggregates: the sets of every group are one more table of the exchange each
        from sqlrs_b200.host.expr import AggFunc, InputRef
        from sqlrs_b200.host.plan import PhysicalHashAgg

        col = {f.name: InputRef(i, ffi.dtype_of(f.type)) for i, f in enumerate(schemas[0])}
        plan = PhysicalHashAgg([AggFunc("Count", [col["l_quantity_i64"]], distinct=True), AggFunc("Sum", [col["l_quantity_i64"]]),
                                AggFunc("Sum", [col["l_quantity_i64"]], distinct=True), AggFunc("Count", [col["l_shipdate"]], distinct=True),
                                AggFunc("Count", [col["l_quantity"]])], [col["l_returnflag"], col["l_linestatus"]], plan.child)
    opts = lib.options(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
    n = tpch.num_rows(lib, d, tpch.LINEITEM)
    lo, hi = n * rank // world, n * (rank + 1) // world
    shard = tpch.host_table(lib, d, tpch.LINEITEM, lo, hi, columns=tpch.Q1_COLUMNS)
    p = ExecutorBuilder(lib, opts).build(plan, schemas)
    p.push_table(0, shard.slice(0, shard.num_rows // 2))
    p.push_table(0, shard.slice(shard.num_rows // 2))
    group = sqdist.TorchGroup(dist, torch.device("cpu"))
    result = sqdist.sharded_aggregate(p, group, lo)
    if rank == 0:
        whole = ExecutorBuilder(lib, opts).build(plan, schemas)
        whole.push_table(0, tpch.host_table(lib, d, tpch.LINEITEM, columns=tpch.Q1_COLUMNS))
        expect = whole.run()
        from util import assert_batches_match

        assert_batches_match(result, expect, rtol=1e-9)
        assert result[0].num_rows == 8
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    else:
        assert result == []
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_rank_group_by_matches_single_process(tmp_path, oracle):
    mp.spawn(_worker, args=(2, _free_port(), "q1", str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").read_text() == "ok"


@pytest.mark.parametrize("world", [2, 3])
def test_distinct_aggregates_across_ranks_match_single_process(tmp_path, oracle, world):
    """COUNT(DISTINCT) / SUM(DISTINCT) in the partial/final split: sets merged across ranks == one process over all rows"""
    mp.spawn(_worker, args=(world, _free_port(), "distinct", str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").read_text() == "ok"


def test_partition_by_owner_is_a_partition():
    sys.path.insert(0, ROOT)
    import numpy as np

    from sqlrs_b200.host import distributed as sqdist

    h = np.array([0, 1, 2, 3, -1, -2, 2**63 - 1, -2**63], dtype=np.int64)
    b = pa.RecordBatch.from_arrays([pa.array(h), pa.array(np.arange(len(h)))], names=["hash", "x"])
    parts = sqdist.partition_by_owner(b, 3)
    assert sum(p.num_rows for p in parts) == len(h)
    for r, p in enumerate(parts):
        assert all((int(v) % 2**64) % 3 == r for v in p.column(0).to_pylist())


def _q3_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch

    from sqlrs_b200.host import distributed as sqdist
    from sqlrs_b200.host import ffi, tpch
    from sqlrs_b200.host.plan import ExecutorBuilder

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lib = ffi.Library(os.path.join(ROOT, "oracle", "liboracle.so"), "sqlrs_oracle_")
    d = tpch.dims(0.02)
    opts = lib.options(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
    builder = ExecutorBuilder(lib, opts)
    (s1, s1_schemas), (s2, s2_schemas) = tpch.q3_stage_plans()

    def shard(table, cols):
        n = tpch.num_rows(lib, d, table)
        return tpch.host_table(lib, d, table, n * rank // world, n * (rank + 1) // world, columns=cols)

    customer = tpch.host_table(lib, d, tpch.CUSTOMER, columns=tpch.Q3_CUSTOMER_COLUMNS)  # the small build side: everywhere
    group = sqdist.TorchGroup(dist, torch.device("cpu"))
    result = sqdist.broadcast_build_join_aggregate(builder, group, s1, s1_schemas, {0: customer, 1: shard(tpch.ORDERS, tpch.Q3_ORDERS_COLUMNS)},
                                                   s2, s2_schemas, {1: shard(tpch.LINEITEM, tpch.Q3_LINEITEM_COLUMNS)}, build_slot=0)
    if rank == 0:
        plan, schemas = tpch.q3_plan()
        whole = builder.build(plan, schemas)
        whole.push_table(0, customer)
        whole.push_table(1, tpch.host_table(lib, d, tpch.ORDERS, columns=tpch.Q3_ORDERS_COLUMNS))
        whole.push_table(2, tpch.host_table(lib, d, tpch.LINEITEM, columns=tpch.Q3_LINEITEM_COLUMNS))
        expect = whole.run()
        from util import assert_batches_match

        assert expect[0].num_rows > 50
        assert_batches_match(result, expect, rtol=1e-9)
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_broadcast_build_join_matches_single_process(tmp_path, oracle):
    """Q3' with the join build sides broadcast and the probe sides sharded over 2 ranks == single-process Q3'"""
    mp.spawn(_q3_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").read_text() == "ok"


def _q3_copart_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch

    from sqlrs_b200.host import distributed as sqdist
    from sqlrs_b200.host import ffi, tpch
    from sqlrs_b200.host.plan import ExecutorBuilder

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lib = ffi.Library(os.path.join(ROOT, "oracle", "liboracle.so"), "sqlrs_oracle_")
    d = tpch.dims(0.02)
    opts = lib.options(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
    builder = ExecutorBuilder(lib, opts)
    plan, schemas = tpch.q3_full_plan()
    (o_lo, o_hi), (l_lo, l_hi) = sqdist.copartitioned_shard(int(d.n_orders), rank, world)
    customer = tpch.host_table(lib, d, tpch.CUSTOMER, columns=tpch.Q3_CUSTOMER_COLUMNS)  # dimension table: replicated
    orders = tpch.host_table(lib, d, tpch.ORDERS, o_lo, o_hi, columns=tpch.Q3_ORDERS_COLUMNS)
    lineitem = tpch.host_table(lib, d, tpch.LINEITEM, l_lo, l_hi, columns=tpch.Q3_LINEITEM_COLUMNS)
    # co-partitioned: every line of this rank belongs to one of this rank's orders
    ok, lk = orders.column(0).to_numpy(), lineitem.column(0).to_numpy()
    assert lk.min() >= ok.min() and lk.max() <= ok.max()
    p = builder.build(plan, schemas)
    p.push_table(0, customer)
    p.push_table(1, orders)
    p.push_table(2, lineitem)
    group = sqdist.TorchGroup(dist, torch.device("cpu"))
    result = sqdist.copartitioned_topk(p, group, tpch.q3_tail_order_by(), 10)
    p.close()
    if rank == 0:
        whole = builder.build(plan, schemas)
        whole.push_table(0, customer)
        whole.push_table(1, tpch.host_table(lib, d, tpch.ORDERS, columns=tpch.Q3_ORDERS_COLUMNS))
        whole.push_table(2, tpch.host_table(lib, d, tpch.LINEITEM, columns=tpch.Q3_LINEITEM_COLUMNS))
        expect = whole.run()
        from util import assert_batches_match

        assert sum(b.num_rows for b in expect) == 10
        assert_batches_match(result, expect, rtol=1e-9)
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    else:
        assert result == []
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_copartitioned_q3_full_query_matches_single_process(tmp_path, oracle, world):
    """Q3' incl. ORDER BY / LIMIT with orders and lineitem range-partitioned on orderkey over the ranks and customer
    replicated: every rank runs the whole query locally, rank 0 merges the per-rank top-10 — equals the single-process result"""
    mp.spawn(_q3_copart_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").read_text() == "ok"


def _q3_join_topk_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch

    from sqlrs_b200.host import distributed as sqdist
    from sqlrs_b200.host import ffi, tpch
    from sqlrs_b200.host.plan import ExecutorBuilder

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lib = ffi.Library(os.path.join(ROOT, "oracle", "liboracle.so"), "sqlrs_oracle_")
    d = tpch.dims(0.02)
    builder = ExecutorBuilder(lib, lib.options(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY))
    plan, schemas = tpch.q3_full_plan()
    cust_plan = plan.child.child.child.child.left.left  # Filter(Scan(customer)) of the query
    (o_lo, o_hi), (l_lo, l_hi) = sqdist.copartitioned_shard(int(d.n_orders), rank, world)
    n_c = tpch.num_rows(lib, d, tpch.CUSTOMER)
    shard = tpch.host_table(lib, d, tpch.CUSTOMER, n_c * rank // world, n_c * (rank + 1) // world, columns=tpch.Q3_CUSTOMER_COLUMNS)  # sharded by rows
    orders = tpch.host_table(lib, d, tpch.ORDERS, o_lo, o_hi, columns=tpch.Q3_ORDERS_COLUMNS)
    lineitem = tpch.host_table(lib, d, tpch.LINEITEM, l_lo, l_hi, columns=tpch.Q3_LINEITEM_COLUMNS)
    group = sqdist.TorchGroup(dist, torch.device("cpu"))
    state = {}
    for _ in range(2):  # the second call reuses the plans kept in `state`
        result = sqdist.distributed_join_topk(builder, group, build_plan=cust_plan, build_schemas={0: schemas[0]}, build_tables={0: shard}, query_plan=plan,
                                              query_schemas=schemas, query_tables={1: orders, 2: lineitem}, build_slot=0, order_by=tpch.q3_tail_order_by(),
                                              limit=10, state=state)
    if rank == 0:
        whole = builder.build(plan, schemas)
        whole.push_table(0, tpch.host_table(lib, d, tpch.CUSTOMER, columns=tpch.Q3_CUSTOMER_COLUMNS))
        whole.push_table(1, tpch.host_table(lib, d, tpch.ORDERS, columns=tpch.Q3_ORDERS_COLUMNS))
        whole.push_table(2, tpch.host_table(lib, d, tpch.LINEITEM, columns=tpch.Q3_LINEITEM_COLUMNS))
        expect = whole.run()
        from util import assert_batches_match

        assert sum(b.num_rows for b in expect) == 10
        assert_batches_match(result, expect, rtol=1e-9)
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    else:
        assert result == []
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_join_topk_with_a_sharded_dimension_table(tmp_path, oracle, world):
    """what bench.py times for Q3' at N > 1, through host batches: customer sharded by rows, its filtered rows all-gathered, the
    whole query per rank over co-partitioned orders / lineitem shards, per-rank top-10 merged == single-process result"""
    mp.spawn(_q3_join_topk_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").read_text() == "ok"


def test_packing_index_concatenates_ragged_rank_buffers():
    """the device all-gather moves fixed-size buffers; one gather through this index packs the ranks' rows in rank order"""
    sys.path.insert(0, ROOT)
    import torch

    from sqlrs_b200.host import distributed as sqdist

    for counts in ((2, 0, 3), (0, 0, 0, 1), (5,), (1, 1, 1, 1, 1, 1, 1, 7)):
        W, ncols, most = len(counts), 3, max(max(counts), 1)
        recv = torch.arange(W * ncols * most).view(W, ncols, most)
        index = sqdist.packing_index(torch, counts, ncols, most)
        expect = torch.cat([recv[r, :, :counts[r]] for r in range(W)], dim=1)
        assert index.shape == (ncols, sum(counts))
        assert torch.equal(torch.take(recv, index), expect)
    a, b = object(), object()
    assert sqdist._same_tables({0: a, 1: b}, {0: a, 1: b}) and not sqdist._same_tables({0: a, 1: b}, {0: a, 1: a})
    assert not sqdist._same_tables({0: a}, {0: a, 1: b}) and not sqdist._same_tables({1: a}, {0: a})
    host = pa.RecordBatch.from_pylist([{"x": 1}])
    assert not sqdist._same_tables({0: host}, {0: host})  # a host batch is this call's input: copied to the device every time


def test_copartitioned_shard_covers_everything():
    sys.path.insert(0, ROOT)
    from sqlrs_b200.host import distributed as sqdist

    for n_orders in (7, 100, 30000, 1500001):
        for world in (1, 2, 3, 8):
            o_prev, l_prev = 0, 0
            for r in range(world):
                (o_lo, o_hi), (l_lo, l_hi) = sqdist.copartitioned_shard(n_orders, r, world)
                assert o_lo == o_prev and l_lo == l_prev and o_lo % 7 == 0 and l_lo == 4 * o_lo
                o_prev, l_prev = o_hi, l_hi
            assert o_prev == n_orders and l_prev is None


def _gather_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch

    from sqlrs_b200.host import distributed as sqdist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    group = sqdist.TorchGroup(dist, torch.device("cpu"))
    small = bytes([rank + 1]) * (10 + rank)
    got = group.gather_small(small, dst=0, cap=64)
    assert (got == [bytes([r + 1]) * (10 + r) for r in range(world)]) if rank == 0 else got is None
    # one rank's payload does not fit the fixed-size frame: every rank falls back to the two-phase gather
    big = bytes([rank + 1]) * (200 if rank == 1 else 5)
    got = group.gather_small(big, dst=0, cap=64)
    assert (got == [bytes([r + 1]) * (200 if r == 1 else 5) for r in range(world)]) if rank == 0 else got is None
    assert group.gather_small(b"", dst=0, cap=64) == ([b""] * world if rank == 0 else None)
    if rank == 0:
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_gather_small_and_its_fallback(tmp_path):
    mp.spawn(_gather_worker, args=(3, _free_port(), str(tmp_path)), nprocs=3, join=True)
    assert (tmp_path / "ok").read_text() == "ok"
