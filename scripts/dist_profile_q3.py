"""Per-phase wall times of the N-GPU Q3' step (torchrun; SQLRS_B200_DIST_TRACE=1 synchronises between phases)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from sqlrs_b200.host import distributed as sqdist
from sqlrs_b200.host import ffi, tpch
from sqlrs_b200.host.plan import ExecutorBuilder

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = ffi.load()
sf = float(sys.argv[1]) if len(sys.argv) > 1 else 100
d = tpch.dims(sf)
stream = torch.cuda.Stream(device=dev)
with torch.cuda.stream(stream):
    builder = ExecutorBuilder(lib, lib.options(device_id=local, stream=C.c_void_p(stream.cuda_stream), count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY))
    group = sqdist.TorchGroup(dist, dev)
    n_c = tpch.num_rows(lib, d, tpch.CUSTOMER)
    (o_lo, o_hi), (l_lo, l_hi) = sqdist.copartitioned_shard(int(d.n_orders), rank, world)
    tabs = {0: tpch.device_table(lib, d, tpch.CUSTOMER, n_c * rank // world, n_c * (rank + 1) // world, columns=tpch.Q3_CUSTOMER_COLUMNS, device=dev),
            1: tpch.device_table(lib, d, tpch.ORDERS, o_lo, o_hi, columns=tpch.Q3_ORDERS_COLUMNS, device=dev),
            2: tpch.device_table(lib, d, tpch.LINEITEM, l_lo, l_hi, columns=tpch.Q3_LINEITEM_COLUMNS, device=dev)}
    full, schemas = tpch.q3_full_plan()
    cust = full.child.child.child.child.left.left
    state = {}
    import time

    def one():
        return sqdist.distributed_join_topk(builder, group, build_plan=cust, build_schemas={0: schemas[0]}, build_tables={0: tabs[0]}, query_plan=full, query_schemas=schemas,
                                            query_tables={1: tabs[1], 2: tabs[2]}, build_slot=0, order_by=tpch.q3_tail_order_by(), limit=10, state=state)

    for _ in range(4):
        one()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(stream)
    for _ in range(20):
        one()
    e1.record(stream); dist.barrier(); torch.cuda.synchronize()
    print(f"[dist trace] rank {rank}: untraced 20 steps: events {e0.elapsed_time(e1) / 20:.3f} ms/step, wall {(time.perf_counter() - t0) * 50:.3f} ms/step", flush=True)
    walls = []
    for _ in range(10):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter(); one(); torch.cuda.synchronize(); walls.append((time.perf_counter() - t0) * 1e3)
    print(f"[dist trace] rank {rank}: single steps after a barrier: {sorted(walls)[0]:.3f} .. {sorted(walls)[-1]:.3f} ms", flush=True)
    for it in range(9):
        if it == 4:
            os.environ["SQLRS_B200_DIST_TRACE"] = "1"
        if it == 6:  # the host's timeline of an undisturbed step: where it waits, what it spends between the waits
            os.environ["SQLRS_B200_DIST_TRACE"] = "host"
            if rank == 0:
                print("[dist trace] ---- host timeline (no added synchronisation)", flush=True)
        dist.barrier()
        sqdist.distributed_join_topk(builder, group, build_plan=cust, build_schemas={0: schemas[0]}, build_tables={0: tabs[0]}, query_plan=full, query_schemas=schemas,
                                     query_tables={1: tabs[1], 2: tabs[2]}, build_slot=0, order_by=tpch.q3_tail_order_by(), limit=10, state=state)
dist.barrier()
dist.destroy_process_group()
