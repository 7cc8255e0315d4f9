"""Per-phase wall-clock breakdown of the multi-GPU Q1' step (run under torchrun)."""
import ctypes as C
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from sqlrs_b200.host import distributed as sqdist
from sqlrs_b200.host import ffi, tpch
from sqlrs_b200.host.plan import ExecutorBuilder

sf = float(sys.argv[1]) if len(sys.argv) > 1 else 25
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = ffi.load()
d = tpch.dims(sf)
n = tpch.num_rows(lib, d, tpch.LINEITEM)
lo, hi = n * rank // world, n * (rank + 1) // world
stream = torch.cuda.Stream(device=dev)
plan_root, schemas = tpch.q1_plan()
T = {}
def tick(name, t0):
    torch.cuda.synchronize(dev)
    T[name] = T.get(name, 0.0) + (time.perf_counter() - t0)
    return time.perf_counter()
with torch.cuda.stream(stream):
    table = tpch.device_table(lib, d, tpch.LINEITEM, lo, hi, columns=tpch.Q1_COLUMNS, device=dev)
    opts = lib.options(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY, device_id=local, flags=ffi.FLAG_TIMING,
                       stream=C.c_void_p(stream.cuda_stream))
    plan = ExecutorBuilder(lib, opts).build(plan_root, schemas)
    plan.push_table_device(0, table)
    group = sqdist.TorchGroup(dist, dev)
    cap = 256
    for it in range(25):
        if it == 5:
            T.clear()
        dist.barrier(); torch.cuda.synchronize(dev)
        t = time.perf_counter(); t_all = t
        lib.check(lib.plan_execute_partial(plan.handle, lo)); t = tick("execute_partial", t)
        words = C.c_int32(0); lib.check(lib.plan_partials_row_words(plan.handle, C.byref(words)))
        nw = (cap + 1) * words.value
        if it == 0:
            send = torch.empty(nw, dtype=torch.int64, device=dev); recv = torch.empty(nw * world, dtype=torch.int64, device=dev)
        lib.check(lib.plan_export_partials_device(plan.handle, C.c_void_p(send.data_ptr()), cap)); t = tick("export", t)
        dist.all_gather_into_tensor(recv, send); t = tick("all_gather", t)
        mx = int(recv.view(world, cap + 1, words.value)[:, 0, 0].max().item()); t = tick("max.item", t)
        lib.check(lib.plan_clear_partials(plan.handle)); t = tick("clear", t)
        if rank == 0:
            lib.check(lib.plan_merge_partials_device(plan.handle, C.c_void_p(recv.data_ptr()), world, cap)); t = tick("merge", t)
        lib.check(lib.plan_finish_partial(plan.handle)); t = tick("finish", t)
        res = plan.collect(); t = tick("collect", t)
        T["total"] = T.get("total", 0.0) + (time.perf_counter() - t_all)
        ms, nl = plan.scan_kernel_ms(); T["kernel(events)"] = T.get("kernel(events)", 0.0) + ms / 1e3
    # the same through the library function, unsynchronised phases
    dist.barrier(); torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for it in range(20):
        res = sqdist.sharded_aggregate(plan, group, lo)
    dist.barrier(); torch.cuda.synchronize(dev)
    T["sharded_aggregate()"] = (time.perf_counter() - t0)
for r in range(world):
    dist.barrier()
    if r == rank and rank in (0, world - 1):
        print(f"rank {rank}: " + "  ".join(f"{k}={v / 20 * 1e3:.3f}ms" for k, v in T.items()), flush=True)
dist.destroy_process_group()
