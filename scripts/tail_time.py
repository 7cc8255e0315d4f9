"""Host + device cost of the small plans the multi-GPU Q3' step runs around the query itself (one GPU is enough):
the top-k tail over W x limit candidate rows, and the build-side sub-plan Filter(Scan(customer shard))."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sqlrs_b200.host import distributed as sqdist
from sqlrs_b200.host import ffi, tpch
from sqlrs_b200.host.plan import ExecutorBuilder, PhysicalLimit, PhysicalOrder, PhysicalTableScan

lib = ffi.load()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    opts = lib.options(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY, stream=C.c_void_p(stream.cuda_stream))
    builder = ExecutorBuilder(lib, opts)
    full, schemas = tpch.q3_full_plan()
    schema = full.output_schema(schemas)
    W, limit = 8, 10
    cand = torch.randint(0, 1 << 30, (len(schema), W * limit), dtype=torch.int64, device=dev)
    tail = builder.build(PhysicalLimit(limit, None, PhysicalOrder(tpch.q3_tail_order_by(), PhysicalTableScan(0))), {0: schema})
    tail.push_table_device(0, sqdist.DeviceBatch(schema, [cand[c] for c in range(len(schema))], W * limit, 0))

    def timeit(label, fn, reps=200):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        print(f"{label}: {(time.perf_counter() - t0) / reps * 1e6:.1f} us per call", flush=True)

    timeit("tail plan run (Limit(Order(Scan)) over 80 rows, result to host)", lambda: tail.run())
    timeit("tail plan execute only", lambda: (tail.execute(), tail.collect()))
    l0 = lib.kernel_launches(); tail.run(); print("launches per tail run:", lib.kernel_launches() - l0, tail.describe())
    d = tpch.dims(100)
    n_c = tpch.num_rows(lib, d, tpch.CUSTOMER)
    shard = tpch.device_table(lib, d, tpch.CUSTOMER, 0, n_c // 8, columns=tpch.Q3_CUSTOMER_COLUMNS, device=dev)
    cust = full.child.child.child.child.left.left
    pb = builder.build(cust, {0: schemas[0]})
    pb.push_table_device(0, shard)
    out = torch.empty((2, n_c // 8), dtype=torch.int64, device=dev)

    def build_side():
        pb.execute()
        shape = pb.result_shape()
        pb.next_to_device([out[c].data_ptr() for c in range(2)])
        return shape

    timeit("build-side sub-plan (Filter(Scan) over 1/8 of customer) + next_to_device", build_side)
    l0 = lib.kernel_launches(); build_side(); print("launches:", lib.kernel_launches() - l0, pb.describe())
    os.environ["SQLRS_B200_TRACE"] = "1"
