"""Device time (CUDA events, SQLRS_FLAG_KERNEL_EVENTS) of the fused probe+aggregate kernel of Q3' for one TMA ring shape
(environment: SQLRS_B200_TMA=1, SQLRS_B200_TMA_TROWS/_STAGES/_CONSUMERS) or the register-staged default (no SQLRS_B200_TMA)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sqlrs_b200.host import ffi, tpch
from sqlrs_b200.host.plan import ExecutorBuilder
sf = float(sys.argv[1]) if len(sys.argv) > 1 else 100
lib = ffi.load()
d = tpch.dims(sf)
tabs = {0: tpch.device_table(lib, d, tpch.CUSTOMER, columns=tpch.Q3_CUSTOMER_COLUMNS), 1: tpch.device_table(lib, d, tpch.ORDERS, columns=tpch.Q3_ORDERS_COLUMNS),
        2: tpch.device_table(lib, d, tpch.LINEITEM, columns=tpch.Q3_LINEITEM_COLUMNS)}
plan, schemas = tpch.q3_full_plan()
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    opts = lib.options(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY, flags=ffi.FLAG_KERNEL_EVENTS, stream=C.c_void_p(stream.cuda_stream))
    p = ExecutorBuilder(lib, opts).build(plan, schemas)
    for k, t in tabs.items():
        p.push_table_device(k, t)
    for _ in range(3):
        p.execute(); p.collect()
    torch.cuda.synchronize(); p.kernel_events()
    for _ in range(5):
        p.execute(); p.collect()
    torch.cuda.synchronize()
    ev = p.kernel_events()
    cfg = {k: os.environ.get(k) for k in ("SQLRS_B200_TMA", "SQLRS_B200_TMA_TROWS", "SQLRS_B200_TMA_STAGES", "SQLRS_B200_TMA_CONSUMERS")}
    print(f"SF{sf:g} {cfg}: " + ", ".join(f"{k} {v['ms'] / 5:.3f} ms" for k, v in ev.items() if "joinagg" in k), flush=True)
