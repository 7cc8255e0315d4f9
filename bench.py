#!/usr/bin/env python
"""bench.py — TPC-H-shaped Q1' (Filter + 8-group HashAgg over lineitem) through the sqlrs_b200 C ABI.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port, 1 core)

A "step" is one pass of the hot path (fused scan+filter+group-by+aggregate) over the whole synthetic
lineitem table of scale factor --sf (default 100: the size BASELINE.json's metric is quoted on; it fits
one B200).  With N GPUs the rows are split into N contiguous shards (strong scaling: total work fixed),
each rank aggregates its shard, partial groups are exchanged by key hash (all-to-all) and merged.
`value` = rows/s with inputs resident in HBM; `e2e` = the same call with HOST (pinned) Arrow buffers,
H2D copies inside the timed region.  One JSON line on stdout (rank 0).

    python bench.py --query q3 [--q3-sf 100]                  # secondary line: Q3' (2 hash joins + group-by), one GPU:
                                                              #   `value` = plan up to the aggregate, `full_query` = + ORDER BY / LIMIT on the device
    torchrun ... bench.py --gpus N --query q3 --q3-sf 100     # Q3' whole query on N GPUs, orders/lineitem co-partitioned on orderkey
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Native libraries print there too (NCCL announces its version on fd 1 when
# the first communicator comes up), so fd 1 is pointed at stderr for the whole run and the line goes to the saved fd.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def load_oracle():
    """CPU restatement of the reference's executor — used ONLY for cpu_baseline / --impl reference."""
    from sqlrs_b200.host import ffi

    path = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    return ffi.Library(path, "sqlrs_oracle_")


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


_SAMPLER_SRC = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
print("max", nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM), flush=True)
bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
        "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
while True:
    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
    print("s", time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), ",".join(k for k, b in bits.items() if r & b), flush=True)
    time.sleep(0.02)
"""


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region by a separate process (NVML every 20 ms), so
    that the sampling neither holds this process's GIL nor sits between its CUDA calls."""

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.proc.stdout.readline()  # "max ..." = the sampler is up
        except Exception:
            self.proc = None

    def stop(self, t_begin=0.0, t_end=1e300):
        """Summary of the samples taken in [t_begin, t_end] (time.time() of the timed region)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.03)
        self.proc.terminate()
        out = self.proc.stdout.read()
        sm, reasons = [], set()
        for line in out.splitlines():
            parts = line.split()
            if len(parts) >= 3 and parts[0] == "s" and t_begin <= float(parts[1]) <= t_end:
                sm.append(float(parts[2]))
                if len(parts) > 3:
                    reasons.update(x for x in parts[3].split(",") if x)
        mx = None
        try:
            import pynvml as nv

            nv.nvmlInit()
            mx = float(nv.nvmlDeviceGetMaxClockInfo(nv.nvmlDeviceGetHandleByIndex(self.index), nv.NVML_CLOCK_SM))
        except Exception:
            pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def q1_on_host_batches(lib, plan_root, schemas, table, batch_rows, opts):
    """One Q1' pass of a HOST build of the ABI over `table`, fed in batches like the reference's scan."""
    import pyarrow as pa

    from sqlrs_b200.host.plan import ExecutorBuilder

    p = ExecutorBuilder(lib, lib.options(**opts)).build(plan_root, schemas)
    t0 = time.perf_counter()
    for off in range(0, table.num_rows, batch_rows):
        p.push_table(0, table.slice(off, batch_rows))
    res = pa.Table.from_batches(p.run())
    dt = time.perf_counter() - t0
    p.close()
    return dt, res


def cpu_port_run(sample_rows, batch_rows, steps, warmup, sf, kind_note=""):
    """Times the oracle port (1 core, like the reference's single-threaded executor) on a bounded sample."""
    from sqlrs_b200.host import ffi, tpch

    oracle = load_oracle()
    d = tpch.dims(sf)
    n_total = tpch.num_rows(oracle, d, tpch.LINEITEM)
    n = min(sample_rows, n_total)
    table = tpch.host_table(oracle, d, tpch.LINEITEM, 0, n, columns=tpch.Q1_COLUMNS)
    plan_root, schemas = tpch.q1_plan()
    opts = dict(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
    try:
        os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[0]})
    except Exception:
        pass
    times = []
    for k in range(warmup + steps):
        dt, _ = q1_on_host_batches(oracle, plan_root, schemas, table, batch_rows, opts)
        if k >= warmup:
            times.append(dt)
    total = sum(times)
    return {"rows_per_s": n * len(times) / total, "ms_per_step": 1e3 * total / len(times), "rows": n, "batch_rows": batch_rows}


def run_reference(args, rank, world):
    if rank != 0:
        return
    sample = args.ref_rows
    r = cpu_port_run(sample, 1024, args.steps, max(1, min(args.warmup, 1)), args.sf)
    line = {
        "impl": "reference", "metric": "tpch_q1_rows_per_sec", "value": r["rows_per_s"], "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64+i64", "data": "synthetic",
        "config": {"workload": f"tpch_q1_sf{args.sf:g}", "note": "reference CPU path = C++ restatement of sqlrs v1 Filter+HashAgg (the Rust crate needs nightly-2022-07-29, "
                   "unbuildable here); single-threaded like the reference's executor; 1024-row batches (src/storage/csv.rs:105)"},
        "cpu_baseline": {"value": r["rows_per_s"], "unit": "rows/s", "cores": 1, "kind": "port",
                         "sample": f"first {r['rows']} lineitem rows of SF{args.sf:g} per step, batch {r['batch_rows']} rows"},
        "e2e": {"value": r["rows_per_s"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_gpu(args, rank, world, local_rank):
    import pyarrow as pa
    import torch

    from sqlrs_b200.host import distributed as sqdist
    from sqlrs_b200.host import ffi, tpch
    from sqlrs_b200.host.plan import ExecutorBuilder

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    lib = ffi.load()
    d = tpch.dims(args.sf)
    n_total = tpch.num_rows(lib, d, tpch.LINEITEM)
    lo, hi = n_total * rank // world, n_total * (rank + 1) // world
    n_local = hi - lo
    stream = torch.cuda.Stream(device=dev)
    plan_root, schemas = tpch.q1_plan()
    bytes_per_row = tpch.q1_bytes_per_row()

    with torch.cuda.stream(stream):
        table = tpch.device_table(lib, d, tpch.LINEITEM, lo, hi, columns=tpch.Q1_COLUMNS, device=dev)
        opts = lib.options(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY, device_id=local_rank,
                           flags=ffi.FLAG_TIMING, stream=C.c_void_p(stream.cuda_stream))
        plan = ExecutorBuilder(lib, opts).build(plan_root, schemas)
        kernel_ms, kernel_launches = [0.0], [0]
        group = sqdist.TorchGroup(dist, dev) if world > 1 else None

        plan.push_table_device(0, table)  # zero-copy: the plan scans the table where it lies in HBM

        def step(timed):
            if world > 1:
                res = sqdist.sharded_aggregate(plan, group, lo)
            else:
                plan.execute()
                res = plan.collect()
            if timed:
                ms, nl = plan.scan_kernel_ms()
                kernel_ms[0] += ms
                kernel_launches[0] += nl
            return res

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)

        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()  # before the warm-up: NVML start-up stays out of the timed region
        for _ in range(args.warmup):
            result = step(False)
        barrier()
        launches0 = lib.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin = time.time()
        e0.record(stream)
        step_wall = []
        for _ in range(args.steps):
            t_s = time.perf_counter()
            result = step(True)
            step_wall.append((time.perf_counter() - t_s) * 1e3)
        e1.record(stream)
        barrier()
        t_end = time.time()
        launches = lib.kernel_launches() - launches0
        elapsed_ms = e0.elapsed_time(e1)
        clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
        step_wall.sort()
        log(f"rank {rank}: step wall ms min/median/max = {step_wall[0]:.3f}/{step_wall[len(step_wall) // 2]:.3f}/{step_wall[-1]:.3f}")
        if world > 1:
            t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            elapsed_ms = float(t.item())
        describe = plan.describe()

        # ---- e2e: the same call with HOST (pinned) Arrow buffers; H2D inside the timed region
        e2e = None
        if args.e2e_steps > 0:
            host_cols = []
            for t_dev in table.tensors:
                h = torch.empty(t_dev.shape, dtype=t_dev.dtype, pin_memory=True)
                h.copy_(t_dev)
                host_cols.append(h)
            torch.cuda.synchronize(dev)
            arrays = []
            for h, f in zip(host_cols, table.schema):
                a = h.numpy()
                if pa.types.is_floating(f.type):
                    a = a.view("float64")
                arrays.append(pa.array(a))
            host_batch = pa.RecordBatch.from_arrays(arrays, schema=table.schema)

            plan.reset()

            def e2e_step():
                plan.push_table(0, host_batch)
                if world > 1:
                    res = sqdist.sharded_aggregate(plan, group, lo)
                else:
                    plan.execute()
                    res = plan.collect()
                plan.reset()
                return res

            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                e2e_result = e2e_step()
            barrier()
            e2e_s = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                e2e_s = float(t.item())
            out_bytes = sum(b.nbytes for b in e2e_result) if e2e_result else 0
            e2e = {"value": n_total * args.e2e_steps / e2e_s, "unit": "rows/s", "h2d_bytes_per_step": n_local * bytes_per_row,
                   "d2h_bytes_per_step": int(out_bytes), "steps": args.e2e_steps, "ms_per_step": 1e3 * e2e_s / args.e2e_steps,
                   "note": "pinned host Arrow buffers -> sqlrs_plan_push_table (H2D) -> execute -> result to host, per rank shard"}
            del host_batch, arrays, host_cols
        plan.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()

    if rank != 0:
        return
    peak, peak_src = measured_peak()
    traffic = None
    try:  # DRAM bytes of the dominant kernel from the committed ncu --set full capture of this workload (per launch)
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f).get(f"tpch_q1_sf{args.sf:g}")
        if t and world == 1:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        pass
    ms_per_step = elapsed_ms / args.steps
    value = n_total / (ms_per_step * 1e-3)
    k_ms = kernel_ms[0] / max(kernel_launches[0], 1)
    achieved = (n_local * bytes_per_row) / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
    line = {
        "metric": "tpch_q1_rows_per_sec", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64+i64",
        "data": "synthetic",
        "config": {"workload": f"tpch_q1_sf{args.sf:g}", "rows": n_total, "rows_per_gpu": n_local, "columns_read": len(tpch.Q1_COLUMNS),
                   "bytes_per_row": bytes_per_row, "groups": len(result[0]) if result else None, "count_mode": "sql_accumulate",
                   "match_mode": "hash_and_key", "l2": "inputs (%.1f GB per GPU) are larger than the 126 MB L2" % (n_local * bytes_per_row / 1e9),
                   "pipeline": describe},
        "hbm_gbs_whole_step": n_total * bytes_per_row / (ms_per_step * 1e-3) / 1e9,
        "roofline": {"bound": "hbm", "kernel": "sq_agg_small", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if achieved else None, "frac_of_8TBs": achieved / 8000.0 if achieved else None, "peak_source": peak_src,
                     "kernel_ms": k_ms, "algorithmic_bytes_per_launch": n_local * bytes_per_row, "traffic": traffic},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    if world == 1 and args.cpu_rows > 0:
        r = cpu_port_run(args.cpu_rows, 1024, 1, 0, args.sf)
        line["cpu_baseline"] = {"value": r["rows_per_s"], "unit": "rows/s", "cores": 1, "kind": "port",
                                "sample": f"first {r['rows']} lineitem rows of SF{args.sf:g}, one pass, batch 1024 rows, C++ restatement of the sqlrs v1 executor"}
    emit(line)


def q3_tables(lib, d, make, cols=None):
    from sqlrs_b200.host import tpch

    return {0: make(lib, d, tpch.CUSTOMER, columns=tpch.Q3_CUSTOMER_COLUMNS), 1: make(lib, d, tpch.ORDERS, columns=tpch.Q3_ORDERS_COLUMNS),
            2: make(lib, d, tpch.LINEITEM, columns=tpch.Q3_LINEITEM_COLUMNS)}


def run_gpu_q3(args):
    """Secondary line (not the driver's default): Q3' = customer ⋈ orders ⋈ lineitem -> group-by, one GPU, tables in HBM."""
    import pyarrow as pa
    import torch

    from sqlrs_b200.host import ffi, tpch
    from sqlrs_b200.host.plan import ExecutorBuilder

    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    lib = ffi.load()
    d = tpch.dims(args.sf)
    stream = torch.cuda.Stream(device=dev)
    plan_root, schemas = tpch.q3_plan()
    mode = dict(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
    with torch.cuda.stream(stream):
        tabs = q3_tables(lib, d, lambda l, dd, t, columns: tpch.device_table(l, dd, t, columns=columns, device=dev))
        rows = {k: t.n_rows for k, t in tabs.items()}
        alg = tpch.q3_algorithmic_bytes(rows[0], rows[1], rows[2])
        plan = ExecutorBuilder(lib, lib.options(device_id=0, stream=C.c_void_p(stream.cuda_stream), **mode)).build(plan_root, schemas)
        for k, t in tabs.items():
            plan.push_table_device(k, t)

        def step():
            plan.execute()
            return plan.collect()

        sampler = ClockSampler(0)
        sampler.start()
        for _ in range(args.warmup):
            result = step()
        torch.cuda.synchronize(dev)
        l0 = lib.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin = time.time()
        e0.record(stream)
        for _ in range(args.steps):
            result = step()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        t_end = time.time()
        ms = e0.elapsed_time(e1) / args.steps
        launches = lib.kernel_launches() - l0
        clocks = sampler.stop(t_begin, t_end)
        describe = plan.describe()
        groups = sum(b.num_rows for b in result)
        plan.close()

        # the whole query on the device: + ORDER BY revenue desc, o_orderdate LIMIT 10 and the select list
        # (Order / Project / Limit nodes) — only the 10 final rows cross PCIe
        full_root, _ = tpch.q3_full_plan()
        fplan = ExecutorBuilder(lib, lib.options(device_id=0, stream=C.c_void_p(stream.cuda_stream), **mode)).build(full_root, schemas)
        for k, t in tabs.items():
            fplan.push_table_device(k, t)
        for _ in range(args.warmup):
            fplan.execute()
            top = fplan.collect()
        torch.cuda.synchronize(dev)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(args.steps):
            fplan.execute()
            top = fplan.collect()
        f1.record(stream)
        torch.cuda.synchronize(dev)
        full_ms = f0.elapsed_time(f1) / args.steps
        full_describe = fplan.describe()
        # the device top-10 must equal the top-10 of the aggregate output the first plan returned
        agg_tab = pa.Table.from_batches(result)
        names = agg_tab.schema.names
        want = agg_tab.sort_by([(names[3], "descending"), (names[1], "ascending")]).slice(0, 10)
        got = pa.Table.from_batches(top)
        top_ok = got.num_rows == want.num_rows and got.column(0).to_pylist() == want.column(0).to_pylist()
        fplan.close()
    peak, peak_src = measured_peak()
    n_in = sum(rows.values())
    line = {
        "metric": "tpch_q3_rows_per_sec", "value": n_in / (ms * 1e-3), "unit": "rows/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "i64+f64", "data": "synthetic",
        "config": {"workload": f"tpch_q3_sf{args.sf:g}", "rows": rows, "groups": groups, "count_mode": "sql_accumulate", "match_mode": "hash_and_key",
                   "l2": "lineitem columns (%.1f GB) are larger than the 126 MB L2" % (rows[2] * 32 / 1e9), "pipeline": describe},
        "roofline": {"bound": "hbm", "kernel": "whole pipeline (2 x join build/probe + aggregate; no single dominant kernel)", "achieved": alg / (ms * 1e-3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg, "traffic": None},
        "full_query": {"ms_per_step": full_ms, "rows_per_s": n_in / (full_ms * 1e-3), "result_rows": got.num_rows, "top10_matches_aggregate_output": bool(top_ok),
                       "note": "same plan + Order(revenue desc, o_orderdate) + Project + Limit 10 on the device", "pipeline": full_describe},
        "e2e": None, "gpu_launches": int(launches), "clocks": clocks,
    }
    if args.cpu_rows > 0:  # CPU port on a bounded sample: the same plan at a smaller scale factor
        oracle = load_oracle()
        dc = tpch.dims(args.cpu_sf)
        host = q3_tables(oracle, dc, lambda l, dd, t, columns: tpch.host_table(l, dd, t, columns=columns))
        p = ExecutorBuilder(oracle, oracle.options(**mode)).build(plan_root, schemas)
        t0 = time.perf_counter()
        for slot, t in host.items():
            for off in range(0, t.num_rows, 1024):
                p.push_table(slot, t.slice(off, 1024))
        p.run()
        dt = time.perf_counter() - t0
        p.close()
        n_cpu = sum(t.num_rows for t in host.values())
        line["cpu_baseline"] = {"value": n_cpu / dt, "unit": "rows/s", "cores": 1, "kind": "port",
                                "sample": f"Q3' at SF{args.cpu_sf:g} ({n_cpu} input rows), one pass, batch 1024 rows, C++ restatement of the sqlrs v1 executor"}
    emit(line)


def run_gpu_q3_multi(args, rank, world, local_rank):
    """Q3' (whole query, tail included) on N GPUs: orders and lineitem range-partitioned on orderkey (co-partitioned: scan
    and both joins stay GPU-local), customer replicated; every rank runs the full plan on its shards, the per-rank top-10
    rows are gathered and ordered again on rank 0 (sqlrs_b200/host/distributed.py: copartitioned_topk).  Strong scaling."""
    import torch
    import torch.distributed as dist

    from sqlrs_b200.host import distributed as sqdist
    from sqlrs_b200.host import ffi, tpch
    from sqlrs_b200.host.plan import ExecutorBuilder

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    lib = ffi.load()
    d = tpch.dims(args.sf)
    stream = torch.cuda.Stream(device=dev)
    plan_root, schemas = tpch.q3_full_plan()
    mode = dict(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
    (o_lo, o_hi), (l_lo, l_hi) = sqdist.copartitioned_shard(int(d.n_orders), rank, world)
    n_c, n_o, n_l = (tpch.num_rows(lib, d, t) for t in (tpch.CUSTOMER, tpch.ORDERS, tpch.LINEITEM))
    with torch.cuda.stream(stream):
        tabs = {0: tpch.device_table(lib, d, tpch.CUSTOMER, columns=tpch.Q3_CUSTOMER_COLUMNS, device=dev),
                1: tpch.device_table(lib, d, tpch.ORDERS, o_lo, o_hi, columns=tpch.Q3_ORDERS_COLUMNS, device=dev),
                2: tpch.device_table(lib, d, tpch.LINEITEM, l_lo, l_hi, columns=tpch.Q3_LINEITEM_COLUMNS, device=dev)}
        plan = ExecutorBuilder(lib, lib.options(device_id=local_rank, stream=C.c_void_p(stream.cuda_stream), **mode)).build(plan_root, schemas)
        for k, t in tabs.items():
            plan.push_table_device(k, t)
        group = sqdist.TorchGroup(dist, dev)
        order_by = tpch.q3_tail_order_by()

        def step():
            return sqdist.copartitioned_topk(plan, group, order_by, 10)

        def barrier():
            dist.barrier()
            torch.cuda.synchronize(dev)

        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        for _ in range(args.warmup):
            result = step()
        barrier()
        l0 = lib.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin = time.time()
        e0.record(stream)
        for _ in range(args.steps):
            result = step()
        e1.record(stream)
        barrier()
        t_end = time.time()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / args.steps
        launches = lib.kernel_launches() - l0
        clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
        describe = plan.describe()
        plan.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank != 0:
        return
    import pyarrow as pa

    peak, peak_src = measured_peak()
    n_in = n_c + n_o + n_l
    alg = tpch.q3_algorithmic_bytes(n_c, n_o, n_l)
    top = pa.Table.from_batches(result)
    line = {
        "metric": "tpch_q3_rows_per_sec", "value": n_in / (ms * 1e-3), "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "i64+f64", "data": "synthetic",
        "config": {"workload": f"tpch_q3_sf{args.sf:g}", "rows": {"0": n_c, "1": n_o, "2": n_l}, "query": "whole query incl. ORDER BY revenue desc, o_orderdate LIMIT 10",
                   "partitioning": "orders + lineitem range-partitioned on orderkey (co-partitioned), customer replicated; no data-path collective, "
                                   "per-rank top-10 gathered and re-ordered on rank 0",
                   "count_mode": "sql_accumulate", "match_mode": "hash_and_key", "result_rows": top.num_rows,
                   "top_row": {k: v[0] for k, v in top.to_pydict().items()} if top.num_rows else None,
                   "l2": "lineitem columns (%.1f GB per GPU) are larger than the 126 MB L2" % (n_l * 32 / world / 1e9), "pipeline": describe},
        "roofline": {"bound": "hbm", "kernel": "whole pipeline (2 x join build/probe + aggregate + order/limit; no single dominant kernel)",
                     "achieved": alg / (ms * 1e-3) / 1e9 / world, "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / world / peak,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg // world, "traffic": None, "note": "per GPU"},
        "e2e": None, "gpu_launches": int(launches), "clocks": clocks,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sf", type=float, default=100.0)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-rows", type=int, default=48_000_000, help="rows of the cpu_baseline sample (0 = skip)")
    ap.add_argument("--ref-rows", type=int, default=12_000_000, help="rows per step of --impl reference")
    ap.add_argument("--query", default="q1", choices=["q1", "q3"], help="q1 = the driver's workload; q3 = secondary line (1 GPU)")
    ap.add_argument("--q3-sf", type=float, default=10.0, help="scale factor of --query q3")
    ap.add_argument("--cpu-sf", type=float, default=1.0, help="scale factor of the q3 cpu_baseline sample")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus:
        log(f"note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    if args.query == "q3":
        args.sf = args.q3_sf  # BASELINE.json configs[2]: Q3 SF10 on one GPU (configs[4]: --q3-sf 100, N GPUs under torchrun)
        if world > 1:
            run_gpu_q3_multi(args, rank, world, local_rank)
        elif rank == 0:
            run_gpu_q3(args)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
