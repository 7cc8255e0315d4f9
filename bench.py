#!/usr/bin/env python
"""bench.py — TPC-H-shaped Q1' and Q3' through the sqlrs_b200 C ABI (BASELINE.json: "TPC-H Q1/Q3 SF100 @1/2/4/8 B200 vs CPU ref").

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port, 1 core)

ONE JSON line on stdout (rank 0).  Headline = Q1' at SF100 (the configuration the metric is quoted on; 38.4 GB, fits one
B200): a "step" is one pass of the fused scan+filter+group-by+aggregate over the whole synthetic lineitem table; with N
GPUs the rows are split into N contiguous shards (strong scaling), partial groups are exchanged over NCCL and merged.
`value` = rows/s with inputs resident in HBM; `e2e` = the same call with HOST (pinned) Arrow buffers, H2D copies inside the
timed region; `roofline` = the dominant kernel against the measured HBM peak; `parity_check` = what the timed plan returned,
checked against reductions over the resident columns and against the CPU oracle on a bounded sample.

The same line carries `q3`: Q3' (customer join orders join lineitem -> group-by -> ORDER BY revenue desc, o_orderdate LIMIT 10, whole
query on the device) at SF10 and SF100 on one GPU, and at SF100 on the N GPUs of a torchrun launch (fact tables sharded by
orderkey ranges, filtered customer rows all-gathered over NVLink), each with ms/step, rows/s, a per-kernel roofline from
CUDA events (SQLRS_FLAG_KERNEL_EVENTS), the CPU port beside it, e2e with host buffers and its own parity check.
`--q3 off` skips it; `--query q3` prints the Q3' object as the line's headline instead.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
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
    """CPU restatement of the reference's executor — used ONLY as the checker / cpu_baseline / --impl reference."""
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


def source_fingerprint(query):
    """Hash of the complete CUDA sources (prelude + generated row program + skeletons) of the kernels the benchmark plan of
    `query` ("q1" | "q3") runs: `roofline.traffic` comes from a committed ncu capture and is only reported while the kernels are
    the ones that were captured (profiles/r02_traffic.json carries the fingerprints of its captures, per query)."""
    if not _FINGERPRINTS:
        import __graft_entry__ as entry
        from sqlrs_b200.host import ffi

        sources = entry.benchmark_kernel_sources(ffi.load(), compile=False)  # [Q1' aggregate x 2 modes, Q3' chain, probe / build, probe + aggregate]
        for name, part in (("q1", sources[:2]), ("q3", sources[2:])):
            h = hashlib.sha1()
            for src in part:
                h.update(src.encode())
            _FINGERPRINTS[name] = h.hexdigest()[:16]
    return _FINGERPRINTS[query]


_FINGERPRINTS = {}


def ncu_traffic(workload, kernel):
    """DRAM bytes (read + write) of one launch of `kernel` in `workload` from this round's ncu --set full capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            t = json.load(f)
        query = "q1" if "_q1_" in workload else "q3"
        if t.get("source_fingerprints", {}).get(query) != source_fingerprint(query):
            return None
        k = t["workloads"][workload][kernel]
        return int(k["dram_bytes_read"] + k["dram_bytes_write"])
    except Exception:
        return None


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

    def stop(self, windows):
        """Summary of the samples taken inside the [t_begin, t_end] windows (time.time()) of the timed regions."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.03)
        self.proc.terminate()
        out = self.proc.stdout.read()
        sm, reasons = [], set()
        for line in out.splitlines():
            parts = line.split()
            if len(parts) >= 3 and parts[0] == "s" and any(a <= float(parts[1]) <= b for a, b in windows):
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


MODE = None  # filled in main(): dict(count_mode=SQL_ACCUMULATE, match_mode=HASH_AND_KEY)


class pin_core:
    """the CPU port runs on ONE core (the reference's executor is single-threaded per query); the affinity is restored afterwards"""

    def __enter__(self):
        try:
            self.prev = os.sched_getaffinity(0)
            os.sched_setaffinity(0, {sorted(self.prev)[0]})
        except Exception:
            self.prev = None

    def __exit__(self, *exc):
        if self.prev:
            try:
                os.sched_setaffinity(0, self.prev)
            except Exception:
                pass


# ------------------------------------------------------------------------------------------------ CPU port (oracle)
def cpu_port_q1(oracle, sf, sample_rows, batch_rows, repeats=1):
    """The oracle port on ONE core over the first `sample_rows` lineitem rows, handed to its executor in `batch_rows`-row
    batches by the library itself (sqlrs_plan_push_table_batched: no Python in the timed loop).  Timed: execute + collect."""
    import pyarrow as pa

    from sqlrs_b200.host import tpch
    from sqlrs_b200.host.plan import ExecutorBuilder

    d = tpch.dims(sf)
    n = min(sample_rows, tpch.num_rows(oracle, d, tpch.LINEITEM))
    table = tpch.host_table(oracle, d, tpch.LINEITEM, 0, n, columns=tpch.Q1_COLUMNS)
    plan_root, schemas = tpch.q1_plan()
    best, res = None, None
    with pin_core():
        for _ in range(repeats):
            p = ExecutorBuilder(oracle, oracle.options(**MODE)).build(plan_root, schemas)
            p.push_table_batched(0, table, batch_rows)
            t0 = time.perf_counter()
            res = pa.Table.from_batches(p.run())
            dt = time.perf_counter() - t0
            p.close()
            best = dt if best is None else min(best, dt)
    return {"rows_per_s": n / best, "seconds": best, "rows": n, "batch_rows": batch_rows}, res


def cpu_port_q3(oracle, sf, batch_rows=1024):
    import pyarrow as pa

    from sqlrs_b200.host import tpch
    from sqlrs_b200.host.plan import ExecutorBuilder

    d = tpch.dims(sf)
    host = {0: tpch.host_table(oracle, d, tpch.CUSTOMER, columns=tpch.Q3_CUSTOMER_COLUMNS), 1: tpch.host_table(oracle, d, tpch.ORDERS, columns=tpch.Q3_ORDERS_COLUMNS),
            2: tpch.host_table(oracle, d, tpch.LINEITEM, columns=tpch.Q3_LINEITEM_COLUMNS)}
    plan_root, schemas = tpch.q3_full_plan()
    with pin_core():
        p = ExecutorBuilder(oracle, oracle.options(**MODE)).build(plan_root, schemas)
        for slot, t in host.items():
            p.push_table_batched(slot, t, batch_rows)
        t0 = time.perf_counter()
        res = pa.Table.from_batches(p.run())
        dt = time.perf_counter() - t0
        p.close()
    n = sum(t.num_rows for t in host.values())
    return {"rows_per_s": n / dt, "seconds": dt, "rows": n, "batch_rows": batch_rows, "sf": sf}, res


def tables_equal(got, want, rtol=1e-9):
    """integer columns and row order exact, Float64 within rtol"""
    import pyarrow as pa

    if got.schema.names != want.schema.names or got.num_rows != want.num_rows:
        return False
    for name in got.schema.names:
        g, w = got.column(name).to_pylist(), want.column(name).to_pylist()
        if pa.types.is_floating(got.schema.field(name).type):
            if not all((a is None and b is None) or (a is not None and b is not None and abs(a - b) <= rtol * max(abs(a), abs(b), 1e-300)) for a, b in zip(g, w)):
                return False
        elif g != w:
            return False
    return True


def q1_config(sf, n_total, n_local=None):
    from sqlrs_b200.host import tpch

    return {"workload": f"tpch_q1_sf{sf:g}", "rows": n_total, "rows_per_gpu": n_local if n_local is not None else n_total, "columns_read": len(tpch.Q1_COLUMNS),
            "bytes_per_row": tpch.q1_bytes_per_row(), "count_mode": "sql_accumulate", "match_mode": "hash_and_key"}


def run_reference(args, rank, world):
    """The reference's CPU path: the C++ restatement of sqlrs v1 Filter+HashAgg (the Rust crate needs nightly-2022-07-29 and is
    unbuildable here), single-threaded like the reference's executor, 1024-row batches (src/storage/csv.rs:105)."""
    if rank != 0:
        return
    from sqlrs_b200.host import tpch

    oracle = load_oracle()
    n_total = tpch.num_rows(oracle, tpch.dims(args.sf), tpch.LINEITEM)
    times = []
    for k in range(max(1, min(args.warmup, 1)) + args.steps):
        r, _ = cpu_port_q1(oracle, args.sf, args.ref_rows, 1024)
        if k >= max(1, min(args.warmup, 1)):
            times.append(r["seconds"])
    rows = r["rows"]
    value = rows * len(times) / sum(times)
    r64, _ = cpu_port_q1(oracle, args.sf, args.ref_rows, 65536)
    cfg = q1_config(args.sf, n_total)
    cfg.update({"groups": 8, "sample_rows_per_step": rows,
                "note": "reference CPU path = C++ restatement of sqlrs v1 Filter+HashAgg, single-threaded like the reference's executor, 1024-row batches "
                        "(src/storage/csv.rs:105); each step = the first sample_rows_per_step rows of the workload (rows/s is scale-free for an 8-group aggregation)"})
    line = {
        "impl": "reference", "metric": "tpch_q1_rows_per_sec", "value": value, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64+i64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": 1, "kind": "port", "sample": f"first {rows} lineitem rows of SF{args.sf:g} per step, batch 1024 rows",
                         "value_batch_65536": r64["rows_per_s"]},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.q3 != "off":
        q3, _ = cpu_port_q3(oracle, args.cpu_sf)
        line["q3"] = {"metric": "tpch_q3_rows_per_sec", "value": q3["rows_per_s"], "unit": "rows/s", "cores": 1, "kind": "port",
                      "sample": f"Q3' whole query at SF{args.cpu_sf:g} ({q3['rows']} input rows), batch 1024 rows"}
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
class Ctx:
    pass


def barrier(ctx):
    if ctx.world > 1:
        ctx.dist.barrier()
    ctx.torch.cuda.synchronize(ctx.dev)


def max_over_ranks(ctx, x):
    if ctx.world == 1:
        return x
    t = ctx.torch.tensor([x], device=ctx.dev, dtype=ctx.torch.float64)
    ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(ctx, x):
    if ctx.world == 1:
        return int(x)
    t = ctx.torch.tensor([int(x)], device=ctx.dev, dtype=ctx.torch.int64)
    ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.SUM)
    return int(t.item())


def timed(ctx, step, steps, warmup):
    """W warm-up steps, then exactly K steps between two events on the plan's stream, barrier + synchronize on both sides,
    max over ranks.  Returns (ms per step, last result, [t_begin, t_end])."""
    torch = ctx.torch
    res = None
    for _ in range(warmup):
        res = step()
    barrier(ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    e0.record(ctx.stream)
    for _ in range(steps):
        res = step()
    e1.record(ctx.stream)
    barrier(ctx)
    t_end = time.time()
    return max_over_ranks(ctx, e0.elapsed_time(e1)) / steps, res, (t_begin, t_end)


def pinned_batch(ctx, table):
    """the device table's columns as a host Arrow batch over PINNED memory"""
    import pyarrow as pa

    torch = ctx.torch
    arrays, keep = [], []
    for t_dev, f in zip(table.tensors, table.schema):
        h = torch.empty(t_dev.shape, dtype=t_dev.dtype, pin_memory=True)
        h.copy_(t_dev)
        keep.append(h)
        a = h.numpy()
        arrays.append(pa.array(a.view("float64") if pa.types.is_floating(f.type) else a))
    torch.cuda.synchronize(ctx.dev)
    return pa.RecordBatch.from_arrays(arrays, schema=table.schema), keep


def run_gpu_q1(args, ctx):
    import pyarrow as pa

    from sqlrs_b200.host import distributed as sqdist
    from sqlrs_b200.host import ffi, tpch
    from sqlrs_b200.host.plan import ExecutorBuilder

    torch, lib, rank, world = ctx.torch, ctx.lib, ctx.rank, ctx.world
    d = tpch.dims(args.sf)
    n_total = tpch.num_rows(lib, d, tpch.LINEITEM)
    lo, hi = n_total * rank // world, n_total * (rank + 1) // world
    n_local = hi - lo
    plan_root, schemas = tpch.q1_plan()
    bytes_per_row = tpch.q1_bytes_per_row()
    table = tpch.device_table(lib, d, tpch.LINEITEM, lo, hi, columns=tpch.Q1_COLUMNS, device=ctx.dev)
    opts = lib.options(device_id=ctx.local_rank, flags=ffi.FLAG_TIMING, stream=C.c_void_p(ctx.stream.cuda_stream), **MODE)
    plan = ExecutorBuilder(lib, opts).build(plan_root, schemas)
    kernel_ms, kernel_launches = [0.0], [0]
    plan.push_table_device(0, table)  # zero-copy: the plan scans the table where it lies in HBM
    counting = [False]

    def step():
        if world > 1:
            res = sqdist.sharded_aggregate(plan, ctx.group, lo)
        else:
            plan.execute()
            res = plan.collect()
        if counting[0]:
            ms, nl = plan.scan_kernel_ms()
            kernel_ms[0] += ms
            kernel_launches[0] += nl
        return res

    for _ in range(args.warmup):
        step()
    counting[0] = True
    launches0 = lib.kernel_launches()
    ms_per_step, result, window = timed(ctx, step, args.steps, 0)
    launches = lib.kernel_launches() - launches0
    describe = plan.describe()

    # ---- parity: what the timed plan returned vs reductions over the resident columns (every N), and vs the oracle on a sample (N = 1)
    ship, qty = table.tensors[tpch.Q1_COLUMNS.index(7)], table.tensors[tpch.Q1_COLUMNS.index(8)]
    keep = ship <= tpch.Q1_SHIPDATE_MAX
    count_expected = sum_over_ranks(ctx, int(keep.sum().item()))
    qty_expected = sum_over_ranks(ctx, int(qty[keep].sum().item()))
    parity = None
    if rank == 0:
        got = pa.Table.from_batches(result).to_pydict()
        names = list(got)
        parity = {"groups": len(got[names[0]]), "count_sum": int(sum(got[names[-1]])), "count_expected": count_expected,
                  "sum_qty_i64": int(sum(got[names[-2]])), "sum_qty_i64_expected": qty_expected}
        parity["ok"] = parity["count_sum"] == count_expected and parity["sum_qty_i64"] == qty_expected and parity["groups"] == 8
    del keep

    # ---- e2e: the same call with HOST (pinned) Arrow buffers; H2D inside the timed region.  Runs AFTER every resident timing of
    # the process (main): pinning / unpinning tens of GB of host memory must not overlap another query's timed region.
    def run_e2e():
        e2e = None
        if args.e2e_steps > 0:
            host_batch, keep_alive = pinned_batch(ctx, table)
            plan.reset()

            def e2e_step():
                plan.push_table(0, host_batch)
                if world > 1:
                    res = sqdist.sharded_aggregate(plan, ctx.group, lo)
                else:
                    plan.execute()
                    res = plan.collect()
                plan.reset()
                return res

            e2e_step()
            barrier(ctx)
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                e2e_result = e2e_step()
            barrier(ctx)
            e2e_s = max_over_ranks(ctx, time.perf_counter() - t0)
            out_bytes = sum(b.nbytes for b in e2e_result) if e2e_result else 0
            e2e_ok = None
            if rank == 0:
                e2e_ok = tables_equal(pa.Table.from_batches(e2e_result), pa.Table.from_batches(result))
            e2e = {"value": n_total * args.e2e_steps / e2e_s, "unit": "rows/s", "h2d_bytes_per_step": n_local * bytes_per_row, "d2h_bytes_per_step": int(out_bytes),
                   "steps": args.e2e_steps, "ms_per_step": 1e3 * e2e_s / args.e2e_steps, "result_equals_resident_run": e2e_ok,
                   "note": "pinned host Arrow buffers -> sqlrs_plan_push_table (H2D) -> execute -> result to host, per rank shard"}
            del host_batch, keep_alive
        plan.close()
        torch.cuda.empty_cache()
        return e2e

    if rank != 0:
        return None, None, run_e2e
    peak, peak_src = measured_peak()
    k_ms = kernel_ms[0] / max(kernel_launches[0], 1)
    achieved = (n_local * bytes_per_row) / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
    cfg = q1_config(args.sf, n_total, n_local)
    cfg.update({"groups": len(result[0]) if result else None, "l2": "inputs (%.1f GB per GPU) are larger than the 126 MB L2" % (n_local * bytes_per_row / 1e9),
                "exchange": "none (1 GPU)" if world == 1 else "packed partial group tables all-gathered device-to-device over NCCL, folded on rank 0", "pipeline": describe})
    line = {
        "metric": "tpch_q1_rows_per_sec", "value": n_total / (ms_per_step * 1e-3), "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64+i64", "data": "synthetic", "config": cfg,
        "hbm_gbs_whole_step": n_total * bytes_per_row / (ms_per_step * 1e-3) / 1e9,
        "roofline": {"bound": "hbm", "kernel": "sq_agg_small", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                     "frac_of_8TBs": achieved / 8000.0 if achieved else None, "peak_source": peak_src, "kernel_ms": k_ms,
                     "algorithmic_bytes_per_launch": n_local * bytes_per_row,
                     "traffic": ncu_traffic(f"tpch_q1_sf{args.sf:g}", "sq_agg_small") if world == 1 else None,
                     "traffic_source": "profiles/r02_traffic.json (ncu --set full of this kernel source; null when the sources changed since, or N > 1)"},
        "e2e": None, "gpu_launches": int(launches), "parity_check": parity,
    }
    return line, window, run_e2e


def q3_device_tables(ctx, d, shards=None):
    from sqlrs_b200.host import tpch

    lib = ctx.lib
    out = {}
    for slot, (table, cols) in enumerate(((tpch.CUSTOMER, tpch.Q3_CUSTOMER_COLUMNS), (tpch.ORDERS, tpch.Q3_ORDERS_COLUMNS), (tpch.LINEITEM, tpch.Q3_LINEITEM_COLUMNS))):
        lo, hi = shards[slot] if shards else (0, None)
        out[slot] = tpch.device_table(lib, d, table, lo, hi, columns=cols, device=ctx.dev)
    return out


# Algorithmic bytes of the Q3' kernels: every input column of the table the kernel consumes, read once
# (customer 2, orders 4, lineitem 4 columns x 8 B).  The fused probe kernels materialise late: they STREAM only the Filter and
# join-key columns (16 B/row) and read the payload columns at matching rows only, so `achieved` on the algorithmic bytes can
# exceed the copy peak; `frac_streamed` is the same time against the bytes actually streamed.
Q3_KERNEL_BYTES = {
    "join 1 build (sq_joinbuild_kernel)": ("customer", 16, 16),
    "sq_joinchain_kernel": ("orders", 32, 16),
    "sq_joinagg_kernel": ("lineitem", 32, 16),
}


def q3_kernel_roofline(events, rows, runs, peak):
    """per-kernel: CUDA-event time per run, algorithmic bytes, achieved GB/s, fraction of the measured peak"""
    per = {k: v["ms"] / runs for k, v in events.items()}
    build = per.get("k_join_insert_kv", 0.0) + per.get("k_join_insert", 0.0) + sum(v for k, v in per.items() if k.startswith("sq_eval_kernel") or k.startswith("sq_joinbuild_kernel"))
    ms = {"join 1 build (sq_joinbuild_kernel)": build, "sq_joinchain_kernel": per.get("sq_joinchain_kernel", 0.0),
          "sq_joinagg_kernel": per.get("sq_joinagg_kernel", 0.0)}
    out = {}
    for name, (table, bpr, streamed) in Q3_KERNEL_BYTES.items():
        b = rows[table] * bpr
        t = ms[name]
        out[name] = {"ms": t, "algorithmic_bytes": b, "achieved": b / (t * 1e-3) / 1e9 if t > 0 else None, "frac": b / (t * 1e-3) / 1e9 / peak if t > 0 else None,
                     "frac_of_8TBs": b / (t * 1e-3) / 1e9 / 8000.0 if t > 0 else None, "streamed_bytes": rows[table] * streamed,
                     "frac_streamed": rows[table] * streamed / (t * 1e-3) / 1e9 / peak if t > 0 else None}
    out["other (finalise, top-k, gathers)"] = {"ms": sum(v for k, v in per.items() if k.startswith("aggregate finalise") or k.startswith("order")),
                                               "algorithmic_bytes": 0}
    return out


def run_gpu_q3(args, ctx, sf, cpu=None):
    """Q3' at scale factor `sf`: on one GPU the plan runs as is; on N GPUs orders / lineitem are sharded by orderkey ranges (the
    shards' key-range statistics show the joins between them need no exchange), customer is sharded by rows and its FILTERED
    rows are all-gathered device-to-device, every rank runs the whole query on its shards, the per-rank top-10 are merged."""
    import pyarrow as pa

    from sqlrs_b200.host import distributed as sqdist
    from sqlrs_b200.host import ffi, tpch
    from sqlrs_b200.host.plan import ExecutorBuilder

    torch, lib, rank, world = ctx.torch, ctx.lib, ctx.rank, ctx.world
    d = tpch.dims(sf)
    n_c, n_o, n_l = (tpch.num_rows(lib, d, t) for t in (tpch.CUSTOMER, tpch.ORDERS, tpch.LINEITEM))
    rows = {"customer": n_c, "orders": n_o, "lineitem": n_l}
    n_in = n_c + n_o + n_l
    alg = tpch.q3_algorithmic_bytes(n_c, n_o, n_l)
    full_root, schemas = tpch.q3_full_plan()
    agg_root, _ = tpch.q3_plan()
    stream_ptr = C.c_void_p(ctx.stream.cuda_stream)
    builder = ExecutorBuilder(lib, lib.options(device_id=ctx.local_rank, stream=stream_ptr, **MODE))
    ev_builder = ExecutorBuilder(lib, lib.options(device_id=ctx.local_rank, stream=stream_ptr, flags=ffi.FLAG_KERNEL_EVENTS, **MODE))
    peak, peak_src = measured_peak()
    out = {"workload": f"tpch_q3_sf{sf:g}", "metric": "tpch_q3_rows_per_sec", "unit": "rows/s", "n_gpus": world, "rows": rows, "input_rows": n_in,
           "query": "customer join orders join lineitem, 3 single-table predicates, GROUP BY l_orderkey, o_orderdate, o_shippriority, SUM(l_extendedprice*(1-l_discount)), "
                    "ORDER BY revenue desc, o_orderdate LIMIT 10 — whole query on the device, 10 rows to the host",
           "count_mode": "sql_accumulate", "match_mode": "hash_and_key", "steps": args.steps, "warmup": args.warmup}
    if world == 1:
        tabs = q3_device_tables(ctx, d)
        plan = builder.build(full_root, schemas)
        for k, t in tabs.items():
            plan.push_table_device(k, t)

        def step():
            plan.execute()
            return plan.collect()
    else:
        (o_lo, o_hi), (l_lo, l_hi) = sqdist.copartitioned_shard(int(d.n_orders), rank, world)  # cuts at order boundaries: an orderkey RANGE per rank
        tabs = q3_device_tables(ctx, d, shards=[(n_c * rank // world, n_c * (rank + 1) // world), (o_lo, o_hi), (l_lo, l_hi)])
        # table statistics (min / max of the join key per shard), computed once at load: do the fact-table joins need an exchange?
        copart = sqdist.key_ranges_copartitioned(ctx.group, tabs[1].key_range(0), tabs[2].key_range(0))
        if not copart:
            raise RuntimeError("bench: the orderkey-range shards are expected to be co-partitioned")
        cust_plan = full_root.child.child.child.child.left.left  # Filter(Scan(customer)) of the query
        state = {}

        def step():
            return sqdist.distributed_join_topk(builder, ctx.group, build_plan=cust_plan, build_schemas={0: schemas[0]}, build_tables={0: tabs[0]},
                                                query_plan=full_root, query_schemas=schemas, query_tables={1: tabs[1], 2: tabs[2]}, build_slot=0,
                                                order_by=tpch.q3_tail_order_by(), limit=10, state=state)

        out["partitioning"] = ("orders + lineitem sharded by orderkey ranges (co-partitioned: decided from the shards' key-range statistics), customer sharded by rows; "
                               "filtered customer rows all-gathered device-to-device (NCCL), per-rank top-10 merged on rank 0")
    l0 = lib.kernel_launches()
    ms, result, window = timed(ctx, step, args.steps, args.warmup)
    launches = lib.kernel_launches() - l0
    top = pa.Table.from_batches(result) if (rank == 0 and result) else None
    out.update({"ms_per_step": ms, "value": n_in / (ms * 1e-3), "lineitem_rows_per_s": n_l / (ms * 1e-3), "gpu_launches_per_step": launches / max(args.steps, 1)})
    out["roofline"] = {"bound": "hbm", "whole_query": {"algorithmic_bytes": alg, "achieved": alg / (ms * 1e-3) / 1e9, "frac": alg / (ms * 1e-3) / 1e9 / peak / world,
                                                      "frac_of_8TBs": alg / (ms * 1e-3) / 1e9 / 8000.0 / world, "note": "per GPU; bytes = input columns read once (SURVEY §8d)"},
                       "peak": peak, "unit": "GB/s", "peak_source": peak_src}

    if world == 1:
        out["pipeline"] = plan.describe()
        plan.close()
        # ---- per-kernel device times (CUDA events around the hot kernels; a separate plan so the timed loop above carries no events)
        ev_plan = ev_builder.build(full_root, schemas)
        for k, t in tabs.items():
            ev_plan.push_table_device(k, t)
        for _ in range(2):
            ev_plan.execute()
            ev_plan.collect()
        torch.cuda.synchronize(ctx.dev)
        ev_plan.kernel_events()
        runs = 5
        for _ in range(runs):
            ev_plan.execute()
            ev_plan.collect()
        torch.cuda.synchronize(ctx.dev)
        kernels = q3_kernel_roofline(ev_plan.kernel_events(), rows, runs, peak)
        for name in kernels:
            short = "sq_joinbuild_kernel" if name.startswith("join 1 build") else name.split(" ")[0]
            kernels[name]["traffic"] = ncu_traffic(f"tpch_q3_sf{sf:g}", short) if short.startswith("sq_") else None
        out["roofline"]["kernels"] = kernels
        ev_plan.close()
        # ---- the plan up to the aggregate (SURVEY §8d's timed region: every group to the host, in first-appearance order)
        agg_plan = builder.build(agg_root, schemas)
        for k, t in tabs.items():
            agg_plan.push_table_device(k, t)

        def agg_step():
            agg_plan.execute()
            return agg_plan.collect()

        agg_ms, agg_res, _ = timed(ctx, agg_step, max(3, args.steps // 2), 2)
        agg_tab = pa.Table.from_batches(agg_res)
        names = agg_tab.schema.names
        want = agg_tab.sort_by([(names[3], "descending"), (names[1], "ascending")]).slice(0, 10)
        top_ok = top.num_rows == want.num_rows and top.column(0).to_pylist() == want.column(0).to_pylist()
        out["plan_to_aggregate"] = {"ms_per_step": agg_ms, "rows_per_s": n_in / (agg_ms * 1e-3), "groups": agg_tab.num_rows,
                                    "d2h_bytes_per_step": int(sum(b.nbytes for b in agg_res))}
        agg_plan.close()
        parity = {"top10_equals_top10_of_the_aggregate_output": bool(top_ok), "result_rows": top.num_rows}
        # ---- e2e: host (pinned) Arrow buffers -> push_table (H2D) -> whole query -> 10 rows to the host (deferred: see main)
        def run_e2e():
            if args.e2e_steps <= 0:
                return
            host = {}
            keep_alive = []
            for k, t in tabs.items():
                host[k], ka = pinned_batch(ctx, t)
                keep_alive.append(ka)
            e_plan = builder.build(full_root, schemas)

            def e2e_step():
                for k, b in host.items():
                    e_plan.push_table(k, b)
                e_plan.execute()
                res = e_plan.collect()
                e_plan.reset()
                return res

            e2e_step()
            torch.cuda.synchronize(ctx.dev)
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                e_res = e2e_step()
            torch.cuda.synchronize(ctx.dev)
            e_s = time.perf_counter() - t0
            out["e2e"] = {"value": n_in * args.e2e_steps / e_s, "unit": "rows/s", "ms_per_step": 1e3 * e_s / args.e2e_steps, "h2d_bytes_per_step": int(alg),
                          "d2h_bytes_per_step": int(sum(b.nbytes for b in e_res)), "steps": args.e2e_steps,
                          "result_equals_resident_run": tables_equal(pa.Table.from_batches(e_res), top),
                          "note": "pinned host Arrow buffers -> sqlrs_plan_push_table x3 (H2D) -> execute -> 10 rows to host"}
            out["parity_check"]["ok"] = bool(out["parity_check"]["ok"] and out["e2e"]["result_equals_resident_run"])
            e_plan.close()
            del host, keep_alive

        # ---- CPU port beside it, and the oracle as checker: the same plan at the CPU sample's scale factor on both
        if cpu is not None:
            cpu_stats, cpu_res = cpu
            dc = tpch.dims(cpu_stats["sf"])
            small = q3_device_tables(ctx, dc)
            sp = builder.build(full_root, schemas)
            for k, t in small.items():
                sp.push_table_device(k, t)
            got_small = pa.Table.from_batches(sp.run())
            sp.close()
            parity["oracle_sample"] = f"whole query at SF{cpu_stats['sf']:g} ({cpu_stats['rows']} input rows)"
            parity["equals_oracle_on_sample"] = tables_equal(got_small, cpu_res)
            out["cpu_baseline"] = {"value": cpu_stats["rows_per_s"], "unit": "rows/s", "cores": 1, "kind": "port",
                                   "sample": f"Q3' whole query at SF{cpu_stats['sf']:g} ({cpu_stats['rows']} input rows), one pass, 1024-row batches sliced inside the library, "
                                             "C++ restatement of the sqlrs v1 executor"}
        parity["ok"] = bool(top_ok) and parity.get("equals_oracle_on_sample", True)
        out["parity_check"] = parity
    else:
        out["pipeline"] = state["p_query"].describe()
        # ---- parity: the N-GPU top-10 against the single-GPU plan over the whole tables, on rank 0
        parity = None
        if rank == 0:
            whole = q3_device_tables(ctx, d)
            sp = builder.build(full_root, schemas)
            for k, t in whole.items():
                sp.push_table_device(k, t)
            single = pa.Table.from_batches(sp.run())
            sp.close()
            del whole
            parity = {"equals_single_gpu_result": tables_equal(top, single), "result_rows": top.num_rows}
            parity["ok"] = parity["equals_single_gpu_result"]
        out["parity_check"] = parity
        # ---- e2e: every rank's shards in pinned host memory (deferred: see main)
        def run_e2e():
            if args.e2e_steps > 0:
                host, keep_alive = {}, []
                for k, t in tabs.items():
                    host[k], ka = pinned_batch(ctx, t)
                    keep_alive.append(ka)
                e_state = {}

                def e2e_step():
                    return sqdist.distributed_join_topk(builder, ctx.group, build_plan=cust_plan, build_schemas={0: schemas[0]}, build_tables={0: host[0]},
                                                        query_plan=full_root, query_schemas=schemas, query_tables={1: host[1], 2: host[2]}, build_slot=0,
                                                        order_by=tpch.q3_tail_order_by(), limit=10, state=e_state)

                e2e_step()
                barrier(ctx)
                t0 = time.perf_counter()
                for _ in range(args.e2e_steps):
                    e_res = e2e_step()
                barrier(ctx)
                e_s = max_over_ranks(ctx, time.perf_counter() - t0)
                out["e2e"] = {"value": n_in * args.e2e_steps / e_s, "unit": "rows/s", "ms_per_step": 1e3 * e_s / args.e2e_steps,
                              "h2d_bytes_per_step": int(sum(t.nbytes() for t in tabs.values())), "d2h_bytes_per_step": int(sum(b.nbytes for b in e_res)) if e_res else 0,
                              "steps": args.e2e_steps, "result_equals_resident_run": tables_equal(pa.Table.from_batches(e_res), top) if rank == 0 else None,
                              "note": "per rank: pinned host shards -> push_table (H2D) -> filtered customer rows all-gathered -> whole query -> top-10 merged"}
                for p in e_state.values():
                    if hasattr(p, "close"):
                        p.close()
                del host, keep_alive
            for p in state.values():
                if hasattr(p, "close"):
                    p.close()

    return (out if rank == 0 else None), window, run_e2e


def main():
    global MODE
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sf", type=float, default=100.0)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-rows", type=int, default=48_000_000, help="rows of the Q1' cpu_baseline sample (0 = skip)")
    ap.add_argument("--ref-rows", type=int, default=12_000_000, help="rows per step of --impl reference")
    ap.add_argument("--query", default="q1", choices=["q1", "q3"], help="which query is the line's headline (the other one rides along)")
    ap.add_argument("--q3", default="on", choices=["on", "off"], help="measure Q3' too (the `q3` object of the line)")
    ap.add_argument("--q3-sf", type=float, default=None, help="only this Q3' scale factor (default: 10 and 100 on one GPU, 100 on N)")
    ap.add_argument("--cpu-sf", type=float, default=1.0, help="scale factor of the Q3' cpu_baseline / oracle sample")
    args = ap.parse_args()
    from sqlrs_b200.host import ffi

    MODE = dict(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus:
        log(f"note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    import pyarrow as pa
    import torch

    from sqlrs_b200.host import distributed as sqdist
    from sqlrs_b200.host import tpch

    ctx = Ctx()
    ctx.torch, ctx.rank, ctx.world, ctx.local_rank = torch, rank, world, local_rank
    torch.cuda.set_device(local_rank)
    ctx.dev = torch.device("cuda", local_rank)
    ctx.dist = None
    if world > 1:
        import torch.distributed as dist_mod

        ctx.dist = dist_mod
        ctx.dist.init_process_group("nccl", device_id=ctx.dev)
    ctx.lib = ffi.load()
    ctx.stream = torch.cuda.Stream(device=ctx.dev)
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("SQLRS_BENCH_NO_SAMPLER"):
        sampler.start()  # before any warm-up: NVML start-up stays out of the timed regions
    windows = []
    line = None
    q3 = {}
    with torch.cuda.stream(ctx.stream):
        ctx.group = sqdist.TorchGroup(ctx.dist, ctx.dev) if world > 1 else None
        deferred = []  # the e2e legs: after EVERY resident timing
        if args.query == "q1" or args.q3 == "on":
            line, w, fin = run_gpu_q1(args, ctx)
            if w:
                windows.append(w)
            deferred.append(("q1", fin))
        if args.q3 == "on" or args.query == "q3":
            cpu = None
            if rank == 0 and world == 1 and args.cpu_rows > 0:
                cpu = cpu_port_q3(load_oracle(), args.cpu_sf)
            sfs = [args.q3_sf] if args.q3_sf else ([10.0, 100.0] if world == 1 else [100.0])
            for sf in sfs:
                entry, w, fin = run_gpu_q3(args, ctx, sf, cpu)
                windows.append(w)
                if entry:
                    q3[f"sf{sf:g}"] = entry
                deferred.append((f"q3 sf{sf:g}", fin))
        for name, fin in deferred:
            r = fin()
            if name == "q1" and line is not None:
                line["e2e"] = r
    if world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()
    if rank != 0:
        return
    clocks = sampler.stop(windows)
    if line is not None and world == 1 and args.cpu_rows > 0:
        # the CPU port beside it (1 core: the reference's executor is single-threaded per query), and as the checker: the GPU
        # plan over the SAME sample rows must return what the oracle returns
        oracle = load_oracle()
        r, cpu_res = cpu_port_q1(oracle, args.sf, args.cpu_rows, 1024)
        r64, _ = cpu_port_q1(oracle, args.sf, min(args.cpu_rows, 12_000_000), 65536)
        line["cpu_baseline"] = {"value": r["rows_per_s"], "unit": "rows/s", "cores": 1, "kind": "port", "value_batch_65536": r64["rows_per_s"],
                                "sample": f"first {r['rows']} lineitem rows of SF{args.sf:g}, one pass, 1024-row batches sliced inside the library (no Python in the loop), "
                                          "C++ restatement of the sqlrs v1 executor; value_batch_65536: the same over 12 M rows in 65,536-row batches"}
        from sqlrs_b200.host.plan import ExecutorBuilder

        with torch.cuda.stream(ctx.stream):
            d = tpch.dims(args.sf)
            t = tpch.device_table(ctx.lib, d, tpch.LINEITEM, 0, r["rows"], columns=tpch.Q1_COLUMNS, device=ctx.dev)
            plan_root, schemas = tpch.q1_plan()
            p = ExecutorBuilder(ctx.lib, ctx.lib.options(device_id=local_rank, stream=C.c_void_p(ctx.stream.cuda_stream), **MODE)).build(plan_root, schemas)
            p.push_table_device(0, t)
            got = pa.Table.from_batches(p.run())
            p.close()
        line["parity_check"]["oracle_sample_rows"] = r["rows"]
        line["parity_check"]["equals_oracle_on_sample"] = tables_equal(got, cpu_res)
        line["parity_check"]["ok"] = bool(line["parity_check"]["ok"] and line["parity_check"]["equals_oracle_on_sample"])
    if args.query == "q3" or line is None:
        head = q3.get("sf100") or next(iter(q3.values()))
        line = {"metric": "tpch_q3_rows_per_sec", "value": head["value"], "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "i64+f64", "data": "synthetic",
                "config": {"workload": head["workload"], "rows": head["rows"]}, "roofline": head["roofline"], "e2e": head.get("e2e"),
                "cpu_baseline": head.get("cpu_baseline"), "gpu_launches": int(head["gpu_launches_per_step"] * args.steps), "parity_check": head.get("parity_check")}
    if q3:
        line["q3"] = q3
    line["clocks"] = clocks
    emit(line)


if __name__ == "__main__":
    main()
