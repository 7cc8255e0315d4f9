"""sqlrs_b200 — B200-native execution backend for the hot operators of Fedomn/sqlrs.

`sqlrs_b200.csrc` holds the CUDA kernels and the C ABI (include/sqlrs_b200.h, built into
csrc/libsqlrs_b200.so); `sqlrs_b200.host` is the Python mirror of the reference's operator
interface that drives it through ctypes + the Arrow C Data Interface.  There is no CPU fallback:
`host.ffi.load()` raises if the CUDA library has not been built.
"""
from .host import executor, expr, ffi, plan, tpch  # noqa: F401
from .host.executor import (FilterExecutor, HashAggExecutor, HashJoinExecutor, JoinCondition, SimpleAggExecutor,  # noqa: F401
                            create_hashes, eval_column, try_collect)
from .host.expr import AggFunc, Alias, BinaryOp, Constant, InputRef, TypeCast, bind_binary_op  # noqa: F401
from .host.ffi import ExecutorError, load  # noqa: F401
from .host.plan import (ExecutorBuilder, PhysicalFilter, PhysicalHashAgg, PhysicalHashJoin, PhysicalSimpleAgg,  # noqa: F401
                        PhysicalTableScan)
