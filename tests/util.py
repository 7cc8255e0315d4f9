"""Helpers shared by the tests: batch builders, result comparison."""
import math

import pyarrow as pa


def batch(schema_or_names, *columns, types=None, nullable=True):
    """batch(["a","b"], [1,2], [3,None]) -> RecordBatch (int64 unless `types` says otherwise)."""
    names = list(schema_or_names)
    types = types or [pa.int64()] * len(names)
    fields = [pa.field(n, t, nullable=nullable if isinstance(nullable, bool) else nullable[i]) for i, (n, t) in enumerate(zip(names, types))]
    arrays = [pa.array(c, type=t) for c, t in zip(columns, types)]
    return pa.RecordBatch.from_arrays(arrays, schema=pa.schema(fields))


def rows_of(batches):
    """All rows of a list of batches as tuples (None = NULL), in stream order."""
    out = []
    for b in batches:
        cols = [c.to_pylist() for c in b.columns]
        out.extend(zip(*cols) if cols else [])
    return [tuple(r) for r in out]


def assert_rows_equal(got, want, rtol=0.0, ordered=True, atol=1e-9):
    """Floats: |a-b| <= rtol*max(|a|,|b|) + atol when rtol > 0 (atol covers sums that cancel to ~0, where a
    relative bound is meaningless: summation order differs between the CPU batches and the GPU atomics)."""
    if not ordered:
        key = lambda r: tuple((x is None, 0 if x is None else x) for x in r)
        got, want = sorted(got, key=key), sorted(want, key=key)
    assert len(got) == len(want), f"row count {len(got)} != {len(want)}\n got={got[:10]}\nwant={want[:10]}"
    for i, (g, w) in enumerate(zip(got, want)):
        assert len(g) == len(w), (i, g, w)
        for a, b in zip(g, w):
            if isinstance(b, float) and isinstance(a, float) and rtol > 0:
                if math.isnan(a) and math.isnan(b):
                    continue
                assert abs(a - b) <= rtol * max(abs(a), abs(b)) + atol, f"row {i}: {g} != {w}"
            else:
                assert a == b or (isinstance(a, float) and isinstance(b, float) and math.isnan(a) and math.isnan(b)), f"row {i}: {g} != {w}"


def assert_batches_match(got, want, rtol=0.0, ordered=True, check_names=True):
    """got / want: lists of RecordBatch.  Integer columns bit-exact, floats within rtol (0 = exact)."""
    if got and want and check_names:
        assert got[0].schema.names == want[0].schema.names, (got[0].schema.names, want[0].schema.names)
        assert [str(t) for t in got[0].schema.types] == [str(t) for t in want[0].schema.types], (got[0].schema, want[0].schema)
    assert_rows_equal(rows_of(got), rows_of(want), rtol=rtol, ordered=ordered)
