"""Host-side mirror of the reference's in-memory storage (src/storage/memory.rs), with the tables resident in HBM.

`InMemoryStorage.create_mem_table(id, batches)` / `get_table(id)` as in `memory.rs:38-56`; an `InMemoryTable` is a handle
of `sqlrs_table_*` (include/sqlrs_b200.h): its batches are copied to the device once and scanned by any number of plans
(`GpuPlan.push_table_resident`) without crossing PCIe again.  `read()` mirrors `InMemoryTransaction::next_batch`
(`memory.rs:151-170`).  This file contains no compute.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterator, List, Optional, Sequence

import pyarrow as pa

from . import ffi


class StorageError(RuntimeError):
    """StorageError::TableNotFound (src/storage/mod.rs)"""


class InMemoryTable:
    def __init__(self, lib: ffi.Library, table_id: str, data: Sequence[pa.RecordBatch], options=None):
        self.lib, self.id = lib, table_id
        self.schema: Optional[pa.Schema] = data[0].schema if data else None
        self.handle = C.c_void_p()
        opts = options if options is not None else lib.options()
        lib.check(lib.table_create(C.byref(opts), C.byref(self.handle)))
        for batch in data:
            arr, sch = ffi.export_batch(batch)
            try:
                lib.check(lib.table_append(self.handle, C.byref(arr), C.byref(sch)))
            finally:
                ffi.release_schema(sch)

    @property
    def num_rows(self) -> int:
        return int(self.lib.table_num_rows(self.handle))

    @property
    def num_batches(self) -> int:
        return int(self.lib.table_num_batches(self.handle))

    def read(self, projection: Optional[Sequence[int]] = None) -> Iterator[pa.RecordBatch]:
        """Table::read + Transaction::next_batch until None (memory.rs:137-170)"""
        proj = (C.c_int32 * len(projection))(*projection) if projection is not None else None
        k, has = 0, C.c_int32(0)
        while True:
            arr, sch = ffi.ArrowArray(), ffi.ArrowSchema()
            self.lib.check(self.lib.table_read(self.handle, k, proj, len(projection) if projection is not None else 0, C.byref(arr), C.byref(sch),
                                               C.byref(has)))
            if not has.value:
                return
            yield ffi.import_batch(arr, sch)
            k += 1

    def close(self):
        if self.handle:
            self.lib.table_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CsvTable(InMemoryTable):
    """CsvTable (src/storage/csv.rs:111-170): `CsvStorage::create_csv_table(id, filepath)` infers the schema from the file and
    scans it in 1024-row batches; here the file is parsed on the device once (sqlrs_table_read_csv) and the table is resident.
    `bounds` = (offset, limit) over the whole table and `projection` as in `Table::read(bounds, projection)` (:163-170)."""

    def __init__(self, lib: ffi.Library, table_id: str, filepath: str, options=None, has_header: bool = True, delimiter: str = ",", batch_size: int = 1024,
                 bounds=None, projection: Optional[Sequence[int]] = None):
        self.lib, self.id, self.filepath = lib, table_id, filepath
        self.handle = C.c_void_p()
        opts = options if options is not None else lib.options()
        proj = (C.c_int32 * len(projection))(*projection) if projection is not None else None
        off, lim = bounds if bounds is not None else (-1, -1)
        lib.check(lib.table_read_csv(filepath.encode(), int(has_header), ord(delimiter), batch_size, off, lim, proj, len(projection) if projection is not None else 0,
                                     C.byref(opts), C.byref(self.handle)))
        first = next(iter(self.read()), None)
        self.schema = first.schema if first is not None else None


class InMemoryStorage:
    """InMemoryStorage (memory.rs:9-60): id -> table"""

    def __init__(self, lib: Optional[ffi.Library] = None, options=None):
        self.lib = lib if lib is not None else ffi.load()
        self.options = options
        self.tables: Dict[str, InMemoryTable] = {}

    def create_mem_table(self, table_id: str, data: Sequence[pa.RecordBatch]) -> None:
        self.tables[table_id] = InMemoryTable(self.lib, table_id, list(data), self.options)

    def create_csv_table(self, table_id: str, filepath: str, **kw) -> None:
        """CsvStorage::create_csv_table (csv.rs:60-72)"""
        self.tables[table_id] = CsvTable(self.lib, table_id, filepath, self.options, **kw)

    def get_table(self, table_id: str) -> InMemoryTable:
        try:
            return self.tables[table_id]
        except KeyError:
            raise StorageError(f"table not found: {table_id}") from None
