"""ctypes binding of libpolarway_b200.so (include/polarway_b200.h) + the glue that turns a
``plan.GroupByPlan`` over a pyarrow Table into one C-ABI call.

This is the Python stand-in for the Rust operator binding described in INTEGRATION.md: columns
cross as Arrow C Data Interface structs (``pyarrow.Array._export_to_c``), the query as a ``PwQuery``.
There is no CPU fallback: if the shared library or a CUDA device is missing every call raises.
"""
from __future__ import annotations

import ctypes as C
import datetime as _dt
import os
from typing import Optional

import pyarrow as pa

from . import plan as P

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpolarway_b200.so")

PW_MAX_FACTORS = 4
FLAG_FORCE_HOT = 1 << 0
FLAG_FORCE_GLOBAL = 1 << 1
FLAG_FORCE_SEGMENTED = 1 << 2
FLAG_NO_SEGMENTED = 1 << 3
FLAG_FORCE_PARTITION = 1 << 4
FLAG_NO_PARTITION = 1 << 5
FLAG_NO_DENSE_IDS = 1 << 6
FLAG_NO_BUCKETS = 1 << 7
FLAG_KEYS_SORTED = 1 << 8


class PolarwayError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class _Scalar(C.Union):
    _fields_ = [("i", C.c_int64), ("u", C.c_uint64), ("f", C.c_double)]


class PwPredicate(C.Structure):
    _fields_ = [("column", C.c_int32), ("op", C.c_int32), ("scalar_is_float", C.c_int32), ("reserved", C.c_int32),
                ("scalar", _Scalar)]


class PwFactor(C.Structure):
    _fields_ = [("a", C.c_double), ("b", C.c_double), ("column", C.c_int32), ("reserved", C.c_int32)]


class PwAgg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("column", C.c_int32), ("n_factors", C.c_int32), ("ddof", C.c_int32),
                ("factors", PwFactor * PW_MAX_FACTORS), ("name", C.c_char_p)]


class PwDynamic(C.Structure):
    _fields_ = [("index_column", C.c_int32), ("closed", C.c_int32), ("label", C.c_int32), ("include_boundaries", C.c_int32),
                ("every", C.c_int64), ("period", C.c_int64), ("offset", C.c_int64)]


class PwQuery(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("maintain_order", C.c_int32), ("n_predicates", C.c_int32),
                ("n_keys", C.c_int32), ("n_aggs", C.c_int32), ("hot_table_slots", C.c_int32),
                ("predicates", C.POINTER(PwPredicate)), ("key_columns", C.POINTER(C.c_int32)),
                ("aggs", C.POINTER(PwAgg)), ("dynamic", C.POINTER(PwDynamic)), ("flags", C.c_uint64),
                ("row_offset", C.c_int64), ("initial_table_slots", C.c_int64)]


class PwTimings(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("estimate_ms", C.c_float), ("scan_ms", C.c_float), ("finalize_ms", C.c_float),
                ("d2h_ms", C.c_float), ("total_device_ms", C.c_float), ("n_rows", C.c_int64), ("n_groups", C.c_int64),
                ("table_slots", C.c_int64), ("strategy", C.c_int32), ("retries", C.c_int32),
                ("kernel_launches", C.c_int64), ("spilled_rows", C.c_int64), ("scan_kernel_ms", C.c_float),
                ("reserved", C.c_float), ("host_ms", C.c_float), ("partition_ms", C.c_float),
                ("jit_compiles", C.c_int32), ("jit_cache_hits", C.c_int32)]


class ArrowSchema(C.Structure):
    pass


class ArrowArray(C.Structure):
    pass


ArrowSchema._fields_ = [("format", C.c_char_p), ("name", C.c_char_p), ("metadata", C.c_char_p), ("flags", C.c_int64),
                        ("n_children", C.c_int64), ("children", C.POINTER(C.POINTER(ArrowSchema))),
                        ("dictionary", C.POINTER(ArrowSchema)), ("release", C.c_void_p), ("private_data", C.c_void_p)]
ArrowArray._fields_ = [("length", C.c_int64), ("null_count", C.c_int64), ("offset", C.c_int64), ("n_buffers", C.c_int64),
                       ("n_children", C.c_int64), ("buffers", C.POINTER(C.c_void_p)),
                       ("children", C.POINTER(C.POINTER(ArrowArray))), ("dictionary", C.POINTER(ArrowArray)),
                       ("release", C.c_void_p), ("private_data", C.c_void_p)]

_lib = None


def lib():
    """Loads the CUDA library; raises if it has not been built (there is no other engine)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PolarwayError(-3, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`; "
                                    "polaroid_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.pw_b200_last_error.restype = C.c_char_p
        L.pw_b200_abi_version.restype = C.c_uint32
        L.pw_b200_frame_num_rows.restype = C.c_int64
        L.pw_b200_frame_num_rows.argtypes = [C.c_void_p]
        L.pw_b200_frame_free.argtypes = [C.c_void_p]
        L.pw_b200_set_stream.argtypes = [C.c_void_p]
        L.pw_b200_frame_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.pw_b200_frame_from_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.pw_b200_frame_groupby.argtypes = [C.POINTER(PwQuery), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]
        L.pw_b200_filter_groupby_agg.argtypes = [C.POINTER(PwQuery), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                                 C.c_void_p, C.POINTER(C.c_size_t)]
        L.pw_b200_last_timings.argtypes = [C.POINTER(PwTimings)]
        L.pw_b200_filter.argtypes = [C.POINTER(PwPredicate), C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.pw_b200_frame_filter_select.argtypes = [C.POINTER(PwPredicate), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                                  C.POINTER(C.c_int64)]
        L.pw_b200_frame_group_slices.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pw_b200_frame_group_tuples.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_void_p]
        L.pw_b200_frame_groupby_partial.argtypes = [C.POINTER(PwQuery), C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
        L.pw_b200_partial_row_words.restype = C.c_int64
        L.pw_b200_partial_row_words.argtypes = [C.POINTER(PwQuery), C.c_void_p]
        L.pw_b200_frame_groupby_partial_into.argtypes = [C.POINTER(PwQuery), C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]
        L.pw_b200_merge_gathered.argtypes = [C.POINTER(PwQuery), C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
                                             C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]
        L.pw_b200_partial_row_bytes.restype = C.c_int64
        L.pw_b200_partial_row_bytes.argtypes = [C.c_void_p]
        L.pw_b200_partial_device_rows.restype = C.c_void_p
        L.pw_b200_partial_device_rows.argtypes = [C.c_void_p]
        L.pw_b200_partial_offsets.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.pw_b200_partial_copy_rows.argtypes = [C.c_void_p, C.c_void_p]
        L.pw_b200_partial_free.argtypes = [C.c_void_p]
        L.pw_b200_merge_partials.argtypes = [C.POINTER(PwQuery), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                             C.POINTER(C.c_size_t)]
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise PolarwayError(rc, lib().pw_b200_last_error().decode("utf-8", "replace"))


def last_timings() -> dict:
    t = PwTimings()
    lib().pw_b200_last_timings(C.byref(t))
    return {k: getattr(t, k) for k, _ in PwTimings._fields_}


# ---- Arrow export / import ---------------------------------------------------------------------------
def _is_stringlike(t) -> bool:
    return (pa.types.is_string(t) or pa.types.is_large_string(t) or pa.types.is_binary(t) or pa.types.is_large_binary(t)
            or pa.types.is_string_view(t) or pa.types.is_binary_view(t))


def _to_abi_array(col) -> pa.Array:
    """One contiguous array.  String columns stay as they are here and are turned into Utf8View buffers
    (what Polars itself exports, polars-ffi/src/version_0.rs:64) by ``_ViewColumn``."""
    if isinstance(col, pa.ChunkedArray):
        col = col.combine_chunks() if col.num_chunks != 1 else col.chunk(0)
    return col


class _ViewColumn:
    """A string/binary column laid out as Arrow Utf8View buffers [validity, views, data, variadic sizes],
    built with numpy.  (pyarrow 24's own C export of a view array that has no data buffer segfaults, so the
    struct is filled in by hand; the layout is the one polars-arrow/src/array/binview/view.rs:19-55 defines.)"""

    def __init__(self, arr: pa.Array):
        import numpy as np
        t = arr.type
        if pa.types.is_string_view(t) or pa.types.is_binary_view(t):
            arr = arr.cast(pa.large_binary() if pa.types.is_binary_view(t) else pa.large_string())
            t = arr.type
        self.is_binary = pa.types.is_binary(t) or pa.types.is_large_binary(t)
        if arr.offset != 0 or not (pa.types.is_large_string(t) or pa.types.is_large_binary(t)):
            arr = pa.concat_arrays([arr]).cast(pa.large_binary() if self.is_binary else pa.large_string())
            if arr.offset != 0:
                arr = pa.array(arr.to_pylist(), type=arr.type)
        n = len(arr)
        bufs = arr.buffers()
        off = np.frombuffer(bufs[1], dtype=np.int64, count=n + 1) if n else np.zeros(1, dtype=np.int64)
        data = np.frombuffer(bufs[2], dtype=np.uint8) if bufs[2] is not None and bufs[2].size else np.zeros(1, dtype=np.uint8)
        lens = (off[1:] - off[:-1]).astype(np.int64)
        if arr.null_count:
            valid = np.asarray(arr.is_valid().to_numpy(zero_copy_only=False))
            lens = np.where(valid, lens, 0)
            self.validity = np.packbits(valid, bitorder="little")
        else:
            self.validity = None
        views = np.zeros((n, 16), dtype=np.uint8)
        if n:
            views[:, 0:4] = lens.astype("<u4").view(np.uint8).reshape(n, 4)
            j = np.arange(12, dtype=np.int64)
            idx = np.minimum(off[:-1, None] + j[None, :], max(len(data) - 1, 0))
            inline = lens <= 12
            take = (j[None, :] < lens[:, None]) & inline[:, None]
            views[:, 4:16] = np.where(take, data[idx], 0)
            longm = ~inline
            if longm.any():
                pre = (j[None, :4] < lens[:, None])
                views[longm, 4:8] = np.where(pre[longm], data[idx[longm, :4]], 0)
                views[longm, 8:12] = 0  # buffer index 0
                views[longm, 12:16] = off[:-1][longm].astype("<u4").view(np.uint8).reshape(-1, 4)
        self.views = np.ascontiguousarray(views)
        self.data = np.ascontiguousarray(data)
        self.sizes = np.array([len(self.data)], dtype=np.int64)
        self.length = n
        self.null_count = arr.null_count
        self._bufs = (C.c_void_p * 4)(self.validity.ctypes.data if self.validity is not None else None,
                                      self.views.ctypes.data, self.data.ctypes.data, self.sizes.ctypes.data)

    def fill(self, c_array: "ArrowArray", c_schema: "ArrowSchema"):
        c_array.length = self.length
        c_array.null_count = self.null_count
        c_array.offset = 0
        c_array.n_buffers = 4
        c_array.n_children = 0
        c_array.buffers = C.cast(self._bufs, C.POINTER(C.c_void_p))
        c_array.release = None
        c_schema.format = b"vz" if self.is_binary else b"vu"
        c_schema.flags = 2
        c_schema.release = None


class _Exported:
    """Keeps the exported ArrowArray/ArrowSchema structs (and their owners) alive for a call."""

    def __init__(self, arrays):
        n = len(arrays)
        self.arrays = arrays
        self.c_arrays = (ArrowArray * n)()
        self.c_schemas = (ArrowSchema * n)()
        self.views = {}
        for i, a in enumerate(arrays):
            if _is_stringlike(a.type):
                self.views[i] = _ViewColumn(a)
                self.views[i].fill(self.c_arrays[i], self.c_schemas[i])
            else:
                a._export_to_c(C.addressof(self.c_arrays[i]), C.addressof(self.c_schemas[i]))
        self.arr_ptrs = (C.c_void_p * n)(*[C.addressof(self.c_arrays[i]) for i in range(n)])
        self.sch_ptrs = (C.c_void_p * n)(*[C.addressof(self.c_schemas[i]) for i in range(n)])
        self.n = n

    def set_names(self, names):
        self._names = [n.encode() for n in names]
        for i, b in enumerate(self._names):
            self.c_schemas[i].name = b

    def release(self):
        for i in range(self.n):
            for st in (self.c_arrays[i], self.c_schemas[i]):
                if st.release:
                    C.CFUNCTYPE(None, C.c_void_p)(st.release)(C.addressof(st))


def _import_view_array(c_array: "ArrowArray", c_schema: "ArrowSchema") -> pa.Array:
    """Utf8View result (library output: inline values, plus one data buffer when keys longer than 12 bytes were
    gathered) -> pyarrow large_string, then release the C structs."""
    import numpy as np
    n = int(c_array.length)
    binary = c_schema.format == b"vz"
    if n:
        views = np.ctypeslib.as_array(C.cast(c_array.buffers[1], C.POINTER(C.c_uint8)), shape=(n, 16)).copy()
        lens = views[:, 0:4].copy().view("<u4").reshape(n).astype(np.int64)
        valid = None
        if c_array.null_count and c_array.buffers[0]:
            vb = np.ctypeslib.as_array(C.cast(c_array.buffers[0], C.POINTER(C.c_uint8)), shape=((n + 7) // 8,)).copy()
            valid = np.unpackbits(vb, bitorder="little")[:n].astype(bool)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        longm = lens > 12
        if longm.any():
            # buffers = [validity, views, data, sizes]: views of long values are (length, prefix, buffer 0, offset)
            size = int(C.cast(c_array.buffers[3], C.POINTER(C.c_int64))[0])
            blob = np.ctypeslib.as_array(C.cast(c_array.buffers[2], C.POINTER(C.c_uint8)), shape=(size,))
            data = np.zeros(int(off[-1]), dtype=np.uint8)
            take = (np.arange(12)[None, :] < lens[:, None]) & ~longm[:, None]
            pos = (off[:-1, None] + np.arange(12)[None, :])[take]
            data[pos] = views[:, 4:16][take]
            src = views[:, 12:16].copy().view("<u4").reshape(n).astype(np.int64)
            for i in np.nonzero(longm)[0]:
                data[off[i]:off[i + 1]] = blob[src[i]:src[i] + lens[i]]
        else:
            take = np.arange(12)[None, :] < lens[:, None]
            data = views[:, 4:16][take]
        arr = pa.Array.from_buffers(pa.large_binary() if binary else pa.large_string(), n,
                                    [None, pa.py_buffer(off.tobytes()), pa.py_buffer(data.tobytes())])
        if valid is not None:
            arr = pa.array(arr.to_pylist(), type=arr.type, mask=~valid) if n < 1_000_000 else pa.compute.if_else(pa.array(valid), arr, None)
    else:
        arr = pa.array([], type=pa.large_binary() if binary else pa.large_string())
    for st in (c_array, c_schema):
        if st.release:
            C.CFUNCTYPE(None, C.c_void_p)(st.release)(C.addressof(st))
    return arr


def _import_columns(out_arrays, out_schemas, n) -> tuple:
    names, cols = [], []
    for i in range(n):
        name = out_schemas[i].name.decode() if out_schemas[i].name else ""
        if out_schemas[i].format in (b"vu", b"vz"):
            arr = _import_view_array(out_arrays[i], out_schemas[i])
        else:
            arr = pa.Array._import_from_c(C.addressof(out_arrays[i]), C.addressof(out_schemas[i]))
        names.append(name)
        cols.append(arr)
    return names, cols


def _arrow_format(t: pa.DataType) -> bytes:
    """Arrow C Data Interface format string of the column types this path reads."""
    simple = {pa.int8(): b"c", pa.uint8(): b"C", pa.int16(): b"s", pa.uint16(): b"S", pa.int32(): b"i", pa.uint32(): b"I",
              pa.int64(): b"l", pa.uint64(): b"L", pa.float32(): b"f", pa.float64(): b"g", pa.date32(): b"tdD",
              pa.string_view(): b"vu", pa.binary_view(): b"vz", pa.bool_(): b"b"}
    if t in simple:
        return simple[t]
    unit = {"s": b"s", "ms": b"m", "us": b"u", "ns": b"n"}
    if pa.types.is_timestamp(t):
        return b"ts" + unit[t.unit] + b":" + (t.tz.encode() if t.tz else b"")
    if pa.types.is_duration(t):
        return b"tD" + unit[t.unit]
    raise NotImplementedError(f"device frame column of type {t}")


# ---- plan -> PwQuery -----------------------------------------------------------------------------------
def _physical_scalar(value, typ: pa.DataType):
    if isinstance(value, _dt.datetime):
        if pa.types.is_timestamp(typ):
            return pa.scalar(value, type=pa.timestamp(typ.unit)).cast(pa.int64()).as_py()
        if pa.types.is_date32(typ):
            return pa.scalar(value.date(), type=pa.date32()).cast(pa.int32()).as_py()
    if isinstance(value, _dt.date):
        if pa.types.is_date32(typ):
            return pa.scalar(value, type=pa.date32()).cast(pa.int32()).as_py()
        if pa.types.is_timestamp(typ):
            return pa.scalar(_dt.datetime(value.year, value.month, value.day), type=pa.timestamp(typ.unit)).cast(pa.int64()).as_py()
    if isinstance(value, _dt.timedelta):
        return pa.scalar(value, type=typ).cast(pa.int64()).as_py()
    return value


def _index_unit_ns(typ: pa.DataType) -> int:
    if pa.types.is_timestamp(typ):
        return {"ns": 1, "us": 1_000, "ms": 1_000_000, "s": 1_000_000_000}[typ.unit]
    if pa.types.is_date32(typ):
        return 86_400_000_000_000
    return 1


class _BuiltQuery:
    """Owns every ctypes buffer a PwQuery points to."""

    def __init__(self, schema: pa.Schema, plan: P.GroupByPlan, flags: int = 0, hot_table_slots: int = 0,
                 initial_table_slots: int = 0, row_offset: int = 0):
        names = schema.names
        idx = {n: i for i, n in enumerate(names)}

        def col(name):
            if name not in idx:
                raise KeyError(f"column {name!r} not found")
            return idx[name]

        self.preds = (PwPredicate * max(1, len(plan.predicates)))()
        for i, p in enumerate(plan.predicates):
            typ = schema.field(p.col).type
            v = _physical_scalar(p.value, typ)
            self.preds[i].column = col(p.col)
            self.preds[i].op = P.CMP_OPS[p.op]
            if pa.types.is_floating(typ):
                self.preds[i].scalar_is_float = 1
                self.preds[i].scalar.f = float(v)
            elif isinstance(v, float):
                if v != int(v):
                    raise NotImplementedError("non-integral float scalar compared with an integer column")
                self.preds[i].scalar.i = int(v)
            elif pa.types.is_unsigned_integer(typ):
                self.preds[i].scalar.u = int(v)
            else:
                self.preds[i].scalar.i = int(v)
        self.keys = (C.c_int32 * max(1, len(plan.keys)))(*[col(k) for k in plan.keys])
        self.aggs = (PwAgg * max(1, len(plan.aggs)))()
        self._names = []
        for i, a in enumerate(plan.aggs):
            self.aggs[i].kind = P.AGG_KINDS[a.kind]
            self.aggs[i].ddof = int(getattr(a, "ddof", 0) or 0)
            nm = a.name.encode()
            self._names.append(nm)
            self.aggs[i].name = nm
            if a.expr is None:
                self.aggs[i].column = -1
            elif a.expr.factors is None:
                self.aggs[i].column = col(a.expr.col)
            else:
                if len(a.expr.factors) > PW_MAX_FACTORS:
                    raise NotImplementedError("more than 4 factors in a product expression")
                self.aggs[i].column = -1
                self.aggs[i].n_factors = len(a.expr.factors)
                for j, f in enumerate(a.expr.factors):
                    self.aggs[i].factors[j].a = f.a
                    self.aggs[i].factors[j].b = f.b
                    self.aggs[i].factors[j].column = col(f.col)
        self.dyn = None
        if plan.dynamic is not None:
            d = plan.dynamic
            typ = schema.field(d.index_column).type
            every, period, offset = d.every, d.period, d.offset
            if getattr(plan, "_durations_in_ns", True) and not pa.types.is_integer(typ):
                u = _index_unit_ns(typ)
                for x in (every, period, offset):
                    if x % u:
                        raise ValueError("duration is not a multiple of the index column's time unit")
                every, period, offset = every // u, period // u, offset // u
            self.dyn = PwDynamic(col(d.index_column), P.CLOSED[d.closed], P.LABEL[d.label], int(d.include_boundaries),
                                 every, period, offset)
        q = PwQuery()
        q.abi_version = 1
        q.maintain_order = int(plan.maintain_order)
        q.n_predicates = len(plan.predicates)
        q.n_keys = len(plan.keys)
        q.n_aggs = len(plan.aggs)
        q.hot_table_slots = hot_table_slots
        q.predicates = C.cast(self.preds, C.POINTER(PwPredicate))
        q.key_columns = C.cast(self.keys, C.POINTER(C.c_int32))
        q.aggs = C.cast(self.aggs, C.POINTER(PwAgg))
        q.dynamic = C.pointer(self.dyn) if self.dyn is not None else None
        q.flags = flags | (FLAG_KEYS_SORTED if getattr(plan, "keys_sorted", False) and plan.dynamic is None else 0)
        q.row_offset = row_offset
        q.initial_table_slots = initial_table_slots
        self.q = q


# ---- resident frames -----------------------------------------------------------------------------------
class DeviceFrame:
    """Columns resident in HBM (``pw_b200_frame_upload``)."""

    def __init__(self, table: pa.Table):
        L = lib()
        table, self._dicts = _physical_table(table)
        self.table_schema = table.schema
        arrays = [_to_abi_array(table.column(i)) for i in range(table.num_columns)]
        ex = _Exported(arrays)
        ex.set_names(table.column_names)
        h = C.c_void_p()
        try:
            _check(L.pw_b200_frame_upload(ex.arr_ptrs, ex.sch_ptrs, ex.n, C.byref(h)))
        finally:
            ex.release()
        self.handle = h
        self.num_rows = table.num_rows
        self._queries = {}

    @classmethod
    def from_device(cls, columns) -> "DeviceFrame":
        """Zero-copy frame over buffers that already live in HBM (``pw_b200_frame_from_device``).

        columns: list of ``(name, arrow_type, length, values_ptr, validity_ptr_or_0, null_count, extra)`` where the
        pointers are DEVICE addresses laid out as the Arrow spec prescribes (values: element 0 = row 0; validity:
        LSB-first bitmap) and ``extra`` keeps the owners (e.g. torch tensors) alive as long as the frame."""
        L = lib()
        n = len(columns)
        self = cls.__new__(cls)
        c_arrays = (ArrowArray * n)()
        c_schemas = (ArrowSchema * n)()
        keep = []
        fields = []
        for i, (name, typ, length, vptr, nptr, nulls, extra) in enumerate(columns):
            fmt = _arrow_format(typ)
            nbuf = 4 if fmt in (b"vu", b"vz") else 2
            bufs = (C.c_void_p * nbuf)(nptr or None, vptr, None, None) if nbuf == 4 else (C.c_void_p * 2)(nptr or None, vptr)
            nm = name.encode()
            keep.append((bufs, nm, extra))
            c_arrays[i].length = length
            c_arrays[i].null_count = nulls
            c_arrays[i].offset = 0
            c_arrays[i].n_buffers = nbuf
            c_arrays[i].buffers = C.cast(bufs, C.POINTER(C.c_void_p))
            c_schemas[i].format = fmt
            c_schemas[i].name = nm
            c_schemas[i].flags = 2
            fields.append(pa.field(name, typ))
        arr_ptrs = (C.c_void_p * n)(*[C.addressof(c_arrays[i]) for i in range(n)])
        sch_ptrs = (C.c_void_p * n)(*[C.addressof(c_schemas[i]) for i in range(n)])
        h = C.c_void_p()
        _check(L.pw_b200_frame_from_device(arr_ptrs, sch_ptrs, n, C.byref(h)))
        self.handle = h
        self.table_schema = pa.schema(fields)
        self.num_rows = columns[0][2] if columns else 0
        self._queries = {}
        self._keep = keep
        return self

    def free(self):
        if self.handle:
            lib().pw_b200_frame_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def group_by(self, plan: P.GroupByPlan, **opts) -> pa.Table:
        L = lib()
        # the lowered PwQuery of a plan object is reused across calls (plans are immutable once built)
        ck = (id(plan), tuple(sorted(opts.items())))
        cache = self.__dict__.setdefault("_queries", {})
        hit = cache.get(ck)
        if hit is None or hit[0] is not plan:
            if len(cache) > 64:
                cache.clear()
            hit = (plan, _BuiltQuery(self.table_schema, plan, **opts))
            cache[ck] = hit
        bq = hit[1]
        cap = len(plan.keys) + len(plan.aggs) + 4
        out_arrays = (ArrowArray * cap)()
        out_schemas = (ArrowSchema * cap)()
        n_out = C.c_size_t(cap)
        _check(L.pw_b200_frame_groupby(C.byref(bq.q), self.handle, out_arrays, out_schemas, C.byref(n_out)))
        names, cols = _import_columns(out_arrays, out_schemas, n_out.value)
        cols = _restore_string_types(names, cols, self.table_schema, plan.keys, getattr(self, "_dicts", None))
        return pa.Table.from_arrays(cols, names=names)


def _physical_table(table: pa.Table):
    """Categorical / Enum columns (Arrow dictionary arrays) cross the boundary as their PHYSICAL index column, as the
    reference's operators see them (polars-expr/src/hash_keys.rs:32,83-89: group identity is the u8/u16/u32 code; the
    caller reattaches the categories).  -> (table with index columns, {column name: dictionary values})."""
    dicts = {}
    if not any(pa.types.is_dictionary(f.type) for f in table.schema):
        return table, dicts
    table = table.unify_dictionaries().combine_chunks()
    for i, f in enumerate(table.schema):
        if pa.types.is_dictionary(f.type):
            col = table.column(i)
            arr = col.chunk(0) if col.num_chunks else pa.DictionaryArray.from_arrays(pa.array([], type=f.type.index_type), pa.array([], type=f.type.value_type))
            dicts[f.name] = (arr.dictionary, f.type)
            table = table.set_column(i, f.name, arr.indices)
    return table, dicts


def _restore_string_types(names, cols, schema: pa.Schema, key_names, dicts=None) -> list:
    out = []
    for n, c in zip(names, cols):
        if dicts and n in key_names and n in dicts:
            values, typ = dicts[n]
            c = pa.chunked_array([pa.DictionaryArray.from_arrays(ch.cast(typ.index_type), values) for ch in (c.chunks if isinstance(c, pa.ChunkedArray) else [c])], type=typ)
            out.append(c)
            continue
        if n in key_names and n in schema.names:
            t = schema.field(n).type
            if c.type != t and _is_stringlike(t) and not (pa.types.is_string_view(t) or pa.types.is_binary_view(t)):
                c = c.cast(t)
        out.append(c)
    return out


# ---- the calls the LazyFrame mirror makes ----------------------------------------------------------------
_DEFAULT_OPTS: dict = {}


def set_default_options(**opts):
    """Strategy overrides for tests (the analogue of POLARS_FORCE_PARTITION etc.)."""
    _DEFAULT_OPTS.clear()
    _DEFAULT_OPTS.update(opts)


def run_group_by(table: pa.Table, plan: P.GroupByPlan, **opts) -> pa.Table:
    """Host buffers in, host result out: one ``pw_b200_filter_groupby_agg`` call."""
    L = lib()
    o = dict(_DEFAULT_OPTS)
    o.update(opts)
    # only ship the columns the query touches
    used = []
    for n in ([p.col for p in plan.predicates] + list(plan.keys)
              + ([plan.dynamic.index_column] if plan.dynamic else [])
              + [c for a in plan.aggs if a.expr is not None for c in a.expr.columns()]):
        if n not in used:
            used.append(n)
    if not used:
        used = table.column_names[:1]
    sub, dicts = _physical_table(table.select(used))
    arrays = [_to_abi_array(sub.column(i)) for i in range(sub.num_columns)]
    ex = _Exported(arrays)
    ex.set_names(sub.column_names)
    bq = _BuiltQuery(sub.schema, plan, **o)
    cap = len(plan.keys) + len(plan.aggs) + 4
    out_arrays = (ArrowArray * cap)()
    out_schemas = (ArrowSchema * cap)()
    n_out = C.c_size_t(cap)
    try:
        _check(L.pw_b200_filter_groupby_agg(C.byref(bq.q), ex.arr_ptrs, ex.sch_ptrs, ex.n, out_arrays, out_schemas,
                                            C.byref(n_out)))
    finally:
        ex.release()
    names, cols = _import_columns(out_arrays, out_schemas, n_out.value)
    cols = _restore_string_types(names, cols, sub.schema, plan.keys, dicts)
    return pa.Table.from_arrays(cols, names=names)


def run_filter(table: pa.Table, preds) -> pa.Table:
    """FilterExec alone: ``pw_b200_filter``."""
    L = lib()
    if not preds:
        return table
    arrays = [_to_abi_array(table.column(i)) for i in range(table.num_columns)]
    ex = _Exported(arrays)
    ex.set_names(table.column_names)
    plan = P.GroupByPlan(predicates=list(preds))
    bq = _BuiltQuery(table.schema, plan)
    n = table.num_columns
    out_arrays = (ArrowArray * n)()
    out_schemas = (ArrowSchema * n)()
    try:
        _check(L.pw_b200_filter(bq.preds, len(preds), ex.arr_ptrs, ex.sch_ptrs, n, out_arrays, out_schemas))
    finally:
        ex.release()
    names, cols = _import_columns(out_arrays, out_schemas, n)
    cols = _restore_string_types(names, cols, table.schema, table.column_names)
    return pa.Table.from_arrays(cols, names=names)


def filter_select(table: pa.Table, preds) -> "pa.Array":
    """predicate -> compacted selection vector (ascending UInt32 row ids): ``pw_b200_frame_filter_select``."""
    L = lib()
    frame = DeviceFrame(table)
    try:
        plan = P.GroupByPlan(predicates=list(preds))
        bq = _BuiltQuery(table.schema, plan)
        oa, os_ = (ArrowArray * 1)(), (ArrowSchema * 1)()
        n_sel = C.c_int64(0)
        _check(L.pw_b200_frame_filter_select(bq.preds, len(preds), frame.handle, oa, os_, C.byref(n_sel)))
        _, cols = _import_columns(oa, os_, 1)
        return cols[0]
    finally:
        frame.free()


def group_tuples(table: pa.Table, keys, maintain_order: bool = True):
    """GroupsIdx: (first[g], offsets[g+1], row_ids) — ``pw_b200_frame_group_tuples``."""
    L = lib()
    frame = DeviceFrame(table)
    try:
        kc = (C.c_int32 * max(1, len(keys)))(*[table.column_names.index(k) for k in keys])
        oa, os_ = (ArrowArray * 3)(), (ArrowSchema * 3)()
        _check(L.pw_b200_frame_group_tuples(frame.handle, kc, len(keys), int(maintain_order), C.byref(oa[0]), C.byref(oa[1]),
                                            C.byref(oa[2]), os_))
        _, cols = _import_columns(oa, os_, 3)
        return cols[0], cols[1], cols[2]
    finally:
        frame.free()


def group_slices(table: pa.Table, keys):
    """GroupsSlice of sorted key columns: (first[g], len[g]) per run of equal keys — ``pw_b200_frame_group_slices``
    (polars-arrow/src/legacy/kernels/sort_partition.rs:168 partition_to_groups)."""
    L = lib()
    frame = DeviceFrame(table)
    try:
        kc = (C.c_int32 * max(1, len(keys)))(*[table.column_names.index(k) for k in keys])
        oa, os_ = (ArrowArray * 2)(), (ArrowSchema * 2)()
        _check(L.pw_b200_frame_group_slices(frame.handle, kc, len(keys), C.byref(oa[0]), C.byref(oa[1]), os_))
        _, cols = _import_columns(oa, os_, 2)
        return cols[0], cols[1]
    finally:
        frame.free()
