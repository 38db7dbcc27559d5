"""A Python restatement of the CALL SEQUENCE of the reference's expression-plugin loader
(crates/polars-plan/src/plans/aexpr/function_expr/plugin.rs:75-142, 183-205) so the `_polars_plugin_*` shim of
libpolarway_b200.so can be driven exactly the way an unmodified Polars would drive it:

    version check -> inputs exported as SeriesExport (one ArrowArray per chunk) -> kwargs pickled with protocol 5
    (py-polars/src/polars/plugins.py:111-120) -> symbol call -> `private_data == NULL` means failure and the message
    is fetched from `_polars_plugin_get_last_error_message`.

Used by the tests; INTEGRATION.md shows the one-line `register_plugin_function` call that does the same from Polars.
"""
from __future__ import annotations

import ctypes as C
import pickle

import pyarrow as pa

from . import engine as E
from . import plan as P


class SeriesExport(C.Structure):
    _fields_ = [("field", C.POINTER(E.ArrowSchema)), ("arrays", C.POINTER(C.POINTER(E.ArrowArray))), ("len", C.c_size_t),
                ("release", C.c_void_p), ("private_data", C.c_void_p)]


class CallerContext(C.Structure):
    _fields_ = [("bitflags", C.c_uint64)]


def plan_to_kwargs(schema: pa.Schema, plan: P.GroupByPlan, **extra) -> dict:
    """The kwargs dict the shim understands (column references are positions in the input list)."""
    bq = E._BuiltQuery(schema, plan)
    idx = {n: i for i, n in enumerate(schema.names)}
    kw = {
        "maintain_order": bool(plan.maintain_order),
        "keys": [idx[k] for k in plan.keys],
        "predicates": [(bq.preds[i].column, bq.preds[i].op,
                        float(bq.preds[i].scalar.f) if bq.preds[i].scalar_is_float else int(bq.preds[i].scalar.i))
                       for i in range(len(plan.predicates))],
        "aggs": [(a.name, P.AGG_KINDS[a.kind],
                  None if a.expr is None or a.expr.factors is not None else idx[a.expr.col],
                  None if a.expr is None or a.expr.factors is None else [(f.a, f.b, idx[f.col]) for f in a.expr.factors],
                  int(getattr(a, "ddof", 0) or 0))
                 for a in plan.aggs],
    }
    if bq.dyn is not None:
        kw["dynamic"] = {"index_column": bq.dyn.index_column, "closed": bq.dyn.closed, "label": bq.dyn.label,
                         "include_boundaries": bq.dyn.include_boundaries, "every": bq.dyn.every, "period": bq.dyn.period,
                         "offset": bq.dyn.offset}
    kw.update(extra)
    return kw


def _check_version(L):
    L._polars_plugin_get_version.restype = C.c_uint32
    v = L._polars_plugin_get_version()
    major, minor = v >> 16, v & 0xFFFF
    if major != 0:
        raise RuntimeError(f"this polars version doesn't support plugin version: {major}")  # plugin.rs:139-141
    return major, minor


def plugin_field(schema: pa.Schema, kwargs: dict) -> pa.Schema:
    """`_polars_plugin_field_<fn>`: output dtype inference without data (works without a GPU)."""
    L = E.lib()
    _check_version(L)
    n = len(schema)
    fields = (E.ArrowSchema * n)()
    keep = []
    for i, f in enumerate(schema):
        t = f.type
        if E._is_stringlike(t):
            fields[i].format = b"vu"
        else:
            pa.field(f.name, t)._export_to_c(C.addressof(fields[i]))
        nm = f.name.encode()
        keep.append(nm)
        fields[i].name = nm
    out = E.ArrowSchema()
    blob = pickle.dumps(kwargs, protocol=5)
    L._polars_plugin_field_filter_groupby_agg(fields, C.c_size_t(n), C.byref(out), blob, C.c_size_t(len(blob)))
    if not out.release:
        L._polars_plugin_get_last_error_message.restype = C.c_char_p
        raise E.PolarwayError(-1, "the plugin failed with message: " + L._polars_plugin_get_last_error_message().decode())
    res = []
    for i in range(out.n_children):
        ch = out.children[i].contents
        res.append((ch.name.decode(), ch.format.decode()))
    C.CFUNCTYPE(None, C.c_void_p)(out.release)(C.addressof(out))
    return res


def call_plugin(table: pa.Table, kwargs: dict, release_log: list | None = None) -> pa.Table:
    """`_polars_plugin_<fn>`: every column as a SeriesExport (chunks preserved), result = one Struct series.
    release_log: when given, every input SeriesExport carries a real ``release`` callback (as ``export_column`` gives it,
    polars-ffi/src/version_0.rs:37-60) that appends the input's position; the callee owns the inputs and must call each
    exactly once (plugin.rs:127-130: the caller ``mem::forget``s them)."""
    L = E.lib()
    _check_version(L)
    n = table.num_columns
    inputs = (SeriesExport * n)()
    keep = []
    for i in range(n):
        col = table.column(i)
        chunks = col.chunks if isinstance(col, pa.ChunkedArray) else [col]
        if not chunks:
            chunks = [pa.array([], type=col.type)]
        ex = E._Exported(list(chunks))
        for k in range(len(chunks)):
            ex.c_schemas[k].name = table.column_names[i].encode()
        ptrs = (C.POINTER(E.ArrowArray) * len(chunks))(*[C.pointer(ex.c_arrays[k]) for k in range(len(chunks))])
        inputs[i].field = C.pointer(ex.c_schemas[0])
        inputs[i].arrays = ptrs
        inputs[i].len = len(chunks)
        inputs[i].release = None       # the structs are owned by this Python frame; released below
        if release_log is not None:
            def _mk(pos, exported):
                def _release(ptr):
                    release_log.append(pos)
                    exported.release()
                    C.cast(ptr, C.POINTER(SeriesExport)).contents.release = None
                return C.CFUNCTYPE(None, C.c_void_p)(_release)
            cb = _mk(i, ex)
            inputs[i].release = C.cast(cb, C.c_void_p)
            keep.append((ex, ptrs, cb))
            continue
        keep.append((ex, ptrs))
    blob = pickle.dumps(kwargs, protocol=5)
    out = SeriesExport()
    ctx = CallerContext(0)
    L._polars_plugin_filter_groupby_agg(inputs, C.c_size_t(n), blob, C.c_size_t(len(blob)), C.byref(out), C.byref(ctx))
    if release_log is None:
        for ex, *_ in keep:
            ex.release()
    if not out.private_data:   # plugin.rs:132-138
        L._polars_plugin_get_last_error_message.restype = C.c_char_p
        raise E.PolarwayError(-1, "the plugin failed with message: " + L._polars_plugin_get_last_error_message().decode())
    arr = out.arrays[0].contents
    sch = out.field.contents
    names, cols = [], []
    for i in range(arr.n_children):
        ca, cs = arr.children[i].contents, sch.children[i].contents
        names.append(cs.name.decode())
        if cs.format in (b"vu", b"vz"):
            cols.append(E._import_view_array(ca, cs))
        else:
            cols.append(pa.Array._import_from_c(C.addressof(ca), C.addressof(cs)))
    C.CFUNCTYPE(None, C.c_void_p)(out.release)(C.addressof(out))
    return pa.Table.from_arrays(cols, names=names)
