"""Python driver for the CPU oracle (oracle/pw_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It takes the same logical plan (polaroid_b200.plan.GroupByPlan) as the CUDA
engine and returns a pyarrow Table with the reference's output schema rules:

* sum:   i8/i16/u8/u16 -> Int64, everything else keeps its dtype   (polars-expr/src/reduce/sum.rs:40-47)
* mean:  Float64 (Float32 stays Float32; Date -> Datetime[us]; Datetime/Duration keep type)
         (polars-expr/src/reduce/mean.rs:29-80)
* min/max/first/last keep the input dtype; count/len -> UInt32 (IdxSize, polars-utils/src/index.rs:9-11)
* group_by_dynamic output = [keys..., (_lower_boundary, _upper_boundary), index, aggs...]
  (polars-mem-engine/src/executors/group_by_dynamic.rs:74-82), key slices in ascending key order,
  nulls first (executors/group_by_rolling.rs:17-58).

Parity status: PINNED by tests/golden (see tests/golden/README.md).
"""
from __future__ import annotations

import ctypes as C
import datetime as _dt
import os
import subprocess

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

I64, U64, F64, F32 = 0, 1, 2, 3
_CMP = {"eq": 0, "ne": 1, "lt": 2, "le": 3, "gt": 4, "ge": 5}
_KIND = {"sum": 0, "mean": 1, "min": 2, "max": 3, "count": 4, "len": 5, "first": 6, "last": 7,
         "var": 8, "std": 9, "first_non_null": 10, "last_non_null": 11, "null_count": 12,
         "bitwise_and": 13, "bitwise_or": 14, "bitwise_xor": 15, "any": 16, "all": 17}
_CLOSED = {"left": 0, "right": 1, "both": 2, "none": 3}


class OrcAgg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("vclass", C.c_int32), ("values", C.c_void_p), ("valid", C.c_void_p),
                ("ddof", C.c_int32), ("is_bool", C.c_int32)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libpw_oracle.so")
    src = os.path.join(_HERE, "pw_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpw_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_groupby.restype = C.c_void_p
        L.orc_groupby.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_result_ngroups.restype = C.c_int64
        L.orc_result_ngroups.argtypes = [C.c_void_p]
        L.orc_result_keys.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_result_first_rows.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_result_agg.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_result_free.argtypes = [C.c_void_p]
        L.orc_predicate_mask.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_group_ids.restype = C.c_int64
        L.orc_group_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_partition_to_groups.restype = C.c_int64
        L.orc_partition_to_groups.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_group_by_windows.restype = C.c_int64
        L.orc_group_by_windows.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_agg_slices.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def max_threads() -> int:
    return int(lib().orc_max_threads())


# ---- Arrow -> canonical numpy ---------------------------------------------------------------
def _combine(col):
    if isinstance(col, pa.ChunkedArray):
        col = col.combine_chunks() if col.num_chunks != 1 else col.chunk(0)
    return col


def _physical(arr: pa.Array) -> pa.Array:
    t = arr.type
    if pa.types.is_date32(t):
        return arr.cast(pa.int32())
    if pa.types.is_timestamp(t) or pa.types.is_duration(t) or pa.types.is_date64(t) or pa.types.is_time64(t):
        return arr.cast(pa.int64())
    if pa.types.is_time32(t):
        return arr.cast(pa.int32())
    return arr


def _valid_bytes(arr: pa.Array):
    if arr.null_count == 0:
        return None
    return np.asarray(pc.is_valid(arr).to_numpy(zero_copy_only=False), dtype=np.uint8)


def widen(arr: pa.Array):
    """-> (values ndarray (int64/uint64/float64/float32), vclass, valid bytes|None)."""
    arr = _physical(_combine(arr))
    t = arr.type
    valid = _valid_bytes(arr)
    if pa.types.is_boolean(t):
        filled = arr.fill_null(False) if arr.null_count else arr
        v = np.asarray(filled.to_numpy(zero_copy_only=False)).astype(np.uint64)
        return v, U64, valid
    filled = arr.fill_null(0) if arr.null_count else arr
    v = filled.to_numpy(zero_copy_only=False)
    if pa.types.is_float64(t):
        return np.ascontiguousarray(v, dtype=np.float64), F64, valid
    if pa.types.is_float32(t):
        return np.ascontiguousarray(v, dtype=np.float32), F32, valid
    # widening is exact; a column that already has the 8-byte class type is used in place (no copy)
    if pa.types.is_unsigned_integer(t):
        return np.ascontiguousarray(v, dtype=np.uint64), U64, valid
    if pa.types.is_signed_integer(t):
        return np.ascontiguousarray(v, dtype=np.int64), I64, valid
    raise NotImplementedError(f"value dtype {t}")


def key_words(arr: pa.Array):
    """Key column -> (uint64 words, valid bytes|None).  Strings are dictionary-encoded: equal
    strings <-> equal codes, which is all group identity needs."""
    arr = _combine(arr)
    t = arr.type
    if pa.types.is_dictionary(t):
        # Categorical / Enum: the group identity is the physical code (polars-expr/src/hash_keys.rs:32,83-89)
        idx = arr.indices
        valid = _valid_bytes(idx)
        return np.ascontiguousarray(idx.fill_null(0).to_numpy(zero_copy_only=False).astype(np.uint64)), valid
    if pa.types.is_string(t) or pa.types.is_large_string(t) or pa.types.is_string_view(t) or pa.types.is_binary(t):
        if pa.types.is_string_view(t):
            arr = arr.cast(pa.large_string())
        d = pc.dictionary_encode(arr)
        idx = d.indices
        valid = _valid_bytes(idx)
        v = idx.fill_null(0).to_numpy(zero_copy_only=False).astype(np.uint64)
        return np.ascontiguousarray(v), valid
    if pa.types.is_floating(t):
        v, _, valid = widen(arr)
        v = v.astype(np.float64)
        v = np.where(v == 0.0, 0.0, v)          # -0.0 == 0.0 (TotalEq canonicalisation)
        v = np.where(np.isnan(v), np.nan, v)     # one NaN
        return np.ascontiguousarray(v).view(np.uint64), valid
    v, _, valid = widen(arr)
    return np.ascontiguousarray(v).view(np.uint64), valid


def scalar_to_physical(value, typ: pa.DataType):
    """Python scalar -> physical number in the column's unit."""
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


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def predicate_mask(table: pa.Table, preds) -> np.ndarray | None:
    if not preds:
        return None
    n = table.num_rows
    mask = np.ones(n, dtype=np.uint8)
    for p in preds:
        arr = table.column(p.col)
        v, vc, valid = widen(arr)
        val = scalar_to_physical(p.value, _combine(arr).type)
        if vc == F32:
            v, vc = v.astype(np.float64), F64
        sc = np.array([val], dtype={I64: np.int64, U64: np.uint64, F64: np.float64}[vc])
        m = np.empty(n, dtype=np.uint8)
        lib().orc_predicate_mask(vc, _ptr(v), _ptr(valid), n, _CMP[p.op], _ptr(sc), _ptr(m))
        mask &= m
    return mask


def eval_value_expr(table: pa.Table, expr):
    """-> (values, vclass, valid, source arrow type or None for computed f64)."""
    if expr.factors is None:
        arr = _combine(table.column(expr.col))
        v, vc, valid = widen(arr)
        return v, vc, valid, arr.type
    prod = None
    valid = None
    for f in expr.factors:
        v, vc, vl = widen(table.column(f.col))
        x = v.astype(np.float64)
        t = x if f.b == 1.0 else (np.float64(f.b) * x)
        u = t if f.a == 0.0 else (np.float64(f.a) + t)
        prod = u if prod is None else prod * u
        if vl is not None:
            valid = vl if valid is None else (valid & vl)
    return np.ascontiguousarray(prod), F64, valid, None


# ---- output typing --------------------------------------------------------------------------
def _out_array(kind: str, bits: np.ndarray, ok: np.ndarray, vclass: int, src_type):
    mask = None if ok.all() else ~ok.astype(bool)
    if kind in ("count", "len", "null_count"):
        return pa.array(bits.astype(np.uint32), type=pa.uint32())
    if kind in ("any", "all"):
        return pa.array(bits.astype(bool), type=pa.bool_())
    if kind in ("var", "std"):   # reduce/var_std.rs:104-140: Float32 in -> Float32 out, everything else Float64
        vals = bits.view(np.float64)
        if src_type is not None and pa.types.is_float32(src_type):
            return pa.array(vals.astype(np.float32), mask=mask, type=pa.float32())
        return pa.array(vals, mask=mask, type=pa.float64())
    if kind.startswith("bitwise_") and src_type is not None and pa.types.is_boolean(src_type):
        return pa.array(bits.astype(bool), mask=mask, type=pa.bool_())
    if kind == "mean":
        vals = bits.view(np.float64)
        if src_type is not None and pa.types.is_float32(src_type):
            return pa.array(vals.astype(np.float32), mask=mask, type=pa.float32())
        if src_type is not None and pa.types.is_date32(src_type):
            us = (vals * 86_400_000_000.0)
            us = np.where(ok.astype(bool), us, 0).astype(np.int64)
            return pa.array(us, mask=mask, type=pa.int64()).cast(pa.timestamp("us"))
        if src_type is not None and (pa.types.is_timestamp(src_type) or pa.types.is_duration(src_type)):
            iv = np.where(ok.astype(bool), vals, 0).astype(np.int64)
            return pa.array(iv, mask=mask, type=pa.int64()).cast(src_type)
        return pa.array(vals, mask=mask, type=pa.float64())
    # sum / min / max / first / last
    if src_type is None or pa.types.is_float64(src_type):
        return pa.array(bits.view(np.float64), mask=mask, type=pa.float64())
    if pa.types.is_float32(src_type):
        return pa.array(bits.view(np.float64).astype(np.float32), mask=mask, type=pa.float32())
    phys = _physical(pa.array([], type=src_type)).type
    if kind == "sum":
        if phys in (pa.int8(), pa.int16(), pa.uint8(), pa.uint16()):
            return pa.array(bits.view(np.int64), mask=mask, type=pa.int64())
        if pa.types.is_boolean(src_type):
            return pa.array(bits.astype(np.uint32), mask=mask, type=pa.uint32())
    npdt = phys.to_pandas_dtype()
    with np.errstate(over="ignore"):
        narrowed = bits.astype(npdt) if vclass == U64 else bits.view(np.int64).astype(npdt)
    out = pa.array(narrowed, mask=mask, type=phys)
    if phys != src_type and not (kind == "sum" and pa.types.is_date32(src_type)):
        out = out.cast(src_type)
    return out


def _take_keys(table, keys, first_rows):
    idx = pa.array(first_rows, type=pa.int64())
    return [(_combine(table.column(k)).take(idx)) for k in keys]


def _make_aggs(table, aggs):
    n = table.num_rows
    c_aggs = (OrcAgg * max(1, len(aggs)))()
    keep, meta = [], []
    memo = {}   # every value expression is widened / evaluated once, however many aggregates read it
    for i, a in enumerate(aggs):
        if a.expr is None:
            if None not in memo:
                memo[None] = (np.zeros(1, dtype=np.int64), I64, None, None)   # len() reads no values
            v, vc, valid, st = memo[None]
        else:
            if a.expr not in memo:
                memo[a.expr] = eval_value_expr(table, a.expr)
            v, vc, valid, st = memo[a.expr]
        keep.append((v, valid))
        c_aggs[i].kind = _KIND[a.kind]
        c_aggs[i].vclass = vc
        c_aggs[i].values = v.ctypes.data
        c_aggs[i].valid = valid.ctypes.data if valid is not None else None
        c_aggs[i].ddof = int(getattr(a, "ddof", 0) or 0)
        c_aggs[i].is_bool = 1 if (st is not None and pa.types.is_boolean(st)) else 0
        meta.append((vc, st))
    return c_aggs, keep, meta


def group_by(table: pa.Table, plan, n_threads: int = 1) -> pa.Table:
    """filter -> group_by -> agg.  Output rows in first-occurrence order (== maintain_order)."""
    if plan.dynamic is not None:
        return group_by_dynamic(table, plan)
    L = lib()
    n = table.num_rows
    sel = predicate_mask(table, plan.predicates)
    kws = [key_words(table.column(k)) for k in plan.keys]
    nk = len(kws)
    kp = (C.c_void_p * max(1, nk))(*[w.ctypes.data for w, _ in kws])
    kv = (C.c_void_p * max(1, nk))(*[(v.ctypes.data if v is not None else None) for _, v in kws])
    c_aggs, keep, meta = _make_aggs(table, plan.aggs)
    res = L.orc_groupby(kp, kv, nk, n, _ptr(sel), c_aggs, len(plan.aggs), n_threads)
    try:
        g = L.orc_result_ngroups(res)
        first = np.empty(g, dtype=np.int64)
        L.orc_result_first_rows(res, _ptr(first))
        cols = dict(zip(plan.keys, _take_keys(table, plan.keys, first)))
        for i, a in enumerate(plan.aggs):
            bits = np.empty(g, dtype=np.uint64)
            ok = np.empty(g, dtype=np.uint8)
            L.orc_result_agg(res, c_aggs, i, _ptr(bits), _ptr(ok))
            cols[a.name] = _out_array(a.kind, bits, ok, meta[i][0], meta[i][1])
    finally:
        L.orc_result_free(res)
    return pa.table(cols)


def filter_table(table: pa.Table, preds) -> pa.Table:
    """FilterExec: predicate -> mask (null => false) -> every column compacted."""
    m = predicate_mask(table, preds)
    if m is None:
        return table
    return table.filter(pa.array(m.astype(bool)))


def group_tuples(table: pa.Table, keys, maintain_order: bool = True):
    """GroupsIdx: (first[g], offsets[g+1], row_ids) with groups in first-occurrence order."""
    n = table.num_rows
    kws = [key_words(table.column(k)) for k in keys]
    nk = len(kws)
    kp = (C.c_void_p * max(1, nk))(*[w.ctypes.data for w, _ in kws])
    kv = (C.c_void_p * max(1, nk))(*[(v.ctypes.data if v is not None else None) for _, v in kws])
    gid = np.empty(n, dtype=np.int32)
    first = np.empty(max(n, 1), dtype=np.int64)
    g = lib().orc_group_ids(kp, kv, nk, n, None, _ptr(gid), _ptr(first))
    order = np.argsort(gid, kind="stable")
    counts = np.bincount(gid, minlength=g)
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return first[:g].copy(), offsets, order.astype(np.int64), gid


def partition_to_groups(arr: pa.Array):
    w, valid = key_words(arr)
    n = len(w)
    starts = np.empty(max(n, 1), dtype=np.int64)
    lens = np.empty(max(n, 1), dtype=np.int64)
    g = lib().orc_partition_to_groups(_ptr(w), _ptr(valid), n, _ptr(starts), _ptr(lens))
    return starts[:g].copy(), lens[:g].copy()


def group_by_windows(time: np.ndarray, every: int, period: int, offset: int, closed: str):
    time = np.ascontiguousarray(time, dtype=np.int64)
    n = len(time)
    L = lib()
    g = L.orc_group_by_windows(_ptr(time), n, every, period, offset, _CLOSED[closed], None, None, None, None)
    starts, lens = np.empty(g, dtype=np.int64), np.empty(g, dtype=np.int64)
    lower, upper = np.empty(g, dtype=np.int64), np.empty(g, dtype=np.int64)
    if g:
        L.orc_group_by_windows(_ptr(time), n, every, period, offset, _CLOSED[closed],
                               _ptr(starts), _ptr(lens), _ptr(lower), _ptr(upper))
    return starts, lens, lower, upper


def index_unit_ns(typ: pa.DataType) -> int:
    if pa.types.is_timestamp(typ):
        return {"ns": 1, "us": 1_000, "ms": 1_000_000, "s": 1_000_000_000}[typ.unit]
    if pa.types.is_date32(typ):
        return 86_400_000_000_000
    return 1


def dynamic_durations(plan, typ: pa.DataType):
    d = plan.dynamic
    if getattr(plan, "_durations_in_ns", True) and not pa.types.is_integer(typ):
        u = index_unit_ns(typ)
        for x in (d.every, d.period, d.offset):
            if x % u:
                raise ValueError("duration is not a multiple of the index column's unit")
        return d.every // u, d.period // u, d.offset // u
    return d.every, d.period, d.offset


def group_by_dynamic(table: pa.Table, plan) -> pa.Table:
    """GroupByDynamicExec::execute_impl (polars-mem-engine/src/executors/group_by_dynamic.rs:19-85):
    stable sort by key -> per key slice group_by_windows -> slice aggregations."""
    d = plan.dynamic
    L = lib()
    if plan.predicates:
        table = filter_table(table, plan.predicates)
    n = table.num_rows
    idx_type = _combine(table.column(d.index_column)).type
    every, period, offset = dynamic_durations(plan, idx_type)
    if plan.keys:
        order = pc.sort_indices(table, sort_keys=[(k, "ascending") for k in plan.keys], null_placement="at_start")
        table = table.take(order)
        kstarts, klens = None, None
        # key slices over the sorted frame
        gid_first, offsets, _, gid = group_tuples(table, plan.keys)
        change = np.flatnonzero(np.diff(gid)) + 1 if n else np.array([], dtype=np.int64)
        kstarts = np.concatenate([[0], change]).astype(np.int64) if n else np.array([], dtype=np.int64)
        klens = np.diff(np.concatenate([kstarts, [n]])).astype(np.int64) if n else kstarts
    else:
        kstarts = np.array([0], dtype=np.int64) if n else np.array([], dtype=np.int64)
        klens = np.array([n], dtype=np.int64) if n else kstarts
    tarr = _physical(_combine(table.column(d.index_column)))
    if tarr.null_count:
        raise ValueError("null values in `group_by_dynamic` index column")
    time = tarr.to_numpy(zero_copy_only=False).astype(np.int64)
    S, Ln, Lo, Up = [], [], [], []
    for ks, kl in zip(kstarts, klens):
        sl = time[ks:ks + kl]
        if np.any(np.diff(sl) < 0):
            raise ValueError("argument in operation 'group_by_dynamic' is not sorted")
        s, l, lo, up = group_by_windows(sl, every, period, offset, d.closed)
        S.append(s + ks); Ln.append(l); Lo.append(lo); Up.append(up)
    cat = lambda xs: np.concatenate(xs).astype(np.int64) if xs else np.array([], dtype=np.int64)
    starts, lens, lower, upper = cat(S), cat(Ln), cat(Lo), cat(Up)
    g = len(starts)
    cols = {}
    for k, arr in zip(plan.keys, _take_keys(table, plan.keys, starts)):
        cols[k] = arr
    def _time(vals):
        a = pa.array(vals, type=pa.int64())
        if pa.types.is_date32(idx_type):
            return a.cast(pa.int32()).cast(idx_type)
        return a.cast(idx_type) if not pa.types.is_integer(idx_type) else a.cast(idx_type)
    if d.include_boundaries:
        cols["_lower_boundary"] = _time(lower)
        cols["_upper_boundary"] = _time(upper)
    label = {"left": lower, "right": upper, "datapoint": time[starts] if g else starts}[d.label]
    cols[d.index_column] = _time(label)
    c_aggs, keep, meta = _make_aggs(table, plan.aggs)
    for i, a in enumerate(plan.aggs):
        bits = np.empty(g, dtype=np.uint64)
        ok = np.empty(g, dtype=np.uint8)
        L.orc_agg_slices(C.byref(c_aggs[i]), _ptr(starts), _ptr(lens), g, _ptr(bits), _ptr(ok))
        cols[a.name] = _out_array(a.kind, bits, ok, meta[i][0], meta[i][1])
    return pa.table(cols)


def collect(lazy_result, n_threads: int = 1) -> pa.Table:
    """Run a polaroid_b200.plan.LazyResult on the oracle."""
    return group_by(lazy_result.table, lazy_result.plan, n_threads=n_threads)
