"""Loads tests/golden/vectors.json into pyarrow tables + logical plans (shared by the CPU-oracle
tests and the GPU parity tests)."""
import json
import math
import os

import numpy as np
import pyarrow as pa

import polaroid_b200 as pw
from polaroid_b200 import plan as P

HERE = os.path.dirname(os.path.abspath(__file__))


def load_cases(kinds=None):
    with open(os.path.join(HERE, "golden", "vectors.json")) as f:
        cases = json.load(f)
    return [c for c in cases if kinds is None or c["kind"] in kinds]


def _pa_type(t):
    simple = {"string": pa.string(), "date32": pa.date32()}
    if t in simple:
        return simple[t]
    if t.startswith("timestamp["):
        return pa.timestamp(t[10:-1])
    if t.startswith("duration["):
        return pa.duration(t[9:-1])
    return pa.type_for_alias(t)


def _fl(v):
    if v == "NaN":
        return math.nan
    if v == "inf":
        return math.inf
    if v == "-inf":
        return -math.inf
    return v


def column(spec):
    typ = _pa_type(spec["type"])
    if "repeat" in spec:
        vals = list(spec["repeat"][0]) * spec["repeat"][1]
    elif "repeat_each" in spec:
        vals = [v for v in spec["repeat_each"][0] for _ in range(spec["repeat_each"][1])]
    elif "squares" in spec:
        vals = [i * i for i in range(spec["squares"])]
    elif "range" in spec:
        vals = list(range(spec["range"]))
    elif "first_last_pattern" in spec:
        n = spec["first_last_pattern"]
        # py-polars/tests/unit/operations/test_group_by.py:2647-2659
        vals = (list(range(1, n + 1))
                + [None] * 1 + list(range(2, n - 0)) + [None] * 1
                + [None] * 2 + list(range(3, n - 1)) + [None] * 2
                + [None] * 3 + list(range(4, n - 2)) + [None] * 3
                + [None] * 4 + list(range(5, n - 3)) + [None] * 4)
    else:
        vals = [_fl(v) for v in spec["values"]]
    if pa.types.is_timestamp(typ) or pa.types.is_duration(typ):
        return pa.array(vals, type=pa.int64()).cast(typ)
    if pa.types.is_date32(typ):
        return pa.array(vals, type=pa.int32()).cast(typ)
    return pa.array(vals, type=typ)


def table(frame):
    return pa.table({k: column(v) for k, v in frame.items()})


def build_query(case, tbl, variant=None):
    """-> polaroid_b200.plan.LazyResult"""
    q = case["query"]
    lf = pw.LazyFrame(tbl)
    for c, op, v in q.get("filter", []):
        e = pw.col(c)
        lf = lf.filter({"lt": e < v, "le": e <= v, "gt": e > v, "ge": e >= v, "eq": e == v, "ne": e != v}[op])
    aggs = []
    for name, kind, colname, *extra in q["aggs"]:
        if kind == "len":
            aggs.append(pw.len().alias(name))
        elif kind in ("first_non_null", "last_non_null"):
            aggs.append(getattr(pw.col(colname), kind.split("_")[0])(ignore_nulls=True).alias(name))
        elif kind in ("var", "std"):
            aggs.append(getattr(pw.col(colname), kind)(*(extra[:1] or [1])).alias(name))   # optional 4th element: ddof
        else:
            aggs.append(getattr(pw.col(colname), kind)().alias(name))
    if "index" in q:
        label = (variant or {}).get("label", q.get("label", "left"))
        gb = lf.group_by_dynamic(q["index"], every=q["every"], period=q.get("period"), offset=q.get("offset"),
                                 closed=q.get("closed", "left"), label=label, group_by=q.get("group_by"),
                                 include_boundaries=q.get("include_boundaries", False))
    else:
        gb = lf.group_by(*q["keys"], maintain_order=q.get("maintain_order", False))
    return gb.agg(aggs)


def assert_tables_equal(got: pa.Table, want: pa.Table, sort_by=None, rtol=0.0, check_order=True):
    assert got.column_names == want.column_names, (got.column_names, want.column_names)
    assert got.num_rows == want.num_rows, (got.num_rows, want.num_rows)
    if sort_by:
        got = got.sort_by([(k, "ascending") for k in sort_by])
        want = want.sort_by([(k, "ascending") for k in sort_by])
    for name in want.column_names:
        g, w = got.column(name).combine_chunks(), want.column(name).combine_chunks()
        assert g.type == w.type, f"{name}: dtype {g.type} != {w.type}"
        gm = np.asarray(g.is_valid().to_numpy(zero_copy_only=False))
        wm = np.asarray(w.is_valid().to_numpy(zero_copy_only=False))
        assert (gm == wm).all(), f"{name}: validity differs\n got {g}\nwant {w}"
        if pa.types.is_floating(w.type):
            gv = g.fill_null(0).to_numpy(zero_copy_only=False).astype(np.float64)[wm]
            wv = w.fill_null(0).to_numpy(zero_copy_only=False).astype(np.float64)[wm]
            both_nan = np.isnan(gv) & np.isnan(wv)
            if rtol == 0.0:
                ok = both_nan | (gv == wv)
            else:
                with np.errstate(invalid="ignore"):
                    ok = both_nan | (gv == wv) | (np.abs(gv - wv) <= rtol * np.maximum(np.abs(gv), np.abs(wv)))
            assert ok.all(), f"{name}: values differ (rtol={rtol})\n got {gv[~ok][:5]}\nwant {wv[~ok][:5]}"
        else:
            assert g.equals(w), f"{name}: values differ\n got {g}\nwant {w}"


def expected_table(case, tbl_cols_order, variant=None):
    exp = {k: column(v) for k, v in case["expected"].items()}
    if variant is not None:
        idx = case["query"]["index"]
        src_type = column(case["frame"][idx]).type
        ts = variant["expected_ts"]
        if pa.types.is_timestamp(src_type):
            exp[idx] = pa.array(ts, type=pa.int64()).cast(src_type)
        elif pa.types.is_date32(src_type):
            exp[idx] = pa.array(ts, type=pa.int32()).cast(src_type)
        else:
            exp[idx] = pa.array(ts, type=src_type)
    return pa.table({k: exp[k] for k in tbl_cols_order})


def run_case(case, runner):
    """runner(LazyResult) -> pa.Table.  Asserts the case's expectations."""
    tbl = table(case["frame"])
    if case["kind"] == "group_by":
        got = runner(build_query(case, tbl))
        want = expected_table(case, got.column_names)
        assert_tables_equal(got, want, sort_by=case.get("sort_by"), rtol=case.get("rtol", 0.0))
    elif case["kind"] == "dynamic":
        for var in case["variants"]:
            got = runner(build_query(case, tbl, var))
            want = expected_table(case, got.column_names, var)
            assert_tables_equal(got, want, rtol=case.get("rtol", 0.0))
    elif case["kind"] == "dynamic_total":
        got = runner(build_query(case, tbl))
        total = sum(got.column(case["total_column"]).to_pylist())
        assert total == case["total"], (total, case["total"])
        if "n_rows" in case:
            assert got.num_rows == case["n_rows"], (got.num_rows, case["n_rows"])
    else:
        raise ValueError(case["kind"])
