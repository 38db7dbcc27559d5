"""Seeded synthetic inputs with the shapes BASELINE.json / SURVEY §8d name (C1..C4)."""
import datetime as dt
import json
import os

import numpy as np
import pyarrow as pa

import polaroid_b200 as pw

HERE = os.path.dirname(os.path.abspath(__file__))


def c2_table(n, groups, seed=2, int_value=False):
    rng = np.random.default_rng(seed)
    keys = rng.integers(0, groups, n, dtype=np.int64) * 7919 - 3  # not a dense 0..G range
    if n >= groups:
        keys[:groups] = np.arange(groups, dtype=np.int64) * 7919 - 3   # every group present
    vals = rng.integers(0, 100, n, dtype=np.int64) if int_value else rng.random(n) * 100.0
    return pa.table({"key": keys, "value": vals})


def c3_table(n, groups, seed=3, null_ratio=0.05):
    rng = np.random.default_rng(seed)
    keys = rng.integers(0, groups, n, dtype=np.int64)
    kmask = np.zeros(n, dtype=bool)
    kmask[rng.integers(0, n, max(1, n // 1000))] = True      # one null-key group
    vals = rng.random(n) * 100.0
    return pa.table({"key": pa.array(keys, mask=kmask), "value": pa.array(vals, mask=rng.random(n) < null_ratio)})


def lineitem(n, seed=1):
    """TPC-H lineitem columns Q1 touches, dtypes as in the reference's fixture
    (examples/datasets/pds_heads/lineitem.feather): shipdate timestamp[us], flags strings, qty i64, rest f64."""
    rng = np.random.default_rng(seed)
    u = rng.random(n)
    flag = np.where(u < 0.25, "A", np.where(u < 0.50, "R", "N"))
    status = np.where(flag == "N", np.where(rng.random(n) < 0.014, "F", "O"), "F")
    start = np.datetime64("1992-01-02", "us").astype(np.int64)
    span = (np.datetime64("1998-12-01", "us") - np.datetime64("1992-01-02", "us")).astype(np.int64)
    ship = start + (rng.random(n) * span).astype(np.int64) // 86_400_000_000 * 86_400_000_000
    return pa.table({
        "l_shipdate": pa.array(ship, type=pa.int64()).cast(pa.timestamp("us")),
        "l_returnflag": pa.array(flag), "l_linestatus": pa.array(status),
        "l_quantity": pa.array(rng.integers(1, 51, n, dtype=np.int64)),
        "l_extendedprice": pa.array(np.round(900 + rng.random(n) * 104_100, 2)),
        "l_discount": pa.array(rng.integers(0, 11, n) / 100.0),
        "l_tax": pa.array(rng.integers(0, 9, n) / 100.0)})


def q1_query(t, maintain_order=False):
    c = pw.col
    disc_price = c("l_extendedprice") * (1 - c("l_discount"))
    charge = disc_price * (1 + c("l_tax"))
    return (pw.LazyFrame(t).filter(c("l_shipdate") <= dt.datetime(1998, 9, 2))
            .group_by("l_returnflag", "l_linestatus", maintain_order=maintain_order)
            .agg(c("l_quantity").sum().alias("sum_qty"), c("l_extendedprice").sum().alias("sum_base_price"),
                 disc_price.sum().alias("sum_disc_price"), charge.sum().alias("sum_charge"),
                 c("l_quantity").mean().alias("avg_qty"), c("l_extendedprice").mean().alias("avg_price"),
                 c("l_discount").mean().alias("avg_disc"), pw.len().alias("count_order")))


def lineitem_head():
    with open(os.path.join(HERE, "golden", "lineitem_head.json")) as f:
        d = json.load(f)
    return pa.table({
        "l_shipdate": pa.array(d["l_shipdate_us"], type=pa.int64()).cast(pa.timestamp("us")),
        "l_returnflag": pa.array(d["l_returnflag"]), "l_linestatus": pa.array(d["l_linestatus"]),
        "l_quantity": pa.array(d["l_quantity"], type=pa.int64()),
        "l_extendedprice": pa.array(d["l_extendedprice"], type=pa.float64()),
        "l_discount": pa.array(d["l_discount"], type=pa.float64()), "l_tax": pa.array(d["l_tax"], type=pa.float64())})


def ohlcv(n, n_symbols=100, seed=4, mean_gap_us=1000):
    """C4: ticks sorted by timestamp, symbols interleaved."""
    rng = np.random.default_rng(seed)
    gaps = rng.exponential(mean_gap_us, n).astype(np.int64)
    ts = np.cumsum(gaps) + np.datetime64("2024-01-02T09:30:00", "us").astype(np.int64)
    price = 100.0 + np.cumsum(rng.normal(0, 0.01, n))
    return pa.table({"ts": pa.array(ts, type=pa.int64()).cast(pa.timestamp("us")),
                     "symbol": pa.array(rng.integers(0, n_symbols, n).astype(np.uint32)),
                     "price": pa.array(price), "volume": pa.array(rng.integers(1, 1001, n, dtype=np.int64))})


def ohlcv_query(t, by_symbol=True, every="1m"):
    c = pw.col
    return (pw.LazyFrame(t).group_by_dynamic("ts", every=every, group_by="symbol" if by_symbol else None)
            .agg(c("price").first().alias("open"), c("price").max().alias("high"), c("price").min().alias("low"),
                 c("price").last().alias("close"), c("volume").sum().alias("volume")))
