"""key types beyond plain integers and short strings.
  * Categorical / Enum: group identity is the physical code (polars-expr/src/hash_keys.rs:32,83-89); the codes cross the
    boundary as u8/u16/u32/i32 columns and the categories are reattached on the way out;
  * Boolean keys (row-encoded by the reference, hash_keys.rs:37,114-141);
  * strings longer than 12 bytes group by their bytes (views canonicalised when the frame is created, pw_views.cu).
CUDA vs the oracle, bit-exact."""
import numpy as np
import pyarrow as pa
import pytest

import polaroid_b200 as pw
from oracle import oracle
from polaroid_b200 import engine
from tests import golden_util as G

pytestmark = pytest.mark.gpu

STRATEGIES = [{}, {"flags": engine.FLAG_FORCE_HOT}, {"flags": engine.FLAG_FORCE_GLOBAL}]


def check(q, sort_by=None, **opts):
    got = engine.run_group_by(q.table, q.plan, **opts)
    want = oracle.collect(q)
    if sort_by is None:
        G.assert_tables_equal(got, want)
    else:   # dictionary columns do not sort in pyarrow: compare on the decoded strings
        dec = lambda t: pa.table({n: (t.column(n).cast(t.schema.field(n).type.value_type) if pa.types.is_dictionary(t.schema.field(n).type) else t.column(n))
                                  for n in t.column_names})
        assert got.schema == want.schema
        G.assert_tables_equal(dec(got), dec(want), sort_by=sort_by)
    return got


@pytest.mark.parametrize("strategy", STRATEGIES)
@pytest.mark.parametrize("index_type", ["int8", "uint16", "uint32", "int32"])
def test_categorical_keys_group_on_physical_codes(index_type, strategy):
    rng = np.random.default_rng(61)
    n, cats = 200_000, 100
    values = pa.array([f"category-{i:04d}-with-a-long-name" for i in range(cats)])   # the names never reach the device
    codes = pa.array(rng.integers(0, cats, n).astype(index_type), mask=rng.random(n) < 0.02)
    key = pa.DictionaryArray.from_arrays(codes, values)
    t = pa.table({"k": key, "v": pa.array(rng.integers(-100, 100, n))})
    q = pw.LazyFrame(t).group_by("k").agg(pw.col("v").sum().alias("s"), pw.col("v").min().alias("lo"), pw.len().alias("n"))
    got = check(q, sort_by=["k"], **strategy)
    assert pa.types.is_dictionary(got.schema.field("k").type) and got.num_rows == cats + 1


def test_categorical_and_integer_key_maintain_order():
    rng = np.random.default_rng(62)
    n = 50_000
    key = pa.DictionaryArray.from_arrays(pa.array(rng.integers(0, 7, n).astype(np.uint32)), pa.array(list("abcdefg")))
    t = pa.table({"c": key, "i": pa.array(rng.integers(0, 5, n)), "v": pa.array(rng.random(n))})
    q = pw.LazyFrame(t).group_by("c", "i", maintain_order=True).agg(pw.col("v").max().alias("hi"), pw.len().alias("n"))
    check(q)


@pytest.mark.parametrize("strategy", STRATEGIES)
def test_boolean_key_alone_and_with_nulls(strategy):
    rng = np.random.default_rng(63)
    n = 120_000
    t = pa.table({"b": pa.array(rng.random(n) < 0.3, mask=rng.random(n) < 0.1), "v": pa.array(rng.integers(0, 1000, n)),
                  "f": pa.array(rng.random(n))})
    q = pw.LazyFrame(t).group_by("b").agg(pw.col("v").sum().alias("s"), pw.col("f").max().alias("hi"), pw.col("v").first().alias("first"),
                                          pw.len().alias("n"))
    got = check(q, sort_by=["b"], **strategy)
    assert got.num_rows == 3 and got.schema.field("b").type == pa.bool_()


def test_boolean_key_next_to_a_string_key_sliced_input():
    rng = np.random.default_rng(64)
    n = 80_000
    names = np.array(["alpha", "beta", "gamma", "delta"])
    t = pa.table({"b": pa.array(rng.random(n) < 0.5, mask=rng.random(n) < 0.05), "s": pa.array(names[rng.integers(0, 4, n)]),
                  "v": pa.array(rng.integers(0, 9, n))}).slice(13, n - 40)   # bit offset 13 in the Boolean bitmaps
    q = pw.LazyFrame(t).group_by("b", "s", maintain_order=True).agg(pw.col("v").sum().alias("s_v"), pw.len().alias("n"))
    check(q)


LONG = ["a string of more than twelve bytes", "a string of more than twelve bytes!", "a string of more than twelve bytez",
        "short", "exactly12byt", "thirteen byte", "", "x" * 300, "x" * 299 + "y", "a string of more than twelve bytes"[:-1] + "S"]


@pytest.mark.parametrize("flags", [0, engine.FLAG_FORCE_GLOBAL, engine.FLAG_FORCE_HOT])
def test_long_string_keys_group_by_their_bytes(flags):
    # views of equal long strings point at different offsets of the data buffer: grouping must compare the bytes
    # (binview/view.rs:19-55).  Same prefixes, same lengths, one-byte differences at the end, 300-byte values, nulls.
    rng = np.random.default_rng(71)
    n = 5_000
    pick = rng.integers(0, len(LONG), n)
    keys = pa.array([LONG[i] for i in pick], mask=rng.random(n) < 0.03)
    t = pa.table({"k": keys, "v": pa.array(rng.integers(-100, 100, n)), "w": pa.array(rng.random(n))})
    q = pw.LazyFrame(t).group_by("k").agg(pw.col("v").sum().alias("s"), pw.len().alias("n"), pw.col("w").max().alias("m"),
                                         pw.col("v").first().alias("f"), pw.col("v").last().alias("l"))
    got = engine.run_group_by(q.table, q.plan, flags=flags)
    assert got.num_rows == len(LONG) + 1
    G.assert_tables_equal(got, oracle.collect(q), sort_by=["k"], rtol=1e-12)


def test_long_string_key_next_to_an_integer_key_keeps_first_occurrence_order():
    rng = np.random.default_rng(72)
    n = 40_000
    names = [f"instrument/{i:05d}/with-a-long-symbol-name" for i in range(700)]
    t = pa.table({"name": pa.array([names[i] for i in rng.integers(0, len(names), n)]), "venue": pa.array(rng.integers(0, 3, n).astype("int8")),
                  "px": pa.array(rng.random(n) * 100)})
    q = pw.LazyFrame(t).group_by("venue", "name", maintain_order=True).agg(pw.col("px").mean().alias("avg"), pw.col("px").min().alias("lo"), pw.len().alias("n"))
    got = engine.run_group_by(q.table, q.plan)
    G.assert_tables_equal(got, oracle.collect(q), rtol=1e-12)


def test_many_distinct_long_strings_take_the_high_cardinality_path():
    rng = np.random.default_rng(73)
    n, g = 300_000, 120_000
    ids = rng.integers(0, g, n)
    t = pa.table({"k": pa.array([f"user-{i:09d}@example.org" for i in ids]), "v": pa.array(rng.integers(0, 1000, n))})
    q = pw.LazyFrame(t).group_by("k").agg(pw.col("v").sum().alias("s"), pw.len().alias("n"))
    got = engine.run_group_by(q.table, q.plan)
    G.assert_tables_equal(got, oracle.collect(q), sort_by=["k"])


def test_resident_frame_with_long_keys_serves_several_queries():
    t = pa.table({"k": pa.array([LONG[i % len(LONG)] for i in range(1000)]), "v": pa.array(np.arange(1000))})
    frame = engine.DeviceFrame(t)
    for agg in (pw.col("v").sum().alias("s"), pw.col("v").max().alias("m")):
        q = pw.LazyFrame(t).group_by("k").agg(agg)
        got = frame.group_by(q.plan)
        G.assert_tables_equal(got, oracle.collect(q), sort_by=["k"])
    frame.free()


def test_group_by_dynamic_refuses_long_string_keys():
    t = pa.table({"t": pa.array(np.arange(100, dtype=np.int64)), "k": pa.array(["a key of more than twelve bytes"] * 100), "v": pa.array(np.arange(100))})
    q = pw.LazyFrame(t).group_by_dynamic("t", every="10i", group_by="k").agg(pw.col("v").sum().alias("s"))
    with pytest.raises(engine.PolarwayError) as e:
        engine.run_group_by(q.table, q.plan)
    assert e.value.code == -2   # PW_ERR_UNSUPPORTED
