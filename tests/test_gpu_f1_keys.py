"""key types beyond plain integers and short strings.
  * Categorical / Enum: group identity is the physical code (polars-expr/src/hash_keys.rs:32,83-89); the codes cross the
    boundary as u8/u16/u32/i32 columns and the categories are reattached on the way out;
  * Boolean keys (row-encoded by the reference, hash_keys.rs:37,114-141);
  * strings longer than 12 bytes are refused loudly, never mis-grouped.
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


def test_long_string_keys_are_refused_not_misgrouped():
    t = pa.table({"k": pa.array(["short", "a string of more than twelve bytes", "a string of more than twelve bytes!", "short"]).cast(pa.string_view()),
                  "v": pa.array([1, 2, 3, 4])})
    q = pw.LazyFrame(t).group_by("k").agg(pw.col("v").sum().alias("s"))
    with pytest.raises(engine.PolarwayError) as e:
        engine.run_group_by(q.table, q.plan)
    assert e.value.code == -2   # PW_ERR_UNSUPPORTED
