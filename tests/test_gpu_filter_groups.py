"""FilterExec alone (predicate -> selection vector -> column compaction) and GroupsIdx construction vs the oracle."""
import numpy as np
import pyarrow as pa
import pytest

import polaroid_b200 as pw
from oracle import oracle
from polaroid_b200 import engine
from tests import golden_util as G

pytestmark = pytest.mark.gpu


def _frame(n, seed=3):
    rng = np.random.default_rng(seed)
    words = np.array(["", "a", "bb", "ccc", "twelve_bytes"])
    return pa.table({"i": pa.array(rng.integers(-100, 100, n), mask=rng.random(n) < 0.1),
                     "f": pa.array(rng.normal(size=n), mask=rng.random(n) < 0.1),
                     "s": pa.array(words[rng.integers(0, len(words), n)], mask=rng.random(n) < 0.1),
                     "u8": pa.array(rng.integers(0, 255, n).astype(np.uint8)),
                     "d": pa.array(rng.integers(0, 20000, n).astype(np.int32)).cast(pa.date32())})


@pytest.mark.parametrize("n", [0, 1, 31, 2048, 2049, 100_003])
def test_filter_compacts_every_column(n):
    t = _frame(n)
    lf = pw.LazyFrame(t).filter((pw.col("i") > -20) & (pw.col("f") <= 0.5))
    got = lf.collect()
    want = oracle.filter_table(t, lf._preds)
    G.assert_tables_equal(got, want)


def test_selection_vector_is_ascending_and_null_is_false():
    t = _frame(50_000).slice(7)   # offset != 0: validity bit offset honoured
    preds = (pw.col("i") >= 0).preds
    ids = engine.filter_select(t, preds).to_numpy()
    mask = oracle.predicate_mask(t, preds).astype(bool)
    assert (ids == np.flatnonzero(mask)).all()
    assert ids.dtype == np.uint32


@pytest.mark.parametrize("maintain_order", [True, False])
def test_group_tuples_match_oracle(maintain_order):
    t = _frame(60_000, seed=4)
    first, offsets, row_ids = engine.group_tuples(t, ["s", "u8"], maintain_order=maintain_order)
    ofirst, ooff, oids, _ = oracle.group_tuples(t, ["s", "u8"])
    first, offsets, row_ids = first.to_numpy(), offsets.to_numpy().astype(np.int64), row_ids.to_numpy()
    assert len(first) == len(ofirst) and len(row_ids) == t.num_rows
    got = {int(first[g]): row_ids[offsets[g]:offsets[g + 1]].tolist() for g in range(len(first))}
    want = {int(ofirst[g]): oids[ooff[g]:ooff[g + 1]].tolist() for g in range(len(ofirst))}
    assert got == want                       # same groups, same ascending row ids, first == smallest id
    if maintain_order:
        assert first.tolist() == ofirst.tolist()   # groups in first-occurrence order


def test_sorted_key_fast_path_equivalence():
    # into_groups.rs:65-129: on a sorted key the groups are the runs
    keys = np.repeat(np.arange(500, dtype=np.int64), 37)
    t = pa.table({"k": pa.array(keys, mask=keys == 3)})
    first, offsets, row_ids = engine.group_tuples(t, ["k"], maintain_order=True)
    starts, lens = oracle.partition_to_groups(t["k"])
    assert first.to_numpy().tolist() == starts.tolist()
    assert np.diff(offsets.to_numpy().astype(np.int64)).tolist() == lens.tolist()
    assert row_ids.to_numpy().tolist() == list(range(len(keys)))


def test_filter_carries_strings_longer_than_twelve_bytes():
    # the selected views of long values point into the frame's data buffers; the result gathers their
    # bytes into a data buffer of its own (binview/view.rs:19-55)
    rng = np.random.default_rng(6)
    n = 30_000
    names = ["short", "exactly12byt", "a value of more than twelve bytes", "a value of more than twelve bytez", "y" * 200, ""]
    t = pa.table({"i": pa.array(rng.integers(-100, 100, n)),
                  "s": pa.array([names[k] for k in rng.integers(0, len(names), n)], mask=rng.random(n) < 0.05),
                  "f": pa.array(rng.random(n))})
    lf = pw.LazyFrame(t).filter((pw.col("i") > 10) & (pw.col("f") <= 0.7))
    got = lf.collect()
    want = oracle.filter_table(t, lf._preds)
    G.assert_tables_equal(got, want)
    # group tuples over the same column: equal long strings are one group
    first, offsets, row_ids = engine.group_tuples(t, ["s"], maintain_order=True)
    ofirst, ooff, oids, _ = oracle.group_tuples(t, ["s"])
    assert first.to_numpy().tolist() == ofirst.tolist() and offsets.to_numpy().tolist() == list(ooff)
