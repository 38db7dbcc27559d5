"""CPU-side checks of the drop-in boundary (no compute): the C-ABI library loads and exports every symbol
include/polarway_b200.h declares, fails loudly without a GPU, the plugin shim's version / error / output-field
functions behave as the reference's loader expects, and the NVRTC specialisation of the scan kernel compiles."""
import ctypes as C
import os
import re

import pyarrow as pa
import pytest

import polaroid_b200 as pw
from polaroid_b200 import engine, plugin_loader
from tests import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "polarway_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pw_b200_\w+|_polars_plugin_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    L = engine.lib()
    syms = declared_symbols()
    assert len(syms) >= 24, syms
    for s in syms:
        assert hasattr(L, s), f"{s} is declared in include/polarway_b200.h but not exported"


def test_abi_version_and_plugin_version():
    L = engine.lib()
    assert L.pw_b200_abi_version() == 1
    assert plugin_loader._check_version(L) == (0, 1)   # polars-ffi/src/lib.rs:12-17


def test_no_gpu_means_a_loud_error_not_a_fallback():
    L = engine.lib()
    if L.pw_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    t = pa.table({"k": [1, 2, 1], "v": [1.0, 2.0, 3.0]})
    with pytest.raises(engine.PolarwayError) as e:
        pw.LazyFrame(t).group_by("k").agg(pw.col("v").sum()).collect()
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)


def test_plugin_field_function_infers_the_output_schema():
    t = synth.lineitem(10, seed=1)
    q = synth.q1_query(t, maintain_order=True)
    sub = t.select(["l_shipdate", "l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax"])
    fields = plugin_loader.plugin_field(sub.schema, plugin_loader.plan_to_kwargs(sub.schema, q.plan))
    assert fields == [("l_returnflag", "vu"), ("l_linestatus", "vu"), ("sum_qty", "l"), ("sum_base_price", "g"),
                      ("sum_disc_price", "g"), ("sum_charge", "g"), ("avg_qty", "g"), ("avg_price", "g"), ("avg_disc", "g"),
                      ("count_order", "I")]


def test_plugin_field_dtype_rules():
    # sum of i8/i16/u8/u16 -> Int64, i32 stays (sum.rs:40-47); mean of Date -> Datetime[us] (mean.rs:62-69); count -> u32
    t = pa.table({"k": pa.array([1], type=pa.int32()), "a": pa.array([1], type=pa.int8()), "b": pa.array([1], type=pa.int32()),
                  "d": pa.array([1], type=pa.int32()).cast(pa.date32())})
    q = pw.LazyFrame(t).group_by("k").agg(pw.col("a").sum().alias("sa"), pw.col("b").sum().alias("sb"), pw.col("d").mean().alias("md"),
                                          pw.col("a").count().alias("c"), pw.col("a").mean().alias("ma"))
    fields = plugin_loader.plugin_field(t.schema, plugin_loader.plan_to_kwargs(t.schema, q.plan))
    assert fields == [("k", "i"), ("sa", "l"), ("sb", "i"), ("md", "tsu:"), ("c", "I"), ("ma", "g")]


def test_plugin_bad_kwargs_sets_the_error_message():
    t = pa.table({"k": [1]})
    with pytest.raises(engine.PolarwayError) as e:
        plugin_loader.plugin_field(t.schema, {"keys": [0], "aggs": [("x", 0, 5, None)]})   # column 5 does not exist
    assert "the plugin failed with message" in str(e.value)


def test_jit_specialisation_compiles_without_a_gpu():
    L = engine.lib()
    buf = C.create_string_buffer(1 << 16)
    rc = L.pw_b200_jit_selftest(buf, C.c_size_t(len(buf)))
    if rc == 1:
        pytest.skip("libnvrtc is not installed on this machine")
    assert rc == 0, buf.value.decode()


def test_host_side_arithmetic_and_result_pool_without_a_gpu():
    # window index by multiply-shift == native 64-bit division (pw_plan.h: div_prepare / div_apply); large result
    # buffers fall back to pageable memory when nothing can be pinned
    L = engine.lib()
    L.pw_b200_host_selftest.restype = C.c_int64
    assert L.pw_b200_host_selftest() == 0


def test_view_columns_cross_the_boundary_with_their_data_buffers():
    """Host side of SURVEY 8-f1 (no device work): a string column is exported as Utf8View buffers
    [validity, views, data, variadic sizes] (polars-arrow/src/array/binview/view.rs:19-55) — values of up to 12 bytes inline,
    longer ones as (length, 4-byte prefix, buffer 0, offset) — and a view array of that layout (what the library returns
    when a result carries long keys) is read back value for value."""
    import numpy as np
    vals = ["", "a", "exactly12byt", "thirteen byte", "a string of more than twelve bytes", None, "x" * 300, "a string of more than twelve bytez"]
    arr = pa.array(vals, type=pa.large_string())
    vc = engine._ViewColumn(arr)
    lens = vc.views[:, 0:4].copy().view("<u4").reshape(-1)
    assert lens.tolist() == [0, 1, 12, 13, 34, 0, 300, 34]
    assert bytes(vc.views[2, 4:16]) == b"exactly12byt"                      # inline: the bytes live in the view
    long_rows = [3, 4, 6, 7]
    for i in long_rows:                                                     # long: prefix + (buffer 0, offset) into the data buffer
        off = int(vc.views[i, 12:16].copy().view("<u4")[0])
        assert int(vc.views[i, 8:12].copy().view("<u4")[0]) == 0
        assert bytes(vc.views[i, 4:8]) == vals[i].encode()[:4]
        assert bytes(vc.data[off:off + lens[i]]) == vals[i].encode()
    assert int(vc.sizes[0]) == len(vc.data)
    c_array, c_schema = engine.ArrowArray(), engine.ArrowSchema()
    vc.fill(c_array, c_schema)
    assert c_array.n_buffers == 4 and c_schema.format == b"vu"
    back = engine._import_view_array(c_array, c_schema)
    assert back.to_pylist() == vals
