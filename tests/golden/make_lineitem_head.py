"""Reads the reference's own fixture examples/datasets/pds_heads/lineitem.feather (10 rows) in THIS container
and commits the columns TPC-H Q1 touches as JSON (the reference tree is not present on the GPU box)."""
import json
import os

import pyarrow as pa
import pyarrow.feather as feather

t = feather.read_table("/root/reference/examples/datasets/pds_heads/lineitem.feather")
print(t.schema)
out = {
    "source": "examples/datasets/pds_heads/lineitem.feather",
    "l_shipdate_us": t["l_shipdate"].cast(pa.timestamp("us")).cast(pa.int64()).to_pylist(),
    "l_returnflag": t["l_returnflag"].to_pylist(), "l_linestatus": t["l_linestatus"].to_pylist(),
    "l_quantity": t["l_quantity"].to_pylist(), "l_extendedprice": t["l_extendedprice"].to_pylist(),
    "l_discount": t["l_discount"].to_pylist(), "l_tax": t["l_tax"].to_pylist(),
}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lineitem_head.json"), "w") as f:
    json.dump(out, f)
