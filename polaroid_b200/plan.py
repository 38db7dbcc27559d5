"""Logical plan for the one path this package accelerates:

    LazyFrame.filter(p).group_by(k).agg(a)        and
    LazyFrame.group_by_dynamic(index, every=...).agg(a)

The classes mirror the reference's front-end for that path (same method names and
argument meaning) so the parity tests read like the reference's own tests:

* ``LazyFrame.filter``            py-polars/src/polars/lazyframe/frame.py:4225
* ``LazyFrame.group_by``          frame.py:4678   (``maintain_order``)
* ``LazyFrame.group_by_dynamic``  frame.py:4926   (``every/period/offset/closed/label/group_by``)
* ``LazyGroupBy.agg``             py-polars/src/polars/lazyframe/group_by.py:76
* ``collect``                     frame.py:2385

Only the expression shapes this path needs are modelled: ``col(x) <cmp> scalar`` conjunctions
for the predicate, plain columns as keys, and ``sum/mean/min/max/count/len/first/last`` over a
column or over a product of affine factors of columns (TPC-H Q1's ``price*(1-disc)*(1+tax)``).
Anything else raises ``NotImplementedError`` — there is no CPU fallback.

This module is pure Python (no CUDA, no oracle imports): it only *describes* the query.
"""
from __future__ import annotations

import datetime as _dt
from dataclasses import dataclass, field
from typing import Any, Optional, Sequence

# ---- small enums shared with include/polarway_b200.h ------------------------------------
CMP_OPS = {"eq": 0, "ne": 1, "lt": 2, "le": 3, "gt": 4, "ge": 5}
AGG_KINDS = {"sum": 0, "mean": 1, "min": 2, "max": 3, "count": 4, "len": 5, "first": 6, "last": 7,
             "var": 8, "std": 9, "first_non_null": 10, "last_non_null": 11, "null_count": 12,
             "bitwise_and": 13, "bitwise_or": 14, "bitwise_xor": 15, "any": 16, "all": 17}
CLOSED = {"left": 0, "right": 1, "both": 2, "none": 3}
LABEL = {"left": 0, "right": 1, "datapoint": 2}


@dataclass(frozen=True)
class Factor:
    """One affine factor ``a + b * col``."""
    a: float
    b: float
    col: str


@dataclass(frozen=True)
class ValueExpr:
    """Either a plain column (``factors is None``; any dtype, value passed through exactly) or a
    product of affine factors evaluated in f64 left to right, each operation rounded once
    (no FMA contraction): ``((a0+b0*c0) * (a1+b1*c1)) * ...``.  Null if any input is null."""
    col: Optional[str] = None
    factors: Optional[tuple] = None

    def columns(self) -> list:
        if self.factors is None:
            return [self.col] if self.col is not None else []
        out = []
        for f in self.factors:
            if f.col not in out:
                out.append(f.col)
        return out


@dataclass(frozen=True)
class AggSpec:
    name: str
    kind: str            # key of AGG_KINDS
    expr: Optional[ValueExpr]  # None for len()
    ddof: int = 0        # var / std (py-polars Expr.var/std default: 1)


@dataclass(frozen=True)
class Predicate:
    col: str
    op: str              # key of CMP_OPS
    value: Any           # python scalar (int/float/date/datetime/timedelta)


@dataclass(frozen=True)
class DynamicOptions:
    """polars-time/src/group_by/dynamic.rs:19-39 (DynamicGroupOptions), fixed durations only."""
    index_column: str
    every: int           # in the index column's own unit (ns/us/ms, days for Date, 1 for ints)
    period: int
    offset: int
    closed: str = "left"
    label: str = "left"
    include_boundaries: bool = False


@dataclass
class GroupByPlan:
    predicates: list = field(default_factory=list)   # conjunction
    keys: list = field(default_factory=list)
    aggs: list = field(default_factory=list)
    maintain_order: bool = False
    dynamic: Optional[DynamicOptions] = None
    keys_sorted: bool = False    # every key column carries the sorted flag (LazyFrame.set_sorted): the GroupsSlice path


# ---- expression front-end -----------------------------------------------------------------
class Expr:
    """A restricted expression: product of affine factors of columns (or one plain column)."""

    def __init__(self, factors: Sequence[Factor], plain: Optional[str] = None, name: Optional[str] = None):
        self._factors = tuple(factors)
        self._plain = plain
        self._name = name

    # -- naming
    def alias(self, name: str) -> "Expr":
        e = Expr(self._factors, self._plain, name)
        return e

    def _out_name(self) -> str:
        if self._name:
            return self._name
        if self._plain:
            return self._plain
        return self._factors[0].col

    def _value_expr(self) -> ValueExpr:
        if self._plain is not None:
            return ValueExpr(col=self._plain)
        return ValueExpr(factors=self._factors)

    # -- arithmetic (closed under what a product of affine factors can express)
    def _single_affine(self) -> Optional[Factor]:
        if self._plain is not None:
            return Factor(0.0, 1.0, self._plain)
        if len(self._factors) == 1:
            return self._factors[0]
        return None

    def __mul__(self, other):
        if isinstance(other, Expr):
            lf = self._factors if self._plain is None else (Factor(0.0, 1.0, self._plain),)
            rf = other._factors if other._plain is None else (Factor(0.0, 1.0, other._plain),)
            return Expr(lf + rf, None, self._name)
        f = self._single_affine()
        if f is None or f.a != 0.0:
            raise NotImplementedError("only column*scalar or products of affine factors are supported")
        return Expr((Factor(0.0, f.b * float(other), f.col),), None, self._name)

    __rmul__ = __mul__

    def __add__(self, other):
        f = self._single_affine()
        if isinstance(other, Expr) or f is None:
            raise NotImplementedError("only scalar + column affine terms are supported")
        return Expr((Factor(f.a + float(other), f.b, f.col),), None, self._name)

    __radd__ = __add__

    def __sub__(self, other):
        return self.__add__(-float(other))

    def __rsub__(self, other):
        f = self._single_affine()
        if f is None:
            raise NotImplementedError("only scalar - column affine terms are supported")
        return Expr((Factor(float(other) - f.a, -f.b, f.col),), None, self._name)

    def __neg__(self):
        return self.__rsub__(0.0)

    # -- comparisons -> predicate
    def _cmp(self, op, value) -> "PredExpr":
        if self._plain is None:
            raise NotImplementedError("predicates compare a plain column with a scalar")
        return PredExpr([Predicate(self._plain, op, value)])

    def __lt__(self, v): return self._cmp("lt", v)
    def __le__(self, v): return self._cmp("le", v)
    def __gt__(self, v): return self._cmp("gt", v)
    def __ge__(self, v): return self._cmp("ge", v)
    def __eq__(self, v): return self._cmp("eq", v)      # type: ignore[override]
    def __ne__(self, v): return self._cmp("ne", v)      # type: ignore[override]
    __hash__ = None  # type: ignore[assignment]

    # -- aggregations
    def _agg(self, kind) -> "AggExpr":
        return AggExpr(kind, self._value_expr(), self._out_name())

    def sum(self): return self._agg("sum")
    def mean(self): return self._agg("mean")
    def min(self): return self._agg("min")
    def max(self): return self._agg("max")
    def count(self): return self._agg("count")
    def len(self): return self._agg("len")
    # py-polars Expr.first/last(ignore_nulls=...): expr.py (IRAggExpr::FirstNonNull / LastNonNull)
    def first(self, *, ignore_nulls: bool = False): return self._agg("first_non_null" if ignore_nulls else "first")
    def last(self, *, ignore_nulls: bool = False): return self._agg("last_non_null" if ignore_nulls else "last")
    # the remaining pre-aggregatable reductions (polars-expr/src/reduce/convert.rs:46-150)
    def var(self, ddof: int = 1): return AggExpr("var", self._value_expr(), self._out_name(), ddof)
    def std(self, ddof: int = 1): return AggExpr("std", self._value_expr(), self._out_name(), ddof)
    def null_count(self): return self._agg("null_count")
    def bitwise_and(self): return self._agg("bitwise_and")
    def bitwise_or(self): return self._agg("bitwise_or")
    def bitwise_xor(self): return self._agg("bitwise_xor")

    def any(self, *, ignore_nulls: bool = True):
        if not ignore_nulls:
            raise NotImplementedError("any(ignore_nulls=False) (Kleene logic) is outside this path")
        return self._agg("any")

    def all(self, *, ignore_nulls: bool = True):
        if not ignore_nulls:
            raise NotImplementedError("all(ignore_nulls=False) (Kleene logic) is outside this path")
        return self._agg("all")


class AggExpr:
    def __init__(self, kind: str, expr: Optional[ValueExpr], name: str, ddof: int = 0):
        self.kind, self.expr, self.name, self.ddof = kind, expr, name, ddof

    def alias(self, name: str) -> "AggExpr":
        return AggExpr(self.kind, self.expr, name, self.ddof)

    def spec(self) -> AggSpec:
        return AggSpec(self.name, self.kind, self.expr, self.ddof)


class PredExpr:
    def __init__(self, preds: list):
        self.preds = list(preds)

    def __and__(self, other: "PredExpr") -> "PredExpr":
        return PredExpr(self.preds + other.preds)


def col(name: str) -> Expr:
    return Expr((), plain=name)


def len_() -> AggExpr:  # pl.len()
    return AggExpr("len", None, "len")


def sum_(name: str) -> AggExpr:  # pl.sum("b")
    return col(name).sum()


def count_(name: str) -> AggExpr:  # pl.count("a")
    return col(name).count()


# ---- durations ------------------------------------------------------------------------------
_NS = {"ns": 1, "us": 1_000, "ms": 1_000_000, "s": 1_000_000_000, "m": 60_000_000_000,
       "h": 3_600_000_000_000, "d": 86_400_000_000_000, "w": 604_800_000_000_000}


def parse_duration_ns(s) -> tuple:
    """Parse a Polars duration string made of fixed-length units (polars-time/src/windows/
    duration.rs `Duration::parse`): returns (nanoseconds, is_index_unit).  ``"3i"`` is the integer
    index unit.  Calendar units (mo, q, y) are out of scope."""
    if isinstance(s, _dt.timedelta):
        return (s.days * 86400 + s.seconds) * 1_000_000_000 + s.microseconds * 1000, False
    if isinstance(s, int):
        return s, True
    txt = s.strip()
    neg = txt.startswith("-")
    if neg or txt.startswith("+"):
        txt = txt[1:]
    total, num, i, index_unit = 0, "", 0, False
    while i < len(txt):
        c = txt[i]
        if c.isdigit():
            num += c
            i += 1
            continue
        j = i
        while j < len(txt) and not txt[j].isdigit():
            j += 1
        unit = txt[i:j]
        if unit == "i":
            index_unit = True
            total += int(num)
        elif unit in _NS:
            total += int(num) * _NS[unit]
        else:
            raise NotImplementedError(f"duration unit {unit!r} (calendar durations are out of scope)")
        num = ""
        i = j
    if num:
        raise ValueError(f"duration {s!r} is missing a unit")
    return (-total if neg else total), index_unit


# ---- LazyFrame mirror ----------------------------------------------------------------------
class LazyFrame:
    """Holds a pyarrow Table plus the (at most one) pending filter."""

    def __init__(self, data):
        import pyarrow as pa
        if isinstance(data, dict):
            data = pa.table(data)
        self._table = data
        self._preds: list = []
        self._sorted: set = set()

    def filter(self, pred: PredExpr) -> "LazyFrame":
        out = LazyFrame(self._table)
        out._preds = self._preds + list(pred.preds)
        out._sorted = set(self._sorted)
        return out

    def set_sorted(self, column: str) -> "LazyFrame":
        """Flag a column as sorted (py-polars LazyFrame.set_sorted): a group_by over flagged key columns takes the
        run-boundary path (polars-core/src/frame/group_by/into_groups.rs:65-129).  The flag is a promise, not a check."""
        out = LazyFrame(self._table)
        out._preds = list(self._preds)
        out._sorted = set(self._sorted) | {column}
        return out

    def group_by(self, *keys, maintain_order: bool = False) -> "LazyGroupBy":
        flat = []
        for k in keys:
            flat.extend(k if isinstance(k, (list, tuple)) else [k])
        names = [k if isinstance(k, str) else k._out_name() for k in flat]
        return LazyGroupBy(self, GroupByPlan(predicates=list(self._preds), keys=names, maintain_order=maintain_order,
                                             keys_sorted=bool(names) and all(k in self._sorted for k in names)))

    def group_by_dynamic(self, index_column: str, *, every, period=None, offset=None,
                         closed: str = "left", label: str = "left", group_by=None,
                         include_boundaries: bool = False) -> "LazyGroupBy":
        # period defaults to every, offset to 0: frame.py:5240-5262
        e, e_idx = parse_duration_ns(every)
        p, _ = parse_duration_ns(period) if period is not None else (e, e_idx)
        o, _ = parse_duration_ns(offset) if offset is not None else (0, e_idx)
        if e <= 0:
            raise ValueError("'every' argument must be positive")  # dynamic.rs:271
        keys = [] if group_by is None else ([group_by] if isinstance(group_by, str) else list(group_by))
        dyn = DynamicOptions(index_column, e, p, o, closed, label, include_boundaries)
        plan = GroupByPlan(predicates=list(self._preds), keys=keys, maintain_order=True, dynamic=dyn)
        plan._durations_in_ns = not e_idx  # converted to the column's unit at collect time
        return LazyGroupBy(self, plan)

    def collect(self, engine=None):
        """A bare ``filter`` (FilterExec, polars-mem-engine/src/executors/filter.rs)."""
        from . import engine as _engine
        return _engine.run_filter(self._table, self._preds)


class LazyGroupBy:
    def __init__(self, lf: LazyFrame, plan: GroupByPlan):
        self._lf, self._plan = lf, plan

    def agg(self, *aggs) -> "LazyResult":
        flat = []
        for a in aggs:
            flat.extend(a if isinstance(a, (list, tuple)) else [a])
        plan = self._plan
        plan.aggs = [a.spec() for a in flat]
        return LazyResult(self._lf._table, plan)

    # convenience forms used by the reference tests
    def _all(self, kind):
        skip = set(self._plan.keys)
        if self._plan.dynamic:
            skip.add(self._plan.dynamic.index_column)
        cols = [c for c in self._lf._table.column_names if c not in skip]
        return self.agg([AggExpr(kind, ValueExpr(col=c), c) for c in cols])

    def sum(self): return self._all("sum")
    def mean(self): return self._all("mean")
    def min(self): return self._all("min")
    def max(self): return self._all("max")
    def first(self): return self._all("first")
    def last(self): return self._all("last")
    def len(self): return self.agg(len_())


class LazyResult:
    def __init__(self, table, plan: GroupByPlan):
        self.table, self.plan = table, plan

    def collect(self, engine=None):
        """Runs on the B200 through the C ABI.  There is no CPU engine in this package."""
        from . import engine as _engine
        return _engine.run_group_by(self.table, self.plan)
