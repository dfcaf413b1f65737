"""Aggregation constructors — the reference's public API surface, host side.

Each constructor returns an `Agg` node.  Where the reference monomorphises the typed tree
(generics + tuples, src/tuple.rs, src/agg.rs:10-36) and runs `SegmentAgg::collect` per
document, this facade *lowers* the tree once per query into the flat pre-order `tagg_node`
array of include/tagg.h (`Agg.lower`) and *decodes* the device result back into the same
fruit shapes (`Agg.decode`).  Tuples of aggs are plain Python tuples (arity 2..=10 as in
src/tuple.rs:73-81).
"""
import numpy as np

from . import _ffi as F
from . import codec
from .fruits import Histogram, Percentiles, Terms

MAX_TUPLE = 10
LUT_MAX_DOMAIN = 1 << 22  # closures are tabulated over the code domain up to this many codes


class LowerCtx:
    """Accumulates the flat plan while walking the typed tree."""

    def __init__(self, searcher=None):
        self.nodes = []       # list[_ffi.Node]
        self.blobs = []       # list[bytes]
        self.filters = []     # list[Query] — one per FILTER node, aux = index
        self.searcher = searcher

    def emit(self, **kw):
        n = F.Node()
        for k, v in kw.items():
            setattr(n, k, v)
        self.nodes.append(n)
        return len(self.nodes) - 1


class Agg:
    """Mirror of `trait Agg` (src/agg.rs:10-17); `requires_scoring` is false for every built-in."""

    def requires_scoring(self):
        return False

    def lower(self, ctx):  # -> index of the emitted node
        raise NotImplementedError

    def decode(self, reader, bucket):  # -> fruit of `bucket` (index into the enclosing scope)
        raise NotImplementedError

    def create_fruit(self):
        """`PreparedAgg::create_fruit` (src/agg.rs:23)."""
        raise NotImplementedError

    def merge(self, acc, fruit):
        """`PreparedAgg::merge(&mut acc, fruit)` (src/agg.rs:27) on decoded fruits; returns the new acc.
        Used when fruits of different processes are combined on the host (e.g. hashed bucket tables)."""
        raise NotImplementedError


def as_agg(a):
    if isinstance(a, Agg):
        return a
    if isinstance(a, tuple):
        return TupleAgg(a)
    raise TypeError(f"not an aggregation: {a!r}")


class TupleAgg(Agg):
    """(a1, .., an) — src/tuple.rs:5-81"""

    def __init__(self, members):
        if not 2 <= len(members) <= MAX_TUPLE:
            raise TypeError(f"tuple aggregations have arity 2..={MAX_TUPLE} (src/tuple.rs:73-81)")
        self.members = [as_agg(m) for m in members]

    def lower(self, ctx):
        i = ctx.emit(op=F.OP_TUPLE, n_children=len(self.members))
        for m in self.members:
            m.lower(ctx)
        return i

    def decode(self, reader, bucket):
        return tuple(m.decode(reader, bucket) for m in self.members)

    def create_fruit(self):
        return tuple(m.create_fruit() for m in self.members)

    def merge(self, acc, fruit):
        return tuple(m.merge(a, f) for m, a, f in zip(self.members, acc, fruit))


class CountAgg(Agg):
    """count_agg() — src/metric/count.rs:7-9; Fruit = u64"""

    def lower(self, ctx):
        self.node = ctx.emit(op=F.OP_COUNT)
        return self.node

    def decode(self, reader, bucket):
        values, _ = reader.metric(self.node)
        return int(values[bucket])

    def create_fruit(self):
        return 0

    def merge(self, acc, fruit):  # count.rs:39-41
        return acc + fruit


class _FoldAgg(Agg):
    """sum / min / max — Fruit = Option<T> (src/metric/sum.rs, src/metric/minmax.rs)"""
    op = None

    def __init__(self, field, kind, multi):
        self.field, self.kind, self.multi = int(field), kind, int(multi)

    def lower(self, ctx):
        self.node = ctx.emit(op=self.op, kind=self.kind, multi=self.multi, field_id=self.field)
        return self.node

    def decode(self, reader, bucket):
        values, seen = reader.metric(self.node)
        if not seen[bucket]:
            return None
        return codec.bits_to_value(self.kind, values[bucket])

    def create_fruit(self):
        return None

    def merge(self, acc, fruit):  # sum.rs:59-70, minmax.rs:59-72
        if fruit is None:
            return acc
        if acc is None:
            return fruit
        if self.op == F.OP_SUM:
            if self.kind == F.F64:
                return acc + fruit
            r = (acc + fruit) & ((1 << 64) - 1)  # wrapping, like release-mode Rust
            return r - (1 << 64) if self.kind == F.I64 and r >> 63 else r
        if self.op == F.OP_MIN:
            return fruit if fruit < acc else acc
        return fruit if fruit > acc else acc


class SumAgg(_FoldAgg):
    op = F.OP_SUM


class MinAgg(_FoldAgg):
    op = F.OP_MIN


class MaxAgg(_FoldAgg):
    op = F.OP_MAX


class PercentilesAgg(Agg):
    """percentiles_agg_f64[s] — src/metric/percentile.rs:130-138; Fruit = Percentiles<f64>"""

    def __init__(self, field, multi):
        self.field, self.multi = int(field), int(multi)

    def lower(self, ctx):
        self.node = ctx.emit(op=F.OP_PERCENTILES, kind=F.F64, multi=self.multi, field_id=self.field)
        return self.node

    def decode(self, reader, bucket):
        n, ranks, bits = reader.percentiles(self.node, bucket)
        return Percentiles(n, ranks.tolist(), bits.view(np.float64).tolist())


class TermsAgg(Agg):
    """terms_agg_{u64,i64}[s](field, sub) and filtered_terms_agg_* — src/bucket/terms.rs:185-195,391-401.
    Fruit = Terms<K, SubFruit>.  The key filter of filtered_terms_agg_* only decides which buckets
    exist (terms.rs:322-330), so it is applied while decoding."""

    def __init__(self, field, kind, multi, sub, key_filter=None):
        self.field, self.kind, self.multi = int(field), kind, int(multi)
        self.sub = as_agg(sub)
        self.key_filter = key_filter

    def lower(self, ctx):
        self.node = ctx.emit(op=F.OP_TERMS, kind=self.kind, multi=self.multi, field_id=self.field, n_children=1)
        self.sub.lower(ctx)
        return self.node

    def decode(self, reader, bucket):
        keys, children = reader.scope_children(self.node, bucket)
        res = {}
        for key_bits, child in zip(keys, children):
            key = codec.bits_to_value(self.kind, key_bits)
            if self.key_filter is not None and not self.key_filter(key):
                continue
            res[key] = self.sub.decode(reader, child)
        return Terms(res)

    def create_fruit(self):
        return Terms()

    def merge(self, acc, fruit):  # terms.rs:85-92
        for key, bucket in fruit.res.items():
            acc.res[key] = self.sub.merge(acc.res[key] if key in acc.res else self.sub.create_fruit(), bucket)
        return acc


class HistogramAgg(Agg):
    """histogram_agg_f64(field, start, interval, sub) — src/bucket/histogram.rs:9-21"""

    def __init__(self, field, start, interval, sub, kind=F.F64):
        self.field, self.start, self.interval = int(field), float(start), float(interval)
        self.sub = as_agg(sub)
        self.kind = kind  # f64 in the reference; date / i64 keys: date_histogram_agg below

    def lower(self, ctx):
        self.node = ctx.emit(op=F.OP_HISTOGRAM, kind=self.kind, field_id=self.field, n_children=1,
                             f0=self.start, f1=self.interval)
        self.sub.lower(ctx)
        return self.node

    def decode(self, reader, bucket):
        ords, children = reader.scope_children(self.node, bucket)
        return Histogram(self.start, self.interval,
                         {int(o): self.sub.decode(reader, c) for o, c in zip(ords, children)})

    def create_fruit(self):
        return Histogram(self.start, self.interval)

    def merge(self, acc, fruit):  # histogram.rs:90-97
        b = dict(acc._buckets)
        for o, bucket in fruit._buckets.items():
            b[o] = self.sub.merge(b[o] if o in b else self.sub.create_fruit(), bucket)
        return Histogram(acc.start, acc.interval, b)


class FilterAgg(Agg):
    """filter_agg(&query, sub) — src/filter.rs:8-16: narrows the doc stream by a second query."""

    def __init__(self, query, sub):
        self.query = query
        self.sub = as_agg(sub)

    def lower(self, ctx):
        aux = len(ctx.filters)
        ctx.filters.append(self.query)
        self.node = ctx.emit(op=F.OP_FILTER, n_children=1, aux=aux)
        self.sub.lower(ctx)
        return self.node

    def decode(self, reader, bucket):
        return self.sub.decode(reader, bucket)

    def create_fruit(self):
        return self.sub.create_fruit()

    def merge(self, acc, fruit):
        return self.sub.merge(acc, fruit)


# ---- predicates for post_filter_agg_* ---------------------------------------------------
class Pred:
    """A declarative predicate over a fast-field value; also callable like the Rust closure."""

    def code_ranges(self, kind):
        raise NotImplementedError


class _Cmp(Pred):
    def __init__(self, op, x):
        self.op, self.x = op, x

    def __call__(self, v):
        return {"gt": v > self.x, "ge": v >= self.x, "lt": v < self.x, "le": v <= self.x, "eq": v == self.x}[self.op]

    def code_range(self, kind):
        """Inclusive [lo, hi] on codes.  For f64 the NaN codes lie outside [code(-inf), code(+inf)],
        which reproduces IEEE comparisons with NaN being false."""
        if kind == F.F64:
            lo_all, hi_all = codec.scalar_code(kind, -np.inf), codec.scalar_code(kind, np.inf)
            if self.x != self.x:
                return 1, 0  # comparisons with NaN never pass
            c = codec.scalar_code(kind, self.x)
            if self.op == "eq" and self.x == 0.0:  # -0.0 == +0.0
                return codec.scalar_code(kind, -0.0), codec.scalar_code(kind, 0.0)
            if self.x == 0.0:  # order around the two zeros
                cneg, cpos = codec.scalar_code(kind, -0.0), codec.scalar_code(kind, 0.0)
                return {"gt": (cpos + 1, hi_all), "ge": (cneg, hi_all), "lt": (lo_all, cneg - 1),
                        "le": (lo_all, cpos)}[self.op]
        else:
            lo_all, hi_all = 0, (1 << 64) - 1
            c = codec.scalar_code(kind, self.x)
        return {"gt": (c + 1, hi_all), "ge": (c, hi_all), "lt": (lo_all, c - 1), "le": (lo_all, c),
                "eq": (c, c)}[self.op]


def gt(x):
    return _Cmp("gt", x)


def ge(x):
    return _Cmp("ge", x)


def lt(x):
    return _Cmp("lt", x)


def le(x):
    return _Cmp("le", x)


def eq(x):
    return _Cmp("eq", x)


class in_set(Pred):
    def __init__(self, values):
        self.values = set(values)

    def __call__(self, v):
        return v in self.values


class PostFilterAgg(Agg):
    """post_filter_agg_{u64,i64,f64}[s](field, filter, sub) — src/post_filter.rs:303-315.

    The reference evaluates a Rust closure per document.  Lowering (SURVEY §7 "Closures"):
      comparison predicates (gt/ge/lt/le/eq)            -> PRED_RANGE on codes (device)
      any callable when the column's code domain is small -> PRED_LUT, tabulated on the host (device)
      otherwise                                          -> host evaluates per value into a
                                                            per-segment bitset -> FILTER node.
    """

    def __init__(self, field, kind, multi, pred, sub):
        self.field, self.kind, self.multi = int(field), kind, int(multi)
        self.pred = pred
        self.sub = as_agg(sub)

    def lower(self, ctx):
        p = self.pred
        if isinstance(p, _Cmp):
            lo, hi = p.code_range(self.kind)
            if lo > hi:
                lo, hi = 1, 0
            self.node = ctx.emit(op=F.OP_POST_FILTER, kind=self.kind, multi=self.multi, field_id=self.field,
                                 n_children=1, pred=F.PRED_RANGE, u0=lo, u1=hi)
            self.sub.lower(ctx)
            return self.node
        dom = ctx.searcher.code_domain(self.field, self.multi) if ctx.searcher is not None else None
        if dom is not None and dom[1] - dom[0] + 1 <= LUT_MAX_DOMAIN:
            lo, hi = dom
            n = hi - lo + 1
            codes = np.arange(lo, hi + 1, dtype=np.uint64) if n else np.zeros(0, np.uint64)
            vals = [codec.bits_to_value(self.kind, b) for b in _codes_to_bits(self.kind, codes)]
            bits = np.fromiter((1 if p(v) else 0 for v in vals), dtype=np.uint8, count=n)
            blob = np.packbits(bits, bitorder="little").tobytes() or b"\0"
            aux = len(ctx.blobs)
            ctx.blobs.append(blob)
            self.node = ctx.emit(op=F.OP_POST_FILTER, kind=self.kind, multi=self.multi, field_id=self.field,
                                 n_children=1, pred=F.PRED_LUT, u0=lo, u1=n, aux=aux)
            self.sub.lower(ctx)
            return self.node
        # host route: the closure runs on the host, the device sees one more docset (FILTER)
        from .index import HostPredicateQuery
        aux = len(ctx.filters)
        ctx.filters.append(HostPredicateQuery(self.field, self.kind, self.multi, p))
        self.node = ctx.emit(op=F.OP_FILTER, n_children=1, aux=aux)
        self.sub.lower(ctx)
        return self.node

    def decode(self, reader, bucket):
        return self.sub.decode(reader, bucket)

    def create_fruit(self):
        return self.sub.create_fruit()

    def merge(self, acc, fruit):
        return self.sub.merge(acc, fruit)


def _codes_to_bits(kind, codes):
    if kind == F.U64:
        return codes
    if kind in (F.I64, F.DATE):
        return codec.code_to_i64(codes).view(np.uint64)
    return codec.code_to_f64(codes).view(np.uint64)


class GenericPostFilterAgg(Agg):
    """post_filter_agg(fetcher, filter, sub) — src/post_filter.rs:11-22: an arbitrary closure over
    any readers.  `filter(readers, doc)` runs on the host per document (readers = the segment's
    host view); the device sees the resulting bitset as a FILTER docset."""

    def __init__(self, fetcher, filt, sub):
        self.fetcher, self.filt = fetcher, filt
        self.sub = as_agg(sub)

    def lower(self, ctx):
        from .index import HostClosureQuery
        aux = len(ctx.filters)
        ctx.filters.append(HostClosureQuery(self.fetcher, self.filt))
        self.node = ctx.emit(op=F.OP_FILTER, n_children=1, aux=aux)
        self.sub.lower(ctx)
        return self.node

    def decode(self, reader, bucket):
        return self.sub.decode(reader, bucket)

    def create_fruit(self):
        return self.sub.create_fruit()

    def merge(self, acc, fruit):
        return self.sub.merge(acc, fruit)


class EitherAgg(Agg):
    """either_agg / one_of_agg — src/either.rs: a runtime choice between two aggs, resolved on the
    host before lowering (pure dispatch, no kernel).  `which` is 'left' or 'right'."""

    def __init__(self, which, agg, tag_result):
        self.which, self.agg, self.tag = which, as_agg(agg), tag_result

    def lower(self, ctx):
        return self.agg.lower(ctx)

    def decode(self, reader, bucket):
        f = self.agg.decode(reader, bucket)
        return (self.which, f) if self.tag else f

    def create_fruit(self):
        f = self.agg.create_fruit()
        return (self.which, f) if self.tag else f

    def merge(self, acc, fruit):
        if self.tag:
            return (self.which, self.agg.merge(acc[1], fruit[1]))
        return self.agg.merge(acc, fruit)


# ---- the reference's constructor functions --------------------------
def count_agg():
    return CountAgg()


def _fold_ctors(cls, name, kinds):
    out = {}
    for suffix, kind in kinds:
        out[f"{name}_agg_{suffix}"] = (lambda k: lambda field: cls(field, k, 0))(kind)
        out[f"{name}_agg_{suffix}s"] = (lambda k: lambda field: cls(field, k, 1))(kind)
    return out


_g = globals()
_g.update(_fold_ctors(SumAgg, "sum", [("u64", F.U64), ("i64", F.I64), ("f64", F.F64)]))            # sum.rs:146-158
_g.update(_fold_ctors(MinAgg, "min", [("u64", F.U64), ("i64", F.I64), ("f64", F.F64), ("date", F.DATE)]))  # minmax.rs:152-181
_g.update(_fold_ctors(MaxAgg, "max", [("u64", F.U64), ("i64", F.I64), ("f64", F.F64), ("date", F.DATE)]))


def percentiles_agg_f64(field):
    return PercentilesAgg(field, 0)


def percentiles_agg_f64s(field):
    return PercentilesAgg(field, 1)


def terms_agg_u64(field, sub):
    return TermsAgg(field, F.U64, 0, sub)


def terms_agg_i64(field, sub):
    return TermsAgg(field, F.I64, 0, sub)


def terms_agg_u64s(field, sub):
    return TermsAgg(field, F.U64, 1, sub)


def terms_agg_i64s(field, sub):
    return TermsAgg(field, F.I64, 1, sub)


def filtered_terms_agg_u64(field, sub, key_filter):
    return TermsAgg(field, F.U64, 0, sub, key_filter)


def filtered_terms_agg_i64(field, sub, key_filter):
    return TermsAgg(field, F.I64, 0, sub, key_filter)


def filtered_terms_agg_u64s(field, sub, key_filter):
    return TermsAgg(field, F.U64, 1, sub, key_filter)


def filtered_terms_agg_i64s(field, sub, key_filter):
    return TermsAgg(field, F.I64, 1, sub, key_filter)


def histogram_agg_f64(field, start, interval, sub):
    return HistogramAgg(field, start, interval, sub)


def filter_agg(query, sub):
    return FilterAgg(query, sub)


def post_filter_agg(fetcher, filt, sub):
    return GenericPostFilterAgg(fetcher, filt, sub)


def post_filter_agg_u64(field, pred, sub):
    return PostFilterAgg(field, F.U64, 0, pred, sub)


def post_filter_agg_i64(field, pred, sub):
    return PostFilterAgg(field, F.I64, 0, pred, sub)


def post_filter_agg_f64(field, pred, sub):
    return PostFilterAgg(field, F.F64, 0, pred, sub)


def post_filter_agg_u64s(field, pred, sub):
    return PostFilterAgg(field, F.U64, 1, pred, sub)


def post_filter_agg_i64s(field, pred, sub):
    return PostFilterAgg(field, F.I64, 1, pred, sub)


def post_filter_agg_f64s(field, pred, sub):
    return PostFilterAgg(field, F.F64, 1, pred, sub)


def either_agg(which, left, right):
    """src/either.rs:58-64 — different fruit types: result is ('left'|'right', fruit)."""
    return EitherAgg(which, left if which == "left" else right, True)


def one_of_agg(which, left, right):
    """src/either.rs:173-181 — same fruit type on both arms."""
    return EitherAgg(which, left if which == "left" else right, False)


# ---- beyond the reference: two entries of its TODO list (README.md:31-45) that are pure compositions -----------------
class Stats:
    """Fruit of stats_agg_*: count / sum / min / max of a field (avg derived) — README.md:35 "stat"."""

    def __init__(self, count, sum_, min_, max_):
        self.count, self.sum, self.min, self.max = count, sum_, min_, max_

    @property
    def avg(self):
        return None if not self.count or self.sum is None else self.sum / self.count

    def canon(self):
        return ("stats", self.count, self.sum, self.min, self.max)

    def __repr__(self):
        return f"Stats(count={self.count}, sum={self.sum}, min={self.min}, max={self.max})"


class StatsAgg(Agg):
    """stats_agg_{u64,i64,f64}[s](field): lowers to (count, sum, min, max) on one column — the fused root / bucket shape of
    the streaming kernel reads the column once for all four.  `count` counts VALUES for a multi-valued field (documents
    that reach the leaf for a single-valued one)."""

    def __init__(self, field, kind, multi):
        self.field, self.kind, self.multi = int(field), kind, int(multi)
        self.inner = TupleAgg((CountAgg(), SumAgg(field, kind, multi), MinAgg(field, kind, multi), MaxAgg(field, kind, multi)))

    def lower(self, ctx):
        return self.inner.lower(ctx)

    def decode(self, reader, bucket):
        c, s, mn, mx = self.inner.decode(reader, bucket)
        return Stats(c, s, mn, mx)

    def create_fruit(self):
        return Stats(0, None, None, None)

    def merge(self, acc, fruit):
        c, s, mn, mx = self.inner.merge((acc.count, acc.sum, acc.min, acc.max), (fruit.count, fruit.sum, fruit.min, fruit.max))
        return Stats(c, s, mn, mx)


def stats_agg_u64(field):
    return StatsAgg(field, F.U64, 0)


def stats_agg_i64(field):
    return StatsAgg(field, F.I64, 0)


def stats_agg_f64(field):
    return StatsAgg(field, F.F64, 0)


def filters_agg(named_queries, sub_factory):
    """filters_agg({name: query, ...}, lambda: sub) — README.md:40 "filters": one filter_agg per named query over the same
    sub-aggregation, evaluated in ONE pass (a tuple of FILTER nodes; at most 8 per plan).  Fruit: {name: sub fruit}."""
    names = list(named_queries)
    if not 2 <= len(names) <= 8:
        raise TypeError("filters_agg takes 2..=8 named queries")
    return _FiltersAgg(names, [FilterAgg(named_queries[n], sub_factory()) for n in names])


class _FiltersAgg(Agg):
    def __init__(self, names, members):
        self.names, self.inner = names, TupleAgg(tuple(members))

    def lower(self, ctx):
        return self.inner.lower(ctx)

    def decode(self, reader, bucket):
        return dict(zip(self.names, self.inner.decode(reader, bucket)))

    def create_fruit(self):
        return dict(zip(self.names, self.inner.create_fruit()))

    def merge(self, acc, fruit):
        merged = self.inner.merge(tuple(acc[n] for n in self.names), tuple(fruit[n] for n in self.names))
        return dict(zip(self.names, merged))


def date_histogram_agg(field, interval_seconds, sub, start=0, kind=F.DATE):
    """date_histogram_agg(field, interval, sub) — README.md:41 "date_histogram": fixed-width buckets over a date fast field
    (seconds since the epoch, i64 codec — SURVEY §8 E1), ordinal = floor((t - start) / interval) by the arithmetic of
    histogram.rs:136-152 on the timestamp as f64 (exact below 2^53); documents before `start` are skipped like the
    reference's `n < 0`.  Fruit: Histogram whose bucket keys are `start + ord * interval` (timestamps)."""
    if not interval_seconds > 0:
        raise ValueError("date_histogram_agg needs a positive interval")
    return HistogramAgg(field, float(start), float(interval_seconds), sub, kind=kind)


class Cardinality:
    """Fruit of cardinality_agg_*: the EXACT number of distinct values (an HLL sketch, README.md:36, would estimate it);
    the distinct keys ride along so fruits of different shards merge exactly."""

    def __init__(self, keys=()):
        self.keys = set(keys)

    @property
    def value(self):
        return len(self.keys)

    def canon(self):
        return ("cardinality", tuple(sorted(self.keys)))

    def __repr__(self):
        return f"Cardinality({self.value})"


class CardinalityAgg(Agg):
    """cardinality_agg_{u64,i64}[s](field) — README.md:36 "cardinality": lowers to terms_agg(field, count_agg()); the
    bucket table the pass builds anyway IS the distinct set (dense tables: one existence byte per value of the column's
    domain; wide domains: the open-addressing spill table), compacted on the device."""

    def __init__(self, field, kind, multi):
        self.inner = TermsAgg(field, kind, multi, CountAgg())

    def lower(self, ctx):
        return self.inner.lower(ctx)

    def decode(self, reader, bucket):
        keys, _ = reader.scope_children(self.inner.node, bucket)
        return Cardinality(codec.bits_to_value(self.inner.kind, k) for k in keys)

    def create_fruit(self):
        return Cardinality()

    def merge(self, acc, fruit):
        return Cardinality(acc.keys | fruit.keys)


def cardinality_agg_u64(field):
    return CardinalityAgg(field, F.U64, 0)


def cardinality_agg_i64(field):
    return CardinalityAgg(field, F.I64, 0)


def cardinality_agg_u64s(field):
    return CardinalityAgg(field, F.U64, 1)


def cardinality_agg_i64s(field):
    return CardinalityAgg(field, F.I64, 1)


FOLD_CTORS = sorted(k for k in _g if k.startswith(("sum_agg_", "min_agg_", "max_agg_")))
