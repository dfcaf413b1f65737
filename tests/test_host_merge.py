"""CPU-only: `PreparedAgg::merge` of the host facade (tantivy_aggregations_b200/agg.py) on decoded fruits — what combines the
fruits of different processes when tables cannot be merged cell by cell (hashed scopes, percentile summaries) and what the
TODO-list compositions add on top (stats, filters, cardinality, date_histogram).  Reference semantics: count.rs:39-41,
sum.rs:59-70, minmax.rs:59-72, terms.rs:85-92, histogram.rs:90-97."""
import tantivy_aggregations_b200 as ta
from tantivy_aggregations_b200 import _ffi as F
from tantivy_aggregations_b200.agg import Cardinality, Stats
from tantivy_aggregations_b200.fruits import Histogram, Terms


def fold(agg, fruits):
    acc = agg.create_fruit()
    for f in fruits:
        acc = agg.merge(acc, f)
    return acc


def test_leaf_merges_follow_the_reference():
    assert fold(ta.count_agg(), [3, 0, 4]) == 7
    assert fold(ta.sum_agg_f64(0), [None, 1.5, None, 2.25]) == 3.75 and fold(ta.sum_agg_f64(0), [None, None]) is None
    assert fold(ta.sum_agg_u64(0), [(1 << 64) - 1, 2]) == 1                      # wrapping, like release-mode Rust
    assert fold(ta.sum_agg_i64(0), [(1 << 63) - 1, 1]) == -(1 << 63)
    assert fold(ta.min_agg_i64(0), [5, None, -7, 3]) == -7 and fold(ta.max_agg_f64(0), [None, 2.0, 9.5]) == 9.5
    # minmax.rs:59-72: replace only on a strict improvement — the first of two equal extremes stays (-0.0 vs +0.0)
    m = fold(ta.min_agg_f64(0), [0.0, -0.0])
    assert m == 0.0 and str(m) == "0.0"


def test_bucket_merges_create_missing_buckets():
    agg = ta.terms_agg_u64(1, (ta.count_agg(), ta.max_agg_f64(2)))
    a = Terms({1: (2, 5.0), 2: (1, None)})
    b = Terms({2: (4, 7.0), 9: (1, 1.0)})
    got = fold(agg, [a, b])
    assert got.res == {1: (2, 5.0), 2: (5, 7.0), 9: (1, 1.0)}
    h = ta.date_histogram_agg(3, 86_400, ta.count_agg(), start=0)
    got = fold(h, [Histogram(0.0, 86400.0, {18261: 2}), Histogram(0.0, 86400.0, {0: 1, 18261: 1})])
    assert dict(got._buckets) == {0: 1, 18261: 3}
    assert [k for k, _ in fold(ta.histogram_agg_f64(2, 0.0, 10.0, ta.count_agg()), [Histogram(0.0, 10.0, {1: 1, 3: 1})]).buckets()] == [10.0, 20.0, 30.0]


def test_todo_list_compositions_merge():
    s = fold(ta.stats_agg_f64(2), [Stats(2, 3.0, 1.0, 2.0), Stats(0, None, None, None), Stats(1, 10.0, 10.0, 10.0)])
    assert (s.count, s.sum, s.min, s.max) == (3, 13.0, 1.0, 10.0) and abs(s.avg - 13.0 / 3) < 1e-15
    c = fold(ta.cardinality_agg_u64(1), [Cardinality([1, 2, 3]), Cardinality([3, 4])])
    assert c.value == 4 and c.keys == {1, 2, 3, 4}
    f = ta.filters_agg({"a": ta.AllQuery(), "b": ta.AllQuery()}, lambda: (ta.count_agg(), ta.min_agg_f64(2)))
    got = fold(f, [{"a": (1, 2.0), "b": (0, None)}, {"a": (2, 1.0), "b": (3, 4.0)}])
    assert got == {"a": (3, 1.0), "b": (3, 4.0)}
    t = ta.terms_agg_u64(1, ta.cardinality_agg_u64s(4))
    got = fold(t, [Terms({7: Cardinality([1])}), Terms({7: Cardinality([1, 2]), 8: Cardinality([5])})])
    assert {k: v.value for k, v in got.res.items()} == {7: 2, 8: 1}
