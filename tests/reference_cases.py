"""The reference's own 19 unit tests, transcribed as known-answer cases.

Each case is `(name, source, run)` where `run(search, S)` performs the reference test's searches
through `search(query, agg) -> fruit` and asserts the reference's expected values.  The same cases
pin the CPU oracle (tests/test_oracle_golden.py, no GPU) and the CUDA path (tests/test_gpu_reference.py).
`search_empty` searches the empty index of `test_empty_terms_agg`.
"""
import tantivy_aggregations_b200 as ta
from tantivy_aggregations_b200 import F64, U64

D = lambda s: s  # dates are unix seconds (chrono DateTime<Utc>::timestamp())


def price_range(S, lo, hi):
    """RangeQuery::new_f64(price, lo..hi) — half-open"""
    return ta.RangeQuery.half_open(S.price, F64, lo, hi)


def cat_term(S, v):
    """ProductIndex::category_query / TermQuery on category_id"""
    return ta.TermQuery(S.category_id, U64, v)


def t_count(search, S, **_):  # src/metric/count.rs:69-80
    assert search(ta.AllQuery(), ta.count_agg()) == 5


def t_sum(search, S, **_):  # src/metric/sum.rs:171-192
    assert search(ta.AllQuery(), ta.sum_agg_u64(S.positive_opinion_percent)) == 437
    assert search(ta.AllQuery(), ta.sum_agg_f64(S.price)) == 170.5
    assert search(ta.AllQuery(), ta.sum_agg_u64s(S.tag_ids)) == 2740


def t_min(search, S, **_):  # src/metric/minmax.rs:197-225
    assert search(ta.AllQuery(), ta.min_agg_u64(S.positive_opinion_percent)) == 71
    assert search(ta.AllQuery(), ta.min_agg_date(S.date_created)) == 0  # 1970-01-01T00:00:00Z: missing value
    assert search(ta.AllQuery(), ta.min_agg_f64(S.price)) == 0.5
    assert search(ta.AllQuery(), ta.min_agg_u64s(S.tag_ids)) == 111


def t_max(search, S, **_):  # src/metric/minmax.rs:228-256
    assert search(ta.AllQuery(), ta.max_agg_f64(S.price)) == 100.01
    assert search(ta.AllQuery(), ta.max_agg_u64(S.positive_opinion_percent)) == 100
    assert search(ta.AllQuery(), ta.max_agg_date(S.date_created)) == 1577840399  # 2020-01-01T00:59:59Z
    assert search(ta.AllQuery(), ta.max_agg_u64s(S.tag_ids)) == 511


def t_tuple(search, S, **_):  # src/tuple.rs:93-110
    agg = (ta.count_agg(), ta.min_agg_f64(S.price), ta.max_agg_f64(S.price))
    assert search(ta.AllQuery(), agg) == (5, 0.5, 100.01)


def t_percentiles(search, S, **_):  # src/metric/percentile.rs:190-221
    p = search(ta.AllQuery(), ta.percentiles_agg_f64(S.price))
    assert p.percentile(0.5) == 10.0
    assert p.percentile(0.33) == 9.99
    assert p.percentile(0.7) == 50.0
    assert p.percentile(0.01) == 0.5
    assert p.percentile(0.99) == 100.01


def t_empty_terms(search, S, search_empty=None, **_):  # src/bucket/terms.rs:473-487
    r = search_empty(ta.AllQuery(), ta.terms_agg_u64(S.category_id, ta.count_agg()))
    assert r.top_k(10, lambda b: b) == []


def t_terms(search, S, **_):  # src/bucket/terms.rs:490-543
    r = search(ta.AllQuery(), ta.terms_agg_u64(S.category_id, (ta.count_agg(), ta.min_agg_f64(S.price))))
    assert r.get(1) == (2, 9.99)
    assert r.get(2) == (3, 0.5)
    assert r.top_k(2, lambda b: b[0]) == [(2, (3, 0.5)), (1, (2, 9.99))]
    import struct
    le = lambda v: struct.pack("<d", v)  # "Floats are hard to sort": the reference sorts by to_le_bytes()

    class Rev:
        def __init__(self, v): self.v = v
        def __lt__(self, o): return o.v < self.v
        def __eq__(self, o): return o.v == self.v
    assert r.top_k(1, lambda b: Rev(le(b[1]))) == [(2, (3, 0.5))]
    assert r.top_k(1, lambda b: le(b[1])) == [(1, (2, 9.99))]


def t_filtered_terms(search, S, **_):  # src/bucket/terms.rs:546-571
    agg = ta.filtered_terms_agg_u64(S.category_id, (ta.count_agg(), ta.min_agg_f64(S.price)), lambda c: c % 2 == 0)
    r = search(ta.AllQuery(), agg)
    assert r.get(1) is None
    assert r.get(2) == (3, 0.5)


def t_histogram(search, S, **_):  # src/bucket/histogram.rs:195-223
    h = search(ta.AllQuery(), ta.histogram_agg_f64(S.price, 0.0, 10.0, ta.count_agg()))
    assert h.buckets() == [(0.0, 2), (10.0, 1), (20.0, None), (30.0, None), (40.0, None), (50.0, 1), (60.0, None),
                           (70.0, None), (80.0, None), (90.0, None), (100.0, 1)]


def t_histogram_custom_start(search, S, **_):  # src/bucket/histogram.rs:226-249
    h = search(ta.AllQuery(), ta.histogram_agg_f64(S.price, 35.0, 10.0, ta.count_agg()))
    assert h.buckets() == [(45.0, 1), (55.0, None), (65.0, None), (75.0, None), (85.0, None), (95.0, 1)]


def t_nested_histogram(search, S, **_):  # src/bucket/histogram.rs:252-338
    agg = ta.terms_agg_u64s(S.tag_ids, (ta.count_agg(), ta.histogram_agg_f64(S.price, 0.0, 10.0, ta.count_agg())))
    r = search(ta.AllQuery(), agg)
    top = r.top_k(3, lambda b: b[0])
    assert top[0][0] == 211 and top[0][1][0] == 3
    assert top[0][1][1].buckets() == [(0.0, 2), (10.0, 1)]
    expected = {
        111: [(0.0, 1), (10.0, 1)],
        311: [(0.0, 1)] + [(float(x), None) for x in range(10, 100, 10)] + [(100.0, 1)],
        320: [(10.0, 1), (20.0, None), (30.0, None), (40.0, None), (50.0, 1)],
    }
    for tag, f in top[1:]:
        assert f[0] == 2
        assert tag in expected, f"Unexpected tag: {tag}"
        assert f[1].buckets() == expected[tag]
    # beyond the reference's top-3 view: every tag's bucket, checked for all four candidates
    for tag, want in expected.items():
        assert r.get(tag)[0] == 2 and r.get(tag)[1].buckets() == want


def t_filtered_histogram(search, S, **_):  # src/bucket/histogram.rs:341-367
    agg = ta.filter_agg(price_range(S, 10.0, 100.0), ta.histogram_agg_f64(S.price, 0.0, 10.0, ta.count_agg()))
    h = search(ta.AllQuery(), agg)
    assert h.buckets() == [(10.0, 1), (20.0, None), (30.0, None), (40.0, None), (50.0, 1)]


def t_filter(search, S, **_):  # src/filter.rs:137-166
    assert search(ta.AllQuery(), ta.filter_agg(cat_term(S, 1), ta.count_agg())) == 2
    assert search(price_range(S, 100.0, 200.0), ta.filter_agg(cat_term(S, 2), ta.count_agg())) == 1


def t_post_filter_f64(search, S, **_):  # src/post_filter.rs:330-345
    assert search(ta.AllQuery(), ta.post_filter_agg_f64(S.price, ta.gt(5.0), ta.count_agg())) == 4
    # the same closure given as an opaque callable (lowered through the LUT / host route)
    assert search(ta.AllQuery(), ta.post_filter_agg_f64(S.price, lambda price: price > 5.0, ta.count_agg())) == 4


def t_post_filter_u64s(search, S, **_):  # src/post_filter.rs:348-366
    tags = {111, 211, 311}
    assert search(ta.AllQuery(), ta.post_filter_agg_u64s(S.tag_ids, lambda t: t in tags, ta.count_agg())) == 4
    assert search(ta.AllQuery(), ta.post_filter_agg_u64s(S.tag_ids, ta.in_set(tags), ta.count_agg())) == 4


def t_post_filter_generic(search, S, **_):  # src/post_filter.rs:369-405
    agg = ta.post_filter_agg(lambda ctx: ctx, lambda ff, doc: ff.get(S.price, doc) > 5.0, ta.count_agg())
    assert search(ta.AllQuery(), agg) == 4
    agg = ta.post_filter_agg(lambda ctx: ctx,
                             lambda ff, doc: ff.get(S.price, doc) >= 10.0 and ff.get(S.category_id, doc) == 1,
                             ta.count_agg())
    assert search(ta.AllQuery(), agg) == 1


def t_either(search, S, **_):  # src/either.rs:294-309
    left, right = ta.count_agg(), (ta.min_agg_f64(S.price), ta.max_agg_f64(S.price))
    assert search(ta.AllQuery(), ta.either_agg("left", left, right)) == ("left", 5)
    assert search(ta.AllQuery(), ta.either_agg("right", left, right)) == ("right", (0.5, 100.01))


def t_one_of(search, S, **_):  # src/either.rs:322-338
    left = ta.count_agg()
    right = ta.filter_agg(cat_term(S, 1), ta.count_agg())
    assert search(ta.AllQuery(), ta.one_of_agg("left", left, right)) == 5
    assert search(ta.AllQuery(), ta.one_of_agg("right", left, right)) == 2


CASES = [
    ("count", "src/metric/count.rs:69-80", t_count),
    ("sum", "src/metric/sum.rs:171-192", t_sum),
    ("min", "src/metric/minmax.rs:197-225", t_min),
    ("max", "src/metric/minmax.rs:228-256", t_max),
    ("tuple", "src/tuple.rs:93-110", t_tuple),
    ("percentiles", "src/metric/percentile.rs:190-221", t_percentiles),
    ("empty_terms", "src/bucket/terms.rs:473-487", t_empty_terms),
    ("terms", "src/bucket/terms.rs:490-543", t_terms),
    ("filtered_terms", "src/bucket/terms.rs:546-571", t_filtered_terms),
    ("histogram", "src/bucket/histogram.rs:195-223", t_histogram),
    ("histogram_custom_start", "src/bucket/histogram.rs:226-249", t_histogram_custom_start),
    ("nested_histogram", "src/bucket/histogram.rs:252-338", t_nested_histogram),
    ("filtered_histogram", "src/bucket/histogram.rs:341-367", t_filtered_histogram),
    ("filter", "src/filter.rs:137-166", t_filter),
    ("post_filter_f64", "src/post_filter.rs:330-345", t_post_filter_f64),
    ("post_filter_u64s", "src/post_filter.rs:348-366", t_post_filter_u64s),
    ("post_filter_generic", "src/post_filter.rs:369-405", t_post_filter_generic),
    ("either", "src/either.rs:294-309", t_either),
    ("one_of", "src/either.rs:322-338", t_one_of),
]
