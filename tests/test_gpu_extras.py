"""GPU: two entries of the reference's TODO list (README.md:31-45) that are compositions of the existing nodes —
stats_agg_* (count + sum + min + max of one column in one fused pass) and filters_agg (several filter_agg over one
sub-aggregation in one pass).  No reference implementation exists: checked against numpy and against the equivalent
explicit tuples through the oracle."""
import numpy as np
import pytest

import tantivy_aggregations_b200 as ta
from helpers import Corpus, SegSpec, assert_fruit_equal
from tantivy_aggregations_b200 import _ffi as F

pytestmark = pytest.mark.gpu
CAT, PRICE, QTY = 1, 2, 3


@pytest.fixture(scope="module")
def world(ctx):
    rng = np.random.default_rng(21)
    segs = []
    for n in (30_000, 17_001):
        s = SegSpec(n)
        s.col(CAT, F.U64, rng.integers(1, 40, size=n, dtype=np.uint64))
        s.col(PRICE, F.F64, 1.0 + 100.0 * rng.random(n))
        s.col(QTY, F.I64, rng.integers(-20, 20, size=n, dtype=np.int64))
        segs.append(s)
    corpus = Corpus(segs)
    return corpus, corpus.build_gpu(ctx), corpus.build_oracle()


def test_stats_root_and_nested(world):
    corpus, searcher, ox = world
    from tantivy_aggregations_b200 import codec
    prices = np.concatenate([codec.code_to_f64(s.cols[PRICE][1]) for s in corpus.segs])
    st = searcher.agg_search(ta.AllQuery(), ta.stats_agg_f64(PRICE))
    assert st.count == len(prices) and st.min == prices.min() and st.max == prices.max()
    assert abs(st.sum - prices.sum()) <= 1e-12 * abs(prices.sum()) and abs(st.avg - prices.mean()) <= 1e-12 * prices.mean()
    # nested under terms == the explicit tuple through the oracle
    got = searcher.agg_search(ta.AllQuery(), ta.terms_agg_u64(CAT, ta.stats_agg_i64(QTY)))
    want, _, _ = ox.search(ta.AllQuery(), ta.terms_agg_u64(CAT, (ta.count_agg(), ta.sum_agg_i64(QTY), ta.min_agg_i64(QTY), ta.max_agg_i64(QTY))))
    assert set(got.res) == set(want.res)
    for k, s in got.res.items():
        assert (s.count, s.sum, s.min, s.max) == want.res[k]


def test_filters_agg_one_pass(world):
    corpus, searcher, ox = world
    named = {"cheap": ta.RangeQuery.half_open(PRICE, F.F64, 0.0, 20.0), "mid": ta.RangeQuery.half_open(PRICE, F.F64, 20.0, 60.0),
             "neg": ta.RangeQuery(QTY, F.I64, -20, -1)}
    sub = lambda: (ta.count_agg(), ta.max_agg_f64(PRICE))
    got, reader = searcher.agg_search_with_executor(ta.AllQuery(), ta.filters_agg(named, sub), ta.SINGLE_THREAD, return_reader=True)
    for name, q in named.items():
        want, _, _ = ox.search(ta.AllQuery(), ta.filter_agg(q, sub()))
        assert_fruit_equal(got[name], want, 1e-12, name)


DATE_F, TAGS_F = 4, 5


@pytest.fixture(scope="module")
def dated(ctx):
    rng = np.random.default_rng(33)
    segs = []
    for n in (40_000, 0, 9_001):
        s = SegSpec(n)
        s.col(CAT, F.U64, rng.integers(1, 12, size=n, dtype=np.uint64))
        # two years of second-resolution timestamps around a day boundary, a few before the epoch
        t = rng.integers(1_500_000_000, 1_563_000_000, size=n, dtype=np.int64)
        t[: n // 100] = rng.integers(-86_400 * 3, 86_400 * 3, size=n // 100)
        s.col(DATE_F, F.DATE, t)
        s.col(PRICE, F.F64, 1.0 + 100.0 * rng.random(n))
        s.mcol(TAGS_F, F.U64, [list(rng.integers(0, 3000, size=rng.integers(0, 4), dtype=np.uint64)) for _ in range(n)])
        s.deleted = range(0, n, 13)
        segs.append(s)
    corpus = Corpus(segs)
    return corpus, corpus.build_gpu(ctx), corpus.build_oracle()


def _alive(s):
    m = np.ones(s.max_doc, dtype=bool)
    m[list(s.deleted)] = False
    return m


def test_date_histogram(dated):
    """README.md:41 date_histogram: daily / 30-day buckets over a date fast field, against the oracle (same arithmetic,
    histogram.rs:136-152 on the timestamp) and against numpy's integer floor division."""
    from tantivy_aggregations_b200 import codec
    corpus, searcher, ox = dated
    ts = np.concatenate([codec.code_to_i64(s.cols[DATE_F][1])[_alive(s)] for s in corpus.segs])
    for interval, start in ((86_400, 0), (30 * 86_400, 1_500_000_000), (3_600, -86_400 * 3)):
        mk = lambda: ta.date_histogram_agg(DATE_F, interval, (ta.count_agg(), ta.max_agg_f64(PRICE)), start=start)
        got = searcher.agg_search(ta.AllQuery(), mk())
        want, _, _ = ox.search(ta.AllQuery(), mk())
        assert_fruit_equal(got, want, 1e-12, f"date_histogram {interval}")
        ok = ts >= start
        ords, counts = np.unique((ts[ok] - start) // interval, return_counts=True)
        assert {o: v[0] for o, v in got._buckets.items()} == {int(o): int(c) for o, c in zip(ords, counts)}
        keys = [k for k, _ in got.buckets()]
        assert keys[0] == start + int(ords[0]) * interval
    # nested under terms, under a filter
    mk = lambda: ta.filter_agg(ta.RangeQuery.half_open(PRICE, F.F64, 10.0, 60.0),
                               ta.terms_agg_u64(CAT, ta.date_histogram_agg(DATE_F, 7 * 86_400, ta.count_agg())))
    want, _, _ = ox.search(ta.AllQuery(), mk())
    assert_fruit_equal(searcher.agg_search(ta.AllQuery(), mk()), want, 1e-12, "nested date_histogram")


def test_cardinality(dated):
    """README.md:36 cardinality: the exact distinct count of a field — root, multi-valued, nested per bucket — against
    numpy, and merged across shards like any fruit."""
    from tantivy_aggregations_b200 import codec
    corpus, searcher, ox = dated
    cats = np.concatenate([s.cols[CAT][1][_alive(s)] for s in corpus.segs])
    assert searcher.agg_search(ta.AllQuery(), ta.cardinality_agg_u64(CAT)).value == len(np.unique(cats))
    tags = set()
    for s in corpus.segs:
        kind, off, codes = s.mcols[TAGS_F]
        al = _alive(s)
        for d in np.nonzero(al)[0]:
            tags.update(codes[int(off[d]):int(off[d + 1])].tolist())
    got = searcher.agg_search(ta.AllQuery(), ta.cardinality_agg_u64s(TAGS_F))
    assert got.value == len(tags) and got.keys == tags
    # nested: distinct days per category == the explicit terms through the oracle
    day = lambda: ta.terms_agg_u64(CAT, ta.cardinality_agg_u64s(TAGS_F))
    got = searcher.agg_search(ta.AllQuery(), day())
    want, _, _ = ox.search(ta.AllQuery(), ta.terms_agg_u64(CAT, ta.terms_agg_u64s(TAGS_F, ta.count_agg())))
    assert {k: c.value for k, c in got.res.items()} == {k: len(t.res) for k, t in want.res.items()}
    # shard merge (PreparedAgg::merge on the host): two halves of the index
    agg = ta.cardinality_agg_u64s(TAGS_F)
    a = ta.Searcher(searcher.ctx, searcher.segments[:1]).agg_search(ta.AllQuery(), ta.cardinality_agg_u64s(TAGS_F))
    b = ta.Searcher(searcher.ctx, searcher.segments[1:]).agg_search(ta.AllQuery(), ta.cardinality_agg_u64s(TAGS_F))
    assert agg.merge(agg.merge(agg.create_fruit(), a), b).value == len(tags)


@pytest.mark.parametrize("n", [30_000, 9_000_000])  # below / above the rank-bin threshold of the percentile pass
def test_top_hits(ctx, n):
    """README.md:42 top_hits: the k matched documents with the largest / smallest values of an f64 field, ties in document
    order, against numpy — deletes and a bitset query in front, duplicates planted around the cut."""
    rng = np.random.default_rng(n)
    half = n // 3
    vals = np.round(1.0 + 100.0 * rng.random(n), 3)
    vals[rng.integers(0, n, size=50)] = 100.999  # ties at the top
    vals[rng.integers(0, n, size=50)] = 1.0      # and at the bottom
    segs = [SegSpec(half).col(PRICE, F.F64, vals[:half]), SegSpec(n - half).col(PRICE, F.F64, vals[half:])]
    segs[0].deleted = range(0, half, 5)
    searcher = Corpus(segs).build_gpu(ctx)
    m = rng.random(n) < 0.7
    q = ta.BitsetQuery({0: np.packbits(m[:half].astype(np.uint8), bitorder="little"), 1: np.packbits(m[half:].astype(np.uint8), bitorder="little")})
    alive = m.copy()
    alive[np.arange(0, half, 5)] = False
    idx = np.nonzero(alive)[0]
    seg_of = (idx >= half).astype(np.int64)
    doc_of = np.where(idx >= half, idx - half, idx)
    for k in (1, 10, 77):
        for desc in (True, False):
            order = np.lexsort((doc_of, seg_of, -vals[idx] if desc else vals[idx]))[:k]
            want = [(float(vals[idx[o]]), int(seg_of[o]), int(doc_of[o])) for o in order]
            assert searcher.top_hits_f64(q, PRICE, k, descending=desc) == want
    assert searcher.top_hits_f64(q, PRICE, 0) == []
    few = searcher.top_hits_f64(ta.DocIdsQuery({0: np.array([1, 2, 3], dtype=np.uint32), 1: np.array([], dtype=np.uint32)}), PRICE, 10)
    assert [h[1:] for h in sorted(few, key=lambda h: h[2])] == [(0, 1), (0, 2), (0, 3)]
