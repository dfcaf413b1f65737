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
