"""GPU: `Terms::top_k` (reference src/bucket/terms.rs:425-457) selected on the device from the HBM-resident fruit image,
with a lazy read-out (only the k winners cross PCIe), against `Terms.top_k` of the oracle's fruit.  The order is the
reference's: sort value descending, equal values ascending by key (which of several buckets tied AT the cut survive is
HashMap-order dependent in the reference; here it is always the smallest keys)."""
import numpy as np
import pytest

import tantivy_aggregations_b200 as ta
from helpers import Corpus, SegSpec
from tantivy_aggregations_b200 import _ffi as F
from tantivy_aggregations_b200 import codec

pytestmark = pytest.mark.gpu
CAT, PRICE, QTY, WIDE = 1, 2, 3, 4


@pytest.fixture(scope="module")
def world(ctx):
    rng = np.random.default_rng(5)
    segs = []
    for n in (60_000, 45_001):
        s = SegSpec(n)
        s.col(CAT, F.U64, rng.zipf(1.3, size=n).astype(np.uint64) % np.uint64(20_000) + np.uint64(1))
        s.col(PRICE, F.F64, np.round(1.0 + 100.0 * rng.random(n), 2))
        s.col(QTY, F.I64, rng.integers(-50, 50, size=n, dtype=np.int64))
        s.col(WIDE, F.U64, (rng.integers(1, 3000, size=n, dtype=np.uint64) << np.uint64(36)) | np.uint64(7))  # hashed scope
        s.deleted = rng.choice(n, size=n // 9, replace=False)
        segs.append(s)
    corpus = Corpus(segs)
    return corpus, corpus.build_gpu(ctx), corpus.build_oracle()


def subs():
    return lambda: (ta.count_agg(), ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE), ta.sum_agg_i64(QTY))


@pytest.mark.parametrize("field", [CAT, WIDE])
@pytest.mark.parametrize("by", [0, 1, 2, 3])
@pytest.mark.parametrize("k", [1, 10, 257, 5000, 100_000])
def test_top_k_matches_the_fruit_accessor(world, field, by, k):
    corpus, searcher, ox = world
    mk = lambda: ta.terms_agg_u64(field, subs()())
    want_fruit, _, _ = ox.search(ta.AllQuery(), mk())
    want = want_fruit.top_k(k, lambda b: (b[by] is not None, b[by]) if by else b[by])
    agg = mk()
    got = searcher.terms_top_k(ta.AllQuery(), agg, agg, agg.sub.members[by], k)
    assert [key for key, _ in got] == [key for key, _ in want]
    for (_, g), (_, w) in zip(got, want):
        assert g == w


def test_top_k_under_a_filter_and_with_ties(world):
    """counts tie massively (a Zipf tail of buckets with 1..3 docs): the cut falls inside a tie group."""
    corpus, searcher, ox = world
    mk = lambda: ta.filter_agg(ta.RangeQuery(QTY, F.I64, 0, 49), (ta.count_agg(), ta.terms_agg_u64(CAT, (ta.count_agg(), ta.sum_agg_i64(QTY)))))
    want_fruit, _, _ = ox.search(ta.AllQuery(), mk())
    for k in (3, 700, 3000):
        agg = mk()
        terms = agg.sub.members[1]
        got = searcher.terms_top_k(ta.AllQuery(), agg, terms, terms.sub.members[0], k)
        want = want_fruit[1].top_k(k, lambda b: b[0])
        assert got == want


def test_lazy_result_still_answers_the_plain_readers(world, ctx):
    corpus, searcher, ox = world
    agg = ta.terms_agg_u64(CAT, subs()())
    plan = searcher.prepare(agg)
    plan.set_readout(F.READOUT_LAZY)
    got = searcher.agg_search(ta.AllQuery(), plan)
    want, _, _ = ox.search(ta.AllQuery(), ta.terms_agg_u64(CAT, subs()()))
    from helpers import assert_fruit_equal
    assert_fruit_equal(got, want)


def test_top_k_edge_sizes_and_a_nested_scope(world):
    """k = 0, a query that matches nothing, and the inner scope of terms-in-terms selected per parent bucket."""
    corpus, searcher, ox = world
    agg = ta.terms_agg_u64(CAT, subs()())
    assert searcher.terms_top_k(ta.AllQuery(), agg, agg, agg.sub.members[0], 0) == []
    none = ta.DocIdsQuery({i: np.zeros(0, np.uint32) for i in range(len(corpus.segs))})
    agg = ta.terms_agg_u64(CAT, subs()())
    assert searcher.terms_top_k(none, agg, agg, agg.sub.members[0], 10) == []
    # nested: outer = sign class of QTY via a histogram-free trick (terms on a small i64 field), inner = terms(CAT, count)
    mk = lambda: ta.terms_agg_i64(QTY, ta.terms_agg_u64(CAT, (ta.count_agg(), ta.max_agg_f64(PRICE))))
    q = ta.RangeQuery(QTY, F.I64, -3, 3, device=False)
    want, _, _ = ox.search(q, mk())
    agg = mk()
    got, reader = searcher.agg_search_with_executor(q, agg, ta.SINGLE_THREAD, return_reader=True)
    inner = agg.sub
    outer_keys, _ = reader.scope(agg.node)
    assert len(outer_keys) == 7
    for p, kb in enumerate(outer_keys.tolist()):
        okey = codec.bits_to_value(F.I64, kb)
        for by, k in ((0, 5), (1, 40), (0, 10**6)):
            idx = reader.top_k(inner.node, inner.sub.members[by].node, k, parent_bucket=p)
            keys, parents = reader.scope_rows(inner.node, idx)
            assert (parents == p).all()
            w = want.get(okey).top_k(k, lambda b: b[by])
            assert keys.tolist() == [key for key, _ in w]
