"""GPU: percentiles_agg_f64[s] nested under bucket aggregations (any Agg can be a sub-aggregation of terms / histogram,
terms.rs:127-132, histogram.rs:136-152).  Per bucket the fruit must hold exact order statistics: equal to the oracle's
CKMS while that is uncompressed (small buckets), inside CKMS's +-eps*q*n rank band otherwise."""
import numpy as np
import pytest

import tantivy_aggregations_b200 as ta
from helpers import Corpus, SegSpec, exact_rank_window
from tantivy_aggregations_b200 import _ffi as F
from tantivy_aggregations_b200 import codec

pytestmark = pytest.mark.gpu
CAT, PRICE, WIDE, FVALS = 0, 1, 2, 3
EPS = 0.01
QS = (0.0, 0.01, 0.1, 0.25, 0.5, 0.75, 0.9, 0.99, 1.0)


def build(seed, seg_sizes, n_cat):
    rng = np.random.default_rng(seed)
    segs, cats, prices, wides, fvals = [], [], [], [], []
    for n in seg_sizes:
        c = rng.integers(1, n_cat + 1, size=n, dtype=np.uint64)
        p = np.round(rng.lognormal(2.0, 1.0, size=n), 3)
        w = (rng.integers(1, 6, size=n, dtype=np.uint64) << np.uint64(40)) | rng.integers(1, 9, size=n, dtype=np.uint64)
        fv = [list(np.round(rng.random(rng.integers(0, 4)) * 100, 2)) for _ in range(n)]
        s = SegSpec(n).col(CAT, F.U64, c).col(PRICE, F.F64, p).col(WIDE, F.U64, w).mcol(FVALS, F.F64, fv)
        if n > 10:
            s.deleted = rng.choice(n, size=n // 10, replace=False)
        segs.append(s)
        alive = np.ones(n, dtype=bool)
        if s.deleted is not None:
            alive[s.deleted] = False
        cats.append(c[alive]); prices.append(p[alive]); wides.append(w[alive]); fvals += [fv[i] for i in np.flatnonzero(alive)]
    return Corpus(segs), np.concatenate(cats), np.concatenate(prices), np.concatenate(wides), fvals


def check_bucket(p, vals, po=None):
    srt = np.sort(np.asarray(vals, dtype=np.float64))
    assert p.n == len(srt)
    for qq in QS:
        v = p.percentile(qq)
        if len(srt) == 0:
            assert v is None
            continue
        lo, hi = exact_rank_window(srt, v)
        assert lo <= hi, ("not an element of the bucket", qq, v)
        k = ta.ckms_target_rank(qq, len(srt))
        band = EPS * qq * len(srt) + 1
        assert lo - band <= k <= hi + band, (qq, k, lo, hi)
        if po is not None and len(srt) <= 40:  # CKMS(0.01) has not compressed yet: exact order statistics on both sides
            assert v == po.percentile(qq), (qq, v, po.percentile(qq))


@pytest.mark.parametrize("executor", ["single", "pool"])
def test_percentiles_under_terms_and_histogram(ctx, executor):
    corpus, cats, prices, wides, fvals = build(3, [4000, 0, 1500, 37, 9000], 40)
    searcher, ox = corpus.build_gpu(ctx), corpus.build_oracle()
    ex = ta.SINGLE_THREAD if executor == "single" else ta.THREAD_POOL
    # dense terms scope
    mk = lambda: ta.terms_agg_u64(CAT, (ta.count_agg(), ta.percentiles_agg_f64(PRICE)))
    got = searcher.agg_search_with_executor(ta.AllQuery(), mk(), ex)
    want, _, _ = ox.search(ta.AllQuery(), mk(), mode=0 if executor == "single" else 1, threads=3)
    assert set(got.res) == set(want.res) == set(int(c) for c in np.unique(cats))
    for c, (cnt, p) in got.res.items():
        assert cnt == int((cats == c).sum()) == want.res[c][0]
        check_bucket(p, prices[cats == c], want.res[c][1])
    # hashed terms scope (40-bit sparse keys) + root percentiles in the same tuple
    mk = lambda: (ta.percentiles_agg_f64(PRICE), ta.terms_agg_u64(WIDE, ta.percentiles_agg_f64(PRICE)))
    root, terms = searcher.agg_search_with_executor(ta.AllQuery(), mk(), ex)
    check_bucket(root, prices)
    assert set(terms.res) == set(int(w) for w in np.unique(wides))
    for w, p in terms.res.items():
        check_bucket(p, prices[wides == w])
    # histogram -> percentiles of a multi-valued field (every value of the document is inserted, percentile.rs:119-124)
    mk = lambda: ta.histogram_agg_f64(PRICE, 0.0, 5.0, ta.percentiles_agg_f64s(FVALS))
    hist = searcher.agg_search_with_executor(ta.AllQuery(), mk(), ex)
    ords = np.floor(prices / 5.0).astype(np.int64)
    for key, p in hist.buckets():
        if p is None:  # a gap bucket synthesised by buckets() (histogram.rs:163-181)
            continue
        o = int(round(key / 5.0))
        vals = [v for i in np.flatnonzero(ords == o) for v in fvals[i]]
        check_bucket(p, vals)


def test_small_buckets_equal_the_oracle(ctx):
    """Buckets of a few documents: the CKMS sketch is still exact, so every percentile must equal the oracle's."""
    corpus, cats, prices, _, _ = build(9, [300, 250], 60)
    searcher, ox = corpus.build_gpu(ctx), corpus.build_oracle()
    mk = lambda: ta.terms_agg_u64(CAT, ta.percentiles_agg_f64(PRICE))
    got = searcher.agg_search(ta.AllQuery(), mk())
    want, _, _ = ox.search(ta.AllQuery(), mk())
    assert set(got.res) == set(want.res)
    for c in got.res:
        check_bucket(got.res[c], prices[cats == c], want.res[c])
