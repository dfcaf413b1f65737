"""GPU: percentiles_agg_f64 on the streaming path (csrc/pct.cu — sampled thresholds, rank bins, exact tails).
Every answer must be an element of the input whose rank lies within CKMS's own +-eps*q*n band around the rank
the sketch targets (percentile.rs:174, eps = 0.01); distributions the bins cannot resolve must fall back to the
exact path and still satisfy the bound."""
import zlib

import numpy as np
import pytest

import tantivy_aggregations_b200 as ta
from helpers import Corpus, SegSpec, exact_rank_window
from tantivy_aggregations_b200 import _ffi as F

pytestmark = pytest.mark.gpu
PRICE, STATUS = 0, 1
N = 9_000_000  # above the 8M-document threshold of the rank-bin mode
EPS = 0.01
QS = (0.0, 1e-6, 1e-4, 0.001, 0.01, 0.1, 0.25, 0.5, 0.75, 0.9, 0.99, 0.999, 0.999999, 1.0)


def distributions(rng):
    return {
        "uniform_price": 1.0 + 100.0 * rng.random(N),
        "lognormal": np.round(rng.lognormal(3.0, 1.0, size=N), 4),
        "normal_signed": rng.normal(0.0, 50.0, size=N),
        "narrow_normal": rng.normal(1000.0, 0.5, size=N),
        "discrete_ties": rng.choice(np.array([0.99, 4.99, 9.99, 19.99, 49.5, 99.0]), size=N),
        "bimodal_far": np.where(rng.random(N) < 0.5, rng.normal(1.0, 0.01, size=N), rng.normal(1e6, 1.0, size=N)),
        "heavy_outliers": np.concatenate([rng.random(N - 5) * 10.0, np.array([1e12, 1e15, -1e9, 1e300, -1e300])]),
    }


def check(p, vals_sel):
    srt = np.sort(vals_sel)
    cnt = len(srt)
    assert p.n == cnt
    assert p.ranks == sorted(p.ranks) and len(set(p.ranks)) == len(p.ranks)
    for qq in QS:
        v = p.percentile(qq)
        lo, hi = exact_rank_window(srt, v)
        assert lo <= hi, ("the answer must be an element of the input", qq, v)
        k = ta.ckms_target_rank(qq, cnt)
        band = EPS * qq * cnt + 1
        assert lo - band <= k <= hi + band, (qq, k, lo, hi)
    # the stored pairs are exact order statistics
    idx = np.linspace(0, len(p.ranks) - 1, 200).astype(int)
    for i in idx:
        lo, hi = exact_rank_window(srt, p.values[i])
        assert lo <= p.ranks[i] <= hi, (i, p.ranks[i], lo, hi)


@pytest.mark.parametrize("name", ["uniform_price", "lognormal", "normal_signed", "narrow_normal", "discrete_ties", "bimodal_far", "heavy_outliers"])
def test_rank_bin_percentiles(ctx, name):
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    vals = distributions(rng)[name]
    half = N // 2
    segs = [SegSpec(half).col(PRICE, F.F64, vals[:half]), SegSpec(N - half).col(PRICE, F.F64, vals[half:])]
    searcher = Corpus(segs).build_gpu(ctx)
    m = rng.random(N) < 0.5
    q = ta.BitsetQuery({0: np.packbits(m[:half].astype(np.uint8), bitorder="little"), 1: np.packbits(m[half:].astype(np.uint8), bitorder="little")})
    p, reader = searcher.agg_search_with_executor(q, ta.percentiles_agg_f64(PRICE), ta.SINGLE_THREAD, return_reader=True)
    check(p, vals[m])
    if name in ("uniform_price", "lognormal", "discrete_ties", "normal_signed", "narrow_normal"):
        assert reader.stats()["path"] == 2, "these distributions must stay on the one-pass streaming path"
    # together with a histogram and root metrics, AllQuery
    (cnt, hist, p2) = searcher.agg_search(ta.AllQuery(), (ta.count_agg(), ta.histogram_agg_f64(PRICE, 0.0, 10.0, ta.count_agg()), ta.percentiles_agg_f64(PRICE)))
    assert cnt == N
    check(p2, vals)
    if name in ("uniform_price", "lognormal", "narrow_normal", "discrete_ties", "normal_signed"):
        # the histogram (fused into the percentile pass where both stream) against numpy: histogram.rs:136-152
        ok = vals - 0.0 >= 0.0
        ords, counts = np.unique(np.floor(vals[ok] / 10.0).astype(np.int64), return_counts=True)
        assert dict(hist._buckets) == {int(o): int(c) for o, c in zip(ords, counts)}


def test_rank_bin_percentiles_under_filters(ctx):
    """The sample must see the same doc stream as the main pass: filter_agg + post_filter in front of the leaf."""
    rng = np.random.default_rng(77)
    vals = 1.0 + 100.0 * rng.random(N)
    status = rng.integers(0, 4, size=N, dtype=np.uint64)
    seg = SegSpec(N).col(PRICE, F.F64, vals).col(STATUS, F.U64, status)
    seg.deleted = rng.choice(N, size=N // 50, replace=False)
    searcher = Corpus([seg]).build_gpu(ctx)
    agg = ta.filter_agg(ta.RangeQuery(STATUS, F.U64, 0, 2), ta.post_filter_agg_f64(PRICE, ta.gt(20.0), (ta.count_agg(), ta.percentiles_agg_f64(PRICE))))
    cnt, p = searcher.agg_search(ta.AllQuery(), agg)
    alive = np.ones(N, dtype=bool)
    alive[seg.deleted] = False
    sel = alive & (status <= 2) & (vals > 20.0)
    assert cnt == int(sel.sum())
    check(p, vals[sel])


@pytest.mark.parametrize("name", ["uniform_price", "lognormal", "heavy_outliers"])
def test_repeated_queries_summarise_the_exact_lists_on_the_device(ctx, name):
    """A prepared plan remembers its thresholds and the length of its exact lists: from the second query on the lists are
    sorted and thinned on the device behind the pass (pct.cu k_tail_pick) — the summary must be the very same pairs the
    host path of the first query produced.  A docset that doubles the lists outruns the prediction and falls back."""
    rng = np.random.default_rng(zlib.crc32(name.encode()) + 1)
    vals = distributions(rng)[name]
    half = N // 2
    segs = [SegSpec(half).col(PRICE, F.F64, vals[:half]), SegSpec(N - half).col(PRICE, F.F64, vals[half:])]
    searcher = Corpus(segs).build_gpu(ctx)
    m = rng.random(N) < 0.4
    q = ta.BitsetQuery({0: np.packbits(m[:half].astype(np.uint8), bitorder="little"), 1: np.packbits(m[half:].astype(np.uint8), bitorder="little")})
    plan = searcher.prepare(ta.percentiles_agg_f64(PRICE))
    first = searcher.agg_search(q, plan)
    check(first, vals[m])
    for _ in range(3):
        again = searcher.agg_search(q, plan)
        assert again.n == first.n and again.ranks == first.ranks and again.values == first.values
    everything = searcher.agg_search(ta.AllQuery(), plan)   # 2.5x the documents: longer lists than predicted
    check(everything, vals)
    again = searcher.agg_search(ta.AllQuery(), plan)
    assert again.ranks == everything.ranks and again.values == everything.values
    back = searcher.agg_search(q, plan)
    assert back.ranks == first.ranks and back.values == first.values
