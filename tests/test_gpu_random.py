"""GPU vs oracle on seeded random corpora: multi-segment, deletes, every docset kind, ragged and empty
inputs, wide/sparse keys (hash scopes), multi-valued fields, nested buckets, i64/date/f64s types the
reference never tests (SURVEY §8c "parity-unpinned")."""
import numpy as np
import pytest

import tantivy_aggregations_b200 as ta
from helpers import Corpus, SegSpec, assert_fruit_equal, exact_rank_window
from tantivy_aggregations_b200 import _ffi as F

pytestmark = pytest.mark.gpu

# fields
CAT, PRICE, STATUS, WIDE, SIGNED, WHEN, TAGS, FVALS, ITAGS, CONST = range(10)
F64_SUM_RTOL = 1e-12  # north_star tolerance for f64 sums (order differs from the doc-order fold)


def make_corpus(seed, seg_sizes, n_cat=50, deletes=True):
    rng = np.random.default_rng(seed)
    segs = []
    for n in seg_sizes:
        s = SegSpec(n)
        s.col(CAT, F.U64, rng.integers(1, n_cat + 1, size=n, dtype=np.uint64))
        s.col(PRICE, F.F64, 1.0 + 100.0 * rng.random(n))
        s.col(STATUS, F.U64, rng.integers(0, 4, size=n, dtype=np.uint64))
        # 37-bit sparse keys like the reference bench's attr_facets (benches/lib.rs:79-83)
        s.col(WIDE, F.U64, (rng.integers(1, 20, size=n, dtype=np.uint64) << np.uint64(32)) | rng.integers(1, 100, size=n, dtype=np.uint64))
        s.col(SIGNED, F.I64, rng.integers(-1000, 1000, size=n, dtype=np.int64))
        s.col(WHEN, F.DATE, rng.integers(1_500_000_000, 1_600_000_000, size=n, dtype=np.int64))
        s.col(CONST, F.U64, np.full(n, 7, dtype=np.uint64))
        s.mcol(TAGS, F.U64, [list(rng.integers(100, 130, size=rng.integers(0, 5))) for _ in range(n)])
        s.mcol(FVALS, F.F64, [list(np.round(rng.random(rng.integers(0, 4)) * 10, 3)) for _ in range(n)])
        s.mcol(ITAGS, F.I64, [list(rng.integers(-5, 6, size=rng.integers(0, 3))) for _ in range(n)])
        if deletes and n:
            s.deleted = rng.choice(n, size=n // 5, replace=False)
        segs.append(s)
    return Corpus(segs)


@pytest.fixture(scope="module")
def world(ctx):
    corpus = make_corpus(1, [7001, 0, 2048, 33, 12345])
    return corpus, corpus.build_gpu(ctx), corpus.build_oracle()


def queries(corpus, seed):
    rng = np.random.default_rng(seed)
    bits, ids = {}, {}
    for i, s in enumerate(corpus.segs):
        m = rng.random(s.max_doc) < 0.5
        bits[i] = np.packbits(m.astype(np.uint8), bitorder="little")
        ids[i] = np.flatnonzero(rng.random(s.max_doc) < 0.03).astype(np.uint32)
    return {
        "all": ta.AllQuery(),
        "bitset": ta.BitsetQuery(bits),
        "ids": ta.DocIdsQuery(ids),
        "range_dev": ta.RangeQuery(STATUS, F.U64, 1, 2, device=True),
        "range_host": ta.RangeQuery(STATUS, F.U64, 1, 2, device=False),
        "none": ta.DocIdsQuery({i: np.zeros(0, np.uint32) for i in range(len(corpus.segs))}),
    }


def aggs():
    return {
        "scalars": lambda: (ta.count_agg(), ta.sum_agg_u64(CAT), ta.sum_agg_i64(SIGNED), ta.sum_agg_f64(PRICE),
                            ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE), ta.min_agg_i64(SIGNED), ta.max_agg_i64(SIGNED),
                            ta.min_agg_date(WHEN), ta.max_agg_date(WHEN)),
        "scalars_multi": lambda: (ta.sum_agg_u64s(TAGS), ta.min_agg_u64s(TAGS), ta.max_agg_u64s(TAGS), ta.sum_agg_f64s(FVALS),
                                  ta.min_agg_f64s(FVALS), ta.max_agg_f64s(FVALS), ta.sum_agg_i64s(ITAGS), ta.min_agg_i64s(ITAGS)),
        "terms": lambda: ta.terms_agg_u64(CAT, (ta.count_agg(), ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE), ta.sum_agg_f64(PRICE))),
        "terms_i64": lambda: ta.terms_agg_i64(SIGNED, (ta.count_agg(), ta.sum_agg_i64(SIGNED))),
        "terms_wide_hash": lambda: ta.terms_agg_u64(WIDE, (ta.count_agg(), ta.max_agg_u64(CAT))),
        "terms_const": lambda: ta.terms_agg_u64(CONST, ta.count_agg()),
        "terms_multi": lambda: ta.terms_agg_u64s(TAGS, (ta.count_agg(), ta.sum_agg_f64s(FVALS), ta.min_agg_f64(PRICE))),
        "terms_multi_i64": lambda: ta.terms_agg_i64s(ITAGS, ta.count_agg()),
        "bench_shape": lambda: ta.filter_agg(ta.TermQuery(STATUS, F.U64, 0),
                                             (ta.count_agg(), ta.terms_agg_u64(CAT, (ta.count_agg(), ta.min_agg_f64(PRICE))))),
        "hist": lambda: ta.histogram_agg_f64(PRICE, 0.0, 10.0, (ta.count_agg(), ta.sum_agg_f64(PRICE))),
        "hist_fine": lambda: ta.histogram_agg_f64(PRICE, 17.3, 0.37, ta.count_agg()),
        "nested": lambda: ta.terms_agg_u64(CAT, (ta.count_agg(), ta.histogram_agg_f64(PRICE, 0.0, 25.0, (ta.count_agg(), ta.terms_agg_u64(STATUS, ta.count_agg()))))),
        "nested_under_hash": lambda: ta.terms_agg_u64(WIDE, ta.terms_agg_u64(STATUS, (ta.count_agg(), ta.min_agg_i64(SIGNED)))),
        "post_filter": lambda: ta.post_filter_agg_u64(STATUS, ta.eq(0), ta.terms_agg_u64(CAT, (ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE), ta.sum_agg_f64(PRICE)))),
        "post_filter_f64": lambda: ta.post_filter_agg_f64(PRICE, ta.gt(50.0), (ta.count_agg(), ta.min_agg_f64(PRICE))),
        "post_filter_lut": lambda: ta.post_filter_agg_u64(CAT, lambda c: c % 3 == 1, ta.count_agg()),
        "post_filter_multi": lambda: ta.post_filter_agg_u64s(TAGS, ta.in_set({101, 117, 129}), (ta.count_agg(), ta.sum_agg_u64s(TAGS))),
        "post_filter_i64s": lambda: ta.post_filter_agg_i64s(ITAGS, ta.lt(0), ta.count_agg()),
        "filter_in_bucket": lambda: ta.terms_agg_u64(CAT, ta.filter_agg(ta.RangeQuery(PRICE, F.F64, 20.0, 60.0), (ta.count_agg(), ta.sum_agg_u64(STATUS)))),
        "two_filters": lambda: (ta.filter_agg(ta.TermQuery(STATUS, F.U64, 1), ta.count_agg()),
                                ta.filter_agg(ta.RangeQuery(PRICE, F.F64, 0.0, 50.0, device=False), ta.max_agg_f64(PRICE))),
    }


@pytest.mark.parametrize("qname", ["all", "bitset", "ids", "range_dev", "range_host", "none"])
@pytest.mark.parametrize("aname", sorted(aggs()))
def test_gpu_matches_oracle(ctx, world, qname, aname):
    corpus, searcher, ox = world
    q = queries(corpus, 5)[qname]
    want, _, _ = ox.search(q, aggs()[aname]())
    got = searcher.agg_search(q, aggs()[aname]())
    assert_fruit_equal(got, want, F64_SUM_RTOL)


@pytest.mark.parametrize("aname", ["scalars", "terms", "terms_wide_hash", "hist", "nested", "terms_multi", "bench_shape"])
def test_thread_pool_merge_matches_oracle(ctx, world, aname):
    """Executor::ThreadPool shape: a fruit per segment merged in segment order (searcher.rs:79-98)."""
    corpus, searcher, ox = world
    q = queries(corpus, 9)["bitset"]
    want, _, _ = ox.search(q, aggs()[aname](), mode=1, threads=3)
    got = searcher.agg_search_with_executor(q, aggs()[aname](), ta.THREAD_POOL)
    assert_fruit_equal(got, want, F64_SUM_RTOL)


def test_hash_table_growth(ctx):
    """More distinct keys than the first table holds: the overflow -> grow -> redo path."""
    rng = np.random.default_rng(3)
    n = 300_000
    s = SegSpec(n)
    keys = rng.integers(0, 1 << 40, size=n, dtype=np.uint64)
    s.col(WIDE, F.U64, keys)
    s.col(PRICE, F.F64, rng.random(n))
    corpus = Corpus([s])
    searcher, ox = corpus.build_gpu(ctx), corpus.build_oracle()
    ids = np.sort(rng.choice(n, size=600, replace=False)).astype(np.uint32)  # tiny candidate bound -> tiny first table
    for q in (ta.DocIdsQuery({0: ids}), ta.AllQuery()):
        agg = lambda: ta.terms_agg_u64(WIDE, (ta.count_agg(), ta.min_agg_f64(PRICE)))
        want, _, _ = ox.search(q, agg())
        got = searcher.agg_search(q, agg())
        assert_fruit_equal(got, want)


def test_percentiles_rank_tolerance(ctx):
    """Large n: the answer's rank lies within CKMS's +-eps*q*n band (here: far tighter), and the
    oracle's CKMS restatement — the tolerance witness — satisfies the same band."""
    rng = np.random.default_rng(11)
    n = 200_000
    vals = np.round(rng.lognormal(3.0, 1.0, size=n), 4)
    s = SegSpec(n).col(PRICE, F.F64, vals)
    corpus = Corpus([s])
    searcher, ox = corpus.build_gpu(ctx), corpus.build_oracle()
    m = rng.random(n) < 0.5
    q = ta.BitsetQuery({0: np.packbits(m.astype(np.uint8), bitorder="little")})
    p = searcher.agg_search(q, ta.percentiles_agg_f64(PRICE))
    po, _, _ = ox.search(q, ta.percentiles_agg_f64(PRICE))
    srt = np.sort(vals[m])
    cnt = len(srt)
    assert p.n == cnt == po.n
    eps = 0.01
    for qq in (0.001, 0.01, 0.25, 0.5, 0.75, 0.95, 0.99, 0.999):
        v = p.percentile(qq)
        lo, hi = exact_rank_window(srt, v)
        assert lo <= hi, "the answer must be an element of the input"
        k = ta.ckms_target_rank(qq, cnt)
        band = eps * qq * cnt + 1
        assert lo - band <= k <= hi + band, (qq, k, lo, hi)
        vo = po.percentile(qq)
        lo2, hi2 = exact_rank_window(srt, vo)
        assert lo2 - 2 * band - 1 <= qq * cnt <= hi2 + 2 * band + 1, ("oracle CKMS out of its own band", qq)


def test_percentiles_small_n_exact(ctx):
    """While CKMS is uncompressed its answer is an exact order statistic: equality with the oracle."""
    rng = np.random.default_rng(12)
    for n in (1, 2, 5, 17, 49):
        vals = np.round(rng.random(n) * 100, 2)
        corpus = Corpus([SegSpec(n).col(PRICE, F.F64, vals)])
        searcher, ox = corpus.build_gpu(ctx), corpus.build_oracle()
        p = searcher.agg_search(ta.AllQuery(), ta.percentiles_agg_f64(PRICE))
        po, _, _ = ox.search(ta.AllQuery(), ta.percentiles_agg_f64(PRICE))
        for qq in (0.0, 0.01, 0.1, 0.33, 0.5, 0.7, 0.9, 0.99, 1.0):
            assert p.percentile(qq) == po.percentile(qq), (n, qq)
    # multi-valued f64s (percentiles_agg_f64s, never tested by the reference)
    corpus = Corpus([SegSpec(4).mcol(FVALS, F.F64, [[1.5, 2.5], [], [0.5], [9.0, 3.0, 4.0]])])
    p = corpus.build_gpu(ctx).agg_search(ta.AllQuery(), ta.percentiles_agg_f64s(FVALS))
    po, _, _ = corpus.build_oracle().search(ta.AllQuery(), ta.percentiles_agg_f64s(FVALS))
    for qq in (0.01, 0.5, 0.99):
        assert p.percentile(qq) == po.percentile(qq)
    # empty -> None
    corpus = Corpus([SegSpec(0).col(PRICE, F.F64, np.zeros(0))])
    assert corpus.build_gpu(ctx).agg_search(ta.AllQuery(), ta.percentiles_agg_f64(PRICE)).percentile(0.5) is None


STREAM_SHAPES = ["scalars", "terms", "terms_i64", "terms_const", "bench_shape", "hist", "hist_fine", "post_filter",
                 "post_filter_f64", "post_filter_lut"]


@pytest.mark.parametrize("qname", ["all", "bitset", "range_dev"])
@pytest.mark.parametrize("aname", STREAM_SHAPES)
def test_streaming_kernel_is_used_and_matches_generic(ctx, world, qname, aname):
    """The flat shapes run on the TMA-staged streaming kernel (path 2); forcing the generic
    tree-walking kernel (path 1) must give the same fruit, and both equal the oracle."""
    corpus, searcher, ox = world
    q = queries(corpus, 5)[qname]
    want, _, _ = ox.search(q, aggs()[aname]())
    ctx.set_path(F.PATH_STREAM)
    try:
        got, reader = searcher.agg_search_with_executor(q, aggs()[aname](), ta.SINGLE_THREAD, return_reader=True)
        assert reader.stats()["path"] == 2
    finally:
        ctx.set_path(F.PATH_AUTO)
    assert_fruit_equal(got, want, F64_SUM_RTOL)


def test_cached_device_docsets(ctx, world):
    """Reusable filters kept resident in HBM (tagg_docset_cache) and docsets produced on the device
    (tagg_docset_to_bitset) agree with the host-evaluated ones."""
    corpus, searcher, ox = world
    host_q = ta.RangeQuery(STATUS, F.U64, 0, 0, device=False)
    for seg in searcher.segments:
        dev_bits = seg.docset_to_bitset(ta.RangeQuery(STATUS, F.U64, 0, 0).docset(seg))
        assert (dev_bits == host_q.docset(seg).buf).all()
    cached = ta.CachedQuery(host_q, searcher.segments)
    agg = lambda fq: ta.filter_agg(fq, (ta.count_agg(), ta.terms_agg_u64(CAT, (ta.count_agg(), ta.min_agg_f64(PRICE)))))
    want, _, _ = ox.search(ta.AllQuery(), agg(host_q))
    assert_fruit_equal(searcher.agg_search(ta.AllQuery(), agg(cached)), want)
    assert_fruit_equal(searcher.agg_search(cached, ta.count_agg()), ox.search(host_q, ta.count_agg())[0])


def test_filter_tables_and_level_filters_on_integer_columns(ctx):
    """Bucket min / max behind the shared-memory filters (u32 rank words when the exact tables would crowd out a consumer
    group, 4-bit levels when the tables live in global memory) on integer and date columns, several segments with different
    column ranges."""
    rng = np.random.default_rng(17)
    segs = []
    for i, n in enumerate((150_000, 90_001, 4096)):
        s = SegSpec(n)
        s.col(CAT, F.U64, rng.integers(1, 12_001, size=n, dtype=np.uint64))
        s.col(WIDE, F.U64, rng.integers(1, 200_001, size=n, dtype=np.uint64))
        s.col(STATUS, F.U64, rng.integers(0, 1000 * (i + 1), size=n, dtype=np.uint64))
        s.col(SIGNED, F.I64, rng.integers(-10_000 * (i + 1), 10_000, size=n, dtype=np.int64))
        s.col(WHEN, F.DATE, rng.integers(1_500_000_000, 1_500_000_000 + 10**6 * (i + 1), size=n, dtype=np.int64))
        s.col(PRICE, F.F64, rng.normal(0.0, 100.0, size=n))
        s.deleted = rng.choice(n, size=n // 20, replace=False)
        segs.append(s)
    corpus = Corpus(segs)
    searcher, ox = corpus.build_gpu(ctx), corpus.build_oracle()
    shapes = {
        "u32_filter": lambda: ta.terms_agg_u64(CAT, (ta.count_agg(), ta.min_agg_u64(STATUS), ta.max_agg_u64(STATUS), ta.min_agg_i64(SIGNED), ta.max_agg_date(WHEN))),
        "u32_filter_f64_signed": lambda: ta.terms_agg_u64(CAT, (ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE), ta.sum_agg_i64(SIGNED))),
        "level_filter": lambda: ta.terms_agg_u64(WIDE, (ta.min_agg_i64(SIGNED), ta.max_agg_i64(SIGNED), ta.sum_agg_i64(SIGNED))),
        "level_filter_f64": lambda: ta.post_filter_agg_u64(STATUS, ta.lt(500), ta.terms_agg_u64(WIDE, (ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE)))),
    }
    for name, mk in shapes.items():
        want, _, _ = ox.search(ta.AllQuery(), mk())
        ctx.set_path(F.PATH_STREAM)
        try:
            got = searcher.agg_search(ta.AllQuery(), mk())
        finally:
            ctx.set_path(F.PATH_AUTO)
        assert_fruit_equal(got, want, F64_SUM_RTOL, name)


def test_page_locked_docsets_are_read_in_place(ctx, world):
    """Page-locked, 16-byte aligned host bitsets (main docset and filter_agg docsets) are pulled over PCIe by the streaming
    kernel's TMA producer without a staging copy; the result must equal the staged-copy path and the oracle, including
    segments whose bitset ends inside its last tile."""
    import torch
    corpus, searcher, ox = world
    rng = np.random.default_rng(23)
    bits, pinned_bits, keep = {}, {}, []
    for i, s in enumerate(corpus.segs):
        m = rng.random(s.max_doc) < 0.4
        b = np.packbits(m.astype(np.uint8), bitorder="little")
        bits[i] = b
        t = torch.zeros((len(b) + 15) // 16 * 16 + 16, dtype=torch.uint8).pin_memory()
        t.numpy()[:len(b)] = b
        keep.append(t)
        pinned_bits[i] = t.numpy()
    for aname in ("scalars", "bench_shape", "terms", "hist", "post_filter", "terms_multi", "nested"):
        mk = aggs()[aname]
        want, _, _ = ox.search(ta.BitsetQuery(bits), mk())
        got = searcher.agg_search(ta.BitsetQuery(pinned_bits), mk())
        assert_fruit_equal(got, want, F64_SUM_RTOL, aname)
    # as a filter_agg docset
    mk = lambda q: ta.filter_agg(q, (ta.count_agg(), ta.terms_agg_u64(CAT, (ta.count_agg(), ta.min_agg_f64(PRICE)))))
    want, _, _ = ox.search(ta.AllQuery(), mk(ta.BitsetQuery(bits)))
    got = searcher.agg_search(ta.AllQuery(), mk(ta.BitsetQuery(pinned_bits)))
    assert_fruit_equal(got, want, F64_SUM_RTOL)


@pytest.mark.parametrize("start,interval", [(-37.5, 3.7), (0.1, 0.3), (1e-3, 12.5)])
def test_histogram_boundary_table_is_bit_exact(ctx, start, interval):
    """Large inputs take the boundary-table histogram (multiply + exact code boundaries instead of the IEEE division of
    histogram.rs:146): the bucket of every value must equal floor((v - start) / interval) computed in IEEE doubles,
    including values a few ulps around every bucket edge, values below start (skipped) and -0.0."""
    rng = np.random.default_rng(31)
    n = 4_400_000  # > 2048 tiles: the table path
    vals = rng.normal(0.0, 20.0, size=n) if start < 0 else np.abs(rng.normal(0.0, 20.0, size=n))
    # plant the edges themselves and their neighbours
    edges = start + interval * np.arange(0, 60, dtype=np.float64)
    planted = np.concatenate([np.nextafter(edges, -np.inf), edges, np.nextafter(edges, np.inf), [-0.0, 0.0, start]])
    vals[:len(planted)] = planted
    seg = SegSpec(n).col(PRICE, F.F64, vals)
    searcher = Corpus([seg]).build_gpu(ctx)
    hist, cnt = searcher.agg_search(ta.AllQuery(), (ta.histogram_agg_f64(PRICE, start, interval, ta.count_agg()), ta.count_agg()))
    assert cnt == n
    d = vals - start
    ok = ~(d < 0.0)
    ords, counts = np.unique(np.floor(d[ok] / interval).astype(np.int64), return_counts=True)
    assert dict(hist._buckets) == {int(o): int(c) for o, c in zip(ords, counts)}
    # and the generic (division) path agrees
    ctx.set_path(F.PATH_GENERIC)
    try:
        hist2 = searcher.agg_search(ta.AllQuery(), ta.histogram_agg_f64(PRICE, start, interval, ta.count_agg()))
    finally:
        ctx.set_path(F.PATH_AUTO)
    assert dict(hist2._buckets) == dict(hist._buckets)


@pytest.mark.parametrize("qname", ["all", "bitset"])
def test_bit_plane_predicates_on_one_and_two_bit_columns(ctx, qname):
    """post_filter / column-range predicates on columns of <= 4 values run on the bit planes of the packed stream (32
    documents per lane, a truth table per tile): every table over the column's values — equality, ranges that clip the
    domain, empty and full sets, arbitrary closures — on 1- and 2-bit columns with a non-zero min_value, segments whose
    widths differ (0, 1 and 2 bits), ragged last tiles and deletes; streaming kernel == oracle."""
    rng = np.random.default_rng(23)
    BIT, QUAD = 20, 21
    segs = []
    for n, quad_vals, bit_vals in ((9001, (5, 6, 7, 8), (3, 4)), (4096, (5, 6), (4, 4)), (2049, (7, 7), (3, 4)), (50_000, (5, 8), (3, 4))):
        s = SegSpec(n)
        s.col(QUAD, F.U64, rng.choice(np.array(quad_vals, dtype=np.uint64), size=n))
        s.col(BIT, F.U64, rng.choice(np.array(bit_vals, dtype=np.uint64), size=n))
        s.col(PRICE, F.F64, 1.0 + 100.0 * rng.random(n))
        s.col(CAT, F.U64, rng.integers(1, 40, size=n, dtype=np.uint64))
        s.deleted = rng.choice(n, size=n // 7, replace=False)
        segs.append(s)
    corpus = Corpus(segs)
    searcher, ox = corpus.build_gpu(ctx), corpus.build_oracle()
    q = queries(corpus, 9)[qname]
    sub = lambda: (ta.count_agg(), ta.terms_agg_u64(CAT, (ta.count_agg(), ta.min_agg_f64(PRICE))))
    preds = [ta.eq(v) for v in (3, 4, 5, 6, 7, 8, 9)] + [ta.ge(6), ta.le(6), ta.lt(5), ta.gt(8), ta.ge(0), ta.in_set({5, 8}), ta.in_set({4}),
                                                           (lambda c: c % 2 == 1), (lambda c: c != 7), (lambda c: False), (lambda c: True)]
    ctx.set_path(F.PATH_STREAM)
    try:
        for field in (QUAD, BIT):
            for pr in preds:
                agg = lambda: ta.post_filter_agg_u64(field, pr, sub())
                got, reader = searcher.agg_search_with_executor(q, agg(), ta.SINGLE_THREAD, return_reader=True)
                assert reader.stats()["path"] == 2
                assert_fruit_equal(got, ox.search(q, agg())[0], F64_SUM_RTOL)
        # nested: both columns narrow the stream; and a device column-range docset as the main query
        agg = lambda: ta.post_filter_agg_u64(QUAD, ta.ge(6), ta.post_filter_agg_u64(BIT, ta.eq(4), sub()))
        assert_fruit_equal(searcher.agg_search(q, agg()), ox.search(q, agg())[0], F64_SUM_RTOL)
        rq = ta.RangeQuery(QUAD, F.U64, 6, 7, device=True)
        assert_fruit_equal(searcher.agg_search(rq, sub()), ox.search(rq, sub())[0], F64_SUM_RTOL)
    finally:
        ctx.set_path(F.PATH_AUTO)
