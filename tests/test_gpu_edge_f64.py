"""GPU vs oracle on f64 edge values: NaN (first / middle / only, with payloads and sign), -0.0 / +0.0 in both orders,
+-inf, sums whose addends are all -0.0 — single- and multi-segment, both executors, at the root and under buckets,
single- and multi-valued columns, with deletes moving the first collected document.

The reference's `PartialOrd` fold (minmax.rs:97-106) lets a FIRST NaN stick and keeps the first-seen zero; sum.rs:95-102
replaces with the first value.  Results must be bit-identical (NaN payload and zero sign included), except f64 sums that
are NaN (any NaN) or order-dependent in the last place (1e-12, none of these sequences)."""
import math

import numpy as np
import pytest

import tantivy_aggregations_b200 as ta
from edge_cases import INF, NAN, NEG_NAN, PAYLOAD_NAN, SEQUENCES, bits, random_sequences, ref_search
from helpers import Corpus, SegSpec, assert_fruit_equal
from tantivy_aggregations_b200 import _ffi as F

pytestmark = pytest.mark.gpu

KEY, PRICE, FVALS = 1, 3, 5
ALL = dict(SEQUENCES)
ALL.update(random_sequences(7))
ALL.update(random_sequences(11, 8))
EXECUTORS = [ta.SINGLE_THREAD, ta.THREAD_POOL]


def corpus_of(segments, keys=None, deleted=None):
    segs = []
    for i, vals in enumerate(segments):
        s = SegSpec(len(vals))
        s.col(PRICE, F.F64, np.array(vals, dtype=np.float64))
        k = keys[i] if keys is not None else [1 + (j % 2) for j in range(len(vals))]
        s.col(KEY, F.U64, np.array(k, dtype=np.uint64))
        if deleted is not None and deleted[i]:
            s.deleted = deleted[i]
        segs.append(s)
    return Corpus(segs)


def check(ctx, corpus, query, make_agg, executor):
    """Bit-exact, NaN payloads and zero signs included (the planted values add up exactly in any order); only an f64 SUM
    that is NaN may differ in payload — callers keep sums out of the plans whose NaNs must match bit for bit."""
    ox = corpus.build_oracle()
    want, _, _ = ox.search(query, make_agg(), mode=0 if executor == ta.SINGLE_THREAD else 1, threads=2)
    got = corpus.build_gpu(ctx).agg_search_with_executor(query, make_agg(), executor)
    assert_fruit_equal(got, want, nan_equal="sum" in repr_ops(make_agg()))
    return got


def repr_ops(agg):
    from tantivy_aggregations_b200 import agg as A
    agg = A.as_agg(agg)
    if isinstance(agg, A.SumAgg):
        return ["sum"]
    out = []
    for child in getattr(agg, "members", []) + ([agg.sub] if hasattr(agg, "sub") else []):
        out += repr_ops(child)
    return out


@pytest.mark.parametrize("executor", EXECUTORS)
@pytest.mark.parametrize("name", sorted(ALL))
def test_root_min_max_sum(ctx, name, executor):
    segments = ALL[name]
    corpus = corpus_of(segments)
    got = check(ctx, corpus, ta.AllQuery(), lambda: (ta.count_agg(), ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE)), executor)
    check(ctx, corpus, ta.AllQuery(), lambda: (ta.count_agg(), ta.sum_agg_f64(PRICE), ta.max_agg_f64(PRICE)), executor)
    # and straight against the restated fold (the oracle is pinned to it in tests/test_oracle_edge.py)
    ex = "SingleThread" if executor == ta.SINGLE_THREAD else "ThreadPool"
    for op, g in zip(("min", "max"), got[1:3]):
        w = ref_search(op, segments, ex)
        assert (g is None) == (w is None) and (w is None or bits(g) == bits(w)), (name, op, g, w)


@pytest.mark.parametrize("executor", EXECUTORS)
@pytest.mark.parametrize("name", sorted(ALL))
def test_under_terms_buckets(ctx, name, executor):
    """Every bucket folds its own documents in doc order: the first NaN / first zero is per bucket."""
    corpus = corpus_of(ALL[name])
    check(ctx, corpus, ta.AllQuery(), lambda: ta.terms_agg_u64(KEY, (ta.count_agg(), ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE))), executor)
    check(ctx, corpus, ta.AllQuery(), lambda: ta.terms_agg_u64(KEY, (ta.sum_agg_f64(PRICE), ta.min_agg_f64(PRICE))), executor)


@pytest.mark.parametrize("executor", EXECUTORS)
def test_deletes_move_the_first_collected_document(ctx, executor):
    segments = [[NAN, 2.0, -0.0, 0.0, 1.0], [0.0, -0.0, NAN]]
    for deleted in ([[0], []], [[0, 2], [0]], [[1, 4], [2]], [[], [0, 1]]):
        corpus = corpus_of(segments, deleted=deleted)
        check(ctx, corpus, ta.AllQuery(), lambda: (ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE)), executor)
        check(ctx, corpus, ta.AllQuery(), lambda: ta.terms_agg_u64(KEY, (ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE))), executor)


@pytest.mark.parametrize("executor", EXECUTORS)
def test_docsets_and_filters_move_the_first_collected_document(ctx, executor):
    segments = [[NAN, 2.0, -0.0, 0.0, 1.0, NAN, -3.0], [0.0, -0.0, NAN, 4.0]]
    corpus = corpus_of(segments)
    bitsets = {0: np.packbits(np.array([0, 1, 1, 1, 0, 1, 0], dtype=np.uint8), bitorder="little"),
               1: np.packbits(np.array([0, 1, 1, 1], dtype=np.uint8), bitorder="little")}
    ids = {0: np.array([2, 3, 5], dtype=np.uint32), 1: np.array([1, 2], dtype=np.uint32)}
    mk = lambda: (ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE), ta.sum_agg_f64(PRICE))
    for q in (ta.BitsetQuery(bitsets), ta.DocIdsQuery(ids)):
        check(ctx, corpus, q, mk, executor)
        check(ctx, corpus, q, lambda: ta.filter_agg(ta.TermQuery(KEY, F.U64, 1), mk()), executor)
        check(ctx, corpus, q, lambda: ta.post_filter_agg_f64(PRICE, ta.le(2.0), mk()), executor)


@pytest.mark.parametrize("executor", EXECUTORS)
def test_multi_valued_f64s(ctx, executor):
    """min / max / sum_agg_f64s fold every value of the document in value order (minmax.rs:135-145, sum.rs:131-140)."""
    cases = [
        [[[NAN, 1.0], [2.0]]],
        [[[1.0, NAN], [0.5], []]],
        [[[], [-0.0, 0.0], [3.0]]],
        [[[0.0], [-0.0, -0.0]]],
        [[[-0.0], [-0.0, -0.0]]],
        [[[5.0, 0.0]], [[-0.0], [PAYLOAD_NAN]]],
        [[[]], [[NEG_NAN, -1.0], [INF]]],
    ]
    for segments in cases:
        segs = []
        for lists in segments:
            s = SegSpec(len(lists))
            s.mcol(FVALS, F.F64, lists)
            s.col(KEY, F.U64, np.array([1 + (j % 2) for j in range(len(lists))], dtype=np.uint64))
            segs.append(s)
        corpus = Corpus(segs)
        check(ctx, corpus, ta.AllQuery(), lambda: (ta.min_agg_f64s(FVALS), ta.max_agg_f64s(FVALS), ta.sum_agg_f64s(FVALS)), executor)
        check(ctx, corpus, ta.AllQuery(),
              lambda: ta.terms_agg_u64(KEY, (ta.min_agg_f64s(FVALS), ta.max_agg_f64s(FVALS), ta.sum_agg_f64s(FVALS))), executor)


def test_sum_of_negative_zeros_keeps_its_sign_on_every_kernel(ctx):
    """sum.rs:95-102: the first value replaces, so a sum whose addends are all -0.0 is -0.0 — on the streaming kernels
    (root registers, CTA-private tables, global tables), on k_mterms and on the generic kernel."""
    n = 5000
    for vals, want_bits in ((np.full(n, -0.0), bits(-0.0)), (np.concatenate([np.full(n - 1, -0.0), [0.0]]), bits(0.0))):
        s = SegSpec(n)
        s.col(PRICE, F.F64, vals)
        s.col(KEY, F.U64, np.arange(n, dtype=np.uint64) % 7)
        s.col(2, F.U64, np.arange(n, dtype=np.uint64) * np.uint64(1 << 30))  # sparse keys: hashed scope
        s.mcol(FVALS, F.F64, [[v] * (j % 3) for j, v in enumerate(vals)])
        s.mcol(4, F.U64, [[j % 5, (j + 1) % 5][: j % 3] for j in range(n)])
        corpus = Corpus([s])
        ox = corpus.build_oracle()
        searcher = corpus.build_gpu(ctx)
        plans = [lambda: (ta.count_agg(), ta.sum_agg_f64(PRICE)),
                 lambda: ta.terms_agg_u64(KEY, (ta.count_agg(), ta.sum_agg_f64(PRICE))),
                 lambda: ta.terms_agg_u64(2, ta.sum_agg_f64(PRICE)),
                 lambda: ta.terms_agg_u64s(4, ta.sum_agg_f64s(FVALS)),
                 lambda: ta.histogram_agg_f64(PRICE, -1.0, 1.0, ta.sum_agg_f64(PRICE))]
        for path in (F.PATH_AUTO, F.PATH_GENERIC):
            ctx.set_path(path)
            try:
                for mk in plans:
                    want, _, _ = ox.search(ta.AllQuery(), mk())
                    got = searcher.agg_search(ta.AllQuery(), mk())
                    assert_fruit_equal(got, want)
            finally:
                ctx.set_path(F.PATH_AUTO)
        got = searcher.agg_search(ta.AllQuery(), (ta.count_agg(), ta.sum_agg_f64(PRICE)))
        assert bits(got[1]) == want_bits


def test_large_signed_column_stays_on_the_streaming_path(ctx):
    """A column that merely SPANS zero keeps the fast kernels: only an actual ambiguous zero result takes the exact path."""
    rng = np.random.default_rng(3)
    n = 300_000
    s = SegSpec(n)
    s.col(PRICE, F.F64, rng.normal(0.0, 10.0, size=n))
    s.col(KEY, F.U64, rng.integers(1, 100, size=n, dtype=np.uint64))
    corpus = Corpus([s])
    ox = corpus.build_oracle()
    searcher = corpus.build_gpu(ctx)
    mk = lambda: (ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE), ta.terms_agg_u64(KEY, (ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE))))
    want, _, _ = ox.search(ta.AllQuery(), mk())
    got, reader = searcher.agg_search_with_executor(ta.AllQuery(), mk(), ta.SINGLE_THREAD, return_reader=True)
    assert_fruit_equal(got, want)
    assert reader.stats()["path"] == 2  # stream
    # plant both zeros into one bucket: that query (and only its ambiguity) moves to the exact path, results stay exact
    vals = rng.normal(0.0, 10.0, size=n)
    vals[vals < 0] *= -1.0
    vals[1000], vals[2000] = 0.0, -0.0
    s2 = SegSpec(n)
    s2.col(PRICE, F.F64, vals)
    keys = rng.integers(1, 100, size=n, dtype=np.uint64)
    keys[1000] = keys[2000] = 5
    s2.col(KEY, F.U64, keys)
    corpus2 = Corpus([s2])
    want, _, _ = corpus2.build_oracle().search(ta.AllQuery(), mk())
    got, reader = corpus2.build_gpu(ctx).agg_search_with_executor(ta.AllQuery(), mk(), ta.SINGLE_THREAD, return_reader=True)
    assert_fruit_equal(got, want)
    assert bits(got[0]) == bits(0.0) and bits(got[2].get(5)[0]) == bits(0.0)  # +0.0 came first
    assert reader.stats()["path"] == 1  # generic (exact)


def test_forced_stream_path_refuses_nan_columns(ctx):
    s = SegSpec(3)
    s.col(PRICE, F.F64, np.array([1.0, NAN, 2.0]))
    searcher = Corpus([s]).build_gpu(ctx)
    ctx.set_path(F.PATH_STREAM)
    try:
        with pytest.raises(F.TaggError):
            searcher.agg_search(ta.AllQuery(), ta.min_agg_f64(PRICE))
        # a sum / count plan on the same column has no ordering question: it still streams
        assert searcher.agg_search(ta.AllQuery(), ta.count_agg()) == 3
    finally:
        ctx.set_path(F.PATH_AUTO)
