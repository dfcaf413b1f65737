"""GPU: the K5 kernel (csrc/mterms.cu) — terms buckets keyed by a multi-valued field or a hashed key domain,
with count / sum / min / max leaves on single- and multi-valued columns (terms.rs:172-179, sum.rs:131-140,
minmax.rs:135-145) — against the oracle and against the generic tree-walking kernel."""
import numpy as np
import pytest

import tantivy_aggregations_b200 as ta
from helpers import Corpus, SegSpec, assert_fruit_equal
from tantivy_aggregations_b200 import _ffi as F
from test_gpu_random import CAT, FVALS, ITAGS, PRICE, SIGNED, STATUS, TAGS, WIDE, make_corpus, queries

pytestmark = pytest.mark.gpu
RTOL = 1e-12
PATH_MTERMS, PATH_MTERMS_GENERIC = 4, 5


@pytest.fixture(scope="module")
def world(ctx):
    corpus = make_corpus(21, [5000, 0, 1024, 1025, 77, 20001])
    return corpus, corpus.build_gpu(ctx), corpus.build_oracle()


def shapes():
    return {
        "c4": lambda: ta.terms_agg_u64s(TAGS, ta.sum_agg_f64s(FVALS)),
        "multi_all_ops": lambda: ta.terms_agg_u64s(TAGS, (ta.count_agg(), ta.sum_agg_f64s(FVALS), ta.min_agg_f64s(FVALS), ta.max_agg_f64s(FVALS))),
        "multi_same_field": lambda: ta.terms_agg_u64s(TAGS, (ta.sum_agg_u64s(TAGS), ta.max_agg_u64s(TAGS), ta.count_agg())),
        "multi_single_leaves": lambda: ta.terms_agg_u64s(TAGS, (ta.count_agg(), ta.min_agg_f64(PRICE), ta.sum_agg_i64(SIGNED))),
        "multi_i64_keys": lambda: ta.terms_agg_i64s(ITAGS, (ta.count_agg(), ta.sum_agg_i64s(ITAGS))),
        "hashed_single_key": lambda: ta.terms_agg_u64(WIDE, (ta.count_agg(), ta.max_agg_u64(CAT), ta.sum_agg_f64s(FVALS))),
        "under_filters": lambda: ta.filter_agg(ta.TermQuery(STATUS, F.U64, 0),
                                               ta.post_filter_agg_f64(PRICE, ta.gt(30.0), ta.terms_agg_u64s(TAGS, (ta.count_agg(), ta.sum_agg_f64s(FVALS))))),
        "under_multi_post_filter": lambda: ta.post_filter_agg_u64s(TAGS, ta.in_set({101, 117, 129}), ta.terms_agg_u64s(TAGS, ta.count_agg())),
        "with_root_metrics": lambda: (ta.count_agg(), ta.terms_agg_u64s(TAGS, ta.sum_agg_f64s(FVALS)), ta.max_agg_f64(PRICE)),
    }


@pytest.mark.parametrize("qname", ["all", "bitset", "range_dev"])
@pytest.mark.parametrize("aname", sorted(shapes()))
def test_mterms_matches_oracle_and_generic(ctx, world, qname, aname):
    corpus, searcher, ox = world
    q = queries(corpus, 5)[qname]
    mk = shapes()[aname]
    want, _, _ = ox.search(q, mk())
    got, reader = searcher.agg_search_with_executor(q, mk(), ta.SINGLE_THREAD, return_reader=True)
    assert reader.stats()["path"] == PATH_MTERMS, "the plan must run on k_mterms (+ streaming launches), not the tree walker"
    assert_fruit_equal(got, want, RTOL)
    ctx.set_path(F.PATH_GENERIC)
    try:
        gen = searcher.agg_search(q, mk())
    finally:
        ctx.set_path(F.PATH_AUTO)
    assert_fruit_equal(gen, want, RTOL)


def test_mterms_leaves_the_rest_to_the_tree_walker(ctx, world):
    """A tuple whose other member has no fast shape: k_mterms takes the terms member, k_generic the nested one."""
    corpus, searcher, ox = world
    mk = lambda: (ta.terms_agg_u64s(TAGS, ta.sum_agg_f64s(FVALS)),
                  ta.terms_agg_u64(CAT, ta.terms_agg_u64(STATUS, ta.count_agg())))
    want, _, _ = ox.search(ta.AllQuery(), mk())
    got, reader = searcher.agg_search_with_executor(ta.AllQuery(), mk(), ta.SINGLE_THREAD, return_reader=True)
    assert reader.stats()["path"] == PATH_MTERMS_GENERIC
    assert_fruit_equal(got, want, RTOL)


def test_mterms_huge_fanout_and_ragged_docs(ctx):
    """Documents with more key occurrences than one expansion chunk (8192), empty documents, a ragged last tile."""
    rng = np.random.default_rng(5)
    n = 3000
    lists = [list(rng.integers(0, 500, size=rng.integers(0, 4))) for _ in range(n)]
    lists[7] = list(rng.integers(0, 500, size=20_000))
    lists[1500] = list(rng.integers(0, 500, size=8192))
    lists[2999] = list(rng.integers(0, 500, size=9000))
    s = SegSpec(n)
    s.mcol(TAGS, F.U64, lists)
    fv = [list(np.round(rng.random(rng.integers(0, 3)) * 10, 3)) for _ in range(n)]
    # leaf values are folded from a 1024-value staging buffer: documents whose values span several fills of it, a
    # range that ends exactly on a fill boundary, values of a deleted document in between
    fv[9] = list(np.round(rng.random(2500) * 10, 3))
    fv[1600] = list(np.round(rng.random(1100) * 10, 3))
    fv[2050] = list(np.round(rng.random(1024 - sum(len(x) for x in fv[2048:2050])) * 10, 3))
    fv[8] = list(np.round(rng.random(700) * 10, 3))
    s.mcol(FVALS, F.F64, fv)
    s.deleted = [0, 8, 2998]
    corpus = Corpus([s])
    searcher, ox = corpus.build_gpu(ctx), corpus.build_oracle()
    for mk in (lambda: ta.terms_agg_u64s(TAGS, (ta.count_agg(), ta.sum_agg_f64s(FVALS), ta.max_agg_f64s(FVALS))),
               lambda: ta.terms_agg_u64s(TAGS, ta.sum_agg_f64s(FVALS)),
               lambda: ta.terms_agg_u64s(TAGS, (ta.min_agg_f64s(FVALS), ta.sum_agg_u64s(TAGS)))):
        want, _, _ = ox.search(ta.AllQuery(), mk())
        got, reader = searcher.agg_search_with_executor(ta.AllQuery(), mk(), ta.SINGLE_THREAD, return_reader=True)
        assert reader.stats()["path"] == PATH_MTERMS
        assert_fruit_equal(got, want, RTOL)


def test_mterms_hashed_growth_multi_valued_keys(ctx):
    """Wide multi-valued keys: the global open-addressing table, including overflow -> grow -> redo."""
    rng = np.random.default_rng(6)
    n = 60_000
    lists = [list(rng.integers(0, 1 << 40, size=rng.integers(0, 6), dtype=np.uint64)) for _ in range(n)]
    s = SegSpec(n)
    s.mcol(TAGS, F.U64, lists)
    s.col(PRICE, F.F64, rng.random(n))
    corpus = Corpus([s])
    searcher, ox = corpus.build_gpu(ctx), corpus.build_oracle()
    mk = lambda: ta.terms_agg_u64s(TAGS, (ta.count_agg(), ta.min_agg_f64(PRICE)))
    want, _, _ = ox.search(ta.AllQuery(), mk())
    got, reader = searcher.agg_search_with_executor(ta.AllQuery(), mk(), ta.SINGLE_THREAD, return_reader=True)
    assert reader.stats()["path"] == PATH_MTERMS
    assert_fruit_equal(got, want, RTOL)
