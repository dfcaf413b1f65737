"""GPU: tagg_execute_begin / tagg_pending_wait — several queries in flight on one context must each return exactly the
fruit of the synchronous call (own stream, own pinned block, own result image per call), including plans that need a
redo after the synchronisation point (hash table growth, exact percentiles, the ambiguous-zero path)."""
import ctypes as C

import numpy as np
import pytest

import tantivy_aggregations_b200 as ta
from helpers import Corpus, SegSpec, assert_fruit_equal
from tantivy_aggregations_b200 import _ffi as F
from tantivy_aggregations_b200 import index as I

pytestmark = pytest.mark.gpu
CAT, PRICE, WIDE = 1, 2, 3


def test_queries_in_flight_match_synchronous_results(ctx):
    rng = np.random.default_rng(9)
    segs = []
    for n in (40_000, 25_000, 3):
        s = SegSpec(n)
        s.col(CAT, F.U64, rng.integers(1, 300, size=n, dtype=np.uint64))
        s.col(PRICE, F.F64, np.round(rng.normal(0.0, 5.0, size=n), 1))  # spans zero, exact zeros of both signs appear
        s.col(WIDE, F.U64, rng.integers(1, 1 << 40, size=n, dtype=np.uint64))  # sparse keys: hashed scope that has to grow
        segs.append(s)
    corpus = Corpus(segs)
    searcher = corpus.build_gpu(ctx)
    ox = corpus.build_oracle()
    plans = [
        lambda: (ta.count_agg(), ta.sum_agg_f64(PRICE), ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE)),
        lambda: ta.terms_agg_u64(CAT, (ta.count_agg(), ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE))),
        lambda: ta.terms_agg_u64(WIDE, ta.count_agg()),
        lambda: ta.histogram_agg_f64(PRICE, -30.0, 2.5, (ta.count_agg(), ta.sum_agg_f64(PRICE))),
        lambda: (ta.percentiles_agg_f64(PRICE), ta.count_agg()),
    ]
    lib = F.lib()
    prepared = [searcher.prepare(mk()) for mk in plans]
    inputs = [I.build_inputs(p, ta.AllQuery(), searcher.segments) for p in prepared]
    for rounds in range(2):
        pend = []
        for p, (arr, keep) in zip(prepared, inputs):
            h = C.c_void_p()
            F.check(lib.tagg_execute_begin(p._h, arr, len(searcher.segments), C.byref(h)))
            pend.append(h)
        order = list(range(len(pend))) if rounds == 0 else list(reversed(range(len(pend))))  # waits need not follow the issue order
        for i in order:
            h = C.c_void_p()
            F.check(lib.tagg_pending_wait(pend[i], C.byref(h)))
            got = prepared[i].agg.decode(I.ResultReader(h), 0)
            want, _, _ = ox.search(ta.AllQuery(), plans[i]())
            assert_fruit_equal(got, want, 1e-12, f"plan{i}")


def test_begin_reports_errors_and_wait_consumes_the_handle(ctx):
    s = SegSpec(10)
    s.col(CAT, F.U64, np.arange(10, dtype=np.uint64))
    searcher = Corpus([s]).build_gpu(ctx)
    plan = searcher.prepare(ta.sum_agg_f64(PRICE))  # PRICE is not a column of the segment
    arr, keep = I.build_inputs(plan, ta.AllQuery(), searcher.segments)
    h = C.c_void_p()
    assert F.lib().tagg_execute_begin(plan._h, arr, 1, C.byref(h)) == F.ERR_NO_SUCH_COLUMN
