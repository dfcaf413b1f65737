"""CPU: the oracle's restatement of the two additions beyond the reference's own aggregations — date_histogram
(integer keys through histogram.rs:136-152's arithmetic) and cardinality (terms + count) — against numpy, so the GPU
tests of tests/test_gpu_extras.py compare with a checker that is itself checked."""
import numpy as np

import tantivy_aggregations_b200 as ta
from helpers import Corpus, SegSpec
from tantivy_aggregations_b200 import _ffi as F

DATE_F, CAT = 0, 1


def test_oracle_date_histogram_and_cardinality_against_numpy():
    rng = np.random.default_rng(9)
    n = 20_000
    t = rng.integers(-200_000, 40_000_000, size=n, dtype=np.int64)
    cat = rng.integers(0, 500, size=n, dtype=np.uint64)
    half = n // 2
    ox = Corpus([SegSpec(half).col(DATE_F, F.DATE, t[:half]).col(CAT, F.U64, cat[:half]),
                 SegSpec(n - half).col(DATE_F, F.DATE, t[half:]).col(CAT, F.U64, cat[half:])]).build_oracle()
    for interval, start in ((86_400, 0), (3_600, -200_000), (7 * 86_400, 1_000_000)):
        for mode in (0, 1):
            got, _, _ = ox.search(ta.AllQuery(), ta.date_histogram_agg(DATE_F, interval, ta.count_agg(), start=start), mode, 2)
            ok = t >= start
            ords, counts = np.unique((t[ok] - start) // interval, return_counts=True)
            assert dict(got._buckets) == {int(o): int(c) for o, c in zip(ords, counts)}
    # cardinality lowers to terms(field, count): the oracle returns that Terms fruit, its key set is the distinct set
    got, _, _ = ox.search(ta.AllQuery(), ta.cardinality_agg_u64(CAT))
    assert set(got.res) == set(np.unique(cat).tolist()) and sum(got.res.values()) == n
