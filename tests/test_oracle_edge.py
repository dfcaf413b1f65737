"""CPU: the oracle's f64 min / max / sum around NaN, -0.0 / +0.0 and the infinities, pinned against a direct Python
restatement of the reference's folds (src/metric/minmax.rs:97-106,59-72; src/metric/sum.rs:95-102,59-70) in both
executor shapes.  No reference test exercises these values (SURVEY §8c "parity-unpinned (5)"): the pin is the text of the
macros, restated twice independently (C++ in oracle/oracle.cpp, Python in tests/edge_cases.py)."""
import math

import numpy as np
import pytest

import tantivy_aggregations_b200 as ta
from edge_cases import SEQUENCES, bits, random_sequences, ref_search
from helpers import Corpus, SegSpec
from tantivy_aggregations_b200 import _ffi as F

PRICE = 3
ALL = dict(SEQUENCES)
ALL.update(random_sequences(7))


def corpus_of(segments):
    segs = []
    for vals in segments:
        s = SegSpec(len(vals))
        s.col(PRICE, F.F64, np.array(vals, dtype=np.float64))
        segs.append(s)
    return Corpus(segs)


@pytest.mark.parametrize("executor", ["SingleThread", "ThreadPool"])
@pytest.mark.parametrize("name", sorted(ALL))
def test_oracle_matches_reference_fold(name, executor):
    segments = ALL[name]
    ox = corpus_of(segments).build_oracle()
    got, _, _ = ox.search(ta.AllQuery(), (ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE), ta.sum_agg_f64(PRICE)),
                          mode=0 if executor == "SingleThread" else 1, threads=2)
    for op, g in zip(("min", "max", "sum"), got):
        w = ref_search(op, segments, executor)
        if w is None:
            assert g is None, (name, op, g)
        elif op == "sum" and math.isnan(w):
            assert g is not None and math.isnan(g), (name, op, g)
        else:
            assert g is not None and bits(g) == bits(w), (name, executor, op, g, w)


def test_codec_keeps_nan_payload_and_zero_sign():
    """The order-preserving code is a bijection on ALL bit patterns (tantivy common::f64_to_u64), so NaN payloads and the
    sign of zero survive the column."""
    from tantivy_aggregations_b200 import codec
    vals = np.array([float("nan"), -0.0, 0.0], dtype=np.float64)
    raw = np.array([0x7FF80000DEADBEEF, 0xFFF8000000000000, 0x8000000000000000, 0], dtype=np.uint64).view(np.float64)
    for arr in (vals, raw):
        codes = codec.values_to_codes(F.F64, arr)
        back = codec.code_to_f64(codes)
        assert (back.view(np.uint64) == arr.view(np.uint64)).all()
