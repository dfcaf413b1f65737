"""Shared by the CPU (oracle) and GPU edge-semantics tests: planted f64 sequences around NaN, the two zeros and the
infinities, and a direct Python restatement of the reference's folds to pin them.

Reference text:
  min / max  src/metric/minmax.rs:97-106 (single), :135-145 (multi), merge :59-72 —
             `if let Some(value) = fruit { if v.lt(value) / v.gt(value) { *value = v } } else { fruit.replace(v) }`
             with `PartialOrd` on f64: nothing is lt / gt a NaN, so a FIRST NaN sticks and later NaNs never replace;
             -0.0 == +0.0, so among the two zeros the first one collected stays.
  sum        src/metric/sum.rs:95-102, :131-140, merge :59-70 — first value replaces, then `+=` in collection order.
"""
import struct

import numpy as np

NAN = float("nan")
NEG_NAN = struct.unpack("<d", struct.pack("<Q", 0xFFF8000000000000))[0]
PAYLOAD_NAN = struct.unpack("<d", struct.pack("<Q", 0x7FF80000DEADBEEF))[0]
INF = float("inf")


def bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def ref_fold(op, values, acc=None):
    """The reference's collect loop on one fruit (Option<f64>); also its merge when `values` are segment fruits."""
    for v in values:
        if v is None:
            continue
        if acc is None:
            acc = v
        elif op == "sum":
            acc = acc + v
        elif op == "min":
            if v < acc:
                acc = v
        else:
            if v > acc:
                acc = v
    return acc


def ref_search(op, segments, executor):
    """segments: list of per-segment value lists in collection order.  SingleThread threads one fruit through all of
    them (searcher.rs:66-78); ThreadPool folds each into its own fruit and merges in segment order (:79-98)."""
    if executor == "SingleThread":
        return ref_fold(op, [v for seg in segments for v in seg])
    return ref_fold(op, [ref_fold(op, seg) for seg in segments])


# per case: list of segments, each a list of f64 values (one per document, in doc order)
SEQUENCES = {
    "nan_first": [[NAN, 1.0, 2.0]],
    "nan_middle": [[1.0, NAN, 0.5, NAN, 3.0]],
    "nan_only": [[NAN]],
    "nan_all": [[NAN, NEG_NAN, PAYLOAD_NAN]],
    "payload_nan_first": [[PAYLOAD_NAN, NAN, 7.0]],
    "neg_nan_first": [[NEG_NAN, -1.0, 1.0]],
    "neg_nan_later": [[4.0, NEG_NAN, 5.0]],
    "poszero_then_negzero": [[0.0, -0.0]],
    "negzero_then_poszero": [[-0.0, 0.0]],
    "zeros_after_positive": [[5.0, 0.0, -0.0, 7.0]],
    "zeros_after_negative": [[-5.0, -0.0, 0.0, -7.0]],
    "zeros_interleaved": [[3.0, -0.0, 0.0, -0.0, 1.0, 0.0]],
    "only_negzero": [[-0.0, -0.0, -0.0]],
    "only_poszero": [[0.0, 0.0]],
    "infinities": [[INF, -INF, 3.0]],
    "inf_and_nan": [[INF, NAN, -INF]],
    "two_segments_nan_heads_second": [[3.0], [NAN, 1.0]],
    "two_segments_nan_heads_first": [[NAN, 1.0], [0.5]],
    "two_segments_zeros": [[1.0, 0.0], [-0.0, 2.0]],
    "two_segments_zeros_rev": [[1.0, -0.0], [0.0, 2.0]],
    "three_segments_mixed": [[2.0, -0.0], [], [NAN, 0.0, -3.0]],
    "empty_then_nan": [[], [NAN], [1.0]],
    "plain_signed": [[-2.5, 3.5, -0.5, 0.25]],
}


def random_sequences(seed, n_cases=12):
    """Random multi-segment sequences over a small alphabet of special and ordinary values."""
    rng = np.random.default_rng(seed)
    alphabet = [NAN, NEG_NAN, PAYLOAD_NAN, 0.0, -0.0, INF, -INF, 1.5, -1.5, 2.0, -2.0, 0.25, -0.25]  # sums are exact in any order
    out = {}
    for c in range(n_cases):
        segs = []
        for _ in range(int(rng.integers(1, 5))):
            n = int(rng.integers(0, 40))
            # a few cases without NaN (zeros only), a few dense in NaN
            pool = alphabet[3:] if c % 3 == 0 else alphabet
            segs.append([pool[i] for i in rng.integers(0, len(pool), size=n)])
        out[f"random_{seed}_{c}"] = segs
    return out
