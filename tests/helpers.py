"""Shared test helpers: one in-memory corpus description that can be materialised both as a
GPU `Searcher` (product) and as an `OracleIndex` (CPU restatement), plus fruit comparison."""
import json
import math
import os
import struct

import numpy as np

import tantivy_aggregations_b200 as ta
from tantivy_aggregations_b200 import _ffi as F
from tantivy_aggregations_b200 import codec
from tantivy_aggregations_b200.fruits import Histogram, Percentiles, Terms

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class SegSpec:
    def __init__(self, max_doc):
        self.max_doc = max_doc
        self.cols = {}    # field -> (kind, codes)
        self.mcols = {}   # field -> (kind, offsets, codes)
        self.deleted = None  # iterable of deleted doc ids

    def col(self, field, kind, values):
        self.cols[field] = (kind, codec.values_to_codes(kind, values))
        return self

    def col_codes(self, field, kind, codes):
        self.cols[field] = (kind, np.ascontiguousarray(codes, dtype=np.uint64))
        return self

    def mcol(self, field, kind, lists):
        lens = np.array([len(l) for l in lists], dtype=np.uint64)
        offsets = np.zeros(len(lists) + 1, dtype=np.uint64)
        np.cumsum(lens, out=offsets[1:])
        flat = [v for l in lists for v in l]
        dt = {F.U64: np.uint64, F.F64: np.float64}.get(kind, np.int64)
        self.mcols[field] = (kind, offsets, codec.values_to_codes(kind, np.array(flat, dtype=dt)))
        return self

    def mcol_codes(self, field, kind, offsets, codes):
        self.mcols[field] = (kind, np.ascontiguousarray(offsets, dtype=np.uint64), np.ascontiguousarray(codes, dtype=np.uint64))
        return self

    def deletes_raw(self):
        if self.deleted is None:
            return None
        bits = np.zeros(self.max_doc, dtype=np.uint8)
        idx = np.asarray(sorted(set(self.deleted)), dtype=np.int64)
        if len(idx):
            bits[idx] = 1
        return np.packbits(bits, bitorder="little").tobytes() if self.max_doc else b""


class Corpus:
    def __init__(self, segs):
        self.segs = segs

    def build_oracle(self):
        from oracle import oracle
        ix = oracle.OracleIndex()
        for s in self.segs:
            o = ix.add_segment(s.max_doc)
            for f, (kind, codes) in s.cols.items():
                ix.set_column_codes(o, f, kind, codes)
            for f, (kind, offsets, codes) in s.mcols.items():
                ix.set_multicolumn_codes(o, f, kind, offsets, codes)
            raw = s.deletes_raw()
            if raw is not None:
                ix.set_deletes(o, raw)
        return ix

    def build_gpu(self, ctx, via="codes", oracle_index=None):
        """via='codes': device re-packs decoded codes; via='bytes': tantivy-layout bytes (produced by
        the oracle's restated serializer) are uploaded unchanged."""
        segments = []
        for i, s in enumerate(self.segs):
            g = ta.Segment(ctx, s.max_doc)
            for f, (kind, codes) in s.cols.items():
                if via == "codes":
                    g.add_column_codes(f, kind, codes)
                else:
                    g.add_column_bytes(f, kind, oracle_index.column_bytes(i, f, 0), host_codes=codes)
            for f, (kind, offsets, codes) in s.mcols.items():
                if via == "codes":
                    g.add_multicolumn_codes(f, kind, offsets, codes)
                else:
                    g.add_multicolumn_bytes(f, kind, oracle_index.column_bytes(i, f, 1), oracle_index.column_bytes(i, f, 0),
                                            host=(offsets, codes))
            raw = s.deletes_raw()
            if raw is not None:
                g.set_deletes(raw=raw)
            segments.append(g)
        return ta.Searcher(ctx, segments)


class ProductSchema:
    """test_fixtures/src/lib.rs:101-128"""
    id, category_id, tag_ids, price, positive_opinion_percent, attr_facets, date_created = range(7)


def product_corpus(empty=False):
    """The reference's 5-document golden corpus (one segment, no deletes)."""
    with open(os.path.join(GOLDEN, "product_fixture.json")) as f:
        fx = json.load(f)
    docs = [] if empty else fx["docs"]
    S = ProductSchema
    seg = SegSpec(len(docs))
    seg.col(S.category_id, F.U64, np.array([d["category_id"] for d in docs], dtype=np.uint64))
    seg.col(S.price, F.F64, np.array([d["price"] for d in docs], dtype=np.float64))
    seg.col(S.positive_opinion_percent, F.U64, np.array([d["positive_opinion_percent"] for d in docs], dtype=np.uint64))
    # a missing single-valued fast field value reads as 0 (tantivy writer default)
    seg.col(S.date_created, F.DATE, np.array([d["date_created"] or 0 for d in docs], dtype=np.int64))
    seg.mcol(S.tag_ids, F.U64, [d["tag_ids"] for d in docs])
    seg.mcol(S.attr_facets, F.U64, [[] for _ in docs])
    return Corpus([seg])


def vint(v):
    """tantivy VInt (common/vint.rs): 7 bits per byte, least significant group first, the LAST byte carries the stop bit."""
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v == 0:
            out.append(b | 0x80)
            return bytes(out)
        out.append(b)


def composite_file(payloads):
    """A tantivy CompositeFile (common/composite_file.rs, restated): payloads = [((field, idx), bytes)] in write order."""
    body, entries = bytearray(), []
    for (field, idx), raw in payloads:
        entries.append((len(body), field, idx))
        body += raw
    footer = bytearray(vint(len(entries)))
    prev = 0
    for off, field, idx in entries:
        footer += vint(off - prev) + struct.pack("<I", field) + vint(idx)
        prev = off
    return bytes(body) + bytes(footer) + struct.pack("<I", len(footer))


def f64_bits(x):
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def assert_fruit_equal(got, want, f64_sum_rtol=0.0, path="fruit", nan_equal=False):
    """Structural equality of two fruits.  Everything is bit-exact except f64 values when
    f64_sum_rtol > 0 (then |got-want| <= rtol*|want| is accepted, for order-dependent f64 sums; a zero must still
    match in sign) and, with nan_equal, any two NaNs (an f64 sum that is NaN carries no defined payload)."""
    if isinstance(want, tuple):
        assert isinstance(got, tuple) and len(got) == len(want), f"{path}: tuple arity {got!r} vs {want!r}"
        for i, (g, w) in enumerate(zip(got, want)):
            assert_fruit_equal(g, w, f64_sum_rtol, f"{path}.{i}", nan_equal)
    elif isinstance(want, Terms):
        assert isinstance(got, Terms), f"{path}: {type(got)}"
        assert set(got.res) == set(want.res), f"{path}: bucket keys differ: {sorted(set(got.res) ^ set(want.res))[:10]}"
        for k in want.res:
            assert_fruit_equal(got.res[k], want.res[k], f64_sum_rtol, f"{path}[{k}]", nan_equal)
    elif isinstance(want, Histogram):
        assert isinstance(got, Histogram), f"{path}: {type(got)}"
        assert f64_bits(got.start) == f64_bits(want.start) and f64_bits(got.interval) == f64_bits(want.interval)
        assert set(got._buckets) == set(want._buckets), f"{path}: bucket ords differ {sorted(got._buckets)} vs {sorted(want._buckets)}"
        for k in want._buckets:
            assert_fruit_equal(got._buckets[k], want._buckets[k], f64_sum_rtol, f"{path}<{k}>", nan_equal)
    elif hasattr(want, "percentile"):
        assert hasattr(got, "percentile"), f"{path}: {type(got)}"
        assert got.n == want.n, f"{path}: percentile n {got.n} vs {want.n}"
        assert_percentiles_agree(got, want, path)
    elif isinstance(want, float):
        assert isinstance(got, float), f"{path}: {got!r} vs {want!r}"
        if f64_bits(got) == f64_bits(want):
            return
        if nan_equal and math.isnan(want) and math.isnan(got):
            return
        if f64_sum_rtol > 0 and not math.isnan(want) and want != 0.0:
            assert abs(got - want) <= f64_sum_rtol * abs(want), f"{path}: {got!r} vs {want!r} (rtol {f64_sum_rtol})"
        else:
            raise AssertionError(f"{path}: f64 bits differ: {got!r} vs {want!r}")
    else:
        assert got == want and type(got) == type(want), f"{path}: {got!r} vs {want!r}"


PCT_QS = (0.01, 0.25, 0.5, 0.75, 0.95, 0.99)


def assert_percentiles_agree(got, want, path="pct", qs=PCT_QS, eps=0.01):
    """percentile(q) of the product fruit (exact order statistics) against the oracle's CKMS sketch (percentile.rs:163-177).

    * while the sketch holds every value (no compression yet: every g == 1, delta == 0) both answer with the same order
      statistic, so the values must be EQUAL;
    * otherwise the sketch only brackets ranks: sample i stands for g_i inserted values in (v_{i-1}, v_i] (compress folds a
      sample into its right neighbour only), so a value x has more than sum(g_j : v_j < x) values below it and fewer than
      sum(g_j : j <= first sample above x) at or below it, each good to eps * rank.  The (rank, value) pair the product
      answers with must sit inside that bracket, and the rank within the estimator's own band eps*k (+1) of the rank k
      CKMS targets.
    Skipped when the sketch's weights do not add up to n (the reference's lossy thread-pool merge, percentile.rs:58-62)."""
    samples = getattr(want, "samples", None)
    if samples is None or not want.n:
        for q in qs:
            assert got.percentile(q) == want.percentile(q), f"{path}: percentile({q})"
        return
    if sum(g for _, g, _ in samples) != want.n:
        return
    uncompressed = len(samples) == want.n and all(g == 1 and d == 0 for _, g, d in samples)
    vals = np.array([v for v, _, _ in samples])
    rmin = np.cumsum([g for _, g, _ in samples])
    for q in qs:
        x = got.percentile(q)
        assert x is not None, f"{path}: percentile({q}) is None"
        if uncompressed:
            assert f64_bits(x) == f64_bits(want.percentile(q)), f"{path}: percentile({q}) {x!r} vs {want.percentile(q)!r}"
            continue
        k, r = got.rank_of_answer(q)
        assert abs(r - k) <= eps * k + 1, f"{path}: percentile({q}) answers rank {r}, CKMS targets {k}"
        below = np.searchsorted(vals, x, side="left")    # samples < x
        above = np.searchsorted(vals, x, side="right")   # first sample > x
        # (CKMS's cumulative weights are themselves only good to eps * rank — measured on the restatement — hence the slack)
        lo = int(rmin[below - 1]) + 1 if below > 0 else 1
        hi = int(rmin[above]) - 1 if above < len(vals) else want.n
        lo, hi = lo - (eps * lo + 1), hi + (eps * hi + 1)
        assert lo <= r <= hi, f"{path}: percentile({q}) = {x!r} claims rank {r}, the oracle sketch brackets it to [{lo}, {hi}]"


def exact_rank_window(sorted_vals, value):
    """1-based [lowest, highest] rank a value occupies in a sorted array."""
    lo = int(np.searchsorted(sorted_vals, value, side="left")) + 1
    hi = int(np.searchsorted(sorted_vals, value, side="right"))
    return lo, hi
