"""The tantivy `.fast` CompositeFile directory: the footer parser is host-only code, checked here without a
GPU against a writer restated from the same source (tantivy@14735ce common/composite_file.rs, common/vint.rs) — a
self-consistency pin, not a pin to real tantivy bytes (none are available offline); the GPU test loads a whole segment
from such a file and compares every aggregation against the oracle."""
import ctypes as C

import numpy as np
import pytest

import tantivy_aggregations_b200 as ta
from helpers import Corpus, SegSpec, assert_fruit_equal, composite_file, product_corpus, ProductSchema, vint
from tantivy_aggregations_b200 import _ffi as F


def entries_of(raw):
    buf = np.frombuffer(raw, dtype=np.uint8)
    n = C.c_uint32()
    F.check(F.lib().tagg_fast_file_entries(buf.ctypes.data_as(C.c_void_p), len(buf), None, None, None, None, 0, C.byref(n)))
    f, i = np.zeros(n.value, np.uint32), np.zeros(n.value, np.uint32)
    b, e = np.zeros(n.value, np.uint64), np.zeros(n.value, np.uint64)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    F.check(F.lib().tagg_fast_file_entries(ptr(buf), len(buf), ptr(f), ptr(i), ptr(b), ptr(e), n.value, C.byref(n)))
    return list(zip(f.tolist(), i.tolist(), b.tolist(), e.tolist()))


def test_vint_has_the_stop_bit_on_the_last_byte():
    assert vint(0) == b"\x80" and vint(5) == b"\x85" and vint(127) == b"\xff" and vint(128) == b"\x00\x81" and vint(300) == b"\x2c\x82"


def test_directory_of_a_composite_file():
    payloads = [((1, 0), b"a" * 24), ((4, 0), b"bb" * 100), ((4, 1), b"c" * 1000), ((9, 0), b""), ((200_000, 0), b"d" * 300)]
    raw = composite_file(payloads)
    got = entries_of(raw)
    at = 0
    for ((field, idx), data), (gf, gi, gb, ge) in zip(payloads, got):
        assert (gf, gi, gb, ge) == (field, idx, at, at + len(data))
        assert raw[gb:ge] == data
        at += len(data)
    assert entries_of(composite_file([])) == []


def test_malformed_files_are_rejected():
    raw = bytearray(composite_file([((1, 0), b"x" * 40)]))
    lib = F.lib()
    n = C.c_uint32()
    for bad in (raw[:3], raw[:-4] + b"\xff\xff\xff\x7f", raw[:40] + b"\x7f\x7f" + raw[-4:]):
        buf = np.frombuffer(bytes(bad), dtype=np.uint8)
        assert lib.tagg_fast_file_entries(buf.ctypes.data_as(C.c_void_p), len(buf), None, None, None, None, 0, C.byref(n)) == F.ERR_BAD_ARG


@pytest.mark.gpu
def test_segment_from_fast_file_matches_oracle(ctx):
    corpus = product_corpus()
    ox = corpus.build_oracle()
    S = ProductSchema
    seg = corpus.segs[0]
    payloads, fields, host = [], [], {}
    for f, (kind, codes) in seg.cols.items():
        payloads.append(((f, 0), ox.column_bytes(0, f, 0)))
        fields.append((f, kind, 0))
        host[f] = codes
    for f, (kind, offsets, codes) in seg.mcols.items():
        payloads.append(((f, 0), ox.column_bytes(0, f, 1)))   # offsets column
        payloads.append(((f, 1), ox.column_bytes(0, f, 0)))   # values column
        fields.append((f, kind, 1))
        host[f] = (offsets, codes)
    g = ta.Segment(ctx, seg.max_doc)
    g.load_fast_file(composite_file(payloads), fields, host)
    searcher = ta.Searcher(ctx, [g])
    for mk in (lambda: (ta.count_agg(), ta.sum_agg_f64(S.price), ta.min_agg_date(S.date_created), ta.max_agg_u64s(S.tag_ids)),
               lambda: ta.terms_agg_u64s(S.tag_ids, (ta.count_agg(), ta.histogram_agg_f64(S.price, 0.0, 10.0, ta.count_agg()))),
               lambda: ta.terms_agg_u64(S.category_id, (ta.count_agg(), ta.min_agg_f64(S.price)))):
        want, _, _ = ox.search(ta.AllQuery(), mk())
        assert_fruit_equal(searcher.agg_search(ta.AllQuery(), mk()), want)
    with pytest.raises(ta.FastFieldNotAvailableError):
        ta.Segment(ctx, seg.max_doc).load_fast_file(composite_file(payloads), [(77, F.U64, 0)])
