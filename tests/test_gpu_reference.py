"""GPU: the reference's 19 known-answer tests through the C ABI (libtagg.so), on both column
ingestion routes and both executor shapes, plus codec parity of the device packer."""
import json
import os

import numpy as np
import pytest

import tantivy_aggregations_b200 as ta
from helpers import GOLDEN, ProductSchema, product_corpus
from reference_cases import CASES
from tantivy_aggregations_b200 import _ffi as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def searchers(ctx):
    full, empty = product_corpus(), product_corpus(empty=True)
    ofull = full.build_oracle()
    return {
        "codes": (full.build_gpu(ctx, "codes"), empty.build_gpu(ctx, "codes")),
        "bytes": (full.build_gpu(ctx, "bytes", ofull), empty.build_gpu(ctx, "codes")),
    }


@pytest.mark.parametrize("name,source,run", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("via", ["codes", "bytes"])
@pytest.mark.parametrize("executor", [ta.SINGLE_THREAD, ta.THREAD_POOL])
@pytest.mark.parametrize("path", [F.PATH_AUTO, F.PATH_GENERIC], ids=["auto", "generic"])
def test_reference_known_answers_gpu(ctx, searchers, name, source, run, via, executor, path):
    full, empty = searchers[via]
    ctx.set_path(path)
    try:
        run(lambda q, a: full.agg_search_with_executor(q, a, executor), ProductSchema,
            search_empty=lambda q, a: empty.agg_search_with_executor(q, a, executor))
    finally:
        ctx.set_path(F.PATH_AUTO)


def test_device_packer_matches_spec_vectors(ctx):
    """tagg_column_upload_codes re-packs on the device into tantivy's layout."""
    from test_oracle_golden import _codes_of
    with open(os.path.join(GOLDEN, "codec_vectors.json")) as f:
        v = json.load(f)
    for c in v["columns"]:
        codes = _codes_of(c)
        seg = ta.Segment(ctx, len(codes))
        seg.add_column_codes(0, F.U64, codes)
        info = seg.column_info(0)
        assert info["min_value"] == int(c["min_value"], 16) and info["amplitude"] == int(c["amplitude"], 16)
        assert info["num_bits"] == c["num_bits"]
        raw = seg.column_bytes(0)
        assert raw[16:].hex() == c["packed_hex"], c["name"]
        seg.close()


@pytest.mark.parametrize("nbits", [0, 1, 2, 7, 8, 13, 31, 32, 33, 55, 56, 64])
def test_device_packer_matches_oracle_all_widths(ctx, nbits):
    from oracle import oracle
    rng = np.random.default_rng(100 + nbits)
    n = 5000
    if nbits == 0:
        codes = np.full(n, 99, dtype=np.uint64)
    else:
        hi = (1 << nbits) - 1
        codes = rng.integers(0, hi, size=n, dtype=np.uint64, endpoint=True)
        codes[3], codes[17] = 0, hi
        if nbits < 64:
            codes = codes + np.uint64(1000)
    seg = ta.Segment(ctx, n)
    seg.add_column_codes(5, F.U64, codes)
    assert seg.column_bytes(5) == oracle.pack(codes)
    # and the kernels read it back exactly: sum / min / max over all docs
    s = ta.Searcher(ctx, [seg])
    got = s.agg_search(ta.AllQuery(), (ta.sum_agg_u64(5), ta.min_agg_u64(5), ta.max_agg_u64(5), ta.count_agg()))
    assert got == (int(codes.astype(object).sum() % (1 << 64)), int(codes.min()), int(codes.max()), n)


def test_missing_fast_field_is_an_error(ctx):
    """FastFieldNotAvailableError (sum.rs:50-55, terms.rs:76-81); histogram panics in the reference
    (histogram.rs:81) — here it is the same error instead of an abort across the ABI."""
    s = product_corpus().build_gpu(ctx)
    for agg in [ta.sum_agg_u64(42), ta.min_agg_f64(42), ta.terms_agg_u64(42, ta.count_agg()),
                ta.histogram_agg_f64(42, 0.0, 1.0, ta.count_agg()), ta.percentiles_agg_f64(42),
                ta.sum_agg_u64s(ProductSchema.price), ta.sum_agg_f64(ProductSchema.tag_ids)]:
        with pytest.raises(ta.FastFieldNotAvailableError):
            s.agg_search(ta.AllQuery(), agg)


def test_bad_plans_are_rejected(ctx):
    import ctypes as C
    lib = F.lib()
    def create(nodes):
        arr = (F.Node * len(nodes))(*nodes)
        h = C.c_void_p()
        return lib.tagg_plan_create(ctx._h, arr, len(nodes), None, 0, C.byref(h))
    assert create([F.Node(op=F.OP_TUPLE, n_children=1), F.Node(op=F.OP_COUNT)]) == F.ERR_BAD_PLAN   # arity < 2
    assert create([F.Node(op=F.OP_TERMS, kind=F.F64, n_children=1), F.Node(op=F.OP_COUNT)]) == F.ERR_BAD_PLAN
    assert create([F.Node(op=F.OP_TERMS, kind=F.U64, n_children=1)]) == F.ERR_BAD_PLAN              # truncated
    assert create([F.Node(op=F.OP_COUNT), F.Node(op=F.OP_COUNT)]) == F.ERR_BAD_PLAN                 # trailing
    assert create([F.Node(op=77)]) == F.ERR_BAD_PLAN
    assert create([F.Node(op=F.OP_HISTOGRAM, kind=F.F64, multi=1, n_children=1), F.Node(op=F.OP_COUNT)]) == F.ERR_BAD_PLAN
