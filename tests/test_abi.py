"""CPU-only: the C-ABI library loads without a GPU, exports every symbol the headers declare, and
fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes as C
import os
import re

import pytest

from tantivy_aggregations_b200 import _ffi as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("tagg.h", "tagg_synth.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(tagg_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(F.LIB_PATH), "libtagg.so not built (python -c 'import __graft_entry__ as g; g.build()')"
    lib = C.CDLL(F.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 35
    for name in sorted(decl):
        assert hasattr(lib, name), f"{name} declared in include/*.h but not exported by libtagg.so"
    # the python binding covers the whole ABI too
    assert decl == set(F.SYMBOLS), sorted(decl ^ set(F.SYMBOLS))


def test_abi_version_and_struct_layout():
    lib = F.lib()
    assert lib.tagg_abi_version() == 2
    # struct layouts mirrored in ctypes must match the C header (x86-64 SysV)
    assert C.sizeof(F.Node) == 48
    assert C.sizeof(F.Docset) == 40
    assert C.sizeof(F.SegmentInput) == 64
    assert C.sizeof(F.Blob) == 16


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = F.lib().tagg_ctx_create(0, C.byref(h))
    assert rc == F.ERR_NO_DEVICE
    assert b"no CPU fallback" in F.lib().tagg_last_error()


def test_plan_lowering_shapes():
    """Host logic: the typed tree lowers to the pre-order node array the header documents."""
    import tantivy_aggregations_b200 as ta
    from tantivy_aggregations_b200.agg import LowerCtx, as_agg
    agg = as_agg(ta.filter_agg(ta.TermQuery(9, ta.U64, 0),
                               (ta.count_agg(), ta.terms_agg_u64(1, (ta.count_agg(), ta.min_agg_f64(3))))))
    lc = LowerCtx()
    agg.lower(lc)
    ops = [n.op for n in lc.nodes]
    assert ops == [F.OP_FILTER, F.OP_TUPLE, F.OP_COUNT, F.OP_TERMS, F.OP_TUPLE, F.OP_COUNT, F.OP_MIN]
    assert [n.n_children for n in lc.nodes] == [1, 2, 0, 1, 2, 0, 0]
    assert lc.nodes[6].kind == F.F64 and lc.nodes[6].field_id == 3 and lc.nodes[3].field_id == 1
    assert len(lc.filters) == 1 and lc.nodes[0].aux == 0
    with pytest.raises(TypeError):
        as_agg((ta.count_agg(),))  # tuples have arity 2..=10 (src/tuple.rs:73-81)
    with pytest.raises(TypeError):
        as_agg(tuple(ta.count_agg() for _ in range(11)))


def test_predicate_code_ranges():
    import numpy as np
    import tantivy_aggregations_b200 as ta
    from tantivy_aggregations_b200 import codec
    vals = np.array([-np.inf, -3.5, -0.0, 0.0, 1.0, 5.0, 5.0000001, 1e300, np.inf, np.nan, -np.nan])
    codes = codec.f64_to_code(vals)
    for pred, ref in [(ta.gt(5.0), vals > 5.0), (ta.ge(5.0), vals >= 5.0), (ta.lt(0.0), vals < 0.0),
                      (ta.le(0.0), vals <= 0.0), (ta.eq(0.0), vals == 0.0), (ta.gt(0.0), vals > 0.0),
                      (ta.ge(-0.0), vals >= 0.0), (ta.eq(1.0), vals == 1.0)]:
        lo, hi = pred.code_range(F.F64)
        got = (codes >= np.uint64(lo)) & (codes <= np.uint64(hi)) if lo <= hi else np.zeros(len(vals), bool)
        assert (got == ref).all(), (pred.op, pred.x, got, ref)
    lo, hi = ta.ge(-5).code_range(F.I64)
    ci = codec.i64_to_code(np.array([-6, -5, 0, 7]))
    assert ((ci >= np.uint64(lo)) & (ci <= np.uint64(hi))).tolist() == [False, True, True, True]


def test_ckms_target_rank_matches_reference_vectors():
    """k = clamp(floor(q*n + max(1, floor(2*eps*q*n))/2), 1, n) reproduces percentile.rs:199-218."""
    import tantivy_aggregations_b200 as ta
    srt = [0.5, 9.99, 10.0, 50.0, 100.01]
    for q, want in [(0.5, 10.0), (0.33, 9.99), (0.7, 50.0), (0.01, 0.5), (0.99, 100.01)]:
        assert srt[ta.ckms_target_rank(q, 5) - 1] == want
