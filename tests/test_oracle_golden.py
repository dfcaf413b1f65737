"""CPU-only: pins the oracle (oracle/oracle.cpp) against the reference's own known-answer tests
(the 19 unit tests on the 5-document fixture) and the spec-derived codec vectors."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, ProductSchema, product_corpus
from oracle import oracle
from reference_cases import CASES
from tantivy_aggregations_b200 import _ffi as F
from tantivy_aggregations_b200 import codec


@pytest.fixture(scope="module")
def indexes():
    return product_corpus().build_oracle(), product_corpus(empty=True).build_oracle()


@pytest.mark.parametrize("name,source,run", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("mode", [0, 1], ids=["single_thread", "thread_pool"])
def test_reference_known_answers(indexes, name, source, run, mode):
    full, empty = indexes
    run(lambda q, a: full.search(q, a, mode=mode, threads=2)[0], ProductSchema,
        search_empty=lambda q, a: empty.search(q, a, mode=mode, threads=2)[0])


def _codes_of(c):
    if "codes" in c:
        return np.array([int(x, 16) for x in c["codes"]], dtype=np.uint64)
    if "values_f64" in c:
        return codec.values_to_codes(F.F64, np.array(c["values_f64"]))
    return codec.values_to_codes(F.I64, np.array(c["values_i64"]))


def test_codec_vectors():
    with open(os.path.join(GOLDEN, "codec_vectors.json")) as f:
        v = json.load(f)
    for c in v["columns"]:
        codes = _codes_of(c)
        raw = oracle.pack(codes)
        assert int.from_bytes(raw[0:8], "little") == int(c["min_value"], 16), c["name"]
        assert int.from_bytes(raw[8:16], "little") == int(c["amplitude"], 16), c["name"]
        assert oracle.lib().orc_num_bits(int(c["amplitude"], 16)) == c["num_bits"]
        assert codec.num_bits(int(c["amplitude"], 16)) == c["num_bits"]
        assert raw[16:].hex() == c["packed_hex"], c["name"]
        assert (oracle.unpack(raw, len(codes)) == codes).all(), c["name"]
    for val, code in v["f64_codes"]:
        assert oracle.lib().orc_f64_to_code(val) == int(code, 16)
        assert int(codec.f64_to_code(np.array([val]))[0]) == int(code, 16)
        back = oracle.lib().orc_code_to_f64(int(code, 16))
        assert np.float64(back).view(np.uint64) == np.float64(val).view(np.uint64)


def test_num_bits_rule():
    """tantivy compute_num_bits: widths above 56 are stored as 64."""
    for amp, nb in [(0, 0), (1, 1), (2, 2), (255, 8), (256, 9), ((1 << 56) - 1, 56), (1 << 56, 64), ((1 << 64) - 1, 64)]:
        assert oracle.lib().orc_num_bits(amp) == nb
        assert codec.num_bits(amp) == nb


@pytest.mark.parametrize("nbits", [0, 1, 2, 7, 8, 13, 31, 32, 33, 55, 56, 64])
def test_pack_roundtrip_widths(nbits):
    rng = np.random.default_rng(nbits)
    n = 1000
    if nbits == 0:
        codes = np.full(n, 12345, dtype=np.uint64)
    else:
        hi = (1 << nbits) - 1
        codes = rng.integers(0, hi, size=n, dtype=np.uint64, endpoint=True)
        codes[0], codes[1] = 0, hi  # force the full amplitude
        if nbits < 64:
            codes = codes + np.uint64(7)
    raw = oracle.pack(codes)
    assert len(raw) == 16 + (n * nbits + 7) // 8 + 7
    assert (oracle.unpack(raw, n) == codes).all()


def test_codec_numpy_matches_oracle():
    rng = np.random.default_rng(7)
    vals = np.concatenate([rng.normal(size=1000) * 1e6, [0.0, -0.0, np.inf, -np.inf, 1.0, 101.0]])
    c = codec.f64_to_code(vals)
    for v, cc in zip(vals, c):
        assert oracle.lib().orc_f64_to_code(float(v)) == int(cc)
    assert (codec.code_to_f64(c).view(np.uint64) == vals.view(np.uint64)).all()
    nz = vals != 0.0  # -0.0 == 0.0 compare equal as values but have distinct codes
    order = np.argsort(vals[nz], kind="stable")
    assert (np.diff(c[nz][order].astype(object)) >= 0).all()  # unsigned code order == value order
    assert int(codec.f64_to_code(np.array([-0.0]))[0]) + 1 == int(codec.f64_to_code(np.array([0.0]))[0])
    iv = rng.integers(-2**62, 2**62, size=1000)
    ci = codec.i64_to_code(iv)
    assert (codec.code_to_i64(ci) == iv).all()
    for v, cc in zip(iv[:50], ci[:50]):
        assert oracle.lib().orc_i64_to_code(int(v)) == int(cc)
