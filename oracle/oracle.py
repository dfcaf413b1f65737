"""ctypes wrapper of the CPU oracle (oracle/oracle.cpp -> oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package.
"""
import ctypes as C
import os
import struct
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

# plain-struct descriptors shared with the product header (include/tagg.h)
import sys
sys.path.insert(0, os.path.dirname(_HERE))
from tantivy_aggregations_b200 import _ffi as F  # noqa: E402  (struct layouts + enums only)
from tantivy_aggregations_b200.fruits import Histogram, Terms  # noqa: E402


class OrcInput(C.Structure):
    _fields_ = [("seg", C.c_uint32), ("docset", F.Docset), ("filters", C.POINTER(F.Docset)), ("n_filters", C.c_uint32)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        l = C.CDLL(LIB_PATH)
        P, U64, SZ = C.c_void_p, C.c_uint64, C.c_size_t
        sig = {
            "orc_index_new": (P, []),
            "orc_index_free": (None, [P]),
            "orc_segment_add": (C.c_int, [P, C.c_uint32]),
            "orc_column_set": (C.c_int, [P, C.c_int, C.c_uint32, C.c_int, P, SZ]),
            "orc_column_set_codes": (C.c_int, [P, C.c_int, C.c_uint32, C.c_int, P, SZ]),
            "orc_multicolumn_set": (C.c_int, [P, C.c_int, C.c_uint32, C.c_int, P, SZ, P, SZ]),
            "orc_multicolumn_set_codes": (C.c_int, [P, C.c_int, C.c_uint32, C.c_int, P, SZ, P, SZ]),
            "orc_deletes_set": (C.c_int, [P, C.c_int, P, SZ]),
            "orc_column_bytes": (SZ, [P, C.c_int, C.c_uint32, C.c_int, P, SZ]),
            "orc_pack": (SZ, [P, SZ, P, SZ]),
            "orc_unpack": (C.c_int, [P, SZ, P, SZ]),
            "orc_num_bits": (C.c_uint32, [U64]),
            "orc_f64_to_code": (U64, [C.c_double]),
            "orc_code_to_f64": (C.c_double, [U64]),
            "orc_i64_to_code": (U64, [C.c_int64]),
            "orc_code_to_i64": (C.c_int64, [U64]),
            "orc_search": (C.c_int, [P, C.POINTER(F.Node), C.c_uint32, C.POINTER(F.Blob), C.c_uint32,
                                     C.POINTER(OrcInput), C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(P)]),
            "orc_result_size": (SZ, [P]),
            "orc_result_copy": (None, [P, P]),
            "orc_result_seconds": (C.c_double, [P]),
            "orc_result_collected": (U64, [P]),
            "orc_result_free": (None, [P]),
            "orc_ckms_new": (P, [C.c_double]),
            "orc_ckms_insert": (None, [P, P, SZ]),
            "orc_ckms_query": (C.c_int, [P, C.c_double, C.POINTER(C.c_double)]),
            "orc_ckms_len": (SZ, [P]),
            "orc_ckms_free": (None, [P]),
            "orc_synth_x": (U64, [U64, U64, U64]),
            "orc_synth_codes": (None, [C.c_int, U64, U64, U64, U64, U64, U64, U64, P]),
            "orc_synth_multi": (U64, [C.c_int, U64, U64, U64, U64, U64, U64, U64, U64, P, P]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and len(a) else None


# ---- codec helpers --------------------------------------------------------------------------------
def pack(codes):
    codes = np.ascontiguousarray(codes, dtype=np.uint64)
    n = lib().orc_pack(_ptr(codes), len(codes), None, 0)
    out = np.zeros(n, dtype=np.uint8)
    lib().orc_pack(_ptr(codes), len(codes), _ptr(out), n)
    return out.tobytes()


def unpack(raw, n):
    buf = np.frombuffer(raw, dtype=np.uint8)
    out = np.zeros(n, dtype=np.uint64)
    rc = lib().orc_unpack(_ptr(buf), len(buf), _ptr(out), n)
    if rc:
        raise ValueError("orc_unpack failed")
    return out


def synth_codes(recipe, seed, tag, doc_base, n, a=0, b=1, c=1):
    out = np.zeros(n, dtype=np.uint64)
    lib().orc_synth_codes(recipe, seed, tag, doc_base, n, a, b, c, _ptr(out))
    return out


def synth_multi(recipe, seed, tag, doc_base, n, count_mod, a=0, b=1, c=1):
    offsets = np.zeros(n + 1, dtype=np.uint64)
    total = lib().orc_synth_multi(recipe, seed, tag, doc_base, n, count_mod, a, b, c, _ptr(offsets), None)
    codes = np.zeros(total, dtype=np.uint64)
    lib().orc_synth_multi(recipe, seed, tag, doc_base, n, count_mod, a, b, c, _ptr(offsets), _ptr(codes) if total else None)
    return offsets, codes


class CKMS:
    """quantiles::ckms::CKMS<f64> restatement (tolerance witness)."""

    def __init__(self, eps=0.01):
        self._h = lib().orc_ckms_new(eps)

    def insert_many(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        lib().orc_ckms_insert(self._h, _ptr(v), len(v))

    def query(self, q):
        out = C.c_double()
        return out.value if lib().orc_ckms_query(self._h, q, C.byref(out)) else None

    def __len__(self):
        return lib().orc_ckms_len(self._h)

    def __del__(self):
        try:
            lib().orc_ckms_free(self._h)
        except Exception:
            pass


class OraclePercentiles:
    """Percentiles fruit of the oracle: the CKMS samples (v, g, delta) + n; query restated from
    quantiles 0.7 `Store::query`."""

    def __init__(self, n, samples):
        self.n = n
        self.samples = samples

    def percentile(self, q):
        s = self.samples
        if not s:
            return None
        import math
        r = 0
        nphi = q * self.n
        inv = max(1, math.floor(2.0 * 0.01 * nphi))
        for i in range(1, len(s)):
            r += s[i - 1][1]
            if r + s[i][1] + s[i][2] > nphi + inv / 2.0:
                return s[i - 1][0]
        return s[-1][0]

    def canon(self):
        return ("pct", self.n)


def _bits_to_value(kind, bits):
    if kind == F.U64:
        return bits
    if kind in (F.I64, F.DATE):
        return bits - (1 << 64) if bits >> 63 else bits
    return struct.unpack("<d", struct.pack("<Q", bits))[0]


def _decode(buf, pos):
    t = buf[pos]
    pos += 1
    if t == 0:
        return struct.unpack_from("<Q", buf, pos)[0], pos + 8
    if t == 1:
        kind, some = buf[pos], buf[pos + 1]
        v = struct.unpack_from("<Q", buf, pos + 2)[0]
        return (_bits_to_value(kind, v) if some else None), pos + 10
    if t == 2:
        n = struct.unpack_from("<I", buf, pos)[0]
        pos += 4
        items = []
        for _ in range(n):
            f, pos = _decode(buf, pos)
            items.append(f)
        return tuple(items), pos
    if t == 3:
        kind = buf[pos]
        n = struct.unpack_from("<Q", buf, pos + 1)[0]
        pos += 9
        res = {}
        for _ in range(n):
            k = struct.unpack_from("<Q", buf, pos)[0]
            f, pos = _decode(buf, pos + 8)
            res[_bits_to_value(kind, k)] = f
        return Terms(res), pos
    if t == 4:
        start, interval = struct.unpack_from("<dd", buf, pos)
        n = struct.unpack_from("<Q", buf, pos + 16)[0]
        pos += 24
        b = {}
        for _ in range(n):
            o = struct.unpack_from("<Q", buf, pos)[0]
            f, pos = _decode(buf, pos + 8)
            b[o] = f
        return Histogram(start, interval, b), pos
    if t == 5:
        n = struct.unpack_from("<Q", buf, pos)[0]
        m = struct.unpack_from("<I", buf, pos + 8)[0]
        pos += 12
        samples = []
        for _ in range(m):
            v, g, d = struct.unpack_from("<dII", buf, pos)
            samples.append((v, g, d))
            pos += 16
        return OraclePercentiles(n, samples), pos
    raise ValueError(f"bad fruit tag {t}")


class _SegView:
    """What product Query objects need from a segment (ord, max_doc, host columns)."""

    def __init__(self, ord_, max_doc):
        self.ord, self.max_doc = ord_, max_doc
        self.host = {}
        self.kinds = {}


class OracleIndex:
    def __init__(self):
        self._h = lib().orc_index_new()
        self.segs = []

    def __del__(self):
        try:
            lib().orc_index_free(self._h)
        except Exception:
            pass

    def add_segment(self, max_doc):
        s = lib().orc_segment_add(self._h, max_doc)
        self.segs.append(_SegView(s, max_doc))
        return s

    def set_column_codes(self, seg, field, kind, codes):
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        assert lib().orc_column_set_codes(self._h, seg, field, kind, _ptr(codes), len(codes)) == 0
        self.segs[seg].host[field] = codes
        self.segs[seg].kinds[field] = (kind, 0)

    def set_column_bytes(self, seg, field, kind, raw, host_codes=None):
        buf = np.frombuffer(raw, dtype=np.uint8)
        assert lib().orc_column_set(self._h, seg, field, kind, _ptr(buf), len(buf)) == 0
        if host_codes is not None:
            self.segs[seg].host[field] = host_codes
        self.segs[seg].kinds[field] = (kind, 0)

    def set_multicolumn_codes(self, seg, field, kind, offsets, codes):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        assert lib().orc_multicolumn_set_codes(self._h, seg, field, kind, _ptr(offsets), len(offsets),
                                               _ptr(codes), len(codes)) == 0
        self.segs[seg].host[field] = (offsets, codes)
        self.segs[seg].kinds[field] = (kind, 1)

    def set_deletes(self, seg, raw):
        buf = np.frombuffer(raw, dtype=np.uint8)
        assert lib().orc_deletes_set(self._h, seg, _ptr(buf), len(buf)) == 0

    def column_bytes(self, seg, field, which=0):
        n = lib().orc_column_bytes(self._h, seg, field, which, None, 0)
        out = np.zeros(n, dtype=np.uint8)
        lib().orc_column_bytes(self._h, seg, field, which, _ptr(out), n)
        return out.tobytes()

    def code_domain(self, field, multi):
        lo = hi = None
        for s in self.segs:
            h = s.host.get(field)
            if h is None:
                return None
            codes = h[1] if multi else h
            if len(codes) == 0:
                continue
            a, b = int(codes.min()), int(codes.max())
            lo = a if lo is None else min(lo, a)
            hi = b if hi is None else max(hi, b)
        return None if lo is None else (lo, hi)

    def search(self, query, agg, mode=0, threads=1, segments=None, decode=True):
        """agg_search on the CPU: mode 0 = Executor::SingleThread, 1 = ThreadPool(threads).
        Returns (fruit, seconds, collected docs)."""
        from tantivy_aggregations_b200.agg import LowerCtx, as_agg
        agg = as_agg(agg)
        lctx = LowerCtx(self)
        agg.lower(lctx)
        nodes = (F.Node * len(lctx.nodes))(*lctx.nodes)
        blob_bufs = [np.frombuffer(b, dtype=np.uint8) for b in lctx.blobs]
        blobs = (F.Blob * max(1, len(blob_bufs)))()
        for i, b in enumerate(blob_bufs):
            blobs[i].data = _ptr(b)
            blobs[i].len = len(b)
        segs = self.segs if segments is None else [self.segs[i] for i in segments]
        keep = []
        inputs = (OrcInput * max(1, len(segs)))()
        for i, sv in enumerate(segs):
            ds = query.docset(sv)
            keep.append(ds)
            inputs[i].seg = sv.ord
            inputs[i].docset = ds.c
            fl = (F.Docset * max(1, len(lctx.filters)))()
            for j, fq in enumerate(lctx.filters):
                fd = fq.docset(sv)
                keep.append(fd)
                fl[j] = fd.c
            keep.append(fl)
            inputs[i].filters = fl
            inputs[i].n_filters = len(lctx.filters)
        h = C.c_void_p()
        rc = lib().orc_search(self._h, nodes, len(lctx.nodes), blobs, len(blob_bufs), inputs, len(segs), mode, threads,
                              1 if decode else 0, C.byref(h))
        if rc:
            raise RuntimeError(f"oracle search failed with status {rc}")
        seconds = lib().orc_result_seconds(h)
        collected = lib().orc_result_collected(h)
        fruit = None
        if decode:
            n = lib().orc_result_size(h)
            buf = (C.c_uint8 * n)()
            lib().orc_result_copy(h, buf)
            fruit, _ = _decode(bytes(buf), 0)
            fruit = self._apply_key_filters(agg, fruit)
        lib().orc_result_free(h)
        return fruit, seconds, collected

    def _apply_key_filters(self, agg, fruit):
        """filtered_terms_agg_*: drop the buckets whose key fails the closure (terms.rs:322-330)."""
        from tantivy_aggregations_b200 import agg as A
        if isinstance(agg, A.TupleAgg):
            return tuple(self._apply_key_filters(m, f) for m, f in zip(agg.members, fruit))
        if isinstance(agg, A.TermsAgg):
            res = {k: self._apply_key_filters(agg.sub, v) for k, v in fruit.res.items()
                   if agg.key_filter is None or agg.key_filter(k)}
            return Terms(res)
        if isinstance(agg, A.HistogramAgg):
            return Histogram(fruit.start, fruit.interval, {k: self._apply_key_filters(agg.sub, v) for k, v in fruit._buckets.items()})
        if isinstance(agg, (A.FilterAgg, A.PostFilterAgg, A.GenericPostFilterAgg)):
            return self._apply_key_filters(agg.sub, fruit)
        if isinstance(agg, A.EitherAgg):
            f = self._apply_key_filters(agg.agg, fruit)
            return (agg.which, f) if agg.tag else f
        return fruit
