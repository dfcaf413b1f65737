"""Host driver: `AggSearcher::{agg_search, agg_search_with_executor}` (reference
src/searcher.rs:12-25,53-101) over device-resident segments, plus the minimal stand-ins for the
tantivy objects the reference touches on this path (Searcher / SegmentReader / Query -> Scorer).

tantivy's inverted index, query parsing and scoring are OUT OF SCOPE: a `Query`
here only has to yield, per segment, the matched-doc set the real scorer would yield — as a
docset handed to the GPU (ALL / bitset / sorted ids), or a COLUMN_RANGE evaluated on the device
when the queried field is also a FAST field.
"""
import ctypes as C

import numpy as np

from . import _ffi as F
from . import codec
from .agg import LowerCtx, as_agg

SINGLE_THREAD = "SingleThread"   # tantivy::Executor::SingleThread
THREAD_POOL = "ThreadPool"       # tantivy::Executor::ThreadPool(_)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """One per GPU / process (tagg_ctx)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        F.check(F.lib().tagg_ctx_create(int(device), C.byref(self._h)))
        self.device = int(device)

    def set_path(self, path):
        F.check(F.lib().tagg_ctx_set_path(self._h, int(path)))

    def synchronize(self):
        F.check(F.lib().tagg_ctx_synchronize(self._h))

    def timer_start(self):
        F.check(F.lib().tagg_ctx_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_double()
        F.check(F.lib().tagg_ctx_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        n = C.c_uint64()
        F.check(F.lib().tagg_ctx_launch_count(self._h, C.byref(n)))
        return n.value

    # multi-GPU (one process per GPU): the host broadcasts rank 0's id, e.g. with torch.distributed
    @staticmethod
    def comm_unique_id():
        buf = (C.c_uint8 * F.UNIQUE_ID_BYTES)()
        F.check(F.lib().tagg_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id, rank, n_ranks):
        buf = (C.c_uint8 * F.UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        F.check(F.lib().tagg_comm_init(self._h, buf, int(rank), int(n_ranks)))
        self.rank, self.n_ranks = rank, n_ranks

    def close(self):
        if self._h:
            F.lib().tagg_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Segment:
    """A tantivy segment's fast fields resident in HBM (tagg_segment).  `keep_host=True` also keeps
    the decoded codes on the host — the stand-in for the SegmentReader the host-side query /
    closure evaluation would consult."""

    def __init__(self, ctx, max_doc, keep_host=True):
        self.ctx = ctx
        self.max_doc = int(max_doc)
        self.keep_host = keep_host
        self.host = {}        # field -> codes (np.uint64) | (offsets, codes)
        self.kinds = {}       # field -> (kind, multi)
        self.deletes = None
        self._h = C.c_void_p()
        F.check(F.lib().tagg_segment_create(ctx._h, self.max_doc, C.byref(self._h)))

    # -- columns ---------------------------------------------------------------------------
    def add_column(self, field, kind, values):
        """Single-valued fast field from values in their natural type (the writer-side view)."""
        self.add_column_codes(field, kind, codec.values_to_codes(kind, values))

    def add_column_codes(self, field, kind, codes):
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        if len(codes) != self.max_doc:
            raise ValueError("a single-valued column has exactly max_doc values")
        F.check(F.lib().tagg_column_upload_codes(self._h, int(field), int(kind), _ptr(codes), len(codes)))
        self.kinds[int(field)] = (kind, 0)
        if self.keep_host:
            self.host[int(field)] = codes

    def add_column_bytes(self, field, kind, raw, host_codes=None):
        """Single-valued fast field from tantivy's own column bytes (zero re-encoding)."""
        buf = np.frombuffer(raw, dtype=np.uint8)
        F.check(F.lib().tagg_column_upload(self._h, int(field), int(kind), _ptr(buf), len(buf)))
        self.kinds[int(field)] = (kind, 0)
        if host_codes is not None and self.keep_host:
            self.host[int(field)] = np.ascontiguousarray(host_codes, dtype=np.uint64)

    def add_multicolumn(self, field, kind, lists):
        """Multi-valued fast field from per-doc value lists."""
        lens = np.fromiter((len(l) for l in lists), dtype=np.uint64, count=len(lists))
        offsets = np.zeros(len(lists) + 1, dtype=np.uint64)
        np.cumsum(lens, out=offsets[1:])
        flat = [v for l in lists for v in l]
        codes = codec.values_to_codes(kind, np.array(flat, dtype={F.U64: np.uint64, F.F64: np.float64}.get(kind, np.int64)))
        self.add_multicolumn_codes(field, kind, offsets, codes)

    def add_multicolumn_codes(self, field, kind, offsets, codes):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        if len(offsets) != self.max_doc + 1:
            raise ValueError("a multi-valued column has max_doc+1 offsets")
        F.check(F.lib().tagg_multicolumn_upload_codes(self._h, int(field), int(kind), _ptr(offsets), len(offsets),
                                                      _ptr(codes), len(codes)))
        self.kinds[int(field)] = (kind, 1)
        if self.keep_host:
            self.host[int(field)] = (offsets, codes)

    def add_multicolumn_bytes(self, field, kind, idx_raw, vals_raw, host=None):
        ib = np.frombuffer(idx_raw, dtype=np.uint8)
        vb = np.frombuffer(vals_raw, dtype=np.uint8)
        F.check(F.lib().tagg_multicolumn_upload(self._h, int(field), int(kind), _ptr(ib), len(ib), _ptr(vb), len(vb)))
        self.kinds[int(field)] = (kind, 1)
        if host is not None and self.keep_host:
            self.host[int(field)] = host

    def load_fast_file(self, raw, fields, host=None):
        """All fast fields of the segment from its tantivy `.fast` CompositeFile bytes.  fields: [(field, kind, multi)]."""
        buf = np.frombuffer(raw, dtype=np.uint8)
        arr = (F.FastField * max(1, len(fields)))(*[F.FastField(int(f), int(k), int(m)) for f, k, m in fields])
        F.check(F.lib().tagg_segment_load_fast_file(self._h, _ptr(buf), len(buf), arr, len(fields)))
        for f, k, m in fields:
            self.kinds[int(f)] = (k, int(m))
        if host and self.keep_host:
            self.host.update(host)

    def synth_column(self, field, kind, recipe, seed, tag, doc_base, a=0, b=1, c=1):
        """Generate a synthetic column on the device (bench / scale tests; SURVEY §8d recipe)."""
        F.check(F.lib().tagg_synth_column(self._h, int(field), int(kind), int(recipe), int(seed), int(tag),
                                          int(doc_base), int(a), int(b), int(c)))
        self.kinds[int(field)] = (kind, 0)

    def synth_multicolumn(self, field, kind, recipe, seed, tag, doc_base, count_mod, a=0, b=1, c=1):
        F.check(F.lib().tagg_synth_multicolumn(self._h, int(field), int(kind), int(recipe), int(seed), int(tag),
                                               int(doc_base), int(count_mod), int(a), int(b), int(c)))
        self.kinds[int(field)] = (kind, 1)

    def add_doc_address_column(self, field, base):
        """A u64 fast field whose value for document d is base + d, generated on the device (tagg_segment_doc_address_column):
        the key column of Searcher.top_hits_f64."""
        F.check(F.lib().tagg_segment_doc_address_column(self._h, int(field), int(base)))
        self.kinds[int(field)] = (F.U64, 0)
        self._doc_addr = (int(field), int(base))

    def set_deletes(self, deleted_docs=None, raw=None):
        """DeleteBitSet: either an iterable of deleted doc ids or the raw bitset bytes."""
        if raw is None:
            bits = np.zeros(self.max_doc, dtype=np.uint8)
            bits[np.asarray(list(deleted_docs), dtype=np.int64)] = 1
            raw = np.packbits(bits, bitorder="little").tobytes()
        buf = np.frombuffer(raw, dtype=np.uint8)
        F.check(F.lib().tagg_segment_set_deletes(self._h, _ptr(buf), len(buf)))
        self.deletes = bytes(raw)

    # -- device-resident docsets -------------------------------------------------------------
    def cache_docset(self, docset):
        """Evaluate / upload a docset once; returns a Docset (DEVICE_BITSET) reusable across queries."""
        out = F.Docset()
        F.check(F.lib().tagg_docset_cache(self._h, C.byref(docset.c), C.byref(out)))
        d = Docset.__new__(Docset)
        d.buf, d.c = None, out
        return d

    def docset_to_bitset(self, docset):
        """Evaluate a docset on the device and return its bitset bytes (np.uint8)."""
        out = np.zeros((self.max_doc + 7) // 8, dtype=np.uint8)
        F.check(F.lib().tagg_docset_to_bitset(self._h, C.byref(docset.c), _ptr(out), len(out)))
        return out

    # -- introspection ---------------------------------------------------------------------
    def column_info(self, field, which=0):
        mn, amp, nv, pl = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        nb = C.c_uint32()
        F.check(F.lib().tagg_column_info(self._h, int(field), which, C.byref(mn), C.byref(amp), C.byref(nb),
                                         C.byref(nv), C.byref(pl)))
        return dict(min_value=mn.value, amplitude=amp.value, num_bits=nb.value, n_values=nv.value,
                    packed_len=pl.value)

    def column_bytes(self, field, which=0):
        info = self.column_info(field, which)
        out = np.zeros(info["packed_len"], dtype=np.uint8)
        F.check(F.lib().tagg_column_download(self._h, int(field), which, _ptr(out), len(out)))
        return out.tobytes()

    def close(self):
        if self._h:
            F.lib().tagg_segment_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- queries: what a tantivy Weight::scorer(segment) yields, as a docset -------------------
class Docset:
    """Keeps the numpy buffer alive next to the C struct."""

    def __init__(self, kind, data=None, n=0, field=0, lo=0, hi=0):
        self.buf = data
        self.c = F.Docset(kind=kind, field_id=field, data=_ptr(data) if data is not None else None, n=n, lo=lo, hi=hi)


class Query:
    def docset(self, seg):
        raise NotImplementedError


class AllQuery(Query):
    """tantivy::query::AllQuery -> AllScorer"""

    def docset(self, seg):
        return Docset(F.DOCSET_ALL)


class BitsetQuery(Query):
    """A query whose scorer output is already a per-segment bitset (e.g. RangeQuery's BitSetDocSet)."""

    def __init__(self, per_segment):
        self.per_segment = per_segment  # dict id(segment)/index -> bytes / np.uint8

    def docset(self, seg):
        raw = self.per_segment[seg.ord]
        buf = np.frombuffer(raw, dtype=np.uint8) if not isinstance(raw, np.ndarray) else raw
        return Docset(F.DOCSET_BITSET, buf, len(buf))


class CachedQuery(Query):
    """A query whose per-segment docsets were made device-resident once (Segment.cache_docset)."""

    def __init__(self, query, segments):
        self.per_segment = {seg.ord: seg.cache_docset(query.docset(seg)) for seg in segments}

    def docset(self, seg):
        return self.per_segment[seg.ord]


class DocIdsQuery(Query):
    """A query whose scorer output is a sorted doc-id list per segment (e.g. TermScorer postings)."""

    def __init__(self, per_segment):
        self.per_segment = per_segment

    def docset(self, seg):
        ids = np.ascontiguousarray(self.per_segment[seg.ord], dtype=np.uint32)
        return Docset(F.DOCSET_SORTED_IDS, ids, len(ids))


class RangeQuery(Query):
    """tantivy RangeQuery / TermQuery on a field that is also FAST: lo <= value <= hi, inclusive,
    in the field's natural type (use `lo == hi` for a term).  device=True evaluates it on the GPU
    from the resident column (COLUMN_RANGE); device=False evaluates on the host into a bitset
    (what decoding the postings would produce) and ships it."""

    def __init__(self, field, kind, lo, hi, device=True):
        self.field, self.kind, self.device = int(field), kind, device
        self.lo, self.hi = codec.scalar_code(kind, lo), codec.scalar_code(kind, hi)

    @classmethod
    def half_open(cls, field, kind, lo, hi, device=True):
        """`lo..hi` (RangeQuery::new_f64(field, lo..hi)): lo <= value < hi."""
        q = cls(field, kind, lo, hi, device)
        q.hi -= 1  # codes are order preserving: the largest code below code(hi)
        return q

    def docset(self, seg):
        if self.lo > self.hi:  # empty range
            return Docset(F.DOCSET_SORTED_IDS, np.zeros(0, dtype=np.uint32), 0)
        if self.device:
            return Docset(F.DOCSET_COLUMN_RANGE, field=self.field, lo=self.lo, hi=self.hi)
        codes = seg.host[self.field]
        m = (codes >= np.uint64(self.lo)) & (codes <= np.uint64(self.hi))
        buf = np.packbits(m.astype(np.uint8), bitorder="little")
        return Docset(F.DOCSET_BITSET, buf, len(buf))


def TermQuery(field, kind, value, device=True):
    return RangeQuery(field, kind, value, value, device)


class HostPredicateQuery(Query):
    """post_filter_agg_* closure that could not be lowered: evaluated on the host per value."""

    def __init__(self, field, kind, multi, pred):
        self.field, self.kind, self.multi, self.pred = field, kind, multi, pred

    def docset(self, seg):
        from .agg import _codes_to_bits
        h = seg.host[self.field]
        if self.multi:
            offsets, codes = h
            bits = _codes_to_bits(self.kind, codes)
            ok = np.fromiter((1 if self.pred(codec.bits_to_value(self.kind, b)) else 0 for b in bits),
                             dtype=np.uint8, count=len(bits))
            csum = np.concatenate([[0], np.cumsum(ok, dtype=np.int64)])
            m = (csum[offsets[1:].astype(np.int64)] - csum[offsets[:-1].astype(np.int64)]) > 0  # any value passes
        else:
            bits = _codes_to_bits(self.kind, h)
            m = np.fromiter((bool(self.pred(codec.bits_to_value(self.kind, b))) for b in bits), dtype=bool,
                            count=len(bits))
        buf = np.packbits(m.astype(np.uint8), bitorder="little")
        return Docset(F.DOCSET_BITSET, buf, len(buf))


class SegmentHostView:
    """What the generic post_filter closure sees instead of tantivy readers: `.get(field, doc)` /
    `.get_vals(field, doc)` in natural value types."""

    def __init__(self, seg):
        self.seg = seg

    def get(self, field, doc):
        kind, _ = self.seg.kinds[int(field)]
        from .agg import _codes_to_bits
        return codec.bits_to_value(kind, _codes_to_bits(kind, self.seg.host[int(field)][doc:doc + 1])[0])

    def get_vals(self, field, doc):
        kind, _ = self.seg.kinds[int(field)]
        from .agg import _codes_to_bits
        offsets, codes = self.seg.host[int(field)]
        return [codec.bits_to_value(kind, b) for b in _codes_to_bits(kind, codes[int(offsets[doc]):int(offsets[doc + 1])])]


class HostClosureQuery(Query):
    def __init__(self, fetcher, filt):
        self.fetcher, self.filt = fetcher, filt

    def docset(self, seg):
        readers = self.fetcher(SegmentHostView(seg))
        m = np.fromiter((bool(self.filt(readers, d)) for d in range(seg.max_doc)), dtype=bool, count=seg.max_doc)
        buf = np.packbits(m.astype(np.uint8), bitorder="little")
        return Docset(F.DOCSET_BITSET, buf, len(buf))


# ---- result reader --------------------------------------------------------------------------
class ResultReader:
    """Typed view over a tagg_result (the `read_fruit` half of the facade)."""

    def __init__(self, handle):
        self._h = handle
        self._scopes = {}
        self._metrics = {}
        self._children = {}

    def scope(self, node):
        if node not in self._scopes:
            n = C.c_uint64()
            F.check(F.lib().tagg_result_scope_len(self._h, node, C.byref(n)))
            keys = np.empty(n.value, dtype=np.uint64)
            parents = np.empty(n.value, dtype=np.uint32)
            F.check(F.lib().tagg_result_scope_read(self._h, node, _ptr(keys), _ptr(parents), n.value))
            self._scopes[node] = (keys, parents)
        return self._scopes[node]

    @staticmethod
    def _view(ptr, n, dtype):
        if not n or not ptr:
            return np.zeros(0, dtype=dtype)
        ctype = {np.uint64: C.c_uint64, np.uint32: C.c_uint32, np.uint8: C.c_uint8}[dtype]
        return np.ctypeslib.as_array((ctype * n).from_address(ptr))

    def scope_view(self, node):
        """(keys, parents) as zero-copy numpy views of the result's page-locked image (valid while the reader lives)."""
        k, p, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        F.check(F.lib().tagg_result_scope_view(self._h, node, C.byref(k), C.byref(p), C.byref(n)))
        return self._view(k.value, n.value, np.uint64), self._view(p.value, n.value, np.uint32)

    def metric_view(self, node):
        v, s, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        F.check(F.lib().tagg_result_metric_view(self._h, node, C.byref(v), C.byref(s), C.byref(n)))
        return self._view(v.value, n.value, np.uint64), self._view(s.value, n.value, np.uint8)

    def top_k(self, scope_node, by_node, k, parent_bucket=0):
        """Bucket indices of `Terms::top_k(k, |b| b.<leaf by_node>)` (terms.rs:425-457), selected on the device."""
        out = np.zeros(max(int(k), 1), dtype=np.uint32)
        n = C.c_uint64()
        F.check(F.lib().tagg_result_top_k(self._h, int(scope_node), int(parent_bucket), int(by_node), int(k), _ptr(out), C.byref(n)))
        return out[:n.value]

    def scope_rows(self, node, buckets):
        buckets = np.ascontiguousarray(buckets, dtype=np.uint32)
        keys = np.zeros(len(buckets), dtype=np.uint64)
        parents = np.zeros(len(buckets), dtype=np.uint32)
        F.check(F.lib().tagg_result_scope_rows(self._h, int(node), _ptr(buckets) if len(buckets) else None, len(buckets), _ptr(keys), _ptr(parents)))
        return keys, parents

    def metric_rows(self, node, buckets):
        buckets = np.ascontiguousarray(buckets, dtype=np.uint32)
        values = np.zeros(len(buckets), dtype=np.uint64)
        seen = np.zeros(len(buckets), dtype=np.uint8)
        F.check(F.lib().tagg_result_metric_rows(self._h, int(node), _ptr(buckets) if len(buckets) else None, len(buckets), _ptr(values), _ptr(seen)))
        return values, seen

    def is_local(self):
        """False on the non-root ranks of a tagg_execute_reduce call: the fruit lives on the root."""
        out = C.c_int()
        F.check(F.lib().tagg_result_is_local(self._h, C.byref(out)))
        return bool(out.value)

    def scope_children(self, node, parent_bucket):
        """(keys, bucket indices) of the buckets of scope `node` whose parent bucket is `parent_bucket`."""
        if node not in self._children:
            keys, parents = self.scope(node)
            order = np.argsort(parents, kind="stable")
            sp = parents[order]
            self._children[node] = (order, sp)
        order, sp = self._children[node]
        keys, _ = self.scope(node)
        lo = np.searchsorted(sp, parent_bucket, side="left")
        hi = np.searchsorted(sp, parent_bucket, side="right")
        idx = order[lo:hi]
        return keys[idx].tolist(), idx.tolist()

    def metric(self, node):
        if node not in self._metrics:
            n = self._metric_len(node)
            values = np.empty(n, dtype=np.uint64)
            seen = np.empty(n, dtype=np.uint8)
            F.check(F.lib().tagg_result_metric_read(self._h, node, _ptr(values), _ptr(seen), n))
            self._metrics[node] = (values, seen)
        return self._metrics[node]

    def _metric_len(self, node):
        n = C.c_uint64()
        F.check(F.lib().tagg_result_metric_len(self._h, node, C.byref(n)))
        return n.value

    def percentiles(self, node, bucket):
        nt, npairs = C.c_uint64(), C.c_uint64()
        F.check(F.lib().tagg_result_percentiles_len(self._h, node, bucket, C.byref(nt), C.byref(npairs)))
        ranks = np.zeros(npairs.value, dtype=np.uint64)
        bits = np.zeros(npairs.value, dtype=np.uint64)
        F.check(F.lib().tagg_result_percentiles_read(self._h, node, bucket, _ptr(ranks), _ptr(bits), npairs.value))
        return nt.value, ranks, bits

    def stats(self):
        ms, by = C.c_double(), C.c_uint64()
        nl, path = C.c_uint32(), C.c_uint32()
        F.check(F.lib().tagg_result_stats(self._h, C.byref(ms), C.byref(by), C.byref(nl), C.byref(path)))
        return dict(kernel_ms=ms.value, alg_bytes=by.value, n_launches=nl.value, path=path.value)

    def free(self):
        if self._h:
            F.lib().tagg_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _RowsReader:
    """A reader over k gathered rows of one scope: lets `Agg.decode` build the sub-fruits of just those buckets."""

    def __init__(self, reader, buckets):
        self.reader, self.buckets = reader, buckets
        self._metrics = {}

    def metric(self, node):
        if node not in self._metrics:
            self._metrics[node] = self.reader.metric_rows(node, self.buckets)
        return self._metrics[node]

    def scope_children(self, node, parent_bucket):
        raise NotImplementedError("top_k on the device decodes buckets whose sub-aggregation is count / sum / min / max leaves")

    def percentiles(self, node, bucket):
        raise NotImplementedError("top_k on the device decodes buckets whose sub-aggregation is count / sum / min / max leaves")


class Plan:
    """A lowered aggregation tree resident as a tagg_plan (the `PreparedAgg`, src/agg.rs:19-28)."""

    def __init__(self, ctx, agg, searcher=None):
        self.agg = as_agg(agg)
        self.lctx = LowerCtx(searcher)
        self.agg.lower(self.lctx)
        nodes = (F.Node * len(self.lctx.nodes))(*self.lctx.nodes)
        self._blob_bufs = [np.frombuffer(b, dtype=np.uint8) for b in self.lctx.blobs]
        blobs = (F.Blob * max(1, len(self._blob_bufs)))()
        for i, b in enumerate(self._blob_bufs):
            blobs[i].data = _ptr(b)
            blobs[i].len = len(b)
        self.nodes = nodes
        self._h = C.c_void_p()
        F.check(F.lib().tagg_plan_create(ctx._h, nodes, len(self.lctx.nodes), blobs, len(self._blob_bufs),
                                         C.byref(self._h)))

    @property
    def filters(self):
        return self.lctx.filters

    def set_readout(self, mode):
        F.check(F.lib().tagg_plan_set_readout(self._h, int(mode)))

    def close(self):
        if self._h:
            F.lib().tagg_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def build_inputs(plan, query, segments):
    """tagg_segment_input[] for `segments` (+ the python objects that must outlive the call)."""
    keep = []
    arr = (F.SegmentInput * max(1, len(segments)))()
    for i, seg in enumerate(segments):
        ds = query.docset(seg)
        keep.append(ds)
        arr[i].segment = seg._h
        arr[i].docset = ds.c
        fl = (F.Docset * max(1, len(plan.filters)))()
        for j, fq in enumerate(plan.filters):
            fd = fq.docset(seg)
            keep.append(fd)
            fl[j] = fd.c
        keep.append(fl)
        arr[i].filters = fl
        arr[i].n_filters = len(plan.filters)
    return arr, keep


class Searcher:
    """`impl AggSearcher for tantivy::Searcher` (src/searcher.rs:53-101)."""

    def __init__(self, ctx, segments):
        self.ctx = ctx
        self.segments = list(segments)
        for i, s in enumerate(self.segments):
            s.ord = i

    def segment_readers(self):
        return self.segments

    def code_domain(self, field, multi):
        """[min code, max code] of a column over all segments (for LUT lowering); None if empty."""
        lo, hi = None, None
        for s in self.segments:
            try:
                info = s.column_info(field, 0)
            except F.TaggError:
                return None
            if info["n_values"] == 0:
                continue
            a, b = info["min_value"], info["min_value"] + info["amplitude"]
            lo = a if lo is None else min(lo, a)
            hi = b if hi is None else max(hi, b)
        return None if lo is None else (lo, hi)

    def prepare(self, agg):
        return Plan(self.ctx, agg, self)

    def terms_top_k(self, query, agg, terms, by, k):
        """agg_search followed by `Terms::top_k(k, |b| <leaf `by`>)` on the fruit of the top-level `terms` node of `agg`
        (terms.rs:425-457) — with the selection on the GPU and only the k winning buckets read back (lazy read-out).
        `terms`: the TermsAgg inside `agg`; `by`: one of its leaf sub-aggregations (count / sum / min / max).
        Returns [(key, sub fruit)] in the reference's order: sort value descending, key ascending."""
        plan = self.prepare(agg)
        plan.set_readout(F.READOUT_LAZY)
        arr, keep = build_inputs(plan, query, self.segments)
        h = C.c_void_p()
        F.check(F.lib().tagg_execute(plan._h, arr, len(self.segments), C.byref(h)))
        reader = ResultReader(h)
        buckets = reader.top_k(terms.node, by.node, k)
        keys, _ = reader.scope_rows(terms.node, buckets)
        rows = _RowsReader(reader, buckets)
        out = []
        for i, kb in enumerate(keys.tolist()):
            key = codec.bits_to_value(terms.kind, kb)
            if terms.key_filter is not None and not terms.key_filter(key):
                continue
            out.append((key, terms.sub.decode(rows, i)))
        return out

    def agg_search(self, query, agg):
        """src/searcher.rs:13-17 — default executor is SingleThread."""
        return self.agg_search_with_executor(query, agg, SINGLE_THREAD)

    DOC_ADDR_FIELD = 0xFFFFFFF0  # field id of the generated address column (segment ordinal << 32 | doc id)

    def top_hits_f64(self, query, field, k, descending=True):
        """top_hits (reference README.md:31-45, TODO there): the k matched documents with the largest (smallest) values of an
        f64 fast field, as [(value, segment ordinal, doc id)] — ties in document order.  Two passes of existing kernels, no
        per-document host work: (1) percentiles_agg_f64(field) — its fruit is a list of EXACT order statistics (rank,
        value), so the stored pair next below rank n - k + 1 bounds the k-th largest value from below; (2)
        post_filter_agg_f64(field >= bound, terms_agg_u64(address column, max_agg_f64(field))) — the few documents at or
        above the bound, keyed by their address, in the open-addressing bucket table."""
        from .agg import ge, le, max_agg_f64, percentiles_agg_f64, post_filter_agg_f64, terms_agg_u64
        if k <= 0:
            return []
        for i, seg in enumerate(self.segments):
            if getattr(seg, "_doc_addr", None) != (self.DOC_ADDR_FIELD, i << 32):
                seg.add_doc_address_column(self.DOC_ADDR_FIELD, i << 32)
        p = self.agg_search(query, percentiles_agg_f64(field))
        if p.n == 0:
            return []
        if descending:
            bound = float("-inf")
            for r, v in zip(p.ranks, p.values):
                if r > p.n - k + 1:
                    break
                bound = v
            pred = ge(bound)
        else:
            bound = float("inf")
            for r, v in zip(reversed(p.ranks), reversed(p.values)):
                if r < k:
                    break
                bound = v
            pred = le(bound)
        t = self.agg_search(query, post_filter_agg_f64(field, pred, terms_agg_u64(self.DOC_ADDR_FIELD, max_agg_f64(field))))
        hits = sorted(((v, a >> 32, a & 0xFFFFFFFF) for a, v in t.res.items()), key=lambda h: (-h[0] if descending else h[0], h[1], h[2]))
        return hits[:k]

    def agg_search_with_executor(self, query, agg, executor, collective=False, return_reader=False, root=None):
        """collective=True: every rank folds its own segments, one NCCL merge step, every rank returns the merged fruit;
        with root=r only rank r does (tagg_execute_reduce) and the other ranks return None."""
        plan = agg if isinstance(agg, Plan) else self.prepare(agg)
        lib = F.lib()
        if collective and root is not None:
            run = lambda p, a, n, out: lib.tagg_execute_reduce(p, a, n, int(root), out)
        else:
            run = lib.tagg_execute_collective if collective else lib.tagg_execute
        if executor == SINGLE_THREAD:
            # one harvest threaded through every segment (searcher.rs:66-78)
            arr, keep = build_inputs(plan, query, self.segments)
            h = C.c_void_p()
            F.check(run(plan._h, arr, len(self.segments), C.byref(h)))
            reader = ResultReader(h)
            if not reader.is_local():
                return (None, reader) if return_reader else None
        else:
            # a fruit per segment, merged in segment order (searcher.rs:79-98)
            if collective:
                raise ValueError("collective execution folds all local segments in one call")
            reader = None
            for seg in self.segments:
                arr, keep = build_inputs(plan, query, [seg])
                h = C.c_void_p()
                F.check(lib.tagg_execute(plan._h, arr, 1, C.byref(h)))
                if reader is None:
                    reader = ResultReader(h)
                else:
                    F.check(lib.tagg_result_merge(reader._h, h))
                    lib.tagg_result_free(h)
            if reader is None:  # no segments: an empty harvest (create_fruit)
                arr, keep = build_inputs(plan, query, [])
                h = C.c_void_p()
                F.check(lib.tagg_execute(plan._h, arr, 0, C.byref(h)))
                reader = ResultReader(h)
        fruit = plan.agg.decode(reader, 0)
        if return_reader:
            return fruit, reader
        return fruit
