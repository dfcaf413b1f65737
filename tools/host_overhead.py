#!/usr/bin/env python
"""Where does the per-step host time go? (python arg marshalling / tagg_execute / result readers)"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tantivy_aggregations_b200 as ta
from tantivy_aggregations_b200 import _ffi as F, index as I
docs, nseg = 100_000_000, 8
ctx = ta.Context(0)
segs = []
for s in range(nseg):
    seg = ta.Segment(ctx, docs // nseg, keep_host=False); base = s * (docs // nseg)
    seg.synth_column(0, ta.U64, 1, 1, 11, base, 0, 4); seg.synth_column(1, ta.U64, 1, 1, 22, base, 1, 10000); seg.synth_column(2, ta.F64, 0, 1, 33, base)
    segs.append(seg)
S = ta.Searcher(ctx, segs)
fq = ta.CachedQuery(ta.TermQuery(0, ta.U64, 0), segs)
agg = ta.filter_agg(fq, (ta.count_agg(), ta.terms_agg_u64(1, (ta.count_agg(), ta.min_agg_f64(2)))))
plan = S.prepare(agg)
allq = ta.AllQuery(); lib = F.lib()
T = [0.0] * 4; N = 50
for it in range(N + 5):
    t0 = time.perf_counter()
    arr, keep = I.build_inputs(plan, allq, segs)
    t1 = time.perf_counter()
    h = C.c_void_p(); F.check(lib.tagg_execute(plan._h, arr, len(segs), C.byref(h)))
    t2 = time.perf_counter()
    r = I.ResultReader(h); r.scope(agg.sub.members[1].node); r.metric(agg.sub.members[0].node); r.metric(agg.sub.members[1].sub.members[0].node); r.metric(agg.sub.members[1].sub.members[1].node)
    st = r.stats()
    t3 = time.perf_counter()
    r.free()
    t4 = time.perf_counter()
    if it >= 5:
        T[0] += t1 - t0; T[1] += t2 - t1; T[2] += t3 - t2; T[3] += st["kernel_ms"] * 1e-3
print(f"build_inputs {T[0]/N*1e6:.1f} us   tagg_execute {T[1]/N*1e6:.1f} us (kernel {T[3]/N*1e6:.1f} us)   readers {T[2]/N*1e6:.1f} us")
