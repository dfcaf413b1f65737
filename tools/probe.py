#!/usr/bin/env python
"""Kernel-time probe: runs a set of aggregation shapes over the bench index and prints kernel ms / GB/s."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tantivy_aggregations_b200 as ta
STATUS, CATEGORY, PRICE = 0, 1, 2
docs = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
nseg = 8
ctx = ta.Context(0)
segs = []
for s in range(nseg):
    seg = ta.Segment(ctx, docs // nseg, keep_host=False)
    base = s * (docs // nseg)
    seg.synth_column(STATUS, ta.U64, 1, 1, 11, base, 0, 4)
    seg.synth_column(CATEGORY, ta.U64, 1, 1, 22, base, 1, 10000)
    seg.synth_column(PRICE, ta.F64, 0, 1, 33, base)
    segs.append(seg)
S = ta.Searcher(ctx, segs)
fq = ta.CachedQuery(ta.TermQuery(STATUS, ta.U64, 0), segs)
shapes = {
    "count(all)": (ta.AllQuery(), lambda: ta.count_agg()),
    "filter->count": (ta.AllQuery(), lambda: ta.filter_agg(fq, ta.count_agg())),
    "filter->sum_u64(cat)": (ta.AllQuery(), lambda: ta.filter_agg(fq, ta.sum_agg_u64(CATEGORY))),
    "filter->min(price)": (ta.AllQuery(), lambda: ta.filter_agg(fq, ta.min_agg_f64(PRICE))),
    "filter->terms(count)": (ta.AllQuery(), lambda: ta.filter_agg(fq, ta.terms_agg_u64(CATEGORY, ta.count_agg()))),
    "filter->terms(min)": (ta.AllQuery(), lambda: ta.filter_agg(fq, ta.terms_agg_u64(CATEGORY, ta.min_agg_f64(PRICE)))),
    "C2": (ta.AllQuery(), lambda: ta.filter_agg(fq, (ta.count_agg(), ta.terms_agg_u64(CATEGORY, (ta.count_agg(), ta.min_agg_f64(PRICE)))))),
    "C2 range-filter": (ta.AllQuery(), lambda: ta.filter_agg(ta.TermQuery(STATUS, ta.U64, 0), (ta.count_agg(), ta.terms_agg_u64(CATEGORY, (ta.count_agg(), ta.min_agg_f64(PRICE)))))),
    "C1 all->(count,sum,min,max price)": (ta.AllQuery(), lambda: (ta.count_agg(), ta.sum_agg_f64(PRICE), ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE))),
    "all->terms(count)": (ta.AllQuery(), lambda: ta.terms_agg_u64(CATEGORY, ta.count_agg())),
    "all->terms(min,max,sum)": (ta.AllQuery(), lambda: ta.terms_agg_u64(CATEGORY, (ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE), ta.sum_agg_f64(PRICE)))),
    "all->hist(price,0,10,count)": (ta.AllQuery(), lambda: ta.histogram_agg_f64(PRICE, 0.0, 10.0, ta.count_agg())),
}
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
for name, (q, mk) in shapes.items():
    if only and not any(o in name for o in only): continue
    plan = S.prepare(mk())
    best = None
    for i in range(4):
        _, r = S.agg_search_with_executor(q, plan, ta.SINGLE_THREAD, return_reader=True)
        st = r.stats()
        best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
    print(f"{name:40s} path={best['path']} kernel={best['kernel_ms']:.3f} ms  alg={best['alg_bytes']/1e6:.0f} MB  {best['alg_bytes']/best['kernel_ms']/1e6:.0f} GB/s", flush=True)
