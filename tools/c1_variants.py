#!/usr/bin/env python
"""C1 x1000 with fewer folds per value: does the streaming pipeline or the per-value work set the time?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tantivy_aggregations_b200 as ta
ctx = ta.Context(0)
PRICE = 2
segs = []
for s in range(64):
    seg = ta.Segment(ctx, 1_000_000_000 // 64, keep_host=False)
    seg.synth_column(PRICE, ta.F64, 0, 1, 33, s * (1_000_000_000 // 64))
    segs.append(seg)
S = ta.Searcher(ctx, segs)
for name, mk in [("min", lambda: ta.min_agg_f64(PRICE)), ("sum", lambda: ta.sum_agg_f64(PRICE)),
                 ("min,max", lambda: (ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE))),
                 ("count,sum,min,max", lambda: (ta.count_agg(), ta.sum_agg_f64(PRICE), ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE)))]:
    plan = S.prepare(mk())
    best = 1e9
    for _ in range(4):
        _, r = S.agg_search_with_executor(ta.AllQuery(), plan, ta.SINGLE_THREAD, return_reader=True)
        best = min(best, r.stats()["kernel_ms"])
    print(f"{name:20s} {best:.3f} ms  {6875/best:.0f} GB/s")
