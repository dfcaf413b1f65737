import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tantivy_aggregations_b200 as ta
SEED=1; STATUS, CATEGORY, PRICE = 0,1,2
ctx = ta.Context(0)
n, nseg = 125_000_000, 8
segs=[]
for s in range(nseg):
    seg = ta.Segment(ctx, n//nseg, keep_host=False); b = s*(n//nseg)
    seg.synth_column(STATUS, ta.U64, 1, SEED, 11, b, 0, 4); seg.synth_column(CATEGORY, ta.U64, 1, SEED, 22, b, 1, 100_000); seg.synth_column(PRICE, ta.F64, 0, SEED, 33, b)
    segs.append(seg)
S = ta.Searcher(ctx, segs)
pf = lambda sub: ta.post_filter_agg_u64(STATUS, ta.eq(0), sub)
fq = ta.CachedQuery(ta.TermQuery(STATUS, ta.U64, 0), segs)
shapes = {
 "C5 pf->terms(min,max,sum)": (ta.AllQuery(), lambda: pf(ta.terms_agg_u64(CATEGORY, (ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE), ta.sum_agg_f64(PRICE))))),
 "pf->count": (ta.AllQuery(), lambda: pf(ta.count_agg())),
 "pf->terms(sum)": (ta.AllQuery(), lambda: pf(ta.terms_agg_u64(CATEGORY, ta.sum_agg_f64(PRICE)))),
 "pf->terms(min)": (ta.AllQuery(), lambda: pf(ta.terms_agg_u64(CATEGORY, ta.min_agg_f64(PRICE)))),
 "pf->terms(count)": (ta.AllQuery(), lambda: pf(ta.terms_agg_u64(CATEGORY, ta.count_agg()))),
 "bitset-filter->terms(min,max,sum)": (ta.AllQuery(), lambda: ta.filter_agg(fq, ta.terms_agg_u64(CATEGORY, (ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE), ta.sum_agg_f64(PRICE))))),
}
only = sys.argv[1].split(",") if len(sys.argv) > 1 else None
for name,(q,mk) in shapes.items():
    if only and not any(o in name for o in only): continue
    plan = S.prepare(mk()); best=None
    for i in range(4):
        _, r = S.agg_search_with_executor(q, plan, ta.SINGLE_THREAD, return_reader=True); st=r.stats()
        best = st if best is None or st["kernel_ms"]<best["kernel_ms"] else best
    print(f"{name:40s} path={best['path']} kernel={best['kernel_ms']:.3f} ms alg={best['alg_bytes']/1e6:.0f} MB {best['alg_bytes']/best['kernel_ms']/1e6:.0f} GB/s", flush=True)
