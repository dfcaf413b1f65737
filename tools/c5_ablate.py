import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tantivy_aggregations_b200 as ta
SEED=1; STATUS, CATEGORY, PRICE = 0, 1, 2
ctx = ta.Context(0)
n, nseg = 1_000_000_000, 64
segs = []
for s in range(nseg):
    seg = ta.Segment(ctx, n // nseg, keep_host=False)
    b = s * (n // nseg)
    seg.synth_column(STATUS, ta.U64, 1, SEED, 11, b, 0, 4)
    seg.synth_column(CATEGORY, ta.U64, 1, SEED, 22, b, 1, 100_000)
    seg.synth_column(PRICE, ta.F64, 0, SEED, 33, b)
    segs.append(seg)
S = ta.Searcher(ctx, segs)
def run(name, mk):
    plan = S.prepare(mk()); best = None
    for _ in range(4):
        _, r = S.agg_search_with_executor(ta.AllQuery(), plan, ta.SINGLE_THREAD, return_reader=True)
        st = r.stats(); best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
    print(f"{name:40s} kernel={best['kernel_ms']:.3f} ms path={best['path']}", flush=True)
pf = lambda sub: ta.post_filter_agg_u64(STATUS, ta.eq(0), sub)
run("min,max,sum", lambda: pf(ta.terms_agg_u64(CATEGORY, (ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE), ta.sum_agg_f64(PRICE)))))
run("sum only", lambda: pf(ta.terms_agg_u64(CATEGORY, ta.sum_agg_f64(PRICE))))
run("min,max", lambda: pf(ta.terms_agg_u64(CATEGORY, (ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE)))))
run("max only", lambda: pf(ta.terms_agg_u64(CATEGORY, ta.max_agg_f64(PRICE))))
run("count only", lambda: pf(ta.terms_agg_u64(CATEGORY, ta.count_agg())))
run("root count only", lambda: pf(ta.count_agg()))
run("root sum", lambda: pf(ta.sum_agg_f64(PRICE)))
import numpy as np
run("AllQuery count (nothing staged)", lambda: ta.count_agg())
def runq(name, q, mk):
    plan = S.prepare(mk()); best = None
    for _ in range(4):
        _, r = S.agg_search_with_executor(q, plan, ta.SINGLE_THREAD, return_reader=True)
        st = r.stats(); best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
    print(f"{name:40s} kernel={best['kernel_ms']:.3f} ms path={best['path']} launches={best['n_launches']}", flush=True)
rng = np.random.default_rng(3)
bq = ta.CachedQuery(ta.BitsetQuery({i: rng.integers(0, 256, size=(s.max_doc + 7) // 8, dtype=np.uint8) for i, s in enumerate(segs)}), segs)
runq("bitset 50% count", bq, lambda: ta.count_agg())
runq("AllQuery sum(price) dense", ta.AllQuery(), lambda: ta.sum_agg_f64(PRICE))
runq("AllQuery count+sum+min+max dense (C1x)", ta.AllQuery(), lambda: (ta.count_agg(), ta.sum_agg_f64(PRICE), ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE)))
runq("AllQuery max(status) dense 2-bit", ta.AllQuery(), lambda: ta.max_agg_u64(STATUS))
