#!/usr/bin/env python
"""Kernel time / achieved algorithmic GB/s of BASELINE configs C1, C3, C4, C5 (C2 is bench.py) on one GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tantivy_aggregations_b200 as ta
SEED = 1
STATUS, CATEGORY, PRICE, KEYS, VALS, KEYS_SPREAD = 0, 1, 2, 3, 4, 5
ctx = ta.Context(0)
which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c1", "c1x", "c3", "c4", "c5"]

def run(name, S, q, mk, reps=4):
    plan = S.prepare(mk())
    best = None
    for _ in range(reps):
        _, r = S.agg_search_with_executor(q, plan, ta.SINGLE_THREAD, return_reader=True)
        st = r.stats()
        best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
    import time
    t0 = time.perf_counter()
    for _ in range(8):
        f = S.agg_search(q, plan)
    step_ms = (time.perf_counter() - t0) / 8 * 1e3
    print(f"{name:52s} path={best['path']} kernel={best['kernel_ms']:9.3f} ms  step={step_ms:7.3f} ms  alg={best['alg_bytes']/1e6:8.0f} MB  {best['alg_bytes']/best['kernel_ms']/1e6:7.0f} GB/s  launches={best['n_launches']}", flush=True)

def segs_of(n, nseg, cols):
    out = []
    for s in range(nseg):
        seg = ta.Segment(ctx, n // nseg, keep_host=False)
        for c in cols: c(seg, s * (n // nseg))
        out.append(seg)
    return out

price = lambda s, b: s.synth_column(PRICE, ta.F64, 0, SEED, 33, b)
c1agg = lambda: (ta.count_agg(), ta.sum_agg_f64(PRICE), ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE))
if "c1" in which:
    segs = segs_of(1_000_000, 1, [price]); run("C1 1M docs (launch-bound)", ta.Searcher(ctx, segs), ta.AllQuery(), c1agg)
if "c1x" in which:
    segs = segs_of(1_000_000_000, 64, [price]); run("C1 x1000 (1G docs)", ta.Searcher(ctx, segs), ta.AllQuery(), c1agg)
    for s in segs: s.close()
if "c2" in which:
    segs = segs_of(100_000_000, 8, [lambda s, b: s.synth_column(STATUS, ta.U64, 1, SEED, 11, b, 0, 4),
                                    lambda s, b: s.synth_column(CATEGORY, ta.U64, 1, SEED, 22, b, 1, 10_000), price])
    S = ta.Searcher(ctx, segs)
    fq = ta.CachedQuery(ta.TermQuery(STATUS, ta.U64, 0), segs)
    run("C2 filter(status=0,(count,terms(cat 10k,(count,min)))) 100M", S, ta.AllQuery(),
        lambda: ta.filter_agg(fq, (ta.count_agg(), ta.terms_agg_u64(CATEGORY, (ta.count_agg(), ta.min_agg_f64(PRICE))))))
    for s in segs: s.close()
if "c3hist" in which:
    segs = segs_of(500_000_000, 8, [price]); S = ta.Searcher(ctx, segs)
    rng = np.random.default_rng(3)
    q = ta.CachedQuery(ta.BitsetQuery({i: rng.integers(0, 256, size=(s.max_doc + 7) // 8, dtype=np.uint8) for i, s in enumerate(segs)}), segs)
    run("C3 hist(price,0,10,count) 500M/50%", S, q, lambda: ta.histogram_agg_f64(PRICE, 0.0, 10.0, ta.count_agg()))
    for s in segs: s.close()
if "c3pair" in which:  # only the fused (histogram, percentiles) tuple — the ncu target
    segs = segs_of(500_000_000, 8, [price]); S = ta.Searcher(ctx, segs)
    rng = np.random.default_rng(3)
    q = ta.CachedQuery(ta.BitsetQuery({i: rng.integers(0, 256, size=(s.max_doc + 7) // 8, dtype=np.uint8) for i, s in enumerate(segs)}), segs)
    run("C3 (hist, percentiles) 500M/50%", S, q, lambda: (ta.histogram_agg_f64(PRICE, 0.0, 10.0, ta.count_agg()), ta.percentiles_agg_f64(PRICE)), reps=3)
    for s in segs: s.close()
if "c3" in which:
    segs = segs_of(500_000_000, 8, [price]); S = ta.Searcher(ctx, segs)
    rng = np.random.default_rng(3)
    q = ta.CachedQuery(ta.BitsetQuery({i: rng.integers(0, 256, size=(s.max_doc + 7) // 8, dtype=np.uint8) for i, s in enumerate(segs)}), segs)
    run("C3 hist(price,0,10,count) 500M/50%", S, q, lambda: ta.histogram_agg_f64(PRICE, 0.0, 10.0, ta.count_agg()))
    run("C3 percentiles(price) 500M/50%", S, q, lambda: ta.percentiles_agg_f64(PRICE), reps=2)
    run("C3 (hist, percentiles) 500M/50%", S, q, lambda: (ta.histogram_agg_f64(PRICE, 0.0, 10.0, ta.count_agg()), ta.percentiles_agg_f64(PRICE)), reps=2)
    for s in segs: s.close()
if "c4" in which or "c4d" in which or "c4z" in which or "c4h" in which:
    segs = segs_of(250_000_000, 16, [lambda s, b: s.synth_multicolumn(KEYS, ta.U64, 1, SEED, 44, b, 9, 0, 1_000_000),
                                     lambda s, b: s.synth_multicolumn(KEYS_SPREAD, ta.U64, 2, SEED, 44, b, 9, 5, 1_000_000, 1 << 20),
                                     lambda s, b: s.synth_multicolumn(6, ta.U64, 3, SEED, 44, b, 9, 0, 1_000_000),
                                     lambda s, b: s.synth_multicolumn(VALS, ta.F64, 0, SEED, 55, b, 3)])
    S = ta.Searcher(ctx, segs)
    if "c4" in which or "c4d" in which:
        run("C4 terms_u64s(keys, sum_f64s(vals)) dense 1M keys", S, ta.AllQuery(), lambda: ta.terms_agg_u64s(KEYS, ta.sum_agg_f64s(VALS)), reps=3)
    if "c4" in which or "c4h" in which:
        run("C4 same, hashed spill table (40-bit key domain)", S, ta.AllQuery(), lambda: ta.terms_agg_u64s(KEYS_SPREAD, ta.sum_agg_f64s(VALS)), reps=3)
    if "c4" in which or "c4z" in which:
        run("C4 same, power-law keys (dense table + hot-key front)", S, ta.AllQuery(), lambda: ta.terms_agg_u64s(6, ta.sum_agg_f64s(VALS)), reps=3)
    for s in segs: s.close()
if "c5" in which:
    segs = segs_of(1_000_000_000, 64, [lambda s, b: s.synth_column(STATUS, ta.U64, 1, SEED, 11, b, 0, 4),
                                       lambda s, b: s.synth_column(CATEGORY, ta.U64, 1, SEED, 22, b, 1, 100_000), price])
    S = ta.Searcher(ctx, segs)
    run("C5 post_filter(status=0)->terms(cat 100k,(min,max,sum)) 1G", S, ta.AllQuery(),
        lambda: ta.post_filter_agg_u64(STATUS, ta.eq(0), ta.terms_agg_u64(CATEGORY, (ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE), ta.sum_agg_f64(PRICE)))))
    run("C5 per GPU at 8 GPUs (8 of 64 segments)", ta.Searcher(ctx, segs[:8]), ta.AllQuery(),
        lambda: ta.post_filter_agg_u64(STATUS, ta.eq(0), ta.terms_agg_u64(CATEGORY, (ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE), ta.sum_agg_f64(PRICE)))))
