#!/usr/bin/env python
"""bench.py — matched docs aggregated / second on BASELINE.json's config C2:

    filter_agg(TermQuery status=0, (count_agg, terms_agg_u64(category_id, (count_agg, min_agg_f64(price)))))
    over AllQuery on a 100M-doc synthetic index (8 segments x 12.5M docs), 10k categories, 25 % selectivity.

A "step" is one pass of the hot path (one agg_search) over the whole index.  At N>1 every rank holds
its own 100M-doc shard (weak scaling) and each step ends with the one exchange step of the path: the
NCCL merge of the bucket tables (tagg_execute_collective).

    value : inputs already resident in HBM (columns + the cached status=0 filter bitset)
    e2e   : the same step through the C ABI with HOST buffers — the filter bitsets are copied
            host->device from pinned memory and the result arrays device->host every step
    --impl reference : the CPU restatement of the reference's collector loop (oracle/), all host threads

One JSON line on stdout (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STATUS, CATEGORY, PRICE = 0, 1, 2
TAG_STATUS, TAG_CATEGORY, TAG_PRICE = 11, 22, 33
SEED = 1
N_CATEGORIES = int(os.environ.get("TAGG_BENCH_NCAT", "10000"))  # 10k = BASELINE config C2
METRIC = "matched docs aggregated/sec"
UNIT = "docs/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--docs", type=int, default=100_000_000, help="documents per GPU (default: the C2 size)")
    ap.add_argument("--segments", type=int, default=8)
    ap.add_argument("--cpu-sample-segments", type=int, default=8, help="segments of the workload the CPU baseline runs on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 force generic kernel, 2 force streaming kernel")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"],
                    help="c2 (default, the bench contract's workload, weak scaling) | c5: BASELINE configs[4], 1B docs in 64 segments "
                         "sharded over the ranks, post_filter + terms + nested min/max/sum, NCCL bucket merge (strong scaling)")
    return ap.parse_args()


def workload_name(args):
    return (f"C2 filter_agg(status=0,(count,terms_u64(category,(count,min_f64 price)))) AllQuery, "
            f"{args.docs} docs/GPU in {args.segments} segments, {N_CATEGORIES} categories, 25% selectivity")


# ---------------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md): sampled DURING the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None
        time.sleep(0.5)  # let the first samples land before the timed region starts

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference's collector loop (the reference itself is Rust with
# un-vendored dependencies and cannot be built in this image — DESIGN.md)
# ---------------------------------------------------------------------------------------------------
def cpu_run(args, n_segments, doc_base0, reps, threads):
    """Builds `n_segments` of the workload on the host (same counter-based recipe as the device
    generator) and times agg_search on them.  Returns (best docs/s single-thread, best docs/s
    thread-pool, docs per run)."""
    import tantivy_aggregations_b200 as ta
    from oracle import oracle
    from tantivy_aggregations_b200 import _ffi as F
    per_seg = args.docs // args.segments
    ix = oracle.OracleIndex()
    bits = {}
    for s in range(n_segments):
        base = doc_base0 + s * per_seg
        o = ix.add_segment(per_seg)
        st = oracle.synth_codes(1, SEED, TAG_STATUS, base, per_seg, 0, 4)
        ix.set_column_codes(o, STATUS, F.U64, st)
        ix.set_column_codes(o, CATEGORY, F.U64, oracle.synth_codes(1, SEED, TAG_CATEGORY, base, per_seg, 1, N_CATEGORIES))
        ix.set_column_codes(o, PRICE, F.F64, oracle.synth_codes(0, SEED, TAG_PRICE, base, per_seg))
        bits[o] = np.packbits((st == 0).astype(np.uint8), bitorder="little")  # TermQuery status=0 postings as a bitset
        ix.segs[o].host.clear()
    agg = lambda: ta.filter_agg(ta.BitsetQuery(bits), (ta.count_agg(), ta.terms_agg_u64(CATEGORY, (ta.count_agg(), ta.min_agg_f64(PRICE)))))
    docs = per_seg * n_segments
    best = {}
    for mode, thr in ((0, 1), (1, threads)):
        t_best = None
        for _ in range(reps):
            _, sec, collected = ix.search(ta.AllQuery(), agg(), mode=mode, threads=thr, decode=False)
            assert collected == docs
            t_best = sec if t_best is None else min(t_best, sec)
        best[mode] = docs / t_best
    return best[0], best[1], docs, ix, agg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    nseg = min(args.cpu_sample_segments, args.segments)
    per_seg = args.docs // args.segments
    import tantivy_aggregations_b200 as ta
    _, _, docs, ix, agg = cpu_run(args, nseg, 0, 1, threads)
    for _ in range(args.warmup):
        ix.search(ta.AllQuery(), agg(), mode=1, threads=threads, decode=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ix.search(ta.AllQuery(), agg(), mode=1, threads=threads, decode=False)
    dt = time.perf_counter() - t0
    value = docs * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64/f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": f"{nseg} of {args.segments} segments ({docs} docs) per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{nseg} of {args.segments} segments ({docs} docs), Executor::ThreadPool shape, {threads} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement of the reference collector loop (oracle/oracle.cpp); the Rust reference cannot be built in this image",
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
def read_result_arrays(reader, plan_nodes):
    """What a host facade reads back: bucket keys + every metric array (numpy, no python dicts)."""
    import tantivy_aggregations_b200 as ta  # noqa
    nbytes = 0
    keys, parents = reader.scope(plan_nodes["terms"])
    nbytes += keys.nbytes + parents.nbytes
    out = {"keys": keys}
    for name in ("root_count", "bucket_count", "bucket_min"):
        v, s = reader.metric(plan_nodes[name])
        nbytes += v.nbytes + s.nbytes
        out[name] = (v, s)
    return out, nbytes


def run_b200(args):
    import torch
    import tantivy_aggregations_b200 as ta
    from tantivy_aggregations_b200 import _ffi as F
    from tantivy_aggregations_b200 import index as I

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = ta.Context(local_rank)
    ctx.set_path(args.path)
    if world > 1:
        ident = [ta.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        ctx.comm_init(ident[0], rank, world)

    # ---- the index: synthetic columns generated in HBM (SURVEY §8d recipe), one shard per rank ----
    per_seg = args.docs // args.segments
    segments = []
    for s in range(args.segments):
        base = rank * args.docs + s * per_seg
        seg = ta.Segment(ctx, per_seg, keep_host=False)
        seg.synth_column(STATUS, ta.U64, 1, SEED, TAG_STATUS, base, 0, 4)
        seg.synth_column(CATEGORY, ta.U64, 1, SEED, TAG_CATEGORY, base, 1, N_CATEGORIES)
        seg.synth_column(PRICE, ta.F64, 0, SEED, TAG_PRICE, base)
        segments.append(seg)
    searcher = ta.Searcher(ctx, segments)
    status_q = ta.TermQuery(STATUS, ta.U64, 0)

    # the filter query's matched docs: on the host as pinned bitsets (what decoding the postings of
    # `status=0` yields; here evaluated once from the fast field) and cached on the device
    # (slices of ONE page-locked buffer at a 256-byte stride: the library reads page-locked, 16-byte aligned bitsets in
    #  place — the TMA producer of the streaming kernel pulls every tile over PCIe once — and otherwise moves equal-sized
    #  strided slices as one 2-D copy)
    host_bits = {}
    need = (per_seg + 7) // 8
    stride = (need + 255) // 256 * 256
    pinned_all = torch.zeros(stride * len(segments), dtype=torch.uint8).pin_memory()
    for i, seg in enumerate(segments):
        b = seg.docset_to_bitset(status_q.docset(seg))
        assert len(b) == need
        pinned = pinned_all[i * stride:(i + 1) * stride]
        pinned.numpy()[:need] = b
        host_bits[seg.ord] = pinned
    host_filter = ta.BitsetQuery({k: v.numpy() for k, v in host_bits.items()})
    dev_filter = ta.CachedQuery(host_filter, segments)

    def make_plan(fq):
        agg = ta.filter_agg(fq, (ta.count_agg(), ta.terms_agg_u64(CATEGORY, (ta.count_agg(), ta.min_agg_f64(PRICE)))))
        plan = searcher.prepare(agg)
        nodes = {"terms": agg.sub.members[1].node, "root_count": agg.sub.members[0].node,
                 "bucket_count": agg.sub.members[1].sub.members[0].node, "bucket_min": agg.sub.members[1].sub.members[1].node}
        return plan, nodes

    plan_dev, nodes = make_plan(dev_filter)
    plan_host, _ = make_plan(host_filter)
    allq = ta.AllQuery()
    lib = F.lib()
    run = lib.tagg_execute_collective if world > 1 else lib.tagg_execute
    docs_per_step = per_seg * args.segments

    def step(plan, read=True):
        arr, keep = I.build_inputs(plan, allq, segments)
        h = C.c_void_p()
        F.check(run(plan._h, arr, len(segments), C.byref(h)))
        reader = I.ResultReader(h)
        out, nbytes = read_result_arrays(reader, nodes) if read else (None, 0)
        st = reader.stats()
        reader.free()
        return out, nbytes, st

    def barrier():
        ctx.synchronize()
        if dist is not None:
            dist.barrier()
        ctx.synchronize()

    def timed(plan):
        for _ in range(args.warmup):
            step(plan)
        barrier()
        ctx.timer_start()
        t0 = time.perf_counter()
        kernel_ms, alg_bytes, launches, d2h, path = 0.0, 0, 0, 0, 0
        for _ in range(args.steps):
            _, nbytes, st = step(plan)
            kernel_ms += st["kernel_ms"]
            alg_bytes = st["alg_bytes"]
            launches += st["n_launches"]
            d2h = nbytes
            path = st["path"]
        dev_ms = ctx.timer_stop()
        wall_ms = 1e3 * (time.perf_counter() - t0)
        barrier()
        ms = max(dev_ms, 0.0)
        if dist is not None:
            t = torch.tensor([ms, wall_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall_ms = float(t[0]), float(t[1])
        return dict(ms=ms, wall_ms=wall_ms, kernel_ms=kernel_ms / args.steps, alg_bytes=alg_bytes, launches=launches,
                    d2h=d2h, path=path)

    # ---- correctness gate before timing: size-independent properties + a sampled oracle check ----
    out, _, st0 = step(plan_dev)
    root_count = int(out["root_count"][0][0])
    bucket_counts = out["bucket_count"][0]
    n_match_host = sum(int(np.unpackbits(v.numpy()[:need], bitorder="little")[:per_seg].sum()) for v in host_bits.values())
    if world == 1:
        assert root_count == n_match_host, (root_count, n_match_host)
    assert int(bucket_counts.sum()) == root_count, "sum of bucket counts != filtered count"
    assert len(out["keys"]) <= N_CATEGORIES and out["keys"].min() >= 1 and out["keys"].max() <= N_CATEGORIES

    sampler = ClockSampler(local_rank)
    res_dev = timed(plan_dev)
    res_e2e = timed(plan_host)
    clocks = sampler.stop()

    value = docs_per_step * world * args.steps / (res_dev["ms"] * 1e-3)
    e2e_value = docs_per_step * world * args.steps / (res_e2e["ms"] * 1e-3)
    h2d = need * len(host_bits)

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = res_dev["alg_bytes"] / (res_dev["kernel_ms"] * 1e-3) / 1e9 if res_dev["kernel_ms"] > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("k_stream_dram_bytes_per_launch")
    except OSError:
        pass

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        nseg = min(args.cpu_sample_segments, args.segments)
        single, pool, docs, ix, agg = cpu_run(args, nseg, 0, 2, threads)
        # parity on the sample: the same segments through the GPU path must equal the oracle bit for bit
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import assert_fruit_equal
        want, _, _ = ix.search(ta.AllQuery(), agg())
        sub = ta.Searcher(ctx, segments[:nseg])
        got = sub.agg_search(allq, ta.filter_agg(status_q, (ta.count_agg(), ta.terms_agg_u64(CATEGORY, (ta.count_agg(), ta.min_agg_f64(PRICE))))))
        for i, s in enumerate(segments):
            s.ord = i
        assert_fruit_equal(got, want)
        cpu_baseline = {"value": pool, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{nseg} of {args.segments} segments ({docs} docs); Executor::ThreadPool shape with {threads} threads; "
                                  f"Executor::SingleThread (agg_search default) = {single:.4g} docs/s; GPU result on the sample == oracle",
                        "single_thread_value": single}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res_dev["ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64/f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "l2": "inputs (875 MB/step) are larger than the 126 MB L2",
                       "timing": "CUDA events on the execute stream around the K steps, max over ranks",
                       "path": {0: "none", 1: "generic", 2: "stream"}.get(res_dev["path"], "?"),
                       "multi_gpu": "one 100M-doc shard per rank; NCCL all-reduce of the bucket tables every step" if world > 1 else "single GPU"},
            "kernel_ms_per_step": res_dev["kernel_ms"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "traffic": traffic, "kernel": "k_stream",
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "algorithmic_bytes_per_launch": res_dev["alg_bytes"],
                         "frac_of_8TBs_contract": achieved / 8000.0},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": res_e2e["d2h"],
                    "ms_per_step": res_e2e["ms"] / args.steps, "kernel_ms_per_step": res_e2e["kernel_ms"]},
            "gpu_launches": res_dev["launches"] + res_e2e["launches"],
            "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"]},
            "cpu_baseline": cpu_baseline,
            "matched_docs_per_step": docs_per_step * world, "filtered_docs_per_step": root_count,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
# BASELINE configs[4] (C5): 1B docs / 64 segments sharded over the ranks, strong scaling.  Not the bench
# contract's workload (that is C2 above); run by hand for the scaling table in DESIGN.md / profiles/.
# ---------------------------------------------------------------------------------------------------
def run_c5(args):
    import torch
    import tantivy_aggregations_b200 as ta
    from tantivy_aggregations_b200 import _ffi as F
    from tantivy_aggregations_b200 import index as I

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = ta.Context(local_rank)
    if world > 1:
        ident = [ta.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        ctx.comm_init(ident[0], rank, world)
    n_total, n_seg, n_cat = 1_000_000_000, 64, 100_000
    per_seg = n_total // n_seg
    mine = ta.assign_segments([per_seg] * n_seg, world)[rank]
    segments = []
    for s in mine:
        seg = ta.Segment(ctx, per_seg, keep_host=False)
        seg.synth_column(STATUS, ta.U64, 1, SEED, TAG_STATUS, s * per_seg, 0, 4)
        seg.synth_column(CATEGORY, ta.U64, 1, SEED, TAG_CATEGORY, s * per_seg, 1, n_cat)
        seg.synth_column(PRICE, ta.F64, 0, SEED, TAG_PRICE, s * per_seg)
        segments.append(seg)
    searcher = ta.Searcher(ctx, segments)
    agg = ta.post_filter_agg_u64(STATUS, ta.eq(0), ta.terms_agg_u64(CATEGORY, (ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE), ta.sum_agg_f64(PRICE))))
    plan = searcher.prepare(agg)
    terms = agg.sub
    nodes = [m.node for m in terms.sub.members]
    allq = ta.AllQuery()
    lib = F.lib()
    run = lib.tagg_execute_collective if world > 1 else lib.tagg_execute

    def step():
        arr, keep = I.build_inputs(plan, allq, segments)
        h = C.c_void_p()
        F.check(run(plan._h, arr, len(segments), C.byref(h)))
        reader = I.ResultReader(h)
        keys, parents = reader.scope(terms.node)
        nbytes = keys.nbytes + parents.nbytes
        vals = []
        for nd in nodes:
            v, sflag = reader.metric(nd)
            nbytes += v.nbytes + sflag.nbytes
            vals.append(v)
        st = reader.stats()
        reader.free()
        return keys, vals, nbytes, st

    def barrier():
        ctx.synchronize()
        if dist is not None:
            dist.barrier()
        ctx.synchronize()

    keys, vals, _, _ = step()
    assert len(keys) == n_cat, len(keys)
    mn, mx = vals[0].view(np.float64), vals[1].view(np.float64)
    assert (mn >= 1.0).all() and (mx < 101.0).all() and (mn <= mx).all()
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    barrier()
    ctx.timer_start()
    kernel_ms, launches = 0.0, 0
    for _ in range(args.steps):
        _, _, nbytes, st = step()
        kernel_ms += st["kernel_ms"]
        launches += st["n_launches"]
    ms = ctx.timer_stop()
    barrier()
    clocks = sampler.stop()
    alg = st["alg_bytes"]
    if dist is not None:
        t = torch.tensor([ms, kernel_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, kernel_ms = float(t[0]), float(t[1])
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = alg / (kernel_ms / args.steps * 1e-3) / 1e9
        print(json.dumps({
            "metric": METRIC, "value": n_total * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64/f64", "data": "synthetic",
            "config": {"workload": "C5 post_filter_agg_u64(status==0, terms_u64(category 100k, (min,max,sum f64 price))) AllQuery, 1e9 docs in 64 segments "
                                   f"sharded over {world} GPU(s)", "l2": "inputs (>= 1.16 GB per GPU and step) are larger than the 126 MB L2",
                       "multi_gpu": "bucket tables merged by NCCL every step" if world > 1 else "single GPU"},
            "kernel_ms_per_step": kernel_ms / args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "kernel": "k_stream", "algorithmic_bytes_per_launch": alg, "note": "per GPU (slowest rank)"},
            "e2e": {"value": n_total * args.steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": nbytes,
                    "note": "the query has no host docset (post_filter on a fast field); result arrays are read back every step"},
            "gpu_launches": launches, "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"]},
        }))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5":
        run_c5(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
