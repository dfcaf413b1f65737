#!/usr/bin/env python
"""bench.py — matched docs aggregated / second on BASELINE.json's configurations.

Headline workload (the one BASELINE.json quotes "at 1/2/4/8 B200"), STRONG scaling:

    C5  post_filter_agg_u64(status == 0, terms_agg_u64(category, (min_f64 price, max_f64 price, sum_f64 price)))
        over AllQuery on a 1e9-doc synthetic index in 64 segments, 100k categories, sharded by segment over the N ranks;
        the bucket tables are merged by an NCCL reduce into rank 0, which reads the fruit (tagg_execute_reduce).

A "step" is one pass of the hot path (one agg_search) over the whole index.

    value : inputs resident in HBM (columns; AllQuery needs no docset), fruit read back on rank 0 every step
    e2e   : the same step through the C ABI with HOST buffers — the main query's matched-doc set is handed over as
            per-segment bitsets in page-locked host memory (north_star (2); all documents match), the fruit arrays are
            read back every step
    per_config (N = 1 only): kernel time / roofline / e2e / CPU baseline of all five BASELINE configurations
    --impl reference : the CPU restatement of the reference's collector loop (oracle/), all host threads, on a bounded
            sample of the same workload

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
# stdout carries ONE JSON line: keep NCCL's version banner (printed to stdout at NCCL_DEBUG=VERSION) out of it
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

STATUS, CATEGORY, PRICE, KEYS, VALS, KEYS_SPREAD, KEYS_ZIPF = 0, 1, 2, 3, 4, 5, 6
TAG_STATUS, TAG_CATEGORY, TAG_PRICE, TAG_KEYS, TAG_VALS = 11, 22, 33, 44, 55
SEED = 1
METRIC = "matched docs aggregated/sec"
UNIT = "docs/s"
C5_DOCS, C5_SEGS, C5_CATS = 1_000_000_000, 64, 100_000
C5_NAME = ("C5 post_filter_agg_u64(status==0, terms_agg_u64(category 100k, (min,max,sum f64 price))) AllQuery, "
           "1e9 docs in 64 segments sharded over the ranks, NCCL reduce of the bucket tables to rank 0")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample-segments", type=int, default=0,
                    help="C5 segments (15.6M docs each) the CPU arm runs per step; 0 = one per host thread (segments are the unit of "
                         "the reference's thread pool, searcher.rs:79-92), at most 16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-per-config", action="store_true", help="skip the per-configuration table (N = 1)")
    ap.add_argument("--configs", default="c1,c1x,c2,c3,c4", help="per_config entries to run (C5 is the headline)")
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 force generic kernel, 2 force streaming kernel")
    return ap.parse_args()


def config_dict(world):
    return {"workload": C5_NAME, "docs": C5_DOCS, "segments": C5_SEGS, "categories": C5_CATS, "selectivity": 0.25,
            "n_gpus": world, "l2": "inputs (>= 1.16 GB per GPU and step) are larger than the 126 MB L2",
            "queries_in_flight": 1,
            "timing": "CUDA events on the execute stream around the K steps, max over ranks"}


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
            busy = [x for x in sm if x >= 0.9 * max(sm)] or sm
            out.update(sm_mhz=float(np.median(busy)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except OSError:
        return {}


def load_traffic():
    """DRAM bytes per launch of the dominant kernels, from this round's `ncu --set full` captures (profiles/)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except OSError:
        return {}


# ---------------------------------------------------------------------------------------------------
# workload recipes: the same counter-based generator on the device (tagg_synth_*) and on the host (oracle)
# ---------------------------------------------------------------------------------------------------
def c5_agg(ta):
    return ta.post_filter_agg_u64(STATUS, ta.eq(0), ta.terms_agg_u64(CATEGORY, (ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE), ta.sum_agg_f64(PRICE))))


def c5_device_segment(ta, ctx, seg_id):
    per = C5_DOCS // C5_SEGS
    seg = ta.Segment(ctx, per, keep_host=False)
    seg.synth_column(STATUS, ta.U64, 1, SEED, TAG_STATUS, seg_id * per, 0, 4)
    seg.synth_column(CATEGORY, ta.U64, 1, SEED, TAG_CATEGORY, seg_id * per, 1, C5_CATS)
    seg.synth_column(PRICE, ta.F64, 0, SEED, TAG_PRICE, seg_id * per)
    return seg


def c5_oracle_index(seg_ids):
    from oracle import oracle
    from tantivy_aggregations_b200 import _ffi as F
    per = C5_DOCS // C5_SEGS
    ix = oracle.OracleIndex()
    for s in seg_ids:
        o = ix.add_segment(per)
        ix.set_column_codes(o, STATUS, F.U64, oracle.synth_codes(1, SEED, TAG_STATUS, s * per, per, 0, 4))
        ix.set_column_codes(o, CATEGORY, F.U64, oracle.synth_codes(1, SEED, TAG_CATEGORY, s * per, per, 1, C5_CATS))
        ix.set_column_codes(o, PRICE, F.F64, oracle.synth_codes(0, SEED, TAG_PRICE, s * per, per))
        ix.segs[o].host.clear()
    return ix


def cpu_sample_segments(args):
    n = args.cpu_sample_segments or min(os.cpu_count() or 1, 16)
    return max(1, min(n, C5_SEGS))


def cpu_time(ix, query, mk_agg, threads, reps):
    """best-of-reps seconds of agg_search on the oracle in the Executor::ThreadPool shape, and in SingleThread."""
    best = {}
    for mode, thr in ((1, threads), (0, 1)):
        t_best = None
        for _ in range(reps if mode == 1 else 1):
            _, sec, _ = ix.search(query, mk_agg(), mode=mode, threads=thr, decode=False)
            t_best = sec if t_best is None else min(t_best, sec)
        best[mode] = t_best
    return best[1], best[0]


# ---------------------------------------------------------------------------------------------------
# --impl reference: the oracle restatement of the reference's collector loop on the host cores (the reference itself is
# Rust with un-vendored dependencies and cannot be built in this image — DESIGN.md)
# ---------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import tantivy_aggregations_b200 as ta
    threads = os.cpu_count() or 1
    nseg = cpu_sample_segments(args)
    per = C5_DOCS // C5_SEGS
    ix = c5_oracle_index(range(nseg))
    docs = nseg * per
    agg = lambda: c5_agg(ta)
    for _ in range(min(args.warmup, 2)):
        ix.search(ta.AllQuery(), agg(), mode=1, threads=threads, decode=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ix.search(ta.AllQuery(), agg(), mode=1, threads=threads, decode=False)
    dt = time.perf_counter() - t0
    value = docs * args.steps / dt
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    sample = f"{nseg} of {C5_SEGS} segments ({docs} docs) per step, Executor::ThreadPool shape, {threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64/f64", "data": "synthetic", "config": config_dict(world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement of the reference collector loop (oracle/oracle.cpp); the Rust reference cannot be built in this image; "
                "throughput of the sampled segments (segments are independent units: the full index costs 64/sample times as long)",
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import tantivy_aggregations_b200 as ta
        from tantivy_aggregations_b200 import _ffi as F
        from tantivy_aggregations_b200 import index as I
        self.torch, self.ta, self.F, self.I = torch, ta, F, I
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        self.ctx = ta.Context(self.local_rank)
        self.ctx.set_path(args.path)
        if self.world > 1:
            ident = [ta.Context.comm_unique_id() if self.rank == 0 else None]
            self.dist.broadcast_object_list(ident, src=0)
            self.ctx.comm_init(ident[0], self.rank, self.world)
        self.lib = F.lib()
        self.peaks = load_peaks()
        self.peak = float(self.peaks.get("hbm_gbs", 6650.0))
        self.peak_source = "MEASURED_PEAKS.json hbm_gbs" if self.peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        self.traffic = load_traffic()
        self._inputs_key, self._inputs = None, None

    def barrier(self):
        self.ctx.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.ctx.synchronize()

    def pinned_bitsets(self, segments, fill):
        """Per-segment bitsets as slices of ONE page-locked buffer at a 256-byte stride.  fill(seg_index, seg) -> uint8 array."""
        torch = self.torch
        needs = [(s.max_doc + 7) // 8 for s in segments]
        strides = [(n + 255) // 256 * 256 for n in needs]
        buf = torch.zeros(int(sum(strides)), dtype=torch.uint8).pin_memory()
        out, at = {}, 0
        for i, (seg, need, stride) in enumerate(zip(segments, needs, strides)):
            view = buf[at:at + stride].numpy()
            view[:need] = fill(i, seg)
            out[i] = view
            at += stride
        return buf, out, int(sum(needs))

    def step(self, plan, query, segments, read_nodes, collective_root=None):
        """One agg_search through the C ABI + reading the fruit arrays (zero-copy views of the page-locked image)."""
        I, F = self.I, self.F
        # the tagg_segment_input[] of a (plan, query, segment set) is built once: in the Rust facade that is a few stores per
        # segment, here it would be ~2 us of ctypes per segment and step
        key = (id(plan), id(query), len(segments), id(segments[0]) if segments else 0)
        if self._inputs_key != key:
            self._inputs = I.build_inputs(plan, query, segments)
            self._inputs_key = key
        arr, keep = self._inputs
        h = C.c_void_p()
        if collective_root is not None:
            F.check(self.lib.tagg_execute_reduce(plan._h, arr, len(segments), collective_root, C.byref(h)))
        else:
            F.check(self.lib.tagg_execute(plan._h, arr, len(segments), C.byref(h)))
        reader = I.ResultReader(h)
        nbytes, out = 0, {}
        if reader.is_local():
            for name, (kind, node) in read_nodes.items():
                a, b = reader.scope_view(node) if kind == "scope" else reader.metric_view(node)
                nbytes += a.nbytes + b.nbytes
                out[name] = (a, b)
        st = reader.stats()
        return out, nbytes, st, reader

    def step_begin(self, plan, query, segments):
        """tagg_execute_begin: everything of the step is queued on the GPU; returns without waiting."""
        I, F = self.I, self.F
        key = (id(plan), id(query), len(segments), id(segments[0]) if segments else 0)
        if self._inputs_key != key:
            self._inputs = I.build_inputs(plan, query, segments)
            self._inputs_key = key
        arr, keep = self._inputs
        h = C.c_void_p()
        F.check(self.lib.tagg_execute_begin(plan._h, arr, len(segments), C.byref(h)))
        return h

    def step_finish(self, pending, read_nodes):
        I, F = self.I, self.F
        h = C.c_void_p()
        F.check(self.lib.tagg_pending_wait(pending, C.byref(h)))
        reader = I.ResultReader(h)
        nbytes = 0
        for name, (kind, node) in read_nodes.items():
            a, b = reader.scope_view(node) if kind == "scope" else reader.metric_view(node)
            nbytes += a.nbytes + b.nbytes
        st = reader.stats()
        reader.free()
        return nbytes, st

    IN_FLIGHT = int(os.environ.get("TAGG_BENCH_INFLIGHT", "2"))  # depth of the extra `pipelined` measurement (tagg_execute_begin / tagg_pending_wait)

    def timed(self, plan, query, segments, read_nodes, steps, warmup, collective_root=None, in_flight=1):
        for _ in range(warmup):
            _, _, _, r = self.step(plan, query, segments, read_nodes, collective_root)
            r.free()
        if collective_root is None and in_flight > 1:
            # every call slot of the context warms its own scratch (device blocks, page-locked fruit image) on first use
            pending = [self.step_begin(plan, query, segments) for _ in range(in_flight)]
            for _ in range(2 * in_flight):
                self.step_finish(pending.pop(0), read_nodes)
                pending.append(self.step_begin(plan, query, segments))
            while pending:
                self.step_finish(pending.pop(0), read_nodes)
        self.barrier()
        l0 = self.ctx.launch_count()
        self.ctx.timer_start()
        t0 = time.perf_counter()
        kernel_ms, alg_bytes, d2h, path = 0.0, 0, 0, 0
        if collective_root is None and in_flight > 1:
            # the host prepares query i+1 while the GPU runs query i; every step is still one full agg_search whose fruit
            # arrays are read on the host
            pending = []
            for i in range(steps):
                pending.append(self.step_begin(plan, query, segments))
                if len(pending) >= in_flight:
                    nbytes, st = self.step_finish(pending.pop(0), read_nodes)
                    kernel_ms += st["kernel_ms"]; alg_bytes = st["alg_bytes"]; d2h = max(d2h, nbytes); path = st["path"]
            while pending:
                nbytes, st = self.step_finish(pending.pop(0), read_nodes)
                kernel_ms += st["kernel_ms"]; alg_bytes = st["alg_bytes"]; d2h = max(d2h, nbytes); path = st["path"]
        else:
            for _ in range(steps):
                _, nbytes, st, r = self.step(plan, query, segments, read_nodes, collective_root)
                r.free()
                kernel_ms += st["kernel_ms"]
                alg_bytes = st["alg_bytes"]
                d2h = max(d2h, nbytes)
                path = st["path"]
        dev_ms = self.ctx.timer_stop()
        wall_ms = 1e3 * (time.perf_counter() - t0)
        launches = self.ctx.launch_count() - l0
        self.barrier()
        ms = max(dev_ms, 0.0)
        kms = kernel_ms / steps
        if self.dist is not None:
            t = self.torch.tensor([ms, wall_ms, kms, float(d2h)], device="cuda", dtype=self.torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms, wall_ms, kms, d2h = float(t[0]), float(t[1]), float(t[2]), int(t[3])
        return dict(ms=ms / steps, wall_ms=wall_ms / steps, kernel_ms=kms, alg_bytes=alg_bytes, launches=launches, d2h=d2h, path=path)

    def roofline(self, res, kernel, traffic_key):
        achieved = res["alg_bytes"] / (res["kernel_ms"] * 1e-3) / 1e9 if res["kernel_ms"] > 0 else 0.0
        tr = self.traffic.get(traffic_key) or {}
        return {"bound": "hbm", "achieved": achieved, "peak": self.peak, "unit": "GB/s", "frac": achieved / self.peak if self.peak else None,
                "traffic": tr.get("dram_bytes_per_launch"), "traffic_source": tr.get("source"),
                "kernel": kernel, "peak_source": self.peak_source, "algorithmic_bytes_per_launch": res["alg_bytes"],
                "kernel_ms": res["kernel_ms"], "frac_of_8TBs_contract": achieved / 8000.0}

    # ---- the headline: C5, strong scaling ---------------------------------------------------------------------------
    def run_c5(self):
        ta, args = self.ta, self.args
        per = C5_DOCS // C5_SEGS
        parts = ta.assign_segments([per] * C5_SEGS, self.world)
        mine = parts[self.rank]
        segments = [c5_device_segment(ta, self.ctx, s) for s in mine]
        searcher = ta.Searcher(self.ctx, segments)
        agg = c5_agg(ta)
        plan = searcher.prepare(agg)
        terms = agg.sub
        read_nodes = {"keys": ("scope", terms.node), "min": ("metric", terms.sub.members[0].node),
                      "max": ("metric", terms.sub.members[1].node), "sum": ("metric", terms.sub.members[2].node)}
        root = 0 if self.world > 1 else None
        allq = ta.AllQuery()

        # ---- correctness gates before timing -------------------------------------------------------------------------
        # (1) size-independent properties of the full result
        out, _, st0, r0 = self.step(plan, allq, segments, read_nodes, root)
        if self.rank == 0:
            keys = out["keys"][0]
            assert len(keys) == C5_CATS and int(keys.min()) == 1 and int(keys.max()) == C5_CATS, len(keys)
            mn, mx, sm = out["min"][0].view(np.float64), out["max"][0].view(np.float64), out["sum"][0].view(np.float64)
            assert (mn >= 1.0).all() and (mx < 101.0).all() and (mn <= mx).all() and (sm >= mn).all()
            assert out["min"][1].all() and out["max"][1].all() and out["sum"][1].all()
        r0.free()
        # (2) the oracle: one segment PER RANK through the same collective (N = 1: two segments), merged fruit == oracle
        parity = None
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        sample_ids = sorted(p[0] for p in parts) if self.world > 1 else [mine[0], mine[-1]]
        local_sample = [segments[mine.index(s)] for s in sample_ids if s in mine]
        sub = ta.Searcher(self.ctx, local_sample)
        got = sub.agg_search_with_executor(allq, c5_agg(ta), ta.SINGLE_THREAD, collective=self.world > 1, root=root)
        for i, s in enumerate(segments):
            s.ord = i
        if self.rank == 0:
            from helpers import assert_fruit_equal
            ox = c5_oracle_index(sample_ids)
            want, _, _ = ox.search(allq, c5_agg(ta), mode=1, threads=os.cpu_count() or 1)
            assert_fruit_equal(got, want, 1e-12)
            parity = (f"oracle, {self.world} rank(s): merged fruit of {len(sample_ids)} full segments (one per rank, {per} docs each) "
                      "bit-exact buckets / min / max, f64 sums to 1e-12; full index: bucket set, bounds and Option flags")
            del ox, want

        # ---- value: resident inputs ------------------------------------------------------------------------------------
        sampler = ClockSampler(self.local_rank)
        res = self.timed(plan, allq, segments, read_nodes, args.steps, args.warmup, root)
        res_pipe = self.timed(plan, allq, segments, read_nodes, args.steps, 1, root, in_flight=self.IN_FLIGHT) if self.world == 1 else None
        # ---- e2e: the main docset arrives as page-locked host bitsets (every document matches) -------------------------
        buf, bits, h2d = self.pinned_bitsets(segments, lambda i, seg: np.full((seg.max_doc + 7) // 8, 0xFF, dtype=np.uint8))
        for i, seg in enumerate(segments):  # bits past max_doc stay clear
            tail = seg.max_doc % 8
            if tail:
                bits[i][(seg.max_doc + 7) // 8 - 1] = (1 << tail) - 1
        host_q = ta.BitsetQuery(bits)
        res_e2e = self.timed(plan, host_q, segments, read_nodes, args.steps, args.warmup, root)
        clocks = sampler.stop()
        if self.dist is not None:
            t = self.torch.tensor([float(h2d)], device="cuda", dtype=self.torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            h2d = int(t[0])

        line = None
        if self.rank == 0:
            value = C5_DOCS / (res["ms"] * 1e-3)
            e2e_value = C5_DOCS / (res_e2e["ms"] * 1e-3)
            cfg = config_dict(self.world)  # (identical keys and values in the reference arm's line)
            kernel_path = {0: "none", 1: "generic", 2: "stream"}.get(res["path"], "?")
            multi_gpu = ("segments sharded over the ranks; one key-domain agreement + one grouped ncclReduce of the bucket tables to rank 0 per step"
                         if self.world > 1 else "single GPU")
            roof = self.roofline(res, "k_stream<BK_TERMS, global tables, min|max|sum> (C5 shape)", "c5")
            roof["note"] = "per GPU (slowest rank): algorithmic bytes of this rank's shard / its kernel time"
            line = {
                "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": self.world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": res["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u64/f64", "data": "synthetic", "config": cfg, "kernel_path": kernel_path, "multi_gpu": multi_gpu,
                "kernel_ms_per_step": res["kernel_ms"], "roofline": roof,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": res_e2e["d2h"],
                        "ms_per_step": res_e2e["ms"], "kernel_ms_per_step": res_e2e["kernel_ms"],
                        "note": "host docsets: per-segment bitsets of the main query (all docs match) in page-locked memory, read by the GPU "
                                "inside the timed region; fruit arrays (keys + 3 metrics + flags) land in page-locked host memory every step"},
                "gpu_launches": res["launches"] + res_e2e["launches"],
                "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"]},
                "parity": parity, "matched_docs_per_step": C5_DOCS,
            }
            if res_pipe is not None:
                line["pipelined"] = {"queries_in_flight": self.IN_FLIGHT, "ms_per_step": res_pipe["ms"], "kernel_ms_per_step": res_pipe["kernel_ms"], "value": C5_DOCS / (res_pipe["ms"] * 1e-3), "unit": UNIT,
                                     "note": "the same steps with two queries in flight on the one GPU (tagg_execute_begin / tagg_pending_wait): host "
                                             "preparation and the fruit download of query i overlap the pass of query i+1"}
        for s in segments:
            s.close()
        return line

    # ---- per-configuration table (N = 1) --------------------------------------------------------------------------------
    def measure(self, name, workload, docs, segments, mk_agg, query_dev, query_host, h2d, read_nodes_of, kernel, traffic_key, steps=6, warmup=2):
        ta = self.ta
        searcher = ta.Searcher(self.ctx, segments)
        plan = searcher.prepare(mk_agg())
        rn = read_nodes_of(plan.agg)
        res = self.timed(plan, query_dev, segments, rn, steps, warmup)
        pipe = self.timed(plan, query_dev, segments, rn, steps, 1, in_flight=self.IN_FLIGHT) if res["ms"] < 1.0 else None
        entry = {"name": name, "workload": workload, "docs": docs, "kernel_ms": res["kernel_ms"], "ms_per_step": res["ms"],
                 "value": docs / (res["ms"] * 1e-3), "unit": UNIT, "launches_per_step": res["launches"] / steps,
                 "roofline": self.roofline(res, kernel, traffic_key)}
        if pipe is not None:
            entry["pipelined"] = {"queries_in_flight": self.IN_FLIGHT, "ms_per_step": pipe["ms"], "value": docs / (pipe["ms"] * 1e-3)}
        if query_host is not None:
            plan_h = searcher.prepare(mk_agg())
            rh = self.timed(plan_h, query_host, segments, read_nodes_of(plan_h.agg), steps, warmup)
            entry["e2e"] = {"value": docs / (rh["ms"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": rh["d2h"],
                            "ms_per_step": rh["ms"], "kernel_ms_per_step": rh["kernel_ms"]}
        else:
            entry["e2e"] = {"value": docs / (res["ms"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": res["d2h"],
                            "ms_per_step": res["ms"], "note": "AllQuery: the query has no host docset; the fruit is read back every step"}
        return entry

    def per_config(self):
        ta, F, ctx = self.ta, self.F, self.ctx
        from oracle import oracle
        from helpers import assert_fruit_equal
        which = [w for w in self.args.configs.split(",") if w]
        threads = os.cpu_count() or 1
        out = []
        allq = ta.AllQuery()

        def synth(n, nseg, cols):
            segs = []
            for s in range(nseg):
                seg = ta.Segment(ctx, n // nseg, keep_host=False)
                for c in cols:
                    c(seg, s * (n // nseg))
                segs.append(seg)
            return segs

        def close(segs):
            for s in segs:
                s.close()

        price = lambda s, b: s.synth_column(PRICE, ta.F64, 0, SEED, TAG_PRICE, b)
        root_nodes = lambda members: {f"m{i}": ("metric", m.node) for i, m in enumerate(members)}

        def cpu(name, ix, query, mk, docs, sample, check=None):
            pool, single = cpu_time(ix, query, mk, threads, 2)
            d = {"value": docs / pool, "unit": UNIT, "cores": threads, "kind": "port", "single_thread_value": docs / single,
                 "sample": sample + f"; Executor::ThreadPool shape, {threads} threads (SingleThread: {docs / single:.4g} docs/s)"}
            if check is not None:
                want, _, _ = ix.search(query, mk())
                assert_fruit_equal(check(), want, 1e-12)
                d["sample"] += "; GPU result on the sample == oracle"
            return d

        # ---- C1: count + sum/min/max f64 over AllQuery, 1M docs, and the same x1000 ----
        c1agg = lambda: (ta.count_agg(), ta.sum_agg_f64(PRICE), ta.min_agg_f64(PRICE), ta.max_agg_f64(PRICE))
        if "c1" in which:
            n = 1_000_000
            segs = synth(n, 1, [price])
            e = self.measure("C1", "count+sum+min+max f64 price, AllQuery, 1M docs / 1 segment (launch-latency bound: 6.9 MB)", n, segs, c1agg,
                             allq, None, 0, lambda a: root_nodes(a.members), "k_stream<root sum|min|max>", "c1", steps=20, warmup=3)
            ix = oracle.OracleIndex()
            o = ix.add_segment(n)
            ix.set_column_codes(o, PRICE, F.F64, oracle.synth_codes(0, SEED, TAG_PRICE, 0, n))
            e["cpu_baseline"] = cpu("c1", ix, allq, c1agg, n, "the whole configuration (1M docs)", lambda: ta.Searcher(ctx, segs).agg_search(allq, c1agg()))
            out.append(e)
            close(segs)
        if "c1x" in which:
            n = 1_000_000_000
            segs = synth(n, 64, [price])
            e = self.measure("C1x1000", "C1 scaled x1000: 1e9 docs in 64 segments (6.9 GB)", n, segs, c1agg, allq, None, 0,
                             lambda a: root_nodes(a.members), "k_stream<root sum|min|max>", "c1x")
            out.append(e)
            close(segs)
        # ---- C2: filter_agg(status=0, (count, terms(category 10k, (count, min price)))) on 100M docs ----
        if "c2" in which:
            n, nseg, ncat = 100_000_000, 8, 10_000
            segs = synth(n, nseg, [lambda s, b: s.synth_column(STATUS, ta.U64, 1, SEED, TAG_STATUS, b, 0, 4),
                                   lambda s, b: s.synth_column(CATEGORY, ta.U64, 1, SEED, TAG_CATEGORY, b, 1, ncat), price])
            for i, s in enumerate(segs):
                s.ord = i
            status_q = ta.TermQuery(STATUS, ta.U64, 0)
            buf, bits, h2d = self.pinned_bitsets(segs, lambda i, seg: seg.docset_to_bitset(status_q.docset(seg)))
            host_f = ta.BitsetQuery(bits)
            dev_f = ta.CachedQuery(host_f, segs)
            mk = lambda fq: (lambda: ta.filter_agg(fq, (ta.count_agg(), ta.terms_agg_u64(CATEGORY, (ta.count_agg(), ta.min_agg_f64(PRICE))))))
            rn = lambda a: {"keys": ("scope", a.sub.members[1].node), "root_count": ("metric", a.sub.members[0].node),
                            "count": ("metric", a.sub.members[1].sub.members[0].node), "min": ("metric", a.sub.members[1].sub.members[1].node)}
            searcher = ta.Searcher(ctx, segs)
            plan = searcher.prepare(mk(dev_f)())
            res = self.timed(plan, allq, segs, rn(plan.agg), 10, 3)
            pipe = self.timed(plan, allq, segs, rn(plan.agg), 10, 1, in_flight=self.IN_FLIGHT)
            plan_h = searcher.prepare(mk(host_f)())
            rh = self.timed(plan_h, allq, segs, rn(plan_h.agg), 10, 3)
            e = {"name": "C2", "workload": "filter_agg(status=0,(count,terms_u64(category 10k,(count,min_f64 price)))) AllQuery, 100M docs in 8 segments, 25% selectivity",
                 "docs": n, "kernel_ms": res["kernel_ms"], "ms_per_step": res["ms"], "value": n / (res["ms"] * 1e-3), "unit": UNIT,
                 "launches_per_step": res["launches"] / 10, "roofline": self.roofline(res, "k_stream<BK_TERMS, shared tables, min> (C2 shape)", "c2"),
                 "pipelined": {"queries_in_flight": self.IN_FLIGHT, "ms_per_step": pipe["ms"], "value": n / (pipe["ms"] * 1e-3)},
                 "e2e": {"value": n / (rh["ms"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": rh["d2h"], "ms_per_step": rh["ms"],
                         "kernel_ms_per_step": rh["kernel_ms"], "note": "filter bitsets in page-locked host memory, read by the GPU every step"}}
            ix = oracle.OracleIndex()
            cbits = {}
            per = n // nseg
            for s in range(2):
                o = ix.add_segment(per)
                st = oracle.synth_codes(1, SEED, TAG_STATUS, s * per, per, 0, 4)
                ix.set_column_codes(o, STATUS, F.U64, st)
                ix.set_column_codes(o, CATEGORY, F.U64, oracle.synth_codes(1, SEED, TAG_CATEGORY, s * per, per, 1, ncat))
                ix.set_column_codes(o, PRICE, F.F64, oracle.synth_codes(0, SEED, TAG_PRICE, s * per, per))
                cbits[o] = np.packbits((st == 0).astype(np.uint8), bitorder="little")
                ix.segs[o].host.clear()
            cq = ta.BitsetQuery(cbits)

            def gpu_sample():
                r = ta.Searcher(ctx, segs[:2]).agg_search(allq, mk(status_q)())
                for i, s in enumerate(segs):
                    s.ord = i
                return r
            e["cpu_baseline"] = cpu("c2", ix, allq, mk(cq), 2 * per, f"2 of {nseg} segments ({2 * per} docs)", gpu_sample)
            out.append(e)
            close(segs)
        # ---- C3: histogram(price, 0, 10, count) + percentiles(price) over 500M docs, 50 % bitset ----
        if "c3" in which:
            n, nseg = 500_000_000, 8
            segs = synth(n, nseg, [price])
            for i, s in enumerate(segs):
                s.ord = i
            rng = np.random.default_rng(3)
            buf, bits, h2d = self.pinned_bitsets(segs, lambda i, seg: rng.integers(0, 256, size=(seg.max_doc + 7) // 8, dtype=np.uint8))
            host_q = ta.BitsetQuery(bits)
            dev_q = ta.CachedQuery(host_q, segs)
            shapes = [("C3", "(histogram_agg_f64(price,0,10,count), percentiles_agg_f64(price)), 500M docs in 8 segments, 50% bitset — one fused pass",
                       lambda: (ta.histogram_agg_f64(PRICE, 0.0, 10.0, ta.count_agg()), ta.percentiles_agg_f64(PRICE)),
                       lambda a: {"ords": ("scope", a.members[0].node), "count": ("metric", a.members[0].sub.node)}, "k_stream<BK_RANK + fused histogram>", "c3"),
                      ("C3-histogram", "histogram_agg_f64(price,0,10,count) alone, same docset", lambda: ta.histogram_agg_f64(PRICE, 0.0, 10.0, ta.count_agg()),
                       lambda a: {"ords": ("scope", a.node), "count": ("metric", a.sub.node)}, "k_stream<BK_HIST, count>", "c3h"),
                      ("C3-percentiles", "percentiles_agg_f64(price) alone, same docset", lambda: ta.percentiles_agg_f64(PRICE), lambda a: {}, "k_stream<BK_RANK>", "c3p")]
            for nm, wl, mk, rn, kern, tk in shapes:
                e = self.measure(nm, wl, n // 2, segs, mk, dev_q, host_q if nm == "C3" else None, h2d, rn, kern, tk, steps=5, warmup=2)
                e["docs_note"] = "matched docs (50 % of 500M); the pass streams all 500M"
                if nm == "C3":
                    m = 250_000  # the oracle's CKMS restatement inserts ~0.1 M values / s: a small sample
                    ix = oracle.OracleIndex()
                    o = ix.add_segment(m)
                    ix.set_column_codes(o, PRICE, F.F64, oracle.synth_codes(0, SEED, TAG_PRICE, 0, m))
                    cb = {o: bits[0][: (m + 7) // 8].copy()}
                    nsel = int(np.unpackbits(cb[o], bitorder="little")[:m].sum())
                    pool, single = cpu_time(ix, ta.BitsetQuery(cb), mk, threads, 1)
                    e["cpu_baseline"] = {"value": nsel / pool, "unit": UNIT, "cores": 1, "kind": "port",
                                         "sample": f"{m} docs of segment 0 ({nsel} matched), one segment = one thread; CKMS(0.01) insert per matched value dominates"}
                out.append(e)
            close(segs)
        # ---- C4: terms_agg_u64s(keys, sum_agg_f64s(vals)), 250M docs, ~1e9 key values, 1M distinct keys ----
        if "c4" in which:
            n, nseg, nkeys, spread = 250_000_000, 16, 1_000_000, 1 << 20
            segs = synth(n, nseg, [lambda s, b: s.synth_multicolumn(KEYS, ta.U64, 1, SEED, TAG_KEYS, b, 9, 0, nkeys),
                                   lambda s, b: s.synth_multicolumn(KEYS_SPREAD, ta.U64, 2, SEED, TAG_KEYS, b, 9, 5, nkeys, spread),
                                   lambda s, b: s.synth_multicolumn(KEYS_ZIPF, ta.U64, 3, SEED, TAG_KEYS, b, 9, 0, nkeys),
                                   lambda s, b: s.synth_multicolumn(VALS, ta.F64, 0, SEED, TAG_VALS, b, 3)])
            n_values = sum(s.column_info(KEYS, 0)["n_values"] for s in segs)
            rn = lambda a: {"keys": ("scope", a.node), "sum": ("metric", a.sub.node)}
            for nm, wl, field, kern, tk in (
                    ("C4-dense", "terms_agg_u64s(keys 1M distinct, sum_agg_f64s(vals)), 250M docs, 1e9 key values — dense table", KEYS, "k_mterms<dense>", "c4"),
                    ("C4-hashed", "same keys spread over a 40-bit domain — global open-addressing spill table", KEYS_SPREAD, "k_mterms<hashed>", "c4h"),
                    ("C4-zipf", "same shape, Zipf-distributed keys (1M distinct) — dense table behind the shared-memory hot-key front", KEYS_ZIPF, "k_mterms<dense>", "c4z")):
                mk = (lambda f: (lambda: ta.terms_agg_u64s(f, ta.sum_agg_f64s(VALS))))(field)
                e = self.measure(nm, wl, n, segs, mk, allq, None, 0, rn, kern, tk, steps=3, warmup=1)
                e["key_values"] = n_values
                e["values_per_s"] = n_values / (e["ms_per_step"] * 1e-3)
                e["roofline"]["l2_atomic_floor_ms"] = 6.7e8 / 190e9 * 1e3
                e["roofline"]["l2_atomic_note"] = ("6.7e8 RED.F64 (key occurrences of documents that have values) at the measured 190 G/s L2 atomic rate "
                                                  "(profiles/r1_atom_bench_b200.txt) — the kernel's own floor, above its HBM time")
                if nm == "C4-dense":
                    m = 2_000_000
                    ix = oracle.OracleIndex()
                    o = ix.add_segment(m)
                    off, codes = oracle.synth_multi(1, SEED, TAG_KEYS, 0, m, 9, 0, nkeys)
                    ix.set_multicolumn_codes(o, KEYS, F.U64, off, codes)
                    off, codes = oracle.synth_multi(0, SEED, TAG_VALS, 0, m, 3)
                    ix.set_multicolumn_codes(o, VALS, F.F64, off, codes)
                    ix.segs[o].host.clear()
                    pool, single = cpu_time(ix, allq, mk, threads, 1)
                    e["cpu_baseline"] = {"value": m / pool, "unit": UNIT, "cores": 1, "kind": "port",
                                         "sample": f"the first {m} docs of segment 0, one segment = one thread"}
                out.append(e)
            close(segs)
        return out

    def cpu_baseline_c5(self):
        ta, args = self.ta, self.args
        threads = os.cpu_count() or 1
        nseg = cpu_sample_segments(args)
        per = C5_DOCS // C5_SEGS
        ix = c5_oracle_index(range(nseg))
        pool, single = cpu_time(ix, ta.AllQuery(), lambda: c5_agg(ta), threads, 2)
        return {"value": nseg * per / pool, "unit": UNIT, "cores": threads, "kind": "port", "single_thread_value": per * nseg / single,
                "sample": f"{nseg} of {C5_SEGS} segments ({nseg * per} docs); Executor::ThreadPool shape with {threads} threads; "
                          f"Executor::SingleThread (agg_search default) = {per * nseg / single:.4g} docs/s"}

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    b = Bench(args)
    line = b.run_c5()
    if b.rank == 0 and b.world == 1:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = b.cpu_baseline_c5()
        if not args.no_per_config:
            line["per_config"] = b.per_config()
    if b.rank == 0:
        print(json.dumps(line))
    b.close()


if __name__ == "__main__":
    main()
