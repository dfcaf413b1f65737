"""CPU, world_size 2 over gloo: the host-side logic of the N>1 path — segment partitioning, every rank
aggregating only its own segments, and the merge of the per-rank fruits (`PreparedAgg::merge`) — checked
against the single-process oracle.  (The device-side NCCL table merge is covered by
tests/test_gpu_multi.py on a box with >= 2 GPUs.)"""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import tantivy_aggregations_b200 as ta
    from helpers import assert_fruit_equal
    from test_dist_gloo import build_corpus, plans

    corpus = build_corpus()
    ox = corpus.build_oracle()
    parts = ta.assign_segments([s.max_doc for s in corpus.segs], world)
    mine = parts[rank]
    ok = True
    for name, mk in plans().items():
        local, _, _ = ox.search(ta.AllQuery(), mk(), mode=1, threads=2, segments=mine)
        gathered = [None] * world
        dist.all_gather_object(gathered, local)
        merged = ta.merge_fruits(mk(), gathered)
        want, _, _ = ox.search(ta.AllQuery(), mk(), mode=1, threads=2)
        assert_fruit_equal(merged, want, 1e-12, name)
    dist.barrier()
    with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
        f.write(repr(parts))
    dist.destroy_process_group()


def build_corpus():
    from helpers import Corpus, SegSpec
    from tantivy_aggregations_b200 import _ffi as F
    rng = np.random.default_rng(42)
    segs = []
    for n in (4000, 100, 2500, 3000, 0, 1777):
        s = SegSpec(n)
        s.col(1, F.U64, rng.integers(1, 40, size=n, dtype=np.uint64))
        s.col(2, F.F64, 1.0 + 100.0 * rng.random(n))
        s.col(3, F.I64, rng.integers(-50, 50, size=n, dtype=np.int64))
        s.mcol(4, F.U64, [list(rng.integers(0, 9, size=rng.integers(0, 4))) for _ in range(n)])
        s.col(5, F.U64, (rng.integers(1, 6, size=n, dtype=np.uint64) << np.uint64(40)) | rng.integers(1, 30, size=n, dtype=np.uint64))  # sparse 43-bit keys: hashed table
        if n:
            s.deleted = rng.choice(n, size=n // 7, replace=False)
        segs.append(s)
    return Corpus(segs)


def plans():
    import tantivy_aggregations_b200 as ta
    return {
        "scalars": lambda: (ta.count_agg(), ta.sum_agg_f64(2), ta.min_agg_f64(2), ta.max_agg_i64(3), ta.sum_agg_i64(3)),
        "terms": lambda: ta.terms_agg_u64(1, (ta.count_agg(), ta.min_agg_f64(2), ta.sum_agg_f64(2))),
        "nested": lambda: ta.terms_agg_u64s(4, (ta.count_agg(), ta.histogram_agg_f64(2, 0.0, 20.0, (ta.count_agg(), ta.max_agg_f64(2))))),
        "post_filter": lambda: ta.post_filter_agg_i64(3, ta.ge(0), ta.terms_agg_i64(3, ta.count_agg())),
        "date_histogram": lambda: ta.terms_agg_u64(1, ta.date_histogram_agg(3, 7, (ta.count_agg(), ta.sum_agg_f64(2)), start=-20, kind=ta.I64)),
    }


def test_assign_segments_balanced_and_complete():
    import tantivy_aggregations_b200 as ta
    sizes = [15_625_000] * 64
    for n in (1, 2, 4, 8):
        parts = ta.assign_segments(sizes, n)
        assert sorted(x for p in parts for x in p) == list(range(64))
        assert {len(p) for p in parts} == {64 // n}
    parts = ta.assign_segments([10, 1, 1, 1, 7, 3], 2)
    loads = [sum([10, 1, 1, 1, 7, 3][i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 1 and all(p == sorted(p) for p in parts)


def test_two_rank_sharded_aggregation_matches_single_process(tmp_path):
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
