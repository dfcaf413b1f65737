"""GPU, >= 2 devices: segments sharded over ranks (one process per GPU), bucket tables merged by the
NCCL exchange step (tagg_execute_collective), every rank's fruit equal to the single-process oracle.
Skipped on a single-GPU box."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)  # host-side plumbing only (unique id broadcast)
    import tantivy_aggregations_b200 as ta
    from helpers import Corpus, assert_fruit_equal
    from test_dist_gloo import build_corpus

    corpus = build_corpus()
    ox = corpus.build_oracle()
    parts = ta.assign_segments([s.max_doc for s in corpus.segs], world)
    ctx = ta.Context(rank)
    ident = [ta.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    ctx.comm_init(ident[0], rank, world)
    local = Corpus([corpus.segs[i] for i in parts[rank]]).build_gpu(ctx)
    shapes = {
        "scalars": lambda: (ta.count_agg(), ta.sum_agg_f64(2), ta.min_agg_f64(2), ta.max_agg_i64(3), ta.sum_agg_i64(3)),
        "terms": lambda: ta.terms_agg_u64(1, (ta.count_agg(), ta.min_agg_f64(2), ta.max_agg_f64(2), ta.sum_agg_f64(2))),
        "bench_shape": lambda: ta.filter_agg(ta.RangeQuery(3, ta.I64, 0, 49), (ta.count_agg(), ta.terms_agg_u64(1, (ta.count_agg(), ta.min_agg_f64(2))))),
        "hist": lambda: ta.histogram_agg_f64(2, 0.0, 10.0, (ta.count_agg(), ta.sum_agg_f64(2))),
        "nested_dense": lambda: ta.terms_agg_u64(1, ta.histogram_agg_f64(2, 0.0, 50.0, ta.count_agg())),
        "multi": lambda: ta.terms_agg_u64s(4, (ta.count_agg(), ta.min_agg_f64(2))),
    }
    for name, mk in shapes.items():
        want, _, _ = ox.search(ta.AllQuery(), mk(), mode=1, threads=2)
        for path in (0, 1):
            ctx.set_path(path)
            got = local.agg_search_with_executor(ta.AllQuery(), mk(), ta.SINGLE_THREAD, collective=True)
            assert_fruit_equal(got, want, 1e-12, f"{name}/path{path}/rank{rank}")
    # tagg_execute_reduce: ncclReduce of the bucket tables into ONE root; the other ranks get a placeholder
    for root in range(world):
        for name, mk in shapes.items():
            want, _, _ = ox.search(ta.AllQuery(), mk(), mode=1, threads=2)
            got = local.agg_search_with_executor(ta.AllQuery(), mk(), ta.SINGLE_THREAD, collective=True, root=root)
            if rank == root:
                assert_fruit_equal(got, want, 1e-12, f"reduce/{name}/root{root}")
            else:
                assert got is None or name in ("multi",), (name, rank, root)  # (a plan that exchanges compact results returns the fruit everywhere)
    # ONE rank's segment set changes between two calls on the SAME plan (an index refresh on that rank only): every call
    # re-agrees the table layout, the unchanged ranks notice that their cached agreement moved and redo the pass
    for name in ("terms", "bench_shape", "nested_dense"):
        plan = local.prepare(shapes[name]())
        for drop in (0, 1, 0, 2):
            mine = list(parts[rank])
            if rank == world - 1 and drop:
                mine = mine[:-drop] if drop < len(mine) else []
            sub = ta.Searcher(ctx, [local.segments[parts[rank].index(i)] for i in mine])
            kept = sorted(i for r in range(world) for i in (parts[r] if r != world - 1 or not drop else (parts[r][:-drop] if drop < len(parts[r]) else [])))
            want, _, _ = ox.search(ta.AllQuery(), shapes[name](), mode=1, threads=2, segments=kept)
            got = sub.agg_search_with_executor(ta.AllQuery(), plan, ta.SINGLE_THREAD, collective=True)
            assert_fruit_equal(got, want, 1e-12, f"refresh/{name}/drop{drop}/rank{rank}")
            got = sub.agg_search_with_executor(ta.AllQuery(), plan, ta.SINGLE_THREAD, collective=True, root=0)
            if rank == 0:
                assert_fruit_equal(got, want, 1e-12, f"refresh-reduce/{name}/drop{drop}")
        for i, sgm in enumerate(local.segments):
            sgm.ord = i
    # f64 min / max over NaN and signed zeros: per-rank exact fold, merged in rank order (PreparedAgg::merge, minmax.rs:59-72)
    import numpy as np
    from helpers import SegSpec
    from tantivy_aggregations_b200 import _ffi as F
    nan = float("nan")
    edge_vals = [[3.0, -0.0, 1.0], [nan, 0.0, 2.0], [0.0, -0.0], [nan]]
    edge_segs = []
    for v in edge_vals:
        sp = SegSpec(len(v))
        sp.col(2, F.F64, np.array(v))
        sp.col(1, F.U64, np.arange(len(v), dtype=np.uint64) % 2)
        edge_segs.append(sp)
    eparts = [[0, 1], [2, 3]] if world == 2 else ta.assign_segments([len(v) for v in edge_vals], world)
    elocal = Corpus([edge_segs[i] for i in eparts[rank]]).build_gpu(ctx)
    emk = lambda: (ta.min_agg_f64(2), ta.max_agg_f64(2), ta.terms_agg_u64(1, (ta.min_agg_f64(2), ta.max_agg_f64(2))))
    # the reference shape of "one fruit per unit, merged in order" with the ranks as units
    per_rank = [Corpus([edge_segs[i] for i in eparts[r]]).build_oracle().search(ta.AllQuery(), emk())[0] for r in range(world)]
    want = ta.merge_fruits(emk(), per_rank)
    got = elocal.agg_search_with_executor(ta.AllQuery(), emk(), ta.SINGLE_THREAD, collective=True)
    assert_fruit_equal(got, want, 0.0, f"edge/rank{rank}")
    # hashed bucket tables and percentile summaries: merged as compact results (all-gather + PreparedAgg::merge in rank order)
    from helpers import exact_rank_window
    ctx.set_path(0)
    sparse_key = lambda: ta.terms_agg_u64(5, (ta.count_agg(), ta.min_agg_f64(2)))
    want, _, _ = ox.search(ta.AllQuery(), sparse_key(), mode=1, threads=2)
    got = local.agg_search_with_executor(ta.AllQuery(), sparse_key(), ta.SINGLE_THREAD, collective=True)
    assert_fruit_equal(got, want, 1e-12, f"hashed/rank{rank}")
    pct_plan = lambda: (ta.count_agg(), ta.percentiles_agg_f64(2), ta.terms_agg_u64(1, ta.percentiles_agg_f64(2)))
    cnt, pct, terms = local.agg_search_with_executor(ta.AllQuery(), pct_plan(), ta.SINGLE_THREAD, collective=True)
    alive_vals, alive_cat = [], []
    for sg in corpus.segs:
        alive = np.ones(sg.max_doc, dtype=bool)
        if sg.deleted is not None and sg.max_doc:
            alive[np.asarray(sg.deleted)] = False
        from tantivy_aggregations_b200 import codec
        alive_vals.append(codec.code_to_f64(sg.cols[2][1])[alive])
        alive_cat.append(sg.cols[1][1][alive])
    vals, cats = np.concatenate(alive_vals), np.concatenate(alive_cat)
    assert cnt == len(vals) == pct.n

    def check(p, v):
        srt = np.sort(v)
        assert p.n == len(srt)
        for qq in (0.01, 0.25, 0.5, 0.75, 0.99):
            x = p.percentile(qq)
            lo, hi = exact_rank_window(srt, x)
            k = ta.ckms_target_rank(qq, len(srt))
            band = 0.01 * qq * len(srt) + 1
            assert lo <= hi and lo - band <= k <= hi + band, (qq, k, lo, hi)

    check(pct, vals)
    assert set(terms.res) == set(int(c) for c in np.unique(cats))
    for c, p in terms.res.items():
        check(p, vals[cats == c])
    dist.barrier()
    with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
        f.write("ok")
    dist.destroy_process_group()


def test_collective_merge_matches_oracle(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    port = 31000 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
