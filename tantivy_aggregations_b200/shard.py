"""Segment -> GPU partitioning: segments are independent units and fruits form a
commutative monoid under `PreparedAgg::merge`, so the path shards with no data-path collective; the
one exchange step is the merge of the bucket tables (tagg_execute_collective / Agg.merge)."""


def assign_segments(sizes, n_ranks):
    """Greedy longest-processing-time partition of segments (by max_doc) over ranks.
    Returns a list of n_ranks lists of segment ordinals, each ascending (segment order is the merge
    order of the reference's Executor::ThreadPool path, src/searcher.rs:93-96)."""
    if n_ranks < 1:
        raise ValueError("n_ranks must be >= 1")
    loads = [0] * n_ranks
    out = [[] for _ in range(n_ranks)]
    for ord_ in sorted(range(len(sizes)), key=lambda i: (-sizes[i], i)):
        r = min(range(n_ranks), key=lambda k: (loads[k], k))
        out[r].append(ord_)
        loads[r] += sizes[ord_]
    return [sorted(x) for x in out]


def merge_fruits(agg, fruits):
    """Fold per-rank (or per-segment) fruits with PreparedAgg::merge, in the given order."""
    from .agg import as_agg
    agg = as_agg(agg)
    acc = agg.create_fruit()
    for f in fruits:
        acc = agg.merge(acc, f)
    return acc
