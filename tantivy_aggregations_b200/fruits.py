"""Result ("fruit") types of the aggregation API — mirrors of the reference's
`Terms<K,T>` (src/bucket/terms.rs:403-458), `Histogram<T>` (src/bucket/histogram.rs:156-181)
and `Percentiles<T>` (src/metric/percentile.rs:152-177).  `Option<T>` is None / value and a
tuple fruit is a Python tuple.
"""
import heapq
import math

CKMS_EPS = 0.01  # percentile.rs:174


class Terms:
    """`Terms<K, T>`: bucket key -> sub fruit."""

    def __init__(self, res=None):
        self.res = dict(res or {})

    def get(self, key):
        """terms.rs:421-423"""
        return self.res.get(key)

    def __len__(self):
        return len(self.res)

    def top_k(self, k, sort_by=lambda b: b):
        """terms.rs:425-457: the k buckets with the largest sort key, descending; equal sort
        keys ascending by bucket key.  (Which of several buckets tied AT the cut survive is
        HashMap-iteration-order dependent in the reference; here iteration is by ascending key.)"""
        if not self.res or k == 0:
            return []

        class _Rev:  # std::cmp::Reverse
            __slots__ = ("v",)

            def __init__(self, v):
                self.v = v

            def __lt__(self, o):
                return o.v < self.v

            def __eq__(self, o):
                return o.v == self.v

        # max-heap of (Reverse(sort), key)  ==  python min-heap on the negation: keep it simple and
        # restate the observable result: sort by (sort key desc, key asc), the heap only bounds memory.
        items = sorted(self.res.items(), key=lambda kv: kv[0])
        best = heapq.nsmallest(k, items, key=lambda kv: (_Rev(sort_by(kv[1])), kv[0]))
        return [(key, fruit) for key, fruit in best]

    def canon(self):
        return {k: _canon(v) for k, v in self.res.items()}

    def __repr__(self):
        return f"Terms({self.res!r})"


class Histogram:
    """`Histogram<T>`: bucket_ord -> sub fruit, plus start / interval."""

    def __init__(self, start, interval, buckets=None):
        self.start = float(start)
        self.interval = float(interval)
        self._buckets = dict(sorted((buckets or {}).items()))  # BTreeMap<u64, T>

    def buckets(self):
        """histogram.rs:163-181: ascending (bucket key, Some(fruit)) with gap buckets as None."""
        res = []
        last = None
        for ord_, agg in self._buckets.items():
            if last is not None:
                gap = ord_ - last
                if gap > 1:
                    for i in range(gap - 1):
                        res.append((float(last + i + 1) * self.interval + self.start, None))
            res.append((float(ord_) * self.interval + self.start, agg))
            last = ord_
        return res

    def canon(self):
        return ("hist", self.start, self.interval, {k: _canon(v) for k, v in self._buckets.items()})

    def __repr__(self):
        return f"Histogram(start={self.start}, interval={self.interval}, {self._buckets!r})"


def ckms_target_rank(q, n, eps=CKMS_EPS):
    """1-based rank the reference's CKMS::query(q) returns while the sketch is uncompressed
    (SURVEY §8a): k = clamp(floor(q*n + max(1, floor(2*eps*q*n)) / 2), 1, n)."""
    nphi = q * n
    inv = max(1, math.floor(2.0 * eps * nphi))
    k = math.floor(nphi + inv / 2.0)
    return min(max(k, 1), n)


class Percentiles:
    """`Percentiles<f64>`.  The reference stores a CKMS(eps=0.01) sketch; this fruit stores exact
    order statistics (rank, value) and answers percentile(q) with the stored statistic whose
    rank is nearest the rank CKMS targets — zero rank error whenever that rank is stored
    (always for small inputs), else within the sketch's own +-eps*q*n band (checked in tests)."""

    def __init__(self, n_total, ranks, values):
        self.n = int(n_total)
        self.ranks = list(ranks)
        self.values = list(values)

    def percentile(self, q):
        """percentile.rs:163-165"""
        if self.n == 0 or not self.ranks:
            return None
        k = ckms_target_rank(q, self.n)
        return self.values[self._nearest(k)]

    def _nearest(self, k):
        import bisect
        i = bisect.bisect_left(self.ranks, k)
        if i == len(self.ranks):
            return i - 1
        if i == 0 or self.ranks[i] == k:
            return i
        return i if (self.ranks[i] - k) < (k - self.ranks[i - 1]) else i - 1

    def rank_of_answer(self, q):
        """(target rank, rank of the returned statistic) — lets tests bound the rank error."""
        k = ckms_target_rank(q, self.n)
        return k, self.ranks[self._nearest(k)]

    def canon(self):
        return ("pct", self.n)

    def __repr__(self):
        return f"Percentiles(n={self.n}, stored={len(self.ranks)})"


def _canon(f):
    if isinstance(f, tuple):
        return tuple(_canon(x) for x in f)
    if hasattr(f, "canon"):
        return f.canon()
    return f


def canon(fruit):
    """Plain-Python canonical form of a fruit (dicts / tuples / scalars) for comparisons."""
    return _canon(fruit)
