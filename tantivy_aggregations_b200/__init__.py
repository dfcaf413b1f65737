"""tantivy-aggregations hot path on B200 — host-side mirror of the reference's public API
(reference src/lib.rs:10-15) over the C ABI of libtagg.so (include/tagg.h).

Everything that computes runs in hand-written sm_100a CUDA kernels behind the C ABI; this package
only lowers aggregation trees to plans, hands docsets over, and decodes fruits.  There is no CPU
fallback: without libtagg.so / a GPU every compute entry point raises.
"""
from . import _ffi
from ._ffi import DATE, F64, I64, U64, FastFieldNotAvailableError, TaggError
from .agg import *  # noqa: F401,F403  (the reference's constructor functions)
from .agg import Agg, eq, ge, gt, in_set, le, lt
from .fruits import Histogram, Percentiles, Terms, canon, ckms_target_rank
from .shard import assign_segments, merge_fruits
from .index import (SINGLE_THREAD, THREAD_POOL, AllQuery, BitsetQuery, CachedQuery, Context, DocIdsQuery, Plan, RangeQuery,
                    ResultReader, Searcher, Segment, TermQuery)

__all__ = [n for n in dir() if not n.startswith("_")]
