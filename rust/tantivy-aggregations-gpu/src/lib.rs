//! Drop-in replacement of the reference crate's hot path (`src/lib.rs:1-15` re-exports are kept): the same constructor
//! functions and `AggSearcher`, but `agg_search` makes ONE FFI call per search instead of one `collect` per document.
//!
//! UNCOMPILED: the build image has no rustc.  Reviewed by hand against `/root/reference/src/{agg,searcher,tuple}.rs` and the
//! leaf modules; the Python and C++ mirrors of this file are what the test-suite runs.
mod agg;
mod gpu;
mod metric;
mod bucket;
mod searcher;

pub use crate::agg::{Agg, PlanBuilder, ResultReader};
pub use crate::bucket::{filter_agg, histogram_agg_f64, terms_agg_u64, Histogram, Terms};
pub use crate::gpu::{GpuIndex, GpuSegment};
pub use crate::metric::{count_agg, max_agg_f64, min_agg_f64, sum_agg_f64, sum_agg_u64};
pub use crate::searcher::AggSearcher;
