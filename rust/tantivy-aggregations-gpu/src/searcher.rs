//! `AggSearcher` with the reference's signatures (`src/searcher.rs:12-25`).  The reference's `collect_segment`
//! (`searcher.rs:27-51`: `scorer.for_each(|doc, score| segment_agg.collect(doc, score, harvest))`, with the delete check
//! of `:41-46`) is gone: the matched docs of every segment are handed to the GPU as docsets and ONE `tagg_execute` folds
//! all segments into one harvest (`Executor::SingleThread`, `:66-78`), or one call per segment + `tagg_result_merge` in
//! segment order reproduces the thread-pool shape (`:79-98`).
use std::sync::Mutex;

use crate::agg::{check, Agg, PlanBuilder, ResultReader};
use crate::gpu::{all_docset, drain, DrainedDocset, GpuIndex};
use tagg_sys as sys;
use tantivy::query::{AllQuery, Query};
use tantivy::{Executor, Result, Searcher};

pub trait AggSearcher {
    fn agg_search<A: Agg>(&self, query: &dyn Query, agg: &A) -> Result<A::Fruit> {
        self.agg_search_with_executor(query, agg, &Executor::SingleThread)
    }
    fn agg_search_with_executor<A: Agg>(&self, query: &dyn Query, agg: &A, executor: &Executor) -> Result<A::Fruit>;
}

/// The searcher plus the device mirror of its segments.
pub struct GpuSearcher<'a> { pub searcher: &'a Searcher, pub gpu: &'a Mutex<GpuIndex> }

impl<'a> AggSearcher for GpuSearcher<'a> {
    fn agg_search_with_executor<A: Agg>(&self, query: &dyn Query, agg: &A, executor: &Executor) -> Result<A::Fruit> {
        let searcher = self.searcher;
        let weight = query.weight(searcher, agg.requires_scoring())?;          // searcher.rs:62
        let mut pb = PlanBuilder::default();                                   // Agg::prepare (searcher.rs:63) == lowering
        let root = agg.emit_plan(&mut pb);
        let filter_weights = pb.filters.iter().map(|q| q.weight(searcher, false)).collect::<Result<Vec<_>>>()?;  // filter.rs:34-39
        let mut gpu = self.gpu.lock().unwrap();
        let mut plan = std::ptr::null_mut();
        check(unsafe { sys::tagg_plan_create(gpu.ctx, pb.nodes.as_ptr(), pb.nodes.len() as u32, std::ptr::null(), 0, &mut plan) })?;

        let is_all = query.as_any().is::<AllQuery>();
        let mut drained: Vec<DrainedDocset> = Vec::new();           // keep the host buffers alive over the call
        let mut per_segment: Vec<(usize, Vec<usize>)> = Vec::new();  // indices into `drained`
        let mut segs = Vec::new();
        for reader in searcher.segment_readers() {
            segs.push(gpu.ensure_segment(reader, &searcher.schema())?.raw as *const sys::tagg_segment);
            let main = if is_all { usize::MAX } else { drained.push(drain(weight.scorer(reader)?, reader.max_doc())); drained.len() - 1 };
            let mut fl = Vec::new();
            for w in &filter_weights {                               // filter.rs:65-73: the filter's own scorer per segment
                drained.push(drain(w.scorer(reader)?, reader.max_doc()));
                fl.push(drained.len() - 1);
            }
            per_segment.push((main, fl));
        }
        let filter_docsets: Vec<Vec<sys::tagg_docset>> = per_segment.iter().map(|(_, fl)| fl.iter().map(|&i| drained[i].as_docset()).collect()).collect();
        let inputs: Vec<sys::tagg_segment_input> = per_segment.iter().enumerate().map(|(i, (main, _))| sys::tagg_segment_input {
            segment: segs[i],
            docset: if *main == usize::MAX { all_docset() } else { drained[*main].as_docset() },
            filters: filter_docsets[i].as_ptr(),
            n_filters: filter_docsets[i].len() as u32,
        }).collect();

        let mut res = std::ptr::null_mut();
        let status = match executor {
            Executor::SingleThread => unsafe { sys::tagg_execute(plan, inputs.as_ptr(), inputs.len() as u32, &mut res) },
            _ => unsafe {  // a fruit per segment, merged in segment order (searcher.rs:93-96)
                let mut st = sys::tagg_execute(plan, inputs.as_ptr(), inputs.len().min(1) as u32, &mut res);
                for input in inputs.iter().skip(1) {
                    if st != sys::TAGG_OK { break; }
                    let mut part = std::ptr::null_mut();
                    st = sys::tagg_execute(plan, input, 1, &mut part);
                    if st == sys::TAGG_OK { st = sys::tagg_result_merge(res, part); sys::tagg_result_free(part); }
                }
                st
            },
        };
        unsafe { sys::tagg_plan_destroy(plan) };
        check(status)?;
        let reader = ResultReader { raw: res };
        agg.read_fruit(&reader, root, 0)
    }
}
