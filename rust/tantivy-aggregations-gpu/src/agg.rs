//! The plugin contract.  The reference's `Agg` (`src/agg.rs:10-17`) keeps its name, `Fruit` and `requires_scoring`;
//! `prepare -> PreparedAgg -> SegmentAgg::collect(doc)` (`agg.rs:19-36`) is replaced by lowering to a flat plan
//! (`emit_plan`) and decoding the device result (`read_fruit`): per-document `collect` is no longer the execution model.
use std::os::raw::c_int;
use tagg_sys as sys;
use tantivy::query::Query;
use tantivy::schema::Field;
use tantivy::{Result, TantivyError};

pub trait Agg {
    type Fruit: Send;
    /// every built-in returns false (count.rs:19, sum.rs:32, minmax.rs:32, terms.rs:51 ...)
    fn requires_scoring(&self) -> bool { false }
    /// Append this node (pre-order) and its children to the flat plan; remember the node index for `read_fruit`.
    fn emit_plan<'q>(&'q self, plan: &mut PlanBuilder<'q>) -> u32;
    /// Rebuild the typed fruit of `bucket` (an index into the enclosing scope's bucket order) from the device result.
    fn read_fruit(&self, res: &ResultReader, node: u32, bucket: u32) -> Result<Self::Fruit>;
    /// Nodes this sub-tree emits (so that a tuple can find its members' node indices).
    fn n_nodes(&self) -> u32;
}

/// `tagg_node[]` under construction + the queries of the `filter_agg` nodes (one docset per segment and filter).
#[derive(Default)]
pub struct PlanBuilder<'q> {
    pub nodes: Vec<sys::tagg_node>,
    pub filters: Vec<&'q dyn Query>,
}

impl<'q> PlanBuilder<'q> {
    pub fn emit(&mut self, node: sys::tagg_node) -> u32 {
        self.nodes.push(node);
        (self.nodes.len() - 1) as u32
    }
    pub fn leaf(op: u8, kind: u8, multi: bool, field: Field) -> sys::tagg_node {
        sys::tagg_node { op, kind, multi: multi as u8, field_id: field.0, ..Default::default() }
    }
}

pub(crate) fn check(status: c_int) -> Result<()> {
    if status == sys::TAGG_OK {
        return Ok(());
    }
    let msg = unsafe { std::ffi::CStr::from_ptr(sys::tagg_last_error()) }.to_string_lossy().into_owned();
    // TAGG_ERR_NO_SUCH_COLUMN is raised exactly where the reference raises FastFieldNotAvailableError
    // (sum.rs:50-55, minmax.rs:50-55, terms.rs:76-81, percentile.rs:49-54; histogram.rs:81 panics instead)
    Err(match status {
        sys::TAGG_ERR_NO_SUCH_COLUMN => TantivyError::SchemaError(msg),
        sys::TAGG_ERR_BAD_ARG | sys::TAGG_ERR_BAD_PLAN => TantivyError::InvalidArgument(msg),
        _ => TantivyError::SystemError(msg),
    })
}

/// Owning, typed view over a `tagg_result` (freed on drop).  Bucket arrays are read through the zero-copy views.
pub struct ResultReader {
    pub(crate) raw: *mut sys::tagg_result,
}
unsafe impl Send for ResultReader {}

impl ResultReader {
    pub fn scope(&self, scope_node: u32) -> Result<(&[u64], &[u32])> {
        let (mut k, mut p, mut n) = (std::ptr::null(), std::ptr::null(), 0u64);
        check(unsafe { sys::tagg_result_scope_view(self.raw, scope_node, &mut k, &mut p, &mut n) })?;
        if n == 0 {
            return Ok((&[], &[]));
        }
        Ok(unsafe { (std::slice::from_raw_parts(k, n as usize), std::slice::from_raw_parts(p, n as usize)) })
    }
    pub fn metric(&self, node: u32) -> Result<(&[u64], &[u8])> {
        let (mut v, mut s, mut n) = (std::ptr::null(), std::ptr::null(), 0u64);
        check(unsafe { sys::tagg_result_metric_view(self.raw, node, &mut v, &mut s, &mut n) })?;
        if n == 0 {
            return Ok((&[], &[]));
        }
        Ok(unsafe { (std::slice::from_raw_parts(v, n as usize), std::slice::from_raw_parts(s, n as usize)) })
    }
    /// `Terms::top_k` on the device (terms.rs:425-457): bucket indices, sort value descending, key ascending.
    pub fn top_k(&self, scope_node: u32, parent_bucket: u64, by_node: u32, k: usize) -> Result<Vec<u32>> {
        let mut out = vec![0u32; k.max(1)];
        let mut n = 0u64;
        check(unsafe { sys::tagg_result_top_k(self.raw, scope_node, parent_bucket, by_node, k as u64, out.as_mut_ptr(), &mut n) })?;
        out.truncate(n as usize);
        Ok(out)
    }
}

impl Drop for ResultReader {
    fn drop(&mut self) {
        unsafe { sys::tagg_result_free(self.raw) };
    }
}

// tuples of arity 2..=10 (tuple.rs:73-81): TUPLE node + members, fruit = tuple of member fruits
macro_rules! impl_agg_for_tuple {
    ($n:expr; $($idx:tt $t:ident),+) => {
        impl<$($t: Agg),+> Agg for ($($t,)+) {
            type Fruit = ($($t::Fruit,)+);
            fn requires_scoring(&self) -> bool { false $(|| self.$idx.requires_scoring())+ }
            fn emit_plan<'q>(&'q self, plan: &mut PlanBuilder<'q>) -> u32 {
                let me = plan.emit(sys::tagg_node { op: sys::TAGG_OP_TUPLE, n_children: $n, ..Default::default() });
                $( self.$idx.emit_plan(plan); )+
                me
            }
            fn read_fruit(&self, res: &ResultReader, node: u32, bucket: u32) -> Result<Self::Fruit> {
                let mut at = node + 1;
                Ok(($({ let f = self.$idx.read_fruit(res, at, bucket)?; at += self.$idx.n_nodes(); let _ = at; f },)+))
            }
            fn n_nodes(&self) -> u32 { 1 $(+ self.$idx.n_nodes())+ }
        }
    };
}
impl_agg_for_tuple!(2; 0 A1, 1 A2);
impl_agg_for_tuple!(3; 0 A1, 1 A2, 2 A3);
impl_agg_for_tuple!(4; 0 A1, 1 A2, 2 A3, 3 A4);
impl_agg_for_tuple!(5; 0 A1, 1 A2, 2 A3, 3 A4, 4 A5);
impl_agg_for_tuple!(6; 0 A1, 1 A2, 2 A3, 3 A4, 4 A5, 5 A6);
impl_agg_for_tuple!(7; 0 A1, 1 A2, 2 A3, 3 A4, 4 A5, 5 A6, 6 A7);
impl_agg_for_tuple!(8; 0 A1, 1 A2, 2 A3, 3 A4, 4 A5, 5 A6, 6 A7, 7 A8);
impl_agg_for_tuple!(9; 0 A1, 1 A2, 2 A3, 3 A4, 4 A5, 5 A6, 6 A7, 7 A8, 8 A9);
impl_agg_for_tuple!(10; 0 A1, 1 A2, 2 A3, 3 A4, 4 A5, 5 A6, 6 A7, 7 A8, 8 A9, 9 A10);
