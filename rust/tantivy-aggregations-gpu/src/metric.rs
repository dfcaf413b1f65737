//! Leaf metrics: `count_agg` (count.rs:7-9), `sum_agg_*` (sum.rs:146-158), `min/max_agg_*` (minmax.rs:152-181).
//! Fruits are the reference's: `u64`, `Option<T>`.  (The remaining kinds follow the same three-line pattern.)
use crate::agg::{Agg, PlanBuilder, ResultReader};
use tagg_sys as sys;
use tantivy::schema::Field;
use tantivy::Result;

pub struct CountAgg;
pub fn count_agg() -> CountAgg { CountAgg }
impl Agg for CountAgg {
    type Fruit = u64;
    fn emit_plan<'q>(&'q self, plan: &mut PlanBuilder<'q>) -> u32 {
        plan.emit(sys::tagg_node { op: sys::TAGG_OP_COUNT, ..Default::default() })
    }
    fn read_fruit(&self, res: &ResultReader, node: u32, bucket: u32) -> Result<u64> {
        Ok(res.metric(node)?.0[bucket as usize])
    }
    fn n_nodes(&self) -> u32 { 1 }
}

/// sum / min / max over one fast field: `Fruit = Option<T>`, `None` iff nothing was collected (sum.rs:97-101).
pub struct FoldAgg<T> { op: u8, kind: u8, multi: bool, field: Field, decode: fn(u64) -> T }
impl<T: Send> Agg for FoldAgg<T> {
    type Fruit = Option<T>;
    fn emit_plan<'q>(&'q self, plan: &mut PlanBuilder<'q>) -> u32 {
        plan.emit(PlanBuilder::leaf(self.op, self.kind, self.multi, self.field))
    }
    fn read_fruit(&self, res: &ResultReader, node: u32, bucket: u32) -> Result<Option<T>> {
        let (values, seen) = res.metric(node)?;
        Ok(if seen[bucket as usize] != 0 { Some((self.decode)(values[bucket as usize])) } else { None })
    }
    fn n_nodes(&self) -> u32 { 1 }
}
pub fn sum_agg_u64(field: Field) -> FoldAgg<u64> { FoldAgg { op: sys::TAGG_OP_SUM, kind: sys::TAGG_U64, multi: false, field, decode: |b| b } }
pub fn sum_agg_f64(field: Field) -> FoldAgg<f64> { FoldAgg { op: sys::TAGG_OP_SUM, kind: sys::TAGG_F64, multi: false, field, decode: f64::from_bits } }
pub fn min_agg_f64(field: Field) -> FoldAgg<f64> { FoldAgg { op: sys::TAGG_OP_MIN, kind: sys::TAGG_F64, multi: false, field, decode: f64::from_bits } }
pub fn max_agg_f64(field: Field) -> FoldAgg<f64> { FoldAgg { op: sys::TAGG_OP_MAX, kind: sys::TAGG_F64, multi: false, field, decode: f64::from_bits } }
