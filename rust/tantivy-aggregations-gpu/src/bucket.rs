//! Bucket aggregations and doc-stream narrowing: `terms_agg_u64` (terms.rs:185-195), `histogram_agg_f64`
//! (histogram.rs:9-21), `filter_agg` (filter.rs:8-16).  `Terms` / `Histogram` keep their public accessors
//! (`get`, `top_k`, `buckets`); only their private fields differ from the reference (terms.rs:403-458, histogram.rs:156-181).
use std::collections::{BTreeMap, HashMap};

use crate::agg::{Agg, PlanBuilder, ResultReader};
use tagg_sys as sys;
use tantivy::query::Query;
use tantivy::schema::Field;
use tantivy::Result;

#[derive(Default, Debug)]
pub struct Terms<K: std::hash::Hash + Eq, T> { res: HashMap<K, T> }
impl<K: std::hash::Hash + Eq + Ord, T> Terms<K, T> {
    pub fn get(&self, key: &K) -> Option<&T> { self.res.get(key) }
    /// terms.rs:425-457 on the decoded map (a lazily read plan selects on the device instead: `ResultReader::top_k`)
    pub fn top_k<'a, F, U>(&'a self, k: usize, mut sort_by: F) -> Vec<(&'a K, &'a T)>
    where F: FnMut(&'a T) -> U, U: Copy + Ord {
        let mut all: Vec<(U, &K, &T)> = self.res.iter().map(|(key, t)| (sort_by(t), key, t)).collect();
        all.sort_by(|a, b| b.0.cmp(&a.0).then(a.1.cmp(b.1)));
        all.into_iter().take(k).map(|(_, key, t)| (key, t)).collect()
    }
}

pub struct TermsAggU64<A> { field: Field, multi: bool, sub: A }
pub fn terms_agg_u64<A: Agg>(field: Field, sub: A) -> TermsAggU64<A> { TermsAggU64 { field, multi: false, sub } }
impl<A: Agg> Agg for TermsAggU64<A> {
    type Fruit = Terms<u64, A::Fruit>;
    fn requires_scoring(&self) -> bool { self.sub.requires_scoring() }
    fn emit_plan<'q>(&'q self, plan: &mut PlanBuilder<'q>) -> u32 {
        let me = plan.emit(sys::tagg_node { op: sys::TAGG_OP_TERMS, kind: sys::TAGG_U64, multi: self.multi as u8, field_id: self.field.0,
                                            n_children: 1, ..Default::default() });
        self.sub.emit_plan(plan);
        me
    }
    fn read_fruit(&self, res: &ResultReader, node: u32, bucket: u32) -> Result<Self::Fruit> {
        let (keys, parents) = res.scope(node)?;
        let mut out = HashMap::new();
        for (i, (&key, &parent)) in keys.iter().zip(parents).enumerate() {
            if parent == bucket {
                out.insert(key, self.sub.read_fruit(res, node + 1, i as u32)?);
            }
        }
        Ok(Terms { res: out })
    }
    fn n_nodes(&self) -> u32 { 1 + self.sub.n_nodes() }
}

#[derive(Debug)]
pub struct Histogram<T> { start: f64, interval: f64, buckets: BTreeMap<u64, T> }
impl<T> Histogram<T> {
    /// histogram.rs:163-181: ascending (bucket key, Some(fruit)), gap buckets as None
    pub fn buckets(&self) -> Vec<(f64, Option<&T>)> {
        let mut res = Vec::new();
        let mut last: Option<u64> = None;
        for (&ord, agg) in self.buckets.iter() {
            if let Some(l) = last {
                for o in (l + 1)..ord {
                    res.push((o as f64 * self.interval + self.start, None));
                }
            }
            res.push((ord as f64 * self.interval + self.start, Some(agg)));
            last = Some(ord);
        }
        res
    }
}
pub struct HistogramAggF64<A> { field: Field, start: f64, interval: f64, kind: u8, sub: A }
pub fn histogram_agg_f64<A: Agg>(field: Field, start: f64, interval: f64, sub: A) -> HistogramAggF64<A> {
    HistogramAggF64 { field, start, interval, kind: sys::TAGG_F64, sub }
}
/// README.md:41 "date_histogram" (beyond the reference): fixed-width buckets over a date fast field (seconds since the
/// epoch); the same node with a date key — ordinal = floor((t - start) / interval), documents before `start` skipped.
pub fn date_histogram_agg<A: Agg>(field: Field, interval_seconds: u64, start: i64, sub: A) -> HistogramAggF64<A> {
    HistogramAggF64 { field, start: start as f64, interval: interval_seconds as f64, kind: sys::TAGG_DATE, sub }
}
impl<A: Agg> Agg for HistogramAggF64<A> {
    type Fruit = Histogram<A::Fruit>;
    fn emit_plan<'q>(&'q self, plan: &mut PlanBuilder<'q>) -> u32 {
        let me = plan.emit(sys::tagg_node { op: sys::TAGG_OP_HISTOGRAM, kind: self.kind, field_id: self.field.0, n_children: 1,
                                            f0: self.start, f1: self.interval, ..Default::default() });
        self.sub.emit_plan(plan);
        me
    }
    fn read_fruit(&self, res: &ResultReader, node: u32, bucket: u32) -> Result<Self::Fruit> {
        let (ords, parents) = res.scope(node)?;
        let mut buckets = BTreeMap::new();
        for (i, (&ord, &parent)) in ords.iter().zip(parents).enumerate() {
            if parent == bucket {
                buckets.insert(ord, self.sub.read_fruit(res, node + 1, i as u32)?);
            }
        }
        Ok(Histogram { start: self.start, interval: self.interval, buckets })
    }
    fn n_nodes(&self) -> u32 { 1 + self.sub.n_nodes() }
}

/// README.md:36 "cardinality" (beyond the reference): the exact number of distinct values of a u64 fast field — the
/// bucket table of `terms_agg_u64(field, count_agg())` IS the distinct set; only its length is read back.
pub struct CardinalityAggU64 { inner: TermsAggU64<crate::metric::CountAgg> }
pub fn cardinality_agg_u64(field: Field) -> CardinalityAggU64 { CardinalityAggU64 { inner: terms_agg_u64(field, crate::metric::count_agg()) } }
impl Agg for CardinalityAggU64 {
    type Fruit = u64;
    fn emit_plan<'q>(&'q self, plan: &mut PlanBuilder<'q>) -> u32 { self.inner.emit_plan(plan) }
    fn read_fruit(&self, res: &ResultReader, node: u32, bucket: u32) -> Result<Self::Fruit> {
        let (_, parents) = res.scope(node)?;
        Ok(parents.iter().filter(|&&p| p == bucket).count() as u64)
    }
    fn n_nodes(&self) -> u32 { self.inner.n_nodes() }
}

/// filter_agg(&query, sub) (filter.rs:8-16): the second query's matched docs become one more docset per segment.
pub struct FilterAgg<'q, A> { query: &'q dyn Query, sub: A }
pub fn filter_agg<'q, A: Agg>(query: &'q dyn Query, sub: A) -> FilterAgg<'q, A> { FilterAgg { query, sub } }
impl<'f, A: Agg> Agg for FilterAgg<'f, A> {
    type Fruit = A::Fruit;
    fn emit_plan<'q>(&'q self, plan: &mut PlanBuilder<'q>) -> u32 {
        let aux = plan.filters.len() as u32;
        plan.filters.push(self.query);
        let me = plan.emit(sys::tagg_node { op: sys::TAGG_OP_FILTER, n_children: 1, aux, ..Default::default() });
        self.sub.emit_plan(plan);
        me
    }
    fn read_fruit(&self, res: &ResultReader, node: u32, bucket: u32) -> Result<Self::Fruit> { self.sub.read_fruit(res, node + 1, bucket) }
    fn n_nodes(&self) -> u32 { 1 + self.sub.n_nodes() }
}
