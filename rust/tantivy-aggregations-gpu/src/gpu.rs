//! Device-resident mirror of a tantivy index for the hot path: fast-field columns uploaded once per segment
//! (north_star (1)), deletes with them, and the hand-over of a scorer's matched-doc set as a bitset (north_star (2)).
use std::collections::HashMap;
use std::os::raw::c_void;

use crate::agg::check;
use tagg_sys as sys;
use tantivy::query::Scorer;
use tantivy::schema::{FieldType, Schema};
use tantivy::{DocId, Result, SegmentId, SegmentReader};

pub struct GpuSegment { pub(crate) raw: *mut sys::tagg_segment, pub max_doc: DocId }
unsafe impl Send for GpuSegment {}
unsafe impl Sync for GpuSegment {}
impl Drop for GpuSegment { fn drop(&mut self) { unsafe { sys::tagg_segment_destroy(self.raw) }; } }

/// One `tagg_ctx` (one GPU) + the resident segments, keyed by tantivy's SegmentId (a reloaded reader keeps them).
pub struct GpuIndex { pub(crate) ctx: *mut sys::tagg_ctx, segments: HashMap<SegmentId, GpuSegment> }
unsafe impl Send for GpuIndex {}
unsafe impl Sync for GpuIndex {}

impl GpuIndex {
    pub fn new(device: i32) -> Result<Self> {
        let mut ctx = std::ptr::null_mut();
        check(unsafe { sys::tagg_ctx_create(device, &mut ctx) })?;
        Ok(GpuIndex { ctx, segments: HashMap::new() })
    }

    /// Upload every FAST field of `reader` (decoded through tantivy's public readers, re-packed on the device into
    /// tantivy's own layout: `tagg_column_upload_codes`) and its delete bitset.  Done once per segment.
    pub fn ensure_segment(&mut self, reader: &SegmentReader, schema: &Schema) -> Result<&GpuSegment> {
        let id = reader.segment_id();
        if !self.segments.contains_key(&id) {
            let mut raw = std::ptr::null_mut();
            check(unsafe { sys::tagg_segment_create(self.ctx, reader.max_doc(), &mut raw) })?;
            let seg = GpuSegment { raw, max_doc: reader.max_doc() };
            for (field, entry) in schema.fields() {
                if !entry.is_int_fast() { continue; }
                let (kind, codes): (u8, Vec<u64>) = match entry.field_type() {
                    // codes are the fast field's own u64 representation (tantivy common::{i64_to_u64, f64_to_u64})
                    FieldType::U64(_) => (sys::TAGG_U64, { let r = reader.fast_fields().u64(field).unwrap(); (0..reader.max_doc()).map(|d| r.get(d)).collect() }),
                    FieldType::I64(_) => (sys::TAGG_I64, { let r = reader.fast_fields().i64(field).unwrap(); (0..reader.max_doc()).map(|d| tantivy::i64_to_u64(r.get(d))).collect() }),
                    FieldType::F64(_) => (sys::TAGG_F64, { let r = reader.fast_fields().f64(field).unwrap(); (0..reader.max_doc()).map(|d| tantivy::f64_to_u64(r.get(d))).collect() }),
                    FieldType::Date(_) => (sys::TAGG_DATE, { let r = reader.fast_fields().date(field).unwrap(); (0..reader.max_doc()).map(|d| tantivy::i64_to_u64(r.get(d).timestamp())).collect() }),
                    _ => continue,
                };
                check(unsafe { sys::tagg_column_upload_codes(seg.raw, field.0, kind as i32, codes.as_ptr(), codes.len()) })?;
                // (multi-valued fields: MultiValueIntFastFieldReader::get_vals per doc -> tagg_multicolumn_upload_codes)
            }
            if let Some(deletes) = reader.delete_bitset() {
                let mut bytes = vec![0u8; (reader.max_doc() as usize + 7) / 8];
                for d in 0..reader.max_doc() {
                    if deletes.is_deleted(d) { bytes[(d >> 3) as usize] |= 1 << (d & 7); }
                }
                check(unsafe { sys::tagg_segment_set_deletes(seg.raw, bytes.as_ptr(), bytes.len()) })?;
            }
            self.segments.insert(id, seg);
        }
        Ok(&self.segments[&id])
    }
}

impl Drop for GpuIndex {
    fn drop(&mut self) {
        self.segments.clear();
        unsafe { sys::tagg_ctx_destroy(self.ctx) };
    }
}

/// What a scorer yields, as the docset the kernels test: `Scorer::for_each` drained into a bitset
/// (`AllScorer` is recognised by the caller and becomes TAGG_DOCSET_ALL: nothing to hand over).
pub struct DrainedDocset { pub bits: Vec<u8> }
pub fn drain(mut scorer: Box<dyn Scorer>, max_doc: DocId) -> DrainedDocset {
    // 16-byte multiple: a page-locked (cudaHostRegister'ed) arena of such buffers is read in place by the GPU
    let mut bits = vec![0u8; ((max_doc as usize + 7) / 8 + 15) & !15];
    scorer.for_each(&mut |doc, _score| bits[(doc >> 3) as usize] |= 1 << (doc & 7));
    DrainedDocset { bits }
}
impl DrainedDocset {
    pub fn as_docset(&self) -> sys::tagg_docset {
        sys::tagg_docset { kind: sys::TAGG_DOCSET_BITSET, field_id: 0, data: self.bits.as_ptr() as *const c_void, n: self.bits.len() as u64, lo: 0, hi: 0 }
    }
}
pub fn all_docset() -> sys::tagg_docset {
    sys::tagg_docset { kind: sys::TAGG_DOCSET_ALL, field_id: 0, data: std::ptr::null(), n: 0, lo: 0, hi: 0 }
}
